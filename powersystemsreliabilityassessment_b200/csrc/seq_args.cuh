// seq_args.cuh -- kernel arguments shared by the sequential-MC kernels (seq_mc.cu, seq_fast.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

struct SeqArgs {
    int U, H, Wd, seg_words, nseg, ypc, init_mode, K, group, persist, load16, disc, pend_cap, ev_cap, two_halves, pack_shift, static_blocks;
    const int32_t *cap; const float *mttf; const float *mttr; const uint32_t *for_thr;
    const int32_t *load; const int32_t *lmax;
    const int32_t *order;   // units sorted by decreasing transition rate (seq_wide.cu work queue)
    const uint4 *wide_tab;  // per queue position: {capacity, bits of mttf * 2^24, bits of mttr * 2^24, FOR threshold}
    uint32_t k0, k1;
    uint32_t rk[20];        // Philox round keys {k0 + r * 0x9E3779B9, k1 + r * 0xBB67AE85}, r = 0..9 (constant-bank operands)
    long long chain_base;   // absolute index of local chain 0 (Philox counter)
    long long nchains;
    const double *dur;      // injected durations or nullptr
    uint32_t *lol; long long *ens; uint32_t *ent; uint32_t *fail;
    unsigned long long *group_lol; unsigned long long *acc;
    unsigned long long *imp;   // [U] hours with loss of load in which the unit is DOWN (seq_mc.cu only), or nullptr
};

// duration of one sampler draw in ticks of 2^-24 h (DESIGN.md "Sampler"): mean_ticks = mean * 2^24
// in binary32, D = max(1, RN_int64(mean_ticks * E)).  All event-time arithmetic on ticks is exact.
#define PSRA_TICK_SHIFT 24

// seq_fast.cu
size_t seq_fast_smem_bytes(int Wd, int seg_words, int warps_per_block, int ev_cap, bool two_halves, bool load16, bool pack);
int seq_fast_max_threads(bool two_halves);
cudaError_t seq_fast_prepare(bool disc, bool two, bool pack, size_t smem, int threads, int *blocks_per_sm);
void seq_fast_launch(const SeqArgs &a, unsigned grid, int threads, size_t smem, cudaStream_t stream);

// seq_team.cu
size_t seq_team_smem_bytes(int U, int Wd, int seg_words, bool two_halves);
int seq_team_pend_cap(int U);
int seq_team_max_units();
cudaError_t seq_team_prepare(size_t smem, int *blocks_per_sm);
void seq_team_launch(const SeqArgs &a, unsigned grid, size_t smem, cudaStream_t stream);
#define SEQ_TEAM_WARPS 8

// seq_wide.cu
size_t seq_wide_smem_bytes(int Wd);
int seq_wide_threads();
cudaError_t seq_wide_prepare(size_t smem, int *blocks_per_sm);
void seq_wide_launch(const SeqArgs &a, unsigned grid, int threads, size_t smem, cudaStream_t stream);
