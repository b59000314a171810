// seq_args.cuh -- kernel arguments shared by the sequential-MC kernels (seq_mc.cu, seq_fast.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

struct SeqArgs {
    int U, H, Wd, seg_words, nseg, ypc, init_mode, K, group, persist, load16, disc, pend_cap, ev_cap, two_halves, pack_shift, static_blocks;
    uint8_t wide_sblk[64];  // seq_wide.cu: statically scheduled Philox blocks per unit, for every group of 32 queue positions
    int redo_cap;           // capacity of the redo list
    int group_phase;        // local year 0 of this launch is year `group_phase` of its history group (replays only)
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
    // redo list: redo[0] = number of chains a fast kernel handed back (event list of seq_fast.cu full),
    // redo[1 + i] = their absolute chain indices; the host replays them with the generic kernel of seq_mc.cu
    unsigned long long *redo;
    // per-year ENS histogram (tail risk, tail_risk.jl:168-175 / seqMain.m:287): hist[e] = number of years whose ENS is
    // e fixed-point MWh, 1 <= e < hist_bins (years without loss of load are not entered: their number is
    // years - years_with_loss); hist[hist_bins] / hist[hist_bins + 1] = number / ENS sum of the years beyond the range
    unsigned long long *hist;
    long long hist_bins;
};

#ifdef __CUDACC__
__device__ __forceinline__ void seq_redo_push(const SeqArgs &a, long long chain)
{
    const unsigned long long i = atomicAdd(&a.redo[0], 1ull);
    if (i < (unsigned long long)a.redo_cap) a.redo[1 + i] = (unsigned long long)chain;
}

// one year with loss of load into the ENS histogram
__device__ __forceinline__ void seq_hist_add(const SeqArgs &a, long long ens)
{
    if (!a.hist) return;
    if ((unsigned long long)ens < (unsigned long long)a.hist_bins) atomicAdd(&a.hist[ens], 1ull);
    else {
        atomicAdd(&a.hist[a.hist_bins], 1ull);
        atomicAdd(&a.hist[a.hist_bins + 1], (unsigned long long)ens);
    }
}
#endif

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
#define SEQ_WIDE_MAX_UNITS 2048
size_t seq_wide_smem_bytes(int Wd, int U, int nwarps);
int seq_wide_max_warps();
cudaError_t seq_wide_prepare(bool disc, size_t smem, int threads, int *blocks_per_sm);
void seq_wide_launch(const SeqArgs &a, unsigned grid, int threads, size_t smem, cudaStream_t stream);
