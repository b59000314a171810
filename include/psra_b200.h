/*
 * psra_b200.h -- C ABI of libpsra_b200.so, the B200 (sm_100a) implementation of the
 * HL1 generating-adequacy hot path of Matrixeigs/PowerSystemsReliabilityAssessment.
 *
 * The reference has no FFI seam: its boundary is the exported Julia API of
 * GeneratingAdequacy/PowerSystemAdequacy.jl:8-10 (Generator, LoadModel, ReliabilityResult,
 * run_analytical, run_non_sequential_mc, run_sequential_mc) plus the script-level
 * functions named below.  Each entry point here cites the reference function whose body
 * it replaces (paths relative to /root/reference/GeneratingAdequacy).  The Julia shim that
 * binds them with ccall is julia/PowerSystemAdequacyB200.jl; INTEGRATION.md shows the
 * binding.  The Python mirror (ctypes) is powersystemsreliabilityassessment_b200/api.py.
 *
 * Conventions
 *  - plain pointers and sizes only; every buffer is caller-owned HOST memory, borrowed for
 *    the duration of the call (Julia: pass Vectors under GC.@preserve);
 *  - every function returns 0 on success, <0 on failure (PSRA_E_*); the message is
 *    available from psra_last_error(); no exceptions, no exit(), no callbacks;
 *  - calls block until the result is on the host; one handle = one CUDA device + stream,
 *    not thread-safe per handle, independent handles are;
 *  - there is NO CPU fallback: without a CUDA device psra_create fails with PSRA_E_CUDA;
 *  - MW quantities cross the boundary as int32 fixed point (value * fp_scale chosen by the
 *    host; loads and capacities in the same scale), energies as int64 in the same unit
 *    (MW*scale * h).  Loss of load is the strict test cap < load (PSA.jl:192,253).
 */
#ifndef PSRA_B200_H
#define PSRA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PSRA_OK          0
#define PSRA_E_INVALID  -1   /* bad argument / call order */
#define PSRA_E_CUDA     -2   /* CUDA runtime error (no device, OOM, launch failure) */
#define PSRA_E_OVERFLOW -3   /* injected durations exhausted / table too small */
#define PSRA_E_NCCL     -4   /* multi-GPU handle: libnccl.so.2 missing, communicator set-up or all-reduce failed */

#define PSRA_INIT_ALL_UP      0  /* PSA.jl:223-224: every unit UP at hour 0 */
#define PSRA_INIT_STATIONARY  1  /* state ~ Bernoulli(FOR), residual ~ Exp (memoryless) */
/* OR-ed into init_mode: MATLAB next-event discretisation of Montecarlo_seq/seq_mcsampling.m:48-67
 * (time to failure rounded to the nearest hour, time to repair rounded up, DOWN from the hour after
 * the failure time); seqMain.m restarts all-up every year: use PSRA_INIT_ALL_UP, years_per_chain = 1 */
#define PSRA_DISC_MATLAB      0x100

typedef struct psra_handle psra_handle;

typedef struct psra_config {
    int32_t device;           /* CUDA device ordinal (the first one when ngpus > 1) */
    int32_t warps_per_block;  /* sequential kernel: 0 = default */
    int32_t seg_hours;        /* sequential kernel: hours per shared-memory timeline segment,
                                 multiple of 32; 0 = default */
    int32_t blocks_per_sm;    /* 0 = as many as fit */
    int32_t reserved[4];      /* reserved[0] != 0: force the generic sequential kernel (seq_mc.cu) also
                                 for systems of <= 32 units (cross-checks; the default picks seq_fast.cu);
                                 reserved[1] == 1: seq_fast.cu keeps the word sums unpacked (cross-checks);
                                 reserved[2] != 0: systems of > 32 units use seq_team.cu also where seq_wide.cu
                                 applies (cross-checks);
                                 reserved[3] > 0: number of statically scheduled Philox blocks per unit in
                                 seq_fast.cu's single-segment mode (0 = chosen from the expected demand); for
                                 seq_wide.cu: k moves the static-phase threshold to (k - 16) / 8 standard deviations
                                 below the expected demand (0 = 0.5) */
    int32_t ngpus;            /* 0 / 1: one device.  G > 1: the handle spans the devices device .. device + G - 1 of this
                                 process; psra_seq_mc / psra_nonseq_mc split their year / sample range over them
                                 (contiguous shards keyed on the global index, so the integers do not depend on G) and
                                 combine the accumulators, per-hour failure counts and the ENS histogram with
                                 ncclAllReduce(ncclSum) over NVLink (PSRA_E_NCCL if NCCL is unavailable or fails) */
    int32_t ev_cap;           /* cross-checks: event-list slots per warp of seq_fast.cu (0 = sized from the expected
                                 transitions; a small value exercises the redo path) */
    int32_t tail_bins;        /* bins (1 fixed-point MWh wide) of the per-year ENS histogram kept for psra_tail;
                                 0 = 64 x the installed capacity, between 2^16 and 2^24 */
    int32_t reserved2[5];     /* reserved2[0] != 0: the launches of a chunked run (long runs with a convergence history) all go to
                                 one stream instead of alternating between two (comparison builds) */
} psra_config;

/* lifetime ------------------------------------------------------------------------- */
int  psra_create(psra_handle **out, const psra_config *cfg);
void psra_destroy(psra_handle *h);
const char *psra_last_error(const psra_handle *h);
/* (major*1000 + minor) of the library; also proves the .so is the CUDA build */
int  psra_version(void);
/* cudaStream_t of the handle as an integer, for external event timing */
uint64_t psra_stream(const psra_handle *h);
/* SM count and SM clock (kHz) of the handle's device */
int  psra_device_info(const psra_handle *h, int32_t *sm_count, int32_t *sm_clock_khz);

/* diagnostics: raw accumulator slots of the last Monte Carlo call (slot 9 = generation waves,
 * 10 = Philox block jobs, 11 = of which pre-generated ahead, 12 = hour-resolved timeline runs) */
int  psra_last_counters(const psra_handle *h, uint64_t *out, int32_t n);

/* system data: replaces struct Generator / LoadModel, PSA.jl:20-45 ------------------ */
/* cap_fp[U] fixed-point MW; mttf_h / mttr_h in hours (lambda = 1/MTTF, mu = 1/MTTR,
 * FOR = lambda/(lambda+mu), PSA.jl:32-37) */
int psra_set_system(psra_handle *h, const int32_t *cap_fp, const double *mttf_h,
                    const double *mttr_h, int32_t n_units);
/* load_fp[H] fixed-point MW, any H >= 1 (8736 for RTS-79, 8760 in run_full_comparison.jl:19,
 * 1 for the peak-load mode of Montecarlo_nsq_single/nsqMain.m:290-296) */
int psra_set_load(psra_handle *h, const int32_t *load_fp, int32_t n_hours);

/* sequential chronological MC: replaces run_sequential_mc, PSA.jl:214-269 ----------- */
typedef struct psra_seq_summary {
    int64_t  years;             /* simulated years in this call */
    int64_t  sum_lol_hours;     /* sum over years of LOL hours (DLC, seqMain.m:166) */
    int64_t  sum_ens_fp;        /* sum of energy not supplied, fixed-point MWh */
    int64_t  sum_entries;       /* sum of deficit entries (NLC, calnlc.m:22-34) */
    int64_t  years_with_loss;
    uint64_t sum_lol_sq;        /* sum of (LOL hours)^2 */
    uint64_t sum_ens_sq_lo;     /* sum of ENS^2, 128-bit little-endian pair */
    uint64_t sum_ens_sq_hi;
    uint64_t events;            /* state transitions simulated (diagnostic) */
    float    kernel_ms;         /* CUDA-event time of the kernel(s) of this call (multi-GPU: the slowest device) */
    int32_t  redone;            /* chains a fast kernel handed back and the library replayed with the generic kernel
                                   (event list of a warp full); results are exact either way */
} psra_seq_summary;

typedef struct psra_seq_outputs {       /* all optional (NULL = not wanted) */
    uint32_t *lol_hours;    /* [nyears] */
    int64_t  *ens_fp;       /* [nyears] */
    uint32_t *entries;      /* [nyears] */
    uint32_t *fail_count;   /* [H] number of years in which hour h had loss of load
                               (tail_risk.jl:81,88 hourly_failure_prob * n_years) */
    int64_t  *group_lol;    /* [ceil(nyears/group)] LOL-hour sums of consecutive year groups:
                               cumsum/(group*k) = convergence_history, PSA.jl:263-265 */
    int32_t   group;        /* 10 for PSA.jl:263 */
    int32_t   keep_on_device; /* !=0: keep the per-year ENS / LOL vectors on the device for psra_tail */
    double   *history;      /* [nyears/group] running mean of LOL hours after every `group` years, i.e.
                               convergence_history of PSA.jl:263-265, computed on the device */
    int32_t   tail_hist;    /* !=0: the kernel also counts the years by their ENS (bins of 1 fixed-point MWh, see
                               psra_config.tail_bins) -- the histogram of tail_risk.jl:168-175 / seqMain.m:287; it stays on
                               the device (all-reduced over the devices of a multi-GPU handle) and psra_tail(values =
                               NULL) computes exact VaR / CVaR from the counts: no per-year vector in HBM */
    int32_t   reserved;
} psra_seq_outputs;

/* Years [year0, year0+nyears) of the experiment `seed`.  Years are grouped in chains of
 * years_per_chain consecutive years (year0 and nyears must be multiples of it); a chain
 * starts from init_mode and carries unit states across its years like PSA.jl:223-266.
 * Every chain has its own Philox4x32-10 streams keyed (seed; chain, unit), so results do
 * not depend on how years are split over calls, blocks or GPUs. */
int psra_seq_mc(psra_handle *h, int64_t year0, int64_t nyears, uint64_t seed,
                int32_t init_mode, int32_t years_per_chain,
                const psra_seq_outputs *out, psra_seq_summary *summary);

/* psra_seq_mc plus the weak-point statistic of Montecarlo_seq/seqMain.m:140-150,225-231 restricted to generators
 * (HL1): down_in_loss[u] = number of simulated hours with loss of load in which unit u is DOWN, summed over the
 * years of the call; the reference's comp_importance (probability that the component is down given a system
 * failure) is down_in_loss[u] / summary->sum_lol_hours.  Runs the generic kernel (any unit count; more than 32 units
 * need years_per_chain = 1); `out` may be NULL. */
int psra_seq_unit_importance(psra_handle *h, int64_t year0, int64_t nyears, uint64_t seed,
                             int32_t init_mode, int32_t years_per_chain, uint64_t *down_in_loss,
                             const psra_seq_outputs *out, psra_seq_summary *summary);

/* Same kernel fed with injected durations instead of the sampler (bit-exact parity tests):
 * durations[(chain*U + u)*K + k], k = 0 initial TTF (PSA.jl:224), then TTR, TTF, ...
 * (PSA.jl:243,246); all units start UP; all durations must be > 0. */
int psra_seq_eval_injected(psra_handle *h, const double *durations, int64_t nchains,
                           int32_t years_per_chain, int32_t K,
                           const psra_seq_outputs *out, psra_seq_summary *summary);

/* non-sequential state sampling: replaces run_non_sequential_mc, PSA.jl:169-208 ------ */
typedef struct psra_nonseq_summary {
    int64_t  samples;
    int64_t  sum_lol_hours;     /* sum over samples of hours with cap < load */
    int64_t  sum_ens_fp;
    int64_t  samples_with_loss;
    uint64_t sum_lol_sq;
    uint64_t sum_ens_sq_lo, sum_ens_sq_hi;
    float    kernel_ms;
    int32_t  reserved;
} psra_nonseq_summary;

typedef struct psra_nonseq_outputs {    /* all optional */
    uint32_t *lol_hours;    /* [n] */
    int64_t  *ens_fp;       /* [n] */
    int32_t  *cap_avail;    /* [n] fixed-point MW available */
    uint32_t *states;       /* [n * ceil(U/32)] bit u%32 of word u/32 set = unit u UP */
    int64_t  *group_lol;    /* [ceil(n/group)], group = 100 for PSA.jl:202-204 */
    int32_t   group;
    int32_t   reserved;
    double   *history;      /* [n/group] running mean of LOL hours every `group` samples (PSA.jl:202-204) */
} psra_nonseq_outputs;

/* samples [sample0, sample0+n): unit u of sample i UP iff x >= floor(FOR_u * 2^32), x the
 * Philox word keyed (seed; i, u) -- the integer form of rand() >= FOR, PSA.jl:183 */
int psra_nonseq_mc(psra_handle *h, int64_t sample0, int64_t n, uint64_t seed,
                   const psra_nonseq_outputs *out, psra_nonseq_summary *summary);
/* injected bit-packed states, layout of psra_nonseq_outputs.states */
int psra_nonseq_eval_states(psra_handle *h, const uint32_t *states, int64_t n,
                            const psra_nonseq_outputs *out, psra_nonseq_summary *summary);
/* injected uniforms r[i*U + u] standing in for rand(): UP iff r >= FOR (PSA.jl:183) */
int psra_nonseq_eval_uniforms(psra_handle *h, const double *r, int64_t n,
                              const psra_nonseq_outputs *out, psra_nonseq_summary *summary);

/* analytical engine: replaces add_unit_convolution + run_analytical, PSA.jl:67-163 --- */
/* System COPT on the grid i*step (capacities / FOR in real MW as doubles, exactly the
 * reference arithmetic incl. the rounding split PSA.jl:100-107).  probs[max_len] receives
 * the n_states probabilities; *n_states the table length (PSA.jl:73-75 growth rule). */
int psra_copt(psra_handle *h, const double *cap_mw, const double *for_rate, int32_t n_units,
              double step, double *probs, int32_t max_len, int32_t *n_states);
/* LOLE / EUE of a COPT against an hourly load in MW (PSA.jl:123-160) */
int psra_copt_indices(psra_handle *h, const double *probs, int32_t n_states, double step,
                      double total_installed, const double *load_mw, int32_t n_hours,
                      double *lole, double *eue);
/* variant of generating_adequacy_assessment.jl:113-146 (installed = last grid state,
 * strict outage > reserve) */
int psra_copt_indices_strict(psra_handle *h, const double *probs, int32_t n_states, double step,
                             const double *ldc_mw, int32_t n_hours, double *lole, double *eue);
/* frequency & duration recursion, generating_adequacy_frequency.jl:76-129, 1 MW grid;
 * rates per year from MTBF / MTTR hours (:26-31).  cum_prob / cum_freq [max_len]. */
int psra_fd_recursion(psra_handle *h, const double *cap_mw, const double *mtbf_h,
                      const double *mttr_h, int32_t n_units, double *cum_prob, double *cum_freq,
                      int32_t max_len, int32_t *n_states);
/* two-state Markov transient, Markov_process.jl:89-110: prob_down[t], t = 1..steps */
int psra_markov2(psra_handle *h, double lambda, double mu, double dt, int32_t steps,
                 double *prob_down);
/* DTMC capacity series, Markov_process.jl:159-195, injected uniforms r[t*U + i] */
int psra_dtmc_capacity(psra_handle *h, const double *mttf_h, const double *mttr_h,
                       const double *cap_mw, int32_t n_units, const double *r, int32_t n_steps,
                       double *avail_mw);
/* constant-rate failure-time experiment, Markov_process.jl:39-60: n components, each checked every dt hours against
 * rand() < lambda*dt (:54) until it fails or t > max_time (:59).  failure_time[i] = the reference's pushed time, or -1
 * for a component that outlived max_time (the reference pushes nothing).  rand(): r[i*K + k] when r != NULL, else
 * draw k of the Philox stream keyed (seed; component). */
int psra_failure_times(psra_handle *h, double lambda, double dt, double max_time, int64_t n, uint64_t seed,
                       const double *r, int32_t K, double *failure_time);

/* Sampler diagnostic: what the sequential kernels make of a 32-bit draw.  e_bits[i] = the binary32
 * bit pattern of E(draws[i]) = -ln((draw | 1) / 2^32) as the kernels compute it, ticks[i] = the duration
 * RN_int64(max(RN_f32(RN_f32(mean_h * 2^24) * E), 1)) in ticks of 2^-24 h that replaces the reference's
 * -log(rand()) / rate (PowerSystemAdequacy.jl:224,243,246).  Either output may be NULL.  Lets a test compare the
 * device logarithm draw by draw (edge draws included) with its CPU specification. */
int psra_sampler_durations(psra_handle *h, float mean_h, const uint32_t *draws, int64_t n, uint64_t *ticks, uint32_t *e_bits);

/* tail risk over per-year ENS: outputs of tail_risk.jl:18-19,79-90,168-175 plus the
 * VaR / CVaR build-side spec: VaR_a = type-7 quantile
 * (position 1+(N-1)a, linear interpolation), CVaR_a = mean of values >= VaR_a. ---------- */
typedef struct psra_tail_out {
    double  var;        /* fixed-point units */
    double  cvar;
    int64_t n_tail;     /* values >= VaR */
    int64_t x_lo, x_hi; /* the two order statistics bracketing the quantile position */
} psra_tail_out;

/* values[n]: per-year ENS (fixed point).  values == NULL uses what the last psra_seq_mc call left on the device: the
 * ENS histogram (psra_seq_outputs.tail_hist; one kernel launch over the bins, exact order statistics and tail sums
 * from the counts) or else the per-year vector (keep_on_device; radix select).  hist (optional) receives n_bins counts
 * of width bin_width starting at 0, last bin open-ended.  PSRA_E_OVERFLOW if a requested quantile lies beyond the
 * histogram's range (raise psra_config.tail_bins). */
int psra_tail(psra_handle *h, const int64_t *values, int64_t n, const double *alphas,
              int32_t n_alpha, psra_tail_out *out, int64_t *hist, int32_t n_bins,
              int64_t bin_width);

/* The device histogram as plain integers, for callers that shard the years over processes (one handle per process /
 * GPU, torch.distributed or MPI between them): export the counts of this handle, sum them element-wise over the ranks,
 * import the sums and call psra_tail(values = NULL) -- O(bins) bytes per rank instead of 8 B per year.
 * meta[4] = {years, years with loss of load, years beyond the range, their ENS sum}; counts[0 .. *n_used) = bins up to
 * the last non-empty one (n_used <= max_bins, else PSRA_E_OVERFLOW). */
int psra_tail_hist_export(psra_handle *h, int64_t *counts, int64_t max_bins, int64_t *n_used, int64_t *meta);
int psra_tail_hist_import(psra_handle *h, const int64_t *counts, int64_t n, const int64_t *meta);

/* hourly-resampled MC with maintenance / LFU / energy-limited units: replaces run_detailed_mc
 * (tail_risk.jl:12-91) == run_monte_carlo (MCvsMarkovProcess.jl:210-284) ---------------------- */
typedef struct psra_detailed_system {
    const double  *capacity_mw;        /* [n_units] */
    const double  *for_rate;           /* [n_units] mechanical FOR (tail_risk.jl:44) */
    const int32_t *maint_start_week;   /* [n_units] scheduled_outage_start, 1-based week (:39) */
    const int32_t *maint_weeks;        /* [n_units] 0 = no maintenance */
    const double  *energy_limit_mwh;   /* [n_units] +Inf = not energy limited (:46) */
    int32_t n_units;                   /* <= 30 */
    int32_t reserved;
} psra_detailed_system;

/* years [year0, year0+nyears) of experiment `seed`; year_lole[nyears] = hours with deficit per year
 * (yearly_lole_distribution), hourly_fail[H] (optional) = years in which hour h had a deficit
 * (hourly_failure_prob * n_years); lfu_std_mw = maximum(base_load) * lfu_sigma_percent / 100 */
int psra_detailed_mc(psra_handle *h, const psra_detailed_system *sys, const double *base_load_mw,
                     int32_t n_hours, double lfu_std_mw, int64_t year0, int64_t nyears, uint64_t seed,
                     uint32_t *year_lole, uint32_t *hourly_fail, float *kernel_ms);
/* same loop with injected rand() / randn(): uniforms[(y*H + h)*U + u], normals[y*H + h].  The layout is per (year, hour,
 * unit) whatever the maintenance schedule; the reference only calls rand() for a unit that is NOT on maintenance in that
 * hour (tail_risk.jl:39-44), so a stream recorded from the reference is laid out by skipping those units
 * (tests/golden/ref_detailed_mc.npz is produced that way from the reference text) */
int psra_detailed_eval_injected(psra_handle *h, const psra_detailed_system *sys, const double *base_load_mw,
                                int32_t n_hours, double lfu_std_mw, int64_t nyears, const double *uniforms,
                                const double *normals, uint32_t *year_lole, uint32_t *hourly_fail);

/* multi-area sequential adequacy with tie-line support: replaces run_fast_sequential_simulation /
 * solve_curtailment_fast (GeneratingAdequacy/AdequacyAssessmentII.jl:73-179,185-250) -------------- */
#define PSRA_MAX_AREAS 8
#define PSRA_POLICY_ISOLATED        0   /* SupportPolicy ISOLATED       (AdequacyAssessmentII.jl:63,84-92)  */
#define PSRA_POLICY_INTERCONNECTED  1   /* SupportPolicy INTERCONNECTED (AdequacyAssessmentII.jl:96-176)    */

typedef struct psra_area_system {
    int32_t n_areas;               /* <= PSRA_MAX_AREAS; the hour timelines of all areas share one SM's shared memory */
    int32_t n_units;               /* all generators of all areas (Area.generators, :35-40) */
    int32_t n_hours;               /* 8760 in the reference (:200) */
    int32_t reserved;
    const int32_t *unit_area;      /* [n_units] 0-based area of each unit */
    const int32_t *cap_fp;         /* [n_units] Generator.capacity (:17), fixed point */
    const double  *mttf_h;         /* [n_units] */
    const double  *mttr_h;         /* [n_units] */
    const int32_t *load_fp;        /* [n_areas][n_hours] Area.hourly_load (:39), fixed point */
    const int32_t *topology_fp;    /* [n_areas][n_areas] System.topology_matrix (:46,52-61): tie capacities, both directions */
} psra_area_system;

typedef struct psra_area_outputs {  /* optional per-year, per-area vectors [nyears][n_areas] */
    uint32_t *lol_hours;           /* hours with curtailment > 0 (:231-232) */
    int64_t  *ens_fp;              /* curtailed energy (:233) */
} psra_area_outputs;

typedef struct psra_area_summary {
    int64_t  years;
    int64_t  sum_lol_hours[PSRA_MAX_AREAS];   /* area_lole * n_years (:241) */
    int64_t  sum_ens_fp[PSRA_MAX_AREAS];      /* area_eue * n_years (:242) */
    uint64_t events;
    float    kernel_ms;
    int32_t  n_areas;
} psra_area_summary;

/* Years [year0, year0+nyears) of experiment `seed`, independent years (each with its own Philox streams keyed
 * (seed; year, global unit index); init_mode as psra_seq_mc).  Self-contained: the system and load the handle holds
 * from psra_set_system / psra_set_load are neither used nor changed. */
int psra_multi_area_mc(psra_handle *h, const psra_area_system *sys, int32_t policy, int64_t year0,
                       int64_t nyears, uint64_t seed, int32_t init_mode, const psra_area_outputs *out,
                       psra_area_summary *summary);

#ifdef __cplusplus
}
#endif
#endif /* PSRA_B200_H */
