/*
 * psra_oracle.c -- CPU restatement of the reference's HL1 generating-adequacy path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under powersystemsreliabilityassessment_b200/
 * may import, link or execute this file; only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs use it, as the checker / CPU baseline.
 *
 * The reference (Matrixeigs/PowerSystemsReliabilityAssessment) is interpreted Julia +
 * MATLAB; neither runtime exists in this image, so the reference itself cannot be
 * compiled or run here (no oracle/_ref).  Every function below restates, line by line,
 * the arithmetic of the cited reference lines in plain C (FP64, no FMA contraction,
 * same operation order).  Pinning:
 *   (1) known answers: the deterministic functions against the values derived in SURVEY.md
 *       section 8c / BASELINE.md section 3 (classic RTS-79 HL1 LOLE 9.394 h/yr, the script
 *       demos' printed values), tests/test_oracle.py;
 *   (2) the reference's own source text: run_sequential_mc / run_non_sequential_mc /
 *       add_unit_convolution / run_analytical are cut out of the reference checkout and
 *       transliterated line by line into Python (oracle/jl_transliterate.py; the three
 *       `-log(rand())/rate` draws of PSA.jl:224,243,246 read from per-unit lists, rand() of
 *       PSA.jl:183 replayed from a recorded matrix); the vectors they produce are committed
 *       (tests/golden/ref_*.npz, scripts/make_reference_golden.py) and this file reproduces
 *       them bit for bit (tests/test_reference_pin.py), as does an independent naive
 *       transcription (oracle/psa_literal.py); likewise the other Julia files of the path
 *       (multi-area, detailed MC, F&D, COPT demo, Markov script blocks) and the two MATLAB
 *       functions (Montecarlo_seq/seq_mcsampling.m, calnlc.m; oracle/m_transliterate.py,
 *       tests/golden/ref_matlab.npz);
 *   (3) not done here: Julia itself.  tools/patched_reference.jl runs the same three
 *       substitutions in a real Julia for whoever has one.
 * The Monte Carlo STREAMS of the reference (Julia's unversioned default RNG, never seeded)
 * cannot be reproduced by anyone; parity of the sampler paths is per trial on injected
 * duration / state / uniform matrices.
 *
 * All citations are path:line under /root/reference/GeneratingAdequacy unless a
 * directory is given.  PSA.jl = PowerSystemAdequacy.jl.
 *
 * Build:  gcc -O3 -ffp-contract=off -fno-fast-math -shared -fPIC (see oracle/Makefile).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------
 * 1. Sequential chronological MC -- PSA.jl:214-269, literal hour/unit loop.
 *
 * The reference draws every duration as -log(rand())/rate from one global stream
 * (PSA.jl:224,243,246).  Here each draw is replaced by the next entry of a per-unit
 * duration list dur[u*K + k] (k = 0 is the initial TTF of PSA.jl:224, then TTR, TTF,...),
 * the injection protocol of SURVEY.md section 8c.  Everything else is the literal loop:
 * ttf -= 1.0 per hour per unit (PSA.jl:239), while ttf <= 0 toggle and add the next
 * duration (PSA.jl:240-248), capacity of UP units summed in unit order (PSA.jl:249),
 * strict cap < load test and deficit accumulation (PSA.jl:253-257).  State carries
 * across years (PSA.jl:223-224 sit outside the year loop).
 *
 * Per-year outputs additionally hold the number of deficit entries per
 * Montecarlo_seq/calnlc.m:22-34 (0->1 transitions of the hourly flag plus flag(1)),
 * the MATLAB definition of NLC / LOLF (Montecarlo_seq/seqMain.m:160-169).
 *
 * init_status[u] (nullable): 1 = UP at hour 0 (the reference: all UP), 0 = DOWN.
 * Returns 0, or -1 if a unit ran out of injected durations.
 * -------------------------------------------------------------------------------- */
int oracle_seq_literal(int U, const double *cap, int H, const double *load, int years,
                       const double *dur, int K, const unsigned char *init_status,
                       double *year_lol, double *year_eue, double *year_entries,
                       int *draws_used)
{
    unsigned char *status = (unsigned char *)malloc((size_t)U);
    double *ttf = (double *)malloc(sizeof(double) * (size_t)U);
    int *next = (int *)malloc(sizeof(int) * (size_t)U);
    int rc = 0;
    for (int i = 0; i < U; i++) {
        status[i] = init_status ? init_status[i] : 1;
        ttf[i] = dur[(size_t)i * K + 0];
        next[i] = 1;
    }
    for (int y = 0; y < years && rc == 0; y++) {
        double lole = 0.0, eue = 0.0, entries = 0.0;
        int prev_flag = 0;
        for (int h = 0; h < H && rc == 0; h++) {
            double cap_avail = 0.0;
            for (int i = 0; i < U; i++) {
                ttf[i] -= 1.0;
                while (ttf[i] <= 0) {
                    if (next[i] >= K) { rc = -1; break; }
                    status[i] = !status[i];
                    ttf[i] += dur[(size_t)i * K + next[i]];
                    next[i]++;
                }
                if (rc) break;
                if (status[i]) cap_avail += cap[i];
            }
            if (rc) break;
            int flag = 0;
            if (cap_avail < load[h]) {
                double deficit = load[h] - cap_avail;
                lole += 1.0;
                eue += deficit;
                flag = 1;
            }
            if (flag && !prev_flag) entries += 1.0; /* calnlc.m:22-34; prev_flag = 0 at h = 0 */
            prev_flag = flag;
        }
        year_lol[y] = lole;
        year_eue[y] = eue;
        if (year_entries) year_entries[y] = entries;
    }
    if (draws_used) for (int i = 0; i < U; i++) draws_used[i] = next[i];
    free(status); free(ttf); free(next);
    return rc;
}

/* ------------------------------------------------------------------------------------
 * 2. Counter-based sampler specification (build-side; DESIGN.md "Sampler").
 *
 * The reference's random stream is Julia's unpinned default RNG, so the B200 library
 * defines its own: Philox4x32-10 (Salmon et al., SC'11; Random123 1.x), key =
 * (seed_lo, seed_hi), counter = (chain_lo, chain_hi, unit, block).  This is an
 * independent restatement of that published algorithm, pinned against the Random123
 * known-answer vectors in tests/test_oracle.py.
 * -------------------------------------------------------------------------------- */
void oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* E = -ln(u) for the draw x (sampler specification v2), evaluated with integer bit operations and IEEE
 * binary32 add / fma only (every step correctly rounded, so any IEEE machine reproduces it bit for bit):
 *   w = x | 1;  lz = clz32(w);  X = w << lz;  u = w / 2^32 = m * 2^-k,  k = lz + 1,  m = X / 2^31 in [1,2);
 *   m is truncated to 24 bits: bits(m) = 0x3F800000 | ((X >> 8) & 0x7FFFFF);
 *   t = m - 1.5 (exact);  R = P7(t) ~ -ln(1.5 + t)  (Horner with fma; minimax fit on [-0.5, 0.5), |error| < 2.5e-7);
 *   E = fma(k, LN2, R),  LN2 = 0x1.62e430p-1.
 * -1.2e-7 <= E <= 32 ln 2: for the 768 draws nearest 2^32 the approximation error makes E <= 0; the duration
 * clamp (at least one tick) absorbs them.  The coefficients come from a Remez exchange (float64) rounded to
 * binary32; measured over 2 10^7 random draws: max |E + ln u| = 7.2e-7 (at E > 16, rounding of the result),
 * mean error 3.4e-8. */
static float u32_as_float(uint32_t b) { float f; memcpy(&f, &b, 4); return f; }

static const float LOGP[8] = { -0x1.9f324cp-2f, -0x1.555536p-1f, 0x1.c72898p-3f, -0x1.94b470p-4f,
                               0x1.90d388p-5f, -0x1.a7b9fep-6f, 0x1.1d506cp-6f, -0x1.578b02p-7f };
#define ORACLE_LN2_F 0x1.62e430p-1f

float oracle_neglog_u32(uint32_t x)
{
    uint32_t w = x | 1u;
    int lz = __builtin_clz(w);
    uint32_t X = w << lz;
    float m = u32_as_float(0x3F800000u | ((X >> 8) & 0x007FFFFFu));
    float t = m - 1.5f;
    float p = LOGP[7];
    for (int i = 6; i >= 0; i--) p = fmaf(p, t, LOGP[i]);
    float kf = (float)(lz + 1);
    return fmaf(kf, ORACLE_LN2_F, p);
}

/* Duration of one draw in hours: quantised to ticks of 2^-24 h so that every residual of the
 * literal `ttf -= 1.0` / `ttf += D` recurrence below is an exact FP64 number (event times are then
 * plain prefix sums, which is what lets the GPU generate them in parallel):
 *   D = RN_int64(max(RN_f32(RN_f32(mean * 2^24) * E), 1)) / 2^24,  E = oracle_neglog_u32(draw). */
static double duration_hours(float mean_f32, uint32_t x)
{
    float mt = mean_f32 * 16777216.0f;
    float p = fmaxf(mt * oracle_neglog_u32(x), 1.0f);
    long long t = llrintf(p);          /* round to nearest even (default rounding mode) */
    return (double)t / 16777216.0;
}

double oracle_duration_hours(float mean_f32, uint32_t x) { return duration_hours(mean_f32, x); }

/* vector form for the per-draw parity test of the device sampler: ticks of 2^-24 h and the bits of E */
void oracle_sampler_durations(float mean_f32, const uint32_t *draws, long long n, uint64_t *ticks, uint32_t *e_bits)
{
    float mt = mean_f32 * 16777216.0f;
    for (long long i = 0; i < n; i++) {
        float e = oracle_neglog_u32(draws[i]);
        if (e_bits) memcpy(&e_bits[i], &e, 4);
        if (ticks) ticks[i] = (uint64_t)llrintf(fmaxf(mt * e, 1.0f));
    }
}

/* Per-(chain, unit) draw stream: draw j lives in Philox block j/4, word j%4. */
typedef struct { uint32_t key[2]; uint32_t ctr[4]; uint32_t buf[4]; uint32_t j; } draw_stream;

static void stream_init(draw_stream *s, uint64_t seed, uint64_t chain, uint32_t unit)
{
    s->key[0] = (uint32_t)seed; s->key[1] = (uint32_t)(seed >> 32);
    s->ctr[0] = (uint32_t)chain; s->ctr[1] = (uint32_t)(chain >> 32);
    s->ctr[2] = unit; s->ctr[3] = 0; s->j = 0;
}
static uint32_t stream_next(draw_stream *s)
{
    if ((s->j & 3u) == 0) { s->ctr[3] = s->j >> 2; oracle_philox4x32_10(s->ctr, s->key, s->buf); }
    return s->buf[(s->j++) & 3u];
}

/* Sequential MC driven by the sampler above: the same literal loop as section 1
 * (PSA.jl:230-266) where each -log(rand())/rate of PSA.jl:224,243,246 becomes

 * duration_hours(mean_f32, next draw of the unit's stream) (tick-quantised, see above).
 * Chains: years are grouped into chains of years_per_chain consecutive years; each
 * chain starts from a fresh state (PSA.jl:223-224) and carries it across its years
 * (years_per_chain = years reproduces the reference's single chain; 1 = independent
 * years).  init_mode 0 = ALL_UP (the reference; draw 0 is consumed and ignored so that
 * the duration draws keep the same indices in both modes), 1 = STATIONARY: unit DOWN
 * iff draw0 < for_thr[u] (= floor(FOR * 2^32)), first residual ~ Exp(mean of that state),
 * which is the stationary law of the alternating process by memorylessness. */
int oracle_seq_philox_ex(int U, const double *cap, const float *mttf_f, const float *mttr_f,
                         const uint32_t *for_thr, int H, const double *load, uint64_t seed,
                         int64_t chain0, int64_t nchains, int years_per_chain, int init_mode,
                         double *year_lol, double *year_eue, double *year_entries, double *unit_down_in_loss);

int oracle_seq_philox(int U, const double *cap, const float *mttf_f, const float *mttr_f,
                      const uint32_t *for_thr, int H, const double *load, uint64_t seed,
                      int64_t chain0, int64_t nchains, int years_per_chain, int init_mode,
                      double *year_lol, double *year_eue, double *year_entries)
{
    return oracle_seq_philox_ex(U, cap, mttf_f, mttr_f, for_thr, H, load, seed, chain0, nchains, years_per_chain, init_mode,
                                year_lol, year_eue, year_entries, NULL);
}

/* ... plus the weak-point statistic of Montecarlo_seq/seqMain.m:140-150,225-231 restricted to generators:
 * unit_down_in_loss[u] (optional, accumulated over all years) = number of hours with loss of load in which unit u
 * is DOWN; comp_importance = unit_down_in_loss / total loss hours. */
int oracle_seq_philox_ex(int U, const double *cap, const float *mttf_f, const float *mttr_f,
                         const uint32_t *for_thr, int H, const double *load, uint64_t seed,
                         int64_t chain0, int64_t nchains, int years_per_chain, int init_mode,
                         double *year_lol, double *year_eue, double *year_entries, double *unit_down_in_loss)
{
    draw_stream *st = (draw_stream *)malloc(sizeof(draw_stream) * (size_t)U);
    unsigned char *status = (unsigned char *)malloc((size_t)U);
    double *ttf = (double *)malloc(sizeof(double) * (size_t)U);
    for (int64_t c = 0; c < nchains; c++) {
        for (int i = 0; i < U; i++) {
            stream_init(&st[i], seed, (uint64_t)(chain0 + c), (uint32_t)i);
            uint32_t x0 = stream_next(&st[i]);
            status[i] = (init_mode == 1 && x0 < for_thr[i]) ? 0 : 1;
            float mean = status[i] ? mttf_f[i] : mttr_f[i];
            ttf[i] = duration_hours(mean, stream_next(&st[i]));
        }
        for (int y = 0; y < years_per_chain; y++) {
            double lole = 0.0, eue = 0.0, entries = 0.0;
            int prev_flag = 0;
            for (int h = 0; h < H; h++) {
                double cap_avail = 0.0;
                for (int i = 0; i < U; i++) {
                    ttf[i] -= 1.0;
                    while (ttf[i] <= 0) {
                        status[i] = !status[i];
                        float mean = status[i] ? mttf_f[i] : mttr_f[i];
                        ttf[i] += duration_hours(mean, stream_next(&st[i]));
                    }
                    if (status[i]) cap_avail += cap[i];
                }
                int flag = 0;
                if (cap_avail < load[h]) {
                    lole += 1.0;
                    eue += load[h] - cap_avail;
                    flag = 1;
                    if (unit_down_in_loss)
                        for (int i = 0; i < U; i++)
                            if (!status[i]) unit_down_in_loss[i] += 1.0;
                }
                if (flag && !prev_flag) entries += 1.0;
                prev_flag = flag;
            }
            size_t o = (size_t)(c * years_per_chain + y);
            year_lol[o] = lole; year_eue[o] = eue;
            if (year_entries) year_entries[o] = entries;
        }
    }
    free(st); free(status); free(ttf);
    return 0;
}

/* ------------------------------------------------------------------------------------
 * 3. Non-sequential state sampling -- PSA.jl:169-208.
 *
 * Per iteration: unit UP iff r >= FOR in unit order (PSA.jl:183), capacity summed in
 * unit order (PSA.jl:184), then every hour: strict cap < load, count + deficit
 * (PSA.jl:191-197).  r[i*U + u] are the injected uniforms standing in for rand().
 * Per-iteration outputs are returned so tests can compare them one by one.
 * -------------------------------------------------------------------------------- */
void oracle_nonseq_literal(int U, const double *cap, const double *for_rate, int H,
                           const double *load, int64_t iters, const double *r,
                           double *iter_lol, double *iter_eue, double *iter_cap)
{
    for (int64_t i = 0; i < iters; i++) {
        double cap_avail = 0.0;
        for (int u = 0; u < U; u++)
            if (r[(size_t)i * U + u] >= for_rate[u]) cap_avail += cap[u];
        double lole = 0.0, eue = 0.0;
        for (int h = 0; h < H; h++) {
            if (cap_avail < load[h]) {
                double deficit = load[h] - cap_avail;
                lole += 1.0;
                eue += deficit;
            }
        }
        iter_lol[i] = lole; iter_eue[i] = eue;
        if (iter_cap) iter_cap[i] = cap_avail;
    }
}

/* Same loop with packed states: bit u of word states[i*W + u/32] set = unit u UP
 * (the MATLAB twin samples DOWN iff rand < U, Montecarlo_nsq_single/mc_sampling.m:35). */
void oracle_nonseq_states(int U, const double *cap, int H, const double *load, int64_t iters,
                          const uint32_t *states, double *iter_lol, double *iter_eue)
{
    int W = (U + 31) / 32;
    for (int64_t i = 0; i < iters; i++) {
        double cap_avail = 0.0;
        for (int u = 0; u < U; u++)
            if ((states[(size_t)i * W + (u >> 5)] >> (u & 31)) & 1u) cap_avail += cap[u];
        double lole = 0.0, eue = 0.0;
        for (int h = 0; h < H; h++)
            if (cap_avail < load[h]) { lole += 1.0; eue += load[h] - cap_avail; }
        iter_lol[i] = lole; iter_eue[i] = eue;
    }
}

/* Sampler-driven non-sequential MC: sample i, unit u uses draw (u % 4) of Philox block
 * counter = (i_lo, i_hi, u / 4, 0xNS) with key = seed; unit UP iff x >= for_thr[u]
 * (PSA.jl:183 with u = x / 2^32). */
void oracle_nonseq_philox(int U, const double *cap, const uint32_t *for_thr, int H,
                          const double *load, uint64_t seed, int64_t i0, int64_t iters,
                          double *iter_lol, double *iter_eue, uint32_t *states)
{
    int W = (U + 31) / 32;
    uint32_t key[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) };
    for (int64_t i = 0; i < iters; i++) {
        uint64_t s = (uint64_t)(i0 + i);
        double cap_avail = 0.0;
        uint32_t out[4] = {0, 0, 0, 0};
        for (int w = 0; w < W; w++) if (states) states[(size_t)i * W + w] = 0;
        for (int u = 0; u < U; u++) {
            if ((u & 3) == 0) {
                uint32_t ctr[4] = { (uint32_t)s, (uint32_t)(s >> 32), (uint32_t)(u >> 2), 0x4E53u };
                oracle_philox4x32_10(ctr, key, out);
            }
            if (out[u & 3] >= for_thr[u]) {
                cap_avail += cap[u];
                if (states) states[(size_t)i * W + (u >> 5)] |= 1u << (u & 31);
            }
        }
        double lole = 0.0, eue = 0.0;
        for (int h = 0; h < H; h++)
            if (cap_avail < load[h]) { lole += 1.0; eue += load[h] - cap_avail; }
        iter_lol[i] = lole; iter_eue[i] = eue;
    }
}

/* ------------------------------------------------------------------------------------
 * 4. Analytical COPT -- PSA.jl:67-111 (add_unit_convolution) and PSA.jl:113-163.
 * -------------------------------------------------------------------------------- */
static double copt_get(const double *p, int n, double x_val, double step)
{
    /* PSA.jl:81-88: idx = Int(round(X/step)) + 1, 0 outside.  Julia's round() is
     * round-half-even = rint() in the default rounding mode. */
    double q = rint(x_val / step);
    if (q < 0.0 || q > (double)(n - 1)) return 0.0;
    return p[(int)q];
}

/* One convolution step.  old table has n_old states i*step; returns new length, writes
 * new_p (caller sized via oracle_copt_next_len). */
int oracle_copt_next_len(int n_old, double C, double step)
{
    /* PSA.jl:73-75: grid max of the OLD table, not the installed sum */
    double max_old = (n_old > 0) ? (double)(n_old - 1) * step : 0.0;
    double max_new = max_old + C;
    return (int)ceil(max_new / step) + 1;
}

void oracle_copt_add_unit(const double *old_p, int n_old, double C, double q, double step,
                          double *new_p, int n_new)
{
    double p = 1.0 - q;
    int lower_idx = (int)floor(C / step);
    double C_lower = lower_idx * step;
    double C_upper = (lower_idx + 1) * step;
    if (fabs(C - C_lower) < 1e-5) {                 /* PSA.jl:95-98 */
        for (int i = 0; i < n_new; i++) {
            double X = i * step;
            new_p[i] = copt_get(old_p, n_old, X, step) * p + copt_get(old_p, n_old, X - C, step) * q;
        }
    } else {                                        /* PSA.jl:100-107 */
        double alpha = (C - C_lower) / step;
        double q_upper = q * alpha;
        double q_lower = q * (1.0 - alpha);
        for (int i = 0; i < n_new; i++) {
            double X = i * step;
            new_p[i] = copt_get(old_p, n_old, X, step) * p
                     + copt_get(old_p, n_old, X - C_lower, step) * q_lower
                     + copt_get(old_p, n_old, X - C_upper, step) * q_upper;
        }
    }
}

/* Build the system COPT (PSA.jl:118-121).  probs must hold max_len doubles; returns n. */
int oracle_copt_build(int U, const double *cap, const double *for_rate, double step,
                      double *probs, int max_len)
{
    double *a = (double *)calloc((size_t)max_len, sizeof(double));
    double *b = (double *)calloc((size_t)max_len, sizeof(double));
    int n = 1; a[0] = 1.0;
    for (int u = 0; u < U; u++) {
        int n2 = oracle_copt_next_len(n, cap[u], step);
        if (n2 > max_len) { free(a); free(b); return -1; }
        oracle_copt_add_unit(a, n, cap[u], for_rate[u], step, b, n2);
        double *t = a; a = b; b = t; n = n2;
    }
    memcpy(probs, a, sizeof(double) * (size_t)n);
    free(a); free(b);
    return n;
}

/* PSA.jl:123-160: LOLE / EUE of a COPT against the hourly load, literal tail loops. */
void oracle_analytical_indices(const double *probs, int n, double step, double total_installed,
                               int H, const double *load, double *lole_out, double *eue_out)
{
    double *cum = (double *)malloc(sizeof(double) * (size_t)n);
    /* reverse(cumsum(reverse(p))): running sum from the last state downwards (PSA.jl:130) */
    double acc = 0.0;
    for (int i = n - 1; i >= 0; i--) { acc += probs[i]; cum[i] = acc; }
    double lole = 0.0, eue = 0.0;
    for (int h = 0; h < H; h++) {
        double reserve = total_installed - load[h];
        long idx = (long)floor(reserve / step) + 2;       /* 1-based, PSA.jl:140 */
        if (idx <= n && idx >= 1) {
            lole += cum[idx - 1];
            for (long k = idx; k <= n; k++) {
                double outage = (double)(k - 1) * step;
                eue += (outage - reserve) * probs[k - 1];
            }
        } else if (idx < 1) {                             /* PSA.jl:152-159 */
            lole += 1.0;
            eue += (load[h] - total_installed);
            double avg = 0.0;
            for (int k = 0; k < n; k++) avg += ((double)k * step) * probs[k];
            eue += avg;
        }
    }
    *lole_out = lole; *eue_out = eue;
    free(cum);
}

/* ------------------------------------------------------------------------------------
 * 5. generating_adequacy_assessment.jl:30-146 -- the stand-alone COPT variant:
 *    tolerance lookup |x - X| < 1e-5 (:48-56), C_upper = ceil (:70-75), strict
 *    outage > reserve with installed = last grid state (:125-139).
 * -------------------------------------------------------------------------------- */
static double copt_get_tol(const double *p, int n, double x_val, double step)
{
    for (int k = 0; k < n; k++)
        if (fabs((double)k * step - x_val) < 1e-5) return p[k];
    return 0.0;
}

void oracle_gaa_add_unit(const double *old_p, int n_old, double C, double q, double step,
                         double *new_p, int n_new)
{
    double p = 1.0 - q;
    int lower_idx = (int)floor(C / step);
    int upper_idx = (int)ceil(C / step);
    double C_lower = lower_idx * step, C_upper = upper_idx * step;
    if (C_lower == C_upper) {
        for (int i = 0; i < n_new; i++) {
            double X = i * step;
            double term1 = copt_get_tol(old_p, n_old, X, step) * p;
            double term2 = copt_get_tol(old_p, n_old, X - C, step) * q;
            new_p[i] = term1 + term2;
        }
    } else {
        double alpha = (C - C_lower) / step;
        double q_upper = q * alpha, q_lower = q * (1.0 - alpha);
        for (int i = 0; i < n_new; i++) {
            double X = i * step;
            double term1 = copt_get_tol(old_p, n_old, X, step) * p;
            double term2 = copt_get_tol(old_p, n_old, X - C_lower, step) * q_lower;
            double term3 = copt_get_tol(old_p, n_old, X - C_upper, step) * q_upper;
            new_p[i] = term1 + term2 + term3;
        }
    }
}

void oracle_gaa_calculate_indices(const double *probs, int n, double step, int H,
                                  const double *ldc, double *lole_out, double *eue_out)
{
    double lole = 0.0, eue = 0.0;
    double installed = (double)(n - 1) * step;            /* :125 */
    for (int h = 0; h < H; h++) {
        double reserve = installed - ldc[h];
        double pl = 0.0, es = 0.0;
        for (int i = 0; i < n; i++) {
            double outage = (double)i * step;
            if (outage > reserve) {
                pl += probs[i];
                es += (outage - reserve) * probs[i];
            }
        }
        lole += pl; eue += es;
    }
    *lole_out = lole; *eue_out = eue;
}

/* ------------------------------------------------------------------------------------
 * 6. Frequency & duration recursion -- generating_adequacy_frequency.jl:76-129.
 *    Cumulative tables on a 1 MW grid 0..max (:69-70); boundary P(x<0)=1, F=0 (:77-82);
 *    lookup = first level >= x, (0,0) beyond the table (:85-98).
 *    lam is the failure rate per YEAR (8760/MTBF, :26-27), p/q availability (:30-31).
 * -------------------------------------------------------------------------------- */
static void fd_get(const double *P, const double *F, int n, double x, double *p_out, double *f_out)
{
    if (x < 0) { *p_out = 1.0; *f_out = 0.0; return; }
    long idx = (long)ceil(x);              /* first integer level >= x on the 1 MW grid */
    if (idx >= n) { *p_out = 0.0; *f_out = 0.0; return; }
    *p_out = P[idx]; *f_out = F[idx];
}

void oracle_fd_add_unit(const double *P_old, const double *F_old, int n_old, double C, double p,
                        double q, double lam, double *P_new, double *F_new, int n_new)
{
    for (int i = 0; i < n_new; i++) {
        double X = (double)i;
        double PX, FX, PXC, FXC;
        fd_get(P_old, F_old, n_old, X, &PX, &FX);
        fd_get(P_old, F_old, n_old, X - C, &PXC, &FXC);
        P_new[i] = (p * PX) + (q * PXC);                   /* :110 */
        double term1 = p * FX, term2 = q * FXC;            /* :113-116 */
        double term3 = lam * p * (PXC - PX);
        F_new[i] = term1 + term2 + term3;
    }
}

/* evaluate_risk, generating_adequacy_frequency.jl:155-186 */
void oracle_fd_evaluate(const double *P, const double *F, int n, double peak, double installed,
                        double *lole_h, double *lolf, double *lold)
{
    double reserve = installed - peak;
    *lole_h = 0.0; *lolf = 0.0; *lold = 0.0;
    for (int i = 0; i < n; i++) {
        if ((double)i > reserve) {
            *lole_h = P[i] * 8760.0;
            *lolf = F[i];
            *lold = (*lolf > 0) ? (*lole_h / *lolf) : 0.0;
            return;
        }
    }
}

/* ------------------------------------------------------------------------------------
 * 7. Markov_process.jl:89-110 -- two-state chain, pi(t+1) = pi(t) P, P(down) per step;
 *    Markov_process.jl:159-195 -- DTMC capacity series with injected uniforms
 *    r[t*U + i] (unit order inside hour order, :172-175).
 * -------------------------------------------------------------------------------- */
void oracle_markov2(double lambda, double mu, double dt, int steps, double *prob_down)
{
    double p01 = 1 - exp(-lambda * dt), p10 = 1 - exp(-mu * dt);
    double p00 = 1 - p01, p11 = 1 - p10;
    double up = 1.0, down = 0.0;
    for (int t = 0; t < steps; t++) {
        double nu = up * p00 + down * p10;     /* row vector times P */
        double nd = up * p01 + down * p11;
        up = nu; down = nd;
        prob_down[t] = down;
    }
}

void oracle_dtmc_capacity(int U, const double *mttf, const double *mttr, const double *cap,
                          int T, const double *r, double *avail)
{
    int *state = (int *)calloc((size_t)U, sizeof(int));
    for (int t = 0; t < T; t++) {
        for (int i = 0; i < U; i++) {
            double p01 = 1 - exp(-(1 / mttf[i])), p10 = 1 - exp(-(1 / mttr[i]));
            double x = r[(size_t)t * U + i];
            if (state[i] == 0) { if (x < p01) state[i] = 1; }
            else               { if (x < p10) state[i] = 0; }
        }
        double c = 0.0;
        for (int i = 0; i < U; i++) if (state[i] == 0) c += cap[i];
        avail[t] = c;
    }
    free(state);
}

/* ------------------------------------------------------------------------------------
 * 8. RTS-79 hourly load factors -- Montecarlo_seq/anloducurve.m:24-88 (1-based hour).
 *    weekly[52], daily[7], hourly[24*6] row-major (hour, column) as in
 *    Montecarlo_seq/case24_loadprofile.m:23-73.
 * -------------------------------------------------------------------------------- */
void oracle_load_factors(int total_hours, const double *weekly, const double *daily,
                         const double *hourly, double *factors)
{
    for (int h = 1; h <= total_hours; h++) {
        int week = (int)ceil(h / 168.0);                        /* :27 */
        int season;                                             /* 0 winter 1 summer 2 spring/fall */
        if (week <= 8 || week >= 44) season = 0;
        else if (week >= 18 && week <= 30) season = 1;
        else season = 2;
        int day = (int)ceil(fmod(h / 24.0, 7.0));               /* :39 */
        if (day == 0) day = 7;
        int weekend = (day <= 5) ? 0 : 1;
        int hod = h % 24; if (hod == 0) hod = 24;               /* :49-50 */
        int col = 2 * season + weekend;                         /* :62-84 */
        factors[h - 1] = weekly[week - 1] * daily[day - 1] * hourly[(hod - 1) * 6 + col];
    }
}

/* ------------------------------------------------------------------------------------
 * 9. Hourly-resampled MC with maintenance / LFU / energy-limited units --
 *    tail_risk.jl:12-91 (run_detailed_mc) == MCvsMarkovProcess.jl:210-284 (run_monte_carlo)
 *    == generating_adequancy_comparative.jl:15-120.  SURVEY.md section 8 row f-1.
 *
 * Literal restatement: per year the ELU energy state is reset (:27); per hour, week =
 * div(h-1,168)+1 (:31); per unit in order: skipped while on maintenance (:39-42), out iff
 * rand() < FOR (:44), energy-limited units are unavailable once their state reached the limit
 * (:46-55); load = base + randn()*sigma (:59); unserved = max(0, load - unlimited capacity)
 * (:60); full or proportional drain of the available ELUs (:63-76); an hour with deficit > 0
 * counts once for the year and for the hour (:79-84).
 *
 * Injection: unif[(y*H + h)*U + u] stands for the rand() of unit u in hour h (consumed only
 * when the unit is not on maintenance, like the reference), norm[y*H + h] for randn().
 * energy_limit[u] = INFINITY for ordinary units.
 * -------------------------------------------------------------------------------- */
static void detailed_hour(int U, const double *cap, const double *elim, const unsigned char *avail_flag,
                          double load, double *energy, int *deficit_flag)
{
    /* avail_flag[u]: unit passed the maintenance and outage tests this hour */
    double cap_unlimited = 0.0, cap_elu = 0.0;
    int elu_idx[64]; int n_elu = 0;
    for (int i = 0; i < U; i++) {
        if (!avail_flag[i]) continue;
        if (elim[i] < INFINITY) {
            if (energy[i] >= elim[i]) continue;           /* exhausted */
            cap_elu += cap[i];
            elu_idx[n_elu++] = i;
        } else {
            cap_unlimited += cap[i];
        }
    }
    double unserved = load - cap_unlimited;
    if (unserved < 0.0) unserved = 0.0;                   /* max(0.0, ...) */
    double deficit = 0.0;
    if (unserved > 0) {
        if (unserved > cap_elu) {
            deficit = unserved - cap_elu;
            for (int k = 0; k < n_elu; k++) energy[elu_idx[k]] += cap[elu_idx[k]];
        } else {
            double needed = unserved;
            for (int k = 0; k < n_elu; k++) {
                double share = needed * (cap[elu_idx[k]] / cap_elu);
                energy[elu_idx[k]] += share;
            }
        }
    }
    *deficit_flag = deficit > 0;
}

int oracle_detailed_mc_injected(int U, const double *cap, const double *for_rate, const int *maint_start,
                                const int *maint_weeks, const double *elim, int H, const double *base_load,
                                double lfu_std, int n_years, const double *unif, const double *norm,
                                double *year_lole, double *hourly_fail)
{
    if (U > 64) return -1;
    double energy[64]; unsigned char avail[64];
    for (int h = 0; h < H; h++) hourly_fail[h] = 0.0;
    for (int y = 0; y < n_years; y++) {
        for (int i = 0; i < U; i++) energy[i] = 0.0;
        double count = 0.0;
        for (int h = 1; h <= H; h++) {
            int week = (h - 1) / 168 + 1;
            for (int i = 0; i < U; i++) {
                avail[i] = 0;
                if (week >= maint_start[i] && week < maint_start[i] + maint_weeks[i]) continue;
                if (unif[((size_t)y * H + (h - 1)) * U + i] < for_rate[i]) continue;
                avail[i] = 1;
            }
            double load = base_load[h - 1] + norm[(size_t)y * H + (h - 1)] * lfu_std;
            int d;
            detailed_hour(U, cap, elim, avail, load, energy, &d);
            if (d) { count += 1.0; hourly_fail[h - 1] += 1.0; }
        }
        year_lole[y] = count;
    }
    return 0;
}

/* Standard normal from two sampler words, fixed binary32 operation sequence (Box-Muller):
 *   r = sqrt(2 * max(E(x1), 0)),  E = oracle_neglog_u32 (which may dip to -1.2e-7 for draws next to 2^32);
 *   angle = 2 pi (4k + f) / 4 with k = x2 >> 30 and f = (2*((x2 >> 7) & 0x7FFFFF) + 1) / 2^24 in (0,1);
 *   phi = f * pi/2 folded to [0, pi/4] (swap sin/cos), Taylor polynomials by fma;  z = r * cos(angle). */
float oracle_normal_u32x2(uint32_t x1, uint32_t x2)
{
    float r = sqrtf(2.0f * fmaxf(oracle_neglog_u32(x1), 0.0f));
    uint32_t k = x2 >> 30;
    float f = (float)(2u * ((x2 >> 7) & 0x7FFFFFu) + 1u) * 5.9604644775390625e-08f;   /* exact */
    int swap = f > 0.5f;
    float g = swap ? 1.0f - f : f;                 /* exact */
    float x = g * 1.57079637f;
    float x2f = x * x;
    float s = fmaf(x2f, 2.75573192e-06f, -1.98412701e-04f);
    s = fmaf(x2f, s, 8.33333377e-03f);
    s = fmaf(x2f, s, -1.66666672e-01f);
    s = fmaf(x * x2f, s, x);                       /* sin x */
    float c = fmaf(x2f, 2.48015876e-05f, -1.38888892e-03f);
    c = fmaf(x2f, c, 4.16666679e-02f);
    c = fmaf(x2f, c, -0.5f);
    c = fmaf(x2f, c, 1.0f);                        /* cos x */
    float sn = swap ? c : s, cs = swap ? s : c;    /* sin(phi), cos(phi) */
    float v = (k == 0) ? cs : (k == 1) ? -sn : (k == 2) ? -cs : sn;
    return r * v;
}

/* Sampler-driven version: words of Philox blocks keyed (seed; year, hour, 0x444D0000 | blk):
 * word u decides unit u (OUT iff word < floor(FOR*2^32)), words U and U+1 give the normal. */
int oracle_detailed_mc_philox(int U, const double *cap, const uint32_t *for_thr, const int *maint_start,
                              const int *maint_weeks, const double *elim, int H, const double *base_load,
                              double lfu_std, uint64_t seed, int64_t year0, int n_years,
                              double *year_lole, double *hourly_fail)
{
    if (U > 62) return -1;
    double energy[64]; unsigned char avail[64]; uint32_t words[64 + 4];
    uint32_t key[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) };
    int nblk = (U + 2 + 3) / 4;
    for (int h = 0; h < H; h++) hourly_fail[h] = 0.0;
    for (int y = 0; y < n_years; y++) {
        uint64_t yy = (uint64_t)(year0 + y);
        for (int i = 0; i < U; i++) energy[i] = 0.0;
        double count = 0.0;
        for (int h = 1; h <= H; h++) {
            for (int b = 0; b < nblk; b++) {
                uint32_t ctr[4] = { (uint32_t)yy, (uint32_t)(yy >> 32), (uint32_t)(h - 1), 0x444D0000u | (uint32_t)b };
                oracle_philox4x32_10(ctr, key, &words[4 * b]);
            }
            int week = (h - 1) / 168 + 1;
            for (int i = 0; i < U; i++) {
                avail[i] = 0;
                if (week >= maint_start[i] && week < maint_start[i] + maint_weeks[i]) continue;
                if (words[i] < for_thr[i]) continue;
                avail[i] = 1;
            }
            double z = (double)oracle_normal_u32x2(words[U], words[U + 1]);
            double load = base_load[h - 1] + z * lfu_std;
            int d;
            detailed_hour(U, cap, elim, avail, load, energy, &d);
            if (d) { count += 1.0; hourly_fail[h - 1] += 1.0; }
        }
        year_lole[y] = count;
    }
    return 0;
}


/* ------------------------------------------------------------------------------------
 * 10. MATLAB next-event discretisation -- Montecarlo_seq/seq_mcsampling.m:40-74 (SURVEY a-8),
 *     evaluated at HL1 (capacity vs load, no network): every year all components start UP at
 *     time 0 (:40-41, seqMain.m:91); time to failure = round(-MTTF ln u) (:52-53, MATLAB round =
 *     half away from zero), time to repair = ceil(-MTTR ln u) (:59-60); the DOWN hours are
 *     start = round(t)+1 ... min(start+dur-1, H) (:63-67).  Durations come from the sampler
 *     streams of section 2 (draw 0 of a unit is consumed and ignored, as in ALL_UP mode).
 * -------------------------------------------------------------------------------- */
int oracle_seq_matlab_philox(int U, const double *cap, const float *mttf_f, const float *mttr_f, int H,
                             const double *load, uint64_t seed, int64_t year0, int64_t nyears,
                             double *year_lol, double *year_eue, double *year_entries)
{
    unsigned char *down = (unsigned char *)malloc((size_t)U * (size_t)H);
    draw_stream st;
    for (int64_t y = 0; y < nyears; y++) {
        memset(down, 0, (size_t)U * (size_t)H);
        for (int i = 0; i < U; i++) {
            stream_init(&st, seed, (uint64_t)(year0 + y), (uint32_t)i);
            (void)stream_next(&st);
            double current_time = 0; int is_up = 1;
            while (current_time < H) {
                if (is_up) {
                    double duration = duration_hours(mttf_f[i], stream_next(&st));
                    double duration_int = round(duration);
                    current_time = current_time + duration_int;
                } else {
                    double duration = duration_hours(mttr_f[i], stream_next(&st));
                    double duration_int = ceil(duration);
                    long start_idx = (long)round(current_time) + 1;
                    long end_idx = start_idx + (long)duration_int - 1;
                    if (end_idx > H) end_idx = H;
                    if (start_idx <= H)
                        for (long h = start_idx; h <= end_idx; h++) down[(size_t)i * H + (h - 1)] = 1;
                    current_time = current_time + duration_int;
                }
                is_up = !is_up;
            }
        }
        double lole = 0.0, eue = 0.0, entries = 0.0; int prev = 0;
        for (int h = 0; h < H; h++) {
            double cap_avail = 0.0;
            for (int i = 0; i < U; i++) if (!down[(size_t)i * H + h]) cap_avail += cap[i];
            int flag = 0;
            if (cap_avail < load[h]) { lole += 1.0; eue += load[h] - cap_avail; flag = 1; }
            if (flag && !prev) entries += 1.0;
            prev = flag;
        }
        year_lol[y] = lole; year_eue[y] = eue; year_entries[y] = entries;
    }
    free(down);
    return 0;
}

/* ------------------------------------------------------------------------------------
 * 11. Analytical side of tail_risk.jl / generating_adequacy_comprehensive.jl (host-side planning):
 *     calculate_expected_generation (comprehensive.jl:118-142) and the hourly LFU risk loop
 *     (tail_risk.jl:124-136 == comprehensive.jl:251-265), literal loops over the 7-step LFU table
 *     (comprehensive.jl:76-80).  `installed` / `cap_rest` are the LAST GRID STATE of the table.
 * -------------------------------------------------------------------------------- */
static const double LFU_Z[7] = { -3.0, -2.0, -1.0, 0.0, 1.0, 2.0, 3.0 };
static const double LFU_P[7] = { 0.006, 0.061, 0.242, 0.382, 0.242, 0.061, 0.006 };

double oracle_expected_generation(const double *probs, int n, double step, double unit_cap,
                                  const double *loads, int H, double lfu_sigma)
{
    double total_energy = 0.0;
    double cap_rest = (double)(n - 1) * step;
    for (int h = 0; h < H; h++) {
        double hourly_e = 0.0;
        for (int k = 0; k < 7; k++) {
            double actual_load = loads[h] + (LFU_Z[k] * lfu_sigma);
            double reserve_thresh = cap_rest - actual_load;
            double term_e = 0.0;
            for (int i = 0; i < n; i++) {
                double outage = (double)i * step;
                if (outage > reserve_thresh) {
                    double deficit = outage - reserve_thresh;
                    term_e += (unit_cap < deficit ? unit_cap : deficit) * probs[i];
                }
            }
            hourly_e += term_e * LFU_P[k];
        }
        total_energy += hourly_e;
    }
    return total_energy;
}

void oracle_lfu_hourly_risk(const double *probs, int n, double step, const double *loads, int H,
                            double lfu_mw, double *risk)
{
    double installed = (double)(n - 1) * step;
    for (int h = 0; h < H; h++) {
        double risk_h = 0.0;
        for (int k = 0; k < 7; k++) {
            double res = installed - (loads[h] + LFU_Z[k] * lfu_mw);
            for (int i = 0; i < n; i++)
                if ((double)i * step > res) risk_h += probs[i] * LFU_P[k];
        }
        risk[h] = risk_h;
    }
}

/* ------------------------------------------------------------------------------------
 * 9. Multi-area adequacy -- GeneratingAdequacy/AdequacyAssessmentII.jl.
 *
 * oracle_solve_curtailment: solve_curtailment_fast (:73-179) verbatim in FP64, 0-based:
 *   all margins >= 0 -> zeros (:78-80); ISOLATED -> max(0, -margin) (:84-92);
 *   INTERCONNECTED: repeat { source = first area with margin > 1e-4, sink = first with
 *   margin < -1e-4 (:107-108); stop if either is missing (:111-113); BFS from the source in
 *   area order over residual capacities > 1e-4, never re-entering the source (:116-134);
 *   stop if the sink was not reached (:136-147); path flow = min(surplus, deficit, residuals)
 *   (:150-156); apply, with reverse residuals (:159-168) }; curtailment = remaining deficits (:171-176).
 * topo[i*n+j] is System.topology_matrix (:52-61: tie capacities added in both directions).
 * -------------------------------------------------------------------------------- */
#define ORACLE_MAX_AREAS 16
void oracle_solve_curtailment(int n, const double *topo, const double *margins, int policy, double *curt)
{
    int all_ok = 1;
    for (int i = 0; i < n; i++) { curt[i] = 0.0; if (!(margins[i] >= 0)) all_ok = 0; }
    if (all_ok) return;
    if (policy == 0) {                                       /* ISOLATED */
        for (int i = 0; i < n; i++) if (margins[i] < 0) curt[i] = -margins[i];
        return;
    }
    double res[ORACLE_MAX_AREAS * ORACLE_MAX_AREAS], m[ORACLE_MAX_AREAS];
    for (int i = 0; i < n * n; i++) res[i] = topo[i];
    for (int i = 0; i < n; i++) m[i] = margins[i];
    for (;;) {
        int src = -1, snk = -1;
        for (int i = 0; i < n; i++) if (m[i] > 1e-4) { src = i; break; }
        for (int i = 0; i < n; i++) if (m[i] < -1e-4) { snk = i; break; }
        if (src < 0 || snk < 0) break;
        int parent[ORACLE_MAX_AREAS], queue[ORACLE_MAX_AREAS], qh = 0, qt = 0, found = 0;
        for (int i = 0; i < n; i++) parent[i] = -1;
        queue[qt++] = src;
        while (qh < qt) {
            int u = queue[qh++];
            if (u == snk) { found = 1; break; }
            for (int v = 0; v < n; v++)
                if (res[u * n + v] > 1e-4 && parent[v] < 0 && v != src) { parent[v] = u; queue[qt++] = v; }
        }
        if (!found) break;
        double f = m[src] < -m[snk] ? m[src] : -m[snk];
        for (int c = snk; c != src; c = parent[c]) { double r = res[parent[c] * n + c]; if (r < f) f = r; }
        m[src] -= f; m[snk] += f;
        for (int c = snk; c != src; c = parent[c]) { res[parent[c] * n + c] -= f; res[c * n + parent[c]] += f; }
    }
    for (int i = 0; i < n; i++) if (m[i] < 0) curt[i] = -m[i];
}

/* run_fast_sequential_simulation (:185-250): literal hour / area / generator loop; every -log(rand()) * mean of
 * :25,209,212 becomes the sampler duration of the unit's own (chain, global unit index) stream, exactly as in
 * oracle_seq_philox (section 2), with the same init modes and one chain per year.  Units are listed area by
 * area (unit_area[u] non-decreasing).  Outputs per year and area: hours with curtailment > 0 (:231-232) and
 * the curtailed energy (:233). */
int oracle_multi_area_philox(int n_areas, int U, const int *unit_area, const double *cap, const float *mttf_f,
                             const float *mttr_f, const uint32_t *for_thr, int H, const double *load /*[n_areas][H]*/,
                             const double *topo, int policy, uint64_t seed, int64_t year0, int64_t nyears,
                             int init_mode, double *year_lol /*[nyears][n_areas]*/, double *year_eue)
{
    if (n_areas > ORACLE_MAX_AREAS) return -1;
    draw_stream *st = (draw_stream *)malloc(sizeof(draw_stream) * (size_t)U);
    unsigned char *status = (unsigned char *)malloc((size_t)U);
    double *ttf = (double *)malloc(sizeof(double) * (size_t)U);
    double margins[ORACLE_MAX_AREAS], curt[ORACLE_MAX_AREAS];
    for (int64_t y = 0; y < nyears; y++) {
        for (int i = 0; i < U; i++) {
            stream_init(&st[i], seed, (uint64_t)(year0 + y), (uint32_t)i);
            uint32_t x0 = stream_next(&st[i]);
            status[i] = (init_mode == 1 && x0 < for_thr[i]) ? 0 : 1;
            ttf[i] = duration_hours(status[i] ? mttf_f[i] : mttr_f[i], stream_next(&st[i]));
        }
        for (int a = 0; a < n_areas; a++) { year_lol[y * n_areas + a] = 0.0; year_eue[y * n_areas + a] = 0.0; }
        for (int h = 0; h < H; h++) {
            for (int a = 0; a < n_areas; a++) margins[a] = 0.0;
            for (int i = 0; i < U; i++) {
                ttf[i] -= 1.0;
                while (ttf[i] <= 0) {
                    status[i] = !status[i];
                    ttf[i] += duration_hours(status[i] ? mttf_f[i] : mttr_f[i], stream_next(&st[i]));
                }
                if (status[i]) margins[unit_area[i]] += cap[i];
            }
            for (int a = 0; a < n_areas; a++) margins[a] -= load[(size_t)a * H + h];
            oracle_solve_curtailment(n_areas, topo, margins, policy, curt);
            for (int a = 0; a < n_areas; a++)
                if (curt[a] > 0) { year_lol[y * n_areas + a] += 1.0; year_eue[y * n_areas + a] += curt[a]; }
        }
    }
    free(st); free(status); free(ttf);
    return 0;
}

/* ------------------------------------------------------------------------------------
 * 10. Markov_process.jl:39-60 -- constant hourly failure probability => exponential failure times.
 * Literal loop; rand() is r[i*K + k] (injected) or u = word / 2^32 of the Philox stream keyed
 * (seed; component, 0x46540000 | block).  out[i] = pushed time, -1 if the component outlived max_time.
 * -------------------------------------------------------------------------------- */
int oracle_failure_times(double lambda, double dt, double max_time, int64_t n, uint64_t seed, const double *r, int K, double *out)
{
    for (int64_t i = 0; i < n; i++) {
        double t = 0.0;
        uint32_t key[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) }, ctr[4], buf[4] = {0, 0, 0, 0};
        out[i] = -1.0;
        for (int k = 0;; k++) {
            double u;
            if (r) { if (k >= K) return -1; u = r[(size_t)i * K + k]; }
            else {
                if ((k & 3) == 0) {
                    ctr[0] = (uint32_t)i; ctr[1] = (uint32_t)((uint64_t)i >> 32); ctr[2] = 0x46540000u; ctr[3] = (uint32_t)(k >> 2);
                    oracle_philox4x32_10(ctr, key, buf);
                }
                u = (double)buf[k & 3] * 2.3283064365386963e-10;
            }
            if (u < lambda * dt) { out[i] = t; break; }      /* :54-56 */
            t += dt;                                          /* :58 */
            if (t > max_time) break;                          /* :59 */
        }
    }
    return 0;
}
