/* ============================================================================================
 * TEST INFRASTRUCTURE -- NOT PRODUCT CODE.
 *
 * CPU restatement ("oracle") of the SPoNet hot path: the CNVM / CNTM event loops, the sampling
 * primitives, the snapshot rule and the OpinionShares histogram, in plain C.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may load it.  The
 * product (sponet_b200/) never links, imports or calls anything in oracle/.
 *
 * Parity status: PINNED.  The reference ships no golden trajectories, so
 * the oracle is pinned against outputs of the unmodified reference itself, run in the build
 * container: tests/golden/make_golden.py imports /root/reference, runs CNVM/CNTM.simulate with
 * Generator(PCG64(seed)) and stores (raw u64 stream seed, t_traj, x_traj) fixtures under
 * tests/golden/.  tests/test_oracle_golden.py replays them through this file bit-exactly.
 *
 * Third-party arithmetic restated here (not under /root/reference):
 *   numba  (pyproject.toml:25 `numba>=0.56`; poetry.lock pins 0.64.0; validated with 0.65.0)
 *     next_double            numba/np/random/generator_core.py:98-107  -> (u64 >> 11) * 2^-53
 *     random_standard_exponential   numba/np/random/distributions.py:105-119 (256-layer ziggurat)
 *     random_exponential     numba/np/random/distributions.py:244-245  -> scale * std
 *   numpy  (poetry.lock pins 2.2.6; validated with 2.3.5)  PCG64 bit stream (pcg64.h, XSL-RR 128/64)
 *   libm   exp / log1p as called by numba's lowering (glibc 2.39 here)
 * ============================================================================================ */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../sponet_b200/csrc/ziggurat_tables.h"

/* ------------------------------------------------------------------------------------------ */
/* u64 source: an injected buffer (replay) or a live PCG64 (baseline timing).                  */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    const uint64_t* buf; /* NULL => live PCG64 */
    int64_t len, pos;
    int overflow;
    unsigned __int128 state, inc; /* numpy PCG64: state = state * MULT + inc; XSL-RR output */
} ostream;

#define PCG_MULT ((((unsigned __int128)0x2360ED051FC65DA4ULL) << 64) | 0x4385DF649FCCF645ULL)

static inline uint64_t os_u64(ostream* s) {
    if (s->buf) {
        if (s->pos >= s->len) {
            s->overflow = 1;
            return 0;
        }
        return s->buf[s->pos++];
    }
    s->state = s->state * PCG_MULT + s->inc;
    uint64_t hi = (uint64_t)(s->state >> 64), lo = (uint64_t)s->state;
    uint64_t v = hi ^ lo;
    unsigned rot = (unsigned)(hi >> 58);
    s->pos++;
    return (v >> rot) | (v << ((-rot) & 63));
}

/* numba/np/random/generator_core.py:98-107 (numpy's next_double) */
static inline double os_double(ostream* s) { return (double)(os_u64(s) >> 11) * (1.0 / 9007199254740992.0); }

/* numba/np/random/distributions.py:105-119 */
static inline double os_std_exponential(ostream* s) {
    for (;;) {
        uint64_t ri = os_u64(s);
        if (s->overflow) return 0.0;
        ri >>= 3;
        unsigned idx = (unsigned)(ri & 0xFF);
        ri >>= 8;
        double x = (double)ri * sponet_we_double[idx];
        if (ri < sponet_ke_double[idx]) return x;
        if (idx == 0) return SPONET_ZIGGURAT_EXP_R - log1p(-os_double(s));
        if ((sponet_fe_double[idx - 1] - sponet_fe_double[idx]) * os_double(s) + sponet_fe_double[idx] <
            exp(-x))
            return x;
        if (s->overflow) return 0.0;
    }
}

/* sponet/sampling.py:54-96 build_alias_table -- Vose variant with python-list stacks; the output
 * depends on pop order, so the stacks are reproduced literally. */
void oracle_build_alias_table(const double* weights, int64_t n, double* table_prob,
                              int32_t* table_alias) {
    double sum = 0.0;
    for (int64_t i = 0; i < n; ++i) sum += weights[i]; /* numba np.sum: sequential */
    int64_t* small_ids = (int64_t*)malloc((size_t)n * sizeof(int64_t));
    int64_t* large_ids = (int64_t*)malloc((size_t)n * sizeof(int64_t));
    int64_t ns = 0, nl = 0;
    for (int64_t i = 0; i < n; ++i) {
        table_prob[i] = weights[i] / sum * (double)n; /* :71 */
        table_alias[i] = 0;
    }
    for (int64_t i = 0; i < n; ++i) {
        if (table_prob[i] < 1) small_ids[ns++] = i;
        else large_ids[nl++] = i;
    }
    while (ns > 0 && nl > 0) { /* :77-90 */
        int64_t small_id = small_ids[--ns];
        int64_t large_id = large_ids[--nl];
        table_alias[small_id] = (int32_t)large_id;
        table_prob[large_id] = table_prob[large_id] + table_prob[small_id] - 1;
        if (table_prob[large_id] < 1) small_ids[ns++] = large_id;
        else large_ids[nl++] = large_id;
    }
    while (nl > 0) table_prob[large_ids[--nl]] = 1; /* :92-93 */
    while (ns > 0) table_prob[small_ids[--ns]] = 1; /* :95-96 */
    free(small_ids);
    free(large_ids);
}

/* ------------------------------------------------------------------------------------------ */
/* CNVM loops, instantiated for uint8 (M <= 256) and uint16 (M >= 257) agent states            */
/* (sponet/cnvm/model.py:98: opinion_dtype = np.min_scalar_type(M - 1)).                        */
/* ------------------------------------------------------------------------------------------ */
#define XT uint8_t
#define FN(name) name##_u8
#include "oracle_loops.inc"
#undef XT
#undef FN
#define XT uint16_t
#define FN(name) name##_u16
#include "oracle_loops.inc"
#undef XT
#undef FN

/* ------------------------------------------------------------------------------------------ */
/* CNTM (always uint8: sponet/cntm/model.py:63)                                                */
/* ------------------------------------------------------------------------------------------ */

/* sponet/cntm/model.py:163-227 _simulate_teval. */
int64_t oracle_cntm_teval(uint8_t* x, int64_t n, const double* t_eval, int64_t T,
                          double next_event_rate, double noise_prob, double threshold_01,
                          double threshold_10, const int32_t* row_ptr, const int32_t* col,
                          ostream* s, double* t_out, uint8_t* x_out) {
    memcpy(x_out, x, (size_t)n);
    double t = t_eval[0];
    t_out[0] = t_eval[0];
    int64_t previous_agent = 0;
    uint8_t previous_opinion = x[0];
    double previous_t = t;
    int64_t idx = 1, iters = 0;
    while (idx < T) { /* :189 */
        ++iters;
        int64_t agent = (int64_t)(os_double(s) * (double)n); /* :190 */
        if (os_double(s) < noise_prob) { /* :191-195 -- previous_* recorded UNCONDITIONALLY */
            previous_t = t;
            previous_agent = agent;
            previous_opinion = x[agent];
            x[agent] = (uint8_t)(int64_t)(os_double(s) * 2.0);
        } else {
            int64_t beg = row_ptr[agent], deg = row_ptr[agent + 1] - beg;
            if (deg > 0) { /* :198 */
                uint8_t other = (uint8_t)(1 - x[agent]);
                int64_t count = 0;
                for (int64_t j = 0; j < deg; ++j) count += (x[col[beg + j]] == other);
                double share = (double)count / (double)deg; /* :204 */
                double threshold = (x[agent] == 1) ? threshold_10 : threshold_01;
                if (share >= threshold) { /* :207-211 */
                    previous_t = t;
                    previous_agent = agent;
                    previous_opinion = x[agent];
                    x[agent] = other;
                }
            }
        }
        t += next_event_rate * os_std_exponential(s); /* :213 */
        if (s->overflow) { iters = -1; break; }
        if (t >= t_eval[idx]) {
            store_snapshot_u8(t, t_eval[idx], previous_t, x, n, previous_agent, previous_opinion,
                              x_out + idx * n, t_out + idx);
            ++idx;
        }
    }
    return iters;
}

/* sponet/cntm/model.py:111-160 _simulate_all as an event log (t, agent, new opinion). */
int64_t oracle_cntm_all(uint8_t* x, int64_t n, double t_max, double next_event_rate,
                        double noise_prob, double threshold_01, double threshold_10,
                        const int32_t* row_ptr, const int32_t* col, ostream* s, int64_t cap,
                        double* log_t, int32_t* log_agent, int32_t* log_opinion,
                        int64_t* iters_out) {
    double t = 0.0;
    int64_t cnt = 0, iters = 0;
    while (t < t_max) { /* :130 */
        ++iters;
        t += next_event_rate * os_std_exponential(s);        /* :131 */
        int64_t agent = (int64_t)(os_double(s) * (double)n); /* :132 */
        int update = 0;
        if (os_double(s) < 0.5 * noise_prob) { /* :136-140 */
            x[agent] = (uint8_t)(1 - x[agent]);
            update = 1;
        } else {
            int64_t beg = row_ptr[agent], deg = row_ptr[agent + 1] - beg;
            if (deg > 0) {
                uint8_t other = (uint8_t)(1 - x[agent]);
                int64_t count = 0;
                for (int64_t j = 0; j < deg; ++j) count += (x[col[beg + j]] == other);
                double share = (double)count / (double)deg;
                double threshold = (x[agent] == 1) ? threshold_10 : threshold_01;
                if (share >= threshold) {
                    x[agent] = other;
                    update = 1;
                }
            }
        }
        if (s->overflow) { cnt = -1; break; }
        if (update) {
            if (cnt >= cap) { cnt = -2; break; }
            log_t[cnt] = t;
            log_agent[cnt] = (int32_t)agent;
            log_opinion[cnt] = (int32_t)x[agent];
            ++cnt;
        }
    }
    *iters_out = iters;
    return cnt;
}

/* ------------------------------------------------------------------------------------------ */
/* Collective variables                                                                         */
/* ------------------------------------------------------------------------------------------ */

/* sponet/collective_variables.py:400-405 _opinion_shares_numba: np.bincount(x[i], weights,
 * minlength=M) per snapshot -- out[x[i,j]] += w[j] sequentially in agent order. */
void oracle_opinion_shares_u8(const uint8_t* x, int64_t S, int64_t n, int64_t M,
                              const double* weights, double* out) {
    for (int64_t i = 0; i < S; ++i) {
        double* o = out + i * M;
        for (int64_t m = 0; m < M; ++m) o[m] = 0.0;
        const uint8_t* xi = x + i * n;
        if (weights) for (int64_t j = 0; j < n; ++j) o[xi[j]] += weights[j];
        else for (int64_t j = 0; j < n; ++j) o[xi[j]] += 1.0;
    }
}

/* sponet/utils.py:7-45 argmatch. */
void oracle_argmatch(const double* x_ref, int64_t size, const double* x, int64_t nx, int64_t* out) {
    int64_t ref_ind = 0, ind = 0;
    for (int64_t i = 0; i < size; ++i) out[i] = 0;
    while (x_ref[ref_ind] < x[ind]) ++ref_ind;
    while (ref_ind < size && ind < nx - 1) {
        if (x[ind] <= x_ref[ref_ind] && x_ref[ref_ind] <= x[ind + 1]) {
            if (fabs(x[ind] - x_ref[ref_ind]) < fabs(x[ind + 1] - x_ref[ref_ind])) out[ref_ind] = ind;
            else out[ref_ind] = ind + 1;
            ++ref_ind;
        } else {
            ++ind;
        }
    }
    while (ref_ind < size) out[ref_ind++] = nx - 1;
}

/* ------------------------------------------------------------------------------------------ */
/* Native-RNG mode (Philox4x32-10, Poisson interval counts): sequential statement                */
/* ------------------------------------------------------------------------------------------ */
#include "oracle_native.inc"

/* ------------------------------------------------------------------------------------------ */
/* Stream plumbing for the ctypes wrapper                                                       */
/* ------------------------------------------------------------------------------------------ */
ostream* oracle_stream_from_buffer(const uint64_t* buf, int64_t len) {
    ostream* s = (ostream*)calloc(1, sizeof(ostream));
    s->buf = buf;
    s->len = len;
    return s;
}

/* (state, inc) exactly as numpy exposes them in PCG64().state['state'] (128-bit, hi/lo halves). */
ostream* oracle_stream_pcg64(uint64_t state_hi, uint64_t state_lo, uint64_t inc_hi, uint64_t inc_lo) {
    ostream* s = (ostream*)calloc(1, sizeof(ostream));
    s->state = (((unsigned __int128)state_hi) << 64) | state_lo;
    s->inc = (((unsigned __int128)inc_hi) << 64) | inc_lo;
    return s;
}

int64_t oracle_stream_pos(const ostream* s) { return s->pos; }
int oracle_stream_overflow(const ostream* s) { return s->overflow; }
void oracle_stream_free(ostream* s) { free(s); }
uint64_t oracle_stream_next_u64(ostream* s) { return os_u64(s); }
double oracle_stream_next_exponential(ostream* s) { return os_std_exponential(s); }

/* ------------------------------------------------------------------------------------------ */
/* CPU baseline: the ensemble driver of sponet/multiprocessing.py:134-187 (_Worker.__call__) with   */
/* OpinionShares as CV (collective_variables.py:98-124), threads over realizations instead of   */
/* processes.  Each realization gets its own PCG64 stream (numpy's SeedSequence is not restated; */
/* streams are seeded as state = splitmix(seed, run), inc odd).  Used ONLY by bench.py's         */
/* cpu_baseline / --impl reference legs to time the reference algorithm on the host cores.       */
/* ------------------------------------------------------------------------------------------ */
static uint64_t splitmix64(uint64_t* z) {
    uint64_t x = (*z += 0x9E3779B97F4A7C15ULL);
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}

static void seed_run_stream(ostream* s, uint64_t seed, int64_t run) {
    uint64_t z = seed ^ (0xD1B54A32D192ED03ULL * (uint64_t)(run + 1));
    memset(s, 0, sizeof(*s));
    s->state = (((unsigned __int128)splitmix64(&z)) << 64) | splitmix64(&z);
    s->inc = ((((unsigned __int128)splitmix64(&z)) << 64) | splitmix64(&z)) | 1;
    (void)os_u64(s);
    s->pos = 0;
}

/* model: 0 = CNVM network, 1 = CNVM complete, 2 = CNTM.  out_cv: [R, T, M] f64 counts.
 * Returns total loop iterations over all runs.  Threads (pthreads) pull run indices from a shared
 * counter -- the analogue of ProcessPoolExecutor.map over run chunks (multiprocessing.py:96-108). */
typedef struct {
    int model;
    const uint8_t* x_init;
    int64_t n, R, T, M;
    const double* t_eval;
    const int32_t *row_ptr, *col;
    double r_imit, r_noise;
    const double *prob_imit, *prob_noise, *degree_alpha;
    double next_event_rate, noise_prob, threshold_01, threshold_10;
    uint64_t seed;
    double* out_cv;
    int64_t next_run; /* atomic */
    int64_t total;    /* atomic */
} many_runs_job;

static void* many_runs_worker(void* arg) {
    many_runs_job* j = (many_runs_job*)arg;
    const int64_t n = j->n, T = j->T, M = j->M;
    uint8_t* x = (uint8_t*)malloc((size_t)n);
    uint8_t* x_traj = (uint8_t*)malloc((size_t)(T * n));
    double* t_traj = (double*)malloc((size_t)T * sizeof(double));
    int64_t local = 0;
    for (;;) {
        int64_t r = __atomic_fetch_add(&j->next_run, 1, __ATOMIC_RELAXED);
        if (r >= j->R) break;
        ostream s;
        seed_run_stream(&s, j->seed, r);
        memcpy(x, j->x_init, (size_t)n);
        int64_t it;
        if (j->model == 0)
            it = oracle_cnvm_teval_u8(x, n, j->t_eval, T, M, j->row_ptr, j->col, j->r_imit,
                                      j->r_noise, j->prob_imit, j->prob_noise, j->degree_alpha, &s,
                                      t_traj, x_traj);
        else if (j->model == 1)
            it = oracle_cnvm_complete_teval_u8(x, n, j->t_eval, T, M, j->r_imit, j->r_noise,
                                               j->prob_imit, j->prob_noise, j->degree_alpha[0], &s,
                                               t_traj, x_traj);
        else
            it = oracle_cntm_teval(x, n, j->t_eval, T, j->next_event_rate, j->noise_prob,
                                   j->threshold_01, j->threshold_10, j->row_ptr, j->col, &s, t_traj,
                                   x_traj);
        oracle_opinion_shares_u8(x_traj, T, n, M, NULL, j->out_cv + r * T * M);
        local += it;
    }
    __atomic_fetch_add(&j->total, local, __ATOMIC_RELAXED);
    free(x);
    free(x_traj);
    free(t_traj);
    return NULL;
}

int64_t oracle_sample_many_runs_u8(int model, const uint8_t* x_init, int64_t n, int64_t R,
                                   const double* t_eval, int64_t T, int64_t M,
                                   const int32_t* row_ptr, const int32_t* col, double r_imit,
                                   double r_noise, const double* prob_imit, const double* prob_noise,
                                   const double* degree_alpha, double next_event_rate,
                                   double noise_prob, double threshold_01, double threshold_10,
                                   uint64_t seed, int num_threads, double* out_cv) {
    many_runs_job job = {model, x_init, n, R, T, M, t_eval, row_ptr, col, r_imit, r_noise,
                         prob_imit, prob_noise, degree_alpha, next_event_rate, noise_prob,
                         threshold_01, threshold_10, seed, out_cv, 0, 0};
    if (num_threads < 1) num_threads = 1;
    pthread_t* th = (pthread_t*)malloc((size_t)num_threads * sizeof(pthread_t));
    for (int i = 0; i < num_threads; ++i) pthread_create(&th[i], NULL, many_runs_worker, &job);
    for (int i = 0; i < num_threads; ++i) pthread_join(th[i], NULL);
    free(th);
    return job.total;
}
