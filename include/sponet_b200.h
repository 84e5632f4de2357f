/* sponet_b200 -- C ABI of the B200-native SPoNet ensemble path.
 *
 * The reference (lueckem/SPoNet v3.0.0) has no FFI: its boundary is a Python call surface over six
 * private numba loops.  Every entry point below replaces one of those loops (or the host arithmetic
 * wrapped around them) for a whole batch of realizations, and cites the reference site it stands
 * in for.  INTEGRATION.md shows the ctypes stub a SPoNet maintainer would add.
 *
 * Conventions
 *   - Plain pointers and sizes only.  Every data pointer may be a HOST pointer or a DEVICE pointer
 *     (detected with cudaPointerGetAttributes); host buffers are staged through device memory
 *     inside the call.  The caller owns all buffers; the library owns graph handles and scratch.
 *   - All calls run on the CURRENT CUDA device, on `cuda_stream` (a cudaStream_t, NULL = default
 *     stream).  A call that received only device pointers returns without synchronising; a call
 *     with any host output synchronises the stream before returning.
 *   - Return value: SPONET_OK or an error code; sponet_last_error() gives the message (thread
 *     local).  There is NO CPU fallback: without a CUDA device every run call fails.
 *   - Agent states are uint8 for num_opinions <= 256 and uint16 above
 *     (sponet/cnvm/model.py:98, sponet/cntm/model.py:63).
 */
#ifndef SPONET_B200_H
#define SPONET_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPONET_ABI_VERSION 4

enum {
    SPONET_OK = 0,
    SPONET_ERR_INVALID = 1,          /* bad argument (-> ValueError on the Python side)            */
    SPONET_ERR_CUDA = 2,             /* CUDA runtime error or no device                            */
    SPONET_ERR_STREAM_EXHAUSTED = 3, /* replay: injected uint64 stream too short                   */
    SPONET_ERR_UNSUPPORTED = 4,      /* valid request this build cannot serve                      */
    SPONET_ERR_ZERO_RATE = 5         /* all rates zero: 1/0 in next_event_rate (cnvm/model.py:257)  */
};

const char* sponet_last_error(void);
int sponet_abi_version(void);
/* Number of visible CUDA devices (0 and SPONET_ERR_CUDA when there is none). */
int sponet_device_count(int* count);

/* ---------------------------------------------------------------------------------------------
 * Host-side setup arithmetic (sequential float64, bit-compatible with the numba code)
 * ------------------------------------------------------------------------------------------- */

/* sponet/sampling.py:54-96 build_alias_table (stack-based Vose; output depends on pop order). */
int sponet_build_alias_table(const double* weights, int64_t n, double* table_prob, int32_t* table_alias);

/* sponet/cnvm/model.py:257-258 (network) / :383-384 (complete, degree_alpha = NULL and
 * max_degree_alpha used).  Sequential sum like numba's np.sum. */
int sponet_cnvm_event_rates(const double* degree_alpha, int64_t n, double max_degree_alpha,
                            double r_imit, double r_noise, double* next_event_rate,
                            double* noise_probability);

/* ---------------------------------------------------------------------------------------------
 * Graph handle: CSR adjacency resident in HBM (replaces numba.typed.List(params.network),
 * sponet/cnvm/model.py:26-28, sponet/cntm/model.py:29).  Row order is preserved: neighbour slot k
 * is what `int(u * deg)` indexes (cnvm/model.py:286).
 * ------------------------------------------------------------------------------------------- */
typedef struct sponet_graph sponet_graph;

int sponet_graph_create(int64_t num_agents, const int32_t* row_ptr /* [n+1] */,
                        const int32_t* col /* [row_ptr[n]] */, sponet_graph** out);
int sponet_graph_destroy(sponet_graph* g);
int sponet_graph_info(const sponet_graph* g, int64_t* num_agents, int64_t* nnz, int32_t* min_degree,
                      int32_t* max_degree, int* device);

/* Neighbour sampler table of native mode.  For networks with 1 <= degree <= 64 and at most 2^20 agents
 * whose table fits 64 MiB, graph_create also builds L = 2^log2_slots >= max_degree slots per agent:
 * neighbour k of a d-neighbour row owns the slot interval [k L / d, (k+1) L / d).  Entry (agent, slot j)
 * = word0 | word1 << 32 with word0 = primary id | (thr >> 12) << 20, word1 = alias id | (thr & 0xFFF) << 20,
 * primary = floor(j d / L), alias = primary + 1, thr = floor(((primary+1) L - j d) / d * 2^24) (0xFFFFFF and
 * alias = primary when the slot lies inside one neighbour).  The CNVM kernel then draws
 *     slot = r2 >> (32 - log2_slots), frac = ((r2 << log2_slots) >> 8), nbr = frac < thr ? primary : alias
 * -- one 8-byte gather instead of the dependent row -> column pair; each neighbour keeps probability 1/d
 * up to 2^-24 per slot.  *log2_slots = -1 when the graph has no table (the kernel then uses the CSR:
 * nbr = col[begin + mulhi(r2, d)]).  `table` (HOST, [n << log2_slots]) may be NULL. */
int sponet_graph_neighbor_table(const sponet_graph* g, int32_t* log2_slots, uint64_t* table);

/* A BATCH of Erdos-Renyi graphs G(num_agents, p), generated on the device: one graph per chain of the native call
 * that covers the same (num_states, runs [run_begin, run_end) of num_runs_total) -- chain (state j, run i) runs on
 * graph index c = j * num_runs_total + i.  Replaces the per-run `update_network()` of the reference
 * (sponet/cnvm/model.py:95-96 with sponet/network_generator.py:26-85 ErdosRenyiGenerator.__call__, called once per
 * realization by sponet/multiprocessing.py:173-178), i.e. ~0.7 s of networkx per realization at N = 1e5.
 * Graph c depends on (seed, c) only: not on the batch it is part of, nor on the number of GPUs.
 * Construction (restated sequentially by oracle/oracle_native.inc, bit-identical): vertex i owns the pairs (i, j > i);
 * from j = i repeat j += 1 + floor(log1p(-v) / log1p(-p)), v = 53 random bits * 2^-53, edge {i, j} while j < n; draw k
 * of vertex i = words (2 (k & 1), 2 (k & 1) + 1) of Philox4x32-10(counter (k >> 1, i, c_lo, c_hi[16] | attempt << 16 |
 * 3 << 24), key seed).  force_no_isolates = 0: a graph with an isolated vertex is rejected and drawn again (attempt + 1;
 * SPONET_ERR_INVALID "Timeout. ..." after max_attempts, the reference's RuntimeError :75-78).  force_no_isolates = 1
 * (reference :446-454): vertices in order, one that is still isolated gets an edge to j = mulhi(word 0 of Philox(counter
 * (k, i, c_lo, c_hi | 4 << 24)), n), k = 0, 1, ... until j != i.  Rows sorted by neighbour id.
 * layouts: SPONET_BATCH_PADDED_ROWS also builds the padded rows the CNTM kernel scans; SPONET_BATCH_RECORDS also builds
 * one 64-byte record per agent (degree, offset of its row, first 15 neighbours) for the CNVM kernel: the graphs of a
 * batch do not fit the L2, and with the record the degree and the sampled neighbour of most agents come out of one
 * DRAM burst instead of two (skipped silently when a degree exceeds 255).  The neighbour rule is the same with and
 * without records: entry mulhi(r2, degree) of the sorted row.
 * The handle is accepted by sponet_cnvm_run / sponet_cntm_run in native mode (uniform degree_alpha, i.e. alpha = 1, or after
 * sponet_graph_batch_set_weights; no edge collective variables) when it holds num_states * (run_end - run_begin) graphs, and is
 * released with sponet_graph_destroy.  num_states * (run_end - run_begin) * num_agents must stay below 2^32.  A caller
 * that covers the states in blocks adds first_state * num_runs_total to run_begin / run_end (as it adds it to
 * sponet_native_opts.chain_offset of the run call). */
int sponet_graph_batch_create_er(int64_t num_agents, double p, int force_no_isolates, uint64_t seed,
                                 int64_t num_states, int64_t num_runs_total, int64_t run_begin, int64_t run_end,
                                 int32_t max_attempts, int layouts, sponet_graph** out);
#define SPONET_BATCH_PADDED_ROWS 1
#define SPONET_BATCH_RECORDS 2

/* The same for Barabasi-Albert graphs (sponet/network_generator.py:118-144 = networkx.barabasi_albert_graph(n, m)): star on the
 * vertices 0 .. m, then every vertex t > m attaches to m DISTINCT targets drawn uniformly from the endpoints of the edges so
 * far (duplicates are drawn again).  Edge e < m is (0, e + 1), edge e >= m the (e % m)-th edge of vertex e / m + m; endpoint
 * position 2 e is the edge's source, 2 e + 1 its target; candidate k of vertex t is position mulhi(word 0 of
 * Philox4x32-10(counter (k, t, c_lo, c_hi[16] | 5 << 24), key seed), 2 m (t - m)).  One warp per graph resolves 32 vertices at a time;
 * restated sequentially by oracle_ba_graph (bit-identical).  m <= 64. */
int sponet_graph_batch_create_ba(int64_t num_agents, int32_t m, uint64_t seed, int64_t num_states, int64_t num_runs_total,
                                 int64_t run_begin, int64_t run_end, int layouts, sponet_graph** out);

/* The same for Watts-Strogatz graphs (sponet/network_generator.py:147-183 = networkx.connected_watts_strogatz_graph(n, k, p,
 * tries)): ring lattice with k / 2 neighbours on each side; positions (j, u), j = 1 .. k / 2 outer, u inner: with probability p
 * the edge (u, u + j) is replaced by (u, w), w uniform and drawn again while w == u or {u, w} is an edge of the CURRENT graph; the
 * graph is drawn again (attempt + 1) until it is connected, SPONET_ERR_INVALID "Maximum number of tries exceeded" after `tries`.
 * Position q = (j - 1) n + u uses Philox4x32-10(counter (d, q, c_lo, c_hi[16] | attempt << 16 | 6 << 24), key seed): word 0 of
 * draw 0 against floor(p 2^32), w = mulhi(word 0 of draw d >= 1, n).  One warp per graph (the rewiring is sequential by
 * definition); restated by oracle_ws_graph (bit-identical).  A vertex may hold at most k + 32 neighbours. */
int sponet_graph_batch_create_ws(int64_t num_agents, int32_t k, double p, uint64_t seed, int64_t num_states, int64_t num_runs_total,
                                 int64_t run_begin, int64_t run_end, int32_t tries, int layouts, sponet_graph** out);

/* The same for random d-regular graphs (sponet/network_generator.py:87-115 = networkx.random_regular_graph(d, n)): the pairing
 * process with restarts.  Attempt a: n d stubs (position mod n); pass s: Fisher-Yates shuffle of the remaining stubs (i = L - 1
 * .. 1: swap with j = mulhi(word 0 of Philox4x32-10(counter (i, s | a << 16, c_lo, c_hi[16] | 7 << 24), key seed), i + 1)),
 * consecutive stubs paired in order, a loop or an existing edge leaves one potential at each of its vertices; no potential left:
 * done; every two distinct vertices with potentials already adjacent: the attempt failed, the next one starts; else the
 * potentials (in vertex order) are the stubs of the next pass.  One warp per graph; restated by oracle_regular_graph
 * (bit-identical). */
int sponet_graph_batch_create_regular(int64_t num_agents, int32_t d, uint64_t seed, int64_t num_states, int64_t num_runs_total,
                                      int64_t run_begin, int64_t run_end, int32_t tries, int layouts, sponet_graph** out);

/* Degree-weighted agent selection on a batch (alpha != 1): weight_of_degree[d] = d^(1 - alpha) for d = 0 .. max_degree, computed
 * by the caller with the arithmetic of sponet/cnvm/model.py:38 (`degrees ** (1 - alpha)`).  Builds, per graph, the alias table
 * of sponet/sampling.py:54-96 over its agents' weights (same statements as sponet_build_alias_table) and the sequential sum of
 * the weights; sponet_cnvm_run then takes agent weights, event rate and noise probability of every chain from its own graph
 * (sponet/cnvm/model.py:257-258) and ignores params->degree_alpha. */
int sponet_graph_batch_set_weights(sponet_graph* g, const double* weight_of_degree, int64_t num_weights);

/* Device memory of destroyed batches is kept for the next batch (cudaMalloc / cudaFree of tens of gigabytes cost up to
 * seconds; at most 64 GiB idle per process).  This returns it to the driver. */
int sponet_graph_batch_trim(void);

/* Graph `slot` of a batch as a host CSR (row_ptr [n+1]; col may be NULL to query row_ptr[n] first). */
int sponet_graph_batch_get(const sponet_graph* g, int64_t slot, int32_t* row_ptr, int32_t* col, int64_t col_capacity);

/* ---------------------------------------------------------------------------------------------
 * Model parameters as the reference loops receive them
 * ------------------------------------------------------------------------------------------- */
typedef struct {
    int32_t num_opinions;
    double r_imit, r_noise;     /* sponet/cnvm/parameters.py:293-338 convert_rate_to_cnvm          */
    const double* prob_imit;    /* [M*M] C order, HOST pointer                                      */
    const double* prob_noise;   /* [M*M] C order, HOST pointer                                      */
    const double* degree_alpha; /* d_i^(1-alpha), [n], HOST pointer (cnvm/model.py:30-42);          */
                                /* complete graph: only degree_alpha[0] is read (:382)              */
} sponet_cnvm_params;

typedef struct {
    double next_event_rate; /* 1 / (N (r + r_tilde))   sponet/cntm/model.py:32-34                   */
    double noise_prob;      /* r_tilde / (r + r_tilde) sponet/cntm/model.py:31                      */
    double threshold_01, threshold_10;
} sponet_cntm_params;

/* Work list of one replay worker: states [state_begin, state_end) x runs [run_begin, run_end),
 * iterated state-major exactly like _Worker.__call__ (sponet/multiprocessing.py:171-182), consuming
 * ONE sequential uint64 stream.  Outputs are indexed by the global (state, run). */
typedef struct {
    int64_t state_begin, state_end, run_begin, run_end;
    int64_t stream_offset, stream_len; /* slice of `streams` owned by this worker */
} sponet_replay_work;

/* ---------------------------------------------------------------------------------------------
 * REPLAY mode: bit-exact re-execution of the reference loops on an injected raw-uint64 stream
 * (BitGenerator.random_raw).  Uniforms are (u >> 11) * 2^-53, exponentials the 256-layer ziggurat
 * of numba/np/random/distributions.py:105-119.  One GPU warp per worker; lane 0 walks the chain.
 *
 *   x_init      [S, n]        uint8 (M <= 256) or uint16
 *   t_eval      [T]           float64, already validated (sponet/utils.py:173-190)
 *   out_t       [S, R, T]     event times as the reference returns them (t_traj)
 *   out_x       [S, R, T, n]  snapshots (x_traj), may be NULL
 *   stream_used [W]           uint64 values consumed per worker
 *   events      [W]           loop iterations per worker (may be NULL)
 * ------------------------------------------------------------------------------------------- */

/* sponet/cnvm/model.py:239-308 (_simulate_teval, g != NULL) and :365-433 (_simulate_complete_teval,
 * g == NULL). */
int sponet_cnvm_replay(const sponet_graph* g, int64_t num_agents, const sponet_cnvm_params* params,
                       const void* x_init, int64_t S, int64_t R, const double* t_eval, int64_t T,
                       const uint64_t* streams, const sponet_replay_work* work, int64_t W,
                       double* out_t, void* out_x, int64_t* stream_used, int64_t* events,
                       void* cuda_stream);

/* sponet/cntm/model.py:163-227 (_simulate_teval). */
int sponet_cntm_replay(const sponet_graph* g, const sponet_cntm_params* params, const uint8_t* x_init,
                       int64_t S, int64_t R, const double* t_eval, int64_t T, const uint64_t* streams,
                       const sponet_replay_work* work, int64_t W, double* out_t, uint8_t* out_x,
                       int64_t* stream_used, int64_t* events, void* cuda_stream);

/* Store-every-change variants (t_eval = None): sponet/cnvm/model.py:184-236, :311-362 and
 * sponet/cntm/model.py:111-160, for ONE realization, returned as an event log of accepted changes
 * (the reference copies all n agents per change, model.py:233).  log_* have capacity `cap`;
 * *num_changes receives the number of entries, *events the loop iterations. */
int sponet_cnvm_replay_all(const sponet_graph* g, int64_t num_agents, const sponet_cnvm_params* params,
                           const void* x_init, double t_max, const uint64_t* stream,
                           int64_t stream_len, int64_t cap, double* log_t, int32_t* log_agent,
                           int32_t* log_opinion, int64_t* num_changes, int64_t* stream_used,
                           int64_t* events, void* cuda_stream);
int sponet_cntm_replay_all(const sponet_graph* g, const sponet_cntm_params* params,
                           const uint8_t* x_init, double t_max, const uint64_t* stream,
                           int64_t stream_len, int64_t cap, double* log_t, int32_t* log_agent,
                           int32_t* log_opinion, int64_t* num_changes, int64_t* stream_used,
                           int64_t* events, void* cuda_stream);

/* ---------------------------------------------------------------------------------------------
 * NATIVE mode: the production ensemble path.  Chains (state j, run i) are independent; chain id
 * j * R_total + i keys a Philox4x32-10 stream, so results do not depend on how runs are sharded
 * over calls or GPUs.  This call executes runs [run_begin, run_end) of every one of the S initial
 * states.  One warp per chain, 32 candidate events per warp step, agent states bit-packed in shared
 * memory (global-memory fallback when a chain does not fit), collective variables reduced in-kernel.
 * DESIGN.md section 3 gives the sampling scheme; oracle/oracle_native.inc states it sequentially.
 * ------------------------------------------------------------------------------------------- */

/* In-kernel collective variable: the OpinionShares family (sponet/collective_variables.py:63-220).
 * The kernel keeps an integer table count[class * M + opinion] = sum of agent_weight over the agents
 * of that class holding that opinion, updated on every accepted event, and emits
 *     out[d] = count[idx[d]] / (divisors ? divisors[d] : divisor).
 *   OpinionShares(normalize, idx_to_return)       num_classes 1, no weights, divisor 1 or N
 *   OpinionShares(weights = integers) /
 *   DegreeWeightedOpinionShares                    agent_weight = weights / degrees, divisor sum|w|
 *   OpinionSharesByDegree                          agent_class = rank of the degree, divisors = class sizes
 * agent_class / agent_weight are served by the CNVM agent-level kernel only; float (non-integer)
 * weights are not in-kernel (use out_x + sponet_opinion_shares). */
typedef struct {
    int32_t dimension;           /* D                                                           */
    const int32_t* idx;          /* [D] HOST: indices into the table, class * M + opinion       */
    double divisor;
    int32_t num_classes;         /* 0 or 1: a single class                                      */
    const int32_t* agent_class;  /* [n] HOST, NULL = class 0                                    */
    const int32_t* agent_weight; /* [n] HOST integer weights (may be negative), NULL = 1        */
    const double* divisors;      /* [D] HOST per-output divisors, NULL = `divisor` for all      */
    /* Edge collective variables of a two-opinion model, evaluated at every output time by a scan of
     * `graph` against the chain state held on chip (no snapshot leaves the SM):
     *   SPONET_CV_INTERFACES    D = 1: #edges {u,v} with x[u] != x[v]       (collective_variables.py:259-305)
     *   SPONET_CV_PROPENSITIES  D = 2: out[m] = sum over agents j in state m of
     *                           r[m,n] * #{neighbours of j in state n} / degree_alpha[j] + r_tilde[m,n],
     *                           n = 1 - m                                    (collective_variables.py:308-384)
     * The result is divided by `divisor` (normalize: number of edges / number of agents).  idx,
     * agent_class, agent_weight and divisors are ignored for these kinds; out_sum / out_sumsq are
     * available for INTERFACES only (integer counts). */
    int32_t kind;                /* SPONET_CV_SHARES (0) | SPONET_CV_INTERFACES | SPONET_CV_PROPENSITIES */
    const sponet_graph* graph;   /* kinds 1, 2: the CV's own network (Interfaces(network) need not be the model's) */
    const double* degree_alpha;  /* kind 2: [n] HOST, d(j)^alpha (collective_variables.py:353-356)      */
    double rates[4];             /* kind 2: r[0,1], r[1,0], r_tilde[0,1], r_tilde[1,0]                   */
} sponet_shares_cv;

#define SPONET_CV_SHARES 0
#define SPONET_CV_INTERFACES 1
#define SPONET_CV_PROPENSITIES 2

typedef struct {
    uint64_t seed;
    int64_t num_runs_total; /* R_total: stride of the chain id                                      */
    int64_t run_begin, run_end;
    /* optional [S, run_end-run_begin, T-1] int64: candidate-event counts per output interval,
     * overriding the in-kernel Poisson(dt / next_event_rate) sampler (tests, host-side schedules) */
    const int64_t* event_counts;
    int32_t force_global_state; /* debug/tests: keep chain state in global memory                   */
    int32_t warps_per_chain;    /* 0 = auto, W >= 1: the W warps of a chain execute super-steps of 32 W events */
                                /* (CNVM agent-level kernel, shared-memory states only)             */
    /* ---- ABI 3 ---- */
    int32_t final_state_only;   /* out_x holds the state at the LAST output time only: [S, nrun, n] -- what a     */
                                /* caller needs to continue the chains (estimate_steady_state, steady_state.py:64) */
    /* In-kernel histogram of the collective variable over the runs, per initial state, output time and
     * component: out_hist[s, k, d, b] += 1 with b = clamp((c - hist_lo) / hist_width, 0, hist_bins - 1), c the RAW
     * integer value of component d (the count before division by the divisor).  Integer atomics like out_sum:
     * exact, order-free, the per-GPU partial of one all-reduce.  hist_bins = 0: none.  Not available for
     * SPONET_CV_PROPENSITIES (float valued). */
    int32_t hist_bins;
    int64_t hist_lo, hist_width;
    uint64_t* out_hist;         /* [S, T, D, hist_bins], ACCUMULATED into (caller zeroes), or NULL               */
    int64_t chain_offset;       /* added to every chain id: id = chain_offset + s * R_total + run.  A sweep sharded  */
                                /* over GPUs by points passes first_point * R_total, so that results do not depend  */
                                /* on the sharding                                                                  */
} sponet_native_opts;

/* counters: [0] candidate events executed, [1] accepted changes, [2] warp steps, [3] extra commit
 * rounds caused by intra-step hazards.  May be NULL. */
#define SPONET_NUM_COUNTERS 4

/* CNVM on a network (g != NULL; replaces cnvm/model.py:239-308 + utils.py:96-135 +
 * collective_variables.py:400-405 for a batch) or on the complete graph (g == NULL, :365-433).
 *   out_x   [S, nrun, T, n] uint8 snapshots ([S, nrun, n] with opts->final_state_only), or NULL
 *   out_cv  [S, nrun, T, D] float64, or NULL          (needs cv)
 *   out_sum, out_sumsq [S, T, D] int64 sums over runs of count and count^2 (raw counts of the
 *           selected opinions; ACCUMULATED into, caller zeroes) -- the per-GPU partials of
 *           sample_moments (sponet/sample_moments.py:106-108), or NULL                           */
int sponet_cnvm_run(const sponet_graph* g, int64_t num_agents, const sponet_cnvm_params* params,
                    const uint8_t* x_init, int64_t S, const double* t_eval, int64_t T,
                    const sponet_native_opts* opts, const sponet_shares_cv* cv, uint8_t* out_x,
                    double* out_cv, int64_t* out_sum, int64_t* out_sumsq, uint64_t* counters,
                    void* cuda_stream);

/* PARAMETER SWEEP (BASELINE config 5; the per-point loop a user writes around sponet/sample_moments.py:96-108):
 * P rate sets on ONE network in ONE launch.  Point p takes the place of the initial-state index everywhere:
 * chain id = p * R_total + run, out_cv [P, nrun, T, D], out_sum / out_sumsq / out_hist [P, T, D(, bins)],
 * out_x [P, nrun, T, n].  params[p] may differ in r_imit, r_noise, prob_imit, prob_noise (num_opinions and
 * degree_alpha must agree: one network, one alpha); every chain looks its acceptance table and noise
 * probability up by its point.  x_init is [P, n], or [1, n] shared by all points (x_init_shared != 0).
 * Point p of a sweep is bit-identical to sponet_cnvm_run with params[p] and chain_offset = p * R_total. */
int sponet_cnvm_run_sweep(const sponet_graph* g, int64_t num_agents, const sponet_cnvm_params* params,
                          int64_t P, const uint8_t* x_init, int32_t x_init_shared, const double* t_eval,
                          int64_t T, const sponet_native_opts* opts, const sponet_shares_cv* cv,
                          uint8_t* out_x, double* out_cv, int64_t* out_sum, int64_t* out_sumsq,
                          uint64_t* counters, void* cuda_stream);

/* Complete graph, count-only fast path: the chain state is the M opinion counts, kept in registers
 * (valid because agents are exchangeable on the complete graph and OpinionShares only reads counts).
 *   counts_init [S, M] int64 initial opinion counts. */
int sponet_cnvm_run_complete_counts(int64_t num_agents, const sponet_cnvm_params* params,
                                    const int64_t* counts_init, int64_t S, const double* t_eval,
                                    int64_t T, const sponet_native_opts* opts,
                                    const sponet_shares_cv* cv, double* out_cv, int64_t* out_sum,
                                    int64_t* out_sumsq, uint64_t* counters, void* cuda_stream);

/* CNTM (replaces cntm/model.py:163-227 for a batch); num_opinions is 2. */
int sponet_cntm_run(const sponet_graph* g, const sponet_cntm_params* params, const uint8_t* x_init,
                    int64_t S, const double* t_eval, int64_t T, const sponet_native_opts* opts,
                    const sponet_shares_cv* cv, uint8_t* out_x, double* out_cv, int64_t* out_sum,
                    int64_t* out_sumsq, uint64_t* counters, void* cuda_stream);

/* The interval schedule native mode draws: counts[c, k-1] ~ Poisson((t_eval[k]-t_eval[k-1]) /
 * next_event_rate) for chain ids chain_begin .. chain_begin+num_chains-1 (+1 on the first interval,
 * see DESIGN.md).  Exposed so tests can check it against the oracle. */
int sponet_native_event_counts(uint64_t seed, int64_t chain_begin, int64_t num_chains,
                               const double* t_eval, int64_t T, double next_event_rate,
                               int64_t* counts, void* cuda_stream);

/* ---------------------------------------------------------------------------------------------
 * Collective variables on stored states (the `cv(x)` call on user arrays)
 * ------------------------------------------------------------------------------------------- */

/* sponet/collective_variables.py:400-405 _opinion_shares_numba: out[s, m] = sum_j w[j] [x[s,j]==m]
 * (weights NULL = 1).  x is uint8 [S, n]; out float64 [S, M].  Integer-valued weights give exact
 * results; general weights are summed in a fixed tree order (tolerance in tests/test_gpu_api.py::test_collective_variables_known_answers). */
int sponet_opinion_shares(const uint8_t* x, int64_t S, int64_t num_agents, int32_t num_opinions,
                          const double* weights, double* out, void* cuda_stream);

/* sponet/collective_variables.py:259-305 Interfaces: number of edges whose endpoints disagree, per
 * snapshot (2-opinion states).  x uint8 [S, n]; out float64 [S]. */
int sponet_interfaces(const sponet_graph* g, const uint8_t* x, int64_t S, double* out, void* cuda_stream);

/* sponet/collective_variables.py:308-397 Propensities (2 opinions): out[s, m] = sum over agents j in
 * state m of r[m,n] * d(j,n) / degree_alpha[j] + r_tilde[m,n], n = 1 - m.  NOTE degree_alpha here is
 * d_j^alpha (collective_variables.py:355-358), not the d_j^(1-alpha) of the event loop.  g == NULL:
 * complete graph with the scalar complete_degree_alpha = (N-1)^alpha.  r, r_tilde: [2*2] HOST arrays.
 * x uint8 [S, n]; out float64 [S, 2]. */
int sponet_propensities(const sponet_graph* g, int64_t num_agents, const uint8_t* x, int64_t S,
                        const double* degree_alpha, double complete_degree_alpha, const double* r,
                        const double* r_tilde, double* out, void* cuda_stream);

/* sum and sum of squares over the run axis: x [R, L] float64 -> sum[L], sumsq[L]
 * (sponet/sample_moments.py:107-108), fixed-order reduction. */
int sponet_moments(const double* x, int64_t R, int64_t L, double* sum, double* sumsq, void* cuda_stream);

/* ---------------------------------------------------------------------------------------------
 * Initial-state samplers (sponet/states.py:45-164), native RNG (Philox keyed by `seed`)
 * ------------------------------------------------------------------------------------------- */

/* out uint8 [num_states, n] (host or device).
 *   counts == NULL: sponet/states.py:45-77 sample_states_uniform -- every agent uniform in {0..M-1}.
 *   counts [num_states, M] (or [1, M] with counts_shared): states.py:80-164 sample_states_uniform_shares /
 *   sample_states_target_shares -- state s is a uniformly random arrangement of counts[s][m] agents of opinion m
 *   (the reference shuffles np.repeat(opinions, counts); here: sequential draws without replacement).
 * The target counts themselves (counts_from_shares, Dirichlet shares) are host arithmetic of the caller. */
int sponet_sample_states(int64_t num_agents, int32_t num_opinions, const int64_t* counts, int32_t counts_shared,
                         int64_t num_states, uint64_t seed, uint8_t* out, void* cuda_stream);

#ifdef __cplusplus
}
#endif
#endif /* SPONET_B200_H */
