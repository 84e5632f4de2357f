// REPLAY mode kernels: bit-exact re-execution of the reference's numba loops on an injected
// raw-uint64 stream.  One warp per worker; lane 0 walks the (inherently sequential) chain, the
// whole warp copies initial states and snapshots.  This is the parity bridge, not the fast path:
// it exists so that agent-state trajectories and event times can be compared bit for bit with
// `model.simulate(..., rng=Generator(BitGenerator(seed)))`.
//
// Every float64 operation that the reference performs is issued with an explicit round-to-nearest
// intrinsic so that nvcc cannot contract a*b+c into an FMA (numba/LLVM does not contract either).
#include "common.cuh"

#define SPONET_ZIG_TABLE __constant__ const
#include "sponet_log1p.h"
#include "ziggurat_tables.h"

namespace {

struct RStream {
    const uint64_t* p;
    int64_t len, pos;
    bool dry;
    __device__ __forceinline__ uint64_t u64() {
        if (pos >= len) {
            dry = true;
            return 0;
        }
        return p[pos++];
    }
    // numba/np/random/generator_core.py:98-107 -> numpy next_double: (u >> 11) * 2^-53
    __device__ __forceinline__ double dbl() {
        return __dmul_rn((double)(u64() >> 11), 1.0 / 9007199254740992.0);
    }
    // numba/np/random/distributions.py:105-119 random_standard_exponential (256-layer ziggurat)
    __device__ double std_exponential() {
        for (;;) {
            uint64_t ri = u64();
            if (dry) return 0.0;
            ri >>= 3;
            const unsigned idx = (unsigned)(ri & 0xFF);
            ri >>= 8;
            const double x = __dmul_rn((double)ri, sponet_we_double[idx]);
            if (ri < sponet_ke_double[idx]) return x;
            if (idx == 0) return __dsub_rn(SPONET_ZIGGURAT_EXP_R, sponet_log1p(-dbl()));
            const double lhs = __dadd_rn(
                __dmul_rn(__dsub_rn(sponet_fe_double[idx - 1], sponet_fe_double[idx]), dbl()),
                sponet_fe_double[idx]);
            if (lhs < exp(-x)) return x;
            if (dry) return 0.0;
        }
    }
    // sponet/sampling.py:7-25 sample_randint: int(rng.random() * n)
    __device__ __forceinline__ int64_t randint(int64_t n) {
        return (int64_t)__dmul_rn(dbl(), (double)n);
    }
};

struct ReplayArgs {
    int64_t n, S, R, T, M;
    const double* t_eval;
    const uint64_t* streams;
    const sponet_replay_work* work;
    const void* x_init;
    void* scratch;  // [W, n] chain states
    double* out_t;
    void* out_x;
    int64_t* stream_used;
    int64_t* events;
    int* dry_flag;
    // graph
    const int32_t* row_ptr;
    const int32_t* col;
    // cnvm
    double next_event_rate, noise_probability;
    const double* prob_imit;
    const double* prob_noise;
    const double* prob_table;
    const int32_t* alias_table;
    // cntm
    double threshold_01, threshold_10;
};

enum { MODEL_CNVM_NET = 0, MODEL_CNVM_COMPLETE = 1, MODEL_CNTM = 2 };

// One loop iteration of the reference body (everything before `t += rng.exponential(...)`).
// Returns true when the state changed ("previous_*" must be recorded); for the CNTM noise branch
// the reference records unconditionally (cntm/model.py:191-195), which `force_record` reports.
template <typename XT, int MODEL>
__device__ __forceinline__ bool replay_event(const ReplayArgs& A, XT* x, RStream& s, int64_t& agent,
                                             XT& new_opinion) {
    const int64_t n = A.n, M = A.M;
    if (MODEL == MODEL_CNVM_NET) {
        if (s.dbl() < A.noise_probability) {  // cnvm/model.py:275-282
            agent = s.randint(n);
            const int64_t nw = s.randint(M);
            new_opinion = (XT)nw;
            return s.dbl() < A.prob_noise[(int64_t)x[agent] * M + nw];
        }
        // cnvm/model.py:283-292, sampling.py:117-122 (alias sample from ONE uniform)
        const double u = s.dbl();
        const int64_t i = (int64_t)__dmul_rn(u, (double)n);
        const double y = __dsub_rn(__dmul_rn((double)n, u), (double)i);
        agent = (y < A.prob_table[i]) ? i : (int64_t)A.alias_table[i];
        const int64_t beg = A.row_ptr[agent], deg = A.row_ptr[agent + 1] - beg;
        const int64_t nbr = A.col[beg + s.randint(deg)];
        const int64_t nw = x[nbr];
        new_opinion = (XT)nw;
        return s.dbl() < A.prob_imit[(int64_t)x[agent] * M + nw];
    } else if (MODEL == MODEL_CNVM_COMPLETE) {
        agent = s.randint(n);  // cnvm/model.py:400 -- agent first
        if (s.dbl() < A.noise_probability) {
            const int64_t nw = s.randint(M);
            new_opinion = (XT)nw;
            return s.dbl() < A.prob_noise[(int64_t)x[agent] * M + nw];
        }
        int64_t nbr = s.randint(n);  // cnvm/model.py:409-411 rejection
        while (nbr == agent && !s.dry) nbr = s.randint(n);
        const int64_t nw = x[nbr];
        new_opinion = (XT)nw;
        return s.dbl() < A.prob_imit[(int64_t)x[agent] * M + nw];
    } else {
        agent = s.randint(n);             // cntm/model.py:190
        if (s.dbl() < A.noise_probability) {  // :191-195 (noise_probability holds noise_prob)
            new_opinion = (XT)s.randint(2);
            return true;
        }
        const int64_t beg = A.row_ptr[agent], deg = A.row_ptr[agent + 1] - beg;
        if (deg > 0) {  // :198-211
            const XT other = (XT)(1 - x[agent]);
            int64_t count = 0;
            for (int64_t j = 0; j < deg; ++j) count += (x[A.col[beg + j]] == other);
            const double share = __ddiv_rn((double)count, (double)deg);
            const double threshold = (x[agent] == 1) ? A.threshold_10 : A.threshold_01;
            new_opinion = other;
            return share >= threshold;
        }
        return false;
    }
}

template <typename XT, int MODEL>
__global__ void __launch_bounds__(32) replay_teval_kernel(ReplayArgs A) {
    const int w = blockIdx.x, lane = threadIdx.x;
    const sponet_replay_work wk = A.work[w];
    const int64_t n = A.n, T = A.T;
    XT* x = (XT*)A.scratch + (size_t)w * n;
    const XT* x_init = (const XT*)A.x_init;
    XT* out_x = (XT*)A.out_x;
    RStream s{A.streams + wk.stream_offset, wk.stream_len, 0, false};
    int64_t iters = 0;
    bool dry = false;

    for (int64_t j = wk.state_begin; j < wk.state_end && !dry; ++j) {
        for (int64_t i = wk.run_begin; i < wk.run_end && !dry; ++i) {
            const int64_t run = j * A.R + i;
            // x = np.array(x_init) copy; x_traj = [x.copy()], t_traj = [t_eval[0]]  (model.py:105,262-264)
            for (int64_t a = lane; a < n; a += 32) {
                const XT v = x_init[j * n + a];
                x[a] = v;
                if (out_x) out_x[(run * T) * n + a] = v;
            }
            if (lane == 0) A.out_t[run * T] = A.t_eval[0];
            __syncwarp();

            double t = A.t_eval[0], previous_t = t;
            int64_t previous_agent = 0;
            XT previous_opinion = x[0];
            for (int64_t idx = 1; idx < T; ++idx) {
                int undo = 0;
                if (lane == 0) {
                    const double t_store = A.t_eval[idx];
                    for (;;) {  // model.py:274-306 until the next snapshot fires
                        ++iters;
                        int64_t agent;
                        XT nw;
                        if (replay_event<XT, MODEL>(A, x, s, agent, nw)) {
                            previous_t = t;
                            previous_agent = agent;
                            previous_opinion = x[agent];
                            x[agent] = nw;
                        }
                        t = __dadd_rn(t, __dmul_rn(A.next_event_rate, s.std_exponential()));
                        if (s.dry || t >= t_store) break;
                    }
                    // utils.py:128-135 store_snapshot_linspace
                    if (__dsub_rn(t, t_store) <= fabs(__dsub_rn(t_store, previous_t))) {
                        A.out_t[run * T + idx] = t;
                    } else {
                        A.out_t[run * T + idx] = previous_t;
                        undo = 1;
                    }
                }
                __syncwarp();
                dry = __shfl_sync(0xffffffffu, (int)s.dry, 0) != 0;
                if (dry) break;
                undo = __shfl_sync(0xffffffffu, undo, 0);
                const long long pa = __shfl_sync(0xffffffffu, (long long)previous_agent, 0);
                const int po = __shfl_sync(0xffffffffu, (int)previous_opinion, 0);
                if (out_x) {
                    XT* dst = out_x + (run * T + idx) * n;
                    for (int64_t a = lane; a < n; a += 32) {
                        XT v = x[a];
                        if (undo && a == pa) v = (XT)po;
                        dst[a] = v;
                    }
                }
                __syncwarp();
            }
        }
    }
    if (lane == 0) {
        A.stream_used[w] = s.pos;
        if (A.events) A.events[w] = iters;
        if (dry || s.dry) atomicExch(A.dry_flag, 1);
    }
}

struct ReplayAllArgs {
    ReplayArgs base;
    double t_max;
    int64_t cap;
    double* log_t;
    int32_t* log_agent;
    int32_t* log_opinion;
    int64_t* num_changes;
    int* cap_flag;
};

// t_eval = None loops: cnvm/model.py:184-236, :311-362; cntm/model.py:111-160 (single thread).
template <typename XT, int MODEL>
__global__ void replay_all_kernel(ReplayAllArgs B) {
    const ReplayArgs& A = B.base;
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    XT* x = (XT*)A.scratch;
    const XT* x_init = (const XT*)A.x_init;
    for (int64_t a = 0; a < A.n; ++a) x[a] = x_init[a];
    RStream s{A.streams, A.work[0].stream_len, 0, false};
    double t = 0.0;
    int64_t cnt = 0, iters = 0;
    bool overflow = false;
    while (t < B.t_max) {
        ++iters;
        t = __dadd_rn(t, __dmul_rn(A.next_event_rate, s.std_exponential()));  // exponential FIRST
        int64_t agent;
        XT nw;
        bool update;
        if (MODEL == MODEL_CNTM) {
            // cntm/model.py:132-154: agent, then `rng.random() < 0.5 * noise_prob` -> unconditional flip
            agent = s.randint(A.n);
            if (s.dbl() < __dmul_rn(0.5, A.noise_probability)) {
                nw = (XT)(1 - x[agent]);
                update = true;
            } else {
                update = false;
                const int64_t beg = A.row_ptr[agent], deg = A.row_ptr[agent + 1] - beg;
                if (deg > 0) {
                    const XT other = (XT)(1 - x[agent]);
                    int64_t count = 0;
                    for (int64_t j = 0; j < deg; ++j) count += (x[A.col[beg + j]] == other);
                    const double share = __ddiv_rn((double)count, (double)deg);
                    const double threshold = (x[agent] == 1) ? A.threshold_10 : A.threshold_01;
                    nw = other;
                    update = share >= threshold;
                }
            }
        } else {
            update = replay_event<XT, MODEL>(A, x, s, agent, nw);
        }
        if (s.dry) break;
        if (update) {
            x[agent] = nw;
            if (cnt >= B.cap) {
                overflow = true;
                break;
            }
            B.log_t[cnt] = t;
            B.log_agent[cnt] = (int32_t)agent;
            B.log_opinion[cnt] = (int32_t)nw;
            ++cnt;
        }
    }
    *B.num_changes = cnt;
    A.stream_used[0] = s.pos;
    if (A.events) A.events[0] = iters;
    if (s.dry) atomicExch(A.dry_flag, 1);
    if (overflow) atomicExch(B.cap_flag, 1);
}

// ---- host side -----------------------------------------------------------------------------

struct CnvmSetup {
    double next_event_rate, noise_probability;
    std::vector<double> prob_table;
    std::vector<int32_t> alias_table;
};

int check_work(const sponet_replay_work* work, int64_t W, int64_t S, int64_t R) {
    for (int64_t w = 0; w < W; ++w) {
        const sponet_replay_work& k = work[w];
        SP_REQUIRE(0 <= k.state_begin && k.state_begin <= k.state_end && k.state_end <= S,
                   "replay work %lld: state range [%lld,%lld) outside [0,%lld)", (long long)w,
                   (long long)k.state_begin, (long long)k.state_end, (long long)S);
        SP_REQUIRE(0 <= k.run_begin && k.run_begin <= k.run_end && k.run_end <= R,
                   "replay work %lld: run range [%lld,%lld) outside [0,%lld)", (long long)w,
                   (long long)k.run_begin, (long long)k.run_end, (long long)R);
        SP_REQUIRE(k.stream_offset >= 0 && k.stream_len >= 0, "replay work %lld: bad stream slice",
                   (long long)w);
    }
    return SPONET_OK;
}

template <typename XT>
int launch_teval(int model, const ReplayArgs& A, int64_t W, cudaStream_t st) {
    dim3 grid((unsigned)W), block(32);
    if (model == MODEL_CNVM_NET) replay_teval_kernel<XT, MODEL_CNVM_NET><<<grid, block, 0, st>>>(A);
    else if (model == MODEL_CNVM_COMPLETE) replay_teval_kernel<XT, MODEL_CNVM_COMPLETE><<<grid, block, 0, st>>>(A);
    else replay_teval_kernel<XT, MODEL_CNTM><<<grid, block, 0, st>>>(A);
    SP_CUDA(cudaGetLastError());
    return SPONET_OK;
}

template <typename XT>
int launch_all(int model, const ReplayAllArgs& B, cudaStream_t st) {
    if (model == MODEL_CNVM_NET) replay_all_kernel<XT, MODEL_CNVM_NET><<<1, 32, 0, st>>>(B);
    else if (model == MODEL_CNVM_COMPLETE) replay_all_kernel<XT, MODEL_CNVM_COMPLETE><<<1, 32, 0, st>>>(B);
    else replay_all_kernel<XT, MODEL_CNTM><<<1, 32, 0, st>>>(B);
    SP_CUDA(cudaGetLastError());
    return SPONET_OK;
}

// Fills the model part of ReplayArgs (rates, probability matrices, alias tables) on the device.
int setup_cnvm(SpScratch& sc, const sponet_graph* g, int64_t n, const sponet_cnvm_params* p,
               ReplayArgs& A) {
    SP_REQUIRE(p && p->prob_imit && p->prob_noise && p->degree_alpha, "cnvm params: null pointer");
    SP_REQUIRE(p->num_opinions >= 1 && p->num_opinions <= 65536, "num_opinions out of range");
    const int64_t M = p->num_opinions;
    A.M = M;
    if (g) {
        SP_REQUIRE(g->n == n, "graph has %lld agents, expected %lld", (long long)g->n, (long long)n);
        SP_REQUIRE(!g->pooled, "replay mode runs on a single graph, not on a graph batch");
        SP_REQUIRE(g->min_degree >= 1, "Isolated vertices in the network are not allowed.");
        SP_TRY(sponet_cnvm_event_rates(p->degree_alpha, n, 0.0, p->r_imit, p->r_noise,
                                       &A.next_event_rate, &A.noise_probability));
        std::vector<double> prob(n);
        std::vector<int32_t> alias(n);
        SP_TRY(sponet_build_alias_table(p->degree_alpha, n, prob.data(), alias.data()));
        double* d_prob;
        int32_t* d_alias;
        SP_CUDA(sc.alloc((void**)&d_prob, n * sizeof(double)));
        SP_CUDA(sc.alloc((void**)&d_alias, n * sizeof(int32_t)));
        SP_CUDA(cudaMemcpyAsync(d_prob, prob.data(), n * sizeof(double), cudaMemcpyHostToDevice, sc.stream));
        SP_CUDA(cudaMemcpyAsync(d_alias, alias.data(), n * sizeof(int32_t), cudaMemcpyHostToDevice, sc.stream));
        SP_CUDA(cudaStreamSynchronize(sc.stream));  // host vectors die at scope end
        A.prob_table = d_prob;
        A.alias_table = d_alias;
        A.row_ptr = g->d_row_ptr;
        A.col = g->d_col;
    } else {
        SP_TRY(sponet_cnvm_event_rates(nullptr, n, p->degree_alpha[0], p->r_imit, p->r_noise,
                                       &A.next_event_rate, &A.noise_probability));
    }
    double *d_pi, *d_pn;
    SP_CUDA(sc.alloc((void**)&d_pi, M * M * sizeof(double)));
    SP_CUDA(sc.alloc((void**)&d_pn, M * M * sizeof(double)));
    SP_CUDA(cudaMemcpyAsync(d_pi, p->prob_imit, M * M * sizeof(double), cudaMemcpyHostToDevice, sc.stream));
    SP_CUDA(cudaMemcpyAsync(d_pn, p->prob_noise, M * M * sizeof(double), cudaMemcpyHostToDevice, sc.stream));
    A.prob_imit = d_pi;
    A.prob_noise = d_pn;
    return SPONET_OK;
}

int finish_replay(SpScratch& sc, int* d_dry) {
    int dry = 0;
    SP_CUDA(cudaMemcpyAsync(&dry, d_dry, sizeof(int), cudaMemcpyDeviceToHost, sc.stream));
    SP_CUDA(sc.finish(true));
    if (dry)
        return sponet_fail(SPONET_ERR_STREAM_EXHAUSTED,
                           "replay: the injected uint64 stream was exhausted before the last output time");
    return SPONET_OK;
}

int replay_common(int model, SpScratch& sc, ReplayArgs& A, size_t xt_size, const void* x_init,
                  int64_t S, int64_t R, const double* t_eval, int64_t T, const uint64_t* streams,
                  const sponet_replay_work* work, int64_t W, double* out_t, void* out_x,
                  int64_t* stream_used, int64_t* events) {
    SP_REQUIRE(x_init && t_eval && streams && work && out_t && stream_used, "replay: null pointer");
    SP_REQUIRE(S >= 1 && R >= 0 && T >= 1 && W >= 1, "replay: bad sizes");
    SP_TRY(check_work(work, W, S, R));
    const int64_t n = A.n;
    A.S = S;
    A.R = R;
    A.T = T;
    int64_t stream_total = 0;
    for (int64_t w = 0; w < W; ++w)
        stream_total = std::max<int64_t>(stream_total, work[w].stream_offset + work[w].stream_len);
    const uint8_t* d_x;
    SP_CUDA(sc.in((const uint8_t*)x_init, (size_t)S * n * xt_size, &d_x));
    A.x_init = d_x;
    SP_CUDA(sc.in(t_eval, (size_t)T, &A.t_eval));
    SP_CUDA(sc.in(streams, (size_t)stream_total, &A.streams));
    SP_CUDA(sc.in(work, (size_t)W, &A.work));
    SP_CUDA(sc.alloc(&A.scratch, (size_t)W * n * xt_size));
    SP_CUDA(sc.out(out_t, (size_t)S * R * T, &A.out_t));
    uint8_t* d_out_x;
    SP_CUDA(sc.out((uint8_t*)out_x, (size_t)S * R * T * n * xt_size, &d_out_x));
    A.out_x = d_out_x;
    SP_CUDA(sc.out(stream_used, (size_t)W, &A.stream_used));
    SP_CUDA(sc.out(events, (size_t)W, &A.events));
    SP_CUDA(sc.alloc((void**)&A.dry_flag, sizeof(int)));
    SP_CUDA(cudaMemsetAsync(A.dry_flag, 0, sizeof(int), sc.stream));
    if (xt_size == 1) SP_TRY(launch_teval<uint8_t>(model, A, W, sc.stream));
    else SP_TRY(launch_teval<uint16_t>(model, A, W, sc.stream));
    return finish_replay(sc, A.dry_flag);
}

}  // namespace

extern "C" int sponet_cnvm_replay(const sponet_graph* g, int64_t num_agents,
                                  const sponet_cnvm_params* params, const void* x_init, int64_t S,
                                  int64_t R, const double* t_eval, int64_t T, const uint64_t* streams,
                                  const sponet_replay_work* work, int64_t W, double* out_t, void* out_x,
                                  int64_t* stream_used, int64_t* events, void* cuda_stream) {
    SP_REQUIRE(num_agents >= 1, "num_agents must be >= 1");
    SpScratch sc((cudaStream_t)cuda_stream);
    ReplayArgs A{};
    A.n = num_agents;
    SP_TRY(setup_cnvm(sc, g, num_agents, params, A));
    const size_t xt = params->num_opinions <= 256 ? 1 : 2;
    return replay_common(g ? MODEL_CNVM_NET : MODEL_CNVM_COMPLETE, sc, A, xt, x_init, S, R, t_eval, T,
                         streams, work, W, out_t, out_x, stream_used, events);
}

extern "C" int sponet_cntm_replay(const sponet_graph* g, const sponet_cntm_params* params,
                                  const uint8_t* x_init, int64_t S, int64_t R, const double* t_eval,
                                  int64_t T, const uint64_t* streams, const sponet_replay_work* work,
                                  int64_t W, double* out_t, uint8_t* out_x, int64_t* stream_used,
                                  int64_t* events, void* cuda_stream) {
    SP_REQUIRE(g && params, "cntm replay: graph and params are required");
    SP_REQUIRE(!g->pooled, "replay mode runs on a single graph, not on a graph batch");
    SpScratch sc((cudaStream_t)cuda_stream);
    ReplayArgs A{};
    A.n = g->n;
    A.M = 2;
    A.row_ptr = g->d_row_ptr;
    A.col = g->d_col;
    A.next_event_rate = params->next_event_rate;
    A.noise_probability = params->noise_prob;
    A.threshold_01 = params->threshold_01;
    A.threshold_10 = params->threshold_10;
    return replay_common(MODEL_CNTM, sc, A, 1, x_init, S, R, t_eval, T, streams, work, W, out_t, out_x,
                         stream_used, events);
}

namespace {
int replay_all_common(int model, SpScratch& sc, ReplayAllArgs& B, size_t xt_size, const void* x_init,
                      double t_max, const uint64_t* stream, int64_t stream_len, int64_t cap,
                      double* log_t, int32_t* log_agent, int32_t* log_opinion, int64_t* num_changes,
                      int64_t* stream_used, int64_t* events) {
    ReplayArgs& A = B.base;
    SP_REQUIRE(x_init && stream && log_t && log_agent && log_opinion && num_changes && stream_used,
               "replay_all: null pointer");
    SP_REQUIRE(cap >= 0 && stream_len >= 0, "replay_all: bad sizes");
    const int64_t n = A.n;
    sponet_replay_work wk{0, 1, 0, 1, 0, stream_len};
    sponet_replay_work* d_wk;
    SP_CUDA(sc.alloc((void**)&d_wk, sizeof(wk)));
    SP_CUDA(cudaMemcpyAsync(d_wk, &wk, sizeof(wk), cudaMemcpyHostToDevice, sc.stream));
    SP_CUDA(cudaStreamSynchronize(sc.stream));
    A.work = d_wk;
    const uint8_t* d_x;
    SP_CUDA(sc.in((const uint8_t*)x_init, (size_t)n * xt_size, &d_x));
    A.x_init = d_x;
    SP_CUDA(sc.in(stream, (size_t)stream_len, &A.streams));
    SP_CUDA(sc.alloc(&A.scratch, (size_t)n * xt_size));
    SP_CUDA(sc.out(stream_used, 1, &A.stream_used));
    SP_CUDA(sc.out(events, 1, &A.events));
    SP_CUDA(sc.alloc((void**)&A.dry_flag, 2 * sizeof(int)));
    SP_CUDA(cudaMemsetAsync(A.dry_flag, 0, 2 * sizeof(int), sc.stream));
    B.cap_flag = A.dry_flag + 1;
    B.t_max = t_max;
    B.cap = cap;
    SP_CUDA(sc.out(log_t, (size_t)cap, &B.log_t));
    SP_CUDA(sc.out(log_agent, (size_t)cap, &B.log_agent));
    SP_CUDA(sc.out(log_opinion, (size_t)cap, &B.log_opinion));
    SP_CUDA(sc.out(num_changes, 1, &B.num_changes));
    if (xt_size == 1) SP_TRY(launch_all<uint8_t>(model, B, sc.stream));
    else SP_TRY(launch_all<uint16_t>(model, B, sc.stream));
    int flags[2] = {0, 0};
    SP_CUDA(cudaMemcpyAsync(flags, A.dry_flag, sizeof(flags), cudaMemcpyDeviceToHost, sc.stream));
    SP_CUDA(sc.finish(true));
    if (flags[0])
        return sponet_fail(SPONET_ERR_STREAM_EXHAUSTED, "replay: the injected uint64 stream was exhausted before t_max");
    if (flags[1])
        return sponet_fail(SPONET_ERR_INVALID, "replay: event log capacity %lld too small", (long long)cap);
    return SPONET_OK;
}
}  // namespace

extern "C" int sponet_cnvm_replay_all(const sponet_graph* g, int64_t num_agents,
                                      const sponet_cnvm_params* params, const void* x_init,
                                      double t_max, const uint64_t* stream, int64_t stream_len,
                                      int64_t cap, double* log_t, int32_t* log_agent,
                                      int32_t* log_opinion, int64_t* num_changes,
                                      int64_t* stream_used, int64_t* events, void* cuda_stream) {
    SP_REQUIRE(num_agents >= 1, "num_agents must be >= 1");
    SpScratch sc((cudaStream_t)cuda_stream);
    ReplayAllArgs B{};
    B.base.n = num_agents;
    SP_TRY(setup_cnvm(sc, g, num_agents, params, B.base));
    const size_t xt = params->num_opinions <= 256 ? 1 : 2;
    return replay_all_common(g ? MODEL_CNVM_NET : MODEL_CNVM_COMPLETE, sc, B, xt, x_init, t_max, stream,
                             stream_len, cap, log_t, log_agent, log_opinion, num_changes, stream_used,
                             events);
}

extern "C" int sponet_cntm_replay_all(const sponet_graph* g, const sponet_cntm_params* params,
                                      const uint8_t* x_init, double t_max, const uint64_t* stream,
                                      int64_t stream_len, int64_t cap, double* log_t,
                                      int32_t* log_agent, int32_t* log_opinion, int64_t* num_changes,
                                      int64_t* stream_used, int64_t* events, void* cuda_stream) {
    SP_REQUIRE(g && params, "cntm replay: graph and params are required");
    SP_REQUIRE(!g->pooled, "replay mode runs on a single graph, not on a graph batch");
    SpScratch sc((cudaStream_t)cuda_stream);
    ReplayAllArgs B{};
    ReplayArgs& A = B.base;
    A.n = g->n;
    A.M = 2;
    A.row_ptr = g->d_row_ptr;
    A.col = g->d_col;
    A.next_event_rate = params->next_event_rate;
    A.noise_probability = params->noise_prob;
    A.threshold_01 = params->threshold_01;
    A.threshold_10 = params->threshold_10;
    return replay_all_common(MODEL_CNTM, sc, B, 1, x_init, t_max, stream, stream_len, cap, log_t,
                             log_agent, log_opinion, num_changes, stream_used, events);
}
