// NATIVE mode, CNTM (noisy threshold model): one warp per chain, state as 1 bit per agent in shared
// memory.  Each warp step handles 32 consecutive candidate events:
//   phase 1 (parallel, state-independent): lane l draws event l (Philox), fetches its agent's CSR
//            row descriptor and turns both thresholds into integer "minimum opposite-neighbour
//            counts" with the reference's own float64 division (cntm/model.py:204-207);
//   phase 2 (parallel): the first 32 neighbour ids of all 32 rows are gathered, one coalesced
//            row segment per event, into registers -- 32 independent loads in flight per lane;
//   phase 3 (sequential over the 32 events, because every event may read any earlier flip): the
//            warp ballots the neighbour bits from shared memory, compares the popcount with the
//            precomputed minimum and flips the agent's bit.  Rows longer than 32 (BA hubs) stream
//            the rest of the row inline.
// The sequential statement this must reproduce bit for bit is oracle/oracle_native.inc
// (oracle_cntm_native).
#include <algorithm>

#include "common.cuh"

int sp_launch_event_counts(cudaStream_t st, uint64_t seed, int64_t chain_begin, int64_t num_chains,
                           int64_t nrun, int64_t R_total, int64_t run_begin, const double* d_t_eval,
                           int T, double next_event_rate, int64_t* d_counts);
int sp_validate_native_common(const double* t_eval, int64_t T, const sponet_native_opts* o, int64_t S);
int sp_stage_cv(SpScratch& sc, const sponet_shares_cv* cv, int M, SpCvDev* out);

namespace {

struct NativeCntmArgs {
    int32_t n;
    const int2* row;
    const int32_t* col;
    uint32_t noise_thr;
    double threshold_01, threshold_10;
    const uint8_t* x_init;
    const int64_t* counts;
    int64_t S, nrun, R_total, run_begin;
    int32_t T;
    uint64_t seed;
    uint8_t* out_x;
    double* out_cv;
    unsigned long long* out_sum;
    unsigned long long* out_sumsq;
    int32_t D;
    const int32_t* cv_idx;
    double cv_div;
    unsigned long long* counters;
    uint32_t* gstate;
    int32_t words;
    int32_t warps_per_cta;
};

// smallest c in [0, deg] with (double)c / (double)deg >= thr  (deg + 1 when there is none)
__device__ __forceinline__ int min_count(double thr, int deg) {
    if (deg <= 0) return 0x7fffffff;
    if (!(thr > 0.0)) return (0.0 >= thr) ? 0 : 0x7fffffff;  // thr <= 0 -> always; NaN -> never
    double g = ceil(thr * (double)deg);
    int c = g > (double)deg + 1.0 ? deg + 1 : (int)g;
    while (c > 0 && __ddiv_rn((double)(c - 1), (double)deg) >= thr) --c;
    while (c <= deg && !(__ddiv_rn((double)c, (double)deg) >= thr)) ++c;
    return c;
}

template <bool SMEM>
__device__ __forceinline__ uint32_t st_load(const uint32_t* p) {
    if (SMEM) return *p;
    return __ldcg(p);
}

template <bool SMEM>
__global__ void __launch_bounds__(512, 1) cntm_native_kernel(NativeCntmArgs A) {
    extern __shared__ __align__(16) uint32_t smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t n = (uint32_t)A.n;
    uint32_t* state = SMEM ? smem + (size_t)warp * A.words
                           : A.gstate + ((size_t)blockIdx.x * A.warps_per_cta + warp) * A.words;
    const int64_t total_warps = (int64_t)gridDim.x * A.warps_per_cta;
    const int64_t num_chains = A.S * A.nrun;
    const int T = A.T;
    unsigned long long ev_total = 0, acc_total = 0, step_total = 0;

    for (int64_t cl = (int64_t)blockIdx.x * A.warps_per_cta + warp; cl < num_chains; cl += total_warps) {
        const int64_t j = cl / A.nrun, i = cl - j * A.nrun;
        const uint64_t chain = (uint64_t)(j * A.R_total + A.run_begin + i);
        const uint8_t* x0 = A.x_init + j * (int64_t)n;
        int c1 = 0;
        for (int w = lane; w < A.words; w += 32) {
            uint32_t packed = 0;
            const uint32_t base = (uint32_t)w * 32;
#pragma unroll 8
            for (int q = 0; q < 32; ++q) {
                const uint32_t a = base + q;
                if (a < n) packed |= (uint32_t)(x0[a] & 1u) << q;
            }
            c1 += __popc(packed);
            state[w] = packed;
        }
        c1 = __reduce_add_sync(0xffffffffu, c1);
        if (!SMEM) __threadfence_block();
        __syncwarp();

        uint64_t e_base = 0;
        for (int k = 0; k < T; ++k) {
            if (k > 0) {
                int64_t K = A.counts[cl * (int64_t)(T - 1) + (k - 1)];
                while (K > 0) {
                    const int nb = K < 32 ? (int)K : 32;
                    // ---- phase 1 ----
                    const uint4 r = sp_native_draw(A.seed, chain, 0u, e_base + (uint64_t)lane);
                    const uint32_t my_a = __umulhi(r.y, n);
                    const bool my_noise = r.x < A.noise_thr;
                    const uint32_t my_bit = r.z >> 31;
                    int2 my_row = make_int2(0, 0);
                    int my_need01 = 0x7fffffff, my_need10 = 0x7fffffff;
                    if (!my_noise && lane < nb) {
                        my_row = __ldg(A.row + my_a);
                        my_need01 = min_count(A.threshold_01, my_row.y);
                        my_need10 = min_count(A.threshold_10, my_row.y);
                    }
                    // ---- phase 2: gather the first 32 neighbours of every event's row ----
                    uint32_t nbv[32];
#pragma unroll
                    for (int e = 0; e < 32; ++e) {
                        const int beg = __shfl_sync(0xffffffffu, my_row.x, e);
                        const int deg = __shfl_sync(0xffffffffu, my_row.y, e);
                        nbv[e] = (lane < deg) ? (uint32_t)__ldg(A.col + beg + lane) : 0xFFFFFFFFu;
                    }
                    // ---- phase 3: apply the events in order ----
                    int flips = 0;
#pragma unroll
                    for (int e = 0; e < 32; ++e) {
                        if (e < nb) {  // warp-uniform
                            const uint32_t a = __shfl_sync(0xffffffffu, my_a, e);
                            const uint32_t info = __shfl_sync(
                                0xffffffffu, (uint32_t)my_noise | (my_bit << 1), e);
                            const uint32_t wa = a >> 5, ba = a & 31;
                            const uint32_t xa = (st_load<SMEM>(state + wa) >> ba) & 1u;
                            uint32_t nw = xa;
                            if (info & 1u) {
                                nw = info >> 1;
                            } else {
                                const int deg = __shfl_sync(0xffffffffu, my_row.y, e);
                                if (deg > 0) {
                                    const uint32_t other = xa ^ 1u;
                                    const uint32_t v = nbv[e];
                                    bool hit = false;
                                    if (v != 0xFFFFFFFFu)
                                        hit = ((st_load<SMEM>(state + (v >> 5)) >> (v & 31)) & 1u) == other;
                                    int cnt = __popc(__ballot_sync(0xffffffffu, hit));
                                    if (deg > 32) {  // hub row: stream the remainder
                                        const int beg = __shfl_sync(0xffffffffu, my_row.x, e);
                                        for (int off = 32; off < deg; off += 32) {
                                            bool h2 = false;
                                            if (off + lane < deg) {
                                                const uint32_t u = (uint32_t)__ldg(A.col + beg + off + lane);
                                                h2 = ((st_load<SMEM>(state + (u >> 5)) >> (u & 31)) & 1u) == other;
                                            }
                                            cnt += __popc(__ballot_sync(0xffffffffu, h2));
                                        }
                                    }
                                    const int need = __shfl_sync(0xffffffffu, xa ? my_need10 : my_need01, e);
                                    if (cnt >= need) nw = other;
                                }
                            }
                            if (nw != xa) {  // warp-uniform
                                if (lane == 0) state[wa] = st_load<SMEM>(state + wa) ^ (1u << ba);
                                c1 += (int)nw - (int)xa;
                                ++flips;
                                if (!SMEM) __threadfence_block();
                            }
                            __syncwarp();
                        }
                    }
                    acc_total += flips;
                    ++step_total;
                    ev_total += nb;
                    e_base += nb;
                    K -= nb;
                }
            }
            // ---- snapshot k ----
            const int64_t snap = cl * (int64_t)T + k;
            if (A.out_cv || A.out_sum) {
                for (int d = lane; d < A.D; d += 32) {
                    const long long c = A.cv_idx[d] == 1 ? (long long)c1 : (long long)n - c1;
                    if (A.out_cv)
                        A.out_cv[snap * A.D + d] = (A.cv_div == 1.0) ? (double)c : (double)c / A.cv_div;
                    if (A.out_sum) {
                        const int64_t o = (j * (int64_t)T + k) * A.D + d;
                        atomicAdd(A.out_sum + o, (unsigned long long)c);
                        atomicAdd(A.out_sumsq + o, (unsigned long long)(c * c));
                    }
                }
            }
            if (A.out_x) {
                uint8_t* dst = A.out_x + snap * (int64_t)n;
                for (uint32_t a = lane; a < n; a += 32)
                    dst[a] = (uint8_t)((st_load<SMEM>(state + (a >> 5)) >> (a & 31)) & 1u);
            }
        }
        __syncwarp();
    }
    if (lane == 0 && A.counters) {
        atomicAdd(A.counters + 0, ev_total);
        atomicAdd(A.counters + 1, acc_total);
        atomicAdd(A.counters + 2, step_total);
    }
}

template <bool SMEM>
int launch(const NativeCntmArgs& A, int grid, int threads, size_t smem_bytes, cudaStream_t st) {
    auto kern = cntm_native_kernel<SMEM>;
    SP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
    kern<<<grid, threads, smem_bytes, st>>>(A);
    SP_CUDA(cudaGetLastError());
    return SPONET_OK;
}

template <bool SMEM>
int occupancy(int threads, size_t smem_bytes, int* blocks) {
    auto kern = cntm_native_kernel<SMEM>;
    SP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
    SP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks, kern, threads, smem_bytes));
    return SPONET_OK;
}

}  // namespace

extern "C" int sponet_cntm_run(const sponet_graph* g, const sponet_cntm_params* p, const uint8_t* x_init,
                               int64_t S, const double* t_eval, int64_t T, const sponet_native_opts* o,
                               const sponet_shares_cv* cv, uint8_t* out_x, double* out_cv,
                               int64_t* out_sum, int64_t* out_sumsq, uint64_t* counters,
                               void* cuda_stream) {
    SP_REQUIRE(g && p && x_init, "cntm_run: null pointer");
    SP_TRY(sp_validate_native_common(t_eval, T, o, S));
    SP_REQUIRE(p->next_event_rate > 0.0, "cntm_run: next_event_rate must be positive");
    SP_REQUIRE((out_cv || out_sum || out_sumsq) ? cv != nullptr : true, "out_cv/out_sum need a cv descriptor");
    SP_REQUIRE((out_sum == nullptr) == (out_sumsq == nullptr), "out_sum and out_sumsq go together");
    int dev = 0;
    SP_CUDA(cudaGetDevice(&dev));
    SP_REQUIRE(g->device == dev, "graph lives on device %d, current device is %d", g->device, dev);
    const int64_t n = g->n, nrun = o->run_end - o->run_begin, num_chains = S * nrun;
    if (num_chains == 0) return SPONET_OK;

    SpScratch sc((cudaStream_t)cuda_stream);
    NativeCntmArgs A{};
    A.n = (int32_t)n;
    A.row = g->d_row;
    A.col = g->d_col;
    A.noise_thr = sp_prob_to_thr(p->noise_prob);
    A.threshold_01 = p->threshold_01;
    A.threshold_10 = p->threshold_10;
    SpCvDev cvd;
    SP_TRY(sp_stage_cv(sc, cv, 2, &cvd));
    A.D = cvd.D;
    A.cv_idx = cvd.d_idx;
    A.cv_div = cvd.divisor;
    A.S = S;
    A.nrun = nrun;
    A.R_total = o->num_runs_total;
    A.run_begin = o->run_begin;
    A.T = (int32_t)T;
    A.seed = o->seed;
    SP_CUDA(sc.in(x_init, (size_t)S * n, &A.x_init));
    const double* d_t;
    SP_CUDA(sc.in(t_eval, (size_t)T, &d_t));
    if (T > 1) {
        if (o->event_counts) {
            SP_CUDA(sc.in(o->event_counts, (size_t)num_chains * (T - 1), &A.counts));
        } else {
            int64_t* d_c;
            SP_CUDA(sc.alloc((void**)&d_c, (size_t)num_chains * (T - 1) * sizeof(int64_t)));
            SP_TRY(sp_launch_event_counts(sc.stream, o->seed, 0, num_chains, nrun, o->num_runs_total,
                                          o->run_begin, d_t, (int)T, p->next_event_rate, d_c));
            A.counts = d_c;
        }
    }
    SP_CUDA(sc.out(out_x, (size_t)num_chains * T * n, &A.out_x));
    SP_CUDA(sc.out(out_cv, (size_t)num_chains * T * cvd.D, &A.out_cv));
    SP_CUDA(sc.out((unsigned long long*)out_sum, (size_t)S * T * cvd.D, &A.out_sum, true));
    SP_CUDA(sc.out((unsigned long long*)out_sumsq, (size_t)S * T * cvd.D, &A.out_sumsq, true));
    SP_CUDA(sc.out((unsigned long long*)counters, SPONET_NUM_COUNTERS, &A.counters, true));

    A.words = (int32_t)((n + 31) / 32);
    cudaDeviceProp prop;
    SP_CUDA(cudaGetDeviceProperties(&prop, dev));
    const size_t chain_bytes = (size_t)A.words * 4;
    bool smem_state = !o->force_global_state;
    int warps = 0;
    if (smem_state) {
        warps = (int)std::min<size_t>(16, (prop.sharedMemPerBlockOptin - 64) / chain_bytes);
        if (warps < 1) smem_state = false;
    }
    size_t smem_bytes = 0;
    if (smem_state) {
        const int64_t need = (num_chains + prop.multiProcessorCount - 1) / prop.multiProcessorCount;
        warps = (int)std::max<int64_t>(1, std::min<int64_t>(warps, need));
        smem_bytes = chain_bytes * warps;
    } else {
        warps = 8;
    }
    A.warps_per_cta = warps;
    const int threads = warps * 32;
    int blocks_per_sm = 1;
    SP_TRY(smem_state ? occupancy<true>(threads, smem_bytes, &blocks_per_sm)
                      : occupancy<false>(threads, smem_bytes, &blocks_per_sm));
    SP_REQUIRE(blocks_per_sm >= 1, "cntm_run: kernel does not fit on an SM");
    const int64_t ctas_needed = (num_chains + warps - 1) / warps;
    const int grid = (int)std::min<int64_t>(ctas_needed, (int64_t)prop.multiProcessorCount * blocks_per_sm);
    if (!smem_state) SP_CUDA(sc.alloc((void**)&A.gstate, (size_t)grid * warps * chain_bytes));
    SP_TRY(smem_state ? launch<true>(A, grid, threads, smem_bytes, sc.stream)
                      : launch<false>(A, grid, threads, smem_bytes, sc.stream));
    SP_CUDA(sc.finish());
    return SPONET_OK;
}
