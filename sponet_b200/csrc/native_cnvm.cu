// NATIVE mode, CNVM: the production ensemble kernels.
//
// Design:
//   * one warp = one chain (realization).  The uniformised chain's random inputs come from a
//     counter-based Philox4x32-10 stream, so the 32 lanes evaluate 32 CONSECUTIVE candidate events
//     at once: RNG, alias lookup and the CSR gathers (row -> col -> neighbour id) never depend on
//     the chain state and run fully in parallel;
//   * the chain state (one opinion per agent) is bit-packed (1/2/4/8 bits) in SHARED memory --
//     25 KB for N = 1e5, M = 3, i.e. 9 chains per SM in the 227 KB a B200 CTA can address -- so the
//     state-dependent part (read x[a], x[nbr]; accept; write x[a]) costs shared-memory latency;
//   * sequential semantics are restored exactly: a lane has a hazard when an EARLIER lane of the
//     same step writes an agent it reads; lanes before the first hazard commit, the rest re-evaluate
//     against the updated state (expected < 1.02 rounds per step at N = 1e5);
//   * opinion counts (the OpinionShares CV) are maintained incrementally with warp ballots and
//     emitted at the output times; full trajectories never leave the SM unless asked for;
//   * sum / sum-of-squares over runs are accumulated with integer atomics (exact, order-free).
// Chains that do not fit in shared memory use the same code on a global-memory state (L2/HBM).
#include <algorithm>

#include "common.cuh"

namespace {

struct NativeCnvmArgs {
    int32_t n, M;
    const int2* row;  // (begin, degree); nullptr = complete graph
    const int32_t* col;
    uint32_t noise_thr;
    const uint2* alias_tab;  // (threshold, alias) per agent; nullptr = uniform agent weights
    const uint32_t* thr;  // [2*M*M]: imitation thresholds, then noise thresholds
    const uint8_t* x_init;
    const int64_t* counts;  // [S*nrun, T-1]
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
    uint32_t* gstate;  // global chain states (fallback), [total_warps, words]
    int32_t words;     // 32-bit words per chain
    int32_t warps_per_cta;
    int32_t bloom_log2;  // log2 of the per-warp Bloom filter size in bits (0 = disabled)
};

template <int BITS>
struct Pack {
    static constexpr int APW = 32 / BITS;  // agents per word
    static constexpr uint32_t VM = (BITS == 32) ? 0xFFFFFFFFu : ((1u << BITS) - 1u);
    __device__ static __forceinline__ uint32_t word(uint32_t a) { return a / APW; }
    __device__ static __forceinline__ uint32_t shift(uint32_t a) { return (a % APW) * BITS; }
};

template <bool SMEM>
__device__ __forceinline__ uint32_t state_load(const uint32_t* p) {
    if (SMEM) return *p;
    return __ldcg(p);  // global fallback: read at L2, lanes of this warp update it with atomics
}

// ---------------------------------------------------------------------------------------------
// agent-level kernel
//
// Per warp step (32 consecutive candidate events of one chain) the work is split into
//   stage 1  Philox draw -> branch, agent (alias table when alpha != 1); CSR row gather issued
//   stage 2  neighbour slot from the row; column gather issued
//   stage 3  dependency mask: which EARLIER lanes touch an agent this lane reads.  Agents and
//            neighbours never depend on the chain state, so this is computed before the commit:
//            match.any on the agents + a per-warp Bloom filter (shared memory) on the neighbours
//            prove "no overlap" for most steps; only when the filter fires the exact all-pairs
//            shuffle comparison runs
//   commit   read x[a], x[nbr]; accept; hazard = dep & ballot(accepting lanes); lanes before the
//            first hazard write (atomicXor on the packed word), the rest re-evaluate
// Stages 1 and 2 of the NEXT steps are issued before the commit of the current one (software
// pipeline, depth 2), so both L2 gathers are in flight while a step commits.
// ---------------------------------------------------------------------------------------------
struct StepDesc {
    uint64_t e0;  // index of the step's first event within the chain
    int nb;       // number of events in the step (lanes >= nb are idle)
    int snaps;    // output snapshots completed by this step
};

template <int BITS, bool SMEM, bool CSR>
__global__ void __launch_bounds__(512, 1) cnvm_native_kernel(NativeCnvmArgs A) {
    using P = Pack<BITS>;
    constexpr bool REGCNT = (BITS <= 2);    // M <= 4: opinion counts live in registers
    constexpr bool THR_SMEM = (BITS <= 4);  // M <= 16: threshold tables staged in shared memory
    constexpr uint32_t FULL = 0xffffffffu;
    extern __shared__ __align__(16) uint32_t smem[];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int M = A.M, MM = M * M;
    const uint32_t n = (uint32_t)A.n;
    const uint32_t noise_thr = A.noise_thr;
    const uint32_t lane_lt = (1u << lane) - 1u;

    // shared layout: [thr 2*MM (if THR_SMEM)] [counts W*M (if !REGCNT)] [bloom W*BW] [states W*words (if SMEM)]
    uint32_t* sp = smem;
    const uint32_t* thr = A.thr;
    if (THR_SMEM) {
        for (int i = threadIdx.x; i < 2 * MM; i += blockDim.x) sp[i] = A.thr[i];
        thr = sp;
        sp += 2 * MM;
    }
    int32_t* cnt_mem = nullptr;
    if (!REGCNT) {
        cnt_mem = (int32_t*)sp + warp * M;
        sp += A.warps_per_cta * M;
    }
    const int bloom_log2 = A.bloom_log2;  // log2(bits per warp); 0 = no filter (always exact)
    uint32_t* bloom = sp + (size_t)warp * (bloom_log2 ? (1u << (bloom_log2 - 5)) : 0u);
    if (bloom_log2) {
        for (int i = threadIdx.x; i < (A.warps_per_cta << (bloom_log2 - 5)); i += blockDim.x) sp[i] = 0u;
        sp += (size_t)A.warps_per_cta << (bloom_log2 - 5);
    }
    uint32_t* state;
    if (SMEM) state = sp + (size_t)warp * A.words;
    else state = A.gstate + ((size_t)blockIdx.x * A.warps_per_cta + warp) * A.words;
    __syncthreads();

    const int64_t total_warps = (int64_t)gridDim.x * A.warps_per_cta;
    const int64_t num_chains = A.S * A.nrun;
    unsigned long long ev_total = 0, acc_total = 0, step_total = 0, extra_rounds = 0;
    const int T = A.T;

    for (int64_t cl = (int64_t)blockIdx.x * A.warps_per_cta + warp; cl < num_chains; cl += total_warps) {
        const int64_t j = cl / A.nrun, i = cl - j * A.nrun;
        const uint64_t chain = (uint64_t)(j * A.R_total + A.run_begin + i);
        const uint8_t* x0 = A.x_init + j * (int64_t)n;
        const int64_t* kcounts = A.counts + cl * (int64_t)(T - 1);

        // ---- load + pack the initial state, count opinions ----
        int c0 = 0, c1 = 0, c2 = 0, c3 = 0;
        if (!REGCNT) {
            for (int v = lane; v < M; v += 32) cnt_mem[v] = 0;
            __syncwarp();
        }
        for (int w = lane; w < A.words; w += 32) {
            uint32_t packed = 0;
            const uint32_t base = (uint32_t)w * P::APW;
#pragma unroll 4
            for (int q = 0; q < P::APW; ++q) {
                const uint32_t a = base + q;
                if (a < n) {
                    const uint32_t v = x0[a];
                    packed |= v << (q * BITS);
                    if (REGCNT) {
                        c0 += (v == 0);
                        c1 += (v == 1);
                        c2 += (v == 2);
                        c3 += (v == 3);
                    } else {
                        atomicAdd(&cnt_mem[v], 1);
                    }
                }
            }
            state[w] = packed;
        }
        if (REGCNT) {
            c0 = __reduce_add_sync(FULL, c0);
            c1 = __reduce_add_sync(FULL, c1);
            c2 = __reduce_add_sync(FULL, c2);
            c3 = __reduce_add_sync(FULL, c3);
        }
        if (!SMEM) __threadfence_block();
        __syncwarp();

        // Per-lane packed deltas of the opinion counts (4 signed 8-bit fields, M <= 4), folded into
        // c0..c3 before any field can leave [-127, 127].
        uint32_t dpack = 0;
        int dsteps = 0;
        auto fold_counts = [&]() {
            if (REGCNT) {
                uint32_t t = dpack;
                const int d0 = (int)(int8_t)(t & 0xFF);
                t = (uint32_t)(((int)t - d0) >> 8);
                const int d1 = (int)(int8_t)(t & 0xFF);
                t = (uint32_t)(((int)t - d1) >> 8);
                const int d2 = (int)(int8_t)(t & 0xFF);
                t = (uint32_t)(((int)t - d2) >> 8);
                const int d3 = (int)(int8_t)(t & 0xFF);
                c0 += __reduce_add_sync(FULL, d0);
                c1 += __reduce_add_sync(FULL, d1);
                c2 += __reduce_add_sync(FULL, d2);
                c3 += __reduce_add_sync(FULL, d3);
                dpack = 0;
                dsteps = 0;
            }
        };

        int ksnap = 0;
        auto snapshot = [&]() {
            fold_counts();
            const int k = ksnap++;
            const int64_t snap = cl * (int64_t)T + k;
            if (A.out_cv || A.out_sum) {
                for (int d = lane; d < A.D; d += 32) {
                    const int v = A.cv_idx[d];
                    long long c;
                    if (REGCNT) c = (v == 0) ? c0 : (v == 1) ? c1 : (v == 2) ? c2 : c3;
                    else c = cnt_mem[v];
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
                    dst[a] = (uint8_t)((state_load<SMEM>(state + P::word(a)) >> P::shift(a)) & P::VM);
            }
        };

        // ---- step generator: walks the per-interval event counts, two steps ahead of the commit ----
        int gk = 1;
        int64_t gK = 0;
        uint64_t ge = 0;
        bool gdone = (T <= 1);
        snapshot();  // k = 0: the initial state
        if (!gdone) {
            gK = kcounts[0];
            while (gK == 0) {  // leading empty intervals repeat the initial state
                snapshot();
                if (++gk >= T) { gdone = true; break; }
                gK = kcounts[gk - 1];
            }
        }
        auto gen = [&](StepDesc& d) -> bool {
            d.e0 = ge;
            d.snaps = 0;
            if (gdone) { d.nb = 0; return false; }
            d.nb = gK < 32 ? (int)gK : 32;
            ge += (uint64_t)d.nb;
            gK -= d.nb;
            if (gK == 0) {
                d.snaps = 1;
                for (;;) {
                    if (++gk >= T) { gdone = true; break; }
                    gK = kcounts[gk - 1];
                    if (gK > 0) break;
                    ++d.snaps;
                }
            }
            return true;
        };

        // ---- pipeline registers ----
        // stage 1 (row gather in flight)
        uint32_t s1_a = 0, s1_r2 = 0, s1_r3 = 0, s1_meta = 0, s1_same = 0;
        int2 s1_row = make_int2(0, 1);
        // stage 2 (column gather in flight)
        uint32_t s2_a = 0, s2_r3 = 0, s2_meta = 0, s2_nbr = 0, s2_same = 0;
        StepDesc d1, d2;

        auto issue1 = [&](const StepDesc& d) {
            const uint4 r = sp_native_draw(A.seed, chain, 0u, d.e0 + (uint64_t)lane);
            const bool noise = r.x < noise_thr;
            const uint32_t idx = __umulhi(r.y, n);
            uint32_t a = idx;
            if (!noise && A.alias_tab) {  // sampling.py:117-122 with integer thresholds
                const uint2 at = __ldg(A.alias_tab + idx);
                if (!sp_thr_accept(r.y * n, at.x)) a = at.y;
            }
            s1_a = a;
            s1_r2 = r.z;
            s1_r3 = r.w;
            // MATCH.ANY is slow: it is issued here, two steps before its result is first looked at
            s1_same = __match_any_sync(FULL, a);
            s1_meta = (uint32_t)noise | (__umulhi(r.z, (uint32_t)M) << 8);  // bit 0 noise, bits 8.. noise opinion
            if (CSR) s1_row = __ldg(A.row + a);
        };
        auto issue2 = [&]() {
            s2_a = s1_a;
            s2_r3 = s1_r3;
            s2_meta = s1_meta;
            s2_same = s1_same;
            uint32_t nbr = s1_a;
            if (!(s1_meta & 1u)) {
                if (CSR) {
                    nbr = (uint32_t)__ldg(A.col + s1_row.x + __umulhi(s1_r2, (uint32_t)s1_row.y));
                } else {
                    nbr = __umulhi(s1_r2, n - 1);
                    nbr += (nbr >= s1_a);
                }
            }
            s2_nbr = nbr;
        };

        bool v1 = gen(d1);
        if (v1) issue1(d1);
        bool v2 = v1;
        d2 = d1;
        if (v2) issue2();
        v1 = gen(d1);
        if (v1) issue1(d1);

        while (v2) {
            // current step = stage 2 contents
            const uint32_t a = s2_a, nbr = s2_nbr, r3 = s2_r3, meta = s2_meta, same = s2_same;
            const StepDesc dc = d2;
            // advance the pipeline: next step's column gather, the one after's Philox + row gather
            v2 = v1;
            d2 = d1;
            if (v2) issue2();
            v1 = gen(d1);
            if (v1) issue1(d1);

            const bool noise = meta & 1u;
            const uint32_t nw_noise = meta >> 8;
            const int nb = dc.nb;

            // ---- stage 3: dependency mask of this step ----
            uint32_t dep = 0;
            bool suspect;
            {
                if (bloom_log2) {
                    const uint32_t ha = (a * 2654435761u) >> (32 - bloom_log2);
                    atomicOr(bloom + (ha >> 5), 1u << (ha & 31));
                    __syncwarp();
                    bool hit = same != (1u << lane);
                    if (!noise) {
                        const uint32_t hn = (nbr * 2654435761u) >> (32 - bloom_log2);
                        hit |= (bloom[hn >> 5] >> (hn & 31)) & 1u;
                    }
                    suspect = __any_sync(FULL, hit);
                    bloom[ha >> 5] = 0u;  // clear (same value from every lane sharing the word)
                } else {
                    suspect = true;
                }
                if (suspect) {
#pragma unroll
                    for (int jl = 0; jl < 32; ++jl) {
                        const uint32_t aj = __shfl_sync(FULL, a, jl);
                        dep |= (uint32_t)((aj == a) | (aj == nbr)) << jl;
                    }
                    dep &= lane_lt;
                }
            }

            // ---- commit ----
            const uint32_t wa = P::word(a), sa = P::shift(a);
            const uint32_t wn = P::word(nbr), sn = P::shift(nbr);
            const uint32_t tbase = noise ? (uint32_t)MM : 0u;
            bool pending = lane < nb;
            for (;;) {
                const uint32_t m = (state_load<SMEM>(state + wa) >> sa) & P::VM;
                const uint32_t nw = noise ? nw_noise : ((state_load<SMEM>(state + wn) >> sn) & P::VM);
                const uint32_t th = THR_SMEM ? thr[tbase + m * M + nw] : __ldg(thr + tbase + m * M + nw);
                const bool acc = pending && sp_thr_accept(r3, th);
                const uint32_t wmask = __ballot_sync(FULL, acc);
                const uint32_t hmask = suspect ? __ballot_sync(FULL, pending && (dep & wmask) != 0u) : 0u;
                const int first = hmask ? (__ffs(hmask) - 1) : 32;
                const bool commit = acc && lane < first;
                if (commit) {
                    atomicXor(state + wa, (m ^ nw) << sa);
                    if (REGCNT) {
                        dpack += (1u << (8 * nw)) - (1u << (8 * m));
                    } else {
                        atomicAdd(&cnt_mem[m], -1);
                        atomicAdd(&cnt_mem[nw], 1);
                    }
                }
                acc_total += __popc(hmask ? __ballot_sync(FULL, commit) : wmask);
                pending = pending && lane >= first;
                if (!SMEM) __threadfence_block();
                __syncwarp();
                if (first == 32) break;
                ++extra_rounds;
            }
            ++step_total;
            ev_total += nb;
            if (REGCNT && ++dsteps >= 120) fold_counts();
            for (int q = 0; q < dc.snaps; ++q) snapshot();
        }
        __syncwarp();
    }
    if (lane == 0 && A.counters) {
        atomicAdd(A.counters + 0, ev_total);
        atomicAdd(A.counters + 1, acc_total);
        atomicAdd(A.counters + 2, step_total);
        atomicAdd(A.counters + 3, extra_rounds);
    }
}

// ---------------------------------------------------------------------------------------------
// complete graph, count-only: one thread per chain, the M counts in registers
// ---------------------------------------------------------------------------------------------
struct CountsArgs {
    int32_t n, M;
    uint32_t noise_thr;
    const uint32_t* thr;
    const int64_t* counts_init;  // [S, M]
    const int64_t* counts;       // [S*nrun, T-1]
    int64_t S, nrun, R_total, run_begin;
    int32_t T;
    uint64_t seed;
    double* out_cv;
    unsigned long long* out_sum;
    unsigned long long* out_sumsq;
    int32_t D;
    const int32_t* cv_idx;
    double cv_div;
    unsigned long long* counters;
};

template <int MAXM>
__global__ void __launch_bounds__(128) cnvm_counts_kernel(CountsArgs A) {
    extern __shared__ __align__(16) uint32_t smem[];
    const int M = A.M, MM = M * M;
    for (int i = threadIdx.x; i < 2 * MM; i += blockDim.x) smem[i] = A.thr[i];
    int32_t* idx_s = (int32_t*)(smem + 2 * MM);
    for (int i = threadIdx.x; i < A.D; i += blockDim.x) idx_s[i] = A.cv_idx[i];
    __syncthreads();
    const uint32_t* thr = smem;
    const uint32_t n = (uint32_t)A.n;
    const int64_t cl = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long ev = 0, accn = 0;
    if (cl < A.S * A.nrun) {
        const int64_t j = cl / A.nrun, i = cl - j * A.nrun;
        const uint64_t chain = (uint64_t)(j * A.R_total + A.run_begin + i);
        int c[MAXM];
#pragma unroll
        for (int v = 0; v < MAXM; ++v) c[v] = v < M ? (int)A.counts_init[j * M + v] : 0;
        uint64_t e = 0;
        for (int k = 0; k < A.T; ++k) {
            if (k > 0) {
                const int64_t K = A.counts[cl * (int64_t)(A.T - 1) + (k - 1)];
                for (int64_t q = 0; q < K; ++q, ++e) {
                    const uint4 r = sp_native_draw(A.seed, chain, 0u, e);
                    const uint32_t idx = __umulhi(r.y, n);
                    // m = #{v : c_0 + .. + c_v <= idx}
                    uint32_t cum = 0;
                    int m = 0;
#pragma unroll
                    for (int v = 0; v < MAXM - 1; ++v) {
                        cum += (uint32_t)c[v];
                        m += (v < M - 1) & (idx >= cum);
                    }
                    int nw;
                    uint32_t th;
                    if (r.x < A.noise_thr) {
                        nw = (int)__umulhi(r.z, (uint32_t)M);
                        th = thr[MM + m * M + nw];
                    } else {
                        const uint32_t jn = __umulhi(r.z, n - 1);
                        cum = 0;
                        nw = 0;
#pragma unroll
                        for (int v = 0; v < MAXM - 1; ++v) {
                            cum += (uint32_t)(c[v] - (v == m));
                            nw += (v < M - 1) & (jn >= cum);
                        }
                        th = thr[m * M + nw];
                    }
                    if (sp_thr_accept(r.w, th)) {
#pragma unroll
                        for (int v = 0; v < MAXM; ++v) c[v] += (v == nw) - (v == m);
                        ++accn;
                    }
                }
                ev += (unsigned long long)K;
            }
            for (int d = 0; d < A.D; ++d) {
                const int v = idx_s[d];
                long long cv = 0;
#pragma unroll
                for (int u = 0; u < MAXM; ++u) cv = (u == v) ? c[u] : cv;
                if (A.out_cv)
                    A.out_cv[(cl * (int64_t)A.T + k) * A.D + d] =
                        (A.cv_div == 1.0) ? (double)cv : (double)cv / A.cv_div;
                if (A.out_sum) {
                    const int64_t o = (j * (int64_t)A.T + k) * A.D + d;
                    atomicAdd(A.out_sum + o, (unsigned long long)cv);
                    atomicAdd(A.out_sumsq + o, (unsigned long long)(cv * cv));
                }
            }
        }
    }
    if (A.counters) {
        ev = __reduce_add_sync(0xffffffffu, (unsigned)(ev & 0xffffffffu)) +
             ((unsigned long long)__reduce_add_sync(0xffffffffu, (unsigned)(ev >> 32)) << 32);
        accn = __reduce_add_sync(0xffffffffu, (unsigned)(accn & 0xffffffffu)) +
               ((unsigned long long)__reduce_add_sync(0xffffffffu, (unsigned)(accn >> 32)) << 32);
        if ((threadIdx.x & 31) == 0) {
            atomicAdd(A.counters + 0, ev);
            atomicAdd(A.counters + 1, accn);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// interval schedule: K ~ Poisson(dt / next_event_rate) per (chain, interval)
// Hoermann's PTRS transformed rejection for lam >= 10, product of uniforms below.
// ---------------------------------------------------------------------------------------------
__device__ int64_t native_poisson(uint64_t seed, uint64_t chain, uint64_t interval, double lam) {
    if (!(lam > 0.0)) return 0;
    if (lam < 10.0) {
        const double enlam = exp(-lam);
        double prod = 1.0;
        int64_t X = 0;
        for (uint64_t att = 0;; ++att) {
            const uint4 r = sp_native_draw(seed, chain, 1u, interval | (att << 32));
            prod *= sp_u53(r.x, r.y);
            if (prod > enlam) ++X;
            else return X;
        }
    }
    const double slam = sqrt(lam), loglam = log(lam);
    const double b = 0.931 + 2.53 * slam, a = -0.059 + 0.02483 * b;
    const double invalpha = 1.1239 + 1.1328 / (b - 3.4), vr = 0.9277 - 3.6224 / (b - 2.0);
    for (uint64_t att = 0;; ++att) {
        const uint4 r = sp_native_draw(seed, chain, 1u, interval | (att << 32));
        const double U = sp_u53(r.x, r.y) - 0.5, V = sp_u53(r.z, r.w);
        const double us = 0.5 - fabs(U);
        const double kf = floor((2.0 * a / us + b) * U + lam + 0.43);
        if (us >= 0.07 && V <= vr) return (int64_t)kf;
        if (kf < 0.0 || (us < 0.013 && V > us)) continue;
        if (log(V) + log(invalpha) - log(a / (us * us) + b) <= -lam + kf * loglam - lgamma(kf + 1.0))
            return (int64_t)kf;
    }
}

// chain id of local chain cl: (cl / nrun) * R_total + run_begin + cl % nrun; nrun == 0 -> chain_begin + cl
__global__ void event_counts_kernel(uint64_t seed, int64_t chain_begin, int64_t num_chains, int64_t nrun,
                                    int64_t R_total, int64_t run_begin, const double* t_eval, int T,
                                    double next_event_rate, int64_t* counts) {
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t per = T - 1;
    if (tid >= num_chains * per) return;
    const int64_t cl = tid / per;
    const int k = (int)(tid - cl * per) + 1;
    uint64_t chain;
    if (nrun > 0) {
        const int64_t j = cl / nrun, i = cl - j * nrun;
        chain = (uint64_t)(j * R_total + run_begin + i);
    } else {
        chain = (uint64_t)(chain_begin + cl);
    }
    const double lam = (t_eval[k] - t_eval[k - 1]) / next_event_rate;
    counts[tid] = native_poisson(seed, chain, (uint64_t)k, lam) + (k == 1 ? 1 : 0);
}

}  // namespace

// Shared with native_cntm.cu.
int sp_launch_event_counts(cudaStream_t st, uint64_t seed, int64_t chain_begin, int64_t num_chains,
                           int64_t nrun, int64_t R_total, int64_t run_begin, const double* d_t_eval,
                           int T, double next_event_rate, int64_t* d_counts) {
    const int64_t total = num_chains * (int64_t)(T - 1);
    if (total <= 0) return SPONET_OK;
    const int block = 128;
    const int64_t grid = (total + block - 1) / block;
    SP_REQUIRE(grid < ((int64_t)1 << 31), "too many (chain, interval) pairs for one launch");
    event_counts_kernel<<<(unsigned)grid, block, 0, st>>>(seed, chain_begin, num_chains, nrun, R_total,
                                                          run_begin, d_t_eval, T, next_event_rate,
                                                          d_counts);
    SP_CUDA(cudaGetLastError());
    return SPONET_OK;
}

int sp_validate_native_common(const double* t_eval, int64_t T, const sponet_native_opts* o, int64_t S) {
    SP_REQUIRE(o, "opts is NULL");
    SP_REQUIRE(t_eval && T >= 1 && T < ((int64_t)1 << 31), "t_eval/T invalid");
    SP_REQUIRE(S >= 1, "need at least one initial state");
    SP_REQUIRE(o->num_runs_total >= 0 && 0 <= o->run_begin && o->run_begin <= o->run_end &&
                   o->run_end <= o->num_runs_total,
               "run range [%lld,%lld) outside [0,%lld)", (long long)o->run_begin, (long long)o->run_end,
               (long long)o->num_runs_total);
    return SPONET_OK;
}

// Stages the in-kernel CV descriptor.  cv == NULL with moment outputs is an error.
int sp_stage_cv(SpScratch& sc, const sponet_shares_cv* cv, int M, SpCvDev* out) {
    out->D = 0;
    out->d_idx = nullptr;
    out->divisor = 1.0;
    if (!cv) return SPONET_OK;
    SP_REQUIRE(cv->dimension >= 1 && cv->idx, "cv: dimension/idx invalid");
    SP_REQUIRE(cv->divisor != 0.0, "cv: divisor must be non-zero");
    for (int d = 0; d < cv->dimension; ++d)
        SP_REQUIRE(cv->idx[d] >= 0 && cv->idx[d] < M, "cv: opinion index %d out of range", cv->idx[d]);
    int32_t* d_idx;
    SP_CUDA(sc.alloc((void**)&d_idx, cv->dimension * sizeof(int32_t)));
    SP_CUDA(cudaMemcpyAsync(d_idx, cv->idx, cv->dimension * sizeof(int32_t), cudaMemcpyHostToDevice, sc.stream));
    SP_CUDA(cudaStreamSynchronize(sc.stream));
    out->D = cv->dimension;
    out->d_idx = d_idx;
    out->divisor = cv->divisor;
    return SPONET_OK;
}

namespace {

template <int BITS, bool SMEM, bool CSR>
int launch_cnvm(const NativeCnvmArgs& A, int grid, int threads, size_t smem_bytes, cudaStream_t st) {
    auto kern = cnvm_native_kernel<BITS, SMEM, CSR>;
    SP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
    kern<<<grid, threads, smem_bytes, st>>>(A);
    SP_CUDA(cudaGetLastError());
    return SPONET_OK;
}

template <int BITS, bool SMEM>
int launch_cnvm_g(bool csr, const NativeCnvmArgs& A, int grid, int threads, size_t smem_bytes, cudaStream_t st) {
    return csr ? launch_cnvm<BITS, SMEM, true>(A, grid, threads, smem_bytes, st)
               : launch_cnvm<BITS, SMEM, false>(A, grid, threads, smem_bytes, st);
}

template <int BITS>
int launch_cnvm_s(bool smem, bool csr, const NativeCnvmArgs& A, int grid, int threads, size_t smem_bytes,
                  cudaStream_t st) {
    return smem ? launch_cnvm_g<BITS, true>(csr, A, grid, threads, smem_bytes, st)
                : launch_cnvm_g<BITS, false>(csr, A, grid, threads, smem_bytes, st);
}

template <int BITS, bool SMEM, bool CSR>
int occupancy_of(int threads, size_t smem_bytes, int* blocks) {
    auto kern = cnvm_native_kernel<BITS, SMEM, CSR>;
    SP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
    SP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks, kern, threads, smem_bytes));
    return SPONET_OK;
}

int occupancy_dispatch(int bits, bool smem, bool csr, int threads, size_t smem_bytes, int* blocks) {
#define OCC(B)                                                                                     \
    (smem ? (csr ? occupancy_of<B, true, true>(threads, smem_bytes, blocks)                        \
                 : occupancy_of<B, true, false>(threads, smem_bytes, blocks))                      \
          : (csr ? occupancy_of<B, false, true>(threads, smem_bytes, blocks)                       \
                 : occupancy_of<B, false, false>(threads, smem_bytes, blocks)))
    switch (bits) {
        case 1: return OCC(1);
        case 2: return OCC(2);
        case 4: return OCC(4);
        default: return OCC(8);
    }
#undef OCC
}

// Threshold tables: thr[0:MM] imitation, thr[MM:2MM] noise.
int stage_thresholds(SpScratch& sc, const sponet_cnvm_params* p, const uint32_t** d_thr) {
    const int64_t MM = (int64_t)p->num_opinions * p->num_opinions;
    std::vector<uint32_t> h(2 * MM);
    for (int64_t i = 0; i < MM; ++i) {
        SP_REQUIRE(p->prob_imit[i] >= 0.0 && p->prob_imit[i] <= 1.0 && p->prob_noise[i] >= 0.0 &&
                       p->prob_noise[i] <= 1.0,
                   "Probabilities have to be between 0 and 1.");
        h[i] = sp_prob_to_thr(p->prob_imit[i]);
        h[MM + i] = sp_prob_to_thr(p->prob_noise[i]);
    }
    uint32_t* d;
    SP_CUDA(sc.alloc((void**)&d, 2 * MM * sizeof(uint32_t)));
    SP_CUDA(cudaMemcpyAsync(d, h.data(), 2 * MM * sizeof(uint32_t), cudaMemcpyHostToDevice, sc.stream));
    SP_CUDA(cudaStreamSynchronize(sc.stream));
    *d_thr = d;
    return SPONET_OK;
}

}  // namespace

extern "C" int sponet_native_event_counts(uint64_t seed, int64_t chain_begin, int64_t num_chains,
                                          const double* t_eval, int64_t T, double next_event_rate,
                                          int64_t* counts, void* cuda_stream) {
    SP_REQUIRE(t_eval && counts && T >= 1 && num_chains >= 0 && next_event_rate > 0.0,
               "native_event_counts: bad arguments");
    SpScratch sc((cudaStream_t)cuda_stream);
    const double* d_t;
    int64_t* d_c;
    SP_CUDA(sc.in(t_eval, (size_t)T, &d_t));
    SP_CUDA(sc.out(counts, (size_t)(num_chains * (T - 1)), &d_c));
    SP_TRY(sp_launch_event_counts(sc.stream, seed, chain_begin, num_chains, 0, 0, 0, d_t, (int)T,
                                  next_event_rate, d_c));
    SP_CUDA(sc.finish());
    return SPONET_OK;
}

extern "C" int sponet_cnvm_run(const sponet_graph* g, int64_t num_agents, const sponet_cnvm_params* p,
                               const uint8_t* x_init, int64_t S, const double* t_eval, int64_t T,
                               const sponet_native_opts* o, const sponet_shares_cv* cv, uint8_t* out_x,
                               double* out_cv, int64_t* out_sum, int64_t* out_sumsq, uint64_t* counters,
                               void* cuda_stream) {
    SP_REQUIRE(p && p->prob_imit && p->prob_noise && p->degree_alpha && x_init, "cnvm_run: null pointer");
    SP_REQUIRE(num_agents >= 2 && num_agents < ((int64_t)1 << 31), "num_agents out of range");
    SP_TRY(sp_validate_native_common(t_eval, T, o, S));
    const int M = p->num_opinions;
    SP_REQUIRE(M >= 1, "num_opinions must be >= 1");
    if (M > 256)
        return sponet_fail(SPONET_ERR_UNSUPPORTED,
                           "native mode supports num_opinions <= 256 (got %d); use replay mode", M);
    SP_REQUIRE((out_cv || out_sum || out_sumsq) ? cv != nullptr : true, "out_cv/out_sum need a cv descriptor");
    SP_REQUIRE((out_sum == nullptr) == (out_sumsq == nullptr), "out_sum and out_sumsq go together");
    const int64_t n = num_agents, nrun = o->run_end - o->run_begin;
    if (g) {
        SP_REQUIRE(g->n == n, "graph has %lld agents, expected %lld", (long long)g->n, (long long)n);
        SP_REQUIRE(g->min_degree >= 1, "Isolated vertices in the network are not allowed.");
    }
    int dev = 0;
    SP_CUDA(cudaGetDevice(&dev));
    if (g) SP_REQUIRE(g->device == dev, "graph lives on device %d, current device is %d", g->device, dev);
    const int64_t num_chains = S * nrun;
    if (num_chains == 0) return SPONET_OK;

    SpScratch sc((cudaStream_t)cuda_stream);
    NativeCnvmArgs A{};
    A.n = (int32_t)n;
    A.M = M;
    double next_event_rate, noise_probability;
    bool uniform = true;
    if (g) {
        SP_TRY(sponet_cnvm_event_rates(p->degree_alpha, n, 0.0, p->r_imit, p->r_noise, &next_event_rate,
                                       &noise_probability));
        for (int64_t i = 1; i < n && uniform; ++i) uniform = (p->degree_alpha[i] == p->degree_alpha[0]);
        if (!uniform) {
            std::vector<double> prob(n);
            std::vector<int32_t> alias(n);
            std::vector<uint2> tab(n);
            SP_TRY(sponet_build_alias_table(p->degree_alpha, n, prob.data(), alias.data()));
            for (int64_t i = 0; i < n; ++i) tab[i] = make_uint2(sp_prob_to_thr(prob[i]), (uint32_t)alias[i]);
            uint2* d_tab;
            SP_CUDA(sc.alloc((void**)&d_tab, n * sizeof(uint2)));
            SP_CUDA(cudaMemcpyAsync(d_tab, tab.data(), n * sizeof(uint2), cudaMemcpyHostToDevice, sc.stream));
            SP_CUDA(cudaStreamSynchronize(sc.stream));
            A.alias_tab = d_tab;
        }
        A.row = g->d_row;
        A.col = g->d_col;
    } else {
        SP_TRY(sponet_cnvm_event_rates(nullptr, n, p->degree_alpha[0], p->r_imit, p->r_noise,
                                       &next_event_rate, &noise_probability));
    }
    A.noise_thr = sp_prob_to_thr(noise_probability);
    SP_TRY(stage_thresholds(sc, p, &A.thr));
    SpCvDev cvd;
    SP_TRY(sp_stage_cv(sc, cv, M, &cvd));
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
    // interval schedule
    if (T > 1) {
        if (o->event_counts) {
            SP_CUDA(sc.in(o->event_counts, (size_t)num_chains * (T - 1), &A.counts));
        } else {
            int64_t* d_c;
            SP_CUDA(sc.alloc((void**)&d_c, (size_t)num_chains * (T - 1) * sizeof(int64_t)));
            SP_TRY(sp_launch_event_counts(sc.stream, o->seed, 0, num_chains, nrun, o->num_runs_total,
                                          o->run_begin, d_t, (int)T, next_event_rate, d_c));
            A.counts = d_c;
        }
    }
    SP_CUDA(sc.out(out_x, (size_t)num_chains * T * n, &A.out_x));
    SP_CUDA(sc.out(out_cv, (size_t)num_chains * T * cvd.D, &A.out_cv));
    SP_CUDA(sc.out((unsigned long long*)out_sum, (size_t)S * T * cvd.D, &A.out_sum, true));
    SP_CUDA(sc.out((unsigned long long*)out_sumsq, (size_t)S * T * cvd.D, &A.out_sumsq, true));
    SP_CUDA(sc.out((unsigned long long*)counters, SPONET_NUM_COUNTERS, &A.counters, true));

    // ---- launch geometry ----
    const int bits = M <= 2 ? 1 : M <= 4 ? 2 : M <= 16 ? 4 : 8;
    A.words = (int32_t)((n * bits + 31) / 32);
    cudaDeviceProp prop;
    SP_CUDA(cudaGetDeviceProperties(&prop, dev));
    const size_t smem_max = prop.sharedMemPerBlockOptin;
    const size_t chain_bytes = (size_t)A.words * 4;
    const size_t fixed = (bits <= 4 ? (size_t)2 * M * M * 4 : 0);
    const size_t cnt_bytes = (bits >= 4) ? (size_t)M * 4 : 0;  // per-warp opinion counters
    const size_t per_warp = chain_bytes + cnt_bytes;
    int warps = 0;
    bool smem_state = !o->force_global_state;
    if (smem_state) {
        if (smem_max > fixed + 64) warps = (int)std::min<size_t>(16, (smem_max - fixed - 64) / per_warp);
        if (warps < 1) smem_state = false;
    }
    size_t smem_bytes;
    if (smem_state) {
        // no more warps than there are chains to run on this many SMs
        const int64_t need = (num_chains + prop.multiProcessorCount - 1) / prop.multiProcessorCount;
        warps = (int)std::max<int64_t>(1, std::min<int64_t>(warps, need));
        smem_bytes = fixed + per_warp * warps;
    } else {
        warps = 8;
        smem_bytes = fixed + cnt_bytes * warps;
    }
    // Per-warp Bloom filter for the dependency pre-check: the largest power of two (<= 64 Kbit) that the
    // shared memory left over by the chain states can hold; below 1 Kbit it is not worth having.
    {
        const size_t spare = smem_max > smem_bytes + 64 ? (smem_max - smem_bytes - 64) / warps : 0;
        int lg = 0;
        while (lg < 16 && ((size_t)1 << (lg + 1)) / 8 <= spare) ++lg;
        // no larger than ~N/4 bits (false-positive rate ~ 32*25/bits per step vs a true overlap rate of
        // ~1000/N), so that small chains keep their occupancy; only spare bytes are used, a chain is
        // never given up for a larger filter
        int want = 10;
        while (want < 16 && ((int64_t)1 << want) < n / 4) ++want;
        lg = std::min(lg, want);
        A.bloom_log2 = lg >= 10 ? lg : 0;
        if (A.bloom_log2) smem_bytes += ((size_t)1 << A.bloom_log2) / 8 * warps;
    }
    A.warps_per_cta = warps;
    const int threads = warps * 32;
    int blocks_per_sm = 1;
    SP_TRY(occupancy_dispatch(bits, smem_state, g != nullptr, threads, smem_bytes, &blocks_per_sm));
    SP_REQUIRE(blocks_per_sm >= 1, "cnvm_run: kernel does not fit on an SM (smem %zu B)", smem_bytes);
    const int64_t ctas_needed = (num_chains + warps - 1) / warps;
    const int grid = (int)std::min<int64_t>(ctas_needed, (int64_t)prop.multiProcessorCount * blocks_per_sm);
    if (!smem_state) {
        SP_CUDA(sc.alloc((void**)&A.gstate, (size_t)grid * warps * chain_bytes));
    }
    const bool csr = g != nullptr;
    switch (bits) {
        case 1: SP_TRY(launch_cnvm_s<1>(smem_state, csr, A, grid, threads, smem_bytes, sc.stream)); break;
        case 2: SP_TRY(launch_cnvm_s<2>(smem_state, csr, A, grid, threads, smem_bytes, sc.stream)); break;
        case 4: SP_TRY(launch_cnvm_s<4>(smem_state, csr, A, grid, threads, smem_bytes, sc.stream)); break;
        default: SP_TRY(launch_cnvm_s<8>(smem_state, csr, A, grid, threads, smem_bytes, sc.stream)); break;
    }
    SP_CUDA(sc.finish());
    return SPONET_OK;
}

extern "C" int sponet_cnvm_run_complete_counts(int64_t num_agents, const sponet_cnvm_params* p,
                                               const int64_t* counts_init, int64_t S,
                                               const double* t_eval, int64_t T,
                                               const sponet_native_opts* o, const sponet_shares_cv* cv,
                                               double* out_cv, int64_t* out_sum, int64_t* out_sumsq,
                                               uint64_t* counters, void* cuda_stream) {
    SP_REQUIRE(p && p->prob_imit && p->prob_noise && p->degree_alpha && counts_init && cv,
               "cnvm_run_complete_counts: null pointer");
    SP_REQUIRE(num_agents >= 2 && num_agents < ((int64_t)1 << 31), "num_agents out of range");
    SP_TRY(sp_validate_native_common(t_eval, T, o, S));
    const int M = p->num_opinions;
    if (M > 16)
        return sponet_fail(SPONET_ERR_UNSUPPORTED, "count-only path supports num_opinions <= 16 (got %d)", M);
    SP_REQUIRE((out_sum == nullptr) == (out_sumsq == nullptr), "out_sum and out_sumsq go together");
    const int64_t nrun = o->run_end - o->run_begin, num_chains = S * nrun;
    if (num_chains == 0) return SPONET_OK;
    if (!sp_is_device_ptr(counts_init))
        for (int64_t j = 0; j < S; ++j) {
            int64_t tot = 0;
            for (int v = 0; v < M; ++v) {
                SP_REQUIRE(counts_init[j * M + v] >= 0, "counts_init must be non-negative");
                tot += counts_init[j * M + v];
            }
            SP_REQUIRE(tot == num_agents, "counts_init row %lld sums to %lld, expected %lld", (long long)j,
                       (long long)tot, (long long)num_agents);
        }
    SpScratch sc((cudaStream_t)cuda_stream);
    CountsArgs A{};
    A.n = (int32_t)num_agents;
    A.M = M;
    double next_event_rate, noise_probability;
    SP_TRY(sponet_cnvm_event_rates(nullptr, num_agents, p->degree_alpha[0], p->r_imit, p->r_noise,
                                   &next_event_rate, &noise_probability));
    A.noise_thr = sp_prob_to_thr(noise_probability);
    SP_TRY(stage_thresholds(sc, p, &A.thr));
    SpCvDev cvd;
    SP_TRY(sp_stage_cv(sc, cv, M, &cvd));
    A.D = cvd.D;
    A.cv_idx = cvd.d_idx;
    A.cv_div = cvd.divisor;
    A.S = S;
    A.nrun = nrun;
    A.R_total = o->num_runs_total;
    A.run_begin = o->run_begin;
    A.T = (int32_t)T;
    A.seed = o->seed;
    SP_CUDA(sc.in(counts_init, (size_t)S * M, &A.counts_init));
    const double* d_t;
    SP_CUDA(sc.in(t_eval, (size_t)T, &d_t));
    if (T > 1) {
        if (o->event_counts) {
            SP_CUDA(sc.in(o->event_counts, (size_t)num_chains * (T - 1), &A.counts));
        } else {
            int64_t* d_c;
            SP_CUDA(sc.alloc((void**)&d_c, (size_t)num_chains * (T - 1) * sizeof(int64_t)));
            SP_TRY(sp_launch_event_counts(sc.stream, o->seed, 0, num_chains, nrun, o->num_runs_total,
                                          o->run_begin, d_t, (int)T, next_event_rate, d_c));
            A.counts = d_c;
        }
    }
    SP_CUDA(sc.out(out_cv, (size_t)num_chains * T * cvd.D, &A.out_cv));
    SP_CUDA(sc.out((unsigned long long*)out_sum, (size_t)S * T * cvd.D, &A.out_sum, true));
    SP_CUDA(sc.out((unsigned long long*)out_sumsq, (size_t)S * T * cvd.D, &A.out_sumsq, true));
    SP_CUDA(sc.out((unsigned long long*)counters, SPONET_NUM_COUNTERS, &A.counters, true));
    const int block = 128;
    const int64_t grid = (num_chains + block - 1) / block;
    SP_REQUIRE(grid < ((int64_t)1 << 31), "too many chains for one launch");
    const size_t smem_bytes = (size_t)2 * M * M * 4 + (size_t)cvd.D * 4;
    if (M <= 4) cnvm_counts_kernel<4><<<(unsigned)grid, block, smem_bytes, sc.stream>>>(A);
    else if (M <= 8) cnvm_counts_kernel<8><<<(unsigned)grid, block, smem_bytes, sc.stream>>>(A);
    else cnvm_counts_kernel<16><<<(unsigned)grid, block, smem_bytes, sc.stream>>>(A);
    SP_CUDA(cudaGetLastError());
    SP_CUDA(sc.finish());
    return SPONET_OK;
}
