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
//     same step writes an agent it reads.  The commit is optimistic: all lanes evaluate and write, then
//     re-read what they read; a changed input rolls the step back, the exact prefix commits and the rest
//     re-evaluates against the updated state (expected < 1.01 rounds per step at N = 1e5);
//   * opinion counts (the OpinionShares CV) are maintained incrementally with warp ballots and
//     emitted at the output times; full trajectories never leave the SM unless asked for;
//   * sum / sum-of-squares over runs are accumulated with integer atomics (exact, order-free).
// Chains that do not fit in shared memory use the same code on a global-memory state (L2/HBM).
#include <algorithm>
#include <type_traits>

#include "common.cuh"

namespace {

struct NativeCnvmArgs {
    int32_t n, M;
    const int2* row;  // (begin, degree); nullptr = complete graph
    const int32_t* col;
    uint32_t graph_stride;  // n for a BATCH of graphs (chain cl of the call runs on graph cl: rows [cl n, (cl+1) n)), else 0
    const uint32_t* rec;    // batch: 64-byte agent records (sponet_graph::d_rec) or nullptr
    uint32_t rec_deg_bits, rec_deg_mask;  // header = degree | row offset << rec_deg_bits
    const uint2* nbtab;  // neighbour sampler table (GRAPH == 2): [n << nbtab_log2]
    int32_t nbtab_log2;
    const uint32_t* noise_thr;  // [P] noise-branch threshold of the chain's parameter set
    const uint32_t* chain_noise_thr;  // [S * nrun] per chain (a batch with weighted agents: the rate depends on the graph), or nullptr
    const uint2* alias_tab;  // (threshold, alias) per agent; nullptr = uniform agent weights
    const uint32_t* thr;  // [P][2*M*M]: imitation thresholds, then noise thresholds
    int32_t param_stride;  // 0: one parameter set for all initial states; 1: parameter set j for state slot j (sweep)
    const uint32_t* x_packed;  // [Sx, words] initial states bit-packed once per call (pack_states_kernel)
    const int32_t* cnt_init;   // [Sx, cnt_stride] their opinion (x class) counts
    int64_t x_stride;          // 1, or 0 when all state slots share one initial state (sweep): index into the two above
    const int64_t* counts;  // [S*nrun, T-1]
    int64_t S, nrun, R_total, run_begin;
    int64_t cl_begin, cl_end;  // local chains [cl_begin, cl_end) of the S * nrun of the call run in this launch
    int32_t T;
    uint64_t seed;
    uint8_t* out_x;
    double* out_cv;
    unsigned long long* out_sum;
    unsigned long long* out_sumsq;
    unsigned long long* out_hist;  // [S, T, D, hist_bins] or nullptr
    int32_t hist_bins;
    int64_t hist_lo, hist_width;
    int32_t final_only;  // out_x holds the last output time only
    int32_t D;
    const int32_t* cv_idx;
    double cv_div;
    const double* cv_divs;  // per-output divisors (nullptr = cv_div)
    const int2* cw;         // per-agent (class, weight) of a grouped / weighted CV (nullptr = plain counts)
    SpEdgeCv ecv;           // Interfaces / Propensities (kind != 0): replaces the table outputs
    int32_t cnt_stride;     // counters per chain: num_classes * M
    unsigned long long* counters;
    uint32_t* gstate;  // global chain states (fallback), [total_warps, words]
    int32_t words;     // 32-bit words per chain
    int32_t warps_per_cta;
    int32_t warps_per_chain;  // RING kernels: warps serving one chain
    uint32_t rk[20];      // Philox4x32-10 round keys of the seed: (k0 + r*W0, k1 + r*W1), r = 0..9
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
//   stage 1  Philox draw -> branch, agent (alias table when alpha != 1); gather of the agent's sampler-table
//            entry (or CSR row) issued
//   stage 2  neighbour out of the gathered entry (or the column gather issued)
//   commit   write phase: accepting lanes swap their change into the packed state (compare-and-swap on the
//            agent's field); read phase: every lane re-reads its two locations -- an input changed by another
//            lane, or a lost write, rolls the step back, the exact prefix is committed and the rest re-evaluates
//            -- and reads the locations and the threshold of its NEXT step in the same pass
// Stage 1 runs two own steps ahead, stage 2 one (software pipeline, unrolled three ways so that the stage
// registers rotate without copies).
//
// COOP mode (template flag RING; large chains: only a few fit in an SM's shared memory): W warps serve one chain.
// Step t of an output interval belongs to warp t mod W; the W steps u W .. u W + W - 1 form a SUPER-STEP whose
// write and read phases all W warps execute together, separated by named barriers of the chain's 32 W threads
// (bar.sync after the writes, bar.red.or as the vote on the verification).  Nothing is serialized between the
// warps of a chain: 8 chains x 2 warps on workload H, one chain x 16 warps at N = 1e6.
// (Round 1 ordered the warps' commits through a ring of mbarriers; the hand-over latency bounded a chain's rate.)
// ---------------------------------------------------------------------------------------------
// Threads per CTA.  512 leave 128 registers per thread; with 768 the kernel compiles to 80 registers without
// spills in the event loop, but 24 warps per SM measured no faster than 16 (the ALU pipe's two-cycle issue, not
// latency, is the ceiling: profiles/r2_cnvm_h_ncu.md).
#ifndef SP_RING_MAX_THREADS
#define SP_RING_MAX_THREADS 512
#endif

__device__ __forceinline__ uint2 sp_lds2(uint32_t saddr) {  // 8-byte variant
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(saddr) : "memory");
    return v;
}
__device__ __forceinline__ void sp_ring_sync(int id, int threads) {  // named barrier of one chain's warps
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
// the same barrier with a vote: true iff `pred` holds for any of the participating threads
__device__ __forceinline__ bool sp_ring_any(int id, int threads, bool pred) {
    uint32_t r;
    asm volatile(
        "{\n"
        ".reg .pred p, q;\n"
        "setp.ne.u32 q, %3, 0;\n"
        "bar.red.or.pred p, %1, %2, q;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(r)
        : "r"(id), "r"(threads), "r"((uint32_t)pred)
        : "memory");
    return r != 0u;
}

// Bit-packs the initial states once per call and counts their opinions (per class and with the integer weights of a
// grouped CV): a chain then starts with a copy of `words` 32-bit words instead of reading N bytes -- at N = 1e5 a
// tenth of the traffic, which matters when chains are short (parameter sweeps, steady-state rounds: ~1e5 events per
// chain against 1e5 bytes of initial state).  Values >= M (device-resident inputs are not range-checked on the host)
// are clamped so that they can never spill into a neighbour's bits or another chain's counters.
template <int BITS>
__global__ void pack_states_kernel(const uint8_t* x, int64_t num_states, int64_t n, int M, int32_t words, const int2* cw,
                                   int32_t cnt_stride, uint32_t* packed, int32_t* cnt) {
    using P = Pack<BITS>;
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= words) return;
    for (int64_t j = blockIdx.y; j < num_states; j += gridDim.y) {
    const uint8_t* x0 = x + j * n;
    uint32_t out = 0;
    const uint32_t base = (uint32_t)w * P::APW;
    for (int q = 0; q < P::APW; ++q) {
        const uint32_t a = base + q;
        if (a < (uint32_t)n) {
            const uint32_t v = min((uint32_t)x0[a], (uint32_t)(M - 1));
            out |= v << (q * BITS);
            if (cw) {
                const int2 c = cw[a];
                atomicAdd(cnt + j * cnt_stride + c.x * M + v, c.y);
            } else {
                atomicAdd(cnt + j * cnt_stride + v, 1);
            }
        }
    }
    packed[j * words + w] = out;
    }
}

// GRAPH: 0 = complete graph (no gather), 1 = CSR (row gather, then column gather), 2 = neighbour sampler
// table (one gather; bounded-degree graphs whose table fits L2)
// MAXT: threads per CTA the kernel is compiled for.  512 (16 warps, up to 128 registers) is the default; the COOP
// kernels also exist for 576 threads (18 warps, 96 registers: the register file is per scheduler, a fifth warp on a
// scheduler leaves 102 per thread), which lets a NINTH chain of workload H run with two warps like the other eight.
// Per SM that is 2 % slower than 8 x 2, but 1 332 resident chains instead of 1 184 save a whole wave when the chain
// count falls just above a multiple of 1 184 (1 250 chains per GPU when 8 GPUs share 1e4 realizations: +19 %).
#ifndef SP_BATCH_LD64
#define SP_BATCH_LD64 1
#endif
// gather out of a batch's agent records: a random 64-byte record out of gigabytes -- ask the L2 for no more than that
__device__ __forceinline__ uint32_t sp_ldg_rec(const uint32_t* p) {
#if SP_BATCH_LD64
    uint32_t v;
    asm volatile("ld.global.nc.L2::64B.b32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
#else
    return __ldg(p);
#endif
}

#ifndef SP_CSR_SETS
#define SP_CSR_SETS 3
#endif

template <int BITS, bool SMEM, int GRAPH, bool ALIAS, bool RING, int MAXT = SP_RING_MAX_THREADS>
__global__ void __launch_bounds__(MAXT, 1) cnvm_native_kernel(NativeCnvmArgs A) {
    using P = Pack<BITS>;
    constexpr bool THR4 = (BITS <= 2);      // M <= 4: threshold table padded to 4 x 4 per branch
    // M <= 4 and plain counts: count deltas packed per lane, folded lazily; a grouped / weighted CV
    // (A.cw) updates its table with shared-memory atomics on every accepted event instead
    uint32_t packcnt_r = (THR4 && A.cw == nullptr) ? 1u : 0u;
    asm volatile("" : "+r"(packcnt_r));  // keep the flag in a register (otherwise re-derived from the 64-bit pointer per use)
    const bool PACKCNT = THR4 && packcnt_r != 0u;
    constexpr bool THR_SMEM = (BITS <= 4);  // M <= 16: threshold tables staged in shared memory
    constexpr uint32_t FULL = 0xffffffffu;
    constexpr int NS = (GRAPH == 1 && !ALIAS) ? SP_CSR_SETS : 3;  // stage register sets = own steps in flight
    constexpr int ISS2 = NS > 3 ? 2 : 1;
    extern __shared__ __align__(16) uint32_t smem[];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int W = RING ? A.warps_per_chain : 1;  // warps serving one chain
    const int slot = RING ? warp / W : warp;     // chain slot of this warp within the CTA
    const int role = RING ? warp % W : 0;        // RING: this warp takes the steps s with s % W == role
    const int slots = A.warps_per_cta / W;
    const int M = A.M, MM = M * M;
    const uint32_t n = (uint32_t)A.n;

    // shared layout: [thr] [reduction words, 4 per slot (COOP)] [counts slots*M] [states slots*words (if SMEM)]
    // thr: M <= 4 -> 32 (threshold, packed count delta of the move m -> nw) pairs indexed (noise << 4 | m << 2 | nw),
    //      one 8-byte load per event; M <= 16 -> 2*MM thresholds
    uint32_t* sp = smem;
    const uint32_t* thr = A.thr;  // (re-pointed per chain: the chain's parameter set, staged in shared memory)
    uint32_t* thr_stage = sp;     // this chain slot's copy of the table
    uint32_t thr4_saddr = 0;      // THR4: its shared-window address
    if (THR4) {
        thr_stage = sp + (size_t)slot * 64;
        thr4_saddr = (uint32_t)__cvta_generic_to_shared(thr_stage);
        asm volatile("" : "+r"(thr4_saddr));  // keep it in a register (ptxas re-derives it from SR_CgaCtaId per step otherwise)
        sp += (size_t)slots * 64;
    } else if (THR_SMEM) {
        thr_stage = sp + (size_t)slot * ((2 * MM + 1) & ~1);
        sp += (size_t)slots * ((2 * MM + 1) & ~1);
    }
    // COOP: three words per chain slot for the (rare) cross-warp reductions of the hazard path:
    // [0] lowest real writer, [1] lowest flagged lane (positions within the super-step)
    uint32_t* red = sp + (size_t)slot * 4;
    if (RING) {
        if (role == 0 && lane == 0) {
            red[0] = 0xffffffffu;
            red[1] = 0xffffffffu;
        }
        sp += (size_t)slots * 4;
    }
    int32_t* cnt_mem = (int32_t*)sp + slot * A.cnt_stride;
    sp += (size_t)slots * A.cnt_stride;
    uint32_t* state;
    if (SMEM) state = sp + (size_t)slot * A.words;
    else state = A.gstate + ((size_t)blockIdx.x * slots + slot) * A.words;
    __syncthreads();

    unsigned long long ev_total = 0, acc_total = 0, step_total = 0, extra_total = 0;
    const int T = A.T;

    for (int64_t cl = A.cl_begin + (int64_t)blockIdx.x * slots + slot; cl < A.cl_end; cl += (int64_t)gridDim.x * slots) {
        const int64_t j = cl / A.nrun, i = cl - j * A.nrun;
        const uint64_t chain = (uint64_t)(j * A.R_total + A.run_begin + i);
        const uint32_t ctr2 = (uint32_t)chain, ctr3 = (uint32_t)(chain >> 32) & 0x00FFFFFFu;  // tag 0
        const int64_t* kcounts = A.counts + cl * (int64_t)(T - 1);
        // GRAPH: 0 complete graph, 1 CSR, 2 neighbour sampler table, 3 a batch of per-run graphs with agent records
        const uint32_t grow = (GRAPH == 1 || GRAPH == 3) ? (uint32_t)cl * A.graph_stride : 0u;  // first row of the chain's own graph (batch)
        const uint32_t gcol = GRAPH == 3 ? (uint32_t)__ldg(&A.row[grow].x) : 0u;                // ... and its first adjacency entry
        // the chain's parameter set: noise-branch threshold, acceptance tables (staged per chain slot)
        const int64_t pset = j * A.param_stride;
        const uint32_t noise_thr = A.chain_noise_thr ? __ldg(A.chain_noise_thr + cl) : __ldg(A.noise_thr + pset);
        {
            const uint32_t* tab = A.thr + pset * (2 * MM);
            if (THR4) {
                // 32 (threshold, packed count delta of the move m -> nw) pairs indexed (noise << 4 | m << 2 | nw)
                for (int i = role * 32 + lane; i < 32; i += 32 * W) {
                    const int nz = i >> 4, m = (i >> 2) & 3, w = i & 3;
                    thr_stage[2 * i] = (m < M && w < M) ? __ldg(tab + nz * MM + m * M + w) : 0u;
                    thr_stage[2 * i + 1] = (1u << (8 * w)) - (1u << (8 * m));  // +1 in byte w, -1 in byte m (see fold_counts)
                }
                thr = thr_stage;
            } else if (THR_SMEM) {
                for (int i = role * 32 + lane; i < 2 * MM; i += 32 * W) thr_stage[i] = __ldg(tab + i);
                thr = thr_stage;
            } else {
                thr = tab;
            }
        }

        // ---- the initial state: packed once per call (pack_states_kernel), copied per chain ----
        {
            const int64_t jx = j * A.x_stride;
            const uint32_t* src = A.x_packed + jx * A.words;
            for (int w = role * 32 + lane; w < A.words; w += 32 * W) state[w] = __ldg(src + w);
            if (role == 0)
                for (int v = lane; v < A.cnt_stride; v += 32) cnt_mem[v] = __ldg(A.cnt_init + jx * A.cnt_stride + v);
        }
        if (!SMEM) __threadfence_block();
        if (RING) sp_ring_sync(slot + 1, 32 * W);
        else __syncwarp();

        // Per-lane packed deltas of the opinion counts (4 signed 8-bit fields, M <= 4), folded into the
        // shared counters before any field can leave [-127, 127]; 32-bit per-lane statistics with them.
        uint32_t dpack = 0, my_acc = 0, n_extra = 0;
        unsigned long long n_steps = 0, n_events = 0;  // (a chunk has up to 2^30 steps)
        auto fold_counts = [&]() {
            if (PACKCNT) {
                uint32_t t = dpack;
                const int d0 = (int)(int8_t)(t & 0xFF);
                t = (uint32_t)(((int)t - d0) >> 8);
                const int d1 = (int)(int8_t)(t & 0xFF);
                t = (uint32_t)(((int)t - d1) >> 8);
                const int d2 = (int)(int8_t)(t & 0xFF);
                t = (uint32_t)(((int)t - d2) >> 8);
                const int d3 = (int)(int8_t)(t & 0xFF);
                const int s0 = __reduce_add_sync(FULL, d0), s1 = __reduce_add_sync(FULL, d1);
                const int s2 = __reduce_add_sync(FULL, d2), s3 = __reduce_add_sync(FULL, d3);
                if (lane == 0) {
                    atomicAdd(&cnt_mem[0], s0);
                    if (M > 1) atomicAdd(&cnt_mem[1], s1);
                    if (M > 2) atomicAdd(&cnt_mem[2], s2);
                    if (M > 3) atomicAdd(&cnt_mem[3], s3);
                }
                dpack = 0;
                __syncwarp();
            }
        };

        // an agent of (class, weight) moves from opinion m to opinion nw: table update with shared atomics
        auto count_move = [&](uint32_t agent, uint32_t m, uint32_t nw) {
            int base = 0, w = 1;
            if (A.cw) {  // gathered only for accepted events (rare path: grouped / weighted CVs)
                const int2 cw = __ldg(A.cw + agent);
                base = cw.x * M;
                w = cw.y;
            }
            atomicAdd(&cnt_mem[base + (int)m], -w);
            atomicAdd(&cnt_mem[base + (int)nw], w);
        };
        // writes output time k (RING: only the warp that owns the completing step calls this)
        auto snapshot = [&](int k) {
            fold_counts();
            acc_total += __reduce_add_sync(FULL, my_acc);
            ev_total += n_events;
            step_total += n_steps;
            extra_total += n_extra;
            my_acc = n_events = n_steps = n_extra = 0;
            const int64_t snap = cl * (int64_t)T + k;
            if (BITS == 1 && A.ecv.kind != 0) {  // edge CV: scan the CV's graph against the on-chip state
                if (A.out_cv || A.out_sum || A.out_hist) {
                    unsigned long long twice;
                    double p0, p1;
                    sp_edge_cv_scan<SMEM>(state, n, A.ecv, lane, twice, p0, p1);
                    if (lane == 0) {
                        if (A.ecv.kind == SPONET_CV_INTERFACES) {
                            const long long c = (long long)(twice >> 1);  // every edge was seen from both ends
                            if (A.out_cv) A.out_cv[snap] = (A.cv_div == 1.0) ? (double)c : (double)c / A.cv_div;
                            if (A.out_sum) {
                                const int64_t o = j * (int64_t)T + k;
                                atomicAdd(A.out_sum + o, (unsigned long long)c);
                                atomicAdd(A.out_sumsq + o, (unsigned long long)(c * c));
                            }
                            if (A.out_hist) atomicAdd(A.out_hist + (j * (int64_t)T + k) * A.hist_bins + sp_hist_bin(c, A.hist_lo, A.hist_width, A.hist_bins), 1ull);
                        } else if (A.out_cv) {
                            A.out_cv[snap * 2 + 0] = (A.cv_div == 1.0) ? p0 : p0 / A.cv_div;
                            A.out_cv[snap * 2 + 1] = (A.cv_div == 1.0) ? p1 : p1 / A.cv_div;
                        }
                    }
                }
            } else if (A.out_cv || A.out_sum || A.out_hist) {
                for (int d = lane; d < A.D; d += 32) {
                    const long long c = cnt_mem[A.cv_idx[d]];
                    if (A.out_hist)
                        atomicAdd(A.out_hist + ((j * (int64_t)T + k) * A.D + d) * A.hist_bins + sp_hist_bin(c, A.hist_lo, A.hist_width, A.hist_bins), 1ull);
                    if (A.out_cv) {
                        const double dv = A.cv_divs ? A.cv_divs[d] : A.cv_div;
                        A.out_cv[snap * A.D + d] = (dv == 1.0) ? (double)c : (double)c / dv;
                    }
                    if (A.out_sum) {
                        const int64_t o = (j * (int64_t)T + k) * A.D + d;
                        atomicAdd(A.out_sum + o, (unsigned long long)c);
                        atomicAdd(A.out_sumsq + o, (unsigned long long)(c * c));
                    }
                }
            }
            if (A.out_x && (!A.final_only || k == T - 1)) {
                uint8_t* dst = A.out_x + (A.final_only ? cl : snap) * (int64_t)n;
                for (uint32_t a = lane; a < n; a += 32)
                    dst[a] = (uint8_t)((state_load<SMEM>(state + P::word(a)) >> P::shift(a)) & P::VM);
            }
        };

        // ---- stage registers, one set per in-flight own step (indices are compile-time after unrolling) ----
        // NS sets (own steps in flight): 3 -- the gather of an agent's entry is issued one step before it is used, the
        // column gather of the CSR path inside the step that uses it.  A batch of per-run graphs does not fit the L2:
        // with more sets the record header is gathered two steps ahead and the neighbour one step ahead (ISS2 = 2).
        uint32_t q_a[NS], q_r2[NS], q_r3[NS], q_meta[NS], q_nbr[NS];
        int2 q_row[NS];

        // Stage 1 of an own step: Philox, agent, gather of the agent's table entry (or CSR row).  With alias
        // tables (degree-weighted agent selection) the agent itself comes out of a gather, so the stage is
        // split: draw() runs three own steps ahead into one staging set (z_*) and gathers the alias entry,
        // resolve() one step later picks the agent and issues the entry gather -- each gather has a full
        // step to return instead of stalling the warp twice per step.
        uint32_t z_idx = 0, z_frac = 0, z_r2 = 0, z_r3 = 0, z_meta = 0;
        uint2 z_at = make_uint2(0u, 0u);
        auto gather_entry = [&](int s, uint32_t a, uint32_t r2) {
            // every lane gathers (no divergence); noise lanes ignore the result
            if (GRAPH == 1) q_row[s] = __ldg(A.row + (grow + a));
            if (GRAPH == 3) {  // (word index of the agent's record, its header)
                // Only imitation events read the graph here: every record gathered is a random DRAM row, and the
                // rate of those -- not bytes, not latency -- is what bounds this path (profiles/r2_kernels.md 6).
                const uint32_t base = (grow + a) << 4;
                q_row[s] = make_int2((int)base, (q_meta[s] & 1u) ? 1 : (int)sp_ldg_rec(A.rec + base));
            }
            if (GRAPH == 2) {
                // slot index (a << lg) | (top lg bits of r2) in one funnel shift; the table has < 2^26 entries
                const uint2 t = __ldg(A.nbtab + __funnelshift_l(r2, a, (uint32_t)A.nbtab_log2));
                q_row[s] = make_int2((int)t.x, (int)t.y);
            }
        };
        auto draw = [&](uint64_t e0) {  // ALIAS only
            const uint64_t e = e0 + (uint64_t)lane;
            const uint4 r = sp_philox4x32_10_rk((uint32_t)e, (uint32_t)(e >> 32), ctr2, ctr3, A.rk);
            z_idx = __umulhi(r.y, n);
            z_frac = r.y * n;
            z_r2 = r.z;
            z_r3 = r.w;
            z_meta = (uint32_t)(r.x < noise_thr) | (__umulhi(r.z, (uint32_t)M) << 8);
            z_at = __ldg(A.alias_tab + (grow + z_idx));  // (grow: the chain's own table in a batch)
        };
        auto resolve = [&](int s) {  // ALIAS only: sampling.py:117-122 with integer thresholds; noise events pick uniformly
            const uint32_t a = ((z_meta & 1u) | (uint32_t)sp_thr_accept(z_frac, z_at.x)) ? z_idx : z_at.y;
            q_a[s] = a;
            q_r2[s] = z_r2;
            q_r3[s] = z_r3;
            q_meta[s] = z_meta;
            gather_entry(s, a, z_r2);
        };
        auto issue1 = [&](int s, uint64_t e0) {  // uniform agent weights: one stage
            const uint64_t e = e0 + (uint64_t)lane;
            const uint4 r = sp_philox4x32_10_rk((uint32_t)e, (uint32_t)(e >> 32), ctr2, ctr3, A.rk);
            const uint32_t a = __umulhi(r.y, n);
            q_a[s] = a;
            q_r2[s] = r.z;
            q_r3[s] = r.w;
            q_meta[s] = (uint32_t)(r.x < noise_thr) | (__umulhi(r.z, (uint32_t)M) << 8);  // bit 0 noise, bits 8.. noise opinion
            gather_entry(s, a, r.z);
        };
        auto issue2 = [&](int s) {
            uint32_t nbr;
            if (GRAPH == 1) {
                nbr = (uint32_t)__ldg(A.col + ((uint32_t)q_row[s].x + __umulhi(q_r2[s], (uint32_t)q_row[s].y)));  // (32-bit unsigned offsets: a batch)
            } else if (GRAPH == 3) {  // neighbour k < 15 out of the record (same DRAM burst as the header), else out of the row
                const uint32_t hdr = (uint32_t)q_row[s].y, k = __umulhi(q_r2[s], hdr & A.rec_deg_mask);
                const bool in_rec = k < 15u;
                const uint32_t* src = in_rec ? A.rec : (const uint32_t*)A.col;
                nbr = (q_meta[s] & 1u) ? 0u : sp_ldg_rec(src + ((in_rec ? (uint32_t)q_row[s].x + 1u : gcol + (hdr >> A.rec_deg_bits)) + k));
            } else if (GRAPH == 2) {
                const uint32_t w0 = (uint32_t)q_row[s].x, w1 = (uint32_t)q_row[s].y;
                const uint32_t thr = ((w0 >> 20) << 12) | (w1 >> 20);
                const uint32_t frac = (q_r2[s] << A.nbtab_log2) >> 8;
                nbr = (frac < thr ? w0 : w1) & 0xFFFFFu;
            } else {
                nbr = __umulhi(q_r2[s], n - 1);
                nbr += (nbr >= q_a[s]);
            }
            q_nbr[s] = (q_meta[s] & 1u) ? q_a[s] : nbr;  // a noise event reads no neighbour: alias of its own agent
        };

        // ---------------------------------------------------------------------------------------------------
        // COMMIT: optimistic, verified, pipelined.
        //
        // A SUPER-STEP is the 32 * W consecutive events the W warps of a chain hold (warp `role` the events
        // 32 * role ... 32 * role + 31 of it: positions v = 32 * role + lane).  It runs in two phases separated by
        // barriers of the chain's warps (a warp barrier when W = 1):
        //   write phase   every accepting lane swaps its change into the packed state (compare-and-swap on the
        //                 agent's field: of several writers of one agent only the first writes, the others are
        //                 flagged, so an agent changes at most once per phase and changes cannot cancel)
        //   read phase    every lane re-reads the two locations its decision was based on -- and, in the same
        //                 pass, reads the locations of its NEXT step and that step's acceptance threshold.  Nobody
        //                 writes during a read phase, so the next step's decision is ready when the vote on the
        //                 current one comes back, and a super-step costs one shared-memory round trip per phase.
        // The super-step is exact unless some lane read a location that an EARLIER position really changed.  Let
        // w0 be the lowest position that really changes its agent: positions <= w0 have exact inputs in any case.
        // A location written in the phase holds a value different from the one everybody read, so the re-reads are
        // reliable: a changed neighbour, an own agent that is not what this lane left there, or a lost write flags
        // the lane.  The first position i* with stale inputs always flags itself, so with f = the lowest flagged
        // position, positions < max(f, w0 + 1) are exact.
        // On a flag everything is rolled back (XOR is its own inverse), the exact prefix is committed for real and
        // the remaining positions start over (tools/commit_model.py checks this protocol against the sequential
        // semantics on random instances).  At N = 1e5 about 1 % of the 32-event steps are flagged.
        // ---------------------------------------------------------------------------------------------------
        const uint32_t vbase = 32u * (uint32_t)role;
        auto coop_sync = [&]() {
            if (RING) sp_ring_sync(slot + 1, 32 * W);
            else __syncwarp();
        };
        auto coop_any = [&](bool p) -> bool {
            if (RING) return sp_ring_any(slot + 1, 32 * W, p);
            return __any_sync(FULL, p);
        };
        // decision of one lane for one step: locations, what it read, what it wants to write.
        // (The accepted flag lives across loop iterations.  As a bool ptxas packs it into a byte of a shared register
        // (PRMT in the loop), as a uint32_t it costs a register; which is faster depends on the instantiation --
        // measured on a B200: H (2-bit, two warps per chain) +4.9 % with the full register, the one-warp-per-chain and
        // the 1-bit 16-warp kernels +7 % / +6 % with the bool.)
        constexpr bool FLAGS_IN_DEC = !(BITS == 2 && RING);  // bool acc + bool lost inside Dec, or uint32_t acc + a local
        using acc_t = typename std::conditional<FLAGS_IN_DEC, bool, uint32_t>::type;
        struct Dec {
            uint32_t wa, wn;    // word index of the agent / of the location imitated (own agent for a noise event)
            uint32_t sa, sn;    // bit offsets within those words
            uint32_t word;      // the agent's packed word as read (the compare value of the write)
            uint32_t m, nw;     // opinion read, new opinion
            uint32_t delta;     // THR4: packed count delta of m -> nw
            acc_t acc;          // accepted
            bool lost;          // (FLAGS_IN_DEC) write phase: another lane changed this agent first, nothing was written
        };
        auto evaluate = [&](int s, bool active, Dec& d) {
            const uint32_t a = q_a[s], nbr = q_nbr[s], meta = q_meta[s];
            d.wa = P::word(a);
            d.wn = P::word(nbr);
            d.sa = P::shift(a);
            d.sn = P::shift(nbr);
            d.word = state_load<SMEM>(state + d.wa);
            d.m = (d.word >> d.sa) & P::VM;
            d.nw = (meta & 1u) ? (meta >> 8) : ((state_load<SMEM>(state + d.wn) >> d.sn) & P::VM);
            uint32_t th;
            d.delta = 0;
            if (THR4) {
                const uint2 e = sp_lds2(thr4_saddr + ((((meta & 1u) << 4) | (d.m << 2) | d.nw) << 3));
                th = e.x;
                d.delta = e.y;
            } else if (THR_SMEM) th = thr[((meta & 1u) ? MM : 0) + d.m * M + d.nw];
            else th = __ldg(thr + ((meta & 1u) ? MM : 0) + d.m * M + d.nw);
            d.acc = (acc_t)(active && sp_thr_accept(q_r3[s], th));
            if (FLAGS_IN_DEC) d.lost = false;
        };
        auto book = [&](int s, const Dec& d) {  // an accepted event m -> nw of this lane
            ++my_acc;
            if (PACKCNT) dpack += d.delta;
            else count_move(q_a[s], d.m, d.nw);
        };
        auto xmask_of = [&](const Dec& d) -> uint32_t { return (d.m ^ d.nw) << d.sa; };
        // Write phase of one step.  A compare-and-swap on the agent's word, retried while only OTHER agents of the
        // word have changed: the first writer of an agent wins, a later one finds the field changed, writes nothing
        // and is flagged (`lost`).  So an agent is written at most once per write phase and writes cannot cancel.
        auto write = [&](Dec& d) -> bool {  // returns "lost": another lane changed this agent first, nothing written
            bool lost = false;
            if (d.acc && d.m != d.nw) {
                const uint32_t x = xmask_of(d);
                uint32_t expect = d.word;
                for (;;) {
                    const uint32_t old = atomicCAS(state + d.wa, expect, expect ^ x);
                    if (old == expect) break;
                    if (((old >> d.sa) & P::VM) != d.m) {
                        if (FLAGS_IN_DEC) d.lost = true;
                        else lost = true;
                        break;
                    }
                    expect = old;
                }
            }
            return FLAGS_IN_DEC ? d.lost : lost;
        };
        // Between the write phase and the read phase.  Shared memory executes accesses in order: a barrier of the
        // chain's warps is enough.  In global memory (W = 1) the atomics are performed at the L2 and a load issued
        // right behind them could overtake them; every writer has consumed the value its atomic returned, so all
        // of them have been performed once the warp has voted.
        auto settle_writes = [&](bool lost) {
            if (SMEM) {
                coop_sync();
            } else {
                (void)__any_sync(FULL, lost);
                __threadfence_block();
            }
        };
        // re-reads the lane's two locations: has an input changed (or has the lane lost its write)?
        auto changed = [&](int s, bool pending, const Dec& d, bool lost_arg) -> bool {
            const bool lost = FLAGS_IN_DEC ? d.lost : lost_arg;
            uint32_t chk = (state_load<SMEM>(state + d.wa) >> d.sa) ^ ((d.acc && !lost) ? d.nw : d.m);
            if (!(q_meta[s] & 1u)) chk |= (state_load<SMEM>(state + d.wn) >> d.sn) ^ d.nw;
            return (pending && (chk & P::VM) != 0u) || lost;
        };
        // rare path, plain atomics followed by loads of other lanes
        auto settle_plain = [&]() {
            if (!SMEM) __threadfence();
            coop_sync();
        };
        // The step in set `s` was flagged: roll back, commit the exact prefix, re-run the rest until clean.
        auto hazard_rounds = [&](int s, bool active, Dec& d, bool bad, bool lost) {
            bool pending = active;
#pragma unroll 1
            for (;;) {
                const bool writer = d.acc && d.m != d.nw;
                if (writer && !(FLAGS_IN_DEC ? d.lost : lost)) atomicXor(state + d.wa, xmask_of(d));  // roll back
                const uint32_t writers = __ballot_sync(FULL, writer);
                const uint32_t badm = __ballot_sync(FULL, bad);
                uint32_t w0 = writers ? vbase + (uint32_t)__ffs(writers) - 1u : 0xffffffffu;
                uint32_t f = badm ? vbase + (uint32_t)__ffs(badm) - 1u : 0xffffffffu;
                if (RING) {
                    if (lane == 0) {
                        if (writers) atomicMin(red + 0, w0);
                        if (badm) atomicMin(red + 1, f);
                    }
                    coop_sync();
                    w0 = red[0];
                    f = red[1];
                    coop_sync();
                    if (role == 0 && lane == 0) {  // (the next reduction is at least two barriers away)
                        red[0] = 0xffffffffu;
                        red[1] = 0xffffffffu;
                    }
                }
                // positions <= w0 are exact in any case, positions below the first flagged one as well; without a
                // real writer nothing can be flagged (should it ever happen, everything is exact)
                uint32_t g = w0 == 0xffffffffu ? 0xffffffffu : w0 + 1u;
                if (f != 0xffffffffu && f > g) g = f;
                settle_plain();
                const uint32_t v = vbase + (uint32_t)lane;
                if (d.acc && v < g) {
                    if (writer) atomicXor(state + d.wa, xmask_of(d));
                    book(s, d);
                }
                pending = pending && v >= g;
                settle_plain();
                ++n_extra;
                evaluate(s, pending, d);  // the remaining positions against the updated state
                if (RING) coop_sync();    // all reads before anybody writes
                lost = write(d);
                settle_writes(lost);
                bad = changed(s, pending, d, lost);
                if (!coop_any(bad)) {
                    if (d.acc) book(s, d);
                    return;
                }
            }
        };

        // ---- the chain: output interval by output interval ----
        // Inside an interval of K events the 32-event steps are full except the last one; step t of the interval
        // belongs to warp t mod W, super-step u holds the steps u W ... u W + W - 1.  Every warp runs all the
        // super-steps of the interval (its step may be partial or not exist at the interval's end: nb < 32), the
        // software pipeline over the warp's own steps (depth 2) restarts per interval.
        uint64_t E = 0;  // events of the chain so far
        int k = 1;
        if (role == 0) snapshot(0);  // k = 0: the initial state
        while (k < T && kcounts[k - 1] == 0) {  // leading empty intervals repeat the initial state
            if (role == 0) snapshot(k);
            ++k;
        }
        coop_sync();
        int64_t K_left = k < T ? kcounts[k - 1] : 0;
        while (k < T) {
            // at most 2^30 steps at a time (32-bit step counters); the interval ends with the last chunk
            const int64_t Kc = K_left > ((int64_t)1 << 35) ? ((int64_t)1 << 35) : K_left;
            K_left -= Kc;
            const bool ends = K_left == 0;
            int k_next = k;
            if (ends) {  // output times completed by the last step: k and the empty intervals right after it
                k_next = k + 1;
                while (k_next < T && kcounts[k_next - 1] == 0) ++k_next;
            }
            const int nsteps = (int)((Kc + 31) >> 5);
            const int tail = (int)(Kc - 32 * (int64_t)(nsteps - 1));  // events of the last step
            const int n_super = (nsteps + W - 1) / W;
            const int n_own = nsteps > role ? (nsteps - role + W - 1) / W : 0;  // n_super or n_super - 1
            const bool own_last = (nsteps - 1) % W == role;
            // Only the LAST super-step of a chunk can hold a partial step (the chunk's last one, `tail` events) or
            // none at all for this warp; everywhere else the warp's step is full.
            const bool in_last = lane < (n_own == n_super ? (own_last ? tail : 32) : 0);
            // statistics per chunk, not per step: own steps are full except the chunk's last one
            n_steps += (unsigned long long)n_own;
            n_events += 32ull * (unsigned long long)n_own - ((own_last && n_own > 0) ? (unsigned long long)(32 - tail) : 0ull);
            {
                // The pipeline always runs two own steps ahead, also past the end of the chunk: the draws of a
                // step that does not exist are valid gathers that nobody consumes (two per chunk and warp),
                // which keeps the loop body free of prefetch conditions.
                const uint64_t e_stride = 32ull * (uint64_t)W;
                uint64_t e_pre = E + 32ull * (uint64_t)role;  // first event of the next step to draw
                if (ALIAS) {
                    draw(e_pre);
                    e_pre += e_stride;
                    resolve(0);
                    draw(e_pre);
                    e_pre += e_stride;
                    resolve(1);
                    issue2(0);
                    draw(e_pre);  // staged for the third own step
                    e_pre += e_stride;
                } else {
                    issue1(0, e_pre);
                    e_pre += e_stride;
                    if (NS == 3) issue2(0);
                    issue1(1, e_pre);
                    e_pre += e_stride;
                    if (NS > 3) {
#pragma unroll
                        for (int q = 2; q < NS - 1; ++q) {
                            issue1(q, e_pre);
                            e_pre += e_stride;
                        }
                        issue2(0);
                        issue2(1);
                    }
                }
                int rem = n_super;  // super-steps left in this chunk
                Dec dec[NS];  // decisions of the step being committed and of the next one (rotating with the stage sets)
                evaluate(0, rem != 1 || in_last, dec[0]);
                if (RING) coop_sync();  // all reads before anybody writes
                while (rem > 0) {
                    // blocks of at most 120 own steps (a multiple of the 3, 4 or 5 register sets): the packed count
                    // deltas of a lane move by at most 1 per step and are folded after each block
                    const int stop = rem > 120 ? rem - 120 : 0;
                    while (rem > stop) {
#pragma unroll
                        for (int u = 0; u < NS; ++u) {
                            if (rem == stop) break;
                            issue2((u + ISS2) % NS);  // the neighbour of the next own step (ISS2 = 2: of the one after), out of the entry gathered earlier
                            const bool active = rem != 1 || in_last;
                            const bool active_next = rem > 2 || (rem == 2 && in_last);
                            // write phase of this step, read phase: verify it, evaluate the next one
                            Dec& cur = dec[u];
                            Dec& nxt = dec[(u + 1) % NS];
                            const bool lost = write(cur);
                            settle_writes(lost);
                            // The draws of the step after the next one sit in the read phase on purpose: its block
                            // holds the shared-memory round trips (re-reads, next decision, threshold), and the
                            // Philox rounds are independent work to issue while those are in flight.
                            if (ALIAS) {
                                resolve((u + 2) % 3);  // two own steps ahead: agent + entry gather
                                draw(e_pre);           // three own steps ahead: Philox + alias gather
                            } else {
                                issue1((u + NS - 1) % NS, e_pre);  // NS - 1 own steps ahead: Philox + gather
                            }
                            e_pre += e_stride;
                            const bool bad = changed(u, active, cur, lost);
                            evaluate((u + 1) % NS, active_next, nxt);
                            if (!coop_any(bad)) {
                                if (cur.acc) book(u, cur);
                            } else {
                                hazard_rounds(u, active, cur, bad, lost);
                                evaluate((u + 1) % NS, active_next, nxt);  // the state has moved on
                                if (RING) coop_sync();
                            }
                            --rem;
                        }
                    }
                    if (rem > 0) fold_counts();
                }
            }
            E += (uint64_t)Kc;
            if (ends) {
                // every warp's count deltas, then the output times k .. k_next - 1 by the chain's first warp
                fold_counts();
                coop_sync();
                if (role == 0)
                    for (int q = k; q < k_next; ++q) snapshot(q);
                coop_sync();
                k = k_next;
                K_left = k < T ? kcounts[k - 1] : 0;
            }
        }
        ev_total += n_events;
        step_total += n_steps;
        extra_total += n_extra;
        acc_total += __reduce_add_sync(FULL, my_acc);
        coop_sync();
        if (!SMEM) __threadfence_block();
    }
    if (lane == 0 && A.counters) {
        atomicAdd(A.counters + 0, ev_total);
        atomicAdd(A.counters + 1, acc_total);
        atomicAdd(A.counters + 2, step_total);
        atomicAdd(A.counters + 3, extra_total);
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
    unsigned long long* out_hist;
    int32_t hist_bins;
    int64_t hist_lo, hist_width;
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
                if (A.out_hist)
                    atomicAdd(A.out_hist + ((j * (int64_t)A.T + k) * A.D + d) * A.hist_bins + sp_hist_bin(cv, A.hist_lo, A.hist_width, A.hist_bins), 1ull);
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
                                    double next_event_rate, int64_t* counts, const double* state_rates, int rates_per_chain) {
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t per = T - 1;
    if (tid >= num_chains * per) return;
    const int64_t cl = tid / per;
    const int k = (int)(tid - cl * per) + 1;
    uint64_t chain;
    if (nrun > 0) {
        const int64_t j = cl / nrun, i = cl - j * nrun;
        chain = (uint64_t)(j * R_total + run_begin + i);
        if (state_rates) next_event_rate = state_rates[rates_per_chain ? cl : j];  // sweep: one rate per state slot; weighted batch: per chain
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
                           int T, double next_event_rate, int64_t* d_counts, const double* d_state_rates, int rates_per_chain) {
    const int64_t total = num_chains * (int64_t)(T - 1);
    if (total <= 0) return SPONET_OK;
    const int block = 128;
    const int64_t grid = (total + block - 1) / block;
    SP_REQUIRE(grid < ((int64_t)1 << 31), "too many (chain, interval) pairs for one launch");
    event_counts_kernel<<<(unsigned)grid, block, 0, st>>>(seed, chain_begin, num_chains, nrun, R_total,
                                                          run_begin, d_t_eval, T, next_event_rate,
                                                          d_counts, d_state_rates, rates_per_chain);
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

// Initial opinions must lie in [0, M): a larger value would be packed into a neighbouring agent's bits and
// counted in another chain's counters.  Host buffers are checked here (the reference never checks and would
// index prob_*[x, .] out of bounds, cnvm/model.py:278,288); device-resident buffers cannot be checked without
// a synchronisation, the kernels clamp them instead.
int sp_validate_states(const uint8_t* x, int64_t count, int M) {
    if (!x || M >= 256 || sp_is_device_ptr(x)) return SPONET_OK;
    for (int64_t i = 0; i < count; ++i)
        if (x[i] >= M)
            return sponet_fail(SPONET_ERR_INVALID, "x_init[%lld] = %d is not an opinion in [0, %d)", (long long)i,
                               (int)x[i], M);
    return SPONET_OK;
}

// Stages the in-kernel CV descriptor.  `n` > 0 allows the per-agent class / weight arrays (CNVM agent-level
// kernel); kernels that only keep plain opinion counts pass n = 0.
int sp_stage_cv(SpScratch& sc, const sponet_shares_cv* cv, int M, SpCvDev* out, int64_t n) {
    out->D = 0;
    out->d_idx = nullptr;
    out->divisor = 1.0;
    out->d_divisors = nullptr;
    out->num_classes = 1;
    out->d_cw = nullptr;
    out->ecv = SpEdgeCv{0, nullptr, nullptr, nullptr, 0.0, 0.0, 0.0, 0.0};
    if (!cv) return SPONET_OK;
    if (cv->kind != SPONET_CV_SHARES) {  // Interfaces / Propensities: scanned from the on-chip state per output time
        SP_REQUIRE(cv->kind == SPONET_CV_INTERFACES || cv->kind == SPONET_CV_PROPENSITIES, "cv: unknown kind %d", cv->kind);
        if (n <= 0)
            return sponet_fail(SPONET_ERR_UNSUPPORTED, "cv: edge collective variables need the agent states (not the count-only kernel)");
        SP_REQUIRE(M == 2, "cv: Interfaces / Propensities are defined for 2 opinions (model has %d)", M);
        SP_REQUIRE(cv->graph && cv->graph->n == n && !cv->graph->pooled, "cv: edge collective variable needs a graph over the model's %lld agents", (long long)n);
        SP_REQUIRE(cv->dimension == (cv->kind == SPONET_CV_INTERFACES ? 1 : 2), "cv: dimension %d does not match the kind", cv->dimension);
        SP_REQUIRE(cv->divisor != 0.0, "cv: divisor must be non-zero");
        out->ecv.kind = cv->kind;
        out->ecv.row = cv->graph->d_row;
        out->ecv.col = cv->graph->d_col;
        if (cv->kind == SPONET_CV_PROPENSITIES) {
            SP_REQUIRE(cv->degree_alpha, "cv: Propensities need degree_alpha");
            double* d_da;
            SP_CUDA(sc.alloc((void**)&d_da, (size_t)n * sizeof(double)));
            SP_CUDA(cudaMemcpyAsync(d_da, cv->degree_alpha, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, sc.stream));
            SP_CUDA(cudaStreamSynchronize(sc.stream));
            out->ecv.da = d_da;
            out->ecv.r01 = cv->rates[0];
            out->ecv.r10 = cv->rates[1];
            out->ecv.rt01 = cv->rates[2];
            out->ecv.rt10 = cv->rates[3];
        }
        out->D = cv->dimension;
        out->divisor = cv->divisor;
        return SPONET_OK;
    }
    SP_REQUIRE(cv->dimension >= 1 && cv->idx, "cv: dimension/idx invalid");
    SP_REQUIRE(cv->divisors || cv->divisor != 0.0, "cv: divisor must be non-zero");
    const int C = cv->num_classes > 1 ? cv->num_classes : 1;
    const bool grouped = cv->agent_class || cv->agent_weight || C > 1;
    if (grouped && n <= 0)
        return sponet_fail(SPONET_ERR_UNSUPPORTED, "cv: agent classes / weights are served by the CNVM agent-level kernel only");
    SP_REQUIRE((int64_t)C * M <= 16384, "cv: too many classes (%d x %d opinions)", C, M);
    for (int d = 0; d < cv->dimension; ++d)
        SP_REQUIRE(cv->idx[d] >= 0 && cv->idx[d] < C * M, "cv: table index %d out of range", cv->idx[d]);
    int32_t* d_idx;
    SP_CUDA(sc.alloc((void**)&d_idx, cv->dimension * sizeof(int32_t)));
    SP_CUDA(cudaMemcpyAsync(d_idx, cv->idx, cv->dimension * sizeof(int32_t), cudaMemcpyHostToDevice, sc.stream));
    if (cv->divisors) {
        for (int d = 0; d < cv->dimension; ++d) SP_REQUIRE(cv->divisors[d] != 0.0, "cv: divisors must be non-zero");
        double* d_div;
        SP_CUDA(sc.alloc((void**)&d_div, cv->dimension * sizeof(double)));
        SP_CUDA(cudaMemcpyAsync(d_div, cv->divisors, cv->dimension * sizeof(double), cudaMemcpyHostToDevice, sc.stream));
        out->d_divisors = d_div;
    }
    if (grouped) {
        std::vector<int2> cw((size_t)n);
        for (int64_t i = 0; i < n; ++i) {
            const int32_t c = cv->agent_class ? cv->agent_class[i] : 0;
            SP_REQUIRE(c >= 0 && c < C, "cv: agent_class[%lld] = %d out of range", (long long)i, c);
            cw[i] = make_int2(c, cv->agent_weight ? cv->agent_weight[i] : 1);
        }
        int2* d_cw;
        SP_CUDA(sc.alloc((void**)&d_cw, (size_t)n * sizeof(int2)));
        SP_CUDA(cudaMemcpyAsync(d_cw, cw.data(), (size_t)n * sizeof(int2), cudaMemcpyHostToDevice, sc.stream));
        SP_CUDA(cudaStreamSynchronize(sc.stream));
        out->d_cw = d_cw;
    }
    SP_CUDA(cudaStreamSynchronize(sc.stream));
    out->D = cv->dimension;
    out->d_idx = d_idx;
    out->divisor = cv->divisor;
    out->num_classes = C;
    return SPONET_OK;
}

namespace {

// kernel selection over (BITS, SMEM, GRAPH, ALIAS, RING); RING exists for shared-memory states only
typedef void (*cnvm_kernel_t)(NativeCnvmArgs);

#ifdef SP_FAST_BUILD  // development builds: only the kernels of workload H (2-bit states in shared memory, sampler table)
cnvm_kernel_t pick_cnvm_kernel(int, bool, int, bool, bool ring, bool wide) {
    if (ring && wide) return cnvm_native_kernel<2, true, 2, false, true, 576>;
    return ring ? cnvm_native_kernel<2, true, 2, false, true> : cnvm_native_kernel<2, true, 2, false, false>;
}
#else
template <int BITS, bool SMEM, int GRAPH>
cnvm_kernel_t pick_cnvm(bool alias, bool ring, bool wide) {
    // the 576-thread build exists for the COOP kernels on networks (shared-memory states)
    if (SMEM && GRAPH != 0 && ring && wide)
        return alias ? cnvm_native_kernel<BITS, SMEM, GRAPH, true, SMEM, (SMEM && GRAPH != 0) ? 576 : SP_RING_MAX_THREADS>
                     : cnvm_native_kernel<BITS, SMEM, GRAPH, false, SMEM, (SMEM && GRAPH != 0) ? 576 : SP_RING_MAX_THREADS>;
    if (SMEM && ring) return alias ? cnvm_native_kernel<BITS, SMEM, GRAPH, true, SMEM> : cnvm_native_kernel<BITS, SMEM, GRAPH, false, SMEM>;
    return alias ? cnvm_native_kernel<BITS, SMEM, GRAPH, true, false> : cnvm_native_kernel<BITS, SMEM, GRAPH, false, false>;
}

template <int BITS, bool SMEM>
cnvm_kernel_t pick_cnvm_g(int graph, bool alias, bool ring, bool wide) {
    if (graph == 3)  // batches of per-run graphs: uniform agent weights only, 16 warps
        return (SMEM && ring) ? cnvm_native_kernel<BITS, SMEM, 3, false, SMEM> : cnvm_native_kernel<BITS, SMEM, 3, false, false>;
    if (graph == 2) return pick_cnvm<BITS, SMEM, 2>(alias, ring, wide);
    if (graph == 1) return pick_cnvm<BITS, SMEM, 1>(alias, ring, wide);
    return pick_cnvm<BITS, SMEM, 0>(alias, ring, false);
}

template <int BITS>
cnvm_kernel_t pick_cnvm_b(bool smem, int graph, bool alias, bool ring, bool wide) {
    return smem ? pick_cnvm_g<BITS, true>(graph, alias, ring, wide) : pick_cnvm_g<BITS, false>(graph, alias, false, false);
}

cnvm_kernel_t pick_cnvm_kernel(int bits, bool smem, int graph, bool alias, bool ring, bool wide) {
    switch (bits) {
        case 1: return pick_cnvm_b<1>(smem, graph, alias, ring, wide);
        case 2: return pick_cnvm_b<2>(smem, graph, alias, ring, wide);
        case 4: return pick_cnvm_b<4>(smem, graph, alias, ring, wide);
        default: return pick_cnvm_b<8>(smem, graph, alias, ring, wide);
    }
}
#endif

// Threshold tables of P parameter sets: thr[p][0:MM] imitation, thr[p][MM:2MM] noise; noise-branch thresholds [P].
int stage_thresholds(SpScratch& sc, const sponet_cnvm_params* p, int64_t P, const std::vector<double>& noise_prob,
                     const uint32_t** d_thr, const uint32_t** d_noise) {
    const int64_t MM = (int64_t)p->num_opinions * p->num_opinions;
    std::vector<uint32_t> h((size_t)P * 2 * MM + P);
    for (int64_t q = 0; q < P; ++q) {
        for (int64_t i = 0; i < MM; ++i) {
            SP_REQUIRE(p[q].prob_imit[i] >= 0.0 && p[q].prob_imit[i] <= 1.0 && p[q].prob_noise[i] >= 0.0 &&
                           p[q].prob_noise[i] <= 1.0,
                       "Probabilities have to be between 0 and 1.");
            h[q * 2 * MM + i] = sp_prob_to_thr(p[q].prob_imit[i]);
            h[q * 2 * MM + MM + i] = sp_prob_to_thr(p[q].prob_noise[i]);
        }
        h[(size_t)P * 2 * MM + q] = sp_prob_to_thr(noise_prob[q]);
    }
    uint32_t* d;
    SP_CUDA(sc.alloc((void**)&d, h.size() * sizeof(uint32_t)));
    SP_CUDA(cudaMemcpyAsync(d, h.data(), h.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, sc.stream));
    SP_CUDA(cudaStreamSynchronize(sc.stream));
    *d_thr = d;
    *d_noise = d + (size_t)P * 2 * MM;
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
                                  next_event_rate, d_c, nullptr, 0));
    SP_CUDA(sc.finish());
    return SPONET_OK;
}

// Shared implementation of sponet_cnvm_run (P = 1: one parameter set for all S initial states) and
// sponet_cnvm_run_sweep (sweep: S = P state slots, slot j runs parameter set j; x_shared: one initial state).
static int cnvm_run_impl(const sponet_graph* g, int64_t num_agents, const sponet_cnvm_params* p, int64_t P,
                         bool sweep, const uint8_t* x_init, bool x_shared, int64_t S, const double* t_eval,
                         int64_t T, const sponet_native_opts* o, const sponet_shares_cv* cv, uint8_t* out_x,
                         double* out_cv, int64_t* out_sum, int64_t* out_sumsq, uint64_t* counters,
                         void* cuda_stream) {
    SP_REQUIRE(p && x_init && P >= 1, "cnvm_run: null pointer");
    for (int64_t q = 0; q < P; ++q) {
        SP_REQUIRE(p[q].prob_imit && p[q].prob_noise && p[q].degree_alpha, "cnvm_run: null pointer in parameter set %lld", (long long)q);
        SP_REQUIRE(p[q].num_opinions == p[0].num_opinions && p[q].degree_alpha == p[0].degree_alpha,
                   "cnvm_run_sweep: all parameter sets must share num_opinions and degree_alpha (one network, one alpha)");
    }
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
        SP_REQUIRE(!g->pooled || g->batch == S * nrun, "graph batch holds %lld graphs, the call has %lld chains",
                   (long long)g->batch, (long long)(S * nrun));
    }
    int dev = 0;
    SP_CUDA(cudaGetDevice(&dev));
    if (g) SP_REQUIRE(g->device == dev, "graph lives on device %d, current device is %d", g->device, dev);
    const int64_t num_chains = S * nrun;
    if (num_chains == 0) return SPONET_OK;

    SP_TRY(sp_validate_states(x_init, (x_shared ? 1 : S) * n, M));
    SP_REQUIRE(!o->out_hist || (o->hist_bins >= 1 && o->hist_width >= 1 && cv), "out_hist needs hist_bins >= 1, hist_width >= 1 and a cv descriptor");

    SpScratch sc((cudaStream_t)cuda_stream);
    NativeCnvmArgs A{};
    A.n = (int32_t)n;
    A.M = M;
    std::vector<double> rates((size_t)P), noise_prob((size_t)P);
    for (int64_t q = 0; q < P; ++q)
        SP_TRY(sponet_cnvm_event_rates(g ? p->degree_alpha : nullptr, n, g ? 0.0 : p->degree_alpha[0], p[q].r_imit,
                                       p[q].r_noise, &rates[q], &noise_prob[q]));
    bool uniform = true;
    std::vector<double> chain_rates;  // per chain: a graph batch with agent weights
    if (g) {
        for (int64_t i = 1; i < n && uniform; ++i) uniform = (p->degree_alpha[i] == p->degree_alpha[0]);
        const bool weighted_batch = g->pooled && g->d_alias != nullptr;
        if (weighted_batch) {  // agent weights, event rate and noise probability of every chain from its own graph
            SP_REQUIRE(!sweep, "a graph batch with agent weights takes one parameter set");
            uniform = true;  // (no table from params->degree_alpha)
            A.alias_tab = g->d_alias;
            chain_rates.resize((size_t)g->batch);
            std::vector<uint32_t> cthr((size_t)g->batch);
            for (int64_t b = 0; b < g->batch; ++b) {
                const double denom = p->r_imit * g->h_weight_sum[b] + p->r_noise * (double)n;  // sponet_cnvm_event_rates
                if (denom == 0.0) return sponet_fail(SPONET_ERR_ZERO_RATE, "division by zero: all rates are zero");
                chain_rates[b] = 1.0 / denom;
                cthr[b] = sp_prob_to_thr(p->r_noise * (double)n * chain_rates[b]);
            }
            uint32_t* d_cthr;
            SP_CUDA(sc.alloc((void**)&d_cthr, cthr.size() * sizeof(uint32_t)));
            SP_CUDA(cudaMemcpyAsync(d_cthr, cthr.data(), cthr.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, sc.stream));
            SP_CUDA(cudaStreamSynchronize(sc.stream));
            A.chain_noise_thr = d_cthr;
        } else if (!uniform && g->pooled) {
            return sponet_fail(SPONET_ERR_UNSUPPORTED, "a graph batch needs uniform degree_alpha (alpha = 1) or sponet_graph_batch_set_weights");
        }
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
        A.graph_stride = g->pooled ? (uint32_t)n : 0u;
        A.rec = g->pooled ? g->d_rec : nullptr;
        A.rec_deg_bits = (uint32_t)g->rec_deg_bits;
        A.rec_deg_mask = (1u << g->rec_deg_bits) - 1u;
        A.nbtab = g->d_nbtab;
        A.nbtab_log2 = g->nbtab_log2;
    }
    SP_TRY(stage_thresholds(sc, p, P, noise_prob, &A.thr, &A.noise_thr));
    A.param_stride = sweep ? 1 : 0;
    A.x_stride = x_shared ? 0 : 1;
    SpCvDev cvd;
    SP_TRY(sp_stage_cv(sc, cv, M, &cvd, n));
    A.D = cvd.D;
    A.cv_idx = cvd.d_idx;
    A.cv_div = cvd.divisor;
    A.cv_divs = cvd.d_divisors;
    A.cw = cvd.d_cw;
    A.ecv = cvd.ecv;
    if (g && g->pooled && cvd.ecv.kind != 0)
        return sponet_fail(SPONET_ERR_UNSUPPORTED, "edge collective variables are not evaluated on a graph batch");
    SP_REQUIRE(cvd.ecv.kind != SPONET_CV_PROPENSITIES || (!out_sum && !out_sumsq && !o->out_hist),
               "cnvm_run: integer moment sums / histograms are not defined for Propensities");
    A.cnt_stride = cvd.num_classes * M;
    A.final_only = o->final_state_only ? 1 : 0;
    A.hist_bins = o->out_hist ? o->hist_bins : 0;
    A.hist_lo = o->hist_lo;
    A.hist_width = o->hist_width;
    A.S = S;
    A.nrun = nrun;
    A.R_total = o->num_runs_total;
    A.run_begin = o->run_begin + o->chain_offset;
    A.T = (int32_t)T;
    A.seed = o->seed;
    for (int r = 0; r < 10; ++r) {
        A.rk[2 * r] = (uint32_t)o->seed + (uint32_t)r * 0x9E3779B9u;
        A.rk[2 * r + 1] = (uint32_t)(o->seed >> 32) + (uint32_t)r * 0xBB67AE85u;
    }

    const uint8_t* d_x_init;
    SP_CUDA(sc.in(x_init, (size_t)(x_shared ? 1 : S) * n, &d_x_init));
    const double* d_t;
    SP_CUDA(sc.in(t_eval, (size_t)T, &d_t));
    // interval schedule
    if (T > 1) {
        if (o->event_counts) {
            SP_CUDA(sc.in(o->event_counts, (size_t)num_chains * (T - 1), &A.counts));
        } else {
            int64_t* d_c;
            SP_CUDA(sc.alloc((void**)&d_c, (size_t)num_chains * (T - 1) * sizeof(int64_t)));
            const double* d_rates = nullptr;
            if (!chain_rates.empty()) {
                double* d_r;
                SP_CUDA(sc.alloc((void**)&d_r, chain_rates.size() * sizeof(double)));
                SP_CUDA(cudaMemcpyAsync(d_r, chain_rates.data(), chain_rates.size() * sizeof(double), cudaMemcpyHostToDevice, sc.stream));
                SP_CUDA(cudaStreamSynchronize(sc.stream));
                d_rates = d_r;
            } else if (sweep) {  // one event rate per state slot
                double* d_r;
                SP_CUDA(sc.alloc((void**)&d_r, (size_t)P * sizeof(double)));
                SP_CUDA(cudaMemcpyAsync(d_r, rates.data(), (size_t)P * sizeof(double), cudaMemcpyHostToDevice, sc.stream));
                SP_CUDA(cudaStreamSynchronize(sc.stream));
                d_rates = d_r;
            }
            SP_TRY(sp_launch_event_counts(sc.stream, o->seed, 0, num_chains, nrun, o->num_runs_total,
                                          o->run_begin + o->chain_offset, d_t, (int)T, rates[0], d_c, d_rates,
                                          chain_rates.empty() ? 0 : 1));
            A.counts = d_c;
        }
    }
    SP_CUDA(sc.out(out_x, (size_t)num_chains * (A.final_only ? 1 : T) * n, &A.out_x));
    SP_CUDA(sc.out((unsigned long long*)o->out_hist, (size_t)S * T * cvd.D * (size_t)(A.hist_bins > 0 ? A.hist_bins : 0), &A.out_hist, true));
    SP_CUDA(sc.out(out_cv, (size_t)num_chains * T * cvd.D, &A.out_cv));
    SP_CUDA(sc.out((unsigned long long*)out_sum, (size_t)S * T * cvd.D, &A.out_sum, true));
    SP_CUDA(sc.out((unsigned long long*)out_sumsq, (size_t)S * T * cvd.D, &A.out_sumsq, true));
    SP_CUDA(sc.out((unsigned long long*)counters, SPONET_NUM_COUNTERS, &A.counters, true));

    // ---- launch geometry ----
    // slots = chains resident per CTA (one warp each, or a ring of warps_per_chain warps); shared memory per CTA =
    // thresholds + [reduction words] + counters + packed states.
    const int bits = M <= 2 ? 1 : M <= 4 ? 2 : M <= 16 ? 4 : 8;
    A.words = (int32_t)((n * bits + 31) / 32);
    {   // pack the initial states and count their opinions, once per call
        const int64_t Sx = x_shared ? 1 : S;
        uint32_t* d_packed;
        int32_t* d_cnt;
        SP_CUDA(sc.alloc((void**)&d_packed, (size_t)Sx * A.words * sizeof(uint32_t)));
        SP_CUDA(sc.alloc((void**)&d_cnt, (size_t)Sx * A.cnt_stride * sizeof(int32_t)));
        SP_CUDA(cudaMemsetAsync(d_cnt, 0, (size_t)Sx * A.cnt_stride * sizeof(int32_t), sc.stream));
        const dim3 pgrid((unsigned)((A.words + 255) / 256), (unsigned)std::min<int64_t>(Sx, 65535));
        switch (bits) {
            case 1: pack_states_kernel<1><<<pgrid, 256, 0, sc.stream>>>(d_x_init, Sx, n, M, A.words, A.cw, A.cnt_stride, d_packed, d_cnt); break;
            case 2: pack_states_kernel<2><<<pgrid, 256, 0, sc.stream>>>(d_x_init, Sx, n, M, A.words, A.cw, A.cnt_stride, d_packed, d_cnt); break;
            case 4: pack_states_kernel<4><<<pgrid, 256, 0, sc.stream>>>(d_x_init, Sx, n, M, A.words, A.cw, A.cnt_stride, d_packed, d_cnt); break;
            default: pack_states_kernel<8><<<pgrid, 256, 0, sc.stream>>>(d_x_init, Sx, n, M, A.words, A.cw, A.cnt_stride, d_packed, d_cnt); break;
        }
        SP_CUDA(cudaGetLastError());
        A.x_packed = d_packed;
        A.cnt_init = d_cnt;
    }
    SpDeviceLimits prop;
    SP_CUDA(sp_device_limits(dev, &prop));
    const size_t smem_max = prop.sharedMemPerBlockOptin;
    const size_t chain_bytes = (size_t)A.words * 4;
    const size_t fixed = 0;
    const size_t thr_bytes = bits <= 2 ? 256 : (bits <= 4 ? (size_t)((2 * M * M + 1) & ~1) * 4 : 0);  // per chain slot
    const size_t cnt_bytes = (size_t)A.cnt_stride * 4;           // opinion (x class) counters per chain
    const size_t per_slot_min = cnt_bytes + 16 + thr_bytes;      // + the reduction words of a COOP chain
    const int graph_kind = g == nullptr ? 0 : (g->d_nbtab ? 2 : (g->pooled && g->d_rec && !g->d_alias) ? 3 : 1);
    const bool batch_graph = g != nullptr && g->pooled;
    // Launches the local chains [c0, c1) with the geometry that suits their number; returns the chains one wave
    // of that geometry holds (resident chain slots on the whole device).
    auto launch_range = [&](int64_t c0, int64_t c1, bool wide, bool launch, int64_t* wave, int* wpc_out) -> int {
        const int64_t count = c1 - c0;
        const int max_warps = (wide ? 576 : SP_RING_MAX_THREADS) / 32;
        bool smem_state = !o->force_global_state;
        int slots = 0;
        if (smem_state) {
            if (smem_max > fixed + 64)
                slots = (int)std::min<size_t>((size_t)max_warps, (smem_max - fixed - 64) / (chain_bytes + per_slot_min));
            if (slots < 1) smem_state = false;  // a chain beyond 227 KB: states in global memory
        }
        int wpc = 1;  // warps per chain
        if (smem_state) {
            // no more slots than there are chains to run on this many SMs
            const int64_t need = (count + prop.multiProcessorCount - 1) / prop.multiProcessorCount;
            const int fit = (int)std::max<int64_t>(1, std::min<int64_t>(slots, need));
            if (o->warps_per_chain > 0) {
                wpc = std::min(o->warps_per_chain, max_warps);
                slots = std::max(1, std::min(fit, max_warps / wpc));
            } else {
                // Fill the CTA with warps: of all (slots <= fit, warps per chain) the pair with the most warps, more
                // chains on a tie (H, 512 threads: 9 chains fit, 8 x 2 warps beat 9 x 1; a chain that fills the SM's
                // shared memory alone gets every warp of the CTA: COOP super-steps of 32 * wpc events).
                // More warps per chain mean more events per super-step and more hazard rounds: a chain runs fastest
                // with about sqrt(0.7 N) events in flight (measured: N = 1e4 -> 2 warps, 1e5 -> 8, 1e6 -> 16+).
                int w_cap = 1;
                while (w_cap < 16 && (double)(64 * w_cap) * (double)(64 * w_cap) <= 0.7 * (double)n) w_cap *= 2;
                // A batch of per-run graphs lives in HBM and the rate of random DRAM rows bounds the kernel: fewer
                // graphs in flight, each read by more warps, find more open rows (N = 1e5, events/s with 1 / 2 / 4 / 8 /
                // 16 warps per chain: 2.9e10 / 3.4e10 / 4.8e10 / 3.9e10 / 2.2e10) -- half the cap, fewer chains on a tie.
                if (batch_graph) w_cap = std::max(1, w_cap / 2);
                int best = 0;
                for (int sl = fit; sl >= 1; --sl) {
                    int w = std::min(max_warps / sl, w_cap);
                    if (w > 1 && sl > 15) w = 1;  // named barriers 1..15 identify the chains of a CTA
                    if (sl * w > best || (batch_graph && sl * w == best)) {
                        best = sl * w;
                        slots = sl;
                        wpc = w;
                    }
                }
            }
            if (wpc > 1 && slots > 15) slots = 15;
            // the 18-warp build exists for the COOP kernels only
            if (wpc == 1 && slots > SP_RING_MAX_THREADS / 32) slots = SP_RING_MAX_THREADS / 32;
        } else {
            slots = 8;
        }
        const bool pair = wpc > 1;
        const int warps = slots * wpc;
        NativeCnvmArgs B = A;
        B.cl_begin = c0;
        B.cl_end = c1;
        B.warps_per_chain = wpc;
        size_t smem_bytes = fixed + (size_t)slots * thr_bytes + (pair ? (size_t)slots * 16 : 0) + (size_t)slots * cnt_bytes +
                            (smem_state ? (size_t)slots * chain_bytes : 0);
        B.warps_per_cta = warps;
        const int threads = warps * 32;
        if (wpc_out) *wpc_out = wpc;
        cnvm_kernel_t kern = pick_cnvm_kernel(bits, smem_state, graph_kind, A.alias_tab != nullptr, pair, wide && warps > 16);
        SP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
        int blocks_per_sm = 1;
        SP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, kern, threads, smem_bytes));
        SP_REQUIRE(blocks_per_sm >= 1, "cnvm_run: kernel does not fit on an SM (smem %zu B, %d threads)", smem_bytes, threads);
        const int64_t ctas_needed = (count + slots - 1) / slots;
        const int grid = (int)std::min<int64_t>(ctas_needed, (int64_t)prop.multiProcessorCount * blocks_per_sm);
        if (wave) *wave = (int64_t)prop.multiProcessorCount * blocks_per_sm * slots;
        if (!launch || c1 <= c0) return SPONET_OK;  // (geometry query only)
        if (!smem_state) SP_CUDA(sc.alloc((void**)&B.gstate, (size_t)grid * slots * chain_bytes));
        kern<<<grid, threads, smem_bytes, sc.stream>>>(B);
        SP_CUDA(cudaGetLastError());
        return SPONET_OK;
    };
    // Chains are dealt to the resident slots in waves.  A last wave that fills less than three quarters of the
    // slots would leave most of the device idle for a whole wave (1250 chains on 1184 slots: two waves for the work
    // of 1.06): that remainder runs as a launch of its own, whose geometry gives every chain more warps.  Explicit
    // warps_per_chain (tests, A/B runs) keeps the single launch.
    // Two plans: 16 warps per CTA, or (COOP kernels on networks) 18.  Each is a number of full waves plus a tail
    // launch; times are estimated in units of one 16-warp wave -- an 18-warp wave takes 1.15x as long (measured on H:
    // 30.3 vs 26.3 ms), a tail whose chains have W warps instead of 2 runs about {3: 0.74, 4: 0.6, >= 8: 0.45} of a
    // wave -- and the cheaper plan is taken.
    auto plan = [&](bool wide, int64_t* tail_begin, double* cost) -> int {
        int64_t wave = 0;
        int wpc_main = 1, wpc_tail = 1;
        SP_TRY(launch_range(0, num_chains, wide, false, &wave, &wpc_main));  // geometry of the whole call
        *tail_begin = num_chains;
        double t = (double)((num_chains + wave - 1) / std::max<int64_t>(wave, 1));
        if (o->warps_per_chain == 0 && !o->force_global_state && num_chains > wave && wave > 0) {
            const int64_t full = num_chains / wave * wave, rem = num_chains - full;
            if (rem > 0 && 4 * rem < 3 * wave) {
                *tail_begin = full;
                SP_TRY(launch_range(full, num_chains, wide, false, nullptr, &wpc_tail));
                const double ratio = (double)wpc_tail / (double)std::max(wpc_main, 1);
                const double tail = ratio >= 4.0 ? 0.45 : ratio >= 2.0 ? 0.6 : ratio >= 1.5 ? 0.74 : 1.0;
                t = (double)(full / wave) + tail;
            }
        }
        *cost = t * (wide ? 1.15 : 1.0);
        return SPONET_OK;
    };
    int64_t tail_begin = num_chains, tail_wide = num_chains;
    double cost = 0.0, cost_wide = 1e300;
    SP_TRY(plan(false, &tail_begin, &cost));
    bool wide = false;
    if (o->warps_per_chain == 0 && !o->force_global_state && graph_kind != 0 && !batch_graph) {
        int64_t wave_n = 0, wave_w = 0;
        SP_TRY(launch_range(0, num_chains, false, false, &wave_n, nullptr));
        SP_TRY(launch_range(0, num_chains, true, false, &wave_w, nullptr));
        if (wave_w > wave_n) {  // the two extra warps buy resident chains (not just warps per chain)
            SP_TRY(plan(true, &tail_wide, &cost_wide));
            if (cost_wide < cost) {
                wide = true;
                tail_begin = tail_wide;
            }
        }
    }
    SP_TRY(launch_range(0, tail_begin, wide, true, nullptr, nullptr));
    SP_TRY(launch_range(tail_begin, num_chains, wide, true, nullptr, nullptr));
    SP_CUDA(sc.finish());
    return SPONET_OK;
}

extern "C" int sponet_cnvm_run(const sponet_graph* g, int64_t num_agents, const sponet_cnvm_params* p,
                               const uint8_t* x_init, int64_t S, const double* t_eval, int64_t T,
                               const sponet_native_opts* o, const sponet_shares_cv* cv, uint8_t* out_x,
                               double* out_cv, int64_t* out_sum, int64_t* out_sumsq, uint64_t* counters,
                               void* cuda_stream) {
    return cnvm_run_impl(g, num_agents, p, 1, false, x_init, false, S, t_eval, T, o, cv, out_x, out_cv, out_sum, out_sumsq,
                         counters, cuda_stream);
}

extern "C" int sponet_cnvm_run_sweep(const sponet_graph* g, int64_t num_agents, const sponet_cnvm_params* p,
                                     int64_t P, const uint8_t* x_init, int32_t x_init_shared,
                                     const double* t_eval, int64_t T, const sponet_native_opts* o,
                                     const sponet_shares_cv* cv, uint8_t* out_x, double* out_cv,
                                     int64_t* out_sum, int64_t* out_sumsq, uint64_t* counters,
                                     void* cuda_stream) {
    SP_REQUIRE(P >= 1, "cnvm_run_sweep: need at least one parameter set");
    return cnvm_run_impl(g, num_agents, p, P, true, x_init, x_init_shared != 0, P, t_eval, T, o, cv, out_x, out_cv, out_sum,
                         out_sumsq, counters, cuda_stream);
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
    {
        const uint32_t* d_noise;
        SP_TRY(stage_thresholds(sc, p, 1, std::vector<double>(1, noise_probability), &A.thr, &d_noise));
    }
    SP_REQUIRE(!o->out_hist || (o->hist_bins >= 1 && o->hist_width >= 1), "out_hist needs hist_bins >= 1 and hist_width >= 1");
    SP_REQUIRE(!o->final_state_only, "count-only path: there are no agent states to return");
    SpCvDev cvd;
    SP_TRY(sp_stage_cv(sc, cv, M, &cvd, 0));
    if (cvd.d_divisors)
        return sponet_fail(SPONET_ERR_UNSUPPORTED, "count-only path: per-output divisors are not supported");
    A.D = cvd.D;
    A.cv_idx = cvd.d_idx;
    A.cv_div = cvd.divisor;
    A.S = S;
    A.nrun = nrun;
    A.R_total = o->num_runs_total;
    A.run_begin = o->run_begin + o->chain_offset;
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
                                          o->run_begin + o->chain_offset, d_t, (int)T, next_event_rate, d_c, nullptr, 0));
            A.counts = d_c;
        }
    }
    SP_CUDA(sc.out(out_cv, (size_t)num_chains * T * cvd.D, &A.out_cv));
    SP_CUDA(sc.out((unsigned long long*)out_sum, (size_t)S * T * cvd.D, &A.out_sum, true));
    SP_CUDA(sc.out((unsigned long long*)out_sumsq, (size_t)S * T * cvd.D, &A.out_sumsq, true));
    SP_CUDA(sc.out((unsigned long long*)counters, SPONET_NUM_COUNTERS, &A.counters, true));
    A.hist_bins = o->out_hist ? o->hist_bins : 0;
    A.hist_lo = o->hist_lo;
    A.hist_width = o->hist_width;
    SP_CUDA(sc.out((unsigned long long*)o->out_hist, (size_t)S * T * cvd.D * (size_t)A.hist_bins, &A.out_hist, true));
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
