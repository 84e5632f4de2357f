// NATIVE mode, CNTM (noisy threshold model): one warp per chain, 1 bit per agent in shared memory,
// 32 consecutive candidate events per warp step -- the same scheme as the CNVM kernel
// (native_cnvm.cu), with a row SCAN where the voter model has a single neighbour gather:
//   stage 1  Philox draw -> agent, noise flag, noise bit; row descriptor gather issued
//   stage 2  the first two 16-byte segments of the agent's (padded) neighbour row gathered
//   scan     every lane walks its own row: neighbour bits from shared memory, popcount of the
//            opposite opinion, compared with the precomputed minimum count for (threshold, degree)
//            -- the reference's float64 `count / deg >= threshold` (cntm/model.py:204-207) turned into
//            an exact integer table on the host.  The same pass probes the warp's Bloom filter with
//            every neighbour id: a hit means the neighbour may be ANOTHER lane's agent in this step
//   commit   dependency masks only for flagged lanes (exact, by re-walking the flagged lane's row);
//            hazard = dep & ballot(lanes that flip); lanes before the first hazard flip their bit
//            (atomicXor), the rest re-scan against the updated state
// Stages 1 and 2 of the next two steps are issued before the current scan (pipeline of depth 2,
// unrolled three ways).  The sequential statement this must reproduce bit for bit is
// oracle/oracle_native.inc (oracle_cntm_native).
#include <algorithm>

#include "common.cuh"

int sp_launch_event_counts(cudaStream_t st, uint64_t seed, int64_t chain_begin, int64_t num_chains,
                           int64_t nrun, int64_t R_total, int64_t run_begin, const double* d_t_eval,
                           int T, double next_event_rate, int64_t* d_counts, const double* d_state_rates, int rates_per_chain);
int sp_validate_native_common(const double* t_eval, int64_t T, const sponet_native_opts* o, int64_t S);
int sp_stage_cv(SpScratch& sc, const sponet_shares_cv* cv, int M, SpCvDev* out, int64_t n);
int sp_validate_states(const uint8_t* x, int64_t count, int M);

namespace {

struct NativeCntmArgs {
    int32_t n;
    const int2* row4;  // (first int4 index, degree)
    const int4* col4;  // rows padded to multiples of four with the row's own agent
    uint32_t graph_stride;  // n for a BATCH of graphs (chain cl of the call runs on graph cl), else 0
    uint32_t noise_thr;
    const int32_t* need01;  // [max_degree + 1] minimum count of ones for a 0 -> 1 flip (INT_MAX = never)
    const int32_t* need10;  // [max_degree + 1] minimum count of zeros for a 1 -> 0 flip
    const uint8_t* x_init;
    const int64_t* counts;
    int64_t S, nrun, R_total, run_begin;
    int32_t T;
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
    SpEdgeCv ecv;  // Interfaces / Propensities (kind != 0): replaces the share outputs
    unsigned long long* counters;
    uint32_t* gstate;
    int32_t words;
    int32_t warps_per_cta;
    int32_t bloom_words;
    uint32_t rk[20];
};

template <bool SMEM>
__device__ __forceinline__ uint32_t st_load(const uint32_t* p) {
    if (SMEM) return *p;
    return __ldcg(p);
}

#define SP_CNTM_MAX_WARPS 16
template <bool SMEM>
__global__ void __launch_bounds__(32 * SP_CNTM_MAX_WARPS, 1) cntm_native_kernel(NativeCntmArgs A) {
    constexpr uint32_t FULL = 0xffffffffu;
    extern __shared__ __align__(16) uint32_t smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t n = (uint32_t)A.n;
    const uint32_t noise_thr = A.noise_thr;
    const uint32_t lane_bit = 1u << lane, lane_lt = lane_bit - 1u;

    // per warp: 32 partial neighbour counts + one word of Bloom hits for the cooperative row tails
    __shared__ int s_coop[SP_CNTM_MAX_WARPS][33];
    int* coop = s_coop[warp];
    for (int i = lane; i < 33; i += 32) coop[i] = 0;
    // dynamic layout: [bloom W*bloom_words] [states W*words (if SMEM)]
    uint32_t* sp = smem;
    const uint32_t bloom_bits = (uint32_t)A.bloom_words * 32u;
    uint32_t* bloom = sp + (size_t)warp * A.bloom_words;
    for (int i = threadIdx.x; i < A.warps_per_cta * A.bloom_words; i += blockDim.x) sp[i] = 0u;
    sp += (size_t)A.warps_per_cta * A.bloom_words;
    uint32_t* state = SMEM ? sp + (size_t)warp * A.words
                           : A.gstate + ((size_t)blockIdx.x * A.warps_per_cta + warp) * A.words;
    __syncthreads();

    const int64_t total_warps = (int64_t)gridDim.x * A.warps_per_cta;
    const int64_t num_chains = A.S * A.nrun;
    const int T = A.T;
    unsigned long long ev_total = 0, acc_total = 0, step_total = 0, extra_total = 0;

    for (int64_t cl = (int64_t)blockIdx.x * A.warps_per_cta + warp; cl < num_chains; cl += total_warps) {
        const int64_t j = cl / A.nrun, i = cl - j * A.nrun;
        const uint64_t chain = (uint64_t)(j * A.R_total + A.run_begin + i);
        const uint32_t ctr2 = (uint32_t)chain, ctr3 = (uint32_t)(chain >> 32) & 0x00FFFFFFu;
        const uint8_t* x0 = A.x_init + j * (int64_t)n;
        const int64_t* kcounts = A.counts + cl * (int64_t)(T - 1);
        const uint32_t grow = (uint32_t)cl * A.graph_stride;  // first row of the chain's own graph (batch)

        int c1 = 0;  // number of ones (warp-uniform, exact at snapshots)
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
        c1 = __reduce_add_sync(FULL, c1);
        if (!SMEM) __threadfence_block();
        __syncwarp();

        int my_delta = 0;  // per-lane change of the number of ones since the last snapshot
        uint32_t my_acc = 0, n_steps = 0, n_events = 0, n_extra = 0;
        auto snapshot = [&](int k) {
            c1 += __reduce_add_sync(FULL, my_delta);
            acc_total += __reduce_add_sync(FULL, my_acc);
            ev_total += n_events;
            step_total += n_steps;
            extra_total += n_extra;
            my_delta = 0;
            my_acc = n_events = n_steps = n_extra = 0;
            const int64_t snap = cl * (int64_t)T + k;
            if (A.ecv.kind != 0) {  // edge CV: scan the CV's graph against the on-chip state
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
                    const long long c = A.cv_idx[d] == 1 ? (long long)c1 : (long long)n - c1;
                    if (A.out_hist)
                        atomicAdd(A.out_hist + ((j * (int64_t)T + k) * A.D + d) * A.hist_bins + sp_hist_bin(c, A.hist_lo, A.hist_width, A.hist_bins), 1ull);
                    if (A.out_cv)
                        A.out_cv[snap * A.D + d] = (A.cv_div == 1.0) ? (double)c : (double)c / A.cv_div;
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
                    dst[a] = (uint8_t)((st_load<SMEM>(state + (a >> 5)) >> (a & 31)) & 1u);
            }
        };

        // ---- software pipeline: three steps in flight, ONE copy of the step body ----
        // Set 0 is the step being committed (agent, flags, row, first two neighbour chunks), set 1 the next
        // one (its chunks are gathered while set 0 commits), set 2 the one after (Philox + row gather).
        // The sets rotate through register moves instead of a 3x unrolled body: the body is large (the row
        // scan), and three copies of it plus their re-scan paths overflowed the instruction cache
        // (no_instruction was 33 % of the stall samples).
        uint32_t a0 = 0, meta0 = 0, a1 = 0, meta1 = 0, a2 = 0, meta2 = 0;
        int2 row0 = make_int2(0, 0), row1 = row0, row2 = row0;
        const int4 none = make_int4(-1, -1, -1, -1);
        int4 c00 = none, c01 = none, c10 = none, c11 = none;

        auto draw = [&](uint64_t e0, uint32_t& a, uint32_t& meta, int2& row) {
            const uint64_t e = e0 + (uint64_t)lane;
            const uint4 r = sp_philox4x32_10_rk((uint32_t)e, (uint32_t)(e >> 32), ctr2, ctr3, A.rk);
            a = __umulhi(r.y, n);
            meta = (uint32_t)(r.x < noise_thr) | ((r.z >> 31) << 1);  // bit 0 noise, bit 1 the noise opinion
            row = __ldg(A.row4 + (grow + a));
        };
        auto gather = [&](const int2& row, int4& c0, int4& c1) {
            const int chunks = (row.y + 3) >> 2;
            c0 = chunks > 0 ? __ldg(A.col4 + row.x) : none;
            c1 = chunks > 1 ? __ldg(A.col4 + row.x + 1) : none;
        };

        // stage 3 + commit of the step in set 0 (nb events)
        auto process = [&](int nb) {
            const uint32_t a = a0, meta = meta0;
            const bool noise = meta & 1u;
            const int2 row = row0;
            const int deg = noise ? 0 : row.y;  // a noise event reads no neighbour
            const int chunks = (deg + 3) >> 2;
            const uint32_t wa = a >> 5, ba = a & 31;

            // tails of the rows beyond the two gathered chunks: one work list of (lane, chunk) items for
            // the whole warp.  With heavy-tailed degrees the longest row of a step is ~6x the mean and a
            // per-lane loop would run the warp at the pace of that one lane.
            const int t = chunks > 2 ? chunks - 2 : 0;
            int inc = t;  // inclusive prefix sum of the tail lengths over the lanes
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int v = __shfl_up_sync(FULL, inc, d);
                if (lane >= d) inc += v;
            }
            const int total = __shfl_sync(FULL, inc, 31);

            // ---- Bloom insert of the agents ----
            uint32_t ha = 0;
            bool hit = false;
            if (bloom_bits) {
                // Two bits per key inside one word (a blocked Bloom filter with k = 2): with 32 keys in ~13.6 Kbit a
                // probe hits by accident with probability ~1.5e-4 instead of 2.4e-3 for one bit -- 320 probes per
                // step, so ~5 % of the steps are flagged by accident instead of ~50 % (each flagged lane costs a
                // re-walk of its row with dependent gathers).
                const uint32_t h = a * 2654435761u;
                ha = __umulhi(h, (uint32_t)A.bloom_words);
                const uint32_t mask = (1u << (h & 31)) | (1u << ((h >> 5) & 31));
                hit = (atomicOr(bloom + ha, mask) & mask) == mask;  // an equal agent (or a collision)
                __syncwarp();
            }

            // Rounds: the first one scans every row (probing the Bloom filter with every neighbour), flags
            // the lanes that may depend on an earlier lane and builds their exact masks; lanes up to the
            // first hazard commit, the rest re-scan in the next round.  Almost always there is one round.
            bool pending = lane < nb;
            bool first_round = true;
            uint32_t dep = 0, flagged = 0;
#pragma unroll 1
            for (;;) {
                const uint32_t xa = (st_load<SMEM>(state + wa) >> ba) & 1u;
                const uint32_t other = xa ^ 1u;
                int cnt = 0;
                // `self` is the agent whose row the entry belongs to: padding entries hold it.  Its own bit
                // equals its opinion, never the opposite one, so padding adds nothing to the count; only the
                // Bloom probe has to skip it (the agent itself is in the filter).
                auto one = [&](int v, uint32_t self, uint32_t oth, int& c_out, bool& h_out) {
                    const uint32_t vv = (uint32_t)v;
                    const uint32_t bitv = (st_load<SMEM>(state + (vv >> 5)) >> (vv & 31)) & 1u;
                    c_out += (int)(bitv == oth);
                    if (bloom_bits) {  // (a re-scan probes the cleared filter: no hits, and none are used)
                        const uint32_t h = vv * 2654435761u;
                        const uint32_t mask = (1u << (h & 31)) | (1u << ((h >> 5) & 31));
                        h_out |= (vv != self) & ((bloom[__umulhi(h, (uint32_t)A.bloom_words)] & mask) == mask);
                    }
                };
                auto four = [&](const int4& v, uint32_t self, uint32_t oth, int& c_out, bool& h_out) {
                    one(v.x, self, oth, c_out, h_out);
                    one(v.y, self, oth, c_out, h_out);
                    one(v.z, self, oth, c_out, h_out);
                    one(v.w, self, oth, c_out, h_out);
                };
                if (chunks > 0) four(c00, a, other, cnt, hit);
                if (chunks > 1) four(c01, a, other, cnt, hit);
                if (total > 0) {  // warp-uniform
#pragma unroll 1
                    for (int base = 0; base < total; base += 32) {
                        const int it = base + lane;
                        int o = 0;  // owner of item `it`: the number of lanes whose prefix sum is <= it
#pragma unroll
                        for (int sft = 16; sft > 0; sft >>= 1) {
                            const int v = __shfl_sync(FULL, inc, o + sft - 1);
                            if (v <= it) o += sft;
                        }
                        const int first_item = __shfl_sync(FULL, inc - t, o);
                        const int rx = __shfl_sync(FULL, row.x, o);
                        const uint32_t oth = __shfl_sync(FULL, other, o);
                        const uint32_t a_o = __shfl_sync(FULL, a, o);
                        // lanes beyond the work list scan four copies of some owner's agent: no count, no probe
                        const int4 v = it < total ? __ldg(A.col4 + rx + 2 + (it - first_item))
                                                  : make_int4((int)a_o, (int)a_o, (int)a_o, (int)a_o);
                        int c_it = 0;
                        bool h_it = false;
                        four(v, a_o, oth, c_it, h_it);
                        if (c_it) atomicAdd(&coop[o], c_it);
                        if (h_it) atomicOr((uint32_t*)&coop[32], 1u << o);
                    }
                    __syncwarp();
                    cnt += coop[lane];
                    hit |= (bool)(((uint32_t)coop[32] >> lane) & 1u);
                    __syncwarp();
                    coop[lane] = 0;
                    if (lane == 0) coop[32] = 0;
                }
                uint32_t nw;
                if (noise) {
                    nw = meta >> 1;
                } else {
                    const int need = deg > 0 ? __ldg((xa ? A.need10 : A.need01) + deg) : 0x7fffffff;
                    nw = cnt >= need ? other : xa;
                }

                if (first_round) {
                    first_round = false;
                    if (bloom_bits) {
                        flagged = __ballot_sync(FULL, hit);
                        bloom[ha] = 0u;
                    } else {
                        flagged = FULL;
                    }
                    // exact masks for flagged lanes: earlier lanes whose agent is this lane's agent or one
                    // of its neighbours (the flagged lane's row is re-walked by the whole warp)
                    for (uint32_t f = flagged; f; f &= f - 1) {
                        const int il = __ffs(f) - 1;
                        const uint32_t ai = __shfl_sync(FULL, a, il);
                        const int rx = __shfl_sync(FULL, row.x, il), ch = __shfl_sync(FULL, chunks, il);
                        const uint32_t same_a = __ballot_sync(FULL, a == ai);
                        uint32_t touch = same_a;
                        for (int c = 0; c < ch; ++c) {
                            const int4 v = __ldg(A.col4 + rx + c);
                            touch |= __ballot_sync(FULL, ((int)a == v.x) | ((int)a == v.y) | ((int)a == v.z) | ((int)a == v.w));
                        }
                        if (a == ai) dep |= same_a & lane_lt;
                        if (lane == il) dep |= touch & lane_lt;
                    }
                }

                // ---- commit up to the first hazard ----
                const bool chg = pending && nw != xa;
                uint32_t hmask = 0;
                if (flagged) {
                    const uint32_t wmask = __ballot_sync(FULL, chg);
                    hmask = __ballot_sync(FULL, pending && (dep & wmask) != 0u);
                }
                const int first = hmask ? (__ffs(hmask) - 1) : 32;
                if (chg && lane < first) {
                    atomicXor(state + wa, 1u << ba);
                    my_delta += (int)nw - (int)xa;
                    ++my_acc;
                }
                pending = pending && lane >= first;
                if (!SMEM) __threadfence_block();
                __syncwarp();
                if (first == 32) break;
                ++n_extra;
            }
            ++n_steps;
            n_events += (uint32_t)nb;
        };

        // ---- the chain: output interval by output interval (same structure as the CNVM kernel) ----
        // Inside an interval of K events all steps are full except the last; the pipeline restarts per
        // interval and always runs two steps ahead (the draws past the end are valid, unused gathers).
        uint64_t E = 0;  // events of the chain so far
        int k = 1;
        snapshot(0);
        while (k < T && kcounts[k - 1] == 0) snapshot(k++);  // leading empty intervals repeat the initial state
        int64_t K_left = k < T ? kcounts[k - 1] : 0;
        while (k < T) {
            const int64_t Kc = K_left > ((int64_t)1 << 35) ? ((int64_t)1 << 35) : K_left;  // 32-bit step counters
            K_left -= Kc;
            const bool ends = K_left == 0;
            int k_next = k;
            if (ends) {
                k_next = k + 1;
                while (k_next < T && kcounts[k_next - 1] == 0) ++k_next;
            }
            const int nsteps = (int)((Kc + 31) >> 5);
            const int tail = (int)(Kc - 32 * (int64_t)(nsteps - 1));
            uint64_t e_pre = E;
            draw(e_pre, a0, meta0, row0);
            e_pre += 32;
            gather(row0, c00, c01);
            draw(e_pre, a1, meta1, row1);
            e_pre += 32;
            draw(e_pre, a2, meta2, row2);
            e_pre += 32;
#pragma unroll 1
            for (int rem = nsteps; rem > 0; --rem) {
                gather(row1, c10, c11);
                process(rem == 1 ? tail : 32);
                // rotate, then draw: every register that moves was loaded at least a full step ago
                a0 = a1, meta0 = meta1, row0 = row1, c00 = c10, c01 = c11;
                a1 = a2, meta1 = meta2, row1 = row2;
                draw(e_pre, a2, meta2, row2);
                e_pre += 32;
            }
            E += (uint64_t)Kc;
            if (ends) {
                for (int q = k; q < k_next; ++q) snapshot(q);
                k = k_next;
                K_left = k < T ? kcounts[k - 1] : 0;
            }
        }
        c1 += __reduce_add_sync(FULL, my_delta);
        acc_total += __reduce_add_sync(FULL, my_acc);
        ev_total += n_events;
        step_total += n_steps;
        extra_total += n_extra;
        __syncwarp();
    }
    if (lane == 0 && A.counters) {
        atomicAdd(A.counters + 0, ev_total);
        atomicAdd(A.counters + 1, acc_total);
        atomicAdd(A.counters + 2, step_total);
        atomicAdd(A.counters + 3, extra_total);
    }
}

// smallest c in [0, deg] with (double)c / (double)deg >= thr  (INT_MAX when there is none): the
// reference's `share_other_opinion >= threshold` with share = count / len(neighbors), made integer.
int32_t min_count(double thr, int deg) {
    if (deg <= 0) return 0x7fffffff;
    if (!(thr > 0.0)) return (0.0 >= thr) ? 0 : 0x7fffffff;  // thr <= 0 -> always; NaN -> never
    double g = std::ceil(thr * (double)deg);
    int c = g > (double)deg + 1.0 ? deg + 1 : (int)g;
    while (c > 0 && (double)(c - 1) / (double)deg >= thr) --c;
    while (c <= deg && !((double)c / (double)deg >= thr)) ++c;
    return c > deg ? 0x7fffffff : c;
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
    SP_TRY(sp_validate_states(x_init, S * n, 2));

    SpScratch sc((cudaStream_t)cuda_stream);
    NativeCntmArgs A{};
    A.n = (int32_t)n;
    A.row4 = g->d_row4;
    A.col4 = g->d_col4;
    SP_REQUIRE(!g->pooled || g->batch == num_chains, "graph batch holds %lld graphs, the call has %lld chains",
               (long long)g->batch, (long long)num_chains);
    SP_REQUIRE(g->d_row4 && g->d_col4, "cntm_run: the graph batch was created without padded rows");
    A.graph_stride = g->pooled ? (uint32_t)n : 0u;
    A.noise_thr = sp_prob_to_thr(p->noise_prob);
    {
        std::vector<int32_t> need((size_t)2 * (g->max_degree + 1));
        for (int d = 0; d <= g->max_degree; ++d) {
            need[d] = min_count(p->threshold_01, d);
            need[g->max_degree + 1 + d] = min_count(p->threshold_10, d);
        }
        int32_t* d_need;
        SP_CUDA(sc.alloc((void**)&d_need, need.size() * sizeof(int32_t)));
        SP_CUDA(cudaMemcpyAsync(d_need, need.data(), need.size() * sizeof(int32_t), cudaMemcpyHostToDevice, sc.stream));
        SP_CUDA(cudaStreamSynchronize(sc.stream));
        A.need01 = d_need;
        A.need10 = d_need + g->max_degree + 1;
    }
    SpCvDev cvd;
    SP_TRY(sp_stage_cv(sc, cv, 2, &cvd, (cv && cv->kind != SPONET_CV_SHARES) ? (int64_t)g->n : 0));
    if (cvd.d_divisors)
        return sponet_fail(SPONET_ERR_UNSUPPORTED, "cntm_run: per-output divisors are not supported");
    SP_REQUIRE(cvd.ecv.kind != SPONET_CV_PROPENSITIES || (!out_sum && !out_sumsq && !o->out_hist),
               "cntm_run: integer moment sums / histograms are not defined for Propensities");
    SP_REQUIRE(!o->out_hist || (o->hist_bins >= 1 && o->hist_width >= 1 && cv), "out_hist needs hist_bins >= 1, hist_width >= 1 and a cv descriptor");
    A.final_only = o->final_state_only ? 1 : 0;
    A.hist_bins = o->out_hist ? o->hist_bins : 0;
    A.hist_lo = o->hist_lo;
    A.hist_width = o->hist_width;
    A.ecv = cvd.ecv;
    if (g->pooled && cvd.ecv.kind != 0)
        return sponet_fail(SPONET_ERR_UNSUPPORTED, "edge collective variables are not evaluated on a graph batch");
    A.D = cvd.D;
    A.cv_idx = cvd.d_idx;
    A.cv_div = cvd.divisor;
    A.S = S;
    A.nrun = nrun;
    A.R_total = o->num_runs_total;
    A.run_begin = o->run_begin + o->chain_offset;
    A.T = (int32_t)T;
    for (int r = 0; r < 10; ++r) {
        A.rk[2 * r] = (uint32_t)o->seed + (uint32_t)r * 0x9E3779B9u;
        A.rk[2 * r + 1] = (uint32_t)(o->seed >> 32) + (uint32_t)r * 0xBB67AE85u;
    }
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
                                          o->run_begin + o->chain_offset, d_t, (int)T, p->next_event_rate, d_c, nullptr, 0));
            A.counts = d_c;
        }
    }
    SP_CUDA(sc.out(out_x, (size_t)num_chains * (A.final_only ? 1 : T) * n, &A.out_x));
    SP_CUDA(sc.out((unsigned long long*)o->out_hist, (size_t)S * T * cvd.D * (size_t)A.hist_bins, &A.out_hist, true));
    SP_CUDA(sc.out(out_cv, (size_t)num_chains * T * cvd.D, &A.out_cv));
    SP_CUDA(sc.out((unsigned long long*)out_sum, (size_t)S * T * cvd.D, &A.out_sum, true));
    SP_CUDA(sc.out((unsigned long long*)out_sumsq, (size_t)S * T * cvd.D, &A.out_sumsq, true));
    SP_CUDA(sc.out((unsigned long long*)counters, SPONET_NUM_COUNTERS, &A.counters, true));

    A.words = (int32_t)((n + 31) / 32);
    SpDeviceLimits prop;
    SP_CUDA(sp_device_limits(dev, &prop));
    const size_t smem_max = prop.sharedMemPerBlockOptin - SP_CNTM_MAX_WARPS * 33 * 4;  // minus the static arrays
    const size_t chain_bytes = (size_t)A.words * 4;
    bool smem_state = !o->force_global_state;
    int warps = 0;
    if (smem_state) {
        warps = (int)std::min<size_t>(16, (smem_max - 64) / chain_bytes);
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
    {   // Bloom filter per warp from the spare shared memory: a lane probes ~deg neighbours per step, so
        // the flag rate is ~32 * <deg> / bits; up to 64 Kbit, no more than ~N/2 bits for small chains
        const size_t spare = smem_max > smem_bytes + 64 ? (smem_max - smem_bytes - 64) / warps : 0;
        size_t words = std::min<size_t>(spare / 4, 2048);
        words = std::min<size_t>(words, std::max<size_t>(64, (size_t)n / 64));
        A.bloom_words = words >= 32 ? (int32_t)words : 0;
        smem_bytes += (size_t)A.bloom_words * 4 * warps;
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
