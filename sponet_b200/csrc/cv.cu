// Collective variables on stored states and the moment reduction.
//   sponet_opinion_shares  <- sponet/collective_variables.py:400-405 (_opinion_shares_numba:
//                             np.bincount(x[i], weights, minlength=M) per snapshot)
//   sponet_moments         <- sponet/sample_moments.py:107-108 (np.sum(x, 0), np.sum(x**2, 0))
#include "common.cuh"

namespace {

constexpr int CV_THREADS = 256;

// One CTA per snapshot.  Thread t sums agents t, t+256, ... into its private column of a
// [MAXM][256] shared array, then the columns are folded with a fixed binary tree: deterministic,
// and exact whenever the weights are integers (unweighted counts, degree weights, group masks).
template <int MAXM>
__global__ void __launch_bounds__(CV_THREADS) opinion_shares_kernel(const uint8_t* __restrict__ x,
                                                                    int64_t n, int M,
                                                                    const double* __restrict__ w,
                                                                    double* __restrict__ out) {
    __shared__ double acc[MAXM * CV_THREADS];
    const int t = threadIdx.x;
    const int64_t s = blockIdx.x;
#pragma unroll
    for (int m = 0; m < MAXM; ++m) acc[m * CV_THREADS + t] = 0.0;
    const uint8_t* xs = x + s * n;
    for (int64_t a = t; a < n; a += CV_THREADS) {
        const int m = xs[a];
        if (m < M) acc[m * CV_THREADS + t] += w ? w[a] : 1.0;
    }
    __syncthreads();
    for (int stride = CV_THREADS / 2; stride > 0; stride >>= 1) {
        if (t < stride) {
#pragma unroll
            for (int m = 0; m < MAXM; ++m)
                if (m < M) acc[m * CV_THREADS + t] += acc[m * CV_THREADS + t + stride];
        }
        __syncthreads();
    }
    if (t < M) out[s * M + t] = acc[t * CV_THREADS];
}

// Many opinions: shared-memory float64 atomics (order not fixed; exact for integer weights).
__global__ void __launch_bounds__(CV_THREADS) opinion_shares_atomic_kernel(const uint8_t* __restrict__ x,
                                                                           int64_t n, int M,
                                                                           const double* __restrict__ w,
                                                                           double* __restrict__ out) {
    extern __shared__ double hist[];
    const int t = threadIdx.x;
    const int64_t s = blockIdx.x;
    for (int m = t; m < M; m += CV_THREADS) hist[m] = 0.0;
    __syncthreads();
    const uint8_t* xs = x + s * n;
    for (int64_t a = t; a < n; a += CV_THREADS) {
        const int m = xs[a];
        if (m < M) atomicAdd(&hist[m], w ? w[a] : 1.0);
    }
    __syncthreads();
    for (int m = t; m < M; m += CV_THREADS) out[s * M + m] = hist[m];
}

// Thread per column, rows added in order -- the same order numpy uses for an axis-0 reduction.
__global__ void moments_kernel(const double* __restrict__ x, int64_t R, int64_t L,
                               double* __restrict__ sum, double* __restrict__ sumsq) {
    const int64_t l = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= L) return;
    double s = 0.0, q = 0.0;
    for (int64_t r = 0; r < R; ++r) {
        const double v = x[r * L + l];
        s = __dadd_rn(s, v);
        q = __dadd_rn(q, __dmul_rn(v, v));
    }
    sum[l] = s;
    sumsq[l] = q;
}

// One CTA per snapshot, threads stride over the directed CSR entries; every undirected edge is seen from
// both ends, so the count of differing pairs is halved (exact: the total is even).
__global__ void __launch_bounds__(CV_THREADS) interfaces_kernel(const uint8_t* __restrict__ x, int64_t n,
                                                                const int32_t* __restrict__ row_ptr,
                                                                const int32_t* __restrict__ col,
                                                                double* __restrict__ out) {
    __shared__ unsigned long long part[CV_THREADS / 32];
    const int64_t s = blockIdx.x;
    const uint8_t* xs = x + s * n;
    unsigned long long cnt = 0;
    for (int64_t i = threadIdx.x; i < n; i += CV_THREADS) {
        const uint8_t xi = xs[i];
        for (int32_t e = row_ptr[i]; e < row_ptr[i + 1]; ++e) cnt += (xs[col[e]] != xi);
    }
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_down_sync(0xffffffffu, cnt, o);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
        for (int w = 0; w < CV_THREADS / 32; ++w) t += part[w];
        out[s] = (double)(t / 2);
    }
}

// sponet/collective_variables.py:370-397: for every agent j in state m (n = 1 - m)
//   out[s, m] += r[m,n] * #{neighbours of j in state n} / degree_alpha[j] + r_tilde[m,n]
// (complete graph: the neighbour count is the global count of n, degree_alpha a scalar).
// Thread-strided partial sums folded with a fixed tree (the reference sums sequentially over j).
__global__ void __launch_bounds__(CV_THREADS) propensities_kernel(const uint8_t* __restrict__ x, int64_t n,
                                                                  const int32_t* __restrict__ row_ptr,
                                                                  const int32_t* __restrict__ col,
                                                                  const double* __restrict__ degree_alpha,
                                                                  double complete_degree_alpha, double r01,
                                                                  double r10, double rt01, double rt10,
                                                                  double* __restrict__ out) {
    __shared__ double acc[2][CV_THREADS];
    __shared__ unsigned long long ones_s;
    const int64_t s = blockIdx.x;
    const uint8_t* xs = x + s * n;
    const int t = threadIdx.x;
    if (!row_ptr) {  // complete graph: global count of ones first
        if (t == 0) ones_s = 0;
        __syncthreads();
        unsigned long long c = 0;
        for (int64_t i = t; i < n; i += CV_THREADS) c += xs[i];
        for (int o = 16; o > 0; o >>= 1) c += __shfl_down_sync(0xffffffffu, c, o);
        if ((t & 31) == 0) atomicAdd(&ones_s, c);
        __syncthreads();
    }
    double a0 = 0.0, a1 = 0.0;
    for (int64_t j = t; j < n; j += CV_THREADS) {
        const uint8_t m = xs[j];
        const uint8_t other = 1 - m;
        double cnt, da;
        if (row_ptr) {
            int c = 0;
            for (int32_t e = row_ptr[j]; e < row_ptr[j + 1]; ++e) c += (xs[col[e]] == other);
            cnt = (double)c;
            da = degree_alpha[j];
        } else {
            cnt = (double)(other ? ones_s : (unsigned long long)n - ones_s);
            da = complete_degree_alpha;
        }
        if (m == 0) a0 += r01 * cnt / da + rt01;
        else a1 += r10 * cnt / da + rt10;
    }
    acc[0][t] = a0;
    acc[1][t] = a1;
    __syncthreads();
    for (int stride = CV_THREADS / 2; stride > 0; stride >>= 1) {
        if (t < stride) {
            acc[0][t] += acc[0][t + stride];
            acc[1][t] += acc[1][t + stride];
        }
        __syncthreads();
    }
    if (t < 2) out[s * 2 + t] = acc[t][0];
}

}  // namespace

extern "C" int sponet_interfaces(const sponet_graph* g, const uint8_t* x, int64_t S, double* out,
                                 void* cuda_stream) {
    SP_REQUIRE(g && x && out && S >= 0, "interfaces: bad arguments");
    SP_REQUIRE(!g->pooled, "interfaces: a single graph is required, not a graph batch");
    if (S == 0) return SPONET_OK;
    SP_REQUIRE(S < ((int64_t)1 << 31), "interfaces: too many snapshots for one launch");
    SpScratch sc((cudaStream_t)cuda_stream);
    const uint8_t* d_x;
    double* d_out;
    SP_CUDA(sc.in(x, (size_t)S * g->n, &d_x));
    SP_CUDA(sc.out(out, (size_t)S, &d_out));
    interfaces_kernel<<<(unsigned)S, CV_THREADS, 0, sc.stream>>>(d_x, g->n, g->d_row_ptr, g->d_col, d_out);
    SP_CUDA(cudaGetLastError());
    SP_CUDA(sc.finish());
    return SPONET_OK;
}

extern "C" int sponet_propensities(const sponet_graph* g, int64_t num_agents, const uint8_t* x, int64_t S,
                                   const double* degree_alpha, double complete_degree_alpha,
                                   const double* r, const double* r_tilde, double* out, void* cuda_stream) {
    SP_REQUIRE(x && out && r && r_tilde && S >= 0 && num_agents >= 1, "propensities: bad arguments");
    SP_REQUIRE(!g || (g->n == num_agents && degree_alpha), "propensities: graph / degree_alpha mismatch");
    SP_REQUIRE(!g || !g->pooled, "propensities: a single graph is required, not a graph batch");
    if (S == 0) return SPONET_OK;
    SP_REQUIRE(S < ((int64_t)1 << 31), "propensities: too many snapshots for one launch");
    SpScratch sc((cudaStream_t)cuda_stream);
    const uint8_t* d_x;
    const double* d_da = nullptr;
    double* d_out;
    SP_CUDA(sc.in(x, (size_t)S * num_agents, &d_x));
    if (g) SP_CUDA(sc.in(degree_alpha, (size_t)num_agents, &d_da));
    SP_CUDA(sc.out(out, (size_t)S * 2, &d_out));
    propensities_kernel<<<(unsigned)S, CV_THREADS, 0, sc.stream>>>(
        d_x, num_agents, g ? g->d_row_ptr : nullptr, g ? g->d_col : nullptr, d_da, complete_degree_alpha,
        r[1], r[2], r_tilde[1], r_tilde[2], d_out);
    SP_CUDA(cudaGetLastError());
    SP_CUDA(sc.finish());
    return SPONET_OK;
}

extern "C" int sponet_opinion_shares(const uint8_t* x, int64_t S, int64_t n, int32_t M,
                                     const double* weights, double* out, void* cuda_stream) {
    SP_REQUIRE(x && out && S >= 0 && n >= 1 && M >= 1 && M <= 256, "opinion_shares: bad arguments");
    if (S == 0) return SPONET_OK;
    SP_REQUIRE(S < ((int64_t)1 << 31), "opinion_shares: too many snapshots for one launch");
    SpScratch sc((cudaStream_t)cuda_stream);
    const uint8_t* d_x;
    const double* d_w;
    double* d_out;
    SP_CUDA(sc.in(x, (size_t)S * n, &d_x));
    SP_CUDA(sc.in(weights, (size_t)n, &d_w));
    SP_CUDA(sc.out(out, (size_t)S * M, &d_out));
    const unsigned grid = (unsigned)S;
    if (M <= 2) opinion_shares_kernel<2><<<grid, CV_THREADS, 0, sc.stream>>>(d_x, n, M, d_w, d_out);
    else if (M <= 4) opinion_shares_kernel<4><<<grid, CV_THREADS, 0, sc.stream>>>(d_x, n, M, d_w, d_out);
    else if (M <= 8) opinion_shares_kernel<8><<<grid, CV_THREADS, 0, sc.stream>>>(d_x, n, M, d_w, d_out);
    else if (M <= 16) opinion_shares_kernel<16><<<grid, CV_THREADS, 0, sc.stream>>>(d_x, n, M, d_w, d_out);
    else opinion_shares_atomic_kernel<<<grid, CV_THREADS, M * sizeof(double), sc.stream>>>(d_x, n, M, d_w, d_out);
    SP_CUDA(cudaGetLastError());
    SP_CUDA(sc.finish());
    return SPONET_OK;
}

extern "C" int sponet_moments(const double* x, int64_t R, int64_t L, double* sum, double* sumsq,
                              void* cuda_stream) {
    SP_REQUIRE(x && sum && sumsq && R >= 0 && L >= 1, "moments: bad arguments");
    SpScratch sc((cudaStream_t)cuda_stream);
    const double* d_x;
    double *d_s, *d_q;
    SP_CUDA(sc.in(x, (size_t)R * L, &d_x));
    SP_CUDA(sc.out(sum, (size_t)L, &d_s));
    SP_CUDA(sc.out(sumsq, (size_t)L, &d_q));
    const int block = 128;
    moments_kernel<<<(unsigned)((L + block - 1) / block), block, 0, sc.stream>>>(d_x, R, L, d_s, d_q);
    SP_CUDA(cudaGetLastError());
    SP_CUDA(sc.finish());
    return SPONET_OK;
}

// ---------------------------------------------------------------------------------------------
// Initial-state samplers (sponet/states.py:45-164) on the device.  Native RNG: Philox4x32-10 keyed by the seed,
// counter (agent, state, tag 2).  One thread per state walks the agents in order.
//   counts == NULL   sample_states_uniform: x[s, i] = mulhi(r0, M), every opinion with probability 1 / M
//   counts != NULL   a uniformly random arrangement of the multiset with counts[s][m] agents of opinion m
//                    (sample_states_target_shares / sample_states_uniform_shares): sequential draws WITHOUT
//                    replacement from the urn -- agent i takes opinion m with probability left[m] / left_total --
//                    which is exactly the distribution of a shuffled multiset.  The position in the urn comes
//                    from a 64-bit uniform (mul64hi), so the quantisation is left_total / 2^64.
// The sequential statement is oracle_sample_states (oracle/oracle_native.inc).
// ---------------------------------------------------------------------------------------------
namespace {
__global__ void sample_states_kernel(int64_t n, int M, const int64_t* counts, int counts_shared, int64_t S,
                                     uint64_t seed, uint8_t* out) {
    extern __shared__ long long s_left[];  // [blockDim.x][M] remaining counts, thread-major
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    uint8_t* x = out + s * n;
    if (!counts) {
        for (int64_t i = 0; i < n; ++i) x[i] = (uint8_t)__umulhi(sp_native_draw(seed, (uint64_t)s, 2u, (uint64_t)i).x, (uint32_t)M);
        return;
    }
    long long* left = s_left + (size_t)threadIdx.x * M;
    const int64_t* c = counts + (counts_shared ? 0 : s * M);
    long long total = 0;
    for (int m = 0; m < M; ++m) {
        left[m] = c[m];
        total += c[m];
    }
    for (int64_t i = 0; i < n; ++i, --total) {
        const uint4 r = sp_native_draw(seed, (uint64_t)s, 2u, (uint64_t)i);
        unsigned long long pos = __umul64hi(((unsigned long long)r.x << 32) | r.y, (unsigned long long)total);
        int m = 0;
        while (pos >= (unsigned long long)left[m]) pos -= (unsigned long long)left[m++];
        --left[m];
        x[i] = (uint8_t)m;
    }
}
}  // namespace

extern "C" int sponet_sample_states(int64_t num_agents, int32_t num_opinions, const int64_t* counts,
                                    int32_t counts_shared, int64_t num_states, uint64_t seed, uint8_t* out,
                                    void* cuda_stream) {
    SP_REQUIRE(out && num_agents >= 1 && num_opinions >= 1 && num_opinions <= 256 && num_states >= 0, "sample_states: bad arguments");
    if (num_states == 0) return SPONET_OK;
    const int64_t rows = counts_shared ? 1 : num_states;
    if (counts && !sp_is_device_ptr(counts))
        for (int64_t j = 0; j < rows; ++j) {
            int64_t tot = 0;
            for (int v = 0; v < num_opinions; ++v) {
                SP_REQUIRE(counts[j * num_opinions + v] >= 0, "sample_states: counts must be non-negative");
                tot += counts[j * num_opinions + v];
            }
            SP_REQUIRE(tot == num_agents, "sample_states: counts row %lld sums to %lld, expected %lld", (long long)j,
                       (long long)tot, (long long)num_agents);
        }
    SpScratch sc((cudaStream_t)cuda_stream);
    const int64_t* d_c;
    uint8_t* d_out;
    SP_CUDA(sc.in(counts, (size_t)rows * num_opinions, &d_c));
    SP_CUDA(sc.out(out, (size_t)num_states * num_agents, &d_out));
    int block = 32;  // threads (states) per CTA; the remaining counts of a state sit in shared memory (8 M bytes each)
    while (counts && block > 1 && (size_t)block * num_opinions * sizeof(long long) > 40960) block >>= 1;
    const size_t smem = counts ? (size_t)block * num_opinions * sizeof(long long) : 0;
    sample_states_kernel<<<(unsigned)((num_states + block - 1) / block), block, smem, sc.stream>>>(
        num_agents, num_opinions, d_c, counts_shared, num_states, seed, d_out);
    SP_CUDA(cudaGetLastError());
    SP_CUDA(sc.finish());
    return SPONET_OK;
}
