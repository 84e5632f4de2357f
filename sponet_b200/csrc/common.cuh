// Shared host/device helpers of libsponet_b200: error reporting, host<->device staging of the
// caller's buffers, the graph handle, Philox4x32-10 and the draw->decision primitives of native mode.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/sponet_b200.h"

// ---------------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------------
std::string& sponet_err_slot();
int sponet_fail(int code, const char* fmt, ...);

#define SP_CUDA(call)                                                                           \
    do {                                                                                        \
        cudaError_t _e = (call);                                                                \
        if (_e != cudaSuccess)                                                                  \
            return sponet_fail(SPONET_ERR_CUDA, "%s failed: %s (%s:%d)", #call,                 \
                               cudaGetErrorString(_e), __FILE__, __LINE__);                     \
    } while (0)

#define SP_REQUIRE(cond, ...)                                          \
    do {                                                               \
        if (!(cond)) return sponet_fail(SPONET_ERR_INVALID, __VA_ARGS__); \
    } while (0)

#define SP_TRY(expr)                  \
    do {                              \
        int _rc = (expr);             \
        if (_rc != SPONET_OK) return _rc; \
    } while (0)

// ---------------------------------------------------------------------------------------------
// graph handle
// ---------------------------------------------------------------------------------------------
struct sponet_graph {
    int device;
    int64_t n, nnz;
    int32_t min_degree, max_degree;
    int32_t* d_row_ptr;  // [n+1]
    int2* d_row;         // [n] (begin, degree): one 8-byte gather per agent instead of two
    int32_t* d_col;      // [nnz]
    // Row-scanning kernels (CNTM) read whole rows: a second copy with every row padded to a multiple
    // of four entries (-1 sentinels) so that a lane fetches four neighbours per 16-byte load.
    int2* d_row4;        // [n] (first int4 index, degree)
    int4* d_col4;        // [sum_i ceil(deg_i / 4)]
    int64_t n_col4;
    // Neighbour sampler table (bounded-degree graphs that fit L2): 2^nbtab_log2 slots per agent, one
    // 8-byte entry per slot = (primary id | thr_hi << 20, alias id | thr_lo << 20).  A uniformly drawn
    // slot plus 24 further random bits select a uniformly distributed neighbour with ONE gather
    // instead of the dependent row -> column pair (see sponet_graph_neighbor_table in the header).
    uint2* d_nbtab;      // [n << nbtab_log2] or nullptr
    int32_t nbtab_log2;
    // A BATCH of `batch` graphs over the same n agents (graphgen.cu: one graph per chain of a native call):
    // d_row is [batch * n] with ABSOLUTE 32-bit offsets into the shared d_col / d_col4; no d_row_ptr, no table.
    int64_t batch;       // 1 for an ordinary graph
    // Batch, CNVM: one 64-byte record per agent = (degree | (row begin - graph's first entry) << rec_deg_bits, first 15
    // neighbours): the degree and the sampled neighbour of most agents share ONE DRAM burst (a batch never fits the L2);
    // neighbours 15.. are read from d_col.  rec_deg_bits = bits of the largest degree (>= 8).  nullptr when the batch was
    // created without records or degree and offset do not fit 32 bits together.
    uint32_t* d_rec;     // [batch * n * 16]
    int32_t rec_deg_bits;
    bool pooled;         // device memory from sp_pool_alloc (graphgen.cu)
    // Batch with degree-weighted agent selection (alpha != 1, sponet_graph_batch_set_weights): per graph the alias table of
    // the weights w(degree) (sponet/sampling.py:54-96 on every graph) and the sum of the weights (event rates).
    uint2* d_alias;      // [batch * n] (threshold, alias) or nullptr
    double* h_weight_sum;  // [batch] (host) or nullptr
};

// Device memory of graph batches (graphgen.cu): tens of gigabytes per ensemble call, allocated and released block after
// block.  cudaMalloc / cudaFree of that size cost 0.1 - 2 s a time, so released blocks are kept (up to 64 GiB, per
// process) and handed out again; sponet_graph_batch_trim() returns them to the driver.
cudaError_t sp_pool_alloc(void** out, size_t bytes);
void sp_pool_free(void* p);

// ---------------------------------------------------------------------------------------------
// staging: a caller pointer may live on the host or on the device
// ---------------------------------------------------------------------------------------------
inline bool sp_is_device_ptr(const void* p) {
    if (!p) return false;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

// RAII pool of temporary device allocations tied to one call.
// The two device limits the launch geometry needs (cudaGetDeviceProperties fills the whole struct and costs
// about a millisecond per call; these attributes are cheap).
struct SpDeviceLimits {
    int multiProcessorCount;
    size_t sharedMemPerBlockOptin;
};
inline cudaError_t sp_device_limits(int dev, SpDeviceLimits* out) {
    int sms = 0, smem = 0;
    cudaError_t e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return e;
    e = cudaDeviceGetAttribute(&smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (e != cudaSuccess) return e;
    out->multiProcessorCount = sms;
    out->sharedMemPerBlockOptin = (size_t)smem;
    return cudaSuccess;
}

struct SpScratch {
    cudaStream_t stream;
    std::vector<void*> ptrs;
    struct Copyback {
        void* host;
        const void* dev;
        size_t bytes;
    };
    std::vector<Copyback> copybacks;
    explicit SpScratch(cudaStream_t s) : stream(s) {
        // Keep up to 1 GiB of freed scratch in the device's stream-ordered pool.  With the default release
        // threshold (0) every synchronising call hands the memory back to the OS and the next call maps
        // it again, which showed up as 10-100 ms of host time per ensemble call.
        static bool tuned[64] = {};
        int dev = 0;
        if (cudaGetDevice(&dev) == cudaSuccess && dev >= 0 && dev < 64 && !tuned[dev]) {
            cudaMemPool_t pool;
            if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
                uint64_t keep = 1ull << 30;
                cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
            }
            tuned[dev] = true;
        }
    }
    ~SpScratch() {
        for (void* p : ptrs) cudaFreeAsync(p, stream);
    }
    cudaError_t alloc(void** out, size_t bytes) {
        if (bytes == 0) bytes = 16;
        cudaError_t e = cudaMallocAsync(out, bytes, stream);
        if (e == cudaSuccess) ptrs.push_back(*out);
        return e;
    }
    // Input: returns a device pointer holding `bytes` of `p` (copying if p is a host pointer).
    template <typename T>
    cudaError_t in(const T* p, size_t count, const T** out) {
        if (!p) { *out = nullptr; return cudaSuccess; }
        if (sp_is_device_ptr(p)) { *out = p; return cudaSuccess; }
        void* d;
        cudaError_t e = alloc(&d, count * sizeof(T));
        if (e != cudaSuccess) return e;
        *out = (const T*)d;
        return cudaMemcpyAsync(d, p, count * sizeof(T), cudaMemcpyHostToDevice, stream);
    }
    // Output: returns a device pointer; host targets get a device temp that finish() copies back.
    // `preload` copies the host contents in first (accumulate-into outputs).
    template <typename T>
    cudaError_t out(T* p, size_t count, T** dev, bool preload = false) {
        if (!p) { *dev = nullptr; return cudaSuccess; }
        if (sp_is_device_ptr(p)) { *dev = p; return cudaSuccess; }
        void* d;
        cudaError_t e = alloc(&d, count * sizeof(T));
        if (e != cudaSuccess) return e;
        *dev = (T*)d;
        copybacks.push_back({(void*)p, d, count * sizeof(T)});
        if (preload) return cudaMemcpyAsync(d, p, count * sizeof(T), cudaMemcpyHostToDevice, stream);
        return cudaSuccess;
    }
    // Copies host outputs back; synchronises only if there was anything to copy (or `force`).
    cudaError_t finish(bool force_sync = false) {
        for (auto& c : copybacks) {
            cudaError_t e = cudaMemcpyAsync(c.host, c.dev, c.bytes, cudaMemcpyDeviceToHost, stream);
            if (e != cudaSuccess) return e;
        }
        bool need = force_sync || !copybacks.empty();
        copybacks.clear();
        if (need) return cudaStreamSynchronize(stream);
        return cudaSuccess;
    }
};

// ---------------------------------------------------------------------------------------------
// native-mode primitives (device)
// ---------------------------------------------------------------------------------------------
#ifdef __CUDACC__

// Philox4x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3", SC'11).
__device__ __forceinline__ uint4 sp_philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                  uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        c0 = hi1 ^ c1 ^ k0;
        c1 = lo1;
        c2 = hi0 ^ c3 ^ k1;
        c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}

// Same function with the ten round keys precomputed on the host (kernel-parameter constants: the key
// schedule costs no instructions in the event loop).
#ifndef SP_PHILOX_ROUNDS
#define SP_PHILOX_ROUNDS 10  // (development switch: profiles/ has the cost of the generator measured with fewer rounds)
#endif
__device__ __forceinline__ uint4 sp_philox4x32_10_rk(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                     const uint32_t (&rk)[20]) {
#pragma unroll
    for (int r = 0; r < SP_PHILOX_ROUNDS; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        c0 = hi1 ^ c1 ^ rk[2 * r];
        c1 = lo1;
        c2 = hi0 ^ c3 ^ rk[2 * r + 1];
        c3 = lo0;
    }
    return make_uint4(c0, c1, c2, c3);
}

// counter = (e_lo, e_hi, chain_lo, chain_hi[24] | tag << 24), key = seed
__device__ __forceinline__ uint4 sp_native_draw(uint64_t seed, uint64_t chain, uint32_t tag, uint64_t e) {
    return sp_philox4x32_10((uint32_t)e, (uint32_t)(e >> 32), (uint32_t)chain,
                            ((uint32_t)(chain >> 32) & 0x00FFFFFFu) | (tag << 24), (uint32_t)seed,
                            (uint32_t)(seed >> 32));
}

// accept iff r < floor(p * 2^32); thr == 0xFFFFFFFF encodes p >= 1 (always).
__device__ __forceinline__ bool sp_thr_accept(uint32_t r, uint32_t thr) {
    return (r < thr) | (thr == 0xFFFFFFFFu);
}

// histogram bin of a raw integer CV value (sponet_native_opts.hist_*): clamp((c - lo) / width, 0, bins - 1)
__device__ __forceinline__ int64_t sp_hist_bin(long long c, int64_t lo, int64_t width, int32_t bins) {
    long long b = c >= lo ? (c - lo) / width : 0;
    return b < (long long)bins ? b : (long long)bins - 1;
}

__device__ __forceinline__ double sp_u53(uint32_t a, uint32_t b) {
    return (double)(((uint64_t)(a >> 5) << 26) | (uint64_t)(b >> 6)) * (1.0 / 9007199254740992.0);
}

#endif  // __CUDACC__

// p -> floor(p * 2^32) saturated to 0xFFFFFFFF
#ifdef __CUDACC__
__host__ __device__
#endif
inline uint32_t sp_prob_to_thr(double p) {
    if (!(p > 0.0)) return 0u;
    if (p >= 1.0) return 0xFFFFFFFFu;
    double t = p * 4294967296.0;
    if (t >= 4294967295.0) return 0xFFFFFFFFu;
    return (uint32_t)t;
}

// launchers implemented in the kernel translation units --------------------------------------
// Edge collective variable of a two-opinion chain (sponet_shares_cv.kind != 0), evaluated from the 1-bit
// packed state: one warp strides over the agents and walks each agent's CSR row.
struct SpEdgeCv {
    int32_t kind;      // 0 none, SPONET_CV_INTERFACES, SPONET_CV_PROPENSITIES
    const int2* row;   // (begin, degree) of the CV's network
    const int32_t* col;
    const double* da;  // PROPENSITIES: d(j)^alpha
    double r01, r10, rt01, rt10;
};

// iface = number of ordered pairs (j, v in N(j)) with x[j] != x[v] (twice the interface count);
// p0 / p1 = the two propensity sums.  Every lane returns the warp totals (fixed reduction tree).
// (Inlined, arguments by value: a call taking the address of a kernel-parameter member makes the compiler
// copy the whole parameter struct to local memory, which slowed the 1-bit CNVM kernels down 3x.)
template <bool SMEM>
__device__ __forceinline__ void sp_edge_cv_scan(const uint32_t* state, uint32_t n, const SpEdgeCv E, int lane,
                                                unsigned long long& iface, double& p0, double& p1) {
    auto bit = [&](uint32_t a) {
        const uint32_t w = SMEM ? state[a >> 5] : __ldcg(state + (a >> 5));
        return (w >> (a & 31)) & 1u;
    };
    unsigned int diff = 0;
    double a0 = 0.0, a1 = 0.0;
    for (uint32_t j = (uint32_t)lane; j < n; j += 32) {
        const uint32_t xj = bit(j);
        const int2 r = __ldg(E.row + j);
        unsigned int c = 0;
        for (int e = 0; e < r.y; ++e) c += bit((uint32_t)__ldg(E.col + r.x + e)) ^ xj;
        diff += c;
        if (E.kind == SPONET_CV_PROPENSITIES) {
            const double dj = __ldg(E.da + j);
            if (xj == 0) a0 += __dadd_rn(__ddiv_rn(__dmul_rn(E.r01, (double)c), dj), E.rt01);
            else a1 += __dadd_rn(__ddiv_rn(__dmul_rn(E.r10, (double)c), dj), E.rt10);
        }
    }
    unsigned long long t = diff;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        t += __shfl_xor_sync(0xffffffffu, t, o);
        a0 += __shfl_xor_sync(0xffffffffu, a0, o);
        a1 += __shfl_xor_sync(0xffffffffu, a1, o);
    }
    iface = t;
    p0 = a0;
    p1 = a1;
}

struct SpCvDev {
    SpEdgeCv ecv;
    int32_t D;
    const int32_t* d_idx;  // device [D]
    double divisor;
    const double* d_divisors;  // device [D] or nullptr
    int32_t num_classes;       // >= 1
    const int2* d_cw;          // device [n] (class, weight) or nullptr
};
