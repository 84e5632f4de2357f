// Batches of Erdős–Rényi graphs G(n, p) generated ON the device, one per chain of a native call.
//
// What this replaces: sponet/network_generator.py:26-85 (ErdosRenyiGenerator.__call__: networkx
// fast_gnp_random_graph, isolates rejected or patched by _unisolate_vertices :446-454) called once per run by
// cnvm/model.py:95-96 (update_network inside simulate) from sample_many_runs (multiprocessing.py:173-178).
// On the host that is ~0.7 s per graph at N = 1e5 -- several times the GPU time of the chain itself.
//
// The graphs are not networkx's (the reference pins no graph either: its generators draw from a numpy Generator
// through Python's random module); they are G(n, p) by the same geometric-skip construction, keyed by
// (seed, graph index, attempt, vertex), so a graph does not depend on the batch it is generated in nor on the
// number of GPUs.  oracle/oracle_native.inc (oracle_er_graph) restates the construction sequentially; the two
// agree bit for bit (the logarithm is sponet_log1p.h's, identical on both sides).
//
// Construction of graph c, attempt a:
//   vertex i owns the pairs (i, j), j > i.  Starting from j = i, repeat: v = 53 random bits * 2^-53,
//   j += 1 + floor(log1p(-v) / log1p(-p)); while j < n: edge {i, j}.  Draw k of vertex i is half of
//   Philox4x32-10(counter = (k / 2, i, c_lo, c_hi[16] | a[8] << 16 | 3 << 24), key = seed).
//   Without force_no_isolates a graph with an isolated vertex is rejected and attempt a + 1 is generated
//   (reference :64-69); with it, vertices 0 .. n-1 are visited in order and one that is (still) isolated gets an
//   edge to j = floor(u n) != i, u from counter (k, i, c_lo, c_hi | 4 << 24), k = 0, 1, ... until j != i
//   (reference :446-454).
//   Rows are sorted by neighbour id.
#include <cub/cub.cuh>

#include <cstdlib>
#include <mutex>
#include <unordered_map>

#include "common.cuh"
#include "sponet_log1p.h"

// ---- pool of large device blocks (see common.cuh) ----
namespace {
struct PoolBlock {
    void* p;
    size_t bytes;
    int dev;
};
std::mutex pool_mu;
std::vector<PoolBlock> pool_idle;                 // released, oldest first
std::unordered_map<void*, PoolBlock> pool_live;   // handed out
size_t pool_idle_bytes = 0;
constexpr size_t POOL_KEEP = (size_t)64 << 30, POOL_MIN = (size_t)1 << 20;

void pool_evict_locked(size_t keep) {
    while (!pool_idle.empty() && (pool_idle_bytes > keep || pool_idle.size() > 32)) {
        cudaFree(pool_idle.front().p);
        pool_idle_bytes -= pool_idle.front().bytes;
        pool_idle.erase(pool_idle.begin());
    }
}
}  // namespace

cudaError_t sp_pool_alloc(void** out, size_t bytes) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (bytes >= POOL_MIN) {
        std::lock_guard<std::mutex> lk(pool_mu);
        int best = -1;
        for (int i = 0; i < (int)pool_idle.size(); ++i) {
            const PoolBlock& b = pool_idle[i];
            if (b.dev == dev && b.bytes >= bytes && b.bytes <= bytes + bytes / 2 && (best < 0 || b.bytes < pool_idle[best].bytes)) best = i;
        }
        if (best >= 0) {
            const PoolBlock b = pool_idle[best];
            pool_idle.erase(pool_idle.begin() + best);
            pool_idle_bytes -= b.bytes;
            pool_live[b.p] = b;
            *out = b.p;
            return cudaSuccess;
        }
    }
    e = cudaMalloc(out, bytes);
    if (e == cudaErrorMemoryAllocation) {  // give the idle blocks back and try once more
        cudaGetLastError();
        {
            std::lock_guard<std::mutex> lk(pool_mu);
            pool_evict_locked(0);
        }
        e = cudaMalloc(out, bytes);
    }
    if (e == cudaSuccess && bytes >= POOL_MIN) {
        std::lock_guard<std::mutex> lk(pool_mu);
        pool_live[*out] = PoolBlock{*out, bytes, dev};
    }
    return e;
}

void sp_pool_free(void* p) {  // (the caller has synchronised the device)
    if (!p) return;
    std::lock_guard<std::mutex> lk(pool_mu);
    auto it = pool_live.find(p);
    if (it == pool_live.end()) {
        cudaFree(p);
        return;
    }
    pool_idle.push_back(it->second);
    pool_idle_bytes += it->second.bytes;
    pool_live.erase(it);
    pool_evict_locked(POOL_KEEP);
}

extern "C" int sponet_graph_batch_trim(void) {
    cudaDeviceSynchronize();
    std::lock_guard<std::mutex> lk(pool_mu);
    pool_evict_locked(0);
    return SPONET_OK;
}

namespace {

struct ErArgs {
    int64_t n, count;            // vertices per graph, graphs in the batch
    int64_t nrun, R_total, run_begin;  // slot b = j * nrun + i  ->  graph index c = j * R_total + run_begin + i
    double log1mp;               // log1p(-p) (< 0), or 0 for p >= 1 (complete graph)
    uint64_t seed;
};

__device__ __forceinline__ uint64_t er_graph_index(const ErArgs& A, int64_t b) {
    const int64_t j = b / A.nrun, i = b - j * A.nrun;
    return (uint64_t)(j * A.R_total + A.run_begin + i);
}

// the upper-triangle neighbours of vertex i of graph c, attempt a: calls f(j) for each, in increasing order
template <typename F>
__device__ __forceinline__ void er_row(const ErArgs& A, uint64_t c, uint32_t attempt, int64_t i, F&& f) {
    const uint32_t c2 = (uint32_t)c, c3 = ((uint32_t)(c >> 32) & 0xFFFFu) | ((attempt & 0xFFu) << 16) | (3u << 24);
    int64_t j = i;
    const int64_t n = A.n;
    for (uint32_t k = 0;; ++k) {
        const uint4 r = sp_philox4x32_10(k, (uint32_t)i, c2, c3, (uint32_t)A.seed, (uint32_t)(A.seed >> 32));
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const uint32_t hi = h ? r.z : r.x, lo = h ? r.w : r.y;
            int64_t skip = 0;
            if (A.log1mp != 0.0) {
                const double v = (double)((((uint64_t)hi >> 5) << 26) | ((uint64_t)lo >> 6)) * (1.0 / 9007199254740992.0);
                const double q = __ddiv_rn(sponet_log1p(-v), A.log1mp);  // >= 0
                if (!(q < (double)(n - j))) return;                      // beyond the last vertex
                skip = (int64_t)q;                                        // floor
            }
            j += 1 + skip;
            if (j >= n) return;
            f(j);
        }
    }
}

__global__ void er_degree_kernel(ErArgs A, const uint8_t* dirty, const uint32_t* attempt, uint32_t* deg) {
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= A.count * A.n) return;
    const int64_t b = tid / A.n, i = tid - b * A.n;
    if (!dirty[b]) return;
    uint32_t* d = deg + b * A.n;
    uint32_t up = 0;
    er_row(A, er_graph_index(A, b), attempt[b], i, [&](int64_t j) {
        ++up;
        atomicAdd(d + j, 1u);
    });
    if (up) atomicAdd(d + i, up);
}

// isolated vertices per graph (of the dirty slots)
__global__ void er_isolates_kernel(ErArgs A, const uint8_t* dirty, const uint32_t* deg, uint32_t* isolates) {
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= A.count * A.n) return;
    const int64_t b = tid / A.n;
    if (dirty[b] && deg[tid] == 0) atomicAdd(isolates + b, 1u);
}

// rejected graphs: next attempt, degrees cleared
__global__ void er_reset_kernel(ErArgs A, const uint8_t* dirty, uint32_t* deg) {
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= A.count * A.n) return;
    if (dirty[tid / A.n]) deg[tid] = 0;
}

// force_no_isolates: one warp per graph walks the vertices in order (reference :446-454)
__global__ void er_patch_kernel(ErArgs A, uint32_t* deg, int2* patch, uint32_t* n_patch, int64_t cap) {
    const int64_t b = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (b >= A.count) return;
    uint32_t* d = deg + b * A.n;
    const uint64_t c = er_graph_index(A, b);
    uint32_t np = 0;
    for (int64_t i0 = 0; i0 < A.n; i0 += 32) {
        const int64_t i = i0 + lane;
        uint32_t zero = __ballot_sync(0xffffffffu, i < A.n && d[i] == 0);
        if (lane == 0) {
            while (zero) {
                const int64_t v = i0 + __ffs(zero) - 1;
                zero &= zero - 1;
                if (d[v] != 0) continue;  // an earlier patch reached it
                int64_t j = v;
                for (uint32_t k = 0; j == v; ++k) {
                    const uint4 r = sp_philox4x32_10(k, (uint32_t)v, (uint32_t)c, ((uint32_t)(c >> 32) & 0xFFFFu) | (4u << 24),
                                                     (uint32_t)A.seed, (uint32_t)(A.seed >> 32));
                    j = (int64_t)__umulhi(r.x, (uint32_t)A.n);
                }
                d[v] += 1;
                d[j] += 1;
                if ((int64_t)np < cap) patch[b * cap + np] = make_int2((int)v, (int)j);
                ++np;
            }
        }
        __syncwarp();
    }
    if (lane == 0) n_patch[b] = np;
}

__global__ void er_fill_kernel(ErArgs A, const uint32_t* attempt, const uint32_t* off, uint32_t* cursor, int32_t* col) {
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= A.count * A.n) return;
    const int64_t b = tid / A.n, i = tid - b * A.n;
    const uint32_t* o = off + b * A.n;
    uint32_t* cur = cursor + b * A.n;
    er_row(A, er_graph_index(A, b), attempt[b], i, [&](int64_t j) {
        col[o[i] + atomicAdd(cur + i, 1u)] = (int32_t)j;
        col[o[j] + atomicAdd(cur + j, 1u)] = (int32_t)i;
    });
}

__global__ void er_fill_patch_kernel(ErArgs A, const int2* patch, const uint32_t* n_patch, int64_t cap, const uint32_t* off,
                                     uint32_t* cursor, int32_t* col) {
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= A.count * cap) return;
    const int64_t b = tid / cap, k = tid - b * cap;
    if (k >= (int64_t)n_patch[b]) return;
    const int2 e = patch[tid];
    const uint32_t* o = off + b * A.n;
    uint32_t* cur = cursor + b * A.n;
    col[o[e.x] + atomicAdd(cur + e.x, 1u)] = e.y;
    col[o[e.y] + atomicAdd(cur + e.y, 1u)] = e.x;
}

// sorts every row (the atomics above fill them in arbitrary order), writes the (begin, degree) descriptors; degree range of
// the batch and the largest number of adjacency entries of one graph
__global__ void batch_finish_kernel(int64_t rows, int64_t n, const uint32_t* off, const uint32_t* deg, int32_t* col, int2* row,
                                    uint32_t* dmin, uint32_t* dmax, uint32_t* gmax) {
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = tid < rows;
    const uint32_t d = live ? deg[tid] : 0u;
    int32_t* r = col + (live ? off[tid] : 0u);
    if (d <= 48u) {  // insertion sort: rows are short ...
        for (uint32_t a = 1; a < d; ++a) {
            const int32_t key = r[a];
            uint32_t q = a;
            while (q > 0 && r[q - 1] > key) {
                r[q] = r[q - 1];
                --q;
            }
            r[q] = key;
        }
    } else {  // ... except the hubs of a scale-free graph (thousands of neighbours): heap sort
        auto sift = [&](uint32_t root, uint32_t end) {
            const int32_t key = r[root];
            for (;;) {
                uint32_t child = 2 * root + 1;
                if (child >= end) break;
                if (child + 1 < end && r[child + 1] > r[child]) ++child;
                if (r[child] <= key) break;
                r[root] = r[child];
                root = child;
            }
            r[root] = key;
        };
        for (uint32_t a = d / 2; a-- > 0;) sift(a, d);
        for (uint32_t end = d - 1; end > 0; --end) {
            const int32_t top = r[0];
            r[0] = r[end];
            r[end] = top;
            sift(0, end);
        }
    }
    if (live) row[tid] = make_int2((int)off[tid], (int)d);
    if (live && tid % n == n - 1) atomicMax(gmax, off[tid] + d - off[tid - (n - 1)]);
    const uint32_t lo = __reduce_min_sync(0xffffffffu, live ? d : 0xFFFFFFFFu), hi = __reduce_max_sync(0xffffffffu, d);
    if ((threadIdx.x & 31) == 0) {
        atomicMin(dmin, lo);
        atomicMax(dmax, hi);
    }
}

// padded rows of the whole batch for the row-scanning CNTM kernel: see graph_build_kernel (api.cu)
__global__ void er_pad_kernel(int64_t rows, int64_t n, const int2* row, const int32_t* col, const uint32_t* off4, int2* row4,
                              int4* col4) {
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= rows) return;
    const int2 r = row[tid];
    int32_t* dst = reinterpret_cast<int32_t*>(col4 + off4[tid]);
    const int32_t padded = (r.y + 3) / 4 * 4;
    const int32_t self = (int32_t)(tid % n);
    for (int32_t k = 0; k < r.y; ++k) dst[k] = col[(uint32_t)r.x + k];
    for (int32_t k = r.y; k < padded; ++k) dst[k] = self;
    row4[tid] = make_int2((int)off4[tid], r.y);
}

// 64-byte agent records of the CNVM kernel (common.cuh, sponet_graph::d_rec)
__global__ void batch_record_kernel(int64_t rows, int64_t n, const int2* row, const int32_t* col, uint32_t* rec, int deg_bits) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // one thread per record word
    const int64_t a = t >> 4;
    const int w = (int)(t & 15);
    if (a >= rows) return;
    const int2 r = row[a];
    uint32_t v;
    if (w == 0) {
        const uint32_t first = (uint32_t)row[a - a % n].x;  // the graph's first adjacency entry
        v = (uint32_t)r.y | (((uint32_t)r.x - first) << deg_bits);
    } else {
        v = (w - 1 < r.y) ? (uint32_t)col[(uint32_t)r.x + (uint32_t)(w - 1)] : 0u;
    }
    rec[t] = v;
}

__global__ void quarter_kernel(int64_t rows, const uint32_t* deg, uint32_t* q) {
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid < rows) q[tid] = (deg[tid] + 3) / 4;
}

__global__ void sum_u32_kernel(int64_t rows, const uint32_t* v, unsigned long long* total) {
    unsigned long long acc = 0;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < rows; t += (int64_t)gridDim.x * blockDim.x) acc += v[t];
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0 && acc) atomicAdd(total, acc);
}

struct Freer {
    std::vector<void*> p;
    ~Freer() {
        cudaDeviceSynchronize();
        for (void* q : p) sp_pool_free(q);
    }
    cudaError_t alloc(void** out, size_t bytes) {
        cudaError_t e = sp_pool_alloc(out, bytes ? bytes : 16);
        if (e == cudaSuccess) p.push_back(*out);
        return e;
    }
};

// ---- Barabasi-Albert (networkx.barabasi_albert_graph, as sponet/network_generator.py:118-144 calls it) ----
// Star on the vertices 0 .. m; every later vertex t attaches to m DISTINCT targets drawn uniformly from the endpoint list of
// the edges so far (= proportional to degree), duplicates drawn again.  Edge e < m is (0, e + 1), edge e >= m is the
// (e % m)-th edge of vertex e / m + m; position 2 e of the endpoint list is the edge's source, 2 e + 1 its target.
// Candidate k of vertex t: position mulhi(word 0 of Philox(counter (k, t, c_lo, c_hi[16] | 5 << 24), key seed), 2 m (t - m)).
// A vertex depends on earlier vertices only (the endpoint-list formulation of Sanders & Schulz, "Scalable generation of
// scale-free graphs", 2016), which is what lets a warp resolve 32 vertices at a time.
struct BaArgs {
    int64_t n, count, nrun, R_total, run_begin;
    int32_t m;
    uint64_t seed;
};
constexpr int BA_MAX_M = 64;

__device__ __forceinline__ uint64_t ba_graph_index(const BaArgs& A, int64_t b) {
    const int64_t j = b / A.nrun, i = b - j * A.nrun;
    return (uint64_t)(j * A.R_total + A.run_begin + i);
}

// One warp per graph walks the vertices in chunks of 32 (one per lane).  Everything before the chunk is resolved; a lane whose
// candidate points at a vertex of its own chunk that is not resolved yet tries again after the others have written theirs (the
// lowest unresolved lane of a chunk depends on resolved vertices only, so every pass resolves at least one lane).
__global__ void ba_resolve_kernel(BaArgs A, int32_t* targets) {
    const int64_t b = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (b >= A.count) return;
    const int m = A.m;
    const uint64_t c = ba_graph_index(A, b);
    int32_t* T = targets + b * A.n * m;
    for (int64_t t0 = (int64_t)m + 1; t0 < A.n; t0 += 32) {
        const int64_t t = t0 + lane;
        bool done = t >= A.n;
        const uint32_t range = (uint32_t)(2 * (int64_t)m * (t - m));
        for (;;) {
            const uint32_t done_mask = __ballot_sync(0xffffffffu, done);
            if (done_mask == 0xffffffffu) break;
            if (!done) {
                int32_t tg[BA_MAX_M];
                int cnt = 0;
                bool wait = false;
                for (uint32_t k = 0; cnt < m && !wait; ++k) {
                    const uint4 r4 = sp_philox4x32_10(k, (uint32_t)t, (uint32_t)c, ((uint32_t)(c >> 32) & 0xFFFFu) | (5u << 24),
                                                      (uint32_t)A.seed, (uint32_t)(A.seed >> 32));
                    const uint32_t r = __umulhi(r4.x, range);
                    const uint32_t e = r >> 1;
                    int32_t v;
                    if (!(r & 1u)) {
                        v = e < (uint32_t)m ? 0 : (int32_t)(e / (uint32_t)m) + m;
                    } else if (e < (uint32_t)m) {
                        v = (int32_t)e + 1;
                    } else {
                        const int64_t owner = (int64_t)(e / (uint32_t)m) + m;
                        if (owner >= t0 && !((done_mask >> (owner - t0)) & 1u)) {
                            wait = true;  // a vertex of this chunk that is not resolved yet
                            break;
                        }
                        v = __ldcg(T + owner * m + (int64_t)(e % (uint32_t)m));  // (may have been written by another lane: not through L1)
                    }
                    bool dup = false;
                    for (int q = 0; q < cnt; ++q) dup |= (tg[q] == v);
                    if (!dup) tg[cnt++] = v;
                }
                if (!wait) {
                    for (int q = 0; q < m; ++q) T[t * m + q] = tg[q];
                    __threadfence();
                    done = true;
                }
            }
            __syncwarp();
        }
    }
}

// the edges of vertex t (the star's for t <= m): calls f(u, v) once per edge
template <typename F>
__device__ __forceinline__ void ba_edges(const BaArgs& A, const int32_t* targets, int64_t tid, int64_t t, F&& f) {
    if (t == 0) return;
    if (t <= A.m) {
        f(0, (int32_t)t);
        return;
    }
    for (int q = 0; q < A.m; ++q) f((int32_t)t, targets[tid * A.m + q]);
}

__global__ void ba_degree_kernel(BaArgs A, const int32_t* targets, uint32_t* deg) {
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= A.count * A.n) return;
    const int64_t b = tid / A.n, t = tid - b * A.n;
    uint32_t* d = deg + b * A.n;
    ba_edges(A, targets, tid, t, [&](int32_t u, int32_t v) {
        atomicAdd(d + u, 1u);
        atomicAdd(d + v, 1u);
    });
}

__global__ void ba_fill_kernel(BaArgs A, const int32_t* targets, const uint32_t* off, uint32_t* cursor, int32_t* col) {
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= A.count * A.n) return;
    const int64_t b = tid / A.n, t = tid - b * A.n;
    const uint32_t* o = off + b * A.n;
    uint32_t* cur = cursor + b * A.n;
    ba_edges(A, targets, tid, t, [&](int32_t u, int32_t v) {
        col[o[u] + atomicAdd(cur + u, 1u)] = v;
        col[o[v] + atomicAdd(cur + v, 1u)] = u;
    });
}

// Arrays a single warp reads and writes through the L2 (ld.cg / st.cg): what one lane wrote is what another lane reads after
// __syncwarp(), without the cost of volatile (system-scope) accesses.
template <typename T>
struct CgArray {
    T* p;
    struct Ref {
        T* q;
        __device__ __forceinline__ operator T() const { return __ldcg(q); }
        __device__ __forceinline__ const Ref& operator=(T v) const {
            __stcg(q, v);
            return *this;
        }
        __device__ __forceinline__ const Ref& operator=(const Ref& o) const {
            __stcg(q, (T)o);
            return *this;
        }
    };
    __device__ __forceinline__ Ref operator[](int64_t i) const { return Ref{p + i}; }
};

// ---- Watts-Strogatz (networkx.connected_watts_strogatz_graph, as sponet/network_generator.py:147-183 calls it) ----
// Ring lattice with k / 2 neighbours on each side, then the positions (j, u) in networkx's order (j = 1 .. k / 2 outer, u inner):
// with probability p the edge (u, u + j) is replaced by (u, w), w uniform, drawn again while w == u or {u, w} is an edge; the
// graph is drawn again until it is connected.  The rewiring is sequential by definition (a candidate is tested against the
// CURRENT edges), so one warp owns a graph: the lanes draw the decisions of 32 positions at a time and scan adjacency rows
// together, the rewirings are applied in position order.  Random numbers: see oracle_ws_graph (oracle/oracle_native.inc).
struct WsArgs {
    int64_t n, count, nrun, R_total, run_begin;
    int32_t k, tries, cap;
    uint32_t p_thr;
    uint64_t seed;
};

__global__ void __launch_bounds__(128) ws_build_kernel(WsArgs A, int32_t* adj_all, uint32_t* deg_all, int32_t* queue_all,
                                                       uint32_t* seen_all, int32_t* status) {
    __shared__ uint32_t tails[4];
    const int64_t b = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    if (b >= A.count) return;
    const int64_t n = A.n, cap = A.cap;
    const int half = A.k / 2;
    const CgArray<int32_t> adj{adj_all + b * n * cap};
    const CgArray<uint32_t> deg{deg_all + b * n};
    int32_t* queue = queue_all + b * n;
    uint32_t* seen = seen_all + b * ((n + 31) / 32);
    const int64_t jb = b / A.nrun, ib = b - jb * A.nrun;
    const uint64_t c = (uint64_t)(jb * A.R_total + A.run_begin + ib);
    int32_t result = -1;
    for (int attempt = 0; attempt < A.tries && result == -1; ++attempt) {
        const uint32_t c3 = ((uint32_t)(c >> 32) & 0xFFFFu) | (((uint32_t)attempt & 0xFFu) << 16) | (6u << 24);
        for (int64_t u = lane; u < n; u += 32) {  // the lattice
            for (int j = 1; j <= half; ++j) {
                adj[u * cap + 2 * (j - 1)] = (int32_t)((u + j) % n);
                adj[u * cap + 2 * (j - 1) + 1] = (int32_t)((u - j + n) % n);
            }
            deg[u] = (uint32_t)(2 * half);
        }
        __syncwarp();
        bool overflow = false;
        const int64_t positions = (int64_t)half * n;
        for (int64_t q0 = 0; q0 < positions && !overflow; q0 += 32) {
            const int64_t q = q0 + lane;
            bool flip = false;
            if (q < positions) {
                const uint4 r = sp_philox4x32_10(0u, (uint32_t)q, (uint32_t)c, c3, (uint32_t)A.seed, (uint32_t)(A.seed >> 32));
                flip = sp_thr_accept(r.x, A.p_thr);
            }
            uint32_t todo = __ballot_sync(0xffffffffu, flip);
            while (todo) {  // in position order; all lanes work on the same rewiring
                const int64_t qq = q0 + __ffs(todo) - 1;
                todo &= todo - 1;
                const int j = (int)(qq / n) + 1;
                const int64_t u = qq - (int64_t)(j - 1) * n, v = (u + j) % n;
                uint32_t du = deg[u];
                int64_t w = 0;
                bool skip = false;
                for (uint32_t d = 1;; ++d) {
                    const uint4 r = sp_philox4x32_10(d, (uint32_t)qq, (uint32_t)c, c3, (uint32_t)A.seed, (uint32_t)(A.seed >> 32));
                    w = (int64_t)__umulhi(r.x, (uint32_t)n);
                    if (d > 1 && (int64_t)du >= n - 1) {
                        skip = true;
                        break;
                    }
                    bool bad = false;
                    for (uint32_t e = lane; e < du; e += 32) bad |= adj[u * cap + e] == (int32_t)w;
                    if (!__any_sync(0xffffffffu, bad) && w != u) break;
                }
                if (skip) continue;
                // remove {u, v}: the entry moves to the end of its row
                for (int side = 0; side < 2; ++side) {
                    const int64_t a = side ? v : u;
                    const int32_t other = (int32_t)(side ? u : v);
                    const uint32_t da = deg[a];
                    uint32_t where = 0xFFFFFFFFu;
                    for (uint32_t e0 = 0; e0 < da && where == 0xFFFFFFFFu; e0 += 32) {
                        const uint32_t e = e0 + lane;
                        const uint32_t hit = __ballot_sync(0xffffffffu, e < da && adj[a * cap + e] == other);
                        if (hit) where = e0 + __ffs(hit) - 1;
                    }
                    if (lane == 0 && where != 0xFFFFFFFFu) {
                        adj[a * cap + where] = adj[a * cap + da - 1];
                        deg[a] = da - 1;
                    }
                    __syncwarp();
                }
                du = deg[u];
                const uint32_t dw = deg[w];
                if ((int64_t)du >= cap || (int64_t)dw >= cap) {
                    overflow = true;
                    break;
                }
                if (lane == 0) {
                    adj[u * cap + du] = (int32_t)w;
                    adj[w * cap + dw] = (int32_t)u;
                    deg[u] = du + 1;
                    deg[w] = dw + 1;
                }
                __syncwarp();
            }
        }
        if (overflow) {
            result = -3;
            break;
        }
        // connected?  breadth-first from vertex 0, 32 frontier vertices at a time
        for (int64_t i = lane; i < (n + 31) / 32; i += 32) seen[i] = 0u;
        __syncwarp();
        if (lane == 0) {
            seen[0] = 1u;
            queue[0] = 0;
            tails[wib] = 1u;
        }
        __syncwarp();
        uint32_t head = 0;
        for (;;) {
            const uint32_t tail = *(volatile uint32_t*)&tails[wib];
            if (head >= tail) break;
            const uint32_t idx = head + lane;
            if (idx < tail) {
                const int64_t a = __ldcg(queue + idx);
                const uint32_t da = deg[a];
                for (uint32_t e = 0; e < da; ++e) {
                    const int32_t nb = adj[a * cap + e];
                    const uint32_t bit = 1u << (nb & 31);
                    if (!(atomicOr(seen + (nb >> 5), bit) & bit)) queue[atomicAdd(&tails[wib], 1u)] = nb;
                }
            }
            head = min(head + 32u, tail);
            __threadfence_block();
            __syncwarp();
        }
        if ((int64_t)tails[wib] == n) result = attempt + 1;
        __syncwarp();
    }
    if (lane == 0) status[b] = result;
}

__global__ void ws_fill_kernel(int64_t rows, int64_t cap, const int32_t* adj, const uint32_t* deg, const uint32_t* off, int32_t* col) {
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= rows) return;
    const uint32_t d = deg[tid];
    for (uint32_t e = 0; e < d; ++e) col[off[tid] + e] = adj[tid * cap + e];
}

// ---- random regular graphs (networkx.random_regular_graph, as sponet/network_generator.py:87-115 calls it) ----
// The pairing process with restarts: see oracle_regular_graph (oracle/oracle_native.inc) for the definition.  Shuffle and
// pairing are sequential by definition; one warp owns a graph, its lanes draw the random numbers of 32 shuffle steps at a
// time and scan adjacency rows together.
struct RrArgs {
    int64_t n, count, nrun, R_total, run_begin;
    int32_t d, tries;
    uint64_t seed;
};

__global__ void __launch_bounds__(128) rr_build_kernel(RrArgs A, int32_t* stubs_all, int32_t* adj_all, uint32_t* deg_all,
                                                       uint32_t* pot_all, int32_t* status) {
    const int64_t b = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (b >= A.count) return;
    const int64_t n = A.n;
    const int d = A.d;
    const CgArray<int32_t> stubs{stubs_all + b * n * d}, adj{adj_all + b * n * d};
    const CgArray<uint32_t> deg{deg_all + b * n}, pot{pot_all + b * n};
    const int64_t jb = b / A.nrun, ib = b - jb * A.nrun;
    const uint64_t c = (uint64_t)(jb * A.R_total + A.run_begin + ib);
    const uint32_t c3 = ((uint32_t)(c >> 32) & 0xFFFFu) | (7u << 24);
    int32_t result = -1;
    for (int attempt = 0; attempt < A.tries && result == -1; ++attempt) {
        int64_t L = n * d;
        for (int64_t q = lane; q < L; q += 32) stubs[q] = (int32_t)(q % n);
        for (int64_t u = lane; u < n; u += 32) deg[u] = 0u;
        __syncwarp();
        bool failed = false;
        for (uint32_t pass = 0; L > 0 && !failed; ++pass) {
            const uint32_t c1 = (pass & 0xFFFFu) | ((uint32_t)attempt << 16);
            // Fisher-Yates, i = L - 1 .. 1, 32 steps at a time: the lanes draw their targets; steps that touch 64 different
            // positions commute and are done by all lanes at once, otherwise in order by lane 0
            for (int64_t i0 = L - 1; i0 >= 1; i0 -= 32) {
                const int64_t i = i0 - lane;
                const bool live = i >= 1;
                uint32_t j = 0;
                if (live) {
                    const uint4 r = sp_philox4x32_10((uint32_t)i, c1, (uint32_t)c, c3, (uint32_t)A.seed, (uint32_t)(A.seed >> 32));
                    j = __umulhi(r.x, (uint32_t)(i + 1));
                }
                // lane's step touches i and j (j <= i); another lane's i' lies in (i0 - 32, i0]
                bool clash = live && (int64_t)j != i && (int64_t)j > i0 - 32 && (int64_t)j >= 1;
                const uint32_t same = __match_any_sync(0xffffffffu, live ? j : 0xFFFFFFFFu - (uint32_t)lane);
                clash |= live && (same & ~(1u << lane)) != 0u;
                if (!__any_sync(0xffffffffu, clash)) {
                    if (live && (int64_t)j != i) {
                        const int32_t x = stubs[i], y = stubs[j];
                        stubs[i] = y;
                        stubs[j] = x;
                    }
                } else {
                    const int steps = (int)(i0 < 32 ? i0 : 32);
                    for (int t = 0; t < steps; ++t) {
                        const uint32_t jt = __shfl_sync(0xffffffffu, j, t);
                        if (lane == 0) {
                            const int64_t it = i0 - t;
                            const int32_t x = stubs[it];
                            stubs[it] = stubs[jt];
                            stubs[jt] = x;
                        }
                    }
                }
                __syncwarp();
            }
            for (int64_t u = lane; u < n; u += 32) pot[u] = 0u;
            __syncwarp();
            uint32_t npot = 0;
            // the pairs in order, 32 at a time: pairs over 64 different vertices commute (one lane each), otherwise the
            // chunk is done pair by pair with all lanes scanning the row
            for (int64_t q0 = 0; q0 + 1 < L; q0 += 64) {
                const int64_t q = q0 + 2 * lane;
                const bool live = q + 1 < L;
                const int32_t m1 = live ? stubs[q] : -1 - 2 * lane, m2 = live ? stubs[q + 1] : -2 - 2 * lane;
                bool clash = false;
                for (int t = 0; t < 32; ++t) {
                    const int32_t a1 = __shfl_sync(0xffffffffu, m1, t), a2 = __shfl_sync(0xffffffffu, m2, t);
                    clash |= t != lane && (a1 == m1 || a1 == m2 || a2 == m1 || a2 == m2);
                }
                if (!__any_sync(0xffffffffu, clash)) {
                    bool bad = false;
                    if (live) {
                        const uint32_t d1 = deg[m1];
                        bad = m1 == m2;
                        for (uint32_t e = 0; e < d1; ++e) bad |= adj[(int64_t)m1 * d + e] == m2;
                        if (!bad) {
                            const uint32_t d2 = deg[m2];
                            adj[(int64_t)m1 * d + d1] = m2;
                            adj[(int64_t)m2 * d + d2] = m1;
                            deg[m1] = d1 + 1;
                            deg[m2] = d2 + 1;
                        } else if (m1 == m2) {
                            pot[m1] = pot[m1] + 2u;
                        } else {
                            pot[m1] = pot[m1] + 1u;
                            pot[m2] = pot[m2] + 1u;
                        }
                    }
                    npot += 2u * (uint32_t)__popc(__ballot_sync(0xffffffffu, bad));
                } else {
                    const int pairs = (int)(((L - q0) / 2) < 32 ? ((L - q0) / 2) : 32);
                    for (int t = 0; t < pairs; ++t) {
                        const int32_t s1 = __shfl_sync(0xffffffffu, m1, t), s2 = __shfl_sync(0xffffffffu, m2, t);
                        const uint32_t d1 = deg[s1];
                        bool hit = false;
                        for (uint32_t e = lane; e < d1; e += 32) hit |= adj[(int64_t)s1 * d + e] == s2;
                        const bool bad = __any_sync(0xffffffffu, hit) || s1 == s2;
                        if (lane == 0) {
                            if (!bad) {
                                const uint32_t d2 = deg[s2];
                                adj[(int64_t)s1 * d + d1] = s2;
                                adj[(int64_t)s2 * d + d2] = s1;
                                deg[s1] = d1 + 1;
                                deg[s2] = d2 + 1;
                            } else {
                                pot[s1] = pot[s1] + 1u;
                                pot[s2] = pot[s2] + 1u;
                            }
                        }
                        if (bad) npot += 2;
                        __syncwarp();
                    }
                }
                __syncwarp();
            }
            if (npot == 0) {
                L = 0;
                break;
            }
            // the vertices with potentials, in vertex order, back into stubs[0 .. npot) -- and networkx's _suitable on them
            int64_t L2 = 0;
            for (int64_t u0 = 0; u0 < n; u0 += 32) {
                const int64_t u = u0 + lane;
                const uint32_t pu = u < n ? pot[u] : 0u;
                uint32_t have = __ballot_sync(0xffffffffu, pu != 0u);
                while (have) {  // (few vertices: sequential)
                    const int src = __ffs(have) - 1;
                    have &= have - 1;
                    const uint32_t cnt = __shfl_sync(0xffffffffu, pu, src);
                    if (lane == 0)
                        for (uint32_t t = 0; t < cnt; ++t) stubs[L2 + t] = (int32_t)(u0 + src);
                    L2 += cnt;
                }
            }
            __syncwarp();
            bool suitable = false;
            for (int64_t x = 0; x < L2 && !suitable; ++x) {
                const int32_t u = stubs[x];
                if (x > 0 && stubs[x - 1] == u) continue;
                const uint32_t du = deg[u];
                for (int64_t y = x + 1; y < L2 && !suitable; ++y) {
                    const int32_t v = stubs[y];
                    if (v == u) continue;
                    bool hit = false;
                    for (uint32_t e = lane; e < du; e += 32) hit |= adj[(int64_t)u * d + e] == v;
                    suitable = !__any_sync(0xffffffffu, hit);
                }
            }
            if (!suitable) {
                failed = true;
                break;
            }
            L = L2;
        }
        if (!failed) result = attempt + 1;
    }
    if (lane == 0) status[b] = result;
}

// ---- degree-weighted agent selection on a batch (alpha != 1) ----
// One thread per graph: the sum of the weights in agent order (the reference's np.sum inside njit is a sequential loop,
// cnvm/model.py:257) and the stack-based alias construction of sponet/sampling.py:54-96 -- the same statements as
// sponet_build_alias_table (api.cu), whose tables the oracle and the single-graph path use.
__global__ void batch_alias_kernel(int64_t B, int64_t n, const int2* row, const double* weight_of_degree, double* prob_all,
                                   int32_t* small_all, int32_t* large_all, uint2* alias_all, double* weight_sum) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const int2* r = row + b * n;
    double* prob = prob_all + b * n;
    int32_t *small_ids = small_all + b * n, *large_ids = large_all + b * n;
    uint2* out = alias_all + b * n;
    double total = 0.0;
    for (int64_t i = 0; i < n; ++i) total = __dadd_rn(total, weight_of_degree[r[i].y]);
    weight_sum[b] = total;
    int64_t ns = 0, nl = 0;
    for (int64_t i = 0; i < n; ++i) {
        const double pi = __dmul_rn(__ddiv_rn(weight_of_degree[r[i].y], total), (double)n);
        prob[i] = pi;
        out[i].y = 0u;
        if (pi < 1) small_ids[ns++] = (int32_t)i;
        else large_ids[nl++] = (int32_t)i;
    }
    while (ns > 0 && nl > 0) {
        const int32_t sm = small_ids[--ns], lg = large_ids[--nl];
        out[sm].y = (uint32_t)lg;
        const double pl = __dsub_rn(__dadd_rn(prob[lg], prob[sm]), 1.0);
        prob[lg] = pl;
        if (pl < 1) small_ids[ns++] = lg;
        else large_ids[nl++] = lg;
    }
    for (int64_t q = 0; q < nl; ++q) prob[large_ids[q]] = 1.0;
    for (int64_t q = 0; q < ns; ++q) prob[small_ids[q]] = 1.0;
    for (int64_t i = 0; i < n; ++i) out[i].x = sp_prob_to_thr(prob[i]);
}

// Common tail of the batch generators: degrees -> absolute row offsets, graph handle and adjacency array; after the
// generator's fill kernels: rows sorted, (begin, degree) descriptors, degree range, and the layouts of the simulation kernels.
struct BatchBuilder {
    Freer& tmp;
    int64_t n, B, rows;
    const char* who;
    uint32_t *d_off = nullptr, *d_cursor = nullptr;
    uint64_t* d_total = nullptr;
    sponet_graph* g = nullptr;
    BatchBuilder(Freer& t, int64_t n_, int64_t B_, const char* w) : tmp(t), n(n_), B(B_), rows(n_ * B_), who(w) {}
    ~BatchBuilder() {
        if (g) sponet_graph_destroy(g);
    }
    sponet_graph* release() {
        sponet_graph* r = g;
        g = nullptr;
        return r;
    }
    int scan(const uint32_t* d_deg) {
        SP_CUDA(tmp.alloc((void**)&d_total, 8));
        SP_CUDA(cudaMemset(d_total, 0, 8));
        sum_u32_kernel<<<1184, 256>>>(rows, d_deg, (unsigned long long*)d_total);
        SP_CUDA(cudaGetLastError());
        uint64_t nnz = 0;
        SP_CUDA(cudaMemcpy(&nnz, d_total, 8, cudaMemcpyDeviceToHost));
        SP_REQUIRE(nnz < ((uint64_t)1 << 32), "%s: %llu adjacency entries exceed 2^32; use smaller batches", who, (unsigned long long)nnz);
        SP_CUDA(tmp.alloc((void**)&d_off, rows * 4));
        SP_CUDA(tmp.alloc((void**)&d_cursor, rows * 4));
        SP_CUDA(cudaMemset(d_cursor, 0, rows * 4));
        void* d_ws = nullptr;
        size_t ws = 0;
        SP_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, ws, d_deg, d_off, rows));
        SP_CUDA(tmp.alloc(&d_ws, ws));
        SP_CUDA(cub::DeviceScan::ExclusiveSum(d_ws, ws, d_deg, d_off, rows));
        g = new sponet_graph();
        memset(g, 0, sizeof(*g));
        g->batch = B;
        g->pooled = true;
        g->n = n;
        g->nnz = (int64_t)nnz;
        SP_CUDA(cudaGetDevice(&g->device));
        SP_CUDA(sp_pool_alloc((void**)&g->d_row, rows * sizeof(int2)));
        SP_CUDA(sp_pool_alloc((void**)&g->d_col, std::max<uint64_t>(nnz, 1) * sizeof(int32_t)));
        return SPONET_OK;
    }
    int finish(const uint32_t* d_deg, int layouts) {
        const unsigned grid = (unsigned)((rows + 255) / 256);
        SP_CUDA(cudaGetLastError());  // (the generator's fill kernels)
        uint32_t* d_mm;
        SP_CUDA(tmp.alloc((void**)&d_mm, 16));
        uint32_t mm[3] = {0xFFFFFFFFu, 0u, 0u};
        SP_CUDA(cudaMemcpy(d_mm, mm, 12, cudaMemcpyHostToDevice));
        batch_finish_kernel<<<grid, 256>>>(rows, n, d_off, d_deg, g->d_col, g->d_row, d_mm, d_mm + 1, d_mm + 2);
        SP_CUDA(cudaGetLastError());
        SP_CUDA(cudaMemcpy(mm, d_mm, 12, cudaMemcpyDeviceToHost));
        g->min_degree = (int32_t)mm[0];
        g->max_degree = (int32_t)mm[1];
        // record header: the degree in the low bits, the row's offset within its graph above
        int deg_bits = 8;
        while (deg_bits < 31 && ((uint32_t)1 << deg_bits) <= mm[1]) ++deg_bits;
        if ((layouts & SPONET_BATCH_RECORDS) && deg_bits <= 16 && (uint64_t)mm[2] < ((uint64_t)1 << (32 - deg_bits)) &&
            (uint64_t)rows * 16 < ((uint64_t)1 << 32)) {
            g->rec_deg_bits = deg_bits;
            SP_CUDA(sp_pool_alloc((void**)&g->d_rec, (size_t)rows * 64));
            batch_record_kernel<<<(unsigned)((rows * 16 + 255) / 256), 256>>>(rows, n, g->d_row, g->d_col, g->d_rec, deg_bits);
            SP_CUDA(cudaGetLastError());
        }
        if (layouts & SPONET_BATCH_PADDED_ROWS) {  // the CNTM kernels' layout
            uint32_t *d_q, *d_off4;
            uint64_t n4 = 0;
            SP_CUDA(tmp.alloc((void**)&d_q, rows * 4));
            SP_CUDA(tmp.alloc((void**)&d_off4, rows * 4));
            quarter_kernel<<<grid, 256>>>(rows, d_deg, d_q);
            SP_CUDA(cudaMemset(d_total, 0, 8));
            sum_u32_kernel<<<1184, 256>>>(rows, d_q, (unsigned long long*)d_total);
            SP_CUDA(cudaMemcpy(&n4, d_total, 8, cudaMemcpyDeviceToHost));
            SP_REQUIRE(n4 < ((uint64_t)1 << 31), "%s: padded rows exceed 2^31 entries; use smaller batches", who);
            void* d_ws2 = nullptr;
            size_t ws2 = 0;
            SP_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, ws2, d_q, d_off4, rows));
            SP_CUDA(tmp.alloc(&d_ws2, ws2));
            SP_CUDA(cub::DeviceScan::ExclusiveSum(d_ws2, ws2, d_q, d_off4, rows));
            SP_CUDA(sp_pool_alloc((void**)&g->d_row4, rows * sizeof(int2)));
            SP_CUDA(sp_pool_alloc((void**)&g->d_col4, std::max<uint64_t>(n4, 1) * sizeof(int4)));
            g->n_col4 = (int64_t)n4;
            er_pad_kernel<<<grid, 256>>>(rows, n, g->d_row, g->d_col, d_off4, g->d_row4, g->d_col4);
            SP_CUDA(cudaGetLastError());
        }
        SP_CUDA(cudaDeviceSynchronize());
        return SPONET_OK;
    }
};

}  // namespace

extern "C" int sponet_graph_batch_create_er(int64_t num_agents, double p, int force_no_isolates, uint64_t seed,
                                            int64_t num_states, int64_t num_runs_total, int64_t run_begin,
                                            int64_t run_end, int32_t max_attempts, int layouts,
                                            sponet_graph** out) {
    SP_REQUIRE(out, "graph_batch_create_er: out is NULL");
    *out = nullptr;
    const int64_t n = num_agents, nrun = run_end - run_begin, B = num_states * nrun;
    SP_REQUIRE(n >= 2 && n < ((int64_t)1 << 31), "graph_batch_create_er: num_agents out of range");
    SP_REQUIRE(p > 0.0 && p <= 1.0, "graph_batch_create_er: p must be in (0, 1]");
    SP_REQUIRE(num_states >= 1 && nrun >= 1 && run_begin >= 0 && num_runs_total >= 1, "graph_batch_create_er: bad run range");
    SP_REQUIRE(B * n < ((int64_t)1 << 32), "graph_batch_create_er: %lld graphs of %lld agents exceed 2^32 rows; use smaller batches",
               (long long)B, (long long)n);
    SP_REQUIRE(max_attempts >= 1 && max_attempts <= 256, "graph_batch_create_er: max_attempts must be in [1, 256]");
    ErArgs A;
    A.n = n;
    A.count = B;
    A.nrun = nrun;
    A.R_total = num_runs_total;
    A.run_begin = run_begin;
    A.log1mp = p >= 1.0 ? 0.0 : log1p(-p);
    A.seed = seed;
    const int64_t rows = B * n;
    const unsigned grid = (unsigned)((rows + 255) / 256);

    Freer tmp;
    uint32_t *d_deg, *d_attempt, *d_iso;
    uint8_t* d_dirty;
    SP_CUDA(tmp.alloc((void**)&d_deg, rows * 4));
    SP_CUDA(tmp.alloc((void**)&d_attempt, B * 4));
    SP_CUDA(tmp.alloc((void**)&d_iso, B * 4));
    SP_CUDA(tmp.alloc((void**)&d_dirty, B));
    SP_CUDA(cudaMemset(d_deg, 0, rows * 4));
    SP_CUDA(cudaMemset(d_attempt, 0, B * 4));
    std::vector<uint8_t> dirty((size_t)B, 1);
    std::vector<uint32_t> attempt((size_t)B, 0), iso((size_t)B);
    uint32_t max_iso = 0;
    for (;;) {
        SP_CUDA(cudaMemcpy(d_dirty, dirty.data(), B, cudaMemcpyHostToDevice));
        SP_CUDA(cudaMemset(d_iso, 0, B * 4));
        er_degree_kernel<<<grid, 256>>>(A, d_dirty, d_attempt, d_deg);
        er_isolates_kernel<<<grid, 256>>>(A, d_dirty, d_deg, d_iso);
        SP_CUDA(cudaGetLastError());
        SP_CUDA(cudaMemcpy(iso.data(), d_iso, B * 4, cudaMemcpyDeviceToHost));
        bool any = false;
        for (int64_t b = 0; b < B; ++b) {
            if (!dirty[b]) continue;
            if (iso[b] == 0 || force_no_isolates) {
                dirty[b] = 0;
                max_iso = std::max(max_iso, iso[b]);
            } else {
                if ((int32_t)attempt[b] + 1 >= max_attempts)  // reference :75-78 (a time limit there)
                    return sponet_fail(SPONET_ERR_INVALID, "Timeout. Could not generate a graph without isolated vertices.");
                ++attempt[b];
                any = true;
            }
        }
        if (!any) break;
        SP_CUDA(cudaMemcpy(d_dirty, dirty.data(), B, cudaMemcpyHostToDevice));
        SP_CUDA(cudaMemcpy(d_attempt, attempt.data(), B * 4, cudaMemcpyHostToDevice));
        er_reset_kernel<<<grid, 256>>>(A, d_dirty, d_deg);
    }
    int2* d_patch = nullptr;
    uint32_t* d_npatch = nullptr;
    const int64_t cap = max_iso;
    if (cap > 0) {
        SP_CUDA(tmp.alloc((void**)&d_patch, (size_t)B * cap * sizeof(int2)));
        SP_CUDA(tmp.alloc((void**)&d_npatch, B * 4));
        er_patch_kernel<<<(unsigned)((B * 32 + 255) / 256), 256>>>(A, d_deg, d_patch, d_npatch, cap);
        SP_CUDA(cudaGetLastError());
    }
    BatchBuilder bb(tmp, n, B, "graph_batch_create_er");
    SP_TRY(bb.scan(d_deg));
    er_fill_kernel<<<grid, 256>>>(A, d_attempt, bb.d_off, bb.d_cursor, bb.g->d_col);
    if (cap > 0)
        er_fill_patch_kernel<<<(unsigned)((B * cap + 255) / 256), 256>>>(A, d_patch, d_npatch, cap, bb.d_off, bb.d_cursor, bb.g->d_col);
    SP_TRY(bb.finish(d_deg, layouts));
    *out = bb.release();
    return SPONET_OK;
}

extern "C" int sponet_graph_batch_create_ba(int64_t num_agents, int32_t m, uint64_t seed, int64_t num_states,
                                            int64_t num_runs_total, int64_t run_begin, int64_t run_end, int layouts,
                                            sponet_graph** out) {
    SP_REQUIRE(out, "graph_batch_create_ba: out is NULL");
    *out = nullptr;
    const int64_t n = num_agents, nrun = run_end - run_begin, B = num_states * nrun;
    SP_REQUIRE(n >= 2 && n < ((int64_t)1 << 31), "graph_batch_create_ba: num_agents out of range");
    SP_REQUIRE(m >= 1 && m < n, "Barabasi-Albert network must have m >= 1 and m < n, m = %d, n = %lld", m, (long long)n);  // networkx's message
    SP_REQUIRE(m <= BA_MAX_M, "graph_batch_create_ba: m <= %d", BA_MAX_M);
    SP_REQUIRE(num_states >= 1 && nrun >= 1 && run_begin >= 0 && num_runs_total >= 1, "graph_batch_create_ba: bad run range");
    SP_REQUIRE(B * n < ((int64_t)1 << 32), "graph_batch_create_ba: %lld graphs of %lld agents exceed 2^32 rows; use smaller batches",
               (long long)B, (long long)n);
    SP_REQUIRE(2 * (int64_t)m * n < ((int64_t)1 << 32), "graph_batch_create_ba: too many edges per graph");
    BaArgs A;
    A.n = n;
    A.count = B;
    A.nrun = nrun;
    A.R_total = num_runs_total;
    A.run_begin = run_begin;
    A.m = m;
    A.seed = seed;
    const int64_t rows = B * n;
    const unsigned grid = (unsigned)((rows + 255) / 256);
    Freer tmp;
    int32_t* d_targets;
    uint32_t* d_deg;
    SP_CUDA(tmp.alloc((void**)&d_targets, (size_t)rows * m * sizeof(int32_t)));
    SP_CUDA(tmp.alloc((void**)&d_deg, rows * 4));
    SP_CUDA(cudaMemset(d_deg, 0, rows * 4));
    ba_resolve_kernel<<<(unsigned)((B * 32 + 127) / 128), 128>>>(A, d_targets);
    SP_CUDA(cudaGetLastError());
    ba_degree_kernel<<<grid, 256>>>(A, d_targets, d_deg);
    SP_CUDA(cudaGetLastError());
    BatchBuilder bb(tmp, n, B, "graph_batch_create_ba");
    SP_TRY(bb.scan(d_deg));
    ba_fill_kernel<<<grid, 256>>>(A, d_targets, bb.d_off, bb.d_cursor, bb.g->d_col);
    SP_TRY(bb.finish(d_deg, layouts));
    *out = bb.release();
    return SPONET_OK;
}

extern "C" int sponet_graph_batch_create_ws(int64_t num_agents, int32_t k, double p, uint64_t seed, int64_t num_states,
                                            int64_t num_runs_total, int64_t run_begin, int64_t run_end, int32_t tries, int layouts,
                                            sponet_graph** out) {
    SP_REQUIRE(out, "graph_batch_create_ws: out is NULL");
    *out = nullptr;
    const int64_t n = num_agents, nrun = run_end - run_begin, B = num_states * nrun;
    SP_REQUIRE(n >= 3 && n < ((int64_t)1 << 31), "graph_batch_create_ws: num_agents out of range");
    SP_REQUIRE(k >= 2 && k < n, "graph_batch_create_ws: 2 <= k < n required (k = %d, n = %lld)", k, (long long)n);
    SP_REQUIRE(p >= 0.0 && p <= 1.0, "graph_batch_create_ws: p must be in [0, 1]");
    SP_REQUIRE(tries >= 1 && tries <= 256, "graph_batch_create_ws: tries must be in [1, 256]");
    SP_REQUIRE(num_states >= 1 && nrun >= 1 && run_begin >= 0 && num_runs_total >= 1, "graph_batch_create_ws: bad run range");
    SP_REQUIRE(B * n < ((int64_t)1 << 32), "graph_batch_create_ws: %lld graphs of %lld agents exceed 2^32 rows; use smaller batches",
               (long long)B, (long long)n);
    SP_REQUIRE((int64_t)(k / 2) * n < ((int64_t)1 << 32), "graph_batch_create_ws: too many lattice edges per graph");
    WsArgs A;
    A.n = n;
    A.count = B;
    A.nrun = nrun;
    A.R_total = num_runs_total;
    A.run_begin = run_begin;
    A.k = k;
    A.tries = tries;
    A.cap = k + 32;  // k lattice neighbours + rewired-in edges (oracle.ws_row_capacity)
    A.p_thr = sp_prob_to_thr(p);
    A.seed = seed;
    const int64_t rows = B * n;
    const unsigned grid = (unsigned)((rows + 255) / 256);
    Freer tmp;
    int32_t *d_adj, *d_queue, *d_status;
    uint32_t *d_deg, *d_seen;
    SP_CUDA(tmp.alloc((void**)&d_adj, (size_t)rows * A.cap * sizeof(int32_t)));
    SP_CUDA(tmp.alloc((void**)&d_deg, rows * 4));
    SP_CUDA(tmp.alloc((void**)&d_queue, rows * 4));
    SP_CUDA(tmp.alloc((void**)&d_seen, (size_t)B * ((n + 31) / 32) * 4));
    SP_CUDA(tmp.alloc((void**)&d_status, B * 4));
    ws_build_kernel<<<(unsigned)((B * 32 + 127) / 128), 128>>>(A, d_adj, d_deg, d_queue, d_seen, d_status);
    SP_CUDA(cudaGetLastError());
    std::vector<int32_t> status((size_t)B);
    SP_CUDA(cudaMemcpy(status.data(), d_status, B * 4, cudaMemcpyDeviceToHost));
    for (int64_t b = 0; b < B; ++b) {
        if (status[b] == -1) return sponet_fail(SPONET_ERR_INVALID, "Maximum number of tries exceeded");  // networkx's message
        SP_REQUIRE(status[b] > 0, "graph_batch_create_ws: a vertex has more than %d neighbours", A.cap);
    }
    BatchBuilder bb(tmp, n, B, "graph_batch_create_ws");
    SP_TRY(bb.scan(d_deg));
    ws_fill_kernel<<<grid, 256>>>(rows, A.cap, d_adj, d_deg, bb.d_off, bb.g->d_col);
    SP_TRY(bb.finish(d_deg, layouts));
    *out = bb.release();
    return SPONET_OK;
}

extern "C" int sponet_graph_batch_create_regular(int64_t num_agents, int32_t d, uint64_t seed, int64_t num_states,
                                                 int64_t num_runs_total, int64_t run_begin, int64_t run_end, int32_t tries,
                                                 int layouts, sponet_graph** out) {
    SP_REQUIRE(out, "graph_batch_create_regular: out is NULL");
    *out = nullptr;
    const int64_t n = num_agents, nrun = run_end - run_begin, B = num_states * nrun;
    SP_REQUIRE(n >= 2 && n < ((int64_t)1 << 31), "graph_batch_create_regular: num_agents out of range");
    SP_REQUIRE((n * d) % 2 == 0, "n * d must be even");                                  // networkx's messages
    SP_REQUIRE(d >= 1 && d < n, "the 0 <= d < n inequality must be satisfied (and d >= 1: no isolated vertices)");
    SP_REQUIRE(tries >= 1 && tries <= 32767, "graph_batch_create_regular: tries must be in [1, 32767]");
    SP_REQUIRE(num_states >= 1 && nrun >= 1 && run_begin >= 0 && num_runs_total >= 1, "graph_batch_create_regular: bad run range");
    SP_REQUIRE(B * n < ((int64_t)1 << 32), "graph_batch_create_regular: %lld graphs of %lld agents exceed 2^32 rows; use smaller batches",
               (long long)B, (long long)n);
    SP_REQUIRE(n * d < ((int64_t)1 << 32) - 64, "graph_batch_create_regular: too many stubs per graph");
    RrArgs A;
    A.n = n;
    A.count = B;
    A.nrun = nrun;
    A.R_total = num_runs_total;
    A.run_begin = run_begin;
    A.d = d;
    A.tries = tries;
    A.seed = seed;
    const int64_t rows = B * n;
    const unsigned grid = (unsigned)((rows + 255) / 256);
    Freer tmp;
    int32_t *d_stubs, *d_adj, *d_status;
    uint32_t *d_deg, *d_pot;
    SP_CUDA(tmp.alloc((void**)&d_stubs, (size_t)rows * d * sizeof(int32_t)));
    SP_CUDA(tmp.alloc((void**)&d_adj, (size_t)rows * d * sizeof(int32_t)));
    SP_CUDA(tmp.alloc((void**)&d_deg, rows * 4));
    SP_CUDA(tmp.alloc((void**)&d_pot, rows * 4));
    SP_CUDA(tmp.alloc((void**)&d_status, B * 4));
    rr_build_kernel<<<(unsigned)((B * 32 + 127) / 128), 128>>>(A, d_stubs, d_adj, d_deg, d_pot, d_status);
    SP_CUDA(cudaGetLastError());
    std::vector<int32_t> status((size_t)B);
    SP_CUDA(cudaMemcpy(status.data(), d_status, B * 4, cudaMemcpyDeviceToHost));
    for (int64_t b = 0; b < B; ++b)
        SP_REQUIRE(status[b] > 0, "graph_batch_create_regular: no regular graph after %d attempts", tries);
    BatchBuilder bb(tmp, n, B, "graph_batch_create_regular");
    SP_TRY(bb.scan(d_deg));
    ws_fill_kernel<<<grid, 256>>>(rows, d, d_adj, d_deg, bb.d_off, bb.g->d_col);
    SP_TRY(bb.finish(d_deg, layouts));
    *out = bb.release();
    return SPONET_OK;
}

extern "C" int sponet_graph_batch_set_weights(sponet_graph* g, const double* weight_of_degree, int64_t num_weights) {
    SP_REQUIRE(g && weight_of_degree, "graph_batch_set_weights: bad arguments");
    SP_REQUIRE(g->pooled && g->batch >= 1, "graph_batch_set_weights: not a graph batch");
    SP_REQUIRE(num_weights > g->max_degree, "graph_batch_set_weights: need a weight for every degree 0 .. %d", g->max_degree);
    int cur = 0;
    SP_CUDA(cudaGetDevice(&cur));
    SP_REQUIRE(cur == g->device, "graph batch lives on device %d, current device is %d", g->device, cur);
    const int64_t B = g->batch, n = g->n, rows = B * n;
    Freer tmp;
    double *d_w, *d_prob, *d_sum;
    int32_t *d_small, *d_large;
    SP_CUDA(tmp.alloc((void**)&d_w, (size_t)num_weights * sizeof(double)));
    SP_CUDA(tmp.alloc((void**)&d_prob, (size_t)rows * sizeof(double)));
    SP_CUDA(tmp.alloc((void**)&d_small, (size_t)rows * sizeof(int32_t)));
    SP_CUDA(tmp.alloc((void**)&d_large, (size_t)rows * sizeof(int32_t)));
    SP_CUDA(tmp.alloc((void**)&d_sum, (size_t)B * sizeof(double)));
    SP_CUDA(cudaMemcpy(d_w, weight_of_degree, (size_t)num_weights * sizeof(double), cudaMemcpyHostToDevice));
    if (!g->d_alias) SP_CUDA(sp_pool_alloc((void**)&g->d_alias, (size_t)rows * sizeof(uint2)));
    batch_alias_kernel<<<(unsigned)((B + 31) / 32), 32>>>(B, n, g->d_row, d_w, d_prob, d_small, d_large, g->d_alias, d_sum);
    SP_CUDA(cudaGetLastError());
    if (!g->h_weight_sum) g->h_weight_sum = new double[(size_t)B];
    SP_CUDA(cudaMemcpy(g->h_weight_sum, d_sum, (size_t)B * sizeof(double), cudaMemcpyDeviceToHost));
    return SPONET_OK;
}

extern "C" int sponet_graph_batch_get(const sponet_graph* g, int64_t slot, int32_t* row_ptr, int32_t* col, int64_t col_capacity) {
    SP_REQUIRE(g && row_ptr, "graph_batch_get: bad arguments");
    SP_REQUIRE(slot >= 0 && slot < g->batch, "graph_batch_get: slot %lld outside the batch of %lld", (long long)slot, (long long)g->batch);
    int cur = 0;
    SP_CUDA(cudaGetDevice(&cur));
    if (cur != g->device) SP_CUDA(cudaSetDevice(g->device));
    std::vector<int2> row((size_t)g->n);
    cudaError_t e = cudaMemcpy(row.data(), g->d_row + slot * g->n, (size_t)g->n * sizeof(int2), cudaMemcpyDeviceToHost);
    int64_t total = 0;
    if (e == cudaSuccess) {
        for (int64_t i = 0; i < g->n; ++i) {
            row_ptr[i] = (int32_t)total;
            total += row[i].y;
        }
        row_ptr[g->n] = (int32_t)total;
        if (col && total <= col_capacity && total > 0)  // rows of one graph are contiguous in the batch
            e = cudaMemcpy(col, g->d_col + (uint32_t)row[0].x, (size_t)total * sizeof(int32_t), cudaMemcpyDeviceToHost);
    }
    if (cur != g->device) cudaSetDevice(cur);
    SP_CUDA(e);
    SP_REQUIRE(!col || total <= col_capacity, "graph_batch_get: graph has %lld entries, capacity %lld", (long long)total, (long long)col_capacity);
    return SPONET_OK;
}
