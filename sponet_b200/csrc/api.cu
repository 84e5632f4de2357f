// C-ABI plumbing of libsponet_b200: error slot, device query, graph handle, and the host-side setup
// arithmetic that has to stay bit-compatible with the reference's numba code.
#include <algorithm>

#include "common.cuh"

std::string& sponet_err_slot() {
    static thread_local std::string slot;
    return slot;
}

int sponet_fail(int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    sponet_err_slot() = buf;
    return code;
}

extern "C" const char* sponet_last_error(void) { return sponet_err_slot().c_str(); }
extern "C" int sponet_abi_version(void) { return SPONET_ABI_VERSION; }

extern "C" int sponet_device_count(int* count) {
    SP_REQUIRE(count, "count is NULL");
    *count = 0;
    cudaError_t e = cudaGetDeviceCount(count);
    if (e != cudaSuccess || *count == 0) {
        cudaGetLastError();
        *count = 0;
        return sponet_fail(SPONET_ERR_CUDA, "no CUDA device available (%s); sponet_b200 has no CPU fallback",
                           e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    }
    return SPONET_OK;
}

// sponet/sampling.py:54-96.  The two python lists are used as stacks (append/pop at the end), and
// the resulting tables depend on that order, so it is reproduced literally.  np.sum inside njit is a
// sequential left-to-right loop (numpy's pairwise sum would differ in the last bits).
extern "C" int sponet_build_alias_table(const double* weights, int64_t n, double* table_prob,
                                        int32_t* table_alias) {
    SP_REQUIRE(weights && table_prob && table_alias && n >= 1, "build_alias_table: bad arguments");
    double total = 0.0;
    for (int64_t i = 0; i < n; ++i) total += weights[i];
    std::vector<int64_t> small_ids, large_ids;
    small_ids.reserve(n);
    large_ids.reserve(n);
    for (int64_t i = 0; i < n; ++i) {
        table_prob[i] = weights[i] / total * (double)n;
        table_alias[i] = 0;
    }
    for (int64_t i = 0; i < n; ++i) (table_prob[i] < 1 ? small_ids : large_ids).push_back(i);
    while (!small_ids.empty() && !large_ids.empty()) {
        const int64_t s = small_ids.back(), l = large_ids.back();
        small_ids.pop_back();
        large_ids.pop_back();
        table_alias[s] = (int32_t)l;
        table_prob[l] = table_prob[l] + table_prob[s] - 1;
        (table_prob[l] < 1 ? small_ids : large_ids).push_back(l);
    }
    for (int64_t i : large_ids) table_prob[i] = 1;
    for (int64_t i : small_ids) table_prob[i] = 1;
    return SPONET_OK;
}

// sponet/cnvm/model.py:257-258 and :383-384.
extern "C" int sponet_cnvm_event_rates(const double* degree_alpha, int64_t n, double max_degree_alpha,
                                       double r_imit, double r_noise, double* next_event_rate,
                                       double* noise_probability) {
    SP_REQUIRE(next_event_rate && noise_probability && n >= 1, "cnvm_event_rates: bad arguments");
    if (degree_alpha) {
        double total = 0.0;
        for (int64_t i = 0; i < n; ++i) total += degree_alpha[i];
        const double denom = r_imit * total + r_noise * (double)n;
        if (denom == 0.0)
            return sponet_fail(SPONET_ERR_ZERO_RATE, "division by zero: all rates are zero");
        *next_event_rate = 1.0 / denom;
        *noise_probability = r_noise * (double)n * *next_event_rate;
    } else {
        const double denom = r_imit * max_degree_alpha + r_noise;
        if (denom == 0.0)
            return sponet_fail(SPONET_ERR_ZERO_RATE, "division by zero: all rates are zero");
        *next_event_rate = 1.0 / denom / (double)n;
        *noise_probability = r_noise / (r_noise + r_imit * max_degree_alpha);
    }
    return SPONET_OK;
}

// Derived graph layouts, one thread per (agent, slot) -- or per agent when there is no table:
//   col4   rows padded to multiples of four; the padding entries of row i hold i itself (an agent never counts
//          as a neighbour of the opposite opinion of itself, so the CNTM row scan needs no validity test)
//   nbtab  neighbour sampler table: agent i, slot j of L = 2^lg.  Neighbour k owns the slot interval
//          [k L / d, (k+1) L / d); slot j belongs to its primary neighbour floor(j d / L) up to the boundary
//          b = (primary + 1) L / d and to the next neighbour beyond it.  thr = floor(share * 2^24).
__global__ void graph_build_kernel(int64_t n, const int2* row, const int32_t* col, const int2* row4, int4* col4,
                                   uint2* nbtab, int lg) {
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t L = nbtab ? ((int64_t)1 << lg) : 1;
    const int64_t i = tid / L, j = tid - i * L;
    if (i >= n) return;
    const int2 r = row[i];
    if (j == 0) {  // the padded row
        int32_t* dst = reinterpret_cast<int32_t*>(col4 + row4[i].x);
        const int32_t padded = (r.y + 3) / 4 * 4;
        for (int32_t k = 0; k < r.y; ++k) dst[k] = col[r.x + k];
        for (int32_t k = r.y; k < padded; ++k) dst[k] = (int32_t)i;
    }
    if (nbtab) {
        const int64_t d = r.y;
        const int32_t* nb = col + r.x;
        const int64_t k = j * d / L;              // primary neighbour of slot j
        const int64_t num = (k + 1) * L - j * d;  // share of the slot owned by k, times d
        uint32_t id0 = (uint32_t)nb[k], id1 = id0, thr = 0xFFFFFFu;
        if (num < d) {  // the slot continues into neighbour k + 1
            id1 = (uint32_t)nb[k + 1];
            thr = (uint32_t)((num << 24) / d);
        }
        nbtab[i * L + j] = make_uint2(id0 | ((thr >> 12) << 20), id1 | ((thr & 0xFFFu) << 20));
    }
}

extern "C" int sponet_graph_create(int64_t n, const int32_t* row_ptr, const int32_t* col,
                                   sponet_graph** out) {
    SP_REQUIRE(out && row_ptr && n >= 1 && n < (int64_t)1 << 31, "graph_create: bad arguments");
    SP_REQUIRE(!sp_is_device_ptr(row_ptr) && !sp_is_device_ptr(col),
               "graph_create expects host CSR arrays");
    SP_REQUIRE(row_ptr[0] == 0, "graph_create: row_ptr[0] must be 0");
    const int64_t nnz = row_ptr[n];
    SP_REQUIRE(nnz == 0 || col, "graph_create: col is NULL");
    std::vector<int2> row(n);
    int32_t dmin = INT32_MAX, dmax = 0;
    for (int64_t i = 0; i < n; ++i) {
        const int32_t d = row_ptr[i + 1] - row_ptr[i];
        SP_REQUIRE(d >= 0, "graph_create: row_ptr not non-decreasing at %lld", (long long)i);
        dmin = std::min(dmin, d);
        dmax = std::max(dmax, d);
        row[i] = make_int2(row_ptr[i], d);
    }
    for (int64_t e = 0; e < nnz; ++e)
        SP_REQUIRE(col[e] >= 0 && col[e] < n, "graph_create: neighbour index %d out of range", col[e]);
    // padded copy (rows aligned to 16 bytes) and neighbour sampler table: derived ON THE DEVICE from the uploaded
    // CSR (graph_build_kernel below) -- at N = 1e5 the table alone is 3.2e6 entries / 25.6 MB, which a serial host
    // loop plus upload made the dominant cost of every new graph (per-run networks: one graph per realization).
    std::vector<int2> row4(n);
    int64_t n4 = 0;
    for (int64_t i = 0; i < n; ++i) {
        row4[i] = make_int2((int)n4, row[i].y);
        n4 += (row[i].y + 3) / 4;
    }
    SP_REQUIRE(n4 < ((int64_t)1 << 31), "graph_create: network too large");
    int lg = 1;
    while ((1 << lg) < dmax) ++lg;
    const bool use_tab = dmin >= 1 && dmax <= 64 && n <= (1 << 20) && ((int64_t)n << lg) * 8 <= ((int64_t)64 << 20);
    sponet_graph* g = new sponet_graph();
    g->d_nbtab = nullptr;
    g->batch = 1;
    g->d_rec = nullptr;
    g->pooled = false;
    g->d_alias = nullptr;
    g->h_weight_sum = nullptr;
    g->nbtab_log2 = use_tab ? lg : 0;
    g->n_col4 = n4;
    g->n = n;
    g->nnz = nnz;
    g->min_degree = dmin;
    g->max_degree = dmax;
    g->d_row_ptr = nullptr;
    g->d_row = nullptr;
    g->d_col = nullptr;
    g->d_row4 = nullptr;
    g->d_col4 = nullptr;
    cudaError_t e = cudaGetDevice(&g->device);
    if (e == cudaSuccess) e = cudaMalloc((void**)&g->d_row_ptr, (n + 1) * sizeof(int32_t));
    if (e == cudaSuccess) e = cudaMalloc((void**)&g->d_row, n * sizeof(int2));
    if (e == cudaSuccess) e = cudaMalloc((void**)&g->d_col, std::max<int64_t>(nnz, 1) * sizeof(int32_t));
    if (e == cudaSuccess) e = cudaMemcpy(g->d_row_ptr, row_ptr, (n + 1) * sizeof(int32_t), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(g->d_row, row.data(), n * sizeof(int2), cudaMemcpyHostToDevice);
    if (e == cudaSuccess && nnz) e = cudaMemcpy(g->d_col, col, nnz * sizeof(int32_t), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMalloc((void**)&g->d_row4, n * sizeof(int2));
    if (e == cudaSuccess) e = cudaMalloc((void**)&g->d_col4, (size_t)std::max<int64_t>(n4, 1) * sizeof(int4));
    if (e == cudaSuccess) e = cudaMemcpy(g->d_row4, row4.data(), n * sizeof(int2), cudaMemcpyHostToDevice);
    if (e == cudaSuccess && use_tab) e = cudaMalloc((void**)&g->d_nbtab, ((size_t)n << lg) * sizeof(uint2));
    if (e == cudaSuccess) {
        const int64_t work = use_tab ? ((int64_t)n << lg) : n;
        graph_build_kernel<<<(unsigned)((work + 255) / 256), 256>>>(n, g->d_row, g->d_col, g->d_row4, g->d_col4,
                                                                   use_tab ? g->d_nbtab : nullptr, lg);
        e = cudaGetLastError();
        // (the legacy default stream: does not wait for work on the non-blocking streams other graphs' chains run on)
        if (e == cudaSuccess) e = cudaStreamSynchronize(0);
    }
    if (e != cudaSuccess) {
        cudaFree(g->d_nbtab);
        cudaFree(g->d_row_ptr);
        cudaFree(g->d_row);
        cudaFree(g->d_col);
        cudaFree(g->d_row4);
        cudaFree(g->d_col4);
        delete g;
        cudaGetLastError();
        return sponet_fail(SPONET_ERR_CUDA, "graph_create: %s", cudaGetErrorString(e));
    }
    *out = g;
    return SPONET_OK;
}

extern "C" int sponet_graph_destroy(sponet_graph* g) {
    if (!g) return SPONET_OK;
    int cur = 0;
    cudaGetDevice(&cur);
    if (cur != g->device) cudaSetDevice(g->device);
    if (g->pooled) {  // batch memory goes back to the pool it came from (graphgen.cu)
        cudaDeviceSynchronize();
        sp_pool_free(g->d_row);
        sp_pool_free(g->d_col);
        sp_pool_free(g->d_row4);
        sp_pool_free(g->d_col4);
        sp_pool_free(g->d_rec);
        sp_pool_free(g->d_alias);
        delete[] g->h_weight_sum;
    } else {
        cudaFree(g->d_row_ptr);
        cudaFree(g->d_row);
        cudaFree(g->d_col);
        cudaFree(g->d_row4);
        cudaFree(g->d_col4);
        cudaFree(g->d_nbtab);
    }
    if (cur != g->device) cudaSetDevice(cur);
    delete g;
    return SPONET_OK;
}

extern "C" int sponet_graph_info(const sponet_graph* g, int64_t* num_agents, int64_t* nnz,
                                 int32_t* min_degree, int32_t* max_degree, int* device) {
    SP_REQUIRE(g, "graph is NULL");
    if (num_agents) *num_agents = g->n;
    if (nnz) *nnz = g->nnz;
    if (min_degree) *min_degree = g->min_degree;
    if (max_degree) *max_degree = g->max_degree;
    if (device) *device = g->device;
    return SPONET_OK;
}

extern "C" int sponet_graph_neighbor_table(const sponet_graph* g, int32_t* log2_slots, uint64_t* table) {
    SP_REQUIRE(g && log2_slots, "graph_neighbor_table: bad arguments");
    *log2_slots = g->d_nbtab ? g->nbtab_log2 : -1;
    if (g->d_nbtab && table) {
        int cur = 0;
        SP_CUDA(cudaGetDevice(&cur));
        if (cur != g->device) SP_CUDA(cudaSetDevice(g->device));
        cudaError_t e = cudaMemcpy(table, g->d_nbtab, ((size_t)g->n << g->nbtab_log2) * sizeof(uint2), cudaMemcpyDeviceToHost);
        if (cur != g->device) cudaSetDevice(cur);
        SP_CUDA(e);
    }
    return SPONET_OK;
}
