// psfem_pattern.cu -- symbolic phase: structural CSR pattern + assembly plan.
//
// Replaces the dense np.zeros((2n,2n)) target of Polygones.py:715 / :1349.  Every element emits
// its npe^2 ordered node pairs (a,b) as 64-bit keys a<<32|b tagged with the contribution code
// e*npe^2 + i*npe + j.  One stable LSD radix sort (cub) + run-length unique gives
//   - the node-level adjacency (unique keys, sorted by a then b)  -> CSR pattern by 2x2 expansion
//   - for every unique pair the contribution codes in ascending element id -> the assembly plan,
//     i.e. the summation order of the reference's scatter loop Polygones.py:717-726.
#include <cub/cub.cuh>
#include "psfem_common.cuh"

using namespace psfem;

namespace {

struct PatternWs {
    unsigned long long *keys_a, *keys_b;  // keys_b: sorted keys; keys_a: unique keys after RLE
    int32_t *vals_a;                      // sort scratch, then run lengths
    int *num_runs;
    void *cub_tmp;
    size_t cub_bytes;
    size_t total;
};

size_t cub_temp_bytes(int64_t nk, int64_t n_nodes) {
    size_t a = 0, b = 0, c = 0;
    cub::DoubleBuffer<unsigned long long> dk(nullptr, nullptr);
    cub::DoubleBuffer<int32_t> dv(nullptr, nullptr);
    cub::DeviceRadixSort::SortPairs(nullptr, a, dk, dv, (int)nk, 0, 64);
    cub::DeviceRunLengthEncode::Encode(nullptr, b, (unsigned long long *)nullptr, (unsigned long long *)nullptr,
                                       (int32_t *)nullptr, (int *)nullptr, (int)nk);
    cub::DeviceScan::ExclusiveSum(nullptr, c, (int32_t *)nullptr, (int32_t *)nullptr, (int)nk + 1);
    size_t r = a > b ? a : b;
    return align_up(r > c ? r : c) + 256;
}

PatternWs carve(void *ws, int64_t nk, int64_t n_nodes) {
    PatternWs w;
    Carver c(ws);
    w.keys_a = c.take<unsigned long long>(nk);
    w.keys_b = c.take<unsigned long long>(nk);
    w.vals_a = c.take<int32_t>(nk + 1);
    w.num_runs = c.take<int>(4);
    w.cub_bytes = cub_temp_bytes(nk, n_nodes);
    w.cub_tmp = c.take<char>(w.cub_bytes);
    w.total = c.used();
    return w;
}

int key_bits(int64_t n_nodes) {
    int b = 1;
    while ((1ll << b) < n_nodes) ++b;
    return b;
}

template <int NPE>
__global__ void pair_keys_kernel(const int32_t *__restrict__ conn, int64_t nel, unsigned long long *__restrict__ keys,
                                 int32_t *__restrict__ codes) {
    // one thread per (element, i): NPE keys, contiguous codes
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= nel * NPE) return;
    int64_t e = t / NPE;
    int i = (int)(t - e * NPE);
    const int32_t *c = conn + e * NPE;
    unsigned long long a = (unsigned long long)(uint32_t)c[i] << 32;
    int64_t base = t * NPE;
#pragma unroll
    for (int j = 0; j < NPE; ++j) {
        keys[base + j] = a | (uint32_t)c[j];
        codes[base + j] = (int32_t)(base + j);
    }
}

__global__ void conn_range_kernel(const int32_t *__restrict__ conn, int64_t count, int32_t n_nodes, int *bad) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= count) return;
    int32_t v = conn[t];
    if (v < 0 || v >= n_nodes) atomicExch(bad, 1);
}

// node_ptr[a] = first unique pair whose row node is >= a  (lower bound on the sorted unique keys)
__global__ void node_ptr_kernel(const unsigned long long *__restrict__ ukeys, int32_t n_pairs, int64_t n_nodes,
                                int32_t *__restrict__ node_ptr, int32_t *__restrict__ row_ptr) {
    int64_t a = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (a > n_nodes) return;
    unsigned long long key = (unsigned long long)a << 32;
    int32_t lo = 0, hi = n_pairs;
    while (lo < hi) {
        int32_t mid = lo + ((hi - lo) >> 1);
        if (ukeys[mid] < key) lo = mid + 1; else hi = mid;
    }
    node_ptr[a] = lo;
    // DOF rows 2a, 2a+1 each hold 2*deg(a) columns: row_ptr[2a] = 4*node_ptr[a]; row_ptr[2a+1] needs deg -> second kernel
    if (a == n_nodes) row_ptr[2 * n_nodes] = 4 * lo;
}

__global__ void row_ptr_kernel(const int32_t *__restrict__ node_ptr, int64_t n_nodes, int32_t *__restrict__ row_ptr) {
    int64_t a = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (a >= n_nodes) return;
    int32_t p0 = node_ptr[a], deg = node_ptr[a + 1] - p0;
    row_ptr[2 * a] = 4 * p0;
    row_ptr[2 * a + 1] = 4 * p0 + 2 * deg;
}

__global__ void col_idx_kernel(const unsigned long long *__restrict__ ukeys, int32_t n_pairs,
                               const int32_t *__restrict__ node_ptr, int32_t *__restrict__ pair_row,
                               int32_t *__restrict__ col_idx) {
    int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (s >= n_pairs) return;
    unsigned long long k = ukeys[s];
    int32_t a = (int32_t)(k >> 32), b = (int32_t)(k & 0xffffffffu);
    int32_t p0 = node_ptr[a], deg = node_ptr[a + 1] - p0, pos = (int32_t)s - p0;
    pair_row[s] = a;
    int64_t base = 4ll * p0 + 2 * pos;
    *reinterpret_cast<int2 *>(col_idx + base) = make_int2(2 * b, 2 * b + 1);            // row 2a
    *reinterpret_cast<int2 *>(col_idx + base + 2 * deg) = make_int2(2 * b, 2 * b + 1);  // row 2a+1
}

}  // namespace

extern "C" int psfem_pattern_workspace(int64_t nel, int npe, int64_t n_nodes, size_t *bytes) {
    PSFEM_REQUIRE(bytes, "pattern_workspace: null output");
    PSFEM_REQUIRE(nel >= 0 && npe > 0 && npe <= 9, "pattern_workspace: bad sizes");
    int64_t nk = nel * npe * npe;
    PSFEM_REQUIRE(nk < (int64_t)INT32_MAX - 1, "pattern: nel*npe^2 = %lld exceeds int32 (partition the mesh across GPUs)", (long long)nk);
    PatternWs w = carve(nullptr, nk > 0 ? nk : 1, n_nodes);
    *bytes = w.total;
    return PSFEM_OK;
}

extern "C" int psfem_pattern_count(const int32_t *d_conn, int64_t nel, int npe, int64_t n_nodes, int32_t *d_perm,
                                   void *d_ws, size_t ws_bytes, int64_t *h_n_pairs, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    PSFEM_REQUIRE(h_n_pairs, "pattern_count: null output");
    PSFEM_REQUIRE(npe == 3 || npe == 4 || npe == 6 || npe == 9, "pattern_count: nodes per element must be 3, 4, 6 or 9");
    PSFEM_REQUIRE(n_nodes > 0 && n_nodes < (int64_t)INT32_MAX / 2, "pattern_count: node count out of range");
    int64_t nk = nel * npe * npe;
    PSFEM_REQUIRE(nk < (int64_t)INT32_MAX - 1, "pattern: nel*npe^2 = %lld exceeds int32 (partition the mesh across GPUs)", (long long)nk);
    if (nel == 0) { *h_n_pairs = 0; return PSFEM_OK; }
    PatternWs w = carve(d_ws, nk, n_nodes);
    PSFEM_REQUIRE(ws_bytes >= w.total, "pattern_count: workspace too small (%zu < %zu)", ws_bytes, w.total);
    const int T = 256;
    // connectivity must address existing nodes (the reference would raise IndexError)
    PSFEM_CUDA_CHECK(cudaMemsetAsync(w.num_runs, 0, 4 * sizeof(int), stream));
    conn_range_kernel<<<(unsigned)ceil_div(nel * npe, T), T, 0, stream>>>(d_conn, nel * npe, (int32_t)n_nodes, w.num_runs + 1);
    PSFEM_LAUNCH_CHECK();
    unsigned grid = (unsigned)ceil_div(nel * npe, T);
    switch (npe) {
        case 3: pair_keys_kernel<3><<<grid, T, 0, stream>>>(d_conn, nel, w.keys_a, w.vals_a); break;
        case 4: pair_keys_kernel<4><<<grid, T, 0, stream>>>(d_conn, nel, w.keys_a, w.vals_a); break;
        case 6: pair_keys_kernel<6><<<grid, T, 0, stream>>>(d_conn, nel, w.keys_a, w.vals_a); break;
        default: pair_keys_kernel<9><<<grid, T, 0, stream>>>(d_conn, nel, w.keys_a, w.vals_a); break;
    }
    PSFEM_LAUNCH_CHECK();
    cub::DoubleBuffer<unsigned long long> dk(w.keys_a, w.keys_b);
    cub::DoubleBuffer<int32_t> dv(w.vals_a, d_perm);
    size_t tb = w.cub_bytes;
    const int kb = key_bits(n_nodes);
    // only the populated bits of (a,b) are sorted: [0,kb) and [32,32+kb)
    PSFEM_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(w.cub_tmp, tb, dk, dv, (int)nk, 0, kb, stream));
    tb = w.cub_bytes;
    PSFEM_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(w.cub_tmp, tb, dk, dv, (int)nk, 32, 32 + kb, stream));
    // normalise: sorted keys in keys_b, sorted codes in d_perm
    if (dk.Current() != w.keys_b)
        PSFEM_CUDA_CHECK(cudaMemcpyAsync(w.keys_b, dk.Current(), nk * sizeof(unsigned long long), cudaMemcpyDeviceToDevice, stream));
    if (dv.Current() != d_perm)
        PSFEM_CUDA_CHECK(cudaMemcpyAsync(d_perm, dv.Current(), nk * sizeof(int32_t), cudaMemcpyDeviceToDevice, stream));
    tb = w.cub_bytes;
    PSFEM_CUDA_CHECK(cub::DeviceRunLengthEncode::Encode(w.cub_tmp, tb, w.keys_b, w.keys_a, w.vals_a, w.num_runs, (int)nk, stream));
    int h[2] = {0, 0};
    PSFEM_CUDA_CHECK(cudaMemcpyAsync(h, w.num_runs, 2 * sizeof(int), cudaMemcpyDeviceToHost, stream));
    PSFEM_CUDA_CHECK(cudaStreamSynchronize(stream));
    PSFEM_REQUIRE(h[1] == 0, "pattern_count: connectivity references a node outside [0, %lld)", (long long)n_nodes);
    *h_n_pairs = h[0];
    return PSFEM_OK;
}

extern "C" int psfem_pattern_fill(int64_t nel, int npe, int64_t n_nodes, int64_t n_pairs, void *d_ws, size_t ws_bytes,
                                  int32_t *d_row_ptr, int32_t *d_col_idx, int32_t *d_node_ptr, int32_t *d_pair_row,
                                  int32_t *d_pair_ptr, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    int64_t nk = nel * npe * npe;
    PSFEM_REQUIRE(4 * n_pairs < (int64_t)INT32_MAX, "pattern_fill: nnz = %lld exceeds int32", (long long)(4 * n_pairs));
    const int T = 256;
    if (nel == 0 || n_pairs == 0) {
        PSFEM_CUDA_CHECK(cudaMemsetAsync(d_row_ptr, 0, (2 * n_nodes + 1) * sizeof(int32_t), stream));
        PSFEM_CUDA_CHECK(cudaMemsetAsync(d_node_ptr, 0, (n_nodes + 1) * sizeof(int32_t), stream));
        PSFEM_CUDA_CHECK(cudaMemsetAsync(d_pair_ptr, 0, sizeof(int32_t), stream));
        return PSFEM_OK;
    }
    PatternWs w = carve(d_ws, nk, n_nodes);
    PSFEM_REQUIRE(ws_bytes >= w.total, "pattern_fill: workspace too small");
    // pair_ptr = exclusive scan of the run lengths (n_pairs+1 entries; the last is nk)
    PSFEM_CUDA_CHECK(cudaMemsetAsync(w.vals_a + n_pairs, 0, sizeof(int32_t), stream));
    size_t tb = w.cub_bytes;
    PSFEM_CUDA_CHECK(cub::DeviceScan::ExclusiveSum(w.cub_tmp, tb, w.vals_a, d_pair_ptr, (int)n_pairs + 1, stream));
    node_ptr_kernel<<<(unsigned)ceil_div(n_nodes + 1, T), T, 0, stream>>>(w.keys_a, (int32_t)n_pairs, n_nodes, d_node_ptr, d_row_ptr);
    PSFEM_LAUNCH_CHECK();
    row_ptr_kernel<<<(unsigned)ceil_div(n_nodes, T), T, 0, stream>>>(d_node_ptr, n_nodes, d_row_ptr);
    PSFEM_LAUNCH_CHECK();
    col_idx_kernel<<<(unsigned)ceil_div(n_pairs, T), T, 0, stream>>>(w.keys_a, (int32_t)n_pairs, d_node_ptr, d_pair_row, d_col_idx);
    PSFEM_LAUNCH_CHECK();
    return PSFEM_OK;
}
