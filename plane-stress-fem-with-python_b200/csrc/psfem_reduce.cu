// psfem_reduce.cu -- boundary-condition reduction on CSR.
// Polygones.apply_boundary_conditions (Polygones.py:989-1006) and reduce (:898-911) build
// free_dofs with an O(n*c) list scan and slice the dense matrix with np.ix_; here the free map is a
// flag + exclusive scan and the reduced CSR is a two-pass row/column compaction.  Outputs are
// bit-identical to K_csr[free][:,free] (ascending free DOFs, column order preserved).
#include <cub/cub.cuh>
#include "psfem_common.cuh"

using namespace psfem;

namespace {

__global__ void fill_i32_kernel(int32_t *p, int64_t n, int32_t v) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// DOFs outside [0,n) never match `i in constrained_dofs` in the reference (:1000) -> ignored
__global__ void mark_constrained_kernel(const int32_t *__restrict__ constrained, int64_t nc, int64_t n, int32_t *__restrict__ flag) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= nc) return;
    int32_t d = constrained[i];
    if (d >= 0 && d < n) flag[d] = 0;
}

__global__ void free_map_kernel(const int32_t *__restrict__ flag, const int32_t *__restrict__ scan, int64_t n,
                                int32_t *__restrict__ new_id, int32_t *__restrict__ free_dofs) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (flag[i]) {
        new_id[i] = scan[i];
        free_dofs[scan[i]] = (int32_t)i;
    } else {
        new_id[i] = -1;
    }
}

__global__ void reduce_count_kernel(int64_t n, const int32_t *__restrict__ row_ptr, const int32_t *__restrict__ col_idx,
                                    const int32_t *__restrict__ new_id, int32_t *__restrict__ cnt) {
    int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r >= n) return;
    int32_t nr = new_id[r];
    if (nr < 0) return;
    int32_t c = 0;
    for (int32_t k = row_ptr[r]; k < row_ptr[r + 1]; ++k) c += (new_id[col_idx[k]] >= 0);
    cnt[nr] = c;
}

// One thread per row.  The free rows of a block are consecutive rows of the reduced matrix, so the block's output is one
// contiguous piece of col_idx_red / keep: it is compacted in shared memory and leaves with coalesced stores (the per-thread
// 4-byte stores at a stride of one row ran at 1 TB/s: 7 ms at 33.6 M rows).  Blocks whose rows do not fit the staging buffer
// (very long rows) write directly.
constexpr int RF_T = 256, RF_CAP = RF_T * 40;
constexpr size_t RF_SMEM = (size_t)RF_CAP * (sizeof(int32_t) + sizeof(uint16_t));
__global__ void __launch_bounds__(RF_T)
reduce_fill_kernel(int64_t n, const int32_t *__restrict__ row_ptr, const int32_t *__restrict__ col_idx,
                   const int32_t *__restrict__ new_id, const int32_t *__restrict__ row_ptr_red,
                   int32_t *__restrict__ col_idx_red, int32_t *__restrict__ keep) {
    extern __shared__ __align__(16) unsigned char rf_smem[];
    int32_t *s_col = reinterpret_cast<int32_t *>(rf_smem);
    uint16_t *s_k = reinterpret_cast<uint16_t *>(s_col + RF_CAP);
    __shared__ int s_lo, s_hi;
    const int64_t r0 = blockIdx.x * (int64_t)RF_T, r = r0 + threadIdx.x;
    if (threadIdx.x == 0) { s_lo = INT32_MAX; s_hi = -1; }
    __syncthreads();
    const int32_t nr = r < n ? new_id[r] : -1;
    int32_t o = 0;
    if (nr >= 0) {
        o = row_ptr_red[nr];
        atomicMin(&s_lo, o);
        atomicMax(&s_hi, row_ptr_red[nr + 1]);
    }
    __syncthreads();
    if (s_hi < 0) return;                                  // no free row in this block
    const int base = s_lo, cnt = s_hi - s_lo;
    const int64_t r_end = r0 + RF_T < n ? r0 + RF_T : n;
    const int32_t k_base = row_ptr[r0];
    const bool staged = cnt <= RF_CAP && row_ptr[r_end] - k_base <= 65535;   // the same decision in every thread
    if (nr >= 0) {
        for (int32_t k = row_ptr[r]; k < row_ptr[r + 1]; ++k) {
            const int32_t nc = new_id[col_idx[k]];
            if (nc >= 0) {
                if (staged) { s_col[o - base] = nc; s_k[o - base] = (uint16_t)(k - k_base); }
                else { col_idx_red[o] = nc; keep[o] = k; }
                ++o;
            }
        }
    }
    if (!staged) return;
    __syncthreads();
    for (int q = threadIdx.x; q < cnt; q += RF_T) {
        col_idx_red[base + q] = s_col[q];
        keep[base + q] = k_base + (int32_t)s_k[q];
    }
}

// CSR sanity of a matrix that arrives from the host: row_ptr non-decreasing, columns inside [0, n_cols) and strictly
// ascending within each row (what the SpMV / reduction kernels assume).  flags: bit0 = unsorted or duplicate columns,
// bit1 = column out of range, bit2 = row_ptr not monotone.  One pass over col_idx at HBM speed instead of a host scan.
__global__ void csr_check_kernel(int64_t n, int64_t n_cols, const int32_t *__restrict__ row_ptr, const int32_t *__restrict__ col_idx,
                                 int *__restrict__ flags) {
    const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t nnz = row_ptr[n] - row_ptr[0];
    int f = 0;
    if (k < n && row_ptr[k + 1] < row_ptr[k]) f |= 4;
    if (k < nnz) {
        const int64_t q = row_ptr[0] + k;
        const int32_t c = col_idx[q];
        if (c < 0 || c >= n_cols) f |= 2;
        if (k + 1 < nnz && col_idx[q + 1] <= c) {
            // a descent is fine only at a row boundary: find whether q+1 starts a row (binary search in row_ptr)
            int64_t lo = 0, hi = n;
            while (lo < hi) { const int64_t mid = lo + ((hi - lo) >> 1); if (row_ptr[mid] < q + 1) lo = mid + 1; else hi = mid; }
            if (lo > n || row_ptr[lo] != q + 1) f |= 1;
        }
    }
    if (f) atomicOr(flags, f);
}

__global__ void gather_kernel(const double *__restrict__ src, const int32_t *__restrict__ index, int64_t n, double *__restrict__ dst) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[index[i]];
}

__global__ void scatter_free_kernel(const double *__restrict__ red, const int32_t *__restrict__ free_dofs, int64_t n_free,
                                    int64_t n_full, int64_t nvec, double *__restrict__ full) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n_free * nvec) return;
    int64_t v = i / n_free, k = i - v * n_free;
    full[v * n_full + free_dofs[k]] = red[v * n_free + k];
}

size_t scan_temp_bytes(int64_t n) {
    size_t b = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, b, (int32_t *)nullptr, (int32_t *)nullptr, (int)n + 1);
    return align_up(b) + 256;
}

}  // namespace

extern "C" int psfem_scan_workspace(int64_t n, size_t *bytes) {
    PSFEM_REQUIRE(bytes && n >= 0 && n < (int64_t)INT32_MAX - 1, "scan_workspace: bad size");
    *bytes = 2 * align_up((n + 1) * sizeof(int32_t)) + scan_temp_bytes(n);
    return PSFEM_OK;
}

extern "C" int psfem_free_map(int64_t n, const int32_t *d_constrained, int64_t nc, int32_t *d_new_id,
                              int32_t *d_free_dofs, int64_t *h_n_free, void *d_ws, size_t ws_bytes, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    PSFEM_REQUIRE(n > 0 && n < (int64_t)INT32_MAX - 1 && h_n_free, "free_map: bad arguments");
    Carver c(d_ws);
    int32_t *flag = c.take<int32_t>(n + 1);
    int32_t *scan = c.take<int32_t>(n + 1);
    size_t tb = scan_temp_bytes(n);
    void *tmp = c.take<char>(tb);
    PSFEM_REQUIRE(ws_bytes >= c.used(), "free_map: workspace too small");
    const int T = 256;
    fill_i32_kernel<<<(unsigned)ceil_div(n, T), T, 0, stream>>>(flag, n, 1);
    PSFEM_LAUNCH_CHECK();
    PSFEM_CUDA_CHECK(cudaMemsetAsync(flag + n, 0, sizeof(int32_t), stream));
    if (nc > 0) {
        mark_constrained_kernel<<<(unsigned)ceil_div(nc, T), T, 0, stream>>>(d_constrained, nc, n, flag);
        PSFEM_LAUNCH_CHECK();
    }
    PSFEM_CUDA_CHECK(cub::DeviceScan::ExclusiveSum(tmp, tb, flag, scan, (int)n + 1, stream));
    free_map_kernel<<<(unsigned)ceil_div(n, T), T, 0, stream>>>(flag, scan, n, d_new_id, d_free_dofs);
    PSFEM_LAUNCH_CHECK();
    int32_t total = 0;
    PSFEM_CUDA_CHECK(cudaMemcpyAsync(&total, scan + n, sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
    PSFEM_CUDA_CHECK(cudaStreamSynchronize(stream));
    *h_n_free = total;
    return PSFEM_OK;
}

extern "C" int psfem_reduce_count(int64_t n, const int32_t *d_row_ptr, const int32_t *d_col_idx, const int32_t *d_new_id,
                                  int64_t n_free, int32_t *d_row_ptr_red, int64_t *h_nnz_red, void *d_ws, size_t ws_bytes,
                                  void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    PSFEM_REQUIRE(n > 0 && n_free >= 0 && n_free <= n && h_nnz_red, "reduce_count: bad arguments");
    Carver c(d_ws);
    int32_t *cnt = c.take<int32_t>(n + 1);
    (void)c.take<int32_t>(n + 1);
    size_t tb = scan_temp_bytes(n);
    void *tmp = c.take<char>(tb);
    PSFEM_REQUIRE(ws_bytes >= c.used(), "reduce_count: workspace too small");
    const int T = 256;
    PSFEM_CUDA_CHECK(cudaMemsetAsync(cnt, 0, (n_free + 1) * sizeof(int32_t), stream));
    reduce_count_kernel<<<(unsigned)ceil_div(n, T), T, 0, stream>>>(n, d_row_ptr, d_col_idx, d_new_id, cnt);
    PSFEM_LAUNCH_CHECK();
    PSFEM_CUDA_CHECK(cub::DeviceScan::ExclusiveSum(tmp, tb, cnt, d_row_ptr_red, (int)n_free + 1, stream));
    int32_t total = 0;
    PSFEM_CUDA_CHECK(cudaMemcpyAsync(&total, d_row_ptr_red + n_free, sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
    PSFEM_CUDA_CHECK(cudaStreamSynchronize(stream));
    *h_nnz_red = total;
    return PSFEM_OK;
}

extern "C" int psfem_reduce_fill(int64_t n, const int32_t *d_row_ptr, const int32_t *d_col_idx, const int32_t *d_new_id,
                                 const int32_t *d_row_ptr_red, int32_t *d_col_idx_red, int32_t *d_keep, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    PSFEM_CUDA_CHECK(cudaFuncSetAttribute(reduce_fill_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RF_SMEM));
    reduce_fill_kernel<<<(unsigned)ceil_div(n, RF_T), RF_T, RF_SMEM, stream>>>(n, d_row_ptr, d_col_idx, d_new_id, d_row_ptr_red, d_col_idx_red, d_keep);
    PSFEM_LAUNCH_CHECK();
    return PSFEM_OK;
}

extern "C" int psfem_csr_check(int64_t n, int64_t n_cols, const int32_t *d_row_ptr, const int32_t *d_col_idx, int64_t nnz,
                               int64_t *h_flags, void *d_ws, size_t ws_bytes, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    PSFEM_REQUIRE(n >= 0 && d_row_ptr && h_flags && d_ws && ws_bytes >= 16, "csr_check: bad arguments");
    int *flags = static_cast<int *>(d_ws);
    PSFEM_CUDA_CHECK(cudaMemsetAsync(flags, 0, sizeof(int), stream));
    const int64_t work = nnz > n ? nnz : n;
    if (work > 0) {
        csr_check_kernel<<<(unsigned)ceil_div(work, 256), 256, 0, stream>>>(n, n_cols, d_row_ptr, d_col_idx, flags);
        PSFEM_LAUNCH_CHECK();
    }
    int h = 0;
    PSFEM_CUDA_CHECK(cudaMemcpyAsync(&h, flags, sizeof(int), cudaMemcpyDeviceToHost, stream));
    PSFEM_CUDA_CHECK(cudaStreamSynchronize(stream));
    *h_flags = h;
    return PSFEM_OK;
}

extern "C" int psfem_gather(const double *d_src, const int32_t *d_index, int64_t n, double *d_dst, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (n == 0) return PSFEM_OK;
    gather_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, stream>>>(d_src, d_index, n, d_dst);
    PSFEM_LAUNCH_CHECK();
    return PSFEM_OK;
}

extern "C" int psfem_scatter_free(const double *d_red, const int32_t *d_free_dofs, int64_t n_free, int64_t n_full,
                                  int64_t nvec, double *d_full, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    PSFEM_CUDA_CHECK(cudaMemsetAsync(d_full, 0, n_full * nvec * sizeof(double), stream));
    if (n_free * nvec == 0) return PSFEM_OK;
    scatter_free_kernel<<<(unsigned)ceil_div(n_free * nvec, 256), 256, 0, stream>>>(d_red, d_free_dofs, n_free, n_full, nvec, d_full);
    PSFEM_LAUNCH_CHECK();
    return PSFEM_OK;
}
