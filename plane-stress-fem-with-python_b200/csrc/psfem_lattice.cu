// psfem_lattice.cu -- numeric assembly for meshes with the LATTICE TOPOLOGY of Polygones.mesh (Q4).
//
// Polygones.mesh (Polygones.py:62-108) is the reference's only mesh generator: node id = i*m + j, Q4 element
// (i,j) = [i*m+j, (i+1)*m+j, (i+1)*m+j+1, i*m+j+1] (:100).  When the symbolic phase finds that topology
// (psfem_lattice_detect compares every element with the formula; the COORDINATES stay arbitrary -- jittered
// meshes take this path too), everything the general tile plan stores is arithmetic: the neighbours of a node
// are the 3x3 stencil, its CSR block starts at pair (3i-[i>0])*(3m-2) + nci*(3j-[j>0]), and the numeric phase
// needs no plan, no connectivity and no shared-memory scatter:
//
//   * a WARP owns a strip of 31 node rows (j) and MARCHES along i.  Lane l computes element (ci, j0-1+l) ONCE
//     per step (3x3 Gauss products in registers, psfem_elem.cuh).  Lane 0 is the only halo (32/31 redundancy),
//     plus one element column per segment of `seg_len` node columns;
//   * the rows of the element matrix that belong to the node column still open (ci+1) stay in REGISTERS as
//     "pending" 2x2 blocks until the next step completes them; the rows owned by the lane above travel by
//     warp shuffle.  Each CSR entry is summed in ascending element id -- (i-1,j-1), (i-1,j), (i,j-1), (i,j) --
//     the order of the reference's scatter loop Polygones.py:717-726, without atomics: bit-reproducible;
//   * a finished node column of the strip (31 nodes x 288 B, contiguous in the CSR value array) is laid out in
//     the warp's shared buffer in exactly the global layout and leaves with ONE bulk async (TMA) store,
//     double buffered -- HBM sees only full-line writes and the 32 B/node coordinate reads.
//
// HBM traffic is therefore the algorithmic minimum (304 B/element; the connectivity is not read at all) and
// the kernel is bound by the fp64 pipe (about 0.9 k DFMA per element) -- see DESIGN.md section 4.
#include <cstdlib>
#include "psfem_common.cuh"
#include "psfem_elem.cuh"

using namespace psfem;

namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int STRIP = 31;                 // node rows finished per warp and step
constexpr int LBUF_BYTES = 8960;          // 31 nodes * 288 B = 8928, padded to a multiple of 128

template <int ETYPE> __global__ void lattice_check_kernel(const int32_t *__restrict__ conn, int64_t nel, int m, int *__restrict__ flag) {
    const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= nel) return;
    const int M = m - 1;
    bool ok;
    if (ETYPE == PSFEM_Q4) {
        const int64_t i = e / M, j = e % M;
        const int64_t b = i * m + j;
        const int4 c = reinterpret_cast<const int4 *>(conn)[e];
        ok = c.x == b && c.y == b + m && c.z == b + m + 1 && c.w == b + 1;                       // Polygones.py:100
    } else {
        const int64_t cell = e >> 1, i = cell / M, j = cell % M;
        const int64_t b = i * m + j;
        const int32_t *c = conn + 3 * e;
        ok = (e & 1) ? (c[0] == b + m + 1 && c[1] == b + 1 && c[2] == b + m)                     // Polygones.py:85
                     : (c[0] == b && c[1] == b + m && c[2] == b + 1);                            // Polygones.py:83
    }
    if (!ok) atomicExch(flag, 1);
}

// ---- 2x2 blocks: {K(2a,2b), K(2a,2b+1), K(2a+1,2b), K(2a+1,2b+1)}; the mass block is m*I2 -> one scalar ----
struct KB { double x, y, z, w; };
struct MB { double x; };
__device__ __forceinline__ KB up1(const KB &b) {
    KB r;
    r.x = __shfl_up_sync(FULL, b.x, 1); r.y = __shfl_up_sync(FULL, b.y, 1);
    r.z = __shfl_up_sync(FULL, b.z, 1); r.w = __shfl_up_sync(FULL, b.w, 1);
    return r;
}
__device__ __forceinline__ MB up1(const MB &b) { MB r; r.x = __shfl_up_sync(FULL, b.x, 1); return r; }
__device__ __forceinline__ void add(KB &a, const KB &b) { a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; }
__device__ __forceinline__ void add(MB &a, const MB &b) { a.x += b.x; }
__device__ __forceinline__ KB sum(const KB &a, const KB &b) { KB r = a; add(r, b); return r; }
__device__ __forceinline__ MB sum(const MB &a, const MB &b) { MB r = a; add(r, b); return r; }
__device__ __forceinline__ void zero(KB &a) { a.x = a.y = a.z = a.w = 0.0; }
__device__ __forceinline__ void zero(MB &a) { a.x = 0.0; }
__device__ __forceinline__ void put(double2 *row0, int deg, int q, const KB &b) {
    row0[q] = make_double2(b.x, b.y);
    row0[deg + q] = make_double2(b.z, b.w);
}
__device__ __forceinline__ void put(double2 *row0, int deg, int q, const MB &b) {
    row0[q] = make_double2(b.x, 0.0);          // explicit zeros at the u-v couplings (shared K/M pattern)
    row0[deg + q] = make_double2(0.0, b.x);
}

template <bool IS_K> struct BlockOf { using type = KB; };
template <> struct BlockOf<false> { using type = MB; };

template <bool IS_K, class EL> __device__ __forceinline__ typename BlockOf<IS_K>::type blk(const EL &el, int r, int c, const Mat &mat) {
    if constexpr (IS_K) {
        const double4 v = el.kblock(r, c, mat);
        KB b; b.x = v.x; b.y = v.y; b.z = v.z; b.w = v.w;
        return b;
    } else {
        MB b; b.x = el.mval(r, c);
        return b;
    }
}

__device__ __forceinline__ uint32_t l_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// pairs of the nodes 0..j-1 of one node column, per neighbour column (the column has 3m-2 in total)
__device__ __forceinline__ int colq(int j, int m) { return 3 * j - (j > 0) - (j == m); }

template <bool IS_K, int WARPS, int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB)
lattice_q4_kernel(const double2 *__restrict__ nodes, Mat mat, int n, int m, int seg_len, int n_seg, int n_strips,
                  double *__restrict__ vals, unsigned long long *__restrict__ bad, unsigned int *__restrict__ counter) {
    using B = typename BlockOf<IS_K>::type;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    unsigned char *wbuf = smem_raw + (size_t)wid * 2 * LBUF_BYTES;
    const int N = n - 1, M = m - 1;
    const int n_items = n_seg * n_strips;
    int parity = 0;
    for (;;) {
        int item = 0;
        if (lane == 0) item = (int)atomicAdd(counter, 1u);
        item = __shfl_sync(FULL, item, 0);
        if (item >= n_items) break;
        const int s = item % n_strips, g = item / n_strips;     // neighbouring warps: neighbouring strips of one segment
        const int ia = g * seg_len, ib = min(ia + seg_len, n);   // node columns finished by this item
        const int js = STRIP * s, je = min(js + STRIP, m);       // node rows finished by this item
        const int jl = js - 1 + lane;                            // lane l: element row jl, node row jl
        const bool el_j_ok = jl >= 0 && jl < M;
        const bool node_out = lane >= 1 && jl < m;
        const bool has_jm = jl > 0, has_jp = jl < M;
        const int njc = 1 + (has_jm ? 1 : 0) + (has_jp ? 1 : 0);
        const int cj1 = has_jm ? 1 : 0;
        const int qrel = colq(max(jl, js), m) - colq(js, m);
        const int qtot = colq(je, m) - colq(js, m);
        const int64_t jc0 = min(max(jl, 0), M), jc1 = min(max(jl + 1, 0), M);   // clamped: loads stay in range
        int ci = max(ia - 1, 0);
        double2 p0 = nodes[(int64_t)ci * m + jc0], p3 = nodes[(int64_t)ci * m + jc1];
        double2 p1 = nodes[(int64_t)min(ci + 1, N) * m + jc0], p2 = nodes[(int64_t)min(ci + 1, N) * m + jc1];
        B acc[9];
#pragma unroll
        for (int q = 0; q < 9; ++q) zero(acc[q]);
        for (; ci < ib; ++ci) {
            const int64_t cpf = min(ci + 2, N);                  // coordinates of the step after this one
            const double2 f1 = nodes[cpf * m + jc0], f2 = nodes[cpf * m + jc1];
            // ptxas consumes f1/f2 right away (it rotates the coordinate registers early), so the column after that
            // is pulled into L1 one step ahead: the loads above then hit L1 instead of waiting for HBM
            asm volatile("prefetch.global.L1 [%0];" ::"l"(nodes + (int64_t)min(ci + 3, N) * m + jc0));
            if (lane == 31) asm volatile("prefetch.global.L1 [%0];" ::"l"(nodes + (int64_t)min(ci + 3, N) * m + jc1));
            const bool el_ok = el_j_ok && ci < N;
            ElemQ4<IS_K, !IS_K> el;
            {
                // lanes without an element integrate a unit square with zero thickness: all products are exact zeros
                double2 p[4];
                p[0] = el_ok ? p0 : make_double2(0.0, 0.0); p[1] = el_ok ? p1 : make_double2(1.0, 0.0);
                p[2] = el_ok ? p2 : make_double2(1.0, 1.0); p[3] = el_ok ? p3 : make_double2(0.0, 1.0);
                Mat ml = mat;
                ml.t = el_ok ? mat.t : 0.0;
                if (!el.compute(p, ml)) atomicMin(bad, (unsigned long long)((int64_t)ci * M + jl));   // Polygones.py:686-687
            }
            if (ci >= ia) {   // ---- finish node column ci: rows 3 of the element below (lane-1), rows 0 of mine ----
                const B b30 = up1(blk<IS_K>(el, 3, 0, mat)), b31 = up1(blk<IS_K>(el, 3, 1, mat));
                const B b32 = up1(blk<IS_K>(el, 3, 2, mat)), b33 = up1(blk<IS_K>(el, 3, 3, mat));
                add(acc[3], b30); acc[6] = b31; acc[7] = b32; add(acc[4], b33);
                add(acc[4], blk<IS_K>(el, 0, 0, mat)); add(acc[7], blk<IS_K>(el, 0, 1, mat));
                acc[8] = blk<IS_K>(el, 0, 2, mat); add(acc[5], blk<IS_K>(el, 0, 3, mat));
                // the bulk store that read this buffer two steps ago has finished reading
                if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                __syncwarp();
                const bool has_im = ci > 0, has_ip = ci < N;
                const int nci = 1 + (has_im ? 1 : 0) + (has_ip ? 1 : 0);
                double2 *buf = reinterpret_cast<double2 *>(wbuf + parity * LBUF_BYTES);
                if (node_out) {
                    const int deg = nci * njc;
                    double2 *row0 = buf + 2 * nci * qrel;
                    int rb = 0;
                    if (has_im) {
                        if (has_jm) put(row0, deg, rb, acc[0]);
                        put(row0, deg, rb + cj1, acc[1]);
                        if (has_jp) put(row0, deg, rb + cj1 + 1, acc[2]);
                        rb += njc;
                    }
                    if (has_jm) put(row0, deg, rb, acc[3]);
                    put(row0, deg, rb + cj1, acc[4]);
                    if (has_jp) put(row0, deg, rb + cj1 + 1, acc[5]);
                    rb += njc;
                    if (has_ip) {
                        if (has_jm) put(row0, deg, rb, acc[6]);
                        put(row0, deg, rb + cj1, acc[7]);
                        if (has_jp) put(row0, deg, rb + cj1 + 1, acc[8]);
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // my generic writes -> visible to the async proxy
                __syncwarp();
                if (lane == 0) {
                    const int64_t pair0 = (int64_t)(3 * ci - (ci > 0)) * (3 * m - 2) + (int64_t)nci * colq(js, m);
                    const uint32_t bytes = (uint32_t)(nci * qtot) * 32u;
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                                 ::"l"(vals + 4 * pair0), "r"(l_smem_u32(buf)), "r"(bytes) : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
                parity ^= 1;
            }
            // ---- open node column ci+1: rows 2 of the element below (lane-1), then rows 1 of mine ----
            {
                const B b20 = up1(blk<IS_K>(el, 2, 0, mat)), b21 = up1(blk<IS_K>(el, 2, 1, mat));
                const B b22 = up1(blk<IS_K>(el, 2, 2, mat)), b23 = up1(blk<IS_K>(el, 2, 3, mat));
                acc[0] = b20;
                acc[1] = sum(b23, blk<IS_K>(el, 1, 0, mat));
                acc[2] = blk<IS_K>(el, 1, 3, mat);
                acc[3] = b21;
                acc[4] = sum(b22, blk<IS_K>(el, 1, 1, mat));
                acc[5] = blk<IS_K>(el, 1, 2, mat);
            }
            p0 = p1; p3 = p2; p1 = f1; p2 = f2;
        }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // all stores complete before the warp exits
}

template <bool IS_K, int WARPS, int MINB>
int launch_lattice_q4(const double *d_nodes, const Mat &mat, int n, int m, int seg_len, double *vals, unsigned long long *bad,
                      unsigned int *counter, cudaStream_t stream) {
    auto kern = lattice_q4_kernel<IS_K, WARPS, MINB>;
    const size_t smem = (size_t)WARPS * 2 * LBUF_BYTES;
    PSFEM_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int dev = 0, sms = 148, per_sm = 1;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    PSFEM_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, WARPS * 32, smem));
    if (per_sm < 1) per_sm = 1;
    const int n_strips = (int)ceil_div(m, STRIP);
    // segments: as close to seg_len as possible with equal lengths
    int n_seg = (int)ceil_div(n, seg_len);
    seg_len = (int)ceil_div(n, n_seg);
    n_seg = (int)ceil_div(n, seg_len);
    const int64_t n_items = (int64_t)n_seg * n_strips;
    int64_t grid = (int64_t)sms * per_sm;
    if (grid * WARPS > n_items) grid = ceil_div(n_items, WARPS);
    PSFEM_CUDA_CHECK(cudaMemsetAsync(counter, 0, sizeof(unsigned int), stream));
    kern<<<(unsigned)grid, WARPS * 32, smem, stream>>>(reinterpret_cast<const double2 *>(d_nodes), mat, n, m, seg_len, n_seg, n_strips,
                                                       vals, bad, counter);
    PSFEM_LAUNCH_CHECK();
    return PSFEM_OK;
}

}  // namespace

extern "C" int psfem_lattice_detect(const int32_t *d_conn, int64_t nel, int etype, int64_t n_nodes, int64_t *h_nm,
                                    void *d_ws, size_t ws_bytes, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    PSFEM_REQUIRE(h_nm && d_ws && ws_bytes >= 16, "lattice_detect: bad arguments");
    h_nm[0] = h_nm[1] = 0;
    if ((etype != PSFEM_Q4 && etype != PSFEM_T3) || nel <= 0 || n_nodes < 4 || !d_conn) return PSFEM_OK;
    const int npe = etype == PSFEM_Q4 ? 4 : 3;
    int32_t first[4] = {0, 0, 0, 0};
    PSFEM_CUDA_CHECK(cudaMemcpyAsync(first, d_conn, npe * sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
    PSFEM_CUDA_CHECK(cudaStreamSynchronize(stream));
    const int64_t m = first[1];                                  // element 0 = [0, m, ...]
    if (first[0] != 0 || m < 2 || n_nodes % m != 0) return PSFEM_OK;
    const int64_t n = n_nodes / m;
    if (n < 2 || nel != (n - 1) * (m - 1) * (etype == PSFEM_T3 ? 2 : 1) || n_nodes >= (1ll << 31)) return PSFEM_OK;
    int *flag = static_cast<int *>(d_ws);
    PSFEM_CUDA_CHECK(cudaMemsetAsync(flag, 0, sizeof(int), stream));
    const int T = 256;
    if (etype == PSFEM_Q4) lattice_check_kernel<PSFEM_Q4><<<(unsigned)ceil_div(nel, T), T, 0, stream>>>(d_conn, nel, (int)m, flag);
    else lattice_check_kernel<PSFEM_T3><<<(unsigned)ceil_div(nel, T), T, 0, stream>>>(d_conn, nel, (int)m, flag);
    PSFEM_LAUNCH_CHECK();
    int h_flag = 1;
    PSFEM_CUDA_CHECK(cudaMemcpyAsync(&h_flag, flag, sizeof(int), cudaMemcpyDeviceToHost, stream));
    PSFEM_CUDA_CHECK(cudaStreamSynchronize(stream));
    if (h_flag == 0) { h_nm[0] = n; h_nm[1] = m; }
    return PSFEM_OK;
}

extern "C" int psfem_assemble_lattice(const double *d_nodes, int64_t n, int64_t m, int etype,
                                      double E, double nu, double rho, double t, int what,
                                      double *d_K_vals, double *d_M_vals, void *d_ws, size_t ws_bytes,
                                      int64_t *h_bad_element, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    PSFEM_REQUIRE(etype == PSFEM_Q4, "assemble_lattice: Q4 only");
    PSFEM_REQUIRE(what >= 1 && what <= 3, "assemble_lattice: what must be 1 (K), 2 (M) or 3 (both)");
    PSFEM_REQUIRE(!(what & 1) || d_K_vals, "assemble_lattice: K requested but d_K_vals is null");
    PSFEM_REQUIRE(!(what & 2) || d_M_vals, "assemble_lattice: M requested but d_M_vals is null");
    PSFEM_REQUIRE(d_nodes && n >= 2 && m >= 2 && n * m < (1ll << 31), "assemble_lattice: bad lattice");
    PSFEM_REQUIRE(d_ws && ws_bytes >= 32, "assemble_lattice: workspace too small");
    if (h_bad_element) *h_bad_element = -1;
    Mat mat;
    const double f = E / (1 - nu * nu);
    mat.d00 = f; mat.d01 = nu * f; mat.d22 = (1 - nu) / 2 * f; mat.t = t; mat.rho = rho;
    unsigned long long *bad = static_cast<unsigned long long *>(d_ws);
    unsigned int *counter = reinterpret_cast<unsigned int *>(bad + 1);
    PSFEM_CUDA_CHECK(cudaMemsetAsync(bad, 0xff, sizeof(unsigned long long), stream));
    const char *seg_env = getenv("PSFEM_LATTICE_SEG");   // tuning knobs (read per call: the tests vary them)
    const char *occ_env = getenv("PSFEM_LATTICE_OCC");
    int seg_len = seg_env ? atoi(seg_env) : 64;
    if (seg_len < 1) seg_len = 64;
    const int occ = occ_env ? atoi(occ_env) : 2;
    int rc = PSFEM_OK;
    if (what & 1) {
        rc = occ == 3 ? launch_lattice_q4<true, 4, 3>(d_nodes, mat, (int)n, (int)m, seg_len, d_K_vals, bad, counter, stream)
                      : launch_lattice_q4<true, 4, 2>(d_nodes, mat, (int)n, (int)m, seg_len, d_K_vals, bad, counter, stream);
        if (rc) return rc;
    }
    if (what & 2) {
        rc = launch_lattice_q4<false, 4, 3>(d_nodes, mat, (int)n, (int)m, seg_len, d_M_vals, bad, counter, stream);
        if (rc) return rc;
    }
    unsigned long long hb = 0;
    PSFEM_CUDA_CHECK(cudaMemcpyAsync(&hb, bad, sizeof(hb), cudaMemcpyDeviceToHost, stream));
    PSFEM_CUDA_CHECK(cudaStreamSynchronize(stream));
    if (hb != ~0ull) {
        if (h_bad_element) *h_bad_element = (int64_t)hb;
        set_error("Jacobian determinant near zero for element %llu", hb);
        return PSFEM_EJACOBIAN;
    }
    return PSFEM_OK;
}
