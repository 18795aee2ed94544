// psfem_lattice.cu -- symbolic and numeric assembly for meshes with the LATTICE TOPOLOGY of Polygones.mesh (T3, Q4).
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

// ---- arithmetic symbolic phase -----------------------------------------------------------------
// Neighbour stencil of a lattice node (i,j): Q4 couples the full 3x3 block; the two triangles per cell of
// Polygones.py:83-85 couple all but the (-1,-1) and (+1,+1) diagonals.  Columns ascend with the node id i*m+j,
// so the slots are visited di-major.  pairs_before(i,j) is the closed-form prefix of the slot counts.
template <int ETYPE> struct Stencil {
    // number of neighbours in node column i+di (di = -1, 0, +1) of the node in row j
    __host__ __device__ static int count(int di, int j, int m) {
        const int jm = j > 0, jp = j < m - 1;
        if (ETYPE == PSFEM_Q4 || di == 0) return 1 + jm + jp;
        return di < 0 ? 1 + jp : 1 + jm;   // T3: (-1,0),(-1,+1) / (+1,-1),(+1,0)
    }
    __host__ __device__ static bool has(int di, int dj) {
        return ETYPE == PSFEM_Q4 || !((di < 0 && dj < 0) || (di > 0 && dj > 0));
    }
    // sum of count(di, j') over j' < j (0 <= j <= m)
    __host__ __device__ static long long prefix(int di, int j, int m) {
        if (ETYPE == PSFEM_Q4 || di == 0) return 3ll * j - (j > 0) - (j == m);
        return di < 0 ? 2ll * j - (j == m) : 2ll * j - (j > 0);
    }
    __host__ __device__ static long long column_total(int i, int n, int m) {
        return (i > 0 ? prefix(-1, m, m) : 0) + prefix(0, m, m) + (i < n - 1 ? prefix(1, m, m) : 0);
    }
    __host__ __device__ static long long pairs_before(int i, int j, int n, int m) {
        // whole columns 0..i-1: column 0 and column n-1 miss one neighbour column
        long long full = 0;
        if (i > 0) {
            const long long inner = prefix(-1, m, m) + prefix(0, m, m) + prefix(1, m, m);
            full = column_total(0, n, m) + (long long)(i - 1) * inner;
        }
        return full + (i > 0 ? prefix(-1, j, m) : 0) + prefix(0, j, m) + (i < n - 1 ? prefix(1, j, m) : 0);
    }
};

// One thread per node; the column indices of a warp's 32 nodes are one contiguous piece of col_idx (up to 32 x 144 B), built in
// shared memory in the global layout and written out with coalesced 16-byte stores (per-thread 8-byte stores at a stride of
// 144 B ran at 0.95 TB/s: 2.67 ms for the 2.5 GB of the 4096 x 4096 plate).
template <int ETYPE>
__global__ void __launch_bounds__(128) lattice_pattern_kernel(int n, int m, int32_t *__restrict__ row_ptr, int32_t *__restrict__ col_idx,
                                                              int32_t *__restrict__ node_ptr) {
    __shared__ __align__(16) int32_t stage[4][32 * 36];     // at most 9 neighbours x 4 indices per node
    const int64_t a = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t n_nodes = (int64_t)n * m;
    const int lane = threadIdx.x & 31;
    int32_t *buf = stage[threadIdx.x >> 5];
    using S = Stencil<ETYPE>;
    if (a == n_nodes) {   // closing entries
        const long long tot = S::pairs_before(n - 1, m, n, m);
        node_ptr[a] = (int32_t)tot;
        row_ptr[2 * a] = (int32_t)(4 * tot);
    }
    const bool node = a < n_nodes;
    long long p0 = 0;
    int deg = 0, i = 0, j = 0;
    if (node) {
        i = (int)(a / m); j = (int)(a % m);
        p0 = S::pairs_before(i, j, n, m);
#pragma unroll
        for (int di = -1; di <= 1; ++di)
            if (i + di >= 0 && i + di < n) deg += S::count(di, j, m);
        node_ptr[a] = (int32_t)p0;
        *reinterpret_cast<int2 *>(row_ptr + 2 * a) = make_int2((int32_t)(4 * p0), (int32_t)(4 * p0 + 2 * deg));
    }
    if (!__shfl_sync(FULL, node ? 1 : 0, 0)) return;        // the whole warp is past the last node
    const long long base = __shfl_sync(FULL, p0, 0);        // consecutive nodes own consecutive pieces of col_idx
    if (node) {
        int2 *r0 = reinterpret_cast<int2 *>(buf + 4 * (int)(p0 - base));
        int2 *r1 = r0 + deg;
        int q = 0;
#pragma unroll
        for (int di = -1; di <= 1; ++di)
#pragma unroll
            for (int dj = -1; dj <= 1; ++dj) {
                if (!S::has(di, dj) || i + di < 0 || i + di >= n || j + dj < 0 || j + dj >= m) continue;
                const int b = (int)(a + (int64_t)di * m + dj);
                const int2 c = make_int2(2 * b, 2 * b + 1);
                r0[q] = c;
                r1[q] = c;
                ++q;
            }
    }
    int pairs = deg;                                         // 4 indices = one int4 per pair
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) pairs += __shfl_xor_sync(FULL, pairs, o);
    __syncwarp();
    int4 *dst = reinterpret_cast<int4 *>(col_idx + 4 * base);
    const int4 *src = reinterpret_cast<const int4 *>(buf);
    for (int q = lane; q < pairs; q += 32) dst[q] = src[q];
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

// Symmetric-lattice output (SL_OUT): instead of the CSR node blocks, a finished node writes its diagonal block and its four
// forward blocks -- 19 doubles -- straight into the structure-of-arrays storage of psfem_spmv_sl.cuh / psfem_cg_sl.cuh,
// sv[((ci - c_lo) * 19 + k) * mpad + j], for the node columns [c_lo, c_lo + S).  Same registers, same sums, same bits as
// psfem_spmv_sl_build on the CSR values; no CSR is ever written.
struct SlTarget { double *sv; int c_lo, S, mpad; };
__device__ __forceinline__ void sl_put3(double *o, size_t mp, const KB &b) { o[0] = b.x; o[mp] = b.y; o[2 * mp] = b.w; }
__device__ __forceinline__ void sl_put3(double *o, size_t mp, const MB &b) { o[0] = b.x; o[mp] = 0.0; o[2 * mp] = b.x; }
__device__ __forceinline__ void sl_put4(double *o, size_t mp, const KB &b) { o[0] = b.x; o[mp] = b.y; o[2 * mp] = b.z; o[3 * mp] = b.w; }
__device__ __forceinline__ void sl_put4(double *o, size_t mp, const MB &b) { o[0] = b.x; o[mp] = 0.0; o[2 * mp] = 0.0; o[3 * mp] = b.x; }

template <bool IS_K, int WARPS, int MINB, bool SL_OUT = false>
__global__ void __launch_bounds__(WARPS * 32, MINB)
lattice_q4_kernel(const double2 *__restrict__ nodes, Mat mat, int n, int m, int seg_len, int n_seg, int n_strips,
                  double *__restrict__ vals, unsigned long long *__restrict__ bad, unsigned int *__restrict__ counter,
                  const SlTarget sl = SlTarget{nullptr, 0, 0, 0}) {
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
                if constexpr (SL_OUT) {
                    const int sc = ci - sl.c_lo;
                    if (node_out && sc >= 0 && sc < sl.S) {
                        const size_t mp = (size_t)sl.mpad;
                        double *o = sl.sv + ((size_t)sc * 19) * mp + jl;
                        sl_put3(o, mp, acc[4]);               // diagonal block: D00, D01, D11
                        sl_put4(o + 3 * mp, mp, acc[5]);      // F1 -> (i, j+1)
                        sl_put4(o + 7 * mp, mp, acc[6]);      // F2 -> (i+1, j-1)
                        sl_put4(o + 11 * mp, mp, acc[7]);     // F3 -> (i+1, j)
                        sl_put4(o + 15 * mp, mp, acc[8]);     // F4 -> (i+1, j+1)
                        if (jl == m - 1)                      // pad rows m .. mpad-1 of this column (never used; kept finite)
                            for (int r = 1; r < sl.mpad - jl; ++r)
#pragma unroll
                                for (int k = 0; k < 19; ++k) o[(size_t)k * mp + r] = 0.0;
                    }
                } else {
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

// ---- T3: two triangles per cell, A = [n00, n10, n01] (element 2c) and B = [n11, n01, n10] (2c+1), Polygones.py:83-85.
// Same marching scheme as Q4; the node rows a cell contributes to are
//   n00: A row 0            n01 (the lane above): A row 2, then B row 1
//   n10: A row 1, B row 2   n11 (the lane above): B row 0
// and every entry is summed in ascending element id: cell(i-1,j-1).B, cell(i-1,j).A, .B, cell(i,j-1).A, .B, cell(i,j).A.
template <bool IS_K, int WARPS, int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB)
lattice_t3_kernel(const double2 *__restrict__ nodes, Mat mat, int n, int m, int seg_len, int n_seg, int n_strips,
                  double *__restrict__ vals, unsigned int *__restrict__ counter) {
    using B = typename BlockOf<IS_K>::type;
    using S = Stencil<PSFEM_T3>;
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
        const int s = item % n_strips, g = item / n_strips;
        const int ia = g * seg_len, ib = min(ia + seg_len, n);
        const int js = STRIP * s, je = min(js + STRIP, m);
        const int jl = js - 1 + lane;
        const bool el_j_ok = jl >= 0 && jl < M;
        const bool node_out = lane >= 1 && jl < m;
        const bool has_jm = jl > 0, has_jp = jl < M;
        const int jn = max(jl, js);
        const int64_t jc0 = min(max(jl, 0), M), jc1 = min(max(jl + 1, 0), M);
        int ci = max(ia - 1, 0);
        double2 p0 = nodes[(int64_t)ci * m + jc0], p3 = nodes[(int64_t)ci * m + jc1];
        double2 p1 = nodes[(int64_t)min(ci + 1, N) * m + jc0], p2 = nodes[(int64_t)min(ci + 1, N) * m + jc1];
        B acc[9];
#pragma unroll
        for (int q = 0; q < 9; ++q) zero(acc[q]);
        for (; ci < ib; ++ci) {
            const int64_t cpf = min(ci + 2, N);
            const double2 f1 = nodes[cpf * m + jc0], f2 = nodes[cpf * m + jc1];
            asm volatile("prefetch.global.L1 [%0];" ::"l"(nodes + (int64_t)min(ci + 3, N) * m + jc0));
            if (lane == 31) asm volatile("prefetch.global.L1 [%0];" ::"l"(nodes + (int64_t)min(ci + 3, N) * m + jc1));
            const bool el_ok = el_j_ok && ci < N;
            ElemT3<IS_K, !IS_K> ea, eb;
            {
                // lanes without a cell integrate unit triangles of zero thickness: exact zeros
                const double2 q0 = el_ok ? p0 : make_double2(0.0, 0.0), q1 = el_ok ? p1 : make_double2(1.0, 0.0);
                const double2 q2 = el_ok ? p2 : make_double2(1.0, 1.0), q3 = el_ok ? p3 : make_double2(0.0, 1.0);
                Mat ml = mat;
                ml.t = el_ok ? mat.t : 0.0;
                const double2 pa[3] = {q0, q1, q3}, pb[3] = {q2, q3, q1};
                ea.compute(pa, ml);
                eb.compute(pb, ml);
            }
            if (ci >= ia) {   // ---- finish node column ci ----
                const B a20 = up1(blk<IS_K>(ea, 2, 0, mat)), a21 = up1(blk<IS_K>(ea, 2, 1, mat)), a22 = up1(blk<IS_K>(ea, 2, 2, mat));
                const B b10 = up1(blk<IS_K>(eb, 1, 0, mat)), b11 = up1(blk<IS_K>(eb, 1, 1, mat)), b12 = up1(blk<IS_K>(eb, 1, 2, mat));
                add(acc[3], a20);
                add(acc[4], a22); add(acc[4], b11); add(acc[4], blk<IS_K>(ea, 0, 0, mat));
                add(acc[5], blk<IS_K>(ea, 0, 2, mat));
                acc[6] = sum(a21, b12);
                acc[7] = sum(b10, blk<IS_K>(ea, 0, 1, mat));
                if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                __syncwarp();
                const bool has_im = ci > 0, has_ip = ci < N;
                double2 *buf = reinterpret_cast<double2 *>(wbuf + parity * LBUF_BYTES);
                long long q_node = S::prefix(0, jn, m) - S::prefix(0, js, m), q_all = S::prefix(0, je, m) - S::prefix(0, js, m);
                if (has_im) { q_node += S::prefix(-1, jn, m) - S::prefix(-1, js, m); q_all += S::prefix(-1, je, m) - S::prefix(-1, js, m); }
                if (has_ip) { q_node += S::prefix(1, jn, m) - S::prefix(1, js, m); q_all += S::prefix(1, je, m) - S::prefix(1, js, m); }
                if (node_out) {
                    const int deg = (has_im ? S::count(-1, jl, m) : 0) + S::count(0, jl, m) + (has_ip ? S::count(1, jl, m) : 0);
                    double2 *row0 = buf + 2 * q_node;
                    int q = 0;
                    if (has_im) {
                        put(row0, deg, q++, acc[1]);
                        if (has_jp) put(row0, deg, q++, acc[2]);
                    }
                    if (has_jm) put(row0, deg, q++, acc[3]);
                    put(row0, deg, q++, acc[4]);
                    if (has_jp) put(row0, deg, q++, acc[5]);
                    if (has_ip) {
                        if (has_jm) put(row0, deg, q++, acc[6]);
                        put(row0, deg, q++, acc[7]);
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) {
                    const long long pair0 = S::pairs_before(ci, js, n, m);
                    const uint32_t bytes = (uint32_t)q_all * 32u;
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                                 ::"l"(vals + 4 * pair0), "r"(l_smem_u32(buf)), "r"(bytes) : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
                parity ^= 1;
            }
            {   // ---- open node column ci+1 ----
                const B b00 = up1(blk<IS_K>(eb, 0, 0, mat)), b01 = up1(blk<IS_K>(eb, 0, 1, mat)), b02 = up1(blk<IS_K>(eb, 0, 2, mat));
                acc[1] = sum(b01, blk<IS_K>(ea, 1, 0, mat));
                acc[2] = sum(blk<IS_K>(ea, 1, 2, mat), blk<IS_K>(eb, 2, 1, mat));
                acc[3] = b02;
                acc[4] = sum(sum(b00, blk<IS_K>(ea, 1, 1, mat)), blk<IS_K>(eb, 2, 2, mat));
                acc[5] = blk<IS_K>(eb, 2, 0, mat);
            }
            p0 = p1; p3 = p2; p1 = f1; p2 = f2;
        }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <bool IS_K, int WARPS, int MINB>
int launch_lattice_t3(const double *d_nodes, const Mat &mat, int n, int m, int seg_len, double *vals, unsigned int *counter,
                      cudaStream_t stream) {
    auto kern = lattice_t3_kernel<IS_K, WARPS, MINB>;
    const size_t smem = (size_t)WARPS * 2 * LBUF_BYTES;
    PSFEM_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int dev = 0, sms = 148, per_sm = 1;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    PSFEM_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, WARPS * 32, smem));
    if (per_sm < 1) per_sm = 1;
    const int n_strips = (int)ceil_div(m, STRIP);
    int n_seg = (int)ceil_div(n, seg_len);
    seg_len = (int)ceil_div(n, n_seg);
    n_seg = (int)ceil_div(n, seg_len);
    const int64_t n_items = (int64_t)n_seg * n_strips;
    int64_t grid = (int64_t)sms * per_sm;
    if (grid * WARPS > n_items) grid = ceil_div(n_items, WARPS);
    PSFEM_CUDA_CHECK(cudaMemsetAsync(counter, 0, sizeof(unsigned int), stream));
    kern<<<(unsigned)grid, WARPS * 32, smem, stream>>>(reinterpret_cast<const double2 *>(d_nodes), mat, n, m, seg_len, n_seg, n_strips, vals, counter);
    PSFEM_LAUNCH_CHECK();
    return PSFEM_OK;
}

template <bool IS_K, int WARPS, int MINB, bool SL_OUT = false>
int launch_lattice_q4(const double *d_nodes, const Mat &mat, int n, int m, int seg_len, double *vals, unsigned long long *bad,
                      unsigned int *counter, cudaStream_t stream, SlTarget sl = SlTarget{nullptr, 0, 0, 0}) {
    auto kern = lattice_q4_kernel<IS_K, WARPS, MINB, SL_OUT>;
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
                                                       vals, bad, counter, sl);
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

extern "C" int psfem_lattice_pattern(int64_t n, int64_t m, int etype, int32_t *d_row_ptr, int32_t *d_col_idx,
                                     int32_t *d_node_ptr, int64_t *h_n_pairs, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    PSFEM_REQUIRE(etype == PSFEM_Q4 || etype == PSFEM_T3, "lattice_pattern: T3 and Q4 only");
    PSFEM_REQUIRE(n >= 2 && m >= 2 && n * m < (1ll << 30) && h_n_pairs, "lattice_pattern: bad lattice");
    const long long np = etype == PSFEM_Q4 ? Stencil<PSFEM_Q4>::pairs_before((int)n - 1, (int)m, (int)n, (int)m)
                                           : Stencil<PSFEM_T3>::pairs_before((int)n - 1, (int)m, (int)n, (int)m);
    PSFEM_REQUIRE(4 * np < (long long)INT32_MAX, "lattice_pattern: more than 2^31 non-zeros on one GPU: partition the mesh");
    *h_n_pairs = np;
    if (!d_row_ptr) return PSFEM_OK;   // size query
    PSFEM_REQUIRE(d_col_idx && d_node_ptr, "lattice_pattern: null output");
    const int T = 128;
    const unsigned grid = (unsigned)ceil_div(n * m + 1, T);
    if (etype == PSFEM_Q4) lattice_pattern_kernel<PSFEM_Q4><<<grid, T, 0, stream>>>((int)n, (int)m, d_row_ptr, d_col_idx, d_node_ptr);
    else lattice_pattern_kernel<PSFEM_T3><<<grid, T, 0, stream>>>((int)n, (int)m, d_row_ptr, d_col_idx, d_node_ptr);
    PSFEM_LAUNCH_CHECK();
    return PSFEM_OK;
}

extern "C" int psfem_assemble_lattice(const double *d_nodes, int64_t n, int64_t m, int etype,
                                      double E, double nu, double rho, double t, int what,
                                      double *d_K_vals, double *d_M_vals, void *d_ws, size_t ws_bytes,
                                      int64_t *h_bad_element, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    PSFEM_REQUIRE(etype == PSFEM_Q4 || etype == PSFEM_T3, "assemble_lattice: T3 and Q4 only");
    PSFEM_REQUIRE(what >= 1 && what <= 3, "assemble_lattice: what must be 1 (K), 2 (M) or 3 (both)");
    PSFEM_REQUIRE(!(what & 1) || d_K_vals, "assemble_lattice: K requested but d_K_vals is null");
    PSFEM_REQUIRE(!(what & 2) || d_M_vals, "assemble_lattice: M requested but d_M_vals is null");
    PSFEM_REQUIRE(d_nodes && n >= 2 && m >= 2 && n * m < (1ll << 31), "assemble_lattice: bad lattice");
    PSFEM_REQUIRE(d_ws && ws_bytes >= 32, "assemble_lattice: workspace too small");
    if (h_bad_element) *h_bad_element = -1;
    const Mat mat = make_mat(E, nu, rho, t, etype);
    unsigned long long *bad = static_cast<unsigned long long *>(d_ws);
    unsigned int *counter = reinterpret_cast<unsigned int *>(bad + 1);
    PSFEM_CUDA_CHECK(cudaMemsetAsync(bad, 0xff, sizeof(unsigned long long), stream));
    const char *seg_env = getenv("PSFEM_LATTICE_SEG");   // tuning knobs (read per call: the tests vary them)
    const char *occ_env = getenv("PSFEM_LATTICE_OCC");
    int seg_len = seg_env ? atoi(seg_env) : 64;
    if (seg_len < 1) seg_len = 64;
    const int occ = occ_env ? atoi(occ_env) : 2;
    int rc = PSFEM_OK;
    if (etype == PSFEM_T3) {   // closed form, no Jacobian guard in the reference (Polygones.py:875-883): nothing to report
        if (what & 1) rc = launch_lattice_t3<true, 4, 3>(d_nodes, mat, (int)n, (int)m, seg_len, d_K_vals, counter, stream);
        if (!rc && (what & 2)) rc = launch_lattice_t3<false, 4, 3>(d_nodes, mat, (int)n, (int)m, seg_len, d_M_vals, counter, stream);
        return rc;
    }
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

// Numeric assembly of a Q4 lattice STRAIGHT INTO the symmetric-lattice operand of the PCG (no CSR values, no column
// indices, no reduction copy): the node columns [c_lo, c_lo + S) of the lattice (n node columns, m nodes per column),
// 19 doubles per node.  c_lo = 1 drops a clamped first column (K[np.ix_(free, free)] of Polygones.py:1003 for
// boundary('left', [0, 0], ...): the forward blocks of the remaining columns are untouched by the reduction); in a strip of
// the multi-GPU partition the first stored column is the halo column before the owned ones.
extern "C" int psfem_assemble_lattice_sl(const double *d_nodes, int64_t n, int64_t m, int etype, double E, double nu, double rho, double t,
                                         int what, int64_t c_lo, int64_t S, double *d_sl_vals, size_t bytes, void *d_ws, size_t ws_bytes,
                                         int64_t *h_bad_element, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    PSFEM_REQUIRE(etype == PSFEM_Q4, "assemble_lattice_sl: Q4 lattices only");
    PSFEM_REQUIRE(what == 1 || what == 2, "assemble_lattice_sl: what must be 1 (K) or 2 (M)");
    PSFEM_REQUIRE(d_nodes && d_sl_vals && n >= 2 && m >= 2 && n * m < (1ll << 31) && c_lo >= 0 && S >= 1 && c_lo + S <= n,
                  "assemble_lattice_sl: bad lattice / column range");
    PSFEM_REQUIRE(d_ws && ws_bytes >= 32, "assemble_lattice_sl: workspace too small");
    const int64_t mpad = (m + 3) & ~3ll;
    const size_t want = (size_t)S * 19 * (size_t)mpad * sizeof(double);
    PSFEM_REQUIRE(bytes >= want, "assemble_lattice_sl: storage too small (%zu < %zu)", bytes, want);
    if (h_bad_element) *h_bad_element = -1;
    const Mat mat = make_mat(E, nu, rho, t, etype);
    unsigned long long *bad = static_cast<unsigned long long *>(d_ws);
    unsigned int *counter = reinterpret_cast<unsigned int *>(bad + 1);
    PSFEM_CUDA_CHECK(cudaMemsetAsync(bad, 0xff, sizeof(unsigned long long), stream));
    const char *seg_env = getenv("PSFEM_LATTICE_SEG");
    int seg_len = seg_env ? atoi(seg_env) : 64;
    if (seg_len < 1) seg_len = 64;
    SlTarget sl{d_sl_vals, (int)c_lo, (int)S, (int)mpad};
    int rc = what == 1 ? launch_lattice_q4<true, 4, 2, true>(d_nodes, mat, (int)n, (int)m, seg_len, nullptr, bad, counter, stream, sl)
                       : launch_lattice_q4<false, 4, 3, true>(d_nodes, mat, (int)n, (int)m, seg_len, nullptr, bad, counter, stream, sl);
    if (rc) return rc;
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
