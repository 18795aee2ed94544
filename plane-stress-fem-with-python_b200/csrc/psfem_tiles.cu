// psfem_tiles.cu -- tile-fused numeric assembly (v2) for T3 and Q4.
//
// v1 (psfem_assemble.cu) writes every element matrix to HBM and reads it back in the segmented
// reduction: 1.4 kB of traffic per Q4 element against 320 B algorithmic, and a 4x redundant Jacobian
// set-up (profiles/r1a_summary.md).  Here a CTA owns a TILE of node rows that are close in space
// (consecutive nodes of a Morton ordering of the coordinates, so the tiling works for any mesh, not
// only structured ones):
//   phase 1  one thread per element touching the tile: Ke (symmetric 2x2-block form) / Me computed ONCE
//            per tile into shared memory (elements on tile borders are recomputed by the neighbour tile:
//            (9*17)/(8*16) = 1.2x for a structured Q4 mesh and 128-node tiles);
//   phase 2  one thread per CSR node pair of the tile: sums its contributions from shared memory in
//            ascending element id (same deterministic order as v1 / as Polygones.py:717) and writes the
//            four CSR values exactly once, in runs of consecutive rows (coalesced).
// No atomics, no intermediate HBM round trip.  The kernel is persistent (2 CTAs per SM) and software
// pipelined: while tile i is computed, the coordinates of tile i+1 are gathered into shared memory with
// cp.async, its plan (pair counts, contribution codes, row offsets) arrives by bulk async copies (TMA)
// on an mbarrier, and the connectivity of tile i+2 / the header of tile i+3 are already in registers --
// so no global-memory latency sits on the critical path.  The symbolic side (psfem_tileplan_build)
// derives the tiles and stores each tile's plan contiguously and 16-byte aligned for those copies.
#include <cstdlib>
#include <cub/cub.cuh>
#include "psfem_common.cuh"

using namespace psfem;

namespace {

// ---------------------------------------------------------------------------------------------
// tile plan (symbolic)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long enc_ordered(double d) {
    unsigned long long b = (unsigned long long)__double_as_longlong(d);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double dec_ordered(unsigned long long e) {
    unsigned long long b = (e >> 63) ? (e & 0x7fffffffffffffffull) : ~e;
    return __longlong_as_double((long long)b);
}

__global__ void bbox_kernel(const double2 *__restrict__ nodes, int64_t n, unsigned long long *__restrict__ box /* minx,miny,maxx,maxy */) {
    double mnx = 1e300, mny = 1e300, mxx = -1e300, mxy = -1e300;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double2 p = nodes[i];
        mnx = fmin(mnx, p.x); mny = fmin(mny, p.y); mxx = fmax(mxx, p.x); mxy = fmax(mxy, p.y);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mnx = fmin(mnx, __shfl_xor_sync(0xffffffffu, mnx, o)); mny = fmin(mny, __shfl_xor_sync(0xffffffffu, mny, o));
        mxx = fmax(mxx, __shfl_xor_sync(0xffffffffu, mxx, o)); mxy = fmax(mxy, __shfl_xor_sync(0xffffffffu, mxy, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(box + 0, enc_ordered(mnx)); atomicMin(box + 1, enc_ordered(mny));
        atomicMax(box + 2, enc_ordered(mxx)); atomicMax(box + 3, enc_ordered(mxy));
    }
}

__device__ __forceinline__ uint32_t spread16(uint32_t v) {
    v &= 0xffff;
    v = (v | (v << 8)) & 0x00ff00ff;
    v = (v | (v << 4)) & 0x0f0f0f0f;
    v = (v | (v << 2)) & 0x33333333;
    v = (v | (v << 1)) & 0x55555555;
    return v;
}

// Morton key of every node; the bounding box is squared so that tiles are square in space
__global__ void morton_kernel(const double2 *__restrict__ nodes, int64_t n, const unsigned long long *__restrict__ box,
                              uint32_t *__restrict__ keys, int32_t *__restrict__ ids) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double mnx = dec_ordered(box[0]), mny = dec_ordered(box[1]), mxx = dec_ordered(box[2]), mxy = dec_ordered(box[3]);
    double ext = fmax(mxx - mnx, mxy - mny);
    if (!(ext > 0)) ext = 1.0;
    const double sc = 65535.0 / ext;
    double2 p = nodes[i];
    uint32_t qx = (uint32_t)fmin(fmax((p.x - mnx) * sc, 0.0), 65535.0);
    uint32_t qy = (uint32_t)fmin(fmax((p.y - mny) * sc, 0.0), 65535.0);
    keys[i] = spread16(qx) | (spread16(qy) << 1);
    ids[i] = (int32_t)i;
}

__global__ void tile_of_kernel(const int32_t *__restrict__ order, int64_t n, int R, int32_t *__restrict__ tile_of,
                               int32_t *__restrict__ ids) {
    int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (p >= n) return;
    tile_of[order[p]] = (int32_t)(p / R);
    ids[p] = (int32_t)p;  // identity values for the second (stable) sort
}

template <int NPE>
__global__ void incidence_keys_kernel(const int32_t *__restrict__ conn, int64_t nel, const int32_t *__restrict__ tile_of,
                                      unsigned long long *__restrict__ keys) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= nel * NPE) return;
    int64_t e = t / NPE;
    keys[t] = ((unsigned long long)(uint32_t)tile_of[conn[t]] << 32) | (uint32_t)e;
}

__global__ void tile_elem_ptr_kernel(const unsigned long long *__restrict__ ukeys, int32_t n_unique, int32_t n_tiles,
                                     int32_t *__restrict__ tile_elem_ptr, int *__restrict__ max_elems) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t > n_tiles) return;
    unsigned long long key = (unsigned long long)t << 32;
    int32_t lo = 0, hi = n_unique;
    while (lo < hi) {
        int32_t mid = lo + ((hi - lo) >> 1);
        if (ukeys[mid] < key) lo = mid + 1; else hi = mid;
    }
    tile_elem_ptr[t] = lo;
}

__global__ void unpack_elems_kernel(const unsigned long long *__restrict__ ukeys, int32_t n, int32_t *__restrict__ tile_elems) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) tile_elems[i] = (int32_t)(ukeys[i] & 0xffffffffull);
}

// per tile-ordered node u (a = tile_nodes[u]): number of CSR pairs (degree) and of contributions
__global__ void tnode_counts_kernel(int64_t n_nodes, const int32_t *__restrict__ tile_nodes, const int32_t *__restrict__ node_ptr,
                                    const int32_t *__restrict__ pair_ptr, int32_t *__restrict__ deg, int32_t *__restrict__ ncodes) {
    int64_t u = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (u > n_nodes) return;
    if (u == n_nodes) { deg[u] = 0; ncodes[u] = 0; return; }
    const int a = tile_nodes[u];
    const int p0 = node_ptr[a], p1 = node_ptr[a + 1];
    deg[u] = p1 - p0;
    ncodes[u] = pair_ptr[p1] - pair_ptr[p0];
}

// padded per-tile sizes (pairs to a multiple of 16 bytes of uint8, codes to 16 bytes of uint16) + maxima
__global__ void tile_sizes_kernel(int32_t n_tiles, int64_t n_nodes, int R, const int32_t *__restrict__ gp, const int32_t *__restrict__ gc,
                                  const int32_t *__restrict__ tile_elem_ptr, int32_t *__restrict__ qsz, int32_t *__restrict__ csz,
                                  int *__restrict__ maxima) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t > n_tiles) return;
    if (t == n_tiles) { qsz[t] = 0; csz[t] = 0; return; }
    const int64_t u0 = (int64_t)t * R, u1 = (u0 + R < n_nodes) ? u0 + R : n_nodes;
    const int np = gp[u1] - gp[u0], nc = gc[u1] - gc[u0];
    qsz[t] = (np + 15) & ~15;
    csz[t] = (nc + 7) & ~7;
    atomicMax(maxima + 0, tile_elem_ptr[t + 1] - tile_elem_ptr[t]);
    atomicMax(maxima + 1, np);
    atomicMax(maxima + 2, nc);
}

__global__ void tile_hdr_kernel(int32_t n_tiles, int64_t n_nodes, int R, const int32_t *__restrict__ gp, const int32_t *__restrict__ gc,
                                const int32_t *__restrict__ tile_elem_ptr, const int32_t *__restrict__ q0, const int32_t *__restrict__ c0,
                                int32_t *__restrict__ hdr) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tiles) return;
    const int64_t u0 = (int64_t)t * R, u1 = (u0 + R < n_nodes) ? u0 + R : n_nodes;
    int32_t *h = hdr + 8ll * t;
    h[0] = tile_elem_ptr[t]; h[1] = tile_elem_ptr[t + 1] - tile_elem_ptr[t];
    h[2] = q0[t]; h[3] = gp[u1] - gp[u0];
    h[4] = c0[t]; h[5] = gc[u1] - gc[u0];
    h[6] = (int)(u1 - u0); h[7] = 0;
}

// tile-ordered compact plan of one node: nmeta = {first CSR pair of the node, pair prefix inside the tile},
// tp_cnt[q] = contributions of pair q (uint8), tcodes = slot*16 + i*npe + j (uint16)
__global__ void tmeta_fill_kernel(int64_t n_nodes, int npe2, int R, const int32_t *__restrict__ tile_nodes,
                                  const int32_t *__restrict__ node_ptr, const int32_t *__restrict__ pair_ptr,
                                  const int32_t *__restrict__ perm, const int32_t *__restrict__ tile_elem_ptr,
                                  const int32_t *__restrict__ tile_elems, const int32_t *__restrict__ gp, const int32_t *__restrict__ gc,
                                  const int32_t *__restrict__ q0, const int32_t *__restrict__ c0, int2 *__restrict__ nmeta,
                                  uint8_t *__restrict__ tp_cnt, uint16_t *__restrict__ tcodes, int *__restrict__ flags) {
    int64_t u = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (u >= n_nodes) return;
    const int t = (int)(u / R);
    const int64_t u0 = (int64_t)t * R;
    const int a = tile_nodes[u];
    const int e0 = tile_elem_ptr[t], e1 = tile_elem_ptr[t + 1];
    const int p0 = node_ptr[a], p1 = node_ptr[a + 1];
    nmeta[u] = make_int2(p0, gp[u] - gp[u0]);
    int q = q0[t] + (gp[u] - gp[u0]), c = c0[t] + (gc[u] - gc[u0]);
    for (int s = p0; s < p1; ++s) {
        const int k0 = pair_ptr[s], k1 = pair_ptr[s + 1];
        if (k1 - k0 > 255) atomicExch(flags, 1);
        tp_cnt[q++] = (uint8_t)(k1 - k0);
        for (int k = k0; k < k1; ++k) {
            const int code = perm[k];
            const int e = code / npe2, ij = code - e * npe2;
            int lo = e0, hi = e1;
            while (lo < hi) {
                int mid = (lo + hi) >> 1;
                if (tile_elems[mid] < e) lo = mid + 1; else hi = mid;
            }
            tcodes[c++] = (uint16_t)((lo - e0) * 16 + ij);
        }
    }
}

template <int NPE>
__global__ void tile_conn_kernel(int32_t n, const int32_t *__restrict__ tile_elems, const int32_t *__restrict__ conn, int4 *__restrict__ tile_conn) {
    int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int32_t *c = conn + (int64_t)tile_elems[k] * NPE;
    tile_conn[k] = make_int4(c[0], c[1], c[2], NPE == 4 ? c[NPE - 1] : 0);
}

struct TileWs {
    unsigned long long *box, *k64a, *k64b;
    uint32_t *k32a, *k32b;
    int32_t *v32a, *v32b, *tile_of, *deg, *ncd, *gp, *gc, *qsz, *csz, *q0, *c0;
    int *scal;
    void *cub_tmp;
    size_t cub_bytes, total;
};

size_t tile_cub_bytes(int64_t n_nodes, int64_t ninc) {
    size_t a = 0, b = 0, c = 0, d = 0;
    cub::DoubleBuffer<uint32_t> k32(nullptr, nullptr);
    cub::DoubleBuffer<int32_t> v32(nullptr, nullptr);
    cub::DeviceRadixSort::SortPairs(nullptr, a, k32, v32, (int)n_nodes, 0, 32);
    cub::DoubleBuffer<unsigned long long> k64(nullptr, nullptr);
    cub::DeviceRadixSort::SortKeys(nullptr, b, k64, (int)ninc, 0, 64);
    cub::DeviceSelect::Unique(nullptr, c, (unsigned long long *)nullptr, (unsigned long long *)nullptr, (int *)nullptr, (int)ninc);
    cub::DeviceScan::ExclusiveSum(nullptr, d, (int32_t *)nullptr, (int32_t *)nullptr, (int)n_nodes + 1);
    size_t r = a > b ? a : b;
    r = r > c ? r : c;
    return align_up(r > d ? r : d) + 256;
}

TileWs carve_tiles(void *ws, int64_t n_nodes, int64_t ninc, int64_t n_tiles) {
    TileWs w;
    Carver c(ws);
    w.box = c.take<unsigned long long>(4);
    w.k64a = c.take<unsigned long long>(ninc);
    w.k64b = c.take<unsigned long long>(ninc);
    w.k32a = c.take<uint32_t>(n_nodes);
    w.k32b = c.take<uint32_t>(n_nodes);
    w.v32a = c.take<int32_t>(n_nodes);
    w.v32b = c.take<int32_t>(n_nodes);
    w.tile_of = c.take<int32_t>(n_nodes);
    w.deg = c.take<int32_t>(n_nodes + 1);
    w.ncd = c.take<int32_t>(n_nodes + 1);
    w.gp = c.take<int32_t>(n_nodes + 1);
    w.gc = c.take<int32_t>(n_nodes + 1);
    w.qsz = c.take<int32_t>(n_tiles + 1);
    w.csz = c.take<int32_t>(n_tiles + 1);
    w.q0 = c.take<int32_t>(n_tiles + 1);
    w.c0 = c.take<int32_t>(n_tiles + 1);
    w.scal = c.take<int>(8);
    w.cub_bytes = tile_cub_bytes(n_nodes, ninc);
    w.cub_tmp = c.take<char>(w.cub_bytes);
    w.total = c.used();
    return w;
}

// ---------------------------------------------------------------------------------------------
// numeric kernel
// ---------------------------------------------------------------------------------------------
struct Mat {
    double d00, d01, d22, t, rho;
};

// upper-triangular block index of (i,j), i<=j, for NPE nodes
template <int NPE> __host__ __device__ constexpr int tri_index(int i, int j) { return i * NPE - (i * (i - 1)) / 2 + (j - i); }

template <int NPE> struct TileLayout {
    static constexpr int NB = NPE * (NPE + 1) / 2;                  // stored blocks per element
    static constexpr int K_STRIDE = NB * 4 + 2;                      // doubles per element (padded: odd number of 16-byte units)
    static constexpr int M_STRIDE = NB + 1;
};

// Q4 element stiffness / mass in symmetric block form.  Same arithmetic as gauss_kernel (Polygones.py:582-595,
// :657-708, :1266-1291) with the element computed once: per Gauss point the 36 gradient products
// Gxx(i<=j), Gyy(i<=j), Gxy(i,j) are accumulated, D is applied once at the end.
template <bool DO_K, bool DO_M>
__device__ __forceinline__ bool q4_element(const double2 (&p)[4], const Mat &mat, double *__restrict__ kout, double *__restrict__ mout) {
    constexpr double A = 0.7745966692414834;  // sqrt(0.6)  (Polygones.py:549)
    const double gp[3] = {-A, 0.0, A};
    const double gw[3] = {5.0 / 9, 8.0 / 9, 5.0 / 9};
    const double sx[4] = {-1, 1, 1, -1}, sy[4] = {-1, -1, 1, 1};
    double gxx[10], gyy[10], gxy[16], mm[10];
#pragma unroll
    for (int q = 0; q < 10; ++q) { gxx[q] = 0; gyy[q] = 0; mm[q] = 0; }
#pragma unroll
    for (int q = 0; q < 16; ++q) gxy[q] = 0;
    bool ok = true;
#pragma unroll
    for (int r = 0; r < 3; ++r) {      // eta (outer, Polygones.py:551-553)
#pragma unroll
        for (int c = 0; c < 3; ++c) {  // xi
            const double xi = gp[c], eta = gp[r];
            double dnx[4], dne[4], nn[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                dnx[i] = sx[i] * (1 + sy[i] * eta) / 4;
                dne[i] = sy[i] * (1 + sx[i] * xi) / 4;
                nn[i] = (1 + sx[i] * xi) * (1 + sy[i] * eta) / 4;
            }
            double j00 = 0, j01 = 0, j10 = 0, j11 = 0;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                j00 += dnx[i] * p[i].x; j01 += dnx[i] * p[i].y;
                j10 += dne[i] * p[i].x; j11 += dne[i] * p[i].y;
            }
            const double detJ = j00 * j11 - j01 * j10;
            const double ad = fabs(detJ);
            if (ad < 1e-10) ok = false;  // Polygones.py:686-687
            const double w = gw[r] * gw[c];
            if (DO_K) {
                const double inv = 1.0 / detJ;
                const double pp = j11 * inv, qq = -j01 * inv, rr = -j10 * inv, ss = j00 * inv;
                const double cc = w * ad * mat.t;
                double dx[4], dy[4], cx[4], cy[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    dx[i] = pp * dnx[i] + qq * dne[i];
                    dy[i] = rr * dnx[i] + ss * dne[i];
                    cx[i] = cc * dx[i];
                    cy[i] = cc * dy[i];
                }
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        if (j >= i) {
                            gxx[tri_index<4>(i, j)] += cx[i] * dx[j];
                            gyy[tri_index<4>(i, j)] += cy[i] * dy[j];
                        }
                        gxy[i * 4 + j] += cx[i] * dy[j];
                    }
            }
            if (DO_M) {
                const double cm = w * mat.rho * mat.t * ad;
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = i; j < 4; ++j) mm[tri_index<4>(i, j)] += (cm * nn[i]) * nn[j];
            }
        }
    }
    if (DO_K) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = i; j < 4; ++j) {
                const int b = tri_index<4>(i, j);
                double *o = kout + 4 * b;
                o[0] = mat.d00 * gxx[b] + mat.d22 * gyy[b];                    // K(2i,2j)
                o[1] = mat.d01 * gxy[i * 4 + j] + mat.d22 * gxy[j * 4 + i];    // K(2i,2j+1)
                o[2] = mat.d01 * gxy[j * 4 + i] + mat.d22 * gxy[i * 4 + j];    // K(2i+1,2j)
                o[3] = mat.d00 * gyy[b] + mat.d22 * gxx[b];                    // K(2i+1,2j+1)
            }
    }
    if (DO_M) {
#pragma unroll
        for (int b = 0; b < 10; ++b) mout[b] = mm[b];
    }
    return ok;
}

// T3 closed form (Polygones.py:840-883, :1218-1235), signed area
template <bool DO_K, bool DO_M>
__device__ __forceinline__ bool t3_element(const double2 (&p)[3], const Mat &mat, double *__restrict__ kout, double *__restrict__ mout) {
    const double delta = p[0].x * (p[1].y - p[2].y) + p[1].x * (p[2].y - p[0].y) + p[2].x * (p[0].y - p[1].y);
    const double vol = 0.5 * delta * mat.t;
    if (DO_K) {
        double b[3], g[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            b[i] = (p[(i + 1) % 3].y - p[(i + 2) % 3].y) / delta;
            g[i] = (p[(i + 2) % 3].x - p[(i + 1) % 3].x) / delta;
        }
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = i; j < 3; ++j) {
                double *o = kout + 4 * tri_index<3>(i, j);
                o[0] = (mat.d00 * b[i] * b[j] + mat.d22 * g[i] * g[j]) * vol;
                o[1] = (mat.d01 * b[i] * g[j] + mat.d22 * g[i] * b[j]) * vol;
                o[2] = (mat.d01 * g[i] * b[j] + mat.d22 * b[i] * g[j]) * vol;
                o[3] = (mat.d00 * g[i] * g[j] + mat.d22 * b[i] * b[j]) * vol;
            }
    }
    if (DO_M) {
        const double base = mat.rho * vol / 12;
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = i; j < 3; ++j) mout[tri_index<3>(i, j)] = (i == j ? 2.0 : 1.0) * base;
    }
    return true;
}

struct TilePlanDev {  // device pointers of the tile plan (see psfem_tileplan_build)
    const int4 *hdr;        // 2 x int4 per tile: {e0, ne, q0, npairs}, {c0, ncodes, nn, 0}
    const int4 *tile_conn;  // connectivity of every tile element (tile order)
    const int2 *nmeta;      // n_tiles*R: {first CSR pair of the node, pair prefix inside the tile}
    const int32_t *tile_elems;  // element id of every tile element (error reporting only)
    const uint8_t *tp_cnt;
    const uint16_t *tcodes;
    int R, max_elems, max_pairs, max_codes;
};

__device__ __forceinline__ uint32_t t_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void t_cp_async16(void *dst, const void *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(t_smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void t_bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(t_smem_u32(dst)), "l"(src), "r"(bytes), "r"(t_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void t_mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(t_smem_u32(bar)), "r"(parity) : "memory");
    } while (!ok);
}

struct TileHdr { int e0, ne, q0, npairs, c0, ncodes, nn, pad; };

template <int ETYPE, int NPE, bool DO_K, bool DO_M, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB)
assemble_tile_kernel(const double2 *__restrict__ nodes, Mat mat, int n_tiles, TilePlanDev tp, double *__restrict__ K_vals,
                     double *__restrict__ M_vals, unsigned long long *__restrict__ bad) {
    using L = TileLayout<NPE>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ uint64_t meta_bar[2];
    __shared__ int s_wsum[THREADS / 32 + 1];
    const int tid = threadIdx.x, R = tp.R;
    // ---- shared-memory carve-up (sizes from the plan's per-tile maxima; every region 16-byte aligned) ----
    const int cap_e = (tp.max_elems + 1) & ~1, cap_p = (tp.max_pairs + 15) & ~15, cap_c = (tp.max_codes + 7) & ~7;
    unsigned char *sp = smem_raw;
    double *kb = reinterpret_cast<double *>(sp); sp += DO_K ? (size_t)cap_e * L::K_STRIDE * 8 : 0;
    double *mb = reinterpret_cast<double *>(sp); sp += DO_M ? (((size_t)cap_e * L::M_STRIDE * 8 + 15) & ~(size_t)15) : 0;
    double2 *xy_s[2]; int2 *nm_s[2]; uint8_t *cnt_s[2]; uint16_t *cod_s[2];
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        xy_s[s] = reinterpret_cast<double2 *>(sp); sp += (size_t)THREADS * NPE * 16;
        nm_s[s] = reinterpret_cast<int2 *>(sp); sp += (size_t)R * 8;
        cnt_s[s] = sp; sp += cap_p;
        cod_s[s] = reinterpret_cast<uint16_t *>(sp); sp += (size_t)cap_c * 2;
    }
    int *s_coff = reinterpret_cast<int *>(sp); sp += (size_t)(cap_p + 4) * 4;
    uint8_t *s_slot = sp;

    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(t_smem_u32(&meta_bar[0])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(t_smem_u32(&meta_bar[1])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int n_mine = (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    auto tile_id = [&](int i) { return (int)blockIdx.x + i * (int)gridDim.x; };
    auto load_hdr = [&](int i) -> TileHdr {
        TileHdr h = {0, 0, 0, 0, 0, 0, 0, 0};
        if (i < n_mine) {
            const int4 a = tp.hdr[2 * tile_id(i)], b = tp.hdr[2 * tile_id(i) + 1];
            h.e0 = a.x; h.ne = a.y; h.q0 = a.z; h.npairs = a.w; h.c0 = b.x; h.ncodes = b.y; h.nn = b.z;
        }
        return h;
    };
    auto load_conn = [&](const TileHdr &h) -> int4 {
        return (tid < h.ne) ? tp.tile_conn[h.e0 + tid] : make_int4(0, 0, 0, 0);
    };
    // coordinates of the first THREADS elements of a tile -> stage (cp.async), its plan -> stage (bulk copies)
    auto prefetch = [&](int i, int stage, const TileHdr &h, const int4 &c) {
        if (i < n_mine) {
            if (tid < h.ne) {
                double2 *dst = xy_s[stage] + (size_t)tid * NPE;
                t_cp_async16(dst + 0, nodes + c.x);
                t_cp_async16(dst + 1, nodes + c.y);
                t_cp_async16(dst + 2, nodes + c.z);
                if (NPE == 4) t_cp_async16(dst + 3, nodes + c.w);
            }
            if (tid == 0) {
                const uint32_t b_nm = (uint32_t)R * 8, b_cnt = (uint32_t)((h.npairs + 15) & ~15), b_cod = (uint32_t)((h.ncodes + 7) & ~7) * 2;
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(t_smem_u32(&meta_bar[stage])), "r"(b_nm + b_cnt + b_cod) : "memory");
                t_bulk_g2s(nm_s[stage], tp.nmeta + (size_t)tile_id(i) * R, b_nm, &meta_bar[stage]);
                if (b_cnt) t_bulk_g2s(cnt_s[stage], tp.tp_cnt + h.q0, b_cnt, &meta_bar[stage]);
                if (b_cod) t_bulk_g2s(cod_s[stage], tp.tcodes + h.c0, b_cod, &meta_bar[stage]);
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    // ---- prologue: tile 0 staged, connectivity of tile 1 and headers up to tile 2 in registers ----
    TileHdr h0 = load_hdr(0), h1 = load_hdr(1), h2 = load_hdr(2);
    int4 conn_a = load_conn(h0);
    prefetch(0, 0, h0, conn_a);
    conn_a = load_conn(h1);

    for (int i = 0; i < n_mine; ++i) {
        const int st = i & 1;
        // ---- software pipeline: inputs of tile i+1, connectivity of tile i+2, header of tile i+3 ----
        prefetch(i + 1, st ^ 1, h1, conn_a);
        const int4 conn_b = load_conn(h2);
        const TileHdr h3 = load_hdr(i + 3);
        asm volatile("cp.async.wait_group 1;" ::: "memory");          // everything but the newest group: tile i is in
        t_mbar_wait(&meta_bar[st], (uint32_t)((i >> 1) & 1));
        __syncthreads();
        const int ne = h0.ne, nn = h0.nn, npairs = h0.npairs;
        const double2 *xy = xy_s[st];
        const int2 *nm = nm_s[st];
        const uint8_t *cnt = cnt_s[st];
        const uint16_t *cod = cod_s[st];
        // ---- phase 1: element matrices of the tile, each computed once ----
        for (int slot = tid; slot < ne; slot += THREADS) {
            double2 p[NPE];
            if (slot < THREADS) {
#pragma unroll
                for (int q = 0; q < NPE; ++q) p[q] = xy[(size_t)slot * NPE + q];
            } else {  // rare: a tile with more elements than threads reads the overflow directly
                const int4 c = tp.tile_conn[h0.e0 + slot];
                p[0] = nodes[c.x]; p[1] = nodes[c.y]; p[2] = nodes[c.z];
                if (NPE == 4) p[NPE - 1] = nodes[c.w];
            }
            bool ok;
            if (ETYPE == PSFEM_Q4)
                ok = q4_element<DO_K, DO_M>(reinterpret_cast<const double2(&)[4]>(p), mat, kb + (size_t)slot * L::K_STRIDE, mb + (size_t)slot * L::M_STRIDE);
            else
                ok = t3_element<DO_K, DO_M>(reinterpret_cast<const double2(&)[3]>(p), mat, kb + (size_t)slot * L::K_STRIDE, mb + (size_t)slot * L::M_STRIDE);
            if (!ok) atomicMin(bad, (unsigned long long)tp.tile_elems[h0.e0 + slot]);   // lowest failing element id, like Polygones.py:686-687
        }
        // node slot of every pair
        for (int s = tid; s < nn; s += THREADS) {
            const int b0 = nm[s].y, b1 = (s + 1 < nn) ? nm[s + 1].y : npairs;
            for (int q = b0; q < b1; ++q) s_slot[q] = (uint8_t)s;
        }
        // exclusive scan of the pair counts -> offset of every pair's codes (chunk per thread + block scan)
        {
            const int chunk = (npairs + THREADS - 1) / THREADS;
            const int b0 = min(tid * chunk, npairs), b1 = min(b0 + chunk, npairs);
            int sum = 0;
            for (int q = b0; q < b1; ++q) sum += cnt[q];
            int v = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int u = __shfl_up_sync(0xffffffffu, v, o);
                if ((tid & 31) >= o) v += u;
            }
            if ((tid & 31) == 31) s_wsum[tid >> 5] = v;
            __syncthreads();
            if (tid == 0) {
                int run = 0;
                for (int w = 0; w < THREADS / 32; ++w) { int x = s_wsum[w]; s_wsum[w] = run; run += x; }
            }
            __syncthreads();
            int off = s_wsum[tid >> 5] + v - sum;
            for (int q = b0; q < b1; ++q) { s_coff[q] = off; off += cnt[q]; }
        }
        __syncthreads();
        // ---- phase 2: one thread per CSR node pair, everything but the output in shared memory ----
        for (int idx = tid; idx < npairs; idx += THREADS) {
            const int s = s_slot[idx];
            const int2 m = nm[s];
            const int pos = idx - m.y;
            const int deg = ((s + 1 < nn) ? nm[s + 1].y : npairs) - m.y;
            const int k0 = s_coff[idx], k1 = k0 + cnt[idx];
            double a0 = 0, a1 = 0, a2 = 0, a3 = 0, mm = 0;
            for (int k = k0; k < k1; ++k) {  // ascending element id
                const int code = cod[k];
                const int slot = code >> 4, ij = code & 15;
                const int ii = ij / NPE, jj = ij - ii * NPE;
                const bool up = ii <= jj;
                const int b = up ? tri_index<NPE>(ii, jj) : tri_index<NPE>(jj, ii);
                if (DO_K) {
                    const double2 *src = reinterpret_cast<const double2 *>(kb + (size_t)slot * L::K_STRIDE + 4 * b);
                    const double2 u = src[0], v = src[1];
                    a0 += u.x; a3 += v.y;
                    a1 += up ? u.y : v.x;   // transpose of the stored block when i > j
                    a2 += up ? v.x : u.y;
                }
                if (DO_M) mm += mb[(size_t)slot * L::M_STRIDE + b];
            }
            const int64_t r0 = 4ll * m.x + 2 * pos, r1 = r0 + 2 * deg;
            if (DO_K) {
                *reinterpret_cast<double2 *>(K_vals + r0) = make_double2(a0, a1);
                *reinterpret_cast<double2 *>(K_vals + r1) = make_double2(a2, a3);
            }
            if (DO_M) {
                *reinterpret_cast<double2 *>(M_vals + r0) = make_double2(mm, 0.0);
                *reinterpret_cast<double2 *>(M_vals + r1) = make_double2(0.0, mm);
            }
        }
        __syncthreads();   // kb / s_coff / stage st are free again
        h0 = h1; h1 = h2; h2 = h3; conn_a = conn_b;
    }
}

template <int NPE, bool DO_K, bool DO_M> size_t tile_smem_bytes(const TilePlanDev &tp, int threads) {
    using L = TileLayout<NPE>;
    const size_t cap_e = (tp.max_elems + 1) & ~1, cap_p = (tp.max_pairs + 15) & ~15, cap_c = (tp.max_codes + 7) & ~7;
    size_t b = DO_K ? cap_e * L::K_STRIDE * 8 : 0;
    b += DO_M ? ((cap_e * L::M_STRIDE * 8 + 15) & ~(size_t)15) : 0;
    b += 2 * ((size_t)threads * NPE * 16 + (size_t)tp.R * 8 + cap_p + cap_c * 2);
    b += (cap_p + 4) * 4 + cap_p + 16;
    return b;
}

template <int ETYPE, int NPE, bool DO_K, bool DO_M, int THREADS, int MINB>
int launch_tiles(const double *d_nodes, Mat mat, int64_t n_tiles, const TilePlanDev &tp, double *K_vals, double *M_vals,
                 unsigned long long *bad, cudaStream_t stream) {
    const size_t smem = tile_smem_bytes<NPE, DO_K, DO_M>(tp, THREADS);
    PSFEM_REQUIRE(smem <= 225 * 1024, "assemble_tiled: tile needs %zu bytes of shared memory", smem);
    auto kern = assemble_tile_kernel<ETYPE, NPE, DO_K, DO_M, THREADS, MINB>;
    PSFEM_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int dev = 0, sms = 148, per_sm = 1;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    PSFEM_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, THREADS, smem));
    if (per_sm < 1) per_sm = 1;
    int64_t grid = (int64_t)sms * per_sm;
    if (grid > n_tiles) grid = n_tiles;
    kern<<<(unsigned)grid, THREADS, smem, stream>>>(reinterpret_cast<const double2 *>(d_nodes), mat, (int)n_tiles, tp, K_vals, M_vals, bad);
    PSFEM_LAUNCH_CHECK();
    return PSFEM_OK;
}

}  // namespace

extern "C" int psfem_tileplan_workspace(int64_t nel, int npe, int64_t n_nodes, int tile_nodes_R, size_t *bytes) {
    PSFEM_REQUIRE(bytes && nel >= 0 && n_nodes > 0 && npe > 0 && tile_nodes_R >= 32, "tileplan_workspace: bad arguments");
    PSFEM_REQUIRE(nel * npe < (int64_t)INT32_MAX, "tileplan: nel*npe exceeds int32");
    *bytes = carve_tiles(nullptr, n_nodes, nel * npe > 0 ? nel * npe : 1, ceil_div(n_nodes, tile_nodes_R)).total;
    return PSFEM_OK;
}

extern "C" int psfem_tileplan_build(const double *d_nodes, const int32_t *d_conn, int64_t nel, int npe, int64_t n_nodes,
                                    int tile_nodes_R, int64_t n_pairs, const int32_t *d_node_ptr, const int32_t *d_pair_ptr,
                                    const int32_t *d_perm, int32_t *d_tile_nodes /* n_nodes */,
                                    int32_t *d_tile_elem_ptr /* n_tiles+1 */, int32_t *d_tile_elems /* nel*npe (upper bound) */,
                                    int32_t *d_tile_hdr /* 8*n_tiles */, int32_t *d_tile_conn /* 4*nel*npe (upper bound) */,
                                    int32_t *d_nmeta /* 2*n_tiles*R */, uint8_t *d_tp_cnt /* n_pairs + 16*n_tiles */,
                                    uint16_t *d_tcodes /* nel*npe^2 + 8*n_tiles */,
                                    int64_t *h_info /* tile elements, max elems, max pairs, max codes, usable, mean elems */,
                                    void *d_ws, size_t ws_bytes, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    PSFEM_REQUIRE(npe == 3 || npe == 4, "tileplan_build: tiled assembly covers T3 and Q4");
    PSFEM_REQUIRE(tile_nodes_R >= 32 && tile_nodes_R <= 256 && tile_nodes_R % 2 == 0, "tileplan_build: tile size must be even and in [32, 256]");
    PSFEM_REQUIRE(h_info, "tileplan_build: null output");
    const int64_t ninc = nel * npe;
    PSFEM_REQUIRE(ninc > 0 && ninc < (int64_t)INT32_MAX, "tileplan_build: bad element count");
    PSFEM_REQUIRE(nel * npe * npe + 8 * ceil_div(n_nodes, tile_nodes_R) < (int64_t)INT32_MAX, "tileplan_build: too many contributions");
    const int R = tile_nodes_R;
    const int64_t n_tiles = ceil_div(n_nodes, R);
    TileWs w = carve_tiles(d_ws, n_nodes, ninc, n_tiles);
    PSFEM_REQUIRE(ws_bytes >= w.total, "tileplan_build: workspace too small");
    const int T = 256;
    const double2 *nodes = reinterpret_cast<const double2 *>(d_nodes);
    // 1. bounding box -> Morton keys -> node order
    unsigned long long init[4] = {~0ull, ~0ull, 0ull, 0ull};
    PSFEM_CUDA_CHECK(cudaMemcpyAsync(w.box, init, sizeof(init), cudaMemcpyHostToDevice, stream));
    bbox_kernel<<<148 * 4, T, 0, stream>>>(nodes, n_nodes, w.box);
    PSFEM_LAUNCH_CHECK();
    morton_kernel<<<(unsigned)ceil_div(n_nodes, T), T, 0, stream>>>(nodes, n_nodes, w.box, w.k32a, w.v32a);
    PSFEM_LAUNCH_CHECK();
    cub::DoubleBuffer<uint32_t> k32(w.k32a, w.k32b);
    cub::DoubleBuffer<int32_t> v32(w.v32a, w.v32b);
    size_t tb = w.cub_bytes;
    PSFEM_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(w.cub_tmp, tb, k32, v32, (int)n_nodes, 0, 32, stream));
    // 2. tile of every node; tile_nodes = node ids sorted by (tile, id): stable sort of the identity by tile
    int32_t *ids = v32.Alternate();
    tile_of_kernel<<<(unsigned)ceil_div(n_nodes, T), T, 0, stream>>>(v32.Current(), n_nodes, R, w.tile_of, ids);
    PSFEM_LAUNCH_CHECK();
    int tile_bits = 1;
    while ((1ll << tile_bits) < n_tiles) ++tile_bits;
    {
        PSFEM_CUDA_CHECK(cudaMemcpyAsync(k32.Current(), w.tile_of, n_nodes * sizeof(int32_t), cudaMemcpyDeviceToDevice, stream));
        cub::DoubleBuffer<uint32_t> tk2(k32.Current(), k32.Alternate());
        cub::DoubleBuffer<int32_t> tv2(ids, v32.Current());
        tb = w.cub_bytes;
        PSFEM_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(w.cub_tmp, tb, tk2, tv2, (int)n_nodes, 0, tile_bits, stream));
        PSFEM_CUDA_CHECK(cudaMemcpyAsync(d_tile_nodes, tv2.Current(), n_nodes * sizeof(int32_t), cudaMemcpyDeviceToDevice, stream));
    }
    // 3. per-tile unique element lists
    unsigned grid = (unsigned)ceil_div(ninc, T);
    if (npe == 3) incidence_keys_kernel<3><<<grid, T, 0, stream>>>(d_conn, nel, w.tile_of, w.k64a);
    else incidence_keys_kernel<4><<<grid, T, 0, stream>>>(d_conn, nel, w.tile_of, w.k64a);
    PSFEM_LAUNCH_CHECK();
    cub::DoubleBuffer<unsigned long long> k64(w.k64a, w.k64b);
    int ebits = 1;
    while ((1ll << ebits) < nel) ++ebits;
    tb = w.cub_bytes;
    PSFEM_CUDA_CHECK(cub::DeviceRadixSort::SortKeys(w.cub_tmp, tb, k64, (int)ninc, 0, ebits, stream));
    tb = w.cub_bytes;
    PSFEM_CUDA_CHECK(cub::DeviceRadixSort::SortKeys(w.cub_tmp, tb, k64, (int)ninc, 32, 32 + tile_bits, stream));
    unsigned long long *sorted = k64.Current(), *uniq = k64.Alternate();
    tb = w.cub_bytes;
    PSFEM_CUDA_CHECK(cub::DeviceSelect::Unique(w.cub_tmp, tb, sorted, uniq, w.scal, (int)ninc, stream));
    int h_unique = 0;
    PSFEM_CUDA_CHECK(cudaMemcpyAsync(&h_unique, w.scal, sizeof(int), cudaMemcpyDeviceToHost, stream));
    PSFEM_CUDA_CHECK(cudaStreamSynchronize(stream));
    tile_elem_ptr_kernel<<<(unsigned)ceil_div(n_tiles + 1, T), T, 0, stream>>>(uniq, h_unique, (int32_t)n_tiles, d_tile_elem_ptr, nullptr);
    PSFEM_LAUNCH_CHECK();
    unpack_elems_kernel<<<(unsigned)ceil_div(h_unique, T), T, 0, stream>>>(uniq, h_unique, d_tile_elems);
    PSFEM_LAUNCH_CHECK();
    if (npe == 3) tile_conn_kernel<3><<<(unsigned)ceil_div(h_unique, T), T, 0, stream>>>(h_unique, d_tile_elems, d_conn, reinterpret_cast<int4 *>(d_tile_conn));
    else tile_conn_kernel<4><<<(unsigned)ceil_div(h_unique, T), T, 0, stream>>>(h_unique, d_tile_elems, d_conn, reinterpret_cast<int4 *>(d_tile_conn));
    PSFEM_LAUNCH_CHECK();
    // 4. tile-ordered, padded plan: global prefixes per tile-ordered node, padded per-tile bases, fill
    PSFEM_CUDA_CHECK(cudaMemsetAsync(w.scal + 2, 0, 6 * sizeof(int), stream));
    tnode_counts_kernel<<<(unsigned)ceil_div(n_nodes + 1, T), T, 0, stream>>>(n_nodes, d_tile_nodes, d_node_ptr, d_pair_ptr, w.deg, w.ncd);
    PSFEM_LAUNCH_CHECK();
    tb = w.cub_bytes;
    PSFEM_CUDA_CHECK(cub::DeviceScan::ExclusiveSum(w.cub_tmp, tb, w.deg, w.gp, (int)n_nodes + 1, stream));
    tb = w.cub_bytes;
    PSFEM_CUDA_CHECK(cub::DeviceScan::ExclusiveSum(w.cub_tmp, tb, w.ncd, w.gc, (int)n_nodes + 1, stream));
    tile_sizes_kernel<<<(unsigned)ceil_div(n_tiles + 1, T), T, 0, stream>>>((int32_t)n_tiles, n_nodes, R, w.gp, w.gc, d_tile_elem_ptr, w.qsz, w.csz, w.scal + 2);
    PSFEM_LAUNCH_CHECK();
    tb = w.cub_bytes;
    PSFEM_CUDA_CHECK(cub::DeviceScan::ExclusiveSum(w.cub_tmp, tb, w.qsz, w.q0, (int)n_tiles + 1, stream));
    tb = w.cub_bytes;
    PSFEM_CUDA_CHECK(cub::DeviceScan::ExclusiveSum(w.cub_tmp, tb, w.csz, w.c0, (int)n_tiles + 1, stream));
    tile_hdr_kernel<<<(unsigned)ceil_div(n_tiles, T), T, 0, stream>>>((int32_t)n_tiles, n_nodes, R, w.gp, w.gc, d_tile_elem_ptr, w.q0, w.c0, d_tile_hdr);
    PSFEM_LAUNCH_CHECK();
    PSFEM_CUDA_CHECK(cudaMemsetAsync(d_nmeta, 0, (size_t)n_tiles * R * 8, stream));
    tmeta_fill_kernel<<<(unsigned)ceil_div(n_nodes, T), T, 0, stream>>>(n_nodes, npe * npe, R, d_tile_nodes, d_node_ptr, d_pair_ptr, d_perm,
                                                                       d_tile_elem_ptr, d_tile_elems, w.gp, w.gc, w.q0, w.c0,
                                                                       reinterpret_cast<int2 *>(d_nmeta), d_tp_cnt, d_tcodes, w.scal + 5);
    PSFEM_LAUNCH_CHECK();
    int h[4] = {0, 0, 0, 0};
    PSFEM_CUDA_CHECK(cudaMemcpyAsync(h, w.scal + 2, sizeof(h), cudaMemcpyDeviceToHost, stream));
    PSFEM_CUDA_CHECK(cudaStreamSynchronize(stream));
    h_info[0] = h_unique; h_info[1] = h[0]; h_info[2] = h[1]; h_info[3] = h[2];
    h_info[4] = (h[3] == 0 && h[0] <= 4095) ? 1 : 0;   // a pair with > 255 contributions / a huge tile: use psfem_assemble
    h_info[5] = ceil_div(h_unique, n_tiles);
    (void)n_pairs;
    return PSFEM_OK;
}

extern "C" int psfem_assemble_tiled(const double *d_nodes, int64_t nel, int etype, int64_t n_nodes,
                                    double E, double nu, double rho, double t, int what, int tile_nodes_R,
                                    const int64_t *h_plan_info, const int32_t *d_tile_hdr, const int32_t *d_tile_conn,
                                    const int32_t *d_tile_elems, const int32_t *d_nmeta, const uint8_t *d_tp_cnt, const uint16_t *d_tcodes,
                                    double *d_K_vals, double *d_M_vals, void *d_ws, size_t ws_bytes,
                                    int64_t *h_bad_element, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    PSFEM_REQUIRE(etype == PSFEM_T3 || etype == PSFEM_Q4, "assemble_tiled: T3 and Q4 only");
    PSFEM_REQUIRE(what >= 1 && what <= 3, "assemble_tiled: what must be 1 (K), 2 (M) or 3 (both)");
    PSFEM_REQUIRE(!(what & 1) || d_K_vals, "assemble_tiled: K requested but d_K_vals is null");
    PSFEM_REQUIRE(!(what & 2) || d_M_vals, "assemble_tiled: M requested but d_M_vals is null");
    PSFEM_REQUIRE(d_ws && ws_bytes >= 16, "assemble_tiled: workspace too small");
    PSFEM_REQUIRE(h_plan_info && h_plan_info[4] == 1, "assemble_tiled: the tile plan is not usable for this mesh");
    if (h_bad_element) *h_bad_element = -1;
    if (nel == 0) return PSFEM_OK;
    Mat mat;
    const double f = E / (1 - nu * nu);
    mat.d00 = f; mat.d01 = nu * f; mat.d22 = (1 - nu) / 2 * f; mat.t = t; mat.rho = rho;
    unsigned long long *bad = static_cast<unsigned long long *>(d_ws);
    PSFEM_CUDA_CHECK(cudaMemsetAsync(bad, 0xff, sizeof(unsigned long long), stream));
    TilePlanDev tp;
    tp.hdr = reinterpret_cast<const int4 *>(d_tile_hdr); tp.tile_conn = reinterpret_cast<const int4 *>(d_tile_conn);
    tp.nmeta = reinterpret_cast<const int2 *>(d_nmeta); tp.tp_cnt = d_tp_cnt; tp.tcodes = d_tcodes; tp.tile_elems = d_tile_elems;
    tp.R = tile_nodes_R; tp.max_elems = (int)h_plan_info[1]; tp.max_pairs = (int)h_plan_info[2]; tp.max_codes = (int)h_plan_info[3];
    const int64_t n_tiles = ceil_div(n_nodes, tp.R);
    const int mean = (int)h_plan_info[5];   // threads follow the TYPICAL tile; bigger tiles take a second round
    int rc;
#define PSFEM_TILE_CASE(ET, NPE, TH, MB)                                                                              \
    (what == 3 ? launch_tiles<ET, NPE, true, true, TH, MB>(d_nodes, mat, n_tiles, tp, d_K_vals, d_M_vals, bad, stream) \
     : what == 1 ? launch_tiles<ET, NPE, true, false, TH, MB>(d_nodes, mat, n_tiles, tp, d_K_vals, d_M_vals, bad, stream) \
                 : launch_tiles<ET, NPE, false, true, TH, MB>(d_nodes, mat, n_tiles, tp, d_K_vals, d_M_vals, bad, stream))
    if (etype == PSFEM_Q4) {
        if (mean <= 96) rc = PSFEM_TILE_CASE(PSFEM_Q4, 4, 96, 3);
        else if (mean <= 160) rc = PSFEM_TILE_CASE(PSFEM_Q4, 4, 160, 2);
        else rc = PSFEM_TILE_CASE(PSFEM_Q4, 4, 320, 1);
    } else {
        if (mean <= 160) rc = PSFEM_TILE_CASE(PSFEM_T3, 3, 160, 3);
        else rc = PSFEM_TILE_CASE(PSFEM_T3, 3, 320, 2);
    }
#undef PSFEM_TILE_CASE
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
