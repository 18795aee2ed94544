// psfem_tiles.cu -- tile-fused numeric assembly (v2) for T3 and Q4.
//
// v1 (psfem_assemble.cu) writes every element matrix to HBM and reads it back in the segmented
// reduction: 1.4 kB of traffic per Q4 element against 320 B algorithmic, and a 4x redundant Jacobian
// set-up (profiles/r1a_summary.md).  Here a CTA owns a TILE of node rows that are close in space
// (consecutive nodes of a Morton ordering of the coordinates, so the tiling works for any mesh, not
// only structured ones):
//   phase 1  one thread per element touching the tile: the gradient products of Ke / Me are computed ONCE
//            per tile and stay in registers (elements on tile borders are recomputed by the neighbour
//            tile: (9*17)/(8*16) = 1.2x for a structured Q4 mesh and 128-node tiles);
//   scatter  the tile's block of CSR values lives in shared memory; elements add their matrix rows into it
//            in ROUNDS: round r of a node is its r-th incident element (ascending id), so all rows added
//            in one round belong to distinct nodes -- plain read-modify-write is race-free, every thread
//            is busy in every round, and each entry is summed in ascending element id (reference order);
//   copy-out the finished rows go to HBM exactly once, in runs of consecutive rows (coalesced).
// No atomics, no intermediate HBM round trip.  The kernel is persistent (2 CTAs per SM) and software
// pipelined: while tile i is computed, the coordinates of tile i+1 are gathered into shared memory with
// cp.async, its plan (scatter offsets, row offsets) arrives by bulk async copies (TMA)
// on an mbarrier, and the connectivity of tile i+2 / the header of tile i+3 are already in registers --
// so no global-memory latency sits on the critical path.  The symbolic side (psfem_tileplan_build)
// derives the tiles and stores each tile's plan contiguously and 16-byte aligned for those copies.
#include <cstdlib>
#include <type_traits>
#include <cub/cub.cuh>
#include "psfem_common.cuh"
#include "psfem_elem.cuh"

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

// per tile-ordered node u (a = tile_nodes[u]): number of CSR pairs (degree)
__global__ void tnode_counts_kernel(int64_t n_nodes, const int32_t *__restrict__ tile_nodes, const int32_t *__restrict__ node_ptr,
                                    int32_t *__restrict__ deg) {
    int64_t u = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (u > n_nodes) return;
    deg[u] = (u == n_nodes) ? 0 : node_ptr[tile_nodes[u] + 1] - node_ptr[tile_nodes[u]];
}

// nmeta[u] = {first CSR pair of the node (global), pair prefix inside its tile}
__global__ void nmeta_kernel(int64_t n_nodes, int R, const int32_t *__restrict__ tile_nodes, const int32_t *__restrict__ node_ptr,
                             const int32_t *__restrict__ gp, int2 *__restrict__ nmeta) {
    int64_t u = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (u >= n_nodes) return;
    const int64_t u0 = (u / R) * R;
    nmeta[u] = make_int2(node_ptr[tile_nodes[u]], gp[u] - gp[u0]);
}

__device__ __forceinline__ int find_sorted(const int32_t *__restrict__ a, int n, int key) {  // index of key in sorted a[0..n), or -1
    int lo = 0, hi = n;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (a[mid] < key) lo = mid + 1; else hi = mid;
    }
    return (lo < n && a[lo] == key) ? lo : -1;
}

// Scatter rounds of one tile.  Element e adds row i of its matrix (the 2x2 blocks (i, j=0..npe-1)) to the
// CSR rows of node a = conn[e][i].  Rows of the same node must not be added concurrently, so every owned
// node hands out round numbers to its incident (element, i) items in ASCENDING ELEMENT ID: round r of
// node a is its r-th incident element.  All items of one round touch distinct nodes => race-free plain
// read-modify-write, all threads busy in every round (one row per element per round on a structured
// mesh), and each CSR entry is summed in ascending element id -- the order of Polygones.py:717-726.
// The round of (e,i) is packed into the 4 spare high bits of the node id in tile_conn (0xF = row not owned).
template <int NPE>
__global__ void tile_rounds_kernel(int32_t n_tiles, int64_t n_nodes, int R, const int32_t *__restrict__ tile_nodes,
                                   const int32_t *__restrict__ tile_elem_ptr, const int32_t *__restrict__ tile_elems,
                                   const int32_t *__restrict__ conn, const int32_t *__restrict__ gp, int4 *__restrict__ tile_conn,
                                   int32_t *__restrict__ hdr, int *__restrict__ flags) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tiles) return;
    const int64_t u0 = (int64_t)t * R;
    const int nn = (int)((n_nodes - u0) < R ? (n_nodes - u0) : R);
    const int32_t *tn = tile_nodes + u0;
    const int e0 = tile_elem_ptr[t], ne = tile_elem_ptr[t + 1] - e0;
    uint8_t cnt[256];
    for (int s = 0; s < nn; ++s) cnt[s] = 0;
    int nrounds = 0;
    for (int k = 0; k < ne; ++k) {
        const int32_t *c = conn + (int64_t)tile_elems[e0 + k] * NPE;
        int v[4] = {0, 0, 0, 0};
#pragma unroll
        for (int i = 0; i < NPE; ++i) {
            const int s = find_sorted(tn, nn, c[i]);
            int r = 15;
            if (s >= 0) {
                r = cnt[s]++;
                if (r >= 15) { atomicExch(flags, 1); r = 14; }
                if (r + 1 > nrounds) nrounds = r + 1;
            }
            v[i] = (int)((uint32_t)c[i] | ((uint32_t)r << 28));
        }
        tile_conn[e0 + k] = make_int4(v[0], v[1], v[2], v[3]);
    }
    int32_t *h = hdr + 8ll * t;
    h[0] = e0; h[1] = ne; h[2] = nn; h[3] = gp[u0 + nn] - gp[u0]; h[4] = nrounds; h[5] = 0; h[6] = 0; h[7] = 0;
}

// Scatter record of a tile element (32 bytes): the tile's CSR block is kept in shared memory in EXACTLY the
// global layout (per node: row 2a = 2*deg doubles, then row 2a+1), nodes in tile order, so that runs of
// consecutive nodes leave with one bulk copy.  Block (i,j) adds {K(2a,2b),K(2a,2b+1)} at double2 index
// base[i]+pos[i][j] and {K(2a+1,2b),K(2a+1,2b+1)} deg[i] double2 further.
struct __align__(16) TileRec {
    uint16_t base[4];   // 2*prefix of row node i inside the tile (double2 units); unused rows: anything
    uint8_t deg[4];     // CSR pairs of row node i
    uint8_t pos[16];    // position of node j among the column nodes of row node i
    uint8_t pad[4];
};

template <int NPE>
__global__ void tile_scatter_plan_kernel(int32_t n_tiles, int64_t n_nodes, int R, const int32_t *__restrict__ tile_nodes,
                                         const int32_t *__restrict__ tile_elem_ptr, const int32_t *__restrict__ sorted_elems,
                                         const int32_t *__restrict__ conn, const int32_t *__restrict__ node_ptr,
                                         const int32_t *__restrict__ col_idx, const int32_t *__restrict__ gp,
                                         TileRec *__restrict__ tile_rec, int *__restrict__ flags) {
    // one thread per tile element; the tile is found by binary search on tile_elem_ptr
    int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int n_te = tile_elem_ptr[n_tiles];
    if (k >= n_te) return;
    int lo = 0, hi = n_tiles;   // last t with tile_elem_ptr[t] <= k
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (tile_elem_ptr[mid] <= k) lo = mid; else hi = mid;
    }
    const int t = lo;
    const int64_t u0 = (int64_t)t * R;
    const int nn = (int)((n_nodes - u0) < R ? (n_nodes - u0) : R);
    const int32_t *tn = tile_nodes + u0;
    const int32_t *c = conn + (int64_t)sorted_elems[k] * NPE;
    int nd[NPE];
#pragma unroll
    for (int i = 0; i < NPE; ++i) nd[i] = c[i];
    TileRec rec;
#pragma unroll
    for (int q = 0; q < 4; ++q) { rec.base[q] = 0; rec.deg[q] = 0; rec.pad[q] = 0; }
#pragma unroll
    for (int q = 0; q < 16; ++q) rec.pos[q] = 0;
#pragma unroll
    for (int i = 0; i < NPE; ++i) {
        const int s = find_sorted(tn, nn, nd[i]);
        if (s < 0) continue;
        const int pfx = gp[u0 + s] - gp[u0];
        const int p0 = node_ptr[nd[i]], deg = node_ptr[nd[i] + 1] - p0;
        if (deg > 255 || 2 * pfx > 65535) atomicExch(flags, 1);
        rec.base[i] = (uint16_t)(2 * pfx);
        rec.deg[i] = (uint8_t)deg;
#pragma unroll
        for (int j = 0; j < NPE; ++j) {
            // position of node nd[j] among the column nodes of row node nd[i]: columns 2b of DOF row 2a
            int l = 0, h = deg;
            const int32_t *cols = col_idx + 4ll * p0;   // row 2a: 2*deg entries {2b, 2b+1}
            const int want = 2 * nd[j];
            while (l < h) {
                int mid = (l + h) >> 1;
                if (cols[2 * mid] < want) l = mid + 1; else h = mid;
            }
            rec.pos[i * NPE + j] = (uint8_t)l;
        }
    }
    tile_rec[k] = rec;
}

__global__ void tile_max_kernel(int32_t n_tiles, const int32_t *__restrict__ hdr, int *__restrict__ maxima) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tiles) return;
    atomicMax(maxima + 0, hdr[8ll * t + 1]);
    atomicMax(maxima + 1, hdr[8ll * t + 3]);
    atomicMax(maxima + 2, hdr[8ll * t + 4]);
}

struct TileWs {
    unsigned long long *box, *k64a, *k64b;
    uint32_t *k32a, *k32b;
    int32_t *v32a, *v32b, *tile_of, *deg, *gp;
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

TileWs carve_tiles(void *ws, int64_t n_nodes, int64_t ninc) {
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
    w.gp = c.take<int32_t>(n_nodes + 1);
    w.scal = c.take<int>(8);
    w.cub_bytes = tile_cub_bytes(n_nodes, ninc);
    w.cub_tmp = c.take<char>(w.cub_bytes);
    w.total = c.used();
    return w;
}

// ---------------------------------------------------------------------------------------------
// numeric kernel
// ---------------------------------------------------------------------------------------------

struct TilePlanDev {  // device pointers of the tile plan (see psfem_tileplan_build)
    const int4 *hdr;            // 2 x int4 per tile: {e0, ne, nn, npairs}, {nrounds,0,0,0}
    const int4 *tile_conn;      // connectivity of every tile element (ascending id); bits 28..31 = scatter round of row i
    const TileRec *tile_rec;    // scatter record of every tile element
    const int2 *nmeta;          // n_tiles*R: {first CSR pair of the node, pair prefix inside the tile}
    const int32_t *tile_elems;  // element id of every tile element (error reporting only)
    int R, max_elems, max_pairs;
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

struct TileHdr { int e0, ne, nn, npairs, nrounds; };

template <int ETYPE, int NPE, bool DO_K, bool DO_M, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB)
assemble_tile_kernel(const double2 *__restrict__ nodes, Mat mat, int n_tiles, TilePlanDev tp, double *__restrict__ K_vals,
                     double *__restrict__ M_vals, unsigned long long *__restrict__ bad) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ uint64_t meta_bar[2];
    const int tid = threadIdx.x, R = tp.R;
    // ---- shared-memory carve-up (sizes from the plan's per-tile maxima; every region 16-byte aligned).
    // Buffers are addressed as smem_raw + offset (never through an array of pointers) so that the compiler
    // keeps them in the shared address space (LDS/STS instead of generic LD/ST).
    const int cap_e = tp.max_elems, cap_p = (tp.max_pairs + 15) & ~15;
    const uint32_t vbuf_bytes = (uint32_t)cap_p * 32;                       // one tile block of CSR values (K or M)
    const uint32_t nbuf = (DO_K ? 1u : 0u) + (DO_M ? 1u : 0u);
    const uint32_t stage0_off = 2 * nbuf * vbuf_bytes;                       // value blocks are double buffered
    const uint32_t xy_bytes = (uint32_t)THREADS * NPE * 16, nm_bytes = (uint32_t)R * 8, rec_bytes = (uint32_t)cap_e * 32;
    const uint32_t cn_bytes = (uint32_t)THREADS * 16;
    const uint32_t stage_bytes = xy_bytes + nm_bytes + rec_bytes + cn_bytes;
    __shared__ __align__(16) int4 s_hdr[4][2];                               // ring of tile headers
#define PSFEM_KBUF(st) reinterpret_cast<double2 *>(smem_raw + (st) * nbuf * vbuf_bytes)
#define PSFEM_MBUF(st) reinterpret_cast<double2 *>(smem_raw + (st) * nbuf * vbuf_bytes + (DO_K ? vbuf_bytes : 0))
#define PSFEM_STAGE_XY(st) reinterpret_cast<double2 *>(smem_raw + stage0_off + (st) * stage_bytes)
#define PSFEM_STAGE_NM(st) reinterpret_cast<int2 *>(smem_raw + stage0_off + (st) * stage_bytes + xy_bytes)
#define PSFEM_STAGE_REC(st) reinterpret_cast<TileRec *>(smem_raw + stage0_off + (st) * stage_bytes + xy_bytes + nm_bytes)
#define PSFEM_STAGE_CONN(st) reinterpret_cast<int4 *>(smem_raw + stage0_off + (st) * stage_bytes + xy_bytes + nm_bytes + rec_bytes)

    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(t_smem_u32(&meta_bar[0])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(t_smem_u32(&meta_bar[1])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int n_mine = (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    auto tile_id = [&](int i) { return (int)blockIdx.x + i * (int)gridDim.x; };
    auto read_hdr = [&](int i) -> TileHdr {        // header of my i-th tile from the shared ring (zeros past the end)
        TileHdr h = {0, 0, 0, 0, 0};
        if (i < n_mine) {
            const int4 a = s_hdr[i & 3][0], b = s_hdr[i & 3][1];
            h.e0 = a.x; h.ne = a.y; h.nn = a.z; h.npairs = a.w; h.nrounds = b.x;
        }
        return h;
    };
    // Everything below is asynchronous and lands in shared memory: no register is held across a tile.
    auto fetch_hdr = [&](int i) {                  // header of tile i -> ring
        if (i < n_mine && tid < 2) t_cp_async16(&s_hdr[i & 3][tid], tp.hdr + 2 * (size_t)tile_id(i) + tid);
    };
    auto fetch_conn = [&](int i, const TileHdr &h) {   // connectivity of the first THREADS elements of tile i
        if (i < n_mine && tid < h.ne) t_cp_async16(PSFEM_STAGE_CONN(i & 1) + tid, tp.tile_conn + h.e0 + tid);
    };
    auto fetch_inputs = [&](int i, const TileHdr &h) {  // coordinates (needs conn(i) in smem) + plan of tile i
        if (i < n_mine) {
            const int stage = i & 1;
            if (tid < h.ne) {
                const int4 c = PSFEM_STAGE_CONN(stage)[tid];
                double2 *dst = PSFEM_STAGE_XY(stage) + (size_t)tid * NPE;
                t_cp_async16(dst + 0, nodes + (c.x & 0x0fffffff));
                t_cp_async16(dst + 1, nodes + (c.y & 0x0fffffff));
                t_cp_async16(dst + 2, nodes + (c.z & 0x0fffffff));
                if (NPE == 4) t_cp_async16(dst + 3, nodes + (c.w & 0x0fffffff));
            }
            if (tid == 0) {
                const uint32_t b_nm = (uint32_t)R * 8, b_rec = (uint32_t)h.ne * 32;
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(t_smem_u32(&meta_bar[stage])), "r"(b_nm + b_rec) : "memory");
                t_bulk_g2s(PSFEM_STAGE_NM(stage), tp.nmeta + (size_t)tile_id(i) * R, b_nm, &meta_bar[stage]);
                if (b_rec) t_bulk_g2s(PSFEM_STAGE_REC(stage), tp.tile_rec + h.e0, b_rec, &meta_bar[stage]);
            }
        }
    };
    auto commit_wait_sync = [&]() {
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
    };
    auto rounds_of = [](const int4 &c) -> uint32_t {   // 4 x 4-bit scatter rounds of the element's rows
        return ((uint32_t)c.x >> 28) | (((uint32_t)c.y >> 28) << 4) | (((uint32_t)c.z >> 28) << 8) | (((uint32_t)c.w >> 28) << 12);
    };
    // one matrix row of an element -> the tile block: the NPE destinations of a row are distinct CSR pairs,
    // so all loads are issued before the adds and stores (the compiler cannot prove it and would serialise)
    auto scatter_row = [&](double2 *vb, const TileRec &rec, int ii, auto &&block) {
        double2 r0[NPE], r1[NPE];
        int q0[NPE];
#pragma unroll
        for (int jj = 0; jj < NPE; ++jj) {
            q0[jj] = (int)rec.base[ii] + (int)rec.pos[ii * NPE + jj];
            r0[jj] = vb[q0[jj]];
            r1[jj] = vb[q0[jj] + rec.deg[ii]];
        }
#pragma unroll
        for (int jj = 0; jj < NPE; ++jj) {
            const double4 v = block(ii, jj);
            r0[jj].x += v.x; r0[jj].y += v.y; r1[jj].x += v.z; r1[jj].y += v.w;
        }
#pragma unroll
        for (int jj = 0; jj < NPE; ++jj) {
            vb[q0[jj]] = r0[jj];
            vb[q0[jj] + rec.deg[ii]] = r1[jj];
        }
    };

    // ---- prologue (latency exposed once per CTA): headers 0..2, conn 0..1, inputs of tile 0 ----
    fetch_hdr(0); fetch_hdr(1); fetch_hdr(2);
    commit_wait_sync();
    fetch_conn(0, read_hdr(0)); fetch_conn(1, read_hdr(1));
    commit_wait_sync();
    fetch_inputs(0, read_hdr(0));
    asm volatile("cp.async.commit_group;" ::: "memory");

    for (int i = 0; i < n_mine; ++i) {
        const int st = i & 1;
        // inputs of tile i were issued one tile ago; the bulk stores that read value buffer `st` two tiles ago are done
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        t_mbar_wait(&meta_bar[st], (uint32_t)((i >> 1) & 1));
        __syncthreads();
        const TileHdr h0 = read_hdr(i);
        const int ne = h0.ne, nn = h0.nn, npairs = h0.npairs;
        const uint32_t rnd_cur = (tid < ne) ? rounds_of(PSFEM_STAGE_CONN(st)[tid]) : 0xffffu;
        double2 *kbuf = PSFEM_KBUF(st), *mbuf = PSFEM_MBUF(st);
        // zero the tile's value block(s)
        for (int q = tid; q < 2 * npairs; q += THREADS) {
            if (DO_K) kbuf[q] = make_double2(0.0, 0.0);
            if (DO_M) mbuf[q] = make_double2(0.0, 0.0);
        }
        __syncthreads();   // conn(i) read by everyone before its buffer is refilled with conn(i+2); zero-fill done
        // ---- software pipeline: inputs of tile i+1, connectivity of tile i+2, header of tile i+3 ----
        fetch_inputs(i + 1, read_hdr(i + 1));
        fetch_conn(i + 2, read_hdr(i + 2));
        fetch_hdr(i + 3);
        asm volatile("cp.async.commit_group;" ::: "memory");
        const double2 *xy = PSFEM_STAGE_XY(st);
        const int2 *nm = PSFEM_STAGE_NM(st);
        const TileRec *recs = PSFEM_STAGE_REC(st);
        // ---- element pass(es): compute once, scatter row by row in rounds ----
        for (int base = 0; base < ne; base += THREADS) {
            const int slot = base + tid;
            const bool active = slot < ne;
            typename std::conditional<ETYPE == PSFEM_Q4, ElemQ4<DO_K, DO_M>, ElemT3<DO_K, DO_M>>::type el;
            uint32_t rnd = 0xffffu;   // all rows "not owned"
            TileRec rec;
            if (active) {
                double2 p[NPE];
                if (base == 0) {
#pragma unroll
                    for (int q = 0; q < NPE; ++q) p[q] = xy[(size_t)slot * NPE + q];
                    rnd = rnd_cur;
                } else {  // rare: a tile with more elements than threads reads the overflow directly
                    const int4 c = tp.tile_conn[h0.e0 + slot];
                    p[0] = nodes[c.x & 0x0fffffff]; p[1] = nodes[c.y & 0x0fffffff]; p[2] = nodes[c.z & 0x0fffffff];
                    if (NPE == 4) p[NPE - 1] = nodes[c.w & 0x0fffffff];
                    rnd = rounds_of(c);
                }
                rec = recs[slot];
                if (!el.compute(p, mat)) atomicMin(bad, (unsigned long long)tp.tile_elems[h0.e0 + slot]);  // lowest id, Polygones.py:686
            }
            for (int r = 0; r < h0.nrounds; ++r) {
#pragma unroll
                for (int ii = 0; ii < NPE; ++ii) {
                    if (((rnd >> (4 * ii)) & 15u) == (uint32_t)r) {
                        if (DO_K) scatter_row(kbuf, rec, ii, [&](int a, int b) { return el.kblock(a, b, mat); });
                        if (DO_M) scatter_row(mbuf, rec, ii, [&](int a, int b) { const double m = el.mval(a, b); return make_double4(m, 0.0, 0.0, m); });
                    }
                }
                __syncthreads();
            }
        }
        // ---- copy-out: runs of consecutive nodes leave with one bulk (TMA) store each; nothing waits for them ----
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic writes of the block -> visible to the async proxy
        __syncthreads();
        if (tid < 32) {
            for (int b0 = 0; b0 < nn; b0 += 32) {
                const int s = b0 + tid;
                int p0 = 0, pfx = 0, deg = 0;
                bool start = false;
                if (s < nn) {
                    const int2 m = nm[s];
                    p0 = m.x; pfx = m.y;
                    deg = ((s + 1 < nn) ? nm[s + 1].y : npairs) - pfx;
                    if (tid == 0) start = true;
                    else { const int2 mp = nm[s - 1]; start = (mp.x + (pfx - mp.y) != p0); }
                }
                const unsigned starts = __ballot_sync(0xffffffffu, start);
                const unsigned valid = __ballot_sync(0xffffffffu, s < nn);
                if (start) {
                    // run = [tid, next start or end of the valid lanes)
                    const unsigned later = starts & ~((2u << tid) - 1u);
                    const int end_lane = later ? (__ffs(later) - 1) : (32 - __clz(valid));
                    const int s_end = b0 + end_lane;   // first node after the run (within this group of 32)
                    const int pfx_end = (s_end < nn) ? nm[s_end].y : npairs;
                    const uint32_t bytes = (uint32_t)(pfx_end - pfx) * 32;
                    if (bytes) {
                        if (DO_K)
                            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(K_vals + 4ll * p0), "r"(t_smem_u32(kbuf + 2 * pfx)), "r"(bytes) : "memory");
                        if (DO_M)
                            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(M_vals + 4ll * p0), "r"(t_smem_u32(mbuf + 2 * pfx)), "r"(bytes) : "memory");
                    }
                }
                (void)deg;
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        // the barrier at the top of the next iteration orders everything else
    }
    if (tid < 32) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // all stores complete before the CTA exits
}

#undef PSFEM_STAGE_XY
#undef PSFEM_STAGE_NM
#undef PSFEM_STAGE_REC
#undef PSFEM_STAGE_CONN
#undef PSFEM_KBUF
#undef PSFEM_MBUF

template <int NPE, bool DO_K, bool DO_M> size_t tile_smem_bytes(const TilePlanDev &tp, int threads) {
    const size_t cap_e = tp.max_elems, cap_p = (tp.max_pairs + 15) & ~15;
    const size_t nbuf = (DO_K ? 1 : 0) + (DO_M ? 1 : 0);
    size_t b = 2 * nbuf * cap_p * 32;
    b += 2 * ((size_t)threads * NPE * 16 + (size_t)tp.R * 8 + cap_e * 32 + (size_t)threads * 16);
    return b + 32;
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
    *bytes = carve_tiles(nullptr, n_nodes, nel * npe > 0 ? nel * npe : 1).total;
    return PSFEM_OK;
}

extern "C" int psfem_tileplan_build(const double *d_nodes, const int32_t *d_conn, int64_t nel, int npe, int64_t n_nodes,
                                    int tile_nodes_R, const int32_t *d_node_ptr, const int32_t *d_col_idx,
                                    int32_t *d_tile_nodes /* n_nodes */, int32_t *d_tile_elem_ptr /* n_tiles+1 */,
                                    int32_t *d_tile_elems /* nel*npe (upper bound) */, int32_t *d_tile_hdr /* 8*n_tiles */,
                                    int32_t *d_tile_conn /* 4*nel*npe (upper bound) */, void *d_tile_rec /* 32 bytes * nel*npe (upper bound) */,
                                    int32_t *d_nmeta /* 2*n_tiles*R */,
                                    int64_t *h_info /* tile elements, max elems, max pairs, max colours, usable, mean elems */,
                                    void *d_ws, size_t ws_bytes, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    PSFEM_REQUIRE(npe == 3 || npe == 4, "tileplan_build: tiled assembly covers T3 and Q4");
    PSFEM_REQUIRE(tile_nodes_R >= 32 && tile_nodes_R <= 256 && tile_nodes_R % 2 == 0, "tileplan_build: tile size must be even and in [32, 256]");
    PSFEM_REQUIRE(h_info, "tileplan_build: null output");
    const int64_t ninc = nel * npe;
    PSFEM_REQUIRE(ninc > 0 && ninc < (int64_t)INT32_MAX, "tileplan_build: bad element count");
    const int R = tile_nodes_R;
    const int64_t n_tiles = ceil_div(n_nodes, R);
    TileWs w = carve_tiles(d_ws, n_nodes, ninc);
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
    // 3. per-tile unique element lists (ascending element id)
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
    // 4. pair prefix per tile-ordered node, row metadata
    PSFEM_CUDA_CHECK(cudaMemsetAsync(w.scal + 2, 0, 6 * sizeof(int), stream));
    tnode_counts_kernel<<<(unsigned)ceil_div(n_nodes + 1, T), T, 0, stream>>>(n_nodes, d_tile_nodes, d_node_ptr, w.deg);
    PSFEM_LAUNCH_CHECK();
    tb = w.cub_bytes;
    PSFEM_CUDA_CHECK(cub::DeviceScan::ExclusiveSum(w.cub_tmp, tb, w.deg, w.gp, (int)n_nodes + 1, stream));
    PSFEM_CUDA_CHECK(cudaMemsetAsync(d_nmeta, 0, (size_t)n_tiles * R * 8, stream));
    nmeta_kernel<<<(unsigned)ceil_div(n_nodes, T), T, 0, stream>>>(n_nodes, R, d_tile_nodes, d_node_ptr, w.gp, reinterpret_cast<int2 *>(d_nmeta));
    PSFEM_LAUNCH_CHECK();
    // 5. scatter rounds of every (tile element, row), packed into the tile connectivity; headers
    PSFEM_REQUIRE(n_nodes < (1ll << 28), "tileplan_build: node ids need 28 bits");
    if (npe == 3)
        tile_rounds_kernel<3><<<(unsigned)ceil_div(n_tiles, 64), 64, 0, stream>>>((int32_t)n_tiles, n_nodes, R, d_tile_nodes, d_tile_elem_ptr, d_tile_elems,
                                                                                d_conn, w.gp, reinterpret_cast<int4 *>(d_tile_conn), d_tile_hdr, w.scal + 5);
    else
        tile_rounds_kernel<4><<<(unsigned)ceil_div(n_tiles, 64), 64, 0, stream>>>((int32_t)n_tiles, n_nodes, R, d_tile_nodes, d_tile_elem_ptr, d_tile_elems,
                                                                                d_conn, w.gp, reinterpret_cast<int4 *>(d_tile_conn), d_tile_hdr, w.scal + 5);
    PSFEM_LAUNCH_CHECK();
    // 6. scatter record of every tile element
    if (npe == 3)
        tile_scatter_plan_kernel<3><<<(unsigned)ceil_div(h_unique, T), T, 0, stream>>>((int32_t)n_tiles, n_nodes, R, d_tile_nodes, d_tile_elem_ptr, d_tile_elems,
                                                                                      d_conn, d_node_ptr, d_col_idx, w.gp, static_cast<TileRec *>(d_tile_rec), w.scal + 5);
    else
        tile_scatter_plan_kernel<4><<<(unsigned)ceil_div(h_unique, T), T, 0, stream>>>((int32_t)n_tiles, n_nodes, R, d_tile_nodes, d_tile_elem_ptr, d_tile_elems,
                                                                                      d_conn, d_node_ptr, d_col_idx, w.gp, static_cast<TileRec *>(d_tile_rec), w.scal + 5);
    PSFEM_LAUNCH_CHECK();
    tile_max_kernel<<<(unsigned)ceil_div(n_tiles, T), T, 0, stream>>>((int32_t)n_tiles, d_tile_hdr, w.scal + 2);
    PSFEM_LAUNCH_CHECK();
    int h[4] = {0, 0, 0, 0};
    PSFEM_CUDA_CHECK(cudaMemcpyAsync(h, w.scal + 2, sizeof(h), cudaMemcpyDeviceToHost, stream));
    PSFEM_CUDA_CHECK(cudaStreamSynchronize(stream));
    h_info[0] = h_unique; h_info[1] = h[0]; h_info[2] = h[1]; h_info[3] = h[2];
    h_info[4] = (h[3] == 0 && h[0] <= 4095 && h[1] < 65535) ? 1 : 0;   // a node with > 14 elements / a huge tile: use psfem_assemble
    h_info[5] = ceil_div(h_unique, n_tiles);
    return PSFEM_OK;
}

extern "C" int psfem_assemble_tiled(const double *d_nodes, int64_t nel, int etype, int64_t n_nodes,
                                    double E, double nu, double rho, double t, int what, int tile_nodes_R,
                                    const int64_t *h_plan_info, const int32_t *d_tile_hdr, const int32_t *d_tile_conn,
                                    const void *d_tile_rec, const int32_t *d_tile_elems, const int32_t *d_nmeta,
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
    const Mat mat = make_mat(E, nu, rho, t, etype);
    unsigned long long *bad = static_cast<unsigned long long *>(d_ws);
    PSFEM_CUDA_CHECK(cudaMemsetAsync(bad, 0xff, sizeof(unsigned long long), stream));
    TilePlanDev tp;
    tp.hdr = reinterpret_cast<const int4 *>(d_tile_hdr); tp.tile_conn = reinterpret_cast<const int4 *>(d_tile_conn);
    tp.tile_rec = static_cast<const TileRec *>(d_tile_rec); tp.nmeta = reinterpret_cast<const int2 *>(d_nmeta); tp.tile_elems = d_tile_elems;
    tp.R = tile_nodes_R; tp.max_elems = (int)h_plan_info[1]; tp.max_pairs = (int)h_plan_info[2];
    const int64_t n_tiles = ceil_div(n_nodes, tp.R);
    const int mean = (int)h_plan_info[5];   // threads follow the TYPICAL tile; bigger tiles take a second pass
    int rc;
#define PSFEM_TILE_CASE(ET, NPE, TH, MB)                                                                              \
    (what == 3 ? launch_tiles<ET, NPE, true, true, TH, MB>(d_nodes, mat, n_tiles, tp, d_K_vals, d_M_vals, bad, stream) \
     : what == 1 ? launch_tiles<ET, NPE, true, false, TH, MB>(d_nodes, mat, n_tiles, tp, d_K_vals, d_M_vals, bad, stream) \
                 : launch_tiles<ET, NPE, false, true, TH, MB>(d_nodes, mat, n_tiles, tp, d_K_vals, d_M_vals, bad, stream))
    if (etype == PSFEM_Q4) {
        if (mean <= 96) rc = PSFEM_TILE_CASE(PSFEM_Q4, 4, 96, 4);
        else if (mean <= 160) {
            static const char *minb_env = getenv("PSFEM_TILE_MINB");   // tuning knob
            if (minb_env && minb_env[0] == '2') rc = PSFEM_TILE_CASE(PSFEM_Q4, 4, 160, 2);
            else rc = PSFEM_TILE_CASE(PSFEM_Q4, 4, 160, 3);
        }
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
