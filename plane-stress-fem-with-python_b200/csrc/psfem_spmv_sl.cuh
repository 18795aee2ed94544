// psfem_spmv_sl.cuh -- SpMV on SYMMETRIC LATTICE storage (included by psfem_solver.cu inside its anonymous namespace).
//
// The matrices of Polygones.mesh meshes (and their reductions by whole node columns, and the strips of the
// multi-GPU partition) couple node (i, j) -- id i*m + j -- with its 3x3 lattice neighbours only, and K, M and
// K + a M are symmetric.  So HALF of the 2x2 blocks carry all the information: per node the diagonal block (3
// values) and the four "forward" blocks to (i, j+1), (i+1, j-1), (i+1, j), (i+1, j+1) -- 19 doubles instead of the
// 36 values + 9 node ids of the node-block kernel.  The SpMV then streams 152 B of matrix per node instead of 324 B
// (152 + 16 x + 16 y = 184 B per node against 360 B: 0.51x the HBM traffic of spmv_nb_kernel).
//
// Storage is a structure of arrays written once per matrix by sl_build_kernel, sv[(s*19 + k)*mpad + j] with s the
// lattice column, k the component and j the node row: a warp reads 32 consecutive doubles per component, no index
// arrays, no shared-memory staging.
//
// Kernel: a warp owns 28 node rows (+1 halo lane below and above; 28 rows start on a 32-byte sector) of a segment of
// columns and MARCHES along i.  The
// transposed products (block (a,b) acting on row b) never touch memory: each row sum travels through the lanes that
// hold its blocks as a CHAIN of shuffles --
//     lane j-1 starts with F4(i-1,j-1)^T x, lane j adds F3(i-1,j)^T x, lane j+1 adds F2(i-1,j+1)^T x   (step i-1)
//     lane j-1 adds F1(i,j-1)^T x, lane j adds diagonal, F1, F2, F3, F4 times x                      (step i)
// -- four hops of two doubles per node.  That is exactly the ascending-column order of the CSR kernels with the same
// fused multiply-adds, so for a bitwise symmetric matrix (sl_build_kernel checks, tol = 0) the result is
// BIT-IDENTICAL to spmv_nb_kernel / spmv2_kernel / scipy.  Column segments start with a "pre-step" that recomputes
// the chain heads of the column before them (12 of 19 components of one column per ~340).  Matrices that are symmetric
// up to rounding only (T3: the reference's B^T D B) are accepted with tol > 0; the upper triangle then acts for both.
//
// Measured (4096 x 4096 Q4 plate, profiles/r1j_summary.md): 0.49 ms per SpMV + dot, DRAM traffic 3.097 GB of 3.088 GB
// algorithmic, 95-96 % of the copy-measured HBM peak; the node-block kernel needs 0.885 ms for 6.04 GB.
//
// x may be an extended vector (strip with halo columns): storage column 0 is then the column BEFORE the first owned
// one, its forward blocks are the transposes of the first owned column's backward blocks.
#pragma once

constexpr int SL_THREADS = 128;
constexpr int SL_WARPS = SL_THREADS / 32;
#ifndef PSFEM_SL_ROWS
#define PSFEM_SL_ROWS 28
#endif
constexpr int SL_ROWS = PSFEM_SL_ROWS;   // owned node rows per warp (lanes 1..SL_ROWS; lanes 0 and SL_ROWS+1 are halo rows)
static_assert(SL_ROWS >= 1 && SL_ROWS <= 30, "a warp holds the owned rows plus one halo row on either side");
constexpr int SL_NC = 19;     // components per node: D00 D01 D11 | F1 | F2 | F3 | F4 (row-major 2x2 blocks)
#ifndef PSFEM_SL_CTAS
#define PSFEM_SL_CTAS 3       // CTAs per SM (register cap 168)
#endif
#ifndef PSFEM_SL_PFD
#define PSFEM_SL_PFD 0        // L2 prefetch distance in columns (0: none); the next column is always loaded into registers
#endif
#ifndef PSFEM_SL_LD
#define PSFEM_SL_LD 0         // 0: ld.global.nc (strip-boundary sectors stay in L2 for the neighbouring warp), 1: ld.global.cs
#endif

struct SlGeom {
    int m, mpad;        // nodes per lattice column, padded to a multiple of 4
    int S, pre;         // storage columns; pre = 1 if storage column 0 precedes the first owned column
    int xcol0, xcols;   // x column of storage column 0, columns in x
    int n_strips, G;    // warp strips per column, column segments
};

__device__ __forceinline__ double sl_ld(const double *p, bool on) {
#if PSFEM_SL_LD
    return on ? __ldcs(p) : 0.0;
#else
    return on ? __ldg(p) : 0.0;
#endif
}
__device__ __forceinline__ void sl_pf(const double *p, bool on) {
    if (on) asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

template <bool DOT>
__global__ void __launch_bounds__(SL_THREADS, PSFEM_SL_CTAS)
spmv_sl_kernel(const double *__restrict__ sv, const SlGeom g, const double *__restrict__ x, double alpha, double gamma,
               const double *__restrict__ z, double *__restrict__ y, double *__restrict__ dot_out, double *__restrict__ partials,
               unsigned int *counter, const int *__restrict__ done, const PeerSync ps) {
    __shared__ double red[SL_THREADS / 32];
    pdl_enter();
    if (done && *done) return;
    constexpr unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int w = blockIdx.x * SL_WARPS + (threadIdx.x >> 5);
    double dot[1] = {0.0};
    if (w < g.n_strips * g.G) {
        const int sidx = w % g.n_strips, seg = w / g.n_strips;
        const int j = sidx * SL_ROWS - 1 + lane;
        const bool valid = j >= 0 && j < g.m && lane <= SL_ROWS + 1;
        const bool owned = valid && lane >= 1 && lane <= SL_ROWS;
        const bool lo = valid && lane <= SL_ROWS;   // lanes 0..SL_ROWS hold F1, F4 (blocks towards the row above)
        const bool hi = valid && lane >= 1;         // lanes 1..SL_ROWS+1 hold F2 (block towards the row below)
        const int owned_cols = g.S - g.pre;
        const int sa = g.pre + (int)((long long)seg * owned_cols / g.G);
        const int sb = g.pre + (int)((long long)(seg + 1) * owned_cols / g.G);
        const size_t mp = (size_t)g.mpad, cstride = (size_t)SL_NC * mp;
        const double2 *__restrict__ x2 = reinterpret_cast<const double2 *>(x);
        // x through L2 (ld.global.cg): the halo columns of a strip are written by the neighbouring GPUs
        auto xload = [&](int s) -> double2 {   // x of storage column s at this lane's row, zero outside the vector
            const int xc = s + g.xcol0;
            if (valid && xc < g.xcols) return __ldcg(x2 + (size_t)xc * g.m + j);
            return make_double2(0.0, 0.0);
        };
        if (ps.halo_in_spmv) {
            // strip of a multi-GPU partition: only the first column segment reads the left halo column (its pre-step)
            // and only the last one the right halo column; their warps wait for the neighbours' flags here, everybody
            // else starts at once (dist_pupdate_kernel then ends without waiting)
            const bool need_l = seg == 0, need_r = seg == g.G - 1;
            if (need_l || need_r) {
                if (lane == 0) peer_wait_halo(ps, need_l, need_r);
                __syncwarp();
            }
        }
        auto load_col = [&](int s, double (&F)[SL_NC]) {
            const double *p = sv + (size_t)s * cstride + j;
#pragma unroll
            for (int k = 0; k < 3; ++k) F[k] = sl_ld(p + k * mp, owned);
#pragma unroll
            for (int k = 3; k < 7; ++k) F[k] = sl_ld(p + k * mp, lo);
#pragma unroll
            for (int k = 7; k < 11; ++k) F[k] = sl_ld(p + k * mp, hi);
#pragma unroll
            for (int k = 11; k < 15; ++k) F[k] = sl_ld(p + k * mp, owned);
#pragma unroll
            for (int k = 15; k < 19; ++k) F[k] = sl_ld(p + k * mp, lo);
        };
        // chain heads of the NEXT column: on return lane l holds the partial sum of row l+1 of column s+1, made of
        // F4(s,l)^T x(s,l), F3(s,l+1)^T x(s,l+1), F2(s,l+2)^T x(s,l+2) in that order
        auto phaseP = [&](const double (&F)[SL_NC], const double2 xc, double &p0, double &p1) {
            double c0 = 0.0, c1 = 0.0;
            c0 += alpha * F[15] * xc.x; c0 += alpha * F[17] * xc.y;
            c1 += alpha * F[16] * xc.x; c1 += alpha * F[18] * xc.y;
            c0 = __shfl_up_sync(FULL, c0, 1); c1 = __shfl_up_sync(FULL, c1, 1);
            c0 += alpha * F[11] * xc.x; c0 += alpha * F[13] * xc.y;
            c1 += alpha * F[12] * xc.x; c1 += alpha * F[14] * xc.y;
            c0 = __shfl_up_sync(FULL, c0, 1); c1 = __shfl_up_sync(FULL, c1, 1);
            c0 += alpha * F[7] * xc.x; c0 += alpha * F[9] * xc.y;
            c1 += alpha * F[8] * xc.x; c1 += alpha * F[10] * xc.y;
            p0 = __shfl_down_sync(FULL, c0, 2); p1 = __shfl_down_sync(FULL, c1, 2);
        };
        double Fc[SL_NC], Fn[SL_NC];
        double p0 = 0.0, p1 = 0.0;
        if (sa > 0) {   // pre-step: chain heads from the column before the segment (another segment's or the halo column)
            const double *p = sv + (size_t)(sa - 1) * cstride + j;
#pragma unroll
            for (int k = 0; k < 7; ++k) Fc[k] = 0.0;
#pragma unroll
            for (int k = 7; k < 11; ++k) Fc[k] = sl_ld(p + k * mp, hi);
#pragma unroll
            for (int k = 11; k < 15; ++k) Fc[k] = sl_ld(p + k * mp, owned);
#pragma unroll
            for (int k = 15; k < 19; ++k) Fc[k] = sl_ld(p + k * mp, lo);
            phaseP(Fc, xload(sa - 1), p0, p1);
        }
        load_col(sa, Fc);
        double2 xc = xload(sa), xn = xload(sa + 1);
        for (int s = sa; s < sb; ++s) {
            const bool more = s + 1 < sb;
            double2 xnn = make_double2(0.0, 0.0);
            if (more) {
                load_col(s + 1, Fn);
                xnn = xload(s + 2);
#if PSFEM_SL_PFD
                if (s + PSFEM_SL_PFD < sb) {
                    const double *pp = sv + (size_t)(s + PSFEM_SL_PFD) * cstride + j;
#pragma unroll
                    for (int k = 0; k < SL_NC; ++k) sl_pf(pp + k * mp, owned);
                }
#endif
            }
            // row (s, l+1): add F1(s,l)^T x(s,l), hand the chain to its own lane
            double s0 = p0, s1 = p1;
            s0 += alpha * Fc[3] * xc.x; s0 += alpha * Fc[5] * xc.y;
            s1 += alpha * Fc[4] * xc.x; s1 += alpha * Fc[6] * xc.y;
            s0 = __shfl_up_sync(FULL, s0, 1); s1 = __shfl_up_sync(FULL, s1, 1);
            double2 xr, xnl, xnr;
            xr.x = __shfl_down_sync(FULL, xc.x, 1); xr.y = __shfl_down_sync(FULL, xc.y, 1);
            xnl.x = __shfl_up_sync(FULL, xn.x, 1); xnl.y = __shfl_up_sync(FULL, xn.y, 1);
            xnr.x = __shfl_down_sync(FULL, xn.x, 1); xnr.y = __shfl_down_sync(FULL, xn.y, 1);
            s0 += alpha * Fc[0] * xc.x; s0 += alpha * Fc[1] * xc.y;
            s1 += alpha * Fc[1] * xc.x; s1 += alpha * Fc[2] * xc.y;
            s0 += alpha * Fc[3] * xr.x; s0 += alpha * Fc[4] * xr.y;
            s1 += alpha * Fc[5] * xr.x; s1 += alpha * Fc[6] * xr.y;
            s0 += alpha * Fc[7] * xnl.x; s0 += alpha * Fc[8] * xnl.y;
            s1 += alpha * Fc[9] * xnl.x; s1 += alpha * Fc[10] * xnl.y;
            s0 += alpha * Fc[11] * xn.x; s0 += alpha * Fc[12] * xn.y;
            s1 += alpha * Fc[13] * xn.x; s1 += alpha * Fc[14] * xn.y;
            s0 += alpha * Fc[15] * xnr.x; s0 += alpha * Fc[16] * xnr.y;
            s1 += alpha * Fc[17] * xnr.x; s1 += alpha * Fc[18] * xnr.y;
            if (owned) {
                const size_t a = (size_t)(s - g.pre) * g.m + j;
                if (z) {
                    const double2 zz = reinterpret_cast<const double2 *>(z)[a];
                    s0 += gamma * zz.x; s1 += gamma * zz.y;
                }
                reinterpret_cast<double2 *>(y)[a] = make_double2(s0, s1);
                if (DOT) { dot[0] += xc.x * s0; dot[0] += xc.y * s1; }
            }
            if (more) {
                phaseP(Fc, xc, p0, p1);
#pragma unroll
                for (int k = 0; k < SL_NC; ++k) Fc[k] = Fn[k];
                xc = xn; xn = xnn;
            }
        }
    }
    if (DOT) {
        double total[1];
        if (grid_sum<SL_THREADS, 1>(dot, partials, counter, red, total)) {
            *dot_out = total[0];
            if (ps.world) {   // my p.q partial -> every rank's board; the global sum (rank order) stays in local memory
                peer_publish<1>(ps, 0, total);
                double pq[1];
                peer_wait_sum<1>(ps, 0, ps.seq, pq);
                ps.scal[4] = pq[0];
            }
        }
    }
}

// One thread per row node: forward blocks into the lattice storage, backward blocks checked against their transposes
// (or, for the column before the first owned one, stored as that column's forward blocks).  bad[0] != 0: the operand
// is not a symmetric 3x3-lattice matrix with this m.
__global__ void sl_build_kernel(int64_t n_row_nodes, int64_t roff, int64_t x_nodes, int m, int mpad, int pre, int c0x,
                                const int32_t *__restrict__ node_ptr, const int32_t *__restrict__ pcol,
                                const double *__restrict__ vals, double tol, double *__restrict__ sv, int *__restrict__ bad) {
    const int64_t a = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (a >= n_row_nodes) return;
    const int64_t e = a + roff;
    const int ci = (int)(e / m), j = (int)(e % m), s = ci - (c0x - pre);
    const int pair_base = node_ptr[0];
    const int q0 = node_ptr[a], deg = node_ptr[a + 1] - q0;
    const double *V = vals + 4ll * q0;
    const int32_t *pc = pcol + (q0 - pair_base);
    bool ok = deg >= 1 && deg <= 9;
    double scale = 0.0;
    for (int k = 0; k < deg && ok; ++k)
        if (pc[k] == (int32_t)e) scale = fmax(fabs(V[2 * k]), fabs(V[2 * deg + 2 * k + 1]));
    auto same = [&](double u, double v) { return tol == 0.0 ? u == v : fabs(u - v) <= tol * scale; };
    bool have_diag = false;
    const size_t mp = (size_t)mpad;
    for (int k = 0; k < deg && ok; ++k) {
        const int64_t b = pc[k];
        if (b < 0 || b >= x_nodes) { ok = false; break; }
        const int bi = (int)(b / m), bj = (int)(b % m), di = bi - ci, dj = bj - j;
        if (di < -1 || di > 1 || dj < -1 || dj > 1) { ok = false; break; }
        const double v00 = V[2 * k], v01 = V[2 * k + 1], v10 = V[2 * deg + 2 * k], v11 = V[2 * deg + 2 * k + 1];
        if (di == 0 && dj == 0) {
            have_diag = true;
            ok = ok && same(v01, v10);
            double *o = sv + ((size_t)s * SL_NC) * mp + j;
            o[0] = v00; o[mp] = v01; o[2 * mp] = v11;
            continue;
        }
        const bool forward = di == 1 || (di == 0 && dj == 1);
        if (forward) {
            const int comp = di == 0 ? 3 : (dj == -1 ? 7 : (dj == 0 ? 11 : 15));
            double *o = sv + ((size_t)s * SL_NC + comp) * mp + j;
            o[0] = v00; o[mp] = v01; o[2 * mp] = v10; o[3 * mp] = v11;
        }
        const int64_t pb = b - roff;
        if (pb >= 0 && pb < n_row_nodes) {   // the partner row is ours: block (b, a) must be the transpose
            const int qb = node_ptr[pb], degb = node_ptr[pb + 1] - qb;
            const double *W = vals + 4ll * qb;
            const int32_t *pcb = pcol + (qb - pair_base);
            bool found = false;
            for (int kk = 0; kk < degb; ++kk)
                if (pcb[kk] == (int32_t)e) {
                    found = true;
                    ok = ok && same(W[2 * kk], v00) && same(W[2 * kk + 1], v10) && same(W[2 * degb + 2 * kk], v01) &&
                         same(W[2 * degb + 2 * kk + 1], v11);
                    break;
                }
            ok = ok && found;
        } else if (!forward) {
            // partner in the column before the first owned one: its forward block towards this node is the transpose
            if (di == -1 && s >= 1) {
                const int comp = dj == 1 ? 7 : (dj == 0 ? 11 : 15);   // direction (1, -dj) seen from the partner
                double *o = sv + ((size_t)(s - 1) * SL_NC + comp) * mp + bj;
                o[0] = v00; o[mp] = v10; o[2 * mp] = v01; o[3 * mp] = v11;
            } else {
                ok = false;
            }
        }
    }
    if (!ok || !have_diag) atomicExch(bad, 1);
}
