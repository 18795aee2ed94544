// psfem_assemble.cu -- numeric phase: element matrices + deterministic segmented reduction.
//
//   element kernels : Ke / Me per element (Polygones.py:566-708, :840-883, :1218-1342), written in
//                     2x2-block layout  kblk[e][i][j] = {K(2i,2j), K(2i,2j+1), K(2i+1,2j), K(2i+1,2j+1)}
//                     (one 32-byte sector per block) and mblk[e][i][j] = M(2i,2j) (= M(2i+1,2j+1);
//                     the u-v couplings of Me are structurally zero, Polygones.py:1285-1288).
//   seg_reduce      : one thread per unique node pair; sums its contributions in ascending element
//                     id (the plan of psfem_pattern.cu) and writes the four CSR values exactly once.
//                     No atomics: the result is bit-reproducible run to run and independent of the grid.
#include "psfem_common.cuh"

using namespace psfem;

namespace {

// ---- quadrature / shape tables (Polygones.py:441-564), filled on the host -------------------
struct ShapeTab {
    int ngp;
    double w[9];
    double N[9][9];
    double dxi[9][9];
    double deta[9][9];
};

__constant__ ShapeTab c_tab[4];  // indexed by element kind (T3 unused)

__host__ __device__ void shape_eval_host(int etype, double xi, double eta, double *N, double *dxi, double *det) {
    if (etype == PSFEM_T3) {  // Polygones.py:460-464 (used by compute_B_matrix_at_point only; Ke/Me of T3 are closed forms)
        N[0] = 1 - xi - eta; N[1] = xi; N[2] = eta;
        dxi[0] = -1; dxi[1] = 1; dxi[2] = 0;
        det[0] = -1; det[1] = 0; det[2] = 1;
    } else if (etype == PSFEM_Q4) {  // Polygones.py:479-484
        const double sx[4] = {-1, 1, 1, -1}, sy[4] = {-1, -1, 1, 1};
        for (int i = 0; i < 4; ++i) {
            N[i] = (1 + sx[i] * xi) * (1 + sy[i] * eta) / 4;
            dxi[i] = sx[i] * (1 + sy[i] * eta) / 4;
            det[i] = sy[i] * (1 + sx[i] * xi) / 4;
        }
    } else if (etype == PSFEM_T6) {  // Polygones.py:467-474
        double s = 1 - xi - eta;
        N[0] = s * (2 * s - 1); N[1] = xi * (2 * xi - 1); N[2] = eta * (2 * eta - 1);
        N[3] = 4 * xi * s; N[4] = 4 * xi * eta; N[5] = 4 * eta * s;
        dxi[0] = 4 * xi + 4 * eta - 3; dxi[1] = 4 * xi - 1; dxi[2] = 0; dxi[3] = 4 - 8 * xi - 4 * eta; dxi[4] = 4 * eta; dxi[5] = -4 * eta;
        det[0] = 4 * xi + 4 * eta - 3; det[1] = 0; det[2] = 4 * eta - 1; det[3] = -4 * xi; det[4] = 4 * xi; det[5] = 4 - 4 * xi - 8 * eta;
    } else {  // Q9, Polygones.py:487-497: products of 1-D quadratic Lagrange polynomials at -1,0,+1
        double fx[3] = {xi * (xi - 1) / 2, 1 - xi * xi, xi * (xi + 1) / 2};
        double fy[3] = {eta * (eta - 1) / 2, 1 - eta * eta, eta * (eta + 1) / 2};
        double gx[3] = {xi - 0.5, -2 * xi, xi + 0.5};
        double gy[3] = {eta - 0.5, -2 * eta, eta + 0.5};
        const int ab[9][2] = {{0, 0}, {2, 0}, {2, 2}, {0, 2}, {1, 0}, {2, 1}, {1, 2}, {0, 1}, {1, 1}};
        for (int i = 0; i < 9; ++i) {
            N[i] = fx[ab[i][0]] * fy[ab[i][1]];
            dxi[i] = gx[ab[i][0]] * fy[ab[i][1]];
            det[i] = fx[ab[i][0]] * gy[ab[i][1]];
        }
    }
}

void build_tab(int etype, ShapeTab &t) {
    double pts[9][2];
    if (etype == PSFEM_T6) {  // Polygones.py:536-546 -- weights sum to 1, not 1/2 (reproduced)
        t.ngp = 3;
        const double p[3][2] = {{1.0 / 6, 1.0 / 6}, {2.0 / 3, 1.0 / 6}, {1.0 / 6, 2.0 / 3}};
        for (int g = 0; g < 3; ++g) { pts[g][0] = p[g][0]; pts[g][1] = p[g][1]; t.w[g] = 1.0 / 3; }
    } else {  // Polygones.py:547-562: 3x3 rule, eta outer / xi inner
        t.ngp = 9;
        const double a = sqrt(0.6);
        const double c[3] = {-a, 0.0, a};
        const double w1 = 5.0 / 9, w2 = 8.0 / 9, w3 = 5.0 / 9;
        const double ww[3] = {w1, w2, w3};
        for (int r = 0; r < 3; ++r)
            for (int q = 0; q < 3; ++q) {
                pts[3 * r + q][0] = c[q];
                pts[3 * r + q][1] = c[r];
                t.w[3 * r + q] = ww[r] * ww[q];
            }
    }
    for (int g = 0; g < t.ngp; ++g) shape_eval_host(etype, pts[g][0], pts[g][1], t.N[g], t.dxi[g], t.deta[g]);
}

int upload_tabs(cudaStream_t stream) {
    static ShapeTab h[4];
    static bool built = false;
    if (!built) {
        for (int e = PSFEM_Q4; e <= PSFEM_Q9; ++e) build_tab(e, h[e]);
        built = true;
    }
    PSFEM_CUDA_CHECK(cudaMemcpyToSymbolAsync(c_tab, h, sizeof(h), 0, cudaMemcpyHostToDevice, stream));
    return PSFEM_OK;
}

struct Material {
    double d00, d01, d22;  // D = E/(1-nu^2) [[1,nu,0],[nu,1,0],[0,0,(1-nu)/2]]   Polygones.py:582-584
    double t, rho;
};

// ---- T3 closed form: Polygones.py:840-883 (Ke), :1218-1235 (Me); SIGNED area ---------------
template <bool DO_K, bool DO_M>
__global__ void t3_kernel(const double2 *__restrict__ nodes, const int32_t *__restrict__ conn, int64_t nel, Material mat,
                          double4 *__restrict__ kblk, double *__restrict__ mblk) {
    int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= nel) return;
    const int32_t *c = conn + e * 3;
    double2 p[3] = {nodes[c[0]], nodes[c[1]], nodes[c[2]]};
    // delta = signed 2*area (:843-845)
    double delta = __dadd_rn(__dadd_rn(__dmul_rn(p[0].x, p[1].y - p[2].y), __dmul_rn(p[1].x, p[2].y - p[0].y)),
                             __dmul_rn(p[2].x, p[0].y - p[1].y));   // no FMA contraction: the reference rounds every product
    double vol = 0.5 * delta * mat.t;  // :877-879
    if (DO_K) {
        double b[3], g[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) {  // :848-850
            b[i] = (p[(i + 1) % 3].y - p[(i + 2) % 3].y) / delta;
            g[i] = (p[(i + 2) % 3].x - p[(i + 1) % 3].x) / delta;
        }
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                double4 v;
                // np.dot(B.T, np.dot(D, B)) * Volume in the reference's operation order (see ElemT3 in psfem_elem.cuh)
                v.x = __dmul_rn(fma(g[i], __dmul_rn(mat.d22, g[j]), __dmul_rn(b[i], __dmul_rn(mat.d00, b[j]))), vol);
                v.y = __dmul_rn(fma(g[i], __dmul_rn(mat.d22, b[j]), __dmul_rn(b[i], __dmul_rn(mat.d01, g[j]))), vol);
                v.z = __dmul_rn(fma(b[i], __dmul_rn(mat.d22, g[j]), __dmul_rn(g[i], __dmul_rn(mat.d01, b[j]))), vol);
                v.w = __dmul_rn(fma(b[i], __dmul_rn(mat.d22, b[j]), __dmul_rn(g[i], __dmul_rn(mat.d00, g[j]))), vol);
                kblk[e * 9 + i * 3 + j] = v;
            }
    }
    if (DO_M) {
        double base = mat.rho * vol / 12;  // Me * rho * Volume / 12 (:1234)
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) mblk[e * 9 + i * 3 + j] = (i == j ? 2.0 : 1.0) * base;
    }
}

// ---- Gauss-integrated elements (Q4, T6, Q9): one thread per (element, row node i) ------------
// Per Gauss point: J (Polygones.py:675-682), detJ guard (:685-687), dN/dx, dN/dy (:697-698),
// Ke += w B^T D B |detJ| t (:595), Me += w rho t N^T N |detJ| (:1291).
template <int ETYPE, int NPE, bool DO_K, bool DO_M>
__global__ void __launch_bounds__(128)
gauss_kernel(const double2 *__restrict__ nodes, const int32_t *__restrict__ conn, int64_t nel, Material mat,
             double4 *__restrict__ kblk, double *__restrict__ mblk, unsigned long long *__restrict__ bad) {
    int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (tid >= nel * NPE) return;
    const int64_t e = tid / NPE;
    const int i = (int)(tid - e * NPE);
    const ShapeTab &tab = c_tab[ETYPE];
    double x[NPE], y[NPE];
#pragma unroll
    for (int j = 0; j < NPE; ++j) {
        double2 p = nodes[conn[e * NPE + j]];
        x[j] = p.x;
        y[j] = p.y;
    }
    double4 acc[NPE];
    double macc[NPE];
#pragma unroll
    for (int j = 0; j < NPE; ++j) { acc[j] = make_double4(0, 0, 0, 0); macc[j] = 0; }
    const int ngp = (ETYPE == PSFEM_T6) ? 3 : 9;
    for (int g = 0; g < ngp; ++g) {
        double j00 = 0, j01 = 0, j10 = 0, j11 = 0;
#pragma unroll
        for (int j = 0; j < NPE; ++j) {
            j00 += tab.dxi[g][j] * x[j];
            j01 += tab.dxi[g][j] * y[j];
            j10 += tab.deta[g][j] * x[j];
            j11 += tab.deta[g][j] * y[j];
        }
        const double detJ = j00 * j11 - j01 * j10;
        const double ad = fabs(detJ);
        if (ad < 1e-10) {  // Polygones.py:686-687
            atomicMin(bad, (unsigned long long)e);
            return;
        }
        if (DO_K) {
            const double inv = 1.0 / detJ;
            const double i00 = j11 * inv, i01 = -j01 * inv, i10 = -j10 * inv, i11 = j00 * inv;
            const double c = tab.w[g] * ad * mat.t;
            double dxi_i = 0, dyi_i = 0;
            double dx[NPE], dy[NPE];
#pragma unroll
            for (int j = 0; j < NPE; ++j) {
                dx[j] = i00 * tab.dxi[g][j] + i01 * tab.deta[g][j];
                dy[j] = i10 * tab.dxi[g][j] + i11 * tab.deta[g][j];
                if (j == i) { dxi_i = dx[j]; dyi_i = dy[j]; }
            }
            const double a0 = c * mat.d00 * dxi_i, a1 = c * mat.d22 * dyi_i;  // row 2i
            const double a2 = c * mat.d01 * dxi_i;
            const double b0 = c * mat.d01 * dyi_i, b1 = c * mat.d22 * dxi_i;  // row 2i+1
            const double b2 = c * mat.d00 * dyi_i;
#pragma unroll
            for (int j = 0; j < NPE; ++j) {
                acc[j].x += a0 * dx[j] + a1 * dy[j];
                acc[j].y += a2 * dy[j] + a1 * dx[j];
                acc[j].z += b0 * dx[j] + b1 * dy[j];
                acc[j].w += b2 * dy[j] + b1 * dx[j];
            }
        }
        if (DO_M) {
            const double cm = tab.w[g] * mat.rho * mat.t * ad * tab.N[g][i];
#pragma unroll
            for (int j = 0; j < NPE; ++j) macc[j] += cm * tab.N[g][j];
        }
    }
    const int64_t base = tid * NPE;
    if (DO_K) {
#pragma unroll
        for (int j = 0; j < NPE; ++j) kblk[base + j] = acc[j];
    }
    if (DO_M) {
#pragma unroll
        for (int j = 0; j < NPE; ++j) mblk[base + j] = macc[j];
    }
}

// ---- compute_B_matrix_at_point (Polygones.py:657-708): B (3 x 2npe), J (2 x 2), detJ of ONE element at (xi, eta) ----
// out: B row-major [3][2npe], then J00 J01 J10 J11, then detJ, then a flag (1.0: |detJ| < 1e-10, B not written)
__global__ void b_matrix_kernel(const double2 *__restrict__ nodes, const int32_t *__restrict__ conn1, int etype, int npe,
                                double xi, double eta, double *__restrict__ out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double N[9], dxi[9], det[9];
    shape_eval_host(etype, xi, eta, N, dxi, det);
    double j00 = 0, j01 = 0, j10 = 0, j11 = 0;
    for (int i = 0; i < npe; ++i) {   // :675-682
        const double2 p = nodes[conn1[i]];
        j00 += dxi[i] * p.x; j01 += dxi[i] * p.y;
        j10 += det[i] * p.x; j11 += det[i] * p.y;
    }
    const double detJ = j00 * j11 - j01 * j10;
    double *B = out, *J = out + 6 * npe;
    J[0] = j00; J[1] = j01; J[2] = j10; J[3] = j11; J[4] = detJ;
    if (fabs(detJ) < 1e-10) { J[5] = 1.0; return; }   // :686-687
    J[5] = 0.0;
    const double inv = 1.0 / detJ;
    const double i00 = j11 * inv, i01 = -j01 * inv, i10 = -j10 * inv, i11 = j00 * inv;
    for (int i = 0; i < npe; ++i) {   // :692-706
        const double dx = i00 * dxi[i] + i01 * det[i], dy = i10 * dxi[i] + i11 * det[i];
        B[2 * i] = dx; B[2 * i + 1] = 0.0;
        B[2 * npe + 2 * i] = 0.0; B[2 * npe + 2 * i + 1] = dy;
        B[4 * npe + 2 * i] = dy; B[4 * npe + 2 * i + 1] = dx;
    }
}

// ---- deterministic segmented reduction into the CSR value arrays ------------------------------
template <bool DO_K, bool DO_M>
__global__ void seg_reduce_kernel(int64_t n_pairs, const int32_t *__restrict__ node_ptr, const int32_t *__restrict__ pair_row,
                                  const int32_t *__restrict__ pair_ptr, const int32_t *__restrict__ perm,
                                  const double4 *__restrict__ kblk, const double *__restrict__ mblk,
                                  double *__restrict__ K_vals, double *__restrict__ M_vals) {
    int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (s >= n_pairs) return;
    const int32_t a = pair_row[s];
    const int32_t p0 = node_ptr[a], deg = node_ptr[a + 1] - p0, pos = (int32_t)s - p0;
    const int32_t k0 = pair_ptr[s], k1 = pair_ptr[s + 1];
    double4 acc = make_double4(0, 0, 0, 0);
    double m = 0;
    for (int32_t k = k0; k < k1; ++k) {  // ascending element id (stable sort) = order of Polygones.py:717
        const int32_t code = perm[k];
        if (DO_K) {
            const double4 v = kblk[code];
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        if (DO_M) m += mblk[code];
    }
    const int64_t r0 = 4ll * p0 + 2 * pos, r1 = r0 + 2 * deg;
    if (DO_K) {
        *reinterpret_cast<double2 *>(K_vals + r0) = make_double2(acc.x, acc.y);
        *reinterpret_cast<double2 *>(K_vals + r1) = make_double2(acc.z, acc.w);
    }
    if (DO_M) {
        *reinterpret_cast<double2 *>(M_vals + r0) = make_double2(m, 0.0);
        *reinterpret_cast<double2 *>(M_vals + r1) = make_double2(0.0, m);
    }
}

// expand block layout to the reference's dense (2npe x 2npe) element matrices
__global__ void expand_blocks_kernel(int64_t nblk, int npe, const double4 *__restrict__ kblk, const double *__restrict__ mblk,
                                     double *__restrict__ Ke, double *__restrict__ Me) {
    int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (b >= nblk) return;
    const int npe2 = npe * npe, nd = 2 * npe;
    int64_t e = b / npe2;
    int ij = (int)(b - e * npe2), i = ij / npe, j = ij - i * npe;
    int64_t o = e * nd * nd + (2 * i) * nd + 2 * j;
    if (Ke) {
        double4 v = kblk[b];
        Ke[o] = v.x; Ke[o + 1] = v.y; Ke[o + nd] = v.z; Ke[o + nd + 1] = v.w;
    }
    if (Me) {
        double m = mblk[b];
        Me[o] = m; Me[o + 1] = 0; Me[o + nd] = 0; Me[o + nd + 1] = m;
    }
}

template <bool DO_K, bool DO_M>
int launch_elements(const double *d_nodes, const int32_t *d_conn, int64_t nel, int etype, Material mat,
                    double4 *kblk, double *mblk, unsigned long long *bad, cudaStream_t stream) {
    const double2 *nodes = reinterpret_cast<const double2 *>(d_nodes);
    if (etype == PSFEM_T3) {
        const int T = 128;
        t3_kernel<DO_K, DO_M><<<(unsigned)ceil_div(nel, T), T, 0, stream>>>(nodes, d_conn, nel, mat, kblk, mblk);
    } else {
        const int T = 128;
        if (etype == PSFEM_Q4)
            gauss_kernel<PSFEM_Q4, 4, DO_K, DO_M><<<(unsigned)ceil_div(nel * 4, T), T, 0, stream>>>(nodes, d_conn, nel, mat, kblk, mblk, bad);
        else if (etype == PSFEM_T6)
            gauss_kernel<PSFEM_T6, 6, DO_K, DO_M><<<(unsigned)ceil_div(nel * 6, T), T, 0, stream>>>(nodes, d_conn, nel, mat, kblk, mblk, bad);
        else
            gauss_kernel<PSFEM_Q9, 9, DO_K, DO_M><<<(unsigned)ceil_div(nel * 9, T), T, 0, stream>>>(nodes, d_conn, nel, mat, kblk, mblk, bad);
    }
    PSFEM_LAUNCH_CHECK();
    return PSFEM_OK;
}

int run_elements(const double *d_nodes, const int32_t *d_conn, int64_t nel, int etype, Material mat, int what,
                 double4 *kblk, double *mblk, unsigned long long *bad, cudaStream_t stream) {
    if (etype != PSFEM_T3) {
        int rc = upload_tabs(stream);
        if (rc) return rc;
    }
    PSFEM_CUDA_CHECK(cudaMemsetAsync(bad, 0xff, sizeof(unsigned long long), stream));
    if (what == 3) return launch_elements<true, true>(d_nodes, d_conn, nel, etype, mat, kblk, mblk, bad, stream);
    if (what == 1) return launch_elements<true, false>(d_nodes, d_conn, nel, etype, mat, kblk, mblk, bad, stream);
    return launch_elements<false, true>(d_nodes, d_conn, nel, etype, mat, kblk, mblk, bad, stream);
}

struct AsmWs {
    double4 *kblk;
    double *mblk;
    unsigned long long *bad;
    size_t total;
};

AsmWs carve_asm(void *ws, int64_t nel, int npe, int what) {
    AsmWs w;
    Carver c(ws);
    int64_t nb = nel * npe * npe;
    w.kblk = c.take<double4>((what & 1) ? nb : 0);
    w.mblk = c.take<double>((what & 2) ? nb : 0);
    w.bad = c.take<unsigned long long>(1);
    w.total = c.used();
    return w;
}

Material make_material(double E, double nu, double rho, double t, int etype) {
    // the T3 closed form and the Gauss path round D differently in the reference (Polygones.py:881 vs :582-584)
    Material m;
    const double den = 1 - nu * nu;
    m.d00 = E / den;
    if (etype == PSFEM_T3) {
        m.d01 = (nu * E) / den;
        m.d22 = (((1 - nu) / 2) * E) / den;
    } else {
        m.d01 = nu * m.d00;
        m.d22 = ((1 - nu) / 2) * m.d00;
    }
    m.t = t;
    m.rho = rho;
    return m;
}

int finish_bad(unsigned long long *d_bad, int64_t *h_bad_element, cudaStream_t stream) {
    unsigned long long hb = 0;
    PSFEM_CUDA_CHECK(cudaMemcpyAsync(&hb, d_bad, sizeof(hb), cudaMemcpyDeviceToHost, stream));
    PSFEM_CUDA_CHECK(cudaStreamSynchronize(stream));
    if (hb != ~0ull) {
        if (h_bad_element) *h_bad_element = (int64_t)hb;
        set_error("Jacobian determinant near zero for element %llu", hb);
        return PSFEM_EJACOBIAN;
    }
    if (h_bad_element) *h_bad_element = -1;
    return PSFEM_OK;
}

}  // namespace

extern "C" int psfem_assemble_workspace(int64_t nel, int etype, int what, size_t *bytes) {
    int npe = psfem_nodes_per_element(etype);
    PSFEM_REQUIRE(bytes && npe > 0 && what >= 1 && what <= 3, "assemble_workspace: bad arguments");
    *bytes = carve_asm(nullptr, nel, npe, what).total;
    return PSFEM_OK;
}

extern "C" int psfem_assemble(const double *d_nodes, const int32_t *d_conn, int64_t nel, int etype, int64_t n_nodes,
                              double E, double nu, double rho, double t, int what, int64_t n_pairs,
                              const int32_t *d_node_ptr, const int32_t *d_pair_row, const int32_t *d_pair_ptr,
                              const int32_t *d_perm, double *d_K_vals, double *d_M_vals, void *d_ws, size_t ws_bytes,
                              int64_t *h_bad_element, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    int npe = psfem_nodes_per_element(etype);
    PSFEM_REQUIRE(npe > 0, "assemble: unknown element kind %d", etype);
    PSFEM_REQUIRE(what >= 1 && what <= 3, "assemble: what must be 1 (K), 2 (M) or 3 (both)");
    PSFEM_REQUIRE(!(what & 1) || d_K_vals, "assemble: K requested but d_K_vals is null");
    PSFEM_REQUIRE(!(what & 2) || d_M_vals, "assemble: M requested but d_M_vals is null");
    if (h_bad_element) *h_bad_element = -1;
    if (nel == 0 || n_pairs == 0) return PSFEM_OK;
    AsmWs w = carve_asm(d_ws, nel, npe, what);
    PSFEM_REQUIRE(ws_bytes >= w.total, "assemble: workspace too small (%zu < %zu)", ws_bytes, w.total);
    int rc = run_elements(d_nodes, d_conn, nel, etype, make_material(E, nu, rho, t, etype), what, w.kblk, w.mblk, w.bad, stream);
    if (rc) return rc;
    const int T = 256;
    unsigned grid = (unsigned)ceil_div(n_pairs, T);
    if (what == 3)
        seg_reduce_kernel<true, true><<<grid, T, 0, stream>>>(n_pairs, d_node_ptr, d_pair_row, d_pair_ptr, d_perm, w.kblk, w.mblk, d_K_vals, d_M_vals);
    else if (what == 1)
        seg_reduce_kernel<true, false><<<grid, T, 0, stream>>>(n_pairs, d_node_ptr, d_pair_row, d_pair_ptr, d_perm, w.kblk, w.mblk, d_K_vals, d_M_vals);
    else
        seg_reduce_kernel<false, true><<<grid, T, 0, stream>>>(n_pairs, d_node_ptr, d_pair_row, d_pair_ptr, d_perm, w.kblk, w.mblk, d_K_vals, d_M_vals);
    PSFEM_LAUNCH_CHECK();
    return finish_bad(w.bad, h_bad_element, stream);
}

extern "C" int psfem_element_matrices(const double *d_nodes, const int32_t *d_conn, int64_t nel, int etype, double E,
                                      double nu, double rho, double t, int what, double *d_Ke, double *d_Me,
                                      void *d_ws, size_t ws_bytes, int64_t *h_bad_element, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    int npe = psfem_nodes_per_element(etype);
    PSFEM_REQUIRE(npe > 0, "element_matrices: unknown element kind %d", etype);
    PSFEM_REQUIRE(what >= 1 && what <= 3, "element_matrices: what must be 1, 2 or 3");
    if (h_bad_element) *h_bad_element = -1;
    if (nel == 0) return PSFEM_OK;
    AsmWs w = carve_asm(d_ws, nel, npe, what);
    PSFEM_REQUIRE(ws_bytes >= w.total, "element_matrices: workspace too small");
    int rc = run_elements(d_nodes, d_conn, nel, etype, make_material(E, nu, rho, t, etype), what, w.kblk, w.mblk, w.bad, stream);
    if (rc) return rc;
    int64_t nb = nel * npe * npe;
    expand_blocks_kernel<<<(unsigned)ceil_div(nb, 256), 256, 0, stream>>>(nb, npe, (what & 1) ? w.kblk : nullptr,
                                                                         (what & 2) ? w.mblk : nullptr,
                                                                         (what & 1) ? d_Ke : nullptr, (what & 2) ? d_Me : nullptr);
    PSFEM_LAUNCH_CHECK();
    return finish_bad(w.bad, h_bad_element, stream);
}

extern "C" int psfem_b_matrix_at_point(const double *d_nodes, const int32_t *d_conn1, int etype, double xi, double eta,
                                       double *d_out, double *h_out, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const int npe = psfem_nodes_per_element(etype);
    PSFEM_REQUIRE(npe > 0 && d_nodes && d_conn1 && d_out && h_out, "b_matrix_at_point: bad arguments");
    b_matrix_kernel<<<1, 32, 0, stream>>>(reinterpret_cast<const double2 *>(d_nodes), d_conn1, etype, npe, xi, eta, d_out);
    PSFEM_LAUNCH_CHECK();
    PSFEM_CUDA_CHECK(cudaMemcpyAsync(h_out, d_out, (size_t)(6 * npe + 6) * sizeof(double), cudaMemcpyDeviceToHost, stream));
    PSFEM_CUDA_CHECK(cudaStreamSynchronize(stream));
    if (h_out[6 * npe + 5] != 0.0) {
        set_error("Jacobian determinant near zero");
        return PSFEM_EJACOBIAN;
    }
    return PSFEM_OK;
}
