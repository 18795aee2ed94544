// psfem_elem.cuh -- per-element arithmetic of the T3 / Q4 kernels (shared by the tile-fused and the
// lattice assembly): gradient products of Ke and shape products of Me, computed once per element in registers.
#pragma once
#include "psfem_common.cuh"

namespace psfem {

struct Mat {
    double d00, d01, d22, t, rho;
};

// Plane-stress D with the reference's own rounding.  The generic path scales the pattern matrix by the
// precomputed factor, [[1,nu,0],[nu,1,0],[0,0,(1-nu)/2]] * (E/(1-nu**2)) (Polygones.py:582-584); the T3 closed
// form multiplies by E first and divides afterwards, [[..]]*E/(1-nu**2) (:881) -- different last bits.
inline Mat make_mat(double E, double nu, double rho, double t, int etype) {
    Mat m;
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

// 1/d without the IEEE division slow path: fp32 reciprocal seed + two Newton steps (relative error ~1e-16;
// |d| >= 1e-10 is guaranteed by the Jacobian guard, far inside the fp32 exponent range)
__device__ __forceinline__ double fast_rcp(double d) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"((float)d));   // one MUFU.RCP, relative error 2^-23
    double x = (double)r;
    x = x * (2.0 - d * x);
    x = x * (2.0 - d * x);
    return x;
}

template <int NPE> __host__ __device__ constexpr int tri_index(int i, int j) { return i * NPE - (i * (i - 1)) / 2 + (j - i); }

// Q4: gradient products of the element, integrated with the 3x3 rule (Polygones.py:534-564, :582-595,
// :657-708; mass :1266-1291).  The 9-point sums are evaluated through their SEPARABLE structure instead of
// point by point: with a_i = dN_i/dxi = +-e(eta), b_i = dN_i/deta = +-f(xi), J = [[j00(eta), j01(eta)],
// [j10(xi), j11(xi)]] and g = w*t/|detJ|,
//   gxx_ij = sum g (j11 a_i - j01 b_i)(j11 a_j - j01 b_j),   gyy_ij = sum g (j00 b_i - j10 a_i)(j00 b_j - j10 a_j),
//   gxy_ij = sum g (j11 a_i - j01 b_i)(j00 b_j - j10 a_j)
// are +-combinations of 34 moments  sum_rc g_rc F(xi_c) G(eta_r) e_u(eta_r) e_v / f_u(xi_c) f_v,  each a
// 3-term sum of 3-term sums.  Same quadrature points and weights as the reference, identical up to rounding
// (1e-15 relative, tests: <= 1e-10), about 0.5 k instead of 0.8 k fp64 operations per element.
template <bool DO_K, bool DO_M> struct ElemQ4 {
    double gxx[10], gyy[10], gxy[16], mm[10];
    __device__ __forceinline__ bool compute(const double2 (&p)[4], const Mat &mat) {
        constexpr double A = 0.7745966692414834;  // sqrt(0.6)  (Polygones.py:549)
        // e_-(eta_r) = (1 - eta_r)/4 and e_+(eta_r) = (1 + eta_r)/4 at the three Gauss abscissae (same tables in xi)
        constexpr double E0[3] = {(1 + A) / 4, 0.25, (1 - A) / 4};
        constexpr double E1[3] = {(1 - A) / 4, 0.25, (1 + A) / 4};
        constexpr double EE[3][3] = {{E0[0] * E0[0], E0[1] * E0[1], E0[2] * E0[2]},   // e_- e_-
                                     {E0[0] * E1[0], E0[1] * E1[1], E0[2] * E1[2]},   // e_- e_+
                                     {E1[0] * E1[0], E1[1] * E1[1], E1[2] * E1[2]}};  // e_+ e_+
        constexpr double GW[3] = {5.0 / 9, 8.0 / 9, 5.0 / 9};
        // Jacobian rows: j00, j01 depend on eta only, j10, j11 on xi only (Polygones.py:675-682)
        const double dx10 = p[1].x - p[0].x, dx23 = p[2].x - p[3].x, dy10 = p[1].y - p[0].y, dy23 = p[2].y - p[3].y;
        const double dx30 = p[3].x - p[0].x, dx21 = p[2].x - p[1].x, dy30 = p[3].y - p[0].y, dy21 = p[2].y - p[1].y;
        double j00[3], j01[3], j10[3], j11[3];
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            j00[q] = E0[q] * dx10 + E1[q] * dx23; j01[q] = E0[q] * dy10 + E1[q] * dy23;   // index r (eta)
            j10[q] = E0[q] * dx30 + E1[q] * dx21; j11[q] = E0[q] * dy30 + E1[q] * dy21;   // index c (xi)
        }
        double ad[3][3];
        bool ok = true;
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                ad[r][c] = fabs(j00[r] * j11[c] - j01[r] * j10[c]);
                if (ad[r][c] < 1e-10) ok = false;  // Polygones.py:686-687
            }
        if (DO_K) {
            double g[3][3];   // w_r w_c t / |detJ|
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int c = 0; c < 3; ++c) g[r][c] = (GW[r] * GW[c]) * mat.t * fast_rcp(ad[r][c]);
            // moments of a xi-only field F: AA[u+v] = sum_r e_u e_v(eta_r) sum_c g_rc F_c
            auto aa = [&](const double (&F)[3], double (&out)[3]) {
                double S[3];
#pragma unroll
                for (int r = 0; r < 3; ++r) S[r] = g[r][0] * F[0] + g[r][1] * F[1] + g[r][2] * F[2];
#pragma unroll
                for (int k = 0; k < 3; ++k) out[k] = EE[k][0] * S[0] + EE[k][1] * S[1] + EE[k][2] * S[2];
            };
            // moments of an eta-only field G: BB[u+v] = sum_c f_u f_v(xi_c) sum_r g_rc G_r
            auto bb = [&](const double (&G)[3], double (&out)[3]) {
                double T[3];
#pragma unroll
                for (int c = 0; c < 3; ++c) T[c] = g[0][c] * G[0] + g[1][c] * G[1] + g[2][c] * G[2];
#pragma unroll
                for (int k = 0; k < 3; ++k) out[k] = EE[k][0] * T[0] + EE[k][1] * T[1] + EE[k][2] * T[2];
            };
            // C_P[v][r] = sum_c g_rc P_c f_v(xi_c)
            auto cc = [&](const double (&P)[3], double (&out)[2][3]) {
                double P0[3], P1[3];
#pragma unroll
                for (int c = 0; c < 3; ++c) { P0[c] = P[c] * E0[c]; P1[c] = P[c] * E1[c]; }
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    out[0][r] = g[r][0] * P0[0] + g[r][1] * P0[1] + g[r][2] * P0[2];
                    out[1][r] = g[r][0] * P1[0] + g[r][1] * P1[1] + g[r][2] * P1[2];
                }
            };
            // mixed moments of P(xi) Q(eta): AB[u][v] = sum_r e_u(eta_r) Q_r C_P[v][r]
            auto ab = [&](const double (&C)[2][3], const double (&Q0)[3], const double (&Q1)[3], double (&out)[2][2]) {
#pragma unroll
                for (int v = 0; v < 2; ++v) {
                    out[0][v] = Q0[0] * C[v][0] + Q0[1] * C[v][1] + Q0[2] * C[v][2];
                    out[1][v] = Q1[0] * C[v][0] + Q1[1] * C[v][1] + Q1[2] * C[v][2];
                }
            };
            double C11[2][3], C10[2][3];
            cc(j11, C11);
            cc(j10, C10);
            double e0j01[3], e1j01[3], e0j00[3], e1j00[3];
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                e0j01[r] = E0[r] * j01[r]; e1j01[r] = E1[r] * j01[r];
                e0j00[r] = E0[r] * j00[r]; e1j00[r] = E1[r] * j00[r];
            }
            double F[3];
            {   // gxx: fields j11^2 (xi), j11 j01 (mixed), j01^2 (eta)
                double AAx[3], BBx[3], ABx[2][2];
#pragma unroll
                for (int q = 0; q < 3; ++q) F[q] = j11[q] * j11[q];
                aa(F, AAx);
#pragma unroll
                for (int q = 0; q < 3; ++q) F[q] = j01[q] * j01[q];
                bb(F, BBx);
                ab(C11, e0j01, e1j01, ABx);
                gxx[0] = AAx[0] - ABx[0][0] - ABx[0][0] + BBx[0];
                gxx[1] = - AAx[0] - ABx[0][1] + ABx[0][0] + BBx[1];
                gxx[2] = - AAx[1] + ABx[0][1] + ABx[1][0] - BBx[1];
                gxx[3] = AAx[1] + ABx[0][0] - ABx[1][0] - BBx[0];
                gxx[4] = AAx[0] + ABx[0][1] + ABx[0][1] + BBx[2];
                gxx[5] = AAx[1] - ABx[0][1] + ABx[1][1] - BBx[2];
                gxx[6] = - AAx[1] - ABx[0][0] - ABx[1][1] - BBx[1];
                gxx[7] = AAx[2] - ABx[1][1] - ABx[1][1] + BBx[2];
                gxx[8] = - AAx[2] - ABx[1][0] + ABx[1][1] + BBx[1];
                gxx[9] = AAx[2] + ABx[1][0] + ABx[1][0] + BBx[0];
            }
            {   // gyy: fields j10^2, j10 j00, j00^2
                double AAy[3], BBy[3], ABy[2][2];
#pragma unroll
                for (int q = 0; q < 3; ++q) F[q] = j10[q] * j10[q];
                aa(F, AAy);
#pragma unroll
                for (int q = 0; q < 3; ++q) F[q] = j00[q] * j00[q];
                bb(F, BBy);
                ab(C10, e0j00, e1j00, ABy);
                gyy[0] = BBy[0] - ABy[0][0] - ABy[0][0] + AAy[0];
                gyy[1] = BBy[1] + ABy[0][0] - ABy[0][1] - AAy[0];
                gyy[2] = - BBy[1] + ABy[1][0] + ABy[0][1] - AAy[1];
                gyy[3] = - BBy[0] - ABy[1][0] + ABy[0][0] + AAy[1];
                gyy[4] = BBy[2] + ABy[0][1] + ABy[0][1] + AAy[0];
                gyy[5] = - BBy[2] + ABy[1][1] - ABy[0][1] + AAy[1];
                gyy[6] = - BBy[1] - ABy[1][1] - ABy[0][0] - AAy[1];
                gyy[7] = BBy[2] - ABy[1][1] - ABy[1][1] + AAy[2];
                gyy[8] = BBy[1] + ABy[1][1] - ABy[1][0] - AAy[2];
                gyy[9] = BBy[0] + ABy[1][0] + ABy[1][0] + AAy[2];
            }
            {   // gxy: fields j11 j10, j01 j00, j11 j00, j10 j01
                double AAm[3], BBm[3], ABp[2][2], ABq[2][2];
#pragma unroll
                for (int q = 0; q < 3; ++q) F[q] = j11[q] * j10[q];
                aa(F, AAm);
#pragma unroll
                for (int q = 0; q < 3; ++q) F[q] = j01[q] * j00[q];
                bb(F, BBm);
                ab(C11, e0j00, e1j00, ABp);
                ab(C10, e0j01, e1j01, ABq);
                gxy[0] = ABp[0][0] - AAm[0] - BBm[0] + ABq[0][0];
                gxy[1] = ABp[0][1] + AAm[0] - BBm[1] - ABq[0][0];
                gxy[2] = - ABp[0][1] + AAm[1] + BBm[1] - ABq[1][0];
                gxy[3] = - ABp[0][0] - AAm[1] + BBm[0] + ABq[1][0];
                gxy[4] = - ABp[0][0] + AAm[0] - BBm[1] + ABq[0][1];
                gxy[5] = - ABp[0][1] - AAm[0] - BBm[2] - ABq[0][1];
                gxy[6] = ABp[0][1] - AAm[1] + BBm[2] - ABq[1][1];
                gxy[7] = ABp[0][0] + AAm[1] + BBm[1] + ABq[1][1];
                gxy[8] = - ABp[1][0] + AAm[1] + BBm[1] - ABq[0][1];
                gxy[9] = - ABp[1][1] - AAm[1] + BBm[2] + ABq[0][1];
                gxy[10] = ABp[1][1] - AAm[2] - BBm[2] + ABq[1][1];
                gxy[11] = ABp[1][0] + AAm[2] - BBm[1] - ABq[1][1];
                gxy[12] = ABp[1][0] - AAm[1] + BBm[0] - ABq[0][0];
                gxy[13] = ABp[1][1] + AAm[1] + BBm[1] + ABq[0][0];
                gxy[14] = - ABp[1][1] + AAm[2] - BBm[1] + ABq[1][0];
                gxy[15] = - ABp[1][0] - AAm[2] - BBm[0] - ABq[1][0];
            }
        }
        if (DO_M) {
            // N_i = 4 f(xi) e(eta): the same separable evaluation is not worth it here (the mass kernel is HBM-bound)
            const double gp[3] = {-A, 0.0, A};
            const double sx[4] = {-1, 1, 1, -1}, sy[4] = {-1, -1, 1, 1};
#pragma unroll
            for (int q = 0; q < 10; ++q) mm[q] = 0;
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    double nn[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) nn[i] = (1 + sx[i] * gp[c]) * (1 + sy[i] * gp[r]) / 4;
                    const double cm = (GW[r] * GW[c]) * mat.rho * mat.t * ad[r][c];
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int j = i; j < 4; ++j) mm[tri_index<4>(i, j)] += (cm * nn[i]) * nn[j];
                }
        }
        return ok;
    }
    // {K(2i,2j), K(2i,2j+1), K(2i+1,2j), K(2i+1,2j+1)}.  The roundings are spelled out (one rounded product, one fused
    // multiply-add) so that every instance of the expression rounds the same way: Ke(2i+p, 2j+q) and Ke(2j+q, 2i+p) are
    // then the same bits, and so are K(r, c) and K(c, r) after the assembly (same contributions, same order) -- the
    // symmetric-lattice SpMV (psfem_spmv_sl.cuh) relies on it.  Left to the compiler, the contraction of a*b + c*d
    // differed between inlined instances (7 % of the entries of K differed from their mirror image by one ulp).
    __device__ __forceinline__ double4 kblock(int i, int j, const Mat &mat) const {
        const int b = i <= j ? tri_index<4>(i, j) : tri_index<4>(j, i);
        return make_double4(fma(mat.d00, gxx[b], __dmul_rn(mat.d22, gyy[b])), fma(mat.d01, gxy[i * 4 + j], __dmul_rn(mat.d22, gxy[j * 4 + i])),
                            fma(mat.d01, gxy[j * 4 + i], __dmul_rn(mat.d22, gxy[i * 4 + j])), fma(mat.d00, gyy[b], __dmul_rn(mat.d22, gxx[b])));
    }
    __device__ __forceinline__ double mval(int i, int j) const { return mm[i <= j ? tri_index<4>(i, j) : tri_index<4>(j, i)]; }
};

// T3 closed form (Polygones.py:840-883, :1218-1235), signed area
template <bool DO_K, bool DO_M> struct ElemT3 {
    double b[3], g[3], vol, mbase;
    double db0[3], db1[3], db2[3], dg0[3], dg1[3], dg2[3];   // D B: {d00, d01, d22} x {b_j, c_j}
    __device__ __forceinline__ bool compute(const double2 (&p)[3], const Mat &mat) {
        // signed 2*area, Polygones.py:843-845, products and sums rounded one by one like the reference (no FMA contraction)
        const double delta = __dadd_rn(__dadd_rn(__dmul_rn(p[0].x, p[1].y - p[2].y), __dmul_rn(p[1].x, p[2].y - p[0].y)),
                                       __dmul_rn(p[2].x, p[0].y - p[1].y));
        vol = 0.5 * delta * mat.t;
        // b_i = (y_j - y_k)/delta, c_i = (x_k - x_j)/delta (Polygones.py:848-850): ONE IEEE division for the
        // correctly rounded reciprocal, then a Markstein correction step per quotient (q' = q + (a - q*delta)*inv),
        // which returns the correctly rounded a/delta -- the reference's six divisions bit for bit, at 3 operations
        // each instead of the ~20 of the division subroutine.  Signed delta as the reference.
        const double inv = 1.0 / delta;
        auto quot = [&](double a) {
            const double q = a * inv;
            return fma(fma(-q, delta, a), inv, q);
        };
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            b[i] = quot(p[(i + 1) % 3].y - p[(i + 2) % 3].y);
            g[i] = quot(p[(i + 2) % 3].x - p[(i + 1) % 3].x);
        }
        if (DO_K) {
#pragma unroll
            for (int j = 0; j < 3; ++j) {   // columns 2j, 2j+1 of D B (:881): single rounded products
                db0[j] = __dmul_rn(mat.d00, b[j]); db1[j] = __dmul_rn(mat.d01, b[j]); db2[j] = __dmul_rn(mat.d22, b[j]);
                dg0[j] = __dmul_rn(mat.d00, g[j]); dg1[j] = __dmul_rn(mat.d01, g[j]); dg2[j] = __dmul_rn(mat.d22, g[j]);
            }
        }
        mbase = mat.rho * vol / 12;   // Me*rho*Volume/12, Polygones.py:1234
        return true;
    }
    // Ke = np.dot(B.T, np.dot(D, B)) * Volume (Polygones.py:877-882) with the operation order of the reference's
    // BLAS: every entry of B^T (D B) is a 3-term dot product accumulated with FMAs in k order, one term of which is a
    // structural zero; the product with the volume is rounded separately (never contracted into the assembly add).
    // The assembled T3 matrix equals the reference's bit for bit (tests/test_gpu_parity.py, golden README mesh).
    __device__ __forceinline__ double4 kblock(int i, int j, const Mat &) const {
        return make_double4(__dmul_rn(fma(g[i], dg2[j], __dmul_rn(b[i], db0[j])), vol),
                            __dmul_rn(fma(g[i], db2[j], __dmul_rn(b[i], dg1[j])), vol),
                            __dmul_rn(fma(b[i], dg2[j], __dmul_rn(g[i], db1[j])), vol),
                            __dmul_rn(fma(b[i], db2[j], __dmul_rn(g[i], dg0[j])), vol));
    }
    __device__ __forceinline__ double mval(int i, int j) const { return (i == j ? 2.0 : 1.0) * mbase; }
};

}  // namespace psfem
