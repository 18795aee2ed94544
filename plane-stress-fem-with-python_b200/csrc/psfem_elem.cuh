// psfem_elem.cuh -- per-element arithmetic of the T3 / Q4 kernels (shared by the tile-fused and the
// lattice assembly): gradient products of Ke and shape products of Me, computed once per element in registers.
#pragma once
#include "psfem_common.cuh"

namespace psfem {

struct Mat {
    double d00, d01, d22, t, rho;
};

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

// Q4: gradient products of the element, integrated with the 3x3 rule.  Same arithmetic as gauss_kernel
// (Polygones.py:582-595, :657-708, :1266-1291) with the element computed once; D is applied when a 2x2
// block is scattered.
template <bool DO_K, bool DO_M> struct ElemQ4 {
    double gxx[10], gyy[10], gxy[16], mm[10];
    __device__ __forceinline__ bool compute(const double2 (&p)[4], const Mat &mat) {
        constexpr double A = 0.7745966692414834;  // sqrt(0.6)  (Polygones.py:549)
        const double gp[3] = {-A, 0.0, A};
        const double gw[3] = {5.0 / 9, 8.0 / 9, 5.0 / 9};
        const double sx[4] = {-1, 1, 1, -1}, sy[4] = {-1, -1, 1, 1};
#pragma unroll
        for (int q = 0; q < 10; ++q) { gxx[q] = 0; gyy[q] = 0; mm[q] = 0; }
#pragma unroll
        for (int q = 0; q < 16; ++q) gxy[q] = 0;
        bool ok = true;
#pragma unroll
        for (int r = 0; r < 3; ++r) {      // eta (outer, Polygones.py:551-553)
            const double eta = gp[r];
            const double wr = gw[r];
#pragma unroll
            for (int c = 0; c < 3; ++c) {  // xi
                const double xi = gp[c];
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
                const double w = wr * gw[c];
                if (DO_K) {
                    const double inv = fast_rcp(detJ);
                    const double pp = j11 * inv, qq = -j01 * inv, rr = -j10 * inv, ss = j00 * inv;
                    const double cc = w * ad * mat.t;
                    double dx[4], dy[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        dx[i] = pp * dnx[i] + qq * dne[i];
                        dy[i] = rr * dnx[i] + ss * dne[i];
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const double cxi = cc * dx[i], cyi = cc * dy[i];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            if (j >= i) {
                                gxx[tri_index<4>(i, j)] += cxi * dx[j];
                                gyy[tri_index<4>(i, j)] += cyi * dy[j];
                            }
                            gxy[i * 4 + j] += cxi * dy[j];
                        }
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
        return ok;
    }
    // {K(2i,2j), K(2i,2j+1), K(2i+1,2j), K(2i+1,2j+1)}
    __device__ __forceinline__ double4 kblock(int i, int j, const Mat &mat) const {
        const int b = i <= j ? tri_index<4>(i, j) : tri_index<4>(j, i);
        return make_double4(mat.d00 * gxx[b] + mat.d22 * gyy[b], mat.d01 * gxy[i * 4 + j] + mat.d22 * gxy[j * 4 + i],
                            mat.d01 * gxy[j * 4 + i] + mat.d22 * gxy[i * 4 + j], mat.d00 * gyy[b] + mat.d22 * gxx[b]);
    }
    __device__ __forceinline__ double mval(int i, int j) const { return mm[i <= j ? tri_index<4>(i, j) : tri_index<4>(j, i)]; }
};

// T3 closed form (Polygones.py:840-883, :1218-1235), signed area
template <bool DO_K, bool DO_M> struct ElemT3 {
    double b[3], g[3], vol, mbase;
    __device__ __forceinline__ bool compute(const double2 (&p)[3], const Mat &mat) {
        const double delta = p[0].x * (p[1].y - p[2].y) + p[1].x * (p[2].y - p[0].y) + p[2].x * (p[0].y - p[1].y);
        vol = 0.5 * delta * mat.t;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            b[i] = (p[(i + 1) % 3].y - p[(i + 2) % 3].y) / delta;
            g[i] = (p[(i + 2) % 3].x - p[(i + 1) % 3].x) / delta;
        }
        mbase = mat.rho * vol / 12;   // T3 keeps IEEE divisions: the arithmetic of Polygones.py:848-850 verbatim
        return true;
    }
    __device__ __forceinline__ double4 kblock(int i, int j, const Mat &mat) const {
        return make_double4((mat.d00 * b[i] * b[j] + mat.d22 * g[i] * g[j]) * vol, (mat.d01 * b[i] * g[j] + mat.d22 * g[i] * b[j]) * vol,
                            (mat.d01 * g[i] * b[j] + mat.d22 * b[i] * g[j]) * vol, (mat.d00 * g[i] * g[j] + mat.d22 * b[i] * b[j]) * vol);
    }
    __device__ __forceinline__ double mval(int i, int j) const { return (i == j ? 2.0 : 1.0) * mbase; }
};

}  // namespace psfem
