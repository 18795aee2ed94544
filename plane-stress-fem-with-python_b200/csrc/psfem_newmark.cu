// psfem_newmark.cu -- Newmark-beta time loop exactly as main.py:264-332 implements it.
//
// NOT the textbook scheme: K_eff = K + M/(beta dt^2) has no damping term (:290), the
// solve result is stored as the acceleration (:319), Rayleigh damping C = alpha_R M + beta_R K enters
// the right-hand side only (:264-266,:310), and a data-dependent rescaling "stabiliser" runs each
// step (:324-332).  Per step: predictor kernel, one fused two-matrix SpMV for the right-hand side
//   F_i - K (u_p + beta_R v_p) - alpha_R M v_p,
// a Jacobi-PCG solve warm-started from the previous acceleration (replaces lu_solve :314), corrector
// kernel with the stabiliser's reductions, store kernel.  Energies (:377-385) are two SpMV+dot.
#include <cstdlib>
#include "psfem_common.cuh"

namespace psfem {
int pcg_solve(const psfem_csr *A, const double *d_b, double *d_x, double rtol, int64_t maxit, int64_t check_every,
              double *h_info, double *d_info, double *h_acc, void *d_ws, size_t ws_bytes, bool have_dinv, cudaStream_t stream,
              bool masked, const int32_t *d_new_id);
double *pcg_dinv_ptr(void *d_ws, int64_t n, int64_t n_blocks);
}
using namespace psfem;

namespace {
constexpr int NT = 256;

int nm_grid(int64_t n) {
    int64_t g = ceil_div(n, NT * 2);
    if (g > 148 * 8) g = 148 * 8;
    return (int)(g < 1 ? 1 : g);
}

struct NmCtl {
    double hist[10];  // sum |U[:,k]| of the last 10 stored steps (main.py:325)
    double scale;     // stabiliser factor of the current step (1 = not triggered)
    double triggers;
};

// main.py:306-307 (+ the vector K is applied to: u_p + beta_R v_p)
__global__ void predictor_kernel(int64_t n, const double *__restrict__ u, const double *__restrict__ v, const double *__restrict__ a,
                                 double dt, double c_u, double c_v, double beta_R, double *__restrict__ up,
                                 double *__restrict__ vp, double *__restrict__ w1) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double ai = a[i], vi = v[i];
        const double upi = u[i] + dt * vi + c_u * ai;   // c_u = (0.5-beta) dt^2
        const double vpi = vi + c_v * ai;               // c_v = (1-gamma) dt
        up[i] = upi; vp[i] = vpi; w1[i] = upi + beta_R * vpi;
    }
}

// start vector of the per-step PCG solve: the linear extrapolation 2 a_{i-1} - a_{i-2} of the last two accelerations instead
// of a_{i-1} (the solve still runs to rtol, so the result is the same to that tolerance; a third fewer iterations per step on
// smooth load histories).  In place: a <- 2a - aprev, aprev <- a.
__global__ void extrapolate_kernel(int64_t n, double *__restrict__ a, double *__restrict__ aprev) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double ai = a[i], pi = aprev[i];
        a[i] = 2.0 * ai - pi;
        aprev[i] = ai;
    }
}

// main.py:319-321 + the reductions of the stabiliser test :324-331
__global__ void __launch_bounds__(NT)
corrector_kernel(int64_t n, const double *__restrict__ up, const double *__restrict__ vp, const double *__restrict__ a,
                 double c_v, double c_u, double *__restrict__ u, double *__restrict__ v, double *__restrict__ partials,
                 unsigned int *counter, NmCtl *ctl, int64_t step, double thresh) {
    __shared__ double red[NT / 32];
    __shared__ bool is_last;
    double mx = 0, sm = 0;
    for (int64_t i = blockIdx.x * (int64_t)NT + threadIdx.x; i < n; i += (int64_t)gridDim.x * NT) {
        const double ai = a[i];
        const double vi = vp[i] + c_v * ai;  // gamma dt
        const double ui = up[i] + c_u * ai;  // beta dt^2
        v[i] = vi; u[i] = ui;
        mx = fmax(mx, fabs(ui)); sm += fabs(ui);
    }
    // block max (order-independent) and deterministic block sum
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x == 0) { for (int w = 1; w < NT / 32; ++w) mx = fmax(mx, red[w]); }
    double s = block_sum<NT>(sm, red);
    const int nb = gridDim.x;
    if (threadIdx.x == 0) {
        partials[blockIdx.x] = mx;
        partials[nb + blockIdx.x] = s;
        __threadfence();
        is_last = (atomicAdd(counter, 1u) == (unsigned)nb - 1);
    }
    __syncthreads();
    if (is_last && threadIdx.x == 0) {
        __threadfence();
        double gmax = 0, gsum = 0;
        for (int b = 0; b < nb; ++b) { gmax = fmax(gmax, __ldcg(partials + b)); gsum += __ldcg(partials + nb + b); }
        double scale = 1.0;
        if (step > 10) {
            double h = 0;
            for (int k = 0; k < 10; ++k) h += ctl->hist[k];
            const double mean = h / (10.0 * (double)n);  // np.mean(np.abs(U[:, i-10:i]))
            if (gmax > thresh * mean) { scale = 10.0 * mean / gmax; ctl->triggers += 1.0; }   // 100 in the Newmark loop (:325), 10 in the explicit one (:365)
        }
        ctl->scale = scale;
        ctl->hist[step % 10] = gsum * scale;  // column i is stored scaled (:329)
        *counter = 0;
    }
}

// apply the stabiliser factor (:329-330) and store the step's columns
__global__ void store_kernel(int64_t n, double *__restrict__ u, double *__restrict__ v, const double *__restrict__ a,
                             const NmCtl *__restrict__ ctl, double *__restrict__ Uc, double *__restrict__ Vc, double *__restrict__ Ac) {
    const double s = ctl->scale;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double ui = u[i], vi = v[i];
        if (s != 1.0) { ui *= s; vi *= s; u[i] = ui; v[i] = vi; }
        if (Uc) Uc[i] = ui;
        if (Vc) Vc[i] = vi;
        if (Ac) Ac[i] = a[i];
    }
}

__global__ void energy_finish_kernel(int64_t n_steps, double *__restrict__ e) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n_steps) return;
    const double kin = 0.5 * e[3 * i], pot = 0.5 * e[3 * i + 1];  // main.py:383-385
    e[3 * i] = kin; e[3 * i + 1] = pot; e[3 * i + 2] = kin + pot;
}


// ---------------------------------------------------------------------------------------------
// Small systems (the reference's own Newmark cases: main.py runs 10 001 steps on a few hundred DOFs): the WHOLE time
// loop in one persistent single-CTA kernel.  K_eff = K + M/(beta dt^2) is banded (node id = column*m + row); its
// Cholesky factor L is computed once per launch in shared memory (band storage Lb[i][k] = L[i][i-k], k = 0..hb) and
// every step solves K_eff x = rhs by a forward and a backward substitution of warp 0 -- the direct solve main.py itself
// does with lu_factor / lu_solve (:295, :314), instead of a PCG solve to 1e-13 per step (1.4 ms per step for the 410
// DOFs of BASELINE config 2; here about 0.1 ms).  Predictor, right-hand side, corrector, stabiliser, stores and the
// two energy products of a step run in the same kernel between block barriers; no launch per step.
// ---------------------------------------------------------------------------------------------
constexpr int NB_T = 1024;          // threads of the single CTA
constexpr int NB_MAX_HB = 32;       // sub-diagonals of the band (one warp lane per term of a substitution row)
constexpr int NB_CHUNK = 2048;      // time steps per launch (their load factors travel through the work space)
constexpr size_t NB_SMEM_MAX = 200 * 1024;

struct NbArgs {
    int n, hb;
    const int32_t *row_ptr, *col_idx;
    const double *K, *M, *Keff, *fshape, *fs;   // fs: load factors of this launch's steps
    long long step0, nsteps, store_every;
    double dt, c_u_pred, c_v_pred, c_v_corr, c_u_corr, alpha_R, beta_R, thresh;
    double *u, *v, *a, *up, *vp, *w1;
    NmCtl *ctl;
    double *U, *V, *A, *energy;
    double *info;                                // [4] = 1: factorisation failed (not positive definite), nothing was done
};

__global__ void band_width_kernel(int n, const int32_t *__restrict__ row_ptr, const int32_t *__restrict__ col_idx, int *__restrict__ out) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    int hb = 0;
    for (int k = row_ptr[r]; k < row_ptr[r + 1]; ++k) hb = max(hb, abs(r - col_idx[k]));
    atomicMax(out, hb);
}

__global__ void __launch_bounds__(NB_T)
newmark_band_kernel(const NbArgs g) {
    extern __shared__ double sm[];
    __shared__ double red[NB_T / 32];
    __shared__ double bcast[4];
    __shared__ int fail;
    const int n = g.n, hb = g.hb, w = hb + 1, t = threadIdx.x, lane = t & 31;
    double *Lb = sm;                      // n * (hb+1)
    double *invd = Lb + (size_t)n * w;    // n
    double *ys = invd + n;                // n: right-hand side / solution of the current step
    // ---- band of K_eff (lower triangle) ----
    for (int i = t; i < n * w; i += NB_T) Lb[i] = 0.0;
    if (t == 0) fail = 0;
    __syncthreads();
    for (int r = t; r < n; r += NB_T)
        for (int k = g.row_ptr[r]; k < g.row_ptr[r + 1]; ++k) {
            const int c = g.col_idx[k];
            if (c <= r) Lb[(size_t)r * w + (r - c)] = g.Keff[k];
        }
    __syncthreads();
    // ---- right-looking banded Cholesky: column j, scale, rank-1 update of the trailing (hb x hb) triangle ----
    for (int j = 0; j < n; ++j) {
        const double ajj = Lb[(size_t)j * w];
        if (!(ajj > 0.0)) { if (t == 0) fail = 1; }
        __syncthreads();
        if (fail) break;
        const double d = sqrt(ajj), id = 1.0 / d;
        const int na = min(hb, n - 1 - j);           // rows below the diagonal in this column
        if (t < na) Lb[(size_t)(j + 1 + t) * w + (t + 1)] *= id;
        if (t == NB_T - 1) { Lb[(size_t)j * w] = d; invd[j] = id; }
        __syncthreads();
        // pairs (a, b), 1 <= b <= a <= na:  A[j+a][j+b] -= L[j+a][j] * L[j+b][j]
        for (int q = t; q < na * (na + 1) / 2; q += NB_T) {
            int a = (int)((sqrt(8.0 * q + 1.0) - 1.0) * 0.5);
            while (a * (a + 1) / 2 > q) --a;
            while ((a + 1) * (a + 2) / 2 <= q) ++a;
            const int b = q - a * (a + 1) / 2;       // 0 <= b <= a
            const int ra = j + 1 + a, rb = j + 1 + b;
            Lb[(size_t)ra * w + (a - b)] -= Lb[(size_t)ra * w + (a + 1)] * Lb[(size_t)rb * w + (b + 1)];
        }
        __syncthreads();
    }
    if (fail) { if (t == 0) g.info[4] = 1.0; return; }
    auto bsum = [&](double x) -> double {          // block sum, broadcast to every thread
        const double s = block_sum<NB_T>(x, red);
        if (t == 0) bcast[0] = s;
        __syncthreads();
        const double o = bcast[0];
        __syncthreads();
        return o;
    };
    for (long long sidx = 0; sidx < g.nsteps; ++sidx) {
        const long long step = g.step0 + sidx;
        const double fsc = g.fs[sidx];
        // predictor (main.py:306-307)
        for (int i = t; i < n; i += NB_T) {
            const double ai = g.a[i], vi = g.v[i];
            const double upi = g.u[i] + g.dt * vi + g.c_u_pred * ai;
            const double vpi = vi + g.c_v_pred * ai;
            g.up[i] = upi; g.vp[i] = vpi; g.w1[i] = upi + g.beta_R * vpi;
        }
        __syncthreads();
        // right-hand side F_i - K (u_p + beta_R v_p) - alpha_R M v_p (main.py:310)
        for (int r = t; r < n; r += NB_T) {
            double s1 = 0.0, s2 = 0.0;
            for (int k = g.row_ptr[r]; k < g.row_ptr[r + 1]; ++k) {
                const int c = g.col_idx[k];
                s1 += g.K[k] * g.w1[c];
                s2 += g.M[k] * g.vp[c];
            }
            ys[r] = (fsc * g.fshape[r] - s1) - g.alpha_R * s2;
        }
        __syncthreads();
        // K_eff a = rhs: L y = rhs, L^T a = y (warp 0; lane k carries term k of the row)
        if (t < 32) {
            for (int i = 0; i < n; ++i) {
                double term = 0.0;
                if (lane < hb && i - 1 - lane >= 0) term = Lb[(size_t)i * w + (lane + 1)] * ys[i - 1 - lane];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) term += __shfl_xor_sync(0xffffffffu, term, o);
                if (lane == 0) ys[i] = (ys[i] - term) * invd[i];
                __syncwarp();
            }
            for (int i = n - 1; i >= 0; --i) {
                double term = 0.0;
                if (lane < hb && i + 1 + lane < n) term = Lb[(size_t)(i + 1 + lane) * w + (lane + 1)] * ys[i + 1 + lane];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) term += __shfl_xor_sync(0xffffffffu, term, o);
                if (lane == 0) ys[i] = (ys[i] - term) * invd[i];
                __syncwarp();
            }
        }
        __syncthreads();
        // corrector (main.py:319-321) + the reductions of the stabiliser test (:324-331)
        double mx = 0.0, sa = 0.0;
        for (int i = t; i < n; i += NB_T) {
            const double ai = ys[i];
            g.a[i] = ai;
            const double vi = g.vp[i] + g.c_v_corr * ai;
            const double ui = g.up[i] + g.c_u_corr * ai;
            g.v[i] = vi; g.u[i] = ui;
            mx = fmax(mx, fabs(ui)); sa += fabs(ui);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        __syncthreads();
        if (lane == 0) red[t >> 5] = mx;
        __syncthreads();
        if (t == 0) { for (int q = 1; q < NB_T / 32; ++q) mx = fmax(mx, red[q]); bcast[1] = mx; }
        __syncthreads();
        const double gmax = bcast[1];
        const double gsum = bsum(sa);
        if (t == 0) {
            double scale = 1.0;
            if (step > 10) {
                double h = 0;
                for (int k = 0; k < 10; ++k) h += g.ctl->hist[k];
                const double mean = h / (10.0 * (double)n);  // np.mean(np.abs(U[:, i-10:i]))
                if (gmax > g.thresh * mean) { scale = 10.0 * mean / gmax; g.ctl->triggers += 1.0; }
            }
            g.ctl->scale = scale;
            g.ctl->hist[step % 10] = gsum * scale;           // column i is stored scaled (:329)
            bcast[2] = scale;
        }
        __syncthreads();
        const double scale = bcast[2];
        const bool st = (step % g.store_every) == 0;
        const long long col = step / g.store_every;
        for (int i = t; i < n; i += NB_T) {
            double ui = g.u[i], vi = g.v[i];
            if (scale != 1.0) { ui *= scale; vi *= scale; g.u[i] = ui; g.v[i] = vi; }
            if (st && g.U) g.U[col * n + i] = ui;
            if (st && g.V) g.V[col * n + i] = vi;
            if (st && g.A) g.A[col * n + i] = ys[i];
        }
        __syncthreads();
        if (g.energy) {   // main.py:383-384: v.(M v), u.(K u); halved and summed by energy_finish_kernel
            double ek = 0.0, ep = 0.0;
            for (int r = t; r < n; r += NB_T) {
                double s1 = 0.0, s2 = 0.0;
                for (int k = g.row_ptr[r]; k < g.row_ptr[r + 1]; ++k) {
                    const int c = g.col_idx[k];
                    s1 += g.M[k] * g.v[c];
                    s2 += g.K[k] * g.u[c];
                }
                ek += g.v[r] * s1; ep += g.u[r] * s2;
            }
            const double ekt = bsum(ek), ept = bsum(ep);
            if (t == 0) { g.energy[3 * step] = ekt; g.energy[3 * step + 1] = ept; }
        }
        __syncthreads();
    }
}

struct NmWs {
    double *u, *v, *a, *up, *vp, *w1, *rhs, *tmp, *aprev, *partials, *info, *fs;
    NmCtl *ctl;
    unsigned int *counter;
    void *pcg;
    size_t pcg_bytes, total;
};

NmWs carve_nm(void *ws, int64_t n, int64_t n_blocks) {
    NmWs w;
    Carver c(ws);
    w.u = c.take<double>(n); w.v = c.take<double>(n); w.a = c.take<double>(n);
    w.up = c.take<double>(n); w.vp = c.take<double>(n); w.w1 = c.take<double>(n);
    w.rhs = c.take<double>(n); w.tmp = c.take<double>(n); w.aprev = c.take<double>(n);
    int64_t np = n_blocks > 148 * 8 ? n_blocks : 148 * 8;
    w.partials = c.take<double>(2 * np);
    w.info = c.take<double>(8);
    w.fs = c.take<double>(NB_CHUNK);
    w.ctl = c.take<NmCtl>(1);
    w.counter = c.take<unsigned int>(4);
    psfem_pcg_workspace(n, n_blocks, &w.pcg_bytes);
    w.pcg = c.take<char>(w.pcg_bytes);
    w.total = c.used();
    return w;
}
}  // namespace

extern "C" int psfem_pcg_spmv_dot(const psfem_csr *A, const double *d_p_ext, int64_t x_offset, double *d_q, double *d_pq,
                                  double *d_partials, uint32_t *d_counter, void *stream_);

extern "C" int psfem_newmark_workspace(int64_t n, int64_t n_blocks, size_t *bytes) {
    PSFEM_REQUIRE(bytes && n >= 0, "newmark_workspace: bad arguments");
    *bytes = carve_nm(nullptr, n, n_blocks).total;
    return PSFEM_OK;
}

// The time loop shared by the two integrators of main.py.  Per step:
//   u_p = u + dt v + c_u_pred a,  v_p = v + c_v_pred a,  rhs = F_i - K (u_p + beta_R v_p) - alpha_R M v_p,
//   a = S^-1 rhs,  v = v_p + c_v_corr a,  u = u_p + c_u_corr a,  stabiliser with threshold `thresh`, stores, energies.
// Newmark-beta (:304-332): S = K_eff, c_u_pred = (1/2 - beta) dt^2, c_v_pred = (1 - gamma) dt, c_v_corr = gamma dt,
// c_u_corr = beta dt^2, thresh = 100.  Modified explicit Euler (:346-372): S = M, u_p = u + dt v is the new
// displacement itself (c_u_pred = c_u_corr = 0), v = v + dt/2 (a_old + a_new) (c_v_pred = c_v_corr = dt/2), no damping,
// thresh = 10.
struct LoopConsts { double c_u_pred, c_v_pred, c_v_corr, c_u_corr, alpha_R, beta_R, thresh; };

static int time_loop(const psfem_csr *K, const psfem_csr *M, const psfem_csr *Keff_, const LoopConsts lc,
                     const double *d_fshape, const double *h_fscale, int64_t n_steps, double dt, const double *d_a0, int64_t store_every,
                     double *d_U, double *d_V, double *d_A, double *d_energy, double rtol, int64_t maxit,
                     double *h_info, void *d_ws, size_t ws_bytes, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const double alpha_R = lc.alpha_R, beta_R = lc.beta_R;
    PSFEM_REQUIRE(K && M && Keff_ && M->vals && Keff_->vals && d_fshape && h_fscale && h_info, "newmark: null argument");
    PSFEM_REQUIRE(M->n_rows == K->n_rows && Keff_->n_rows == K->n_rows && M->nnz == K->nnz && Keff_->nnz == K->nnz,
                  "newmark: K, M and K_eff must share one sparsity pattern");
    PSFEM_REQUIRE(n_steps >= 1 && store_every >= 1, "newmark: bad step counts");
    const int64_t n = K->n_rows;
    NmWs w = carve_nm(d_ws, n, K->n_blocks);
    PSFEM_REQUIRE(ws_bytes >= w.total, "newmark: workspace too small (%zu < %zu)", ws_bytes, w.total);
    // M and K_eff arrive as operands of their own: same pattern as K, their own values and (optionally) their own
    // symmetric-lattice storage (psfem_spmv_sl_build), which the SpMV uses when it matches the value array
    const psfem_csr Keff = *Keff_, Mm = *M;
    const double *d_M_vals = M->vals;
    // right-hand side in two passes of the fast single-matrix kernels when both operands have them
    const bool two_pass = K->sl_vals && K->sl_src_vals == K->vals && Mm.sl_vals && Mm.sl_src_vals == Mm.vals;
    const int g = nm_grid(n);
    PSFEM_CUDA_CHECK(cudaMemsetAsync(w.u, 0, n * sizeof(double), stream));   // main.py:282-283
    PSFEM_CUDA_CHECK(cudaMemsetAsync(w.v, 0, n * sizeof(double), stream));
    if (d_a0) PSFEM_CUDA_CHECK(cudaMemcpyAsync(w.a, d_a0, n * sizeof(double), cudaMemcpyDeviceToDevice, stream));  // :286
    else PSFEM_CUDA_CHECK(cudaMemsetAsync(w.a, 0, n * sizeof(double), stream));
    PSFEM_CUDA_CHECK(cudaMemsetAsync(w.info, 0, 8 * sizeof(double), stream));
    PSFEM_CUDA_CHECK(cudaMemsetAsync(w.ctl, 0, sizeof(NmCtl), stream));
    PSFEM_CUDA_CHECK(cudaMemsetAsync(w.counter, 0, 4 * sizeof(unsigned int), stream));
    if (d_energy) PSFEM_CUDA_CHECK(cudaMemsetAsync(d_energy, 0, 3 * n_steps * sizeof(double), stream));
    // column 0
    if (d_U) PSFEM_CUDA_CHECK(cudaMemsetAsync(d_U, 0, n * sizeof(double), stream));
    if (d_V) PSFEM_CUDA_CHECK(cudaMemsetAsync(d_V, 0, n * sizeof(double), stream));
    if (d_A) PSFEM_CUDA_CHECK(cudaMemcpyAsync(d_A, w.a, n * sizeof(double), cudaMemcpyDeviceToDevice, stream));
    // 1/diag(K_eff) once, into the PCG work space (pcg_solve is told it is there)
    {
        int rc = psfem_extract_diag(&Keff, 0, 1, pcg_dinv_ptr(w.pcg, n, K->n_blocks), stream);
        if (rc) return rc;
    }
    double h_acc[3] = {0, 0, 0};
    const double c_u_pred = lc.c_u_pred, c_v_pred = lc.c_v_pred, c_v_corr = lc.c_v_corr, c_u_corr = lc.c_u_corr;
    int64_t first_step = 1;
    static const bool band_off = getenv("PSFEM_NEWMARK_NO_BAND") != nullptr;   // A/B switch
    if (n <= 16384 && n_steps > 1 && !band_off) {
        // small banded system: the whole time loop in newmark_band_kernel (direct solves with the Cholesky factor of K_eff)
        int *d_hb = reinterpret_cast<int *>(w.counter + 2);
        band_width_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, stream>>>((int)n, K->row_ptr, K->col_idx, d_hb);
        PSFEM_LAUNCH_CHECK();
        int hb = 0;
        PSFEM_CUDA_CHECK(cudaMemcpyAsync(&hb, d_hb, sizeof(int), cudaMemcpyDeviceToHost, stream));
        PSFEM_CUDA_CHECK(cudaStreamSynchronize(stream));
        PSFEM_CUDA_CHECK(cudaMemsetAsync(d_hb, 0, sizeof(int), stream));
        const size_t smem = ((size_t)n * (hb + 1) + 2 * (size_t)n) * sizeof(double);
        if (hb <= NB_MAX_HB && smem <= NB_SMEM_MAX && K->row_ptr && Keff.row_ptr == K->row_ptr && Mm.row_ptr == K->row_ptr) {
            PSFEM_CUDA_CHECK(cudaFuncSetAttribute(newmark_band_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            NbArgs g;
            g.n = (int)n; g.hb = hb; g.row_ptr = K->row_ptr; g.col_idx = K->col_idx;
            g.K = K->vals; g.M = Mm.vals; g.Keff = Keff.vals; g.fshape = d_fshape; g.fs = w.fs;
            g.store_every = store_every; g.dt = dt; g.c_u_pred = c_u_pred; g.c_v_pred = c_v_pred; g.c_v_corr = c_v_corr;
            g.c_u_corr = c_u_corr; g.alpha_R = alpha_R; g.beta_R = beta_R; g.thresh = lc.thresh;
            g.u = w.u; g.v = w.v; g.a = w.a; g.up = w.up; g.vp = w.vp; g.w1 = w.w1; g.ctl = w.ctl;
            g.U = d_U; g.V = d_V; g.A = d_A; g.energy = d_energy; g.info = w.info;
            bool ok = true;
            for (int64_t s0 = 1; s0 < n_steps && ok; s0 += NB_CHUNK) {
                const int64_t cnt = n_steps - s0 < NB_CHUNK ? n_steps - s0 : NB_CHUNK;
                PSFEM_CUDA_CHECK(cudaMemcpyAsync(w.fs, h_fscale + s0, cnt * sizeof(double), cudaMemcpyHostToDevice, stream));
                g.step0 = s0; g.nsteps = cnt;
                newmark_band_kernel<<<1, NB_T, smem, stream>>>(g);
                PSFEM_LAUNCH_CHECK();
                if (s0 == 1) {   // the factorisation of the first launch decides: not positive definite -> PCG path below
                    double flag = 0.0;
                    PSFEM_CUDA_CHECK(cudaMemcpyAsync(&flag, w.info + 4, sizeof(double), cudaMemcpyDeviceToHost, stream));
                    PSFEM_CUDA_CHECK(cudaStreamSynchronize(stream));
                    if (flag != 0.0) { ok = false; PSFEM_CUDA_CHECK(cudaMemsetAsync(w.info + 4, 0, sizeof(double), stream)); }
                }
            }
            if (ok) first_step = n_steps;   // every step done
        }
    }
    int64_t pcg_batch = 50;
    static const bool no_extrap = getenv("PSFEM_NEWMARK_NO_EXTRAPOLATION") != nullptr;   // A/B switch
    if (first_step < n_steps) PSFEM_CUDA_CHECK(cudaMemcpyAsync(w.aprev, w.a, n * sizeof(double), cudaMemcpyDeviceToDevice, stream));
    for (int64_t i = first_step; i < n_steps; ++i) {
        predictor_kernel<<<g, NT, 0, stream>>>(n, w.u, w.v, w.a, dt, c_u_pred, c_v_pred, beta_R, w.up, w.vp, w.w1);
        PSFEM_LAUNCH_CHECK();
        int rc;   // F_i - K (u_p + beta_R v_p) - alpha_R M v_p, main.py:310
        if (alpha_R == 0.0) {
            rc = psfem_spmv(K, w.w1, -1.0, h_fscale[i], d_fshape, w.rhs, stream);
        } else if (two_pass) {
            rc = psfem_spmv(K, w.w1, -1.0, h_fscale[i], d_fshape, w.rhs, stream);
            if (!rc) rc = psfem_spmv(&Mm, w.vp, -alpha_R, 1.0, w.rhs, w.rhs, stream);
        } else {
            rc = psfem_spmv2(K, d_M_vals, w.w1, w.vp, -1.0, -alpha_R, h_fscale[i], d_fshape, w.rhs, stream);
        }
        if (rc) return rc;
        if (!no_extrap) {
            extrapolate_kernel<<<g, NT, 0, stream>>>(n, w.a, w.aprev);
            PSFEM_LAUNCH_CHECK();
        }
        // batches sized by the previous step's iteration count (warm start: a handful of iterations per step)
        const double it_before = h_acc[0];
        rc = pcg_solve(&Keff, w.rhs, w.a, rtol, maxit, pcg_batch, nullptr, w.info, h_acc, w.pcg, w.pcg_bytes, true, stream, false, nullptr);  // :314
        if (h_acc[0] > it_before) { const int64_t used = (int64_t)(h_acc[0] - it_before); pcg_batch = used + 3 < 50 ? used + 3 : 50; }
        if (rc && rc != PSFEM_ENOCONV) return rc;
        corrector_kernel<<<g, NT, 0, stream>>>(n, w.up, w.vp, w.a, c_v_corr, c_u_corr, w.u, w.v, w.partials, w.counter, w.ctl, i, lc.thresh);
        PSFEM_LAUNCH_CHECK();
        const bool st = (i % store_every) == 0;
        const int64_t col = i / store_every;
        store_kernel<<<g, NT, 0, stream>>>(n, w.u, w.v, w.a, w.ctl, (st && d_U) ? d_U + col * n : nullptr,
                                           (st && d_V) ? d_V + col * n : nullptr, (st && d_A) ? d_A + col * n : nullptr);
        PSFEM_LAUNCH_CHECK();
        if (d_energy) {  // :383-384 (step 0 is zero: u0 = v0 = 0)
            rc = psfem_pcg_spmv_dot(&Mm, w.v, 0, w.tmp, d_energy + 3 * i, w.partials, w.counter + 1, stream);
            if (rc) return rc;
            rc = psfem_pcg_spmv_dot(K, w.u, 0, w.tmp, d_energy + 3 * i + 1, w.partials, w.counter + 1, stream);
            if (rc) return rc;
        }
    }
    if (d_energy) {
        energy_finish_kernel<<<(unsigned)ceil_div(n_steps, 256), 256, 0, stream>>>(n_steps, d_energy);
        PSFEM_LAUNCH_CHECK();
    }
    double hi[8];
    NmCtl hc;
    PSFEM_CUDA_CHECK(cudaMemcpyAsync(hi, w.info, sizeof(hi), cudaMemcpyDeviceToHost, stream));
    PSFEM_CUDA_CHECK(cudaMemcpyAsync(&hc, w.ctl, sizeof(hc), cudaMemcpyDeviceToHost, stream));
    PSFEM_CUDA_CHECK(cudaStreamSynchronize(stream));
    h_info[0] = hi[0] + h_acc[0]; h_info[1] = hi[1] > h_acc[1] ? hi[1] : h_acc[1]; h_info[2] = hc.triggers;
    if (hi[3] != 0.0 || h_acc[2] != 0.0) { set_error("newmark: a PCG solve did not converge (max relres %.3e)", hi[1]); return PSFEM_ENOCONV; }
    return PSFEM_OK;
}

extern "C" int psfem_newmark_run(const psfem_csr *K, const psfem_csr *M, const psfem_csr *Keff,
                                 const double *d_fshape, const double *h_fscale, int64_t n_steps, double dt, double gamma,
                                 double beta, double alpha_R, double beta_R, const double *d_a0, int64_t store_every,
                                 double *d_U, double *d_V, double *d_A, double *d_energy, double rtol, int64_t maxit,
                                 double *h_info, void *d_ws, size_t ws_bytes, void *stream) {
    LoopConsts lc;
    lc.c_u_pred = (0.5 - beta) * dt * dt; lc.c_v_pred = (1 - gamma) * dt;
    lc.c_v_corr = gamma * dt; lc.c_u_corr = beta * dt * dt;
    lc.alpha_R = alpha_R; lc.beta_R = beta_R; lc.thresh = 100.0;
    return time_loop(K, M, Keff, lc, d_fshape, h_fscale, n_steps, dt, d_a0, store_every, d_U, d_V, d_A, d_energy, rtol, maxit, h_info,
                     d_ws, ws_bytes, stream);
}

extern "C" int psfem_explicit_run(const psfem_csr *K, const psfem_csr *M, const double *d_fshape, const double *h_fscale,
                                  int64_t n_steps, double dt, const double *d_a0, int64_t store_every, double *d_U, double *d_V,
                                  double *d_A, double *d_energy, double rtol, int64_t maxit, double *h_info, void *d_ws,
                                  size_t ws_bytes, void *stream) {
    LoopConsts lc;
    lc.c_u_pred = 0.0; lc.c_v_pred = 0.5 * dt; lc.c_v_corr = 0.5 * dt; lc.c_u_corr = 0.0;
    lc.alpha_R = 0.0; lc.beta_R = 0.0; lc.thresh = 10.0;
    return time_loop(K, M, M, lc, d_fshape, h_fscale, n_steps, dt, d_a0, store_every, d_U, d_V, d_A, d_energy, rtol, maxit, h_info,
                     d_ws, ws_bytes, stream);
}
