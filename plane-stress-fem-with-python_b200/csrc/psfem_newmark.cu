// psfem_newmark.cu -- Newmark-beta time loop exactly as main.py:264-332 implements it.
//
// NOT the textbook scheme: K_eff = K + M/(beta dt^2) has no damping term (:290), the
// solve result is stored as the acceleration (:319), Rayleigh damping C = alpha_R M + beta_R K enters
// the right-hand side only (:264-266,:310), and a data-dependent rescaling "stabiliser" runs each
// step (:324-332).  Per step: predictor kernel, one fused two-matrix SpMV for the right-hand side
//   F_i - K (u_p + beta_R v_p) - alpha_R M v_p,
// a Jacobi-PCG solve warm-started from the previous acceleration (replaces lu_solve :314), corrector
// kernel with the stabiliser's reductions, store kernel.  Energies (:377-385) are two SpMV+dot.
#include "psfem_common.cuh"

namespace psfem {
int pcg_solve(const psfem_csr *A, const double *d_b, double *d_x, double rtol, int64_t maxit, int64_t check_every,
              double *h_info, double *d_info, double *h_acc, void *d_ws, size_t ws_bytes, bool have_dinv, cudaStream_t stream,
              bool masked);
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

// main.py:319-321 + the reductions of the stabiliser test :324-331
__global__ void __launch_bounds__(NT)
corrector_kernel(int64_t n, const double *__restrict__ up, const double *__restrict__ vp, const double *__restrict__ a,
                 double c_v, double c_u, double *__restrict__ u, double *__restrict__ v, double *__restrict__ partials,
                 unsigned int *counter, NmCtl *ctl, int64_t step) {
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
            if (gmax > 100.0 * mean) { scale = 10.0 * mean / gmax; ctl->triggers += 1.0; }
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

struct NmWs {
    double *u, *v, *a, *up, *vp, *w1, *rhs, *tmp, *partials, *info;
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
    w.rhs = c.take<double>(n); w.tmp = c.take<double>(n);
    int64_t np = n_blocks > 148 * 8 ? n_blocks : 148 * 8;
    w.partials = c.take<double>(2 * np);
    w.info = c.take<double>(8);
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

extern "C" int psfem_newmark_run(const psfem_csr *K, const psfem_csr *M, const psfem_csr *Keff_,
                                 const double *d_fshape, const double *h_fscale, int64_t n_steps, double dt, double gamma,
                                 double beta, double alpha_R, double beta_R, const double *d_a0, int64_t store_every,
                                 double *d_U, double *d_V, double *d_A, double *d_energy, double rtol, int64_t maxit,
                                 double *h_info, void *d_ws, size_t ws_bytes, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
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
    const double c_u_pred = (0.5 - beta) * dt * dt, c_v_pred = (1 - gamma) * dt;
    const double c_v_corr = gamma * dt, c_u_corr = beta * dt * dt;
    for (int64_t i = 1; i < n_steps; ++i) {
        predictor_kernel<<<g, NT, 0, stream>>>(n, w.u, w.v, w.a, dt, c_u_pred, c_v_pred, beta_R, w.up, w.vp, w.w1);
        PSFEM_LAUNCH_CHECK();
        int rc;   // F_i - K (u_p + beta_R v_p) - alpha_R M v_p, main.py:310
        if (two_pass) {
            rc = psfem_spmv(K, w.w1, -1.0, h_fscale[i], d_fshape, w.rhs, stream);
            if (!rc) rc = psfem_spmv(&Mm, w.vp, -alpha_R, 1.0, w.rhs, w.rhs, stream);
        } else {
            rc = psfem_spmv2(K, d_M_vals, w.w1, w.vp, -1.0, -alpha_R, h_fscale[i], d_fshape, w.rhs, stream);
        }
        if (rc) return rc;
        rc = pcg_solve(&Keff, w.rhs, w.a, rtol, maxit, 50, nullptr, w.info, h_acc, w.pcg, w.pcg_bytes, true, stream, false);  // :314
        if (rc && rc != PSFEM_ENOCONV) return rc;
        corrector_kernel<<<g, NT, 0, stream>>>(n, w.up, w.vp, w.a, c_v_corr, c_u_corr, w.u, w.v, w.partials, w.counter, w.ctl, i);
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
