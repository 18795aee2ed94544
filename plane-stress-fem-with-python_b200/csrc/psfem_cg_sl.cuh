// psfem_cg_sl.cuh -- ONE KERNEL PER PCG ITERATION on symmetric-lattice storage (included by psfem_solver.cu inside its
// anonymous namespace, after psfem_spmv_sl.cuh).
//
// The classic recurrence needs two grid-wide reductions per iteration (p.q before the residual update, r.z after it), hence
// three kernels and 344 B of HBM traffic per node: SpMV 184 + update 64 + p-update 96.  The Chronopoulos-Gear form of the
// same Krylov method (one reduction per iteration; iterates equal in exact arithmetic, same iteration counts on the plates,
// measured) removes the first boundary:
//     p = u + beta p ;  s = w + beta s ;  x += alpha p ;  r -= alpha s ;  u = D^-1 r ;  w = A u
//     gamma' = (r, u), delta = (w, u), rr = (r, r)                       <- ONE reduction
//     beta' = gamma'/gamma ;  alpha' = gamma' / (delta - beta' gamma'/alpha)
// Everything before `w = A u` is pointwise, so it rides on the marching warps of the symmetric-lattice SpMV: a warp computes
// u of a lattice column from r, w, s (and the diagonal it streams anyway: D^-1 is NOT read as a vector) when it pulls the
// column in as SpMV input, writes r, s, p, x for the nodes it owns on the way, and writes w when the row sums are done.
// x is touched every OTHER iteration only: p_{i-1} is in hand anyway when p_i = u_i + beta_i p_{i-1} is formed, so
// x += alpha_{i-1} p_{i-1} + alpha_i p_i costs one read-modify-write of x per two iterations (CgCtl::alpha_pend).
// Per node and iteration: 152 B matrix + reads r, w, s, p (64) + writes r, s, w, p (64) + x 32/2 = 296 B instead of 344, in
// one launch instead of three, with one cross-GPU scalar exchange instead of two.
// r, s, w are double buffered (a warp reads the OLD values of the halo rows / columns that other warps update).
//
// Strips of a multi-GPU partition (DIST): the first / last column segment computes u of its boundary column first and
// stores it into the neighbour's halo buffer as LL words {32 data bits | 32-bit sequence number} -- the consumer lanes
// poll exactly the words they need, no fence, no flag; the last block exchanges {gamma', delta, rr} through the peer
// boards like the classic strip solver and leaves alpha', beta' in the control block for the next launch.
#pragma once

struct CgCtl {
    double gamma, alpha, beta, rr, bb, tol2, delta;
    double alpha_pend;        // x lags by alpha_pend * p (the x update of every other iteration is folded into the next one)
    long long iters;          // -1: the next launch is a start pass, -2: an apply pass (w_out = A r_in, nothing else)
    int done, status;         // status 2: breakdown (delta - beta gamma / alpha <= 0)
    long long saved_iters;    // apply pass: iters / done of the finished solve, put back by its last block
    int saved_done, pad_;
};

struct CgVec {
    double *x, *p;                          // owned rows, updated in place
    const double *r_in, *s_in, *w_in;       // state of iteration i   (set `cur`)
    double *r_out, *s_out, *w_out;          // state of iteration i+1 (set `cur ^ 1`)
    const double *dinv;                     // DINV_VEC only: D^-1 of the owned rows, 0 at masked (constrained) rows
    // DIST only: LL halo buffers [m][4] words; mine are written by the neighbours, theirs by me
    const unsigned long long *ll_from_left, *ll_from_right;
    unsigned long long *ll_to_left, *ll_to_right;
};

__device__ __forceinline__ void cg_ll_store(unsigned long long *dst, double2 v, unsigned long long seq) {
    const unsigned long long tag = (seq & 0xffffffffull) << 32;
    const unsigned long long a = (unsigned long long)__double_as_longlong(v.x), b = (unsigned long long)__double_as_longlong(v.y);
    st_relaxed_sys(dst + 0, tag | (a & 0xffffffffull));
    st_relaxed_sys(dst + 1, tag | (a >> 32));
    st_relaxed_sys(dst + 2, tag | (b & 0xffffffffull));
    st_relaxed_sys(dst + 3, tag | (b >> 32));
}
__device__ __forceinline__ double2 cg_ll_load(const unsigned long long *src, unsigned long long seq, int *err, int site = 1) {
    const unsigned long long tag = seq & 0xffffffffull;
    unsigned long long q[4];
    long long spins = 0;
    for (;;) {
#pragma unroll
        for (int k = 0; k < 4; ++k) q[k] = ld_relaxed_sys_u64(src + k);
        if ((q[0] >> 32) == tag && (q[1] >> 32) == tag && (q[2] >> 32) == tag && (q[3] >> 32) == tag) break;
        if (++spins > PEER_SPIN_LIMIT) { *err = site | (int)((seq & 0xfffffull) << 8); break; }   // which wait gave up, at which sequence number
    }
    return make_double2(__longlong_as_double((long long)((q[0] & 0xffffffffull) | (q[1] << 32))),
                        __longlong_as_double((long long)((q[2] & 0xffffffffull) | (q[3] << 32))));
}

// ---------------------------------------------------------------------------------------------
// The operands are staged through shared memory by bulk async copies (TMA) + mbarriers.
//
// A first version that prefetched the next column into registers like spmv_sl_kernel was latency bound (ncu,
// profiles/r2d_cg_regprefetch.md: 8 warps per SM at 182 registers, 7 long-scoreboard stalls per issue, 4.2 TB/s, 1.24 ms per
// iteration at C4 -- slower than the three classic kernels).  Here a CTA = one producer warp + four consumer warps that own
// four adjacent 28-row strips of one column segment.  Per lattice column the producer issues 24 bulk copies (cp.async.bulk, completion
// on an mbarrier) of the CTA's 116-row window: the 17 off-diagonal components of column c, the two diagonal components of
// column c+1, and r, w, s, p of column c+1 -- into a 3-stage ring, two columns ahead of the consumers.  Nothing is
// prefetched in registers (94 registers instead of 182), the loads in flight no longer depend on occupancy.
// Consumer step for stage {M(c), D(c+1), V(c+1)}:  u(c+1) from V and 1/D  ->  r, s, p, x of column c+1 written;  row sums of
// column c from M(c), D(c) (carried in registers), u(c), u(c+1)  ->  w of column c written;  chain heads for column c+1.
// ---------------------------------------------------------------------------------------------
constexpr int CGT_CONSUMERS = 4;                       // consumer warps = strips per CTA
constexpr int CGT_THREADS = 32 * (CGT_CONSUMERS + 1);
constexpr int CGT_RW = SL_ROWS * CGT_CONSUMERS + 4;    // rows of the CTA window: 2 pad/halo rows on either side (116)
#ifndef PSFEM_CG_STAGES
#define PSFEM_CG_STAGES 3
#endif
constexpr int CGT_STAGES = PSFEM_CG_STAGES;
static_assert((CGT_RW * 8) % 16 == 0, "bulk copies move multiples of 16 bytes");

struct CgStage {
    double M[SL_NC][CGT_RW];      // components 1, 3..18 of column c; slots 0 and 2 hold D00 / D11 of column c+1
    double2 R[CGT_RW], W[CGT_RW], S[CGT_RW], P[CGT_RW], DI[CGT_RW];   // column c+1 (DI: DINV_VEC only)
};

template <bool DINV_VEC, bool DIST>
__global__ void __launch_bounds__(CGT_THREADS, 2)
cg_sl_tma_kernel(const double *__restrict__ sv, const SlGeom g, const CgVec v, CgCtl *__restrict__ ctl, double *__restrict__ partials,
                 unsigned int *counter, const PeerSync ps, const int n_groups, const int defer_x) {
    extern __shared__ __align__(128) unsigned char cg_smem[];
    __shared__ double red[CGT_THREADS / 32];
    __shared__ uint64_t full_bar[CGT_STAGES], empty_bar[CGT_STAGES];
    CgStage *stages = reinterpret_cast<CgStage *>(cg_smem);
    pdl_enter();
    if (ctl->done) return;
    const long long mode = ctl->iters;
    const bool apply = mode == -2;                  // y = A x for the true residual of a finished strip solve: u := r_in
    const double alpha = ctl->alpha, beta = ctl->beta, a_pend = ctl->alpha_pend;
    const bool skip_x = defer_x && a_pend == 0.0;   // this iteration leaves x alone; the next one adds both steps
    constexpr unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31, wq = threadIdx.x >> 5;
    const int gidx = blockIdx.x % n_groups, seg = blockIdx.x / n_groups;
    const int own0 = g.pre, own1 = g.S, owned_cols = own1 - own0;
    const int sa = own0 + (int)((long long)seg * owned_cols / g.G);
    const int sb = own0 + (int)((long long)(seg + 1) * owned_cols / g.G);
    const int nsteps = (sb - sa) + 2;
    const size_t mp = (size_t)g.mpad, cstride = (size_t)SL_NC * mp;
    const long long row0 = (long long)SL_ROWS * CGT_CONSUMERS * gidx - 2;   // first row of the CTA window (may be -2)
    if (threadIdx.x == 0) {
        for (int q = 0; q < CGT_STAGES; ++q) { mbar_init(&full_bar[q], 1); mbar_init(&empty_bar[q], CGT_CONSUMERS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    double dots[3] = {0.0, 0.0, 0.0};
    if (wq == CGT_CONSUMERS) {
        // ---------------- producer warp ----------------
        for (int q = 0; q < nsteps; ++q) {
            const int st = q % CGT_STAGES;
            if (q >= CGT_STAGES) mbar_wait(&empty_bar[st], ((q / CGT_STAGES) - 1) & 1);
            CgStage &S = stages[st];
            const int cM = sa + q - 2, cV = sa + q - 1;
            const bool haveM = q >= 1 && cM >= 0 && cM < g.S;
            const bool haveV = cV >= own0 && cV < own1;
            const uint32_t bM = (uint32_t)(CGT_RW * sizeof(double)), bV = (uint32_t)(CGT_RW * sizeof(double2));
            uint32_t bytes = 0;
            if (haveM) bytes += 17u * bM;
            if (haveV) bytes += 2u * bM + (DINV_VEC ? 5u : 4u) * bV;
            if (lane == 0) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the consumers' reads of this stage precede the async writes
                if (bytes) mbar_expect_tx(&full_bar[st], bytes); else mbar_arrive(&full_bar[st]);
            }
            __syncwarp();
            if (haveM && lane < SL_NC && lane != 0 && lane != 2)
                bulk_g2s(&S.M[lane][0], sv + (size_t)cM * cstride + (size_t)lane * mp + row0, bM, &full_bar[st]);
            if (haveV) {
                const long long a0 = (long long)(cV - own0) * g.m + row0;          // first node of the window in the owned-row vectors
                if (lane == 0 || lane == 2)
                    bulk_g2s(&S.M[lane][0], sv + (size_t)cV * cstride + (size_t)lane * mp + row0, bM, &full_bar[st]);
                if (lane == 19) bulk_g2s(&S.R[0], reinterpret_cast<const double2 *>(v.r_in) + a0, bV, &full_bar[st]);
                if (lane == 20) bulk_g2s(&S.W[0], reinterpret_cast<const double2 *>(v.w_in) + a0, bV, &full_bar[st]);
                if (lane == 21) bulk_g2s(&S.S[0], reinterpret_cast<const double2 *>(v.s_in) + a0, bV, &full_bar[st]);
                if (lane == 22) bulk_g2s(&S.P[0], reinterpret_cast<const double2 *>(v.p) + a0, bV, &full_bar[st]);
                if (DINV_VEC && lane == 23) bulk_g2s(&S.DI[0], reinterpret_cast<const double2 *>(v.dinv) + a0, bV, &full_bar[st]);
            }
        }
    } else {
        // ---------------- consumer warps ----------------
        const int sidx = gidx * CGT_CONSUMERS + wq;
        const int j = sidx * SL_ROWS - 2 + lane;                  // lanes 2..29 own rows, lanes 1 and 30 are the halo rows
        const int wrow = wq * SL_ROWS + lane;                     // row inside the CTA window
        const bool active = sidx < g.n_strips;
        const bool valid = active && j >= 0 && j < g.m && lane >= 1 && lane <= SL_ROWS + 2;
        const bool owned = valid && lane >= 2 && lane <= SL_ROWS + 1;
        const bool lo = valid && lane <= SL_ROWS + 1;             // lanes 1..29 hold F1, F4 (blocks towards the row above)
        const bool hi = valid && lane >= 2;                       // lanes 2..30 hold F2 (block towards the row below)
        const bool left_halo = DIST && ps.has_left && g.pre == 1;
        const bool right_halo = DIST && ps.has_right;
        int *err = ps.err;
        auto in_seg = [&](int c) { return c >= sa && c < sb; };
        auto xaddr = [&](int c) { return reinterpret_cast<double2 *>(v.x) + ((size_t)(c - own0) * g.m + j); };
        if (DIST) {
            // the last owned column reaches the right neighbour first (its first segment waits for it): computed here from
            // global memory once more, without the stores
            if (seg == g.G - 1 && ps.has_right && owned_cols > 0 && owned) {
                const int c = own1 - 1;
                const size_t a = (size_t)(c - own0) * g.m + j;
                const double2 r_ = __ldg(reinterpret_cast<const double2 *>(v.r_in) + a), w_ = __ldg(reinterpret_cast<const double2 *>(v.w_in) + a);
                const double2 s_ = __ldg(reinterpret_cast<const double2 *>(v.s_in) + a);
                double2 di;
                if (DINV_VEC) di = __ldg(reinterpret_cast<const double2 *>(v.dinv) + a);
                else { const double *pd = sv + (size_t)c * cstride + j; di = make_double2(1.0 / __ldg(pd), 1.0 / __ldg(pd + 2 * mp)); }
                double2 sn, rn;
                sn.x = w_.x + beta * s_.x; sn.y = w_.y + beta * s_.y;
                rn.x = r_.x - alpha * sn.x; rn.y = r_.y - alpha * sn.y;
                if (DINV_VEC) { rn.x *= (di.x == 0.0 ? 0.0 : 1.0); rn.y *= (di.y == 0.0 ? 0.0 : 1.0); }
                cg_ll_store(v.ll_to_right + 4 * (size_t)j, apply ? r_ : make_double2(di.x * rn.x, di.y * rn.y), ps.seq);
            }
        }
        double2 u_prev = make_double2(0.0, 0.0);      // u of column cM
        double d0 = 0.0, d2 = 0.0;                    // D00, D11 of column cM (arrived one stage earlier)
        double p0 = 0.0, p1 = 0.0;                    // chain heads for column cM
        double2 xpre = make_double2(0.0, 0.0);        // x of the vector column of this step, loaded one step ahead
        for (int q = 0; q < nsteps; ++q) {
            const int st = q % CGT_STAGES;
            const int cM = sa + q - 2, cV = sa + q - 1;
            const bool mineV = in_seg(cV);
            // x of the NEXT vector column (plain load: x is the caller's buffer, not padded for window copies)
            double2 xnext = make_double2(0.0, 0.0);
            if (!skip_x && owned && in_seg(cV + 1)) xnext = *xaddr(cV + 1);
            mbar_wait(&full_bar[st], (q / CGT_STAGES) & 1);
            const CgStage &S = stages[st];
            // ---- u of column cV (+ the pointwise updates of the nodes this lane owns) ----
            double2 un = make_double2(0.0, 0.0);
            double dn0 = 0.0, dn2 = 0.0;
            if (cV >= own0 && cV < own1) {
                if (valid) {
                    const double2 r_ = S.R[wrow], w_ = S.W[wrow], s_ = S.S[wrow];
                    dn0 = S.M[0][wrow]; dn2 = S.M[2][wrow];
                    double2 di;
                    if (DINV_VEC) di = S.DI[wrow];
                    else di = make_double2(1.0 / dn0, 1.0 / dn2);
                    double2 sn, rn;
                    sn.x = w_.x + beta * s_.x; sn.y = w_.y + beta * s_.y;
                    rn.x = r_.x - alpha * sn.x; rn.y = r_.y - alpha * sn.y;
                    if (DINV_VEC) { rn.x *= (di.x == 0.0 ? 0.0 : 1.0); rn.y *= (di.y == 0.0 ? 0.0 : 1.0); }
                    un.x = di.x * rn.x; un.y = di.y * rn.y;
                    if (apply) un = r_;
                    if (!apply && mineV && owned) {
                        const size_t a = (size_t)(cV - own0) * g.m + j;
                        const double2 p_ = S.P[wrow];
                        double2 pn;
                        pn.x = di.x * r_.x + beta * p_.x; pn.y = di.y * r_.y + beta * p_.y;      // p_i = u_i + beta_i p_{i-1}
                        reinterpret_cast<double2 *>(v.p)[a] = pn;
                        if (!skip_x) {   // x += alpha_{i-1} p_{i-1} (pending from the previous iteration, or 0) + alpha_i p_i
                            double2 xn_;
                            xn_.x = (xpre.x + a_pend * p_.x) + alpha * pn.x; xn_.y = (xpre.y + a_pend * p_.y) + alpha * pn.y;
                            reinterpret_cast<double2 *>(v.x)[a] = xn_;
                        }
                        reinterpret_cast<double2 *>(v.r_out)[a] = rn;
                        reinterpret_cast<double2 *>(v.s_out)[a] = sn;
                        dots[0] += rn.x * un.x; dots[0] += rn.y * un.y;
                        dots[2] += rn.x * rn.x; dots[2] += rn.y * rn.y;
                    }
                    if (DIST && mineV && owned && cV == own0 && ps.has_left) cg_ll_store(v.ll_to_left + 4 * (size_t)j, un, ps.seq);
                }
            } else if (DIST && valid) {
                if (cV == own0 - 1 && left_halo) un = cg_ll_load(v.ll_from_left + 4 * (size_t)j, ps.seq, err);
                else if (cV == own1 && right_halo) un = cg_ll_load(v.ll_from_right + 4 * (size_t)j, ps.seq, err, 2);
            }
            // ---- column cM: row sums (q >= 2) and chain heads for the next column (q >= 1) ----
            if (q >= 1 && cM >= 0 && cM < g.S) {
                auto F = [&](int k, bool on) -> double { return on ? S.M[k][wrow] : 0.0; };
                const double2 xc = u_prev, xn = un;
                if (q >= 2) {
                    const double f3 = F(3, lo), f4 = F(4, lo), f5 = F(5, lo), f6 = F(6, lo);
                    double s0 = p0, s1 = p1;
                    s0 += f3 * xc.x; s0 += f5 * xc.y;
                    s1 += f4 * xc.x; s1 += f6 * xc.y;
                    s0 = __shfl_up_sync(FULL, s0, 1); s1 = __shfl_up_sync(FULL, s1, 1);
                    double2 xr, xnl, xnr;
                    xr.x = __shfl_down_sync(FULL, xc.x, 1); xr.y = __shfl_down_sync(FULL, xc.y, 1);
                    xnl.x = __shfl_up_sync(FULL, xn.x, 1); xnl.y = __shfl_up_sync(FULL, xn.y, 1);
                    xnr.x = __shfl_down_sync(FULL, xn.x, 1); xnr.y = __shfl_down_sync(FULL, xn.y, 1);
                    const double f0 = owned ? d0 : 0.0, f1 = F(1, owned), f2 = owned ? d2 : 0.0;
                    s0 += f0 * xc.x; s0 += f1 * xc.y;
                    s1 += f1 * xc.x; s1 += f2 * xc.y;
                    s0 += f3 * xr.x; s0 += f4 * xr.y;
                    s1 += f5 * xr.x; s1 += f6 * xr.y;
                    s0 += F(7, hi) * xnl.x; s0 += F(8, hi) * xnl.y;
                    s1 += F(9, hi) * xnl.x; s1 += F(10, hi) * xnl.y;
                    s0 += F(11, owned) * xn.x; s0 += F(12, owned) * xn.y;
                    s1 += F(13, owned) * xn.x; s1 += F(14, owned) * xn.y;
                    s0 += F(15, lo) * xnr.x; s0 += F(16, lo) * xnr.y;
                    s1 += F(17, lo) * xnr.x; s1 += F(18, lo) * xnr.y;
                    if (owned) {
                        const size_t a = (size_t)(cM - own0) * g.m + j;
                        reinterpret_cast<double2 *>(v.w_out)[a] = make_double2(s0, s1);
                        dots[1] += xc.x * s0; dots[1] += xc.y * s1;
                    }
                }
                if (q + 1 < nsteps) {   // chain heads of column cM + 1 (phaseP)
                    double c0 = 0.0, c1 = 0.0;
                    c0 += F(15, lo) * xc.x; c0 += F(17, lo) * xc.y;
                    c1 += F(16, lo) * xc.x; c1 += F(18, lo) * xc.y;
                    c0 = __shfl_up_sync(FULL, c0, 1); c1 = __shfl_up_sync(FULL, c1, 1);
                    c0 += F(11, owned) * xc.x; c0 += F(13, owned) * xc.y;
                    c1 += F(12, owned) * xc.x; c1 += F(14, owned) * xc.y;
                    c0 = __shfl_up_sync(FULL, c0, 1); c1 = __shfl_up_sync(FULL, c1, 1);
                    c0 += F(7, hi) * xc.x; c0 += F(9, hi) * xc.y;
                    c1 += F(8, hi) * xc.x; c1 += F(10, hi) * xc.y;
                    p0 = __shfl_down_sync(FULL, c0, 2); p1 = __shfl_down_sync(FULL, c1, 2);
                }
            } else {
                p0 = 0.0; p1 = 0.0;
            }
            u_prev = un; d0 = dn0; d2 = dn2; xpre = xnext;
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[st]);
        }
    }
    double total[3];
    if (grid_sum<CGT_THREADS, 3>(dots, partials, counter, red, total)) {
        if (DIST) {
            peer_publish<3>(ps, 1, total);
            double gsum[3];
            peer_wait_sum<3>(ps, 1, ps.seq, gsum);
            total[0] = gsum[0]; total[1] = gsum[1]; total[2] = gsum[2];
        }
        const double gamma_new = total[0], delta = total[1];
        if (apply) {
            // APPLY PASS (psfem_dist_cg_apply): w_out = A r_in is in place; the control block of the finished solve comes back
            ctl->iters = ctl->saved_iters; ctl->done = ctl->saved_done;
        } else if (mode < 0) {
            // START PASS (psfem_dist_cg_start): launched with r = b, w = s = p = x = 0, alpha = beta = 0, the kernel has
            // just produced u_0 = D^-1 b (left in p), w_0 = A u_0 (halo exchange included), gamma_0, delta_0 and b.b
            ctl->gamma = gamma_new; ctl->delta = delta; ctl->rr = total[2]; ctl->bb = total[2];
            ctl->tol2 = ctl->tol2 * total[2];          // rtol^2 was parked here
            ctl->beta = 0.0;
            ctl->alpha = gamma_new / delta;
            ctl->alpha_pend = 0.0;
            ctl->iters = 0;
            ctl->status = 0;
            ctl->done = (total[2] == 0.0 || total[2] <= ctl->tol2) ? 1 : 0;
            if (!ctl->done && !(delta > 0.0)) { ctl->done = 1; ctl->status = 2; }
        } else {
            const double beta_new = gamma_new / ctl->gamma;
            const double den = delta - beta_new * gamma_new / alpha;
            ctl->gamma = gamma_new;
            ctl->delta = delta;
            ctl->rr = total[2];
            ctl->iters += 1;
            ctl->beta = beta_new;
            ctl->alpha = gamma_new / den;
            ctl->alpha_pend = skip_x ? alpha : 0.0;
            if (total[2] <= ctl->tol2) ctl->done = 1;
            else if (!(den > 0.0) || !(gamma_new > 0.0)) { ctl->done = 1; ctl->status = 2; }
        }
        __threadfence();
    }
}

// control block of an apply pass: the kernel must run although the solve is done, and must leave the solve's state readable
__global__ void cg_apply_ctl_kernel(CgCtl *ctl) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    ctl->saved_iters = ctl->iters; ctl->saved_done = ctl->done;
    ctl->alpha = 0.0; ctl->beta = 0.0; ctl->alpha_pend = 0.0;
    ctl->iters = -2;
    ctl->done = 0;
}

// the pending half of the deferred x update: x += alpha_pend * p (then alpha_pend = 0); run before x is read
__global__ void cg_finish_x_kernel(int64_t n, double *__restrict__ x, const double *__restrict__ p, const CgCtl *__restrict__ ctl) {
    const double a = ctl->alpha_pend;
    if (a == 0.0) return;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) x[i] += a * p[i];
}
__global__ void cg_clear_pend_kernel(CgCtl *ctl) { ctl->alpha_pend = 0.0; }

// start of a solve: gamma_0, delta_0 are in place (pcg_init_kernel: p = u_0 = D^-1 r_0, {r.u, r.r, b.b}; the SpMV: w_0 = A u_0,
// delta_0 = u_0 . w_0); alpha_0 = gamma_0 / delta_0, beta_0 = 0
__global__ void cg_start_ctl_kernel(CgCtl *ctl, const double *init3 /* r.u, r.r, b.b */, const double *delta, double rtol) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const double gamma = init3[0], rr = init3[1], bb = init3[2], d = *delta;
    ctl->gamma = gamma; ctl->delta = d; ctl->rr = rr; ctl->bb = bb;
    ctl->tol2 = rtol * rtol * bb;
    ctl->beta = 0.0;
    ctl->alpha = gamma / d;
    ctl->alpha_pend = 0.0;
    ctl->iters = 0;
    ctl->status = 0;
    ctl->done = (bb == 0.0 || rr <= ctl->tol2) ? 1 : 0;
    if (!ctl->done && !(d > 0.0)) { ctl->done = 1; ctl->status = 2; }
}

// control block of a start pass (see the last block of cg_sl_tma_kernel)
__global__ void cg_start_pass_ctl_kernel(CgCtl *ctl, double rtol) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    ctl->gamma = 1.0; ctl->delta = 0.0; ctl->rr = 0.0; ctl->bb = 0.0;
    ctl->tol2 = rtol * rtol;
    ctl->alpha = 0.0; ctl->beta = 0.0; ctl->alpha_pend = 0.0;
    ctl->iters = -1;
    ctl->done = 0; ctl->status = 0;
}

// D^-1 of the owned rows from the streamed diagonal components (0 and 2) of the lattice storage: 16 B per node instead of a
// search through the CSR rows (psfem_extract_diag)
__global__ void cg_sl_dinv_kernel(const double *__restrict__ sv, const SlGeom g, double *__restrict__ dinv) {
    const int64_t nodes = (int64_t)(g.S - g.pre) * g.m;
    const size_t mp = (size_t)g.mpad, cstride = (size_t)SL_NC * mp;
    for (int64_t a = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; a < nodes; a += (int64_t)gridDim.x * blockDim.x) {
        const int64_t c = a / g.m + g.pre, j = a % g.m;
        const double *pd = sv + (size_t)c * cstride + j;
        reinterpret_cast<double2 *>(dinv)[a] = make_double2(1.0 / pd[0], 1.0 / pd[2 * mp]);
    }
}

// ---------------------------------------------------------------------------------------------
// The plain SpMV (+ dot) on symmetric-lattice storage with the same TMA + mbarrier staging: y = alpha A x + gamma z.
// Same arithmetic, same order, same bits as spmv_sl_kernel (and hence as the CSR kernels); the matrix goes through a 4-stage
// shared ring fed by the producer warp instead of a register prefetch (154 registers, 12 warps per SM, 6.2 TB/s there).
// x is the caller's vector (possibly extended by halo columns): plain loads, one column ahead.  Single-GPU operands only:
// the strip solver's classic kernels keep spmv_sl_kernel, which carries the peer waits.
// ---------------------------------------------------------------------------------------------
#ifndef PSFEM_SPMV_TMA_STAGES
#define PSFEM_SPMV_TMA_STAGES 4
#endif
constexpr int SPT_STAGES = PSFEM_SPMV_TMA_STAGES;
struct SpStage { double M[SL_NC][CGT_RW]; };

template <bool DOT>
__global__ void __launch_bounds__(CGT_THREADS, 2)
spmv_sl_tma_kernel(const double *__restrict__ sv, const SlGeom g, const double *__restrict__ x, double alpha, double gamma,
                   const double *__restrict__ z, double *__restrict__ y, double *__restrict__ dot_out, double *__restrict__ partials,
                   unsigned int *counter, const int *__restrict__ done, const int n_groups) {
    extern __shared__ __align__(128) unsigned char cg_smem[];
    __shared__ double red[CGT_THREADS / 32];
    __shared__ uint64_t full_bar[SPT_STAGES], empty_bar[SPT_STAGES];
    SpStage *stages = reinterpret_cast<SpStage *>(cg_smem);
    pdl_enter();
    if (done && *done) return;
    constexpr unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31, wq = threadIdx.x >> 5;
    const int gidx = blockIdx.x % n_groups, seg = blockIdx.x / n_groups;
    const int owned_cols = g.S - g.pre;
    const int sa = g.pre + (int)((long long)seg * owned_cols / g.G);
    const int sb = g.pre + (int)((long long)(seg + 1) * owned_cols / g.G);
    const int q0 = sa > 0 ? 0 : 1;                     // step 0 = the column before the segment (chain heads only)
    const int nsteps = (sb - sa) + 1;
    const size_t mp = (size_t)g.mpad, cstride = (size_t)SL_NC * mp;
    const long long row0 = (long long)SL_ROWS * CGT_CONSUMERS * gidx - 2;
    if (threadIdx.x == 0) {
        for (int q = 0; q < SPT_STAGES; ++q) { mbar_init(&full_bar[q], 1); mbar_init(&empty_bar[q], CGT_CONSUMERS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    double dot[1] = {0.0};
    if (wq == CGT_CONSUMERS) {
        for (int q = q0; q < nsteps; ++q) {
            const int it = q - q0, st = it % SPT_STAGES;
            if (it >= SPT_STAGES) mbar_wait(&empty_bar[st], ((it / SPT_STAGES) - 1) & 1);
            const int cM = sa - 1 + q;
            const uint32_t bM = (uint32_t)(CGT_RW * sizeof(double));
            if (lane == 0) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_expect_tx(&full_bar[st], (uint32_t)SL_NC * bM);
            }
            __syncwarp();
            if (lane < SL_NC) bulk_g2s(&stages[st].M[lane][0], sv + (size_t)cM * cstride + (size_t)lane * mp + row0, bM, &full_bar[st]);
        }
    } else {
        const int sidx = gidx * CGT_CONSUMERS + wq;
        const int j = sidx * SL_ROWS - 2 + lane;
        const int wrow = wq * SL_ROWS + lane;
        const bool active = sidx < g.n_strips;
        const bool valid = active && j >= 0 && j < g.m && lane >= 1 && lane <= SL_ROWS + 2;
        const bool owned = valid && lane >= 2 && lane <= SL_ROWS + 1;
        const bool lo = valid && lane <= SL_ROWS + 1;
        const bool hi = valid && lane >= 2;
        const double2 *__restrict__ x2 = reinterpret_cast<const double2 *>(x);
        auto xload = [&](int s) -> double2 {
            const int xc = s + g.xcol0;
            if (valid && xc >= 0 && xc < g.xcols) return __ldcg(x2 + (size_t)xc * g.m + j);
            return make_double2(0.0, 0.0);
        };
        double p0 = 0.0, p1 = 0.0;
        double2 xc = xload(sa - 1 + q0), xn = xload(sa + q0);
        for (int q = q0; q < nsteps; ++q) {
            const int it = q - q0, st = it % SPT_STAGES;
            const int cM = sa - 1 + q;
            const double2 xnn = xload(cM + 2);
            mbar_wait(&full_bar[st], (it / SPT_STAGES) & 1);
            const SpStage &S = stages[st];
            auto F = [&](int k, bool on) -> double { return on ? S.M[k][wrow] : 0.0; };
            if (q >= 1) {
                const double f3 = F(3, lo), f4 = F(4, lo), f5 = F(5, lo), f6 = F(6, lo);
                double s0 = p0, s1 = p1;
                s0 += alpha * f3 * xc.x; s0 += alpha * f5 * xc.y;
                s1 += alpha * f4 * xc.x; s1 += alpha * f6 * xc.y;
                s0 = __shfl_up_sync(FULL, s0, 1); s1 = __shfl_up_sync(FULL, s1, 1);
                double2 xr, xnl, xnr;
                xr.x = __shfl_down_sync(FULL, xc.x, 1); xr.y = __shfl_down_sync(FULL, xc.y, 1);
                xnl.x = __shfl_up_sync(FULL, xn.x, 1); xnl.y = __shfl_up_sync(FULL, xn.y, 1);
                xnr.x = __shfl_down_sync(FULL, xn.x, 1); xnr.y = __shfl_down_sync(FULL, xn.y, 1);
                const double f0 = F(0, owned), f1 = F(1, owned), f2 = F(2, owned);
                s0 += alpha * f0 * xc.x; s0 += alpha * f1 * xc.y;
                s1 += alpha * f1 * xc.x; s1 += alpha * f2 * xc.y;
                s0 += alpha * f3 * xr.x; s0 += alpha * f4 * xr.y;
                s1 += alpha * f5 * xr.x; s1 += alpha * f6 * xr.y;
                s0 += alpha * F(7, hi) * xnl.x; s0 += alpha * F(8, hi) * xnl.y;
                s1 += alpha * F(9, hi) * xnl.x; s1 += alpha * F(10, hi) * xnl.y;
                s0 += alpha * F(11, owned) * xn.x; s0 += alpha * F(12, owned) * xn.y;
                s1 += alpha * F(13, owned) * xn.x; s1 += alpha * F(14, owned) * xn.y;
                s0 += alpha * F(15, lo) * xnr.x; s0 += alpha * F(16, lo) * xnr.y;
                s1 += alpha * F(17, lo) * xnr.x; s1 += alpha * F(18, lo) * xnr.y;
                if (owned) {
                    const size_t a = (size_t)(cM - g.pre) * g.m + j;
                    if (z) {
                        const double2 zz = reinterpret_cast<const double2 *>(z)[a];
                        s0 += gamma * zz.x; s1 += gamma * zz.y;
                    }
                    reinterpret_cast<double2 *>(y)[a] = make_double2(s0, s1);
                    if (DOT) { dot[0] += xc.x * s0; dot[0] += xc.y * s1; }
                }
            }
            if (q + 1 < nsteps) {   // chain heads of column cM + 1
                double c0 = 0.0, c1 = 0.0;
                c0 += alpha * F(15, lo) * xc.x; c0 += alpha * F(17, lo) * xc.y;
                c1 += alpha * F(16, lo) * xc.x; c1 += alpha * F(18, lo) * xc.y;
                c0 = __shfl_up_sync(FULL, c0, 1); c1 = __shfl_up_sync(FULL, c1, 1);
                c0 += alpha * F(11, owned) * xc.x; c0 += alpha * F(13, owned) * xc.y;
                c1 += alpha * F(12, owned) * xc.x; c1 += alpha * F(14, owned) * xc.y;
                c0 = __shfl_up_sync(FULL, c0, 1); c1 = __shfl_up_sync(FULL, c1, 1);
                c0 += alpha * F(7, hi) * xc.x; c0 += alpha * F(9, hi) * xc.y;
                c1 += alpha * F(8, hi) * xc.x; c1 += alpha * F(10, hi) * xc.y;
                p0 = __shfl_down_sync(FULL, c0, 2); p1 = __shfl_down_sync(FULL, c1, 2);
            }
            xc = xn; xn = xnn;
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[st]);
        }
    }
    if (DOT) {
        double total[1];
        if (grid_sum<CGT_THREADS, 1>(dot, partials, counter, red, total)) *dot_out = total[0];
    }
}
