// psfem_solver.cu -- HBM-bound sparse kernels of the static / Newmark / modal drivers.
//
//   spmv_sl_kernel   (psfem_spmv_sl.cuh) SpMV (+ fused dot) on symmetric-lattice storage: the matrices of Polygones.mesh
//                    meshes as 19 doubles per node, marching warps, transposed products as shuffle chains -- half the
//                    bytes of the node-block kernel, the same bits; the kernel PCG runs on whenever the operand allows
//   spmv_nb_kernel   SpMV (+ fused dot) on the 2x2 node-block structure of FEM matrices: persistent, one producer warp
//                    streams 8 B values + one 4-byte node id per block through a TMA-fed shared ring, one DOF row per
//                    consumer thread, x gathered as double2 -- 9 B per non-zero, the kernel PCG runs on
//   spmv2_kernel     the same pipeline on plain CSR (12 B per non-zero) for matrices without that structure
//   spmv_kernel      v1: one CTA per row block (small matrices, two-matrix right-hand side of Newmark without node blocks)
//   pcg_*_kernel     the vector kernels of a Jacobi-PCG iteration (r -= alpha q with the reductions; x += alpha p,
//                    p = D^-1 r + beta p), scalars on the device, deterministic last-block sums, masked variant for
//                    solves on the unreduced matrix
//   dist_*_kernel    the same iteration on a strip of a multi-GPU partition with the scalar all-reduces and the halo
//                    exchange fused in (peer boards over CUDA-IPC memory, see PeerBoard)
//   pcg_small_kernel whole solve in one single-CTA launch for n <= 16384
//   BLAS-1 helpers for the Lanczos driver.
// Every row is summed in column order (scipy's CSR order): spmv_sl_kernel, spmv_nb_kernel and spmv2_kernel give
// bit-identical results (fused multiply-adds in that order; the one-CTA-per-block spmv_kernel rounds its products first).
// The PCG kernels are chained with programmatic dependent launch (pdl_enter / launch_pdl).
//
// Replaces: np.linalg.solve / scipy.linalg.solve / lu_factor+lu_solve (Statictest.py:69,
// StaticTest2.py:71, Beamexemple.py:63, main.py:286,295,314) and the dense mat-vecs of
// main.py:310,383-384.
#include <cstdlib>
#include <cstring>
#include <type_traits>
#include <utility>
#include <vector>
#include "psfem_common.cuh"

using namespace psfem;

namespace {

constexpr int SPMV_THREADS = 256;
constexpr int SPMV_T = 4096;  // target non-zeros per row block (32 KB of fp64 products)
constexpr int EW_THREADS = 256;

// Programmatic dependent launch: the kernels of a PCG iteration are launched with
// cudaLaunchAttributeProgrammaticStreamSerialization, raise griddepcontrol.launch_dependents on entry and wait for
// their predecessor with griddepcontrol.wait before touching anything it wrote.  The launch of kernel k+1 (command
// processing, CTA scheduling, prologue) then overlaps the tail of kernel k -- the last block's fixed-order reduction
// and, across GPUs, its peer exchange -- instead of following it (about 4 us per kernel boundary).  Without the launch
// attribute both instructions are no-ops.
__device__ __forceinline__ void pdl_enter() {
    asm volatile("griddepcontrol.launch_dependents;");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}
template <typename... KArgs, typename... Args>
cudaError_t launch_pdl(void (*kern)(KArgs...), unsigned grid, unsigned block, size_t smem, cudaStream_t stream, Args &&...args) {
    static const bool off = getenv("PSFEM_NO_PDL") != nullptr;   // A/B switch
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = off ? 0 : 1;
    return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}

int ew_grid(int64_t n) {
    int64_t g = ceil_div(n, EW_THREADS * 4);
    int64_t cap = 148 * 8;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

// ---------------------------------------------------------------------------------------------
// row-block plan
// ---------------------------------------------------------------------------------------------
__global__ void spmv_plan_kernel(const int32_t *__restrict__ row_ptr, int64_t n, int64_t n_blocks, int32_t *__restrict__ rowblk) {
    int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (b > n_blocks) return;
    if (b == n_blocks) { rowblk[b] = (int32_t)n; return; }
    if (b == 0) { rowblk[0] = 0; return; }
    const int64_t target = (int64_t)row_ptr[0] + b * SPMV_T;
    int64_t lo = 0, hi = n;
    while (lo < hi) {  // first row whose start is >= target
        int64_t mid = lo + ((hi - lo) >> 1);
        if (row_ptr[mid] < target) lo = mid + 1; else hi = mid;
    }
    rowblk[b] = (int32_t)lo;
}

__global__ void max_row_kernel(const int32_t *__restrict__ row_ptr, int64_t n, int *__restrict__ out) {
    int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int len = 0;
    if (r < n) len = row_ptr[r + 1] - row_ptr[r];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) len = max(len, __shfl_xor_sync(0xffffffffu, len, o));
    if ((threadIdx.x & 31) == 0 && len > 0) atomicMax(out, len);
}

// ---------------------------------------------------------------------------------------------
// deterministic grid-wide sums: per-block partials, the last block to finish adds them in a fixed
// order (no floating-point atomics).
// ---------------------------------------------------------------------------------------------
template <int THREADS, int NV, bool SYS_FENCE = false>
__device__ __forceinline__ bool grid_sum(double (&v)[NV], double *__restrict__ partials, unsigned int *counter,
                                         double *smem, double (&total)[NV]) {
    __shared__ bool is_last;
    const int nb = gridDim.x;
#pragma unroll
    for (int q = 0; q < NV; ++q) {
        double s = block_sum<THREADS>(v[q], smem);
        if (threadIdx.x == 0) partials[(int64_t)q * nb + blockIdx.x] = s;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (SYS_FENCE) __threadfence_system(); else __threadfence();
        unsigned int ticket = atomicAdd(counter, 1u);
        is_last = (ticket == (unsigned)nb - 1);
    }
    __syncthreads();
    if (!is_last) return false;
    __threadfence();
#pragma unroll
    for (int q = 0; q < NV; ++q) {
        double s = 0;
        for (int i = threadIdx.x; i < nb; i += THREADS) s += __ldcg(partials + (int64_t)q * nb + i);
        total[q] = block_sum<THREADS>(s, smem);
    }
    if (threadIdx.x == 0) *counter = 0;
    return threadIdx.x == 0;  // totals valid in thread 0 of the last block
}

// device-side control block of one PCG solve
struct PcgCtl {
    double rz[2];
    double pq;
    double rr;
    double bb;
    double tol2;
    long long iters;
    int done;
    int parity;
};

// ---------------------------------------------------------------------------------------------
// peer boards: the collectives of the strip-partitioned PCG, fused into its three kernels.
//
// Every rank owns a PeerBoard in memory that all ranks map (CUDA IPC over NVLink/NVSwitch).  A reduction is
// "everybody stores its partial into everybody's board, then a sequence number"; the LAST block of the producing
// kernel then spins on ITS OWN board (one thread, ld.acquire.sys on local HBM), adds the partials in rank order --
// every rank gets the bit-identical sum -- and leaves the result in local memory for the next kernel, so no
// collective is launched and no consumer block ever polls.  The halo of p is stored straight into the
// neighbours' extended p vectors by the kernel that computes p.  Sequence numbers only grow (one per PCG
// iteration / solve start).
// Scalars travel as 8-byte words {32 data bits | 32-bit sequence number} (two per double), the way NCCL's LL protocol
// packs a flag with its data: an 8-byte store is single-copy atomic, so a word that carries the expected sequence
// number also carries its data -- no system-scope fence between "data" and "flag", one NVLink store latency per
// reduction instead of store + fence round trip + store.  A rank cannot publish sequence s+1 of a stage before every
// rank has consumed s (it first needs everybody's partial of the stage in between), so one buffer per stage suffices.
// ---------------------------------------------------------------------------------------------
#ifndef PSFEM_PEER_LL
#define PSFEM_PEER_LL 1     // 0: values, one system-scope fence, then sequence numbers (the protocol this replaced; A/B)
#endif
struct PeerBoard {
    unsigned long long ll[2][PSFEM_MAX_RANKS][8];   // [0: p.q | 1: {r.z, r.r, b.b}][source rank][2 words per double]
    unsigned long long halo_flag[2];                 // [0] left / [1] right neighbour's halo for iteration seq is in
    double val[2][2][PSFEM_MAX_RANKS][4];           // PSFEM_PEER_LL=0: [stage][seq parity][source rank][slot]
    unsigned long long flag[2][PSFEM_MAX_RANKS];    // PSFEM_PEER_LL=0: seq of the value last published by a rank
};
struct PeerSync {
    PeerBoard *board[PSFEM_MAX_RANKS];               // board[w] of rank w (peer mapped), board[rank] is local
    int world, rank, has_left, has_right;
    int halo_in_spmv;                                // the SpMV waits for the halo itself (edge segments of spmv_sl_kernel only)
    unsigned long long seq;
    int *err;                                        // local flag: set when a wait gave up (a peer died), never hangs the GPU
    double *scal;                                    // local scalars: [0] r.z [1] r.r [2] b.b [3] scratch [4] p.q [5+parity] r.z by sequence parity
};
constexpr long long PEER_SPIN_LIMIT = 1ll << 26;     // polls of local memory (~ tens of seconds) before giving up

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
// flag stores are RELAXED: the publisher issues ONE system-scope fence after its data stores and before all of
// its flag stores (fence + relaxed store = release); a st.release per flag would drain the store queues per flag
__device__ __forceinline__ void st_relaxed_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
// one thread: my partial(s), packed with the sequence number, into every rank's board (no fence: see PeerBoard)
template <int NV> __device__ __forceinline__ void peer_publish(const PeerSync &ps, int stage, const double (&v)[NV]) {
    static_assert(NV <= 4, "four doubles per stage");
#if !PSFEM_PEER_LL
    const int par = (int)(ps.seq & 1);
    for (int w = 0; w < ps.world; ++w) {
#pragma unroll
        for (int q = 0; q < NV; ++q) ps.board[w]->val[stage][par][ps.rank][q] = v[q];
    }
    __threadfence_system();
    for (int w = 0; w < ps.world; ++w) st_relaxed_sys(&ps.board[w]->flag[stage][ps.rank], ps.seq);
    return;
#endif
    const unsigned long long tag = (ps.seq & 0xffffffffull) << 32;
    for (int w = 0; w < ps.world; ++w) {
        unsigned long long *dst = ps.board[w]->ll[stage][ps.rank];
#pragma unroll
        for (int q = 0; q < NV; ++q) {
            const unsigned long long bits = (unsigned long long)__double_as_longlong(v[q]);
            st_relaxed_sys(dst + 2 * q, tag | (bits & 0xffffffffull));
            st_relaxed_sys(dst + 2 * q + 1, tag | (bits >> 32));
        }
    }
}
// one thread: wait until every rank's partial of sequence `seq` is on my board, add them in rank order
template <int NV> __device__ __forceinline__ void peer_wait_sum(const PeerSync &ps, int stage, unsigned long long seq, double (&out)[NV]) {
    PeerBoard *me = ps.board[ps.rank];
    long long spins = 0;
#if !PSFEM_PEER_LL
    for (int w = 0; w < ps.world; ++w)
        while (ld_acquire_sys(&me->flag[stage][w]) < seq)
            if (++spins > PEER_SPIN_LIMIT) { *ps.err = 1; break; }
#pragma unroll
    for (int q = 0; q < NV; ++q) {
        double t = 0.0;
        for (int w = 0; w < ps.world; ++w) t += __longlong_as_double((long long)ld_relaxed_sys_u64(reinterpret_cast<const unsigned long long *>(&me->val[stage][(int)(seq & 1)][w][q])));
        out[q] = t;
    }
    return;
#endif
    const unsigned long long tag = seq & 0xffffffffull;
    double s[NV];
#pragma unroll
    for (int q = 0; q < NV; ++q) s[q] = 0.0;
    for (int w = 0; w < ps.world; ++w) {
#pragma unroll
        for (int q = 0; q < NV; ++q) {
            unsigned long long a, b;
            for (;;) {
                a = ld_relaxed_sys_u64(&me->ll[stage][w][2 * q]);
                b = ld_relaxed_sys_u64(&me->ll[stage][w][2 * q + 1]);
                if ((a >> 32) == tag && (b >> 32) == tag) break;
                if (++spins > PEER_SPIN_LIMIT) { *ps.err = 3 | (w << 4) | (int)((seq & 0xfffffull) << 8); break; }
            }
            s[q] += __longlong_as_double((long long)((a & 0xffffffffull) | (b << 32)));
        }
    }
#pragma unroll
    for (int q = 0; q < NV; ++q) out[q] = s[q];
}
__device__ __forceinline__ void peer_wait_halo(const PeerSync &ps, bool left = true, bool right = true) {   // one thread
    PeerBoard *me = ps.board[ps.rank];
    long long spins = 0;
    if (left && ps.has_left) while (ld_acquire_sys(&me->halo_flag[0]) < ps.seq) if (++spins > PEER_SPIN_LIMIT) { *ps.err = 1; break; }
    if (right && ps.has_right) while (ld_acquire_sys(&me->halo_flag[1]) < ps.seq) if (++spins > PEER_SPIN_LIMIT) { *ps.err = 1; break; }
}

// ---------------------------------------------------------------------------------------------
// SpMV
// ---------------------------------------------------------------------------------------------
// y[r] = sum_k (alpha*v1[k]*x1[c[k]] (+ beta*v2[k]*x2[c[k]])) + gamma*z[r];   DOT: dot_out = sum_r xdot[r]*y[r]
template <bool DUAL, bool DOT>
__global__ void __launch_bounds__(SPMV_THREADS)
spmv_kernel(const int32_t *__restrict__ row_ptr, const int32_t *__restrict__ col_idx, const double *__restrict__ v1,
            const double *__restrict__ v2, const int32_t *__restrict__ rowblk, const double *__restrict__ x1,
            const double *__restrict__ x2, double alpha, double beta, double gamma, const double *__restrict__ z,
            double *__restrict__ y, const double *__restrict__ xdot, double *__restrict__ dot_out,
            double *__restrict__ partials, unsigned int *counter, const int *__restrict__ done) {
    extern __shared__ double prod[];
    __shared__ double red[SPMV_THREADS / 32];
    if (done && *done) return;
    const int t = threadIdx.x;
    const int r0 = rowblk[blockIdx.x], r1 = rowblk[blockIdx.x + 1];
    const int k0 = row_ptr[r0], k1 = row_ptr[r1];
    // phase 1: coalesced stream of the block's non-zeros, 4 independent loads in flight per thread
    int k = k0 + t;
    for (; k + 3 * SPMV_THREADS < k1; k += 4 * SPMV_THREADS) {
        int c[4];
        double a[4], b[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            c[u] = __ldg(col_idx + k + u * SPMV_THREADS);
            a[u] = __ldg(v1 + k + u * SPMV_THREADS);
            if (DUAL) b[u] = __ldg(v2 + k + u * SPMV_THREADS);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            double p = alpha * a[u] * x1[c[u]];
            if (DUAL) p += beta * b[u] * x2[c[u]];
            prod[k - k0 + u * SPMV_THREADS] = p;
        }
    }
    for (; k < k1; k += SPMV_THREADS) {
        const int c = __ldg(col_idx + k);
        double p = alpha * __ldg(v1 + k) * x1[c];
        if (DUAL) p += beta * __ldg(v2 + k) * x2[c];
        prod[k - k0] = p;
    }
    __syncthreads();
    // phase 2: one thread per row, products summed in column order
    double dot[1] = {0.0};
    for (int r = r0 + t; r < r1; r += SPMV_THREADS) {
        const int a = row_ptr[r] - k0, b = row_ptr[r + 1] - k0;
        double s = 0.0;
        for (int q = a; q < b; ++q) s += prod[q];
        if (z) s += gamma * z[r];
        y[r] = s;
        if (DOT) dot[0] += xdot[r] * s;
    }
    if (DOT) {
        double total[1];
        if (grid_sum<SPMV_THREADS, 1>(dot, partials, counter, red, total)) *dot_out = total[0];
    }
}

// ---------------------------------------------------------------------------------------------
// SpMV v2: persistent, warp-specialised, TMA-staged.
//
// One producer warp streams the vals / col_idx / row_ptr slices of the CTA's row blocks into a ring
// of shared-memory stages with bulk async copies (cp.async.bulk + mbarrier complete_tx), running
// SPMV2_STAGES blocks ahead of the eight consumer warps, so HBM latency is off the critical path
// (v1 was long-scoreboard bound at 41 % of DRAM peak, profiles/r1a_summary.md).  Consumers turn a
// landed stage into products in place (x gathered through L1/L2), then reduce one row per thread in
// column order.  Row blocks are dealt round-robin to a grid of 2 CTAs per SM; the dot-product partial
// of a CTA is accumulated in registers over all its blocks (296 partials, fixed order).
// ---------------------------------------------------------------------------------------------
constexpr int SPMV2_CONSUMERS = 256;
constexpr int SPMV2_THREADS = SPMV2_CONSUMERS + 32;
#ifndef PSFEM_SPMV_STAGES
#define PSFEM_SPMV_STAGES 2
#endif
constexpr int SPMV2_STAGES = PSFEM_SPMV_STAGES;
constexpr int SPMV2_ROWCAP = 1024;  // row_ptr entries staged per block (blocks with more rows read row_ptr from global)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

struct Spmv2Meta {  // per stage, written by the producer before the stage's bulk copies complete
    int r0, r1, k0a, r0a, rows_staged, kb /* absolute end of the bulk-copied non-zeros */, pad[2];
};

template <bool DOT>
__global__ void __launch_bounds__(SPMV2_THREADS, 2)
spmv2_kernel(const int32_t *__restrict__ row_ptr, const int32_t *__restrict__ col_idx, const double *__restrict__ vals,
             const int32_t *__restrict__ rowblk, int n_blocks, int n_rows, int cap, const double *__restrict__ x, double alpha,
             double gamma, const double *__restrict__ z, double *__restrict__ y, const double *__restrict__ xdot,
             double *__restrict__ dot_out, double *__restrict__ partials, unsigned int *counter, const int *__restrict__ done,
             const PeerSync ps) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t full_bar[SPMV2_STAGES], empty_bar[SPMV2_STAGES];
    __shared__ Spmv2Meta meta[SPMV2_STAGES];
    __shared__ double red[SPMV2_THREADS / 32];
    pdl_enter();
    if (done && *done) return;
    // stage layout: vals[cap] | cols[cap] | rp[ROWCAP+8]
    const size_t stage_bytes = (size_t)cap * 12 + (SPMV2_ROWCAP + 8) * 4;
    auto vals_s = [&](int s) { return reinterpret_cast<double *>(smem_raw + s * stage_bytes); };
    auto cols_s = [&](int s) { return reinterpret_cast<int *>(smem_raw + s * stage_bytes + (size_t)cap * 8); };
    auto rp_s = [&](int s) { return reinterpret_cast<int *>(smem_raw + s * stage_bytes + (size_t)cap * 12); };
    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int s = 0; s < SPMV2_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], SPMV2_CONSUMERS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int n_mine = (n_blocks - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    double dot[1] = {0.0};
    if (tid >= SPMV2_CONSUMERS) {
        // ---------------- producer warp ----------------
        // The 32 lanes fetch the (row, non-zero) ranges of the CTA's next 32 row blocks with two batched
        // loads (rowblk, then row_ptr), so the dependent global-load latency is paid once per 32 blocks
        // instead of once per block; lane 0 then issues the bulk copies block by block.
        const int lane = tid & 31;
        const int kend4 = row_ptr[n_rows] & ~3;
        const int rp_mis = (int)((reinterpret_cast<uintptr_t>(row_ptr) >> 2) & 3);  // row_ptr may be a row slice
        for (int base = 0; base < n_mine; base += 32) {
            int r0_l = 0, r1_l = 0, k0_l = 0, k1_l = 0;
            if (base + lane < n_mine) {
                const int b = blockIdx.x + (base + lane) * gridDim.x;
                r0_l = rowblk[b]; r1_l = rowblk[b + 1];
                k0_l = row_ptr[r0_l]; k1_l = row_ptr[r1_l];
            }
            const int nj = min(32, n_mine - base);
            for (int j = 0; j < nj; ++j) {
                const int r0 = __shfl_sync(0xffffffffu, r0_l, j), r1 = __shfl_sync(0xffffffffu, r1_l, j);
                const int k0 = __shfl_sync(0xffffffffu, k0_l, j), k1 = __shfl_sync(0xffffffffu, k1_l, j);
                if (lane == 0) {
                    // bulk copies need 16-byte aligned addresses and sizes: slices are widened to multiples of
                    // 4 elements but never past the last offset this operand references (row_ptr[n_rows]); the
                    // few non-zeros beyond the last aligned boundary are read from global memory by consumers.
                    const int i = base + j;
                    const int s = i % SPMV2_STAGES, n = i / SPMV2_STAGES;
                    mbar_wait(&empty_bar[s], (n & 1) ^ 1);       // consumers are done with this stage
                    const int k0a = k0 & ~3;
                    int kb = (k1 + 3) & ~3;
                    if (kb > kend4) kb = kend4;
                    if (kb < k0a) kb = k0a;
                    const int r0a = r0 - ((r0 + rp_mis) & 3);        // (row_ptr + r0a) is 16-byte aligned
                    int rows_staged = r1 - r0a + 1;                  // row_ptr[r0a .. r1]
                    int rp_cnt = (rows_staged + 3) & ~3;
                    if (rows_staged > SPMV2_ROWCAP || r0a + rp_cnt > n_rows + 1) { rows_staged = 0; rp_cnt = 0; }
                    Spmv2Meta m;
                    m.r0 = r0; m.r1 = r1; m.k0a = k0a; m.r0a = r0a; m.rows_staged = rows_staged; m.kb = kb;
                    meta[s] = m;
                    const uint32_t nb = (uint32_t)(kb - k0a);
                    mbar_expect_tx(&full_bar[s], nb * 12u + (uint32_t)rp_cnt * 4u);
                    if (nb) {
                        bulk_g2s(vals_s(s), vals + k0a, nb * 8u, &full_bar[s]);
                        bulk_g2s(cols_s(s), col_idx + k0a, nb * 4u, &full_bar[s]);
                    }
                    if (rp_cnt) bulk_g2s(rp_s(s), row_ptr + r0a, (uint32_t)rp_cnt * 4u, &full_bar[s]);
                }
                __syncwarp();
            }
        }
    } else {
        // ---------------- consumer warps ----------------
        for (int i = 0; i < n_mine; ++i) {
            const int s = i % SPMV2_STAGES, n = i / SPMV2_STAGES;
            mbar_wait(&full_bar[s], n & 1);
            const Spmv2Meta m = meta[s];
            double *vs = vals_s(s);
            const int *cs = cols_s(s);
            const int *rp = rp_s(s);
            // one row per thread straight from the staged slices: the gathers of a warp hit consecutive x
            // entries (adjacent rows share / neighbour their columns), every gather of a batch is in flight
            // together, and the row is summed in column order (the order of scipy's CSR mat-vec).
            constexpr int U = 9;
            for (int r = m.r0 + tid; r < m.r1; r += SPMV2_CONSUMERS) {
                int a, b;
                if (m.rows_staged) { a = rp[r - m.r0a]; b = rp[r + 1 - m.r0a]; }
                else { a = row_ptr[r]; b = row_ptr[r + 1]; }
                double sum = 0.0;
                for (int q = a; q < b; q += U) {
                    int c[U];
                    double v[U], xv[U];
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const int kk = q + u;
                        if (kk < b) {
                            const bool staged = kk < m.kb;   // beyond kb: unaligned tail of the last block
                            c[u] = staged ? cs[kk - m.k0a] : col_idx[kk];
                            v[u] = staged ? vs[kk - m.k0a] : vals[kk];
                        }
                    }
#pragma unroll
                    for (int u = 0; u < U; ++u)
                        if (q + u < b) xv[u] = __ldg(x + c[u]);
#pragma unroll
                    for (int u = 0; u < U; ++u)
                        if (q + u < b) sum += alpha * v[u] * xv[u];
                }
                if (z) sum += gamma * z[r];
                y[r] = sum;
                if (DOT) dot[0] += xdot[r] * sum;
            }
            mbar_arrive(&empty_bar[s]);
        }
    }
    if (DOT) {
        double total[1];
        if (grid_sum<SPMV2_THREADS, 1>(dot, partials, counter, red, total)) {
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

// ---------------------------------------------------------------------------------------------
// SpMV on the NODE-BLOCK structure of FEM matrices (2 DOFs per node).
//
// Every matrix this library assembles -- and every reduction of one that removes whole nodes -- has rows 2a, 2a+1
// with the same columns {2b, 2b+1 : b neighbour of a}.  One 4-byte node id per 2x2 block then replaces four 4-byte
// column indices: the kernel streams 8 B (value) + 1 B (index) per non-zero instead of 12 B, 25 % less HBM traffic
// for the kernel that dominates PCG.  Same pipeline as spmv2_kernel (producer warp, TMA bulk copies into a shared
// ring, one row per consumer thread, column-order summation => bit-identical results to the CSR kernels); x is
// gathered as double2 {x[2b], x[2b+1]}.  psfem_spmv_nb_plan detects the structure and builds node_ptr / pcol /
// node blocks; operands without it keep using the CSR kernels.
// ---------------------------------------------------------------------------------------------
#ifndef PSFEM_NB_PAIRS
#define PSFEM_NB_PAIRS 1024
#endif
constexpr int NB_PAIRS = PSFEM_NB_PAIRS;      // target 2x2 blocks per node block (4096 non-zeros at 1024)
constexpr int NB_NODECAP = 1024;    // node_ptr entries staged per block
#ifndef PSFEM_NB_U
#define PSFEM_NB_U 5
#endif
constexpr int NB_U = PSFEM_NB_U;

__global__ void nb_build_kernel(int64_t n_nodes, const int32_t *__restrict__ row_ptr, const int32_t *__restrict__ col_idx,
                                int32_t *__restrict__ node_ptr, int32_t *__restrict__ pcol, int *__restrict__ info /* [0] bad, [1] max pairs, [2] max |column node - row node| */) {
    const int64_t a = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (a > n_nodes) return;
    const int32_t base = row_ptr[0];
    if (a == n_nodes) { node_ptr[a] = row_ptr[2 * a] >> 2; if ((row_ptr[2 * a] | base) & 3) atomicExch(info, 1); return; }
    const int32_t r0 = row_ptr[2 * a], r1 = row_ptr[2 * a + 1], r2 = row_ptr[2 * a + 2];
    const int len = r1 - r0;
    bool ok = (r2 - r1 == len) && !(len & 1) && !(r0 & 3);
    int maxd = 0;
    node_ptr[a] = r0 >> 2;
    if (ok) {
        const int deg = len >> 1;
        int32_t *pc = pcol + ((r0 - base) >> 2);
        for (int k = 0; k < deg; ++k) {
            const int32_t c0 = col_idx[r0 + 2 * k], c1 = col_idx[r0 + 2 * k + 1];
            ok = ok && !(c0 & 1) && c1 == c0 + 1 && col_idx[r1 + 2 * k] == c0 && col_idx[r1 + 2 * k + 1] == c1;
            pc[k] = c0 >> 1;
            const long long d = (long long)(c0 >> 1) - a;
            maxd = max(maxd, (int)min(d < 0 ? -d : d, (long long)INT32_MAX));
        }
        atomicMax(info + 1, deg);
        atomicMax(info + 2, maxd);
    }
    if (!ok) atomicExch(info, 1);
}

// 16-bit form of pcol: column node - row node (FEM numberings keep neighbours close; the plan falls back to 32 bits otherwise)
__global__ void nb_compress_kernel(int64_t n_nodes, const int32_t *__restrict__ node_ptr, const int32_t *__restrict__ pcol,
                                   int16_t *__restrict__ pcol16) {
    const int64_t a = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (a >= n_nodes) return;
    const int base = node_ptr[0];
    for (int k = node_ptr[a] - base; k < node_ptr[a + 1] - base; ++k) pcol16[k] = (int16_t)(pcol[k] - (int32_t)a);
}

__global__ void nb_plan_kernel(const int32_t *__restrict__ node_ptr, int64_t n_nodes, int64_t n_blocks, int32_t *__restrict__ nodeblk) {
    int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (b > n_blocks) return;
    if (b == n_blocks) { nodeblk[b] = (int32_t)n_nodes; return; }
    if (b == 0) { nodeblk[0] = 0; return; }
    const int64_t target = (int64_t)node_ptr[0] + b * NB_PAIRS;
    int64_t lo = 0, hi = n_nodes;
    while (lo < hi) {  // first node whose first pair is >= target
        int64_t mid = lo + ((hi - lo) >> 1);
        if (node_ptr[mid] < target) lo = mid + 1; else hi = mid;
    }
    nodeblk[b] = (int32_t)lo;
}

struct NbMeta {  // per stage
    int a0, a1, P0, pc0 /* first staged pcol entry, relative to the operand's first pair */, a0a, staged_nodes, pad[2];
};

template <bool DOT, bool IDX16>
__global__ void __launch_bounds__(SPMV2_THREADS, 2)
spmv_nb_kernel(const int32_t *__restrict__ node_ptr, const void *__restrict__ pcol_, const double *__restrict__ vals,
               const int32_t *__restrict__ nodeblk, int n_blocks, int n_nodes, int cap /* pairs per stage */,
               const double *__restrict__ x, double alpha, double gamma, const double *__restrict__ z, double *__restrict__ y,
               const double *__restrict__ xdot, double *__restrict__ dot_out, double *__restrict__ partials, unsigned int *counter,
               const int *__restrict__ done, const PeerSync ps) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t full_bar[SPMV2_STAGES], empty_bar[SPMV2_STAGES];
    __shared__ NbMeta meta[SPMV2_STAGES];
    __shared__ double red[SPMV2_THREADS / 32];
    pdl_enter();
    if (done && *done) return;
    // stage layout: vals[4*cap] | pcol[cap+16] (4- or 2-byte entries, 4 bytes reserved each) | np[NB_NODECAP+8]
    using PC = typename std::conditional<IDX16, int16_t, int32_t>::type;
    constexpr int PCA = IDX16 ? 8 : 4;            // pcol entries per 16 bytes: bulk copies start and end on such boundaries
    const PC *__restrict__ pcol = static_cast<const PC *>(pcol_);
    const size_t stage_bytes = (size_t)cap * 32 + (size_t)(cap + 16) * 4 + (NB_NODECAP + 8) * 4;
    auto vals_s = [&](int s) { return reinterpret_cast<double *>(smem_raw + s * stage_bytes); };
    auto pc_s = [&](int s) { return reinterpret_cast<PC *>(smem_raw + s * stage_bytes + (size_t)cap * 32); };
    auto np_s = [&](int s) { return reinterpret_cast<int *>(smem_raw + s * stage_bytes + (size_t)cap * 32 + (size_t)(cap + 16) * 4); };
    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int s = 0; s < SPMV2_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], SPMV2_CONSUMERS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int n_mine = (n_blocks - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int pair_base = node_ptr[0];
    const double2 *__restrict__ x2 = reinterpret_cast<const double2 *>(x);
    double dot[1] = {0.0};
    if (tid >= SPMV2_CONSUMERS) {
        // ---------------- producer warp ----------------
        const int lane = tid & 31;
        for (int base = 0; base < n_mine; base += 32) {
            int a0_l = 0, a1_l = 0, p0_l = 0, p1_l = 0;
            if (base + lane < n_mine) {
                const int b = blockIdx.x + (base + lane) * gridDim.x;
                a0_l = nodeblk[b]; a1_l = nodeblk[b + 1];
                p0_l = node_ptr[a0_l]; p1_l = node_ptr[a1_l];
            }
            const int nj = min(32, n_mine - base);
            for (int j = 0; j < nj; ++j) {
                const int a0 = __shfl_sync(0xffffffffu, a0_l, j), a1 = __shfl_sync(0xffffffffu, a1_l, j);
                const int P0 = __shfl_sync(0xffffffffu, p0_l, j), P1 = __shfl_sync(0xffffffffu, p1_l, j);
                if (lane == 0) {
                    const int i = base + j;
                    const int s = i % SPMV2_STAGES, n = i / SPMV2_STAGES;
                    mbar_wait(&empty_bar[s], (n & 1) ^ 1);
                    // pcol and node_ptr are this plan's own arrays (16-byte aligned, padded by 8 entries): widen freely
                    const int pc0 = (P0 - pair_base) & ~(PCA - 1);
                    const int pc1 = (P1 - pair_base + PCA - 1) & ~(PCA - 1);
                    const int a0a = a0 & ~3;
                    int staged = a1 - a0a + 1;
                    int np_cnt = (staged + 3) & ~3;
                    if (staged > NB_NODECAP) { staged = 0; np_cnt = 0; }
                    NbMeta m;
                    m.a0 = a0; m.a1 = a1; m.P0 = P0; m.pc0 = pc0; m.a0a = a0a; m.staged_nodes = staged;
                    meta[s] = m;
                    const uint32_t vb = (uint32_t)(P1 - P0) * 32u, pb = (uint32_t)(pc1 - pc0) * (uint32_t)sizeof(PC);
                    mbar_expect_tx(&full_bar[s], vb + pb + (uint32_t)np_cnt * 4u);
                    if (vb) bulk_g2s(vals_s(s), vals + 4ll * P0, vb, &full_bar[s]);
                    if (pb) bulk_g2s(pc_s(s), pcol + pc0, pb, &full_bar[s]);
                    if (np_cnt) bulk_g2s(np_s(s), node_ptr + a0a, (uint32_t)np_cnt * 4u, &full_bar[s]);
                }
                __syncwarp();
            }
        }
    } else {
        // ---------------- consumer warps: one DOF row per thread ----------------
        for (int i = 0; i < n_mine; ++i) {
            const int s = i % SPMV2_STAGES, n = i / SPMV2_STAGES;
            mbar_wait(&full_bar[s], n & 1);
            const NbMeta m = meta[s];
            const double *vs = vals_s(s);
            const PC *pc = pc_s(s);
            const int *np = np_s(s);
            constexpr int U = NB_U;   // 2x2 blocks per batch (double2 gathers in flight per thread)
            const int n_rows_blk = 2 * (m.a1 - m.a0);
            for (int rl = tid; rl < n_rows_blk; rl += SPMV2_CONSUMERS) {
                const int a = m.a0 + (rl >> 1), p = rl & 1;
                int q0, q1;
                if (m.staged_nodes) { q0 = np[a - m.a0a]; q1 = np[a + 1 - m.a0a]; }
                else { q0 = node_ptr[a]; q1 = node_ptr[a + 1]; }
                const int deg = q1 - q0;
                const double2 *v = reinterpret_cast<const double2 *>(vs + 4ll * (q0 - m.P0)) + (size_t)p * deg;   // row p of the node block
                const PC *c = pc + (q0 - pair_base - m.pc0);
                const int cbase = IDX16 ? a : 0;       // 16-bit entries are offsets from the row node
                double sum = 0.0;
                for (int k = 0; k < deg; k += U) {
                    double2 xv[U], vv[U];
#pragma unroll
                    for (int u = 0; u < U; ++u)
                        if (k + u < deg) { xv[u] = __ldg(x2 + (cbase + (int)c[k + u])); vv[u] = v[k + u]; }
#pragma unroll
                    for (int u = 0; u < U; ++u)
                        if (k + u < deg) { sum += alpha * vv[u].x * xv[u].x; sum += alpha * vv[u].y * xv[u].y; }
                }
                const int r = 2 * a + p;
                if (z) sum += gamma * z[r];
                y[r] = sum;
                if (DOT) dot[0] += xdot[r] * sum;
            }
            mbar_arrive(&empty_bar[s]);
        }
    }
    if (DOT) {
        double total[1];
        if (grid_sum<SPMV2_THREADS, 1>(dot, partials, counter, red, total)) {
            *dot_out = total[0];
            if (ps.world) {
                peer_publish<1>(ps, 0, total);
                double pq[1];
                peer_wait_sum<1>(ps, 0, ps.seq, pq);
                ps.scal[4] = pq[0];
            }
        }
    }
}

#include "psfem_spmv_sl.cuh"
#include "psfem_cg_sl.cuh"

// geometry of the symmetric-lattice operand attached to A (psfem_spmv_sl_build)
bool sl_geometry(const psfem_csr *A, SlGeom *g, int sms, int ctas_per_sm = PSFEM_SL_CTAS) {
    const int64_t m = A->sl_m, nodes = A->n_rows / 2;
    if (m < 2 || nodes % m || A->sl_roff % m || A->sl_x_nodes % m) return false;
    const int64_t owned_cols = nodes / m, c0x = A->sl_roff / m;
    g->m = (int)m; g->mpad = (int)((m + 3) & ~3ll);
    g->pre = c0x > 0 ? 1 : 0;
    g->S = (int)owned_cols + g->pre;
    g->xcol0 = (int)c0x - g->pre;
    g->xcols = (int)(A->sl_x_nodes / m);
    g->n_strips = (int)ceil_div(m, SL_ROWS);
    int64_t G = (int64_t)sms * ctas_per_sm * SL_WARPS / g->n_strips;     // one resident wave of warps
    if (G > owned_cols / 4) G = owned_cols / 4;                           // segments of at least 4 columns
    if (G < 1) G = 1;
    g->G = (int)G;
    return true;
}

// does launch_spmv run spmv_sl_kernel for this call?  (also decides where the strip solver waits for its halo)
bool sl_usable(const psfem_csr *A, bool dual, const double *x1, const double *y, const double *z, const double *xdot, bool dot) {
    static const bool sl_off = getenv("PSFEM_NO_SYMLATTICE") != nullptr;
    SlGeom g;
    return !dual && !sl_off && A->sl_vals && A->sl_src_vals == A->vals && (!dot || xdot == x1 + 2 * A->sl_roff) && y != x1 &&
           ((reinterpret_cast<uintptr_t>(x1) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(z)) & 15) == 0 &&
           sl_geometry(A, &g, 148);
}

int launch_spmv(const psfem_csr *A, const double *v2, const double *x1, const double *x2, double alpha, double beta,
                double gamma, const double *z, double *y, const double *xdot, double *dot_out, double *partials,
                unsigned int *counter, const int *done, cudaStream_t stream, const PeerSync *peer = nullptr) {
    PSFEM_REQUIRE(A, "spmv: null matrix");
    PeerSync ps;
    memset(&ps, 0, sizeof(ps));
    if (peer) ps = *peer;
    const bool dual = v2 != nullptr, dot = dot_out != nullptr;
    // operands assembled straight into the symmetric-lattice storage (psfem_assemble_lattice_sl) have no CSR side at all
    PSFEM_REQUIRE((A->rowblk && A->n_blocks > 0) || sl_usable(A, dual, x1, y, z, xdot, dot),
                  "spmv: matrix has no row-block plan (psfem_spmv_plan) and no usable symmetric-lattice storage");
    const size_t smem = (size_t)(SPMV_T + A->max_row_nnz + 8) * sizeof(double);
    PSFEM_REQUIRE(smem <= 200 * 1024, "spmv: a row with %lld non-zeros does not fit the shared-memory stage", (long long)A->max_row_nnz);
    static const bool nb_off = getenv("PSFEM_NO_NODEBLOCK") != nullptr;   // A/B switch for measurements
    // symmetric-lattice storage of exactly these values (psfem_spmv_sl_build): half the HBM traffic, same result
    PSFEM_REQUIRE(!ps.halo_in_spmv || sl_usable(A, dual, x1, y, z, xdot, dot), "spmv: the halo wait was delegated to a kernel that does not run");
    if (sl_usable(A, dual, x1, y, z, xdot, dot)) {
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        SlGeom g;
        static const bool no_tma = getenv("PSFEM_SPMV_NO_TMA") != nullptr;   // A/B switch: the register-prefetch kernel
        // TMA-staged kernel (single-GPU operands whose storage carries the window slack: see psfem_spmv_sl_size)
        if (!no_tma && !peer && sl_geometry(A, &g, sms, 2) && g.S - g.pre >= 8) {
            const int n_groups = (int)ceil_div(g.n_strips, CGT_CONSUMERS);
            int64_t G = (int64_t)sms * 2 / n_groups;
            const int64_t owned_cols = g.S - g.pre;
            if (G > owned_cols / 4) G = owned_cols / 4;
            if (G < 1) G = 1;
            g.G = (int)G;
            const unsigned grid = (unsigned)(n_groups * g.G);
            const size_t smem_t = sizeof(SpStage) * SPT_STAGES + 128;
            if (dot) {
                PSFEM_CUDA_CHECK(cudaFuncSetAttribute(spmv_sl_tma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_t));
                PSFEM_CUDA_CHECK(launch_pdl(spmv_sl_tma_kernel<true>, grid, CGT_THREADS, smem_t, stream, A->sl_vals, g, x1, alpha, gamma, z, y, dot_out, partials, counter, done, n_groups));
            } else {
                PSFEM_CUDA_CHECK(cudaFuncSetAttribute(spmv_sl_tma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_t));
                PSFEM_CUDA_CHECK(launch_pdl(spmv_sl_tma_kernel<false>, grid, CGT_THREADS, smem_t, stream, A->sl_vals, g, x1, alpha, gamma, z, y, dot_out, partials, counter, done, n_groups));
            }
            PSFEM_LAUNCH_CHECK();
            return PSFEM_OK;
        }
        if (sl_geometry(A, &g, sms)) {
            const unsigned grid = (unsigned)ceil_div((int64_t)g.n_strips * g.G, SL_WARPS);
            if (dot)
                PSFEM_CUDA_CHECK(launch_pdl(spmv_sl_kernel<true>, grid, SL_THREADS, 0, stream, A->sl_vals, g, x1, alpha, gamma, z, y, dot_out, partials, counter, done, ps));
            else
                PSFEM_CUDA_CHECK(launch_pdl(spmv_sl_kernel<false>, grid, SL_THREADS, 0, stream, A->sl_vals, g, x1, alpha, gamma, z, y, dot_out, partials, counter, done, ps));
            PSFEM_LAUNCH_CHECK();
            return PSFEM_OK;
        }
    }
    if (!dual && !nb_off && A->nb_node_ptr && (A->nb_pcol || A->nb_pcol16) && A->nb_n_blocks >= 8 && ((reinterpret_cast<uintptr_t>(A->vals) | reinterpret_cast<uintptr_t>(x1)) & 15) == 0) {
        const int cap = (int)((NB_PAIRS + A->nb_max_pairs + 8 + 3) & ~3ll);
        const size_t stage_bytes = (size_t)cap * 32 + (size_t)(cap + 16) * 4 + (NB_NODECAP + 8) * 4;
        const bool idx16 = A->nb_pcol16 != nullptr;
        const void *pcol = idx16 ? static_cast<const void *>(A->nb_pcol16) : static_cast<const void *>(A->nb_pcol);
        const size_t smem_nb = SPMV2_STAGES * stage_bytes;
        if (smem_nb <= 110 * 1024) {
            int dev = 0, sms = 148;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
            int64_t grid = 2ll * sms;
            if (grid > A->nb_n_blocks) grid = A->nb_n_blocks;
#define PSFEM_NB_LAUNCH(D, I)                                                                                              \
    do {                                                                                                                   \
        PSFEM_CUDA_CHECK(cudaFuncSetAttribute(spmv_nb_kernel<D, I>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_nb)); \
        PSFEM_CUDA_CHECK(launch_pdl(spmv_nb_kernel<D, I>, (unsigned)grid, SPMV2_THREADS, smem_nb, stream, A->nb_node_ptr, pcol, A->vals, A->nb_nodeblk,  \
            (int)A->nb_n_blocks, (int)(A->n_rows / 2), cap, x1, alpha, gamma, z, y, xdot, dot_out, partials, counter, done, ps)); \
    } while (0)
            if (dot && idx16) PSFEM_NB_LAUNCH(true, true);
            else if (dot) PSFEM_NB_LAUNCH(true, false);
            else if (idx16) PSFEM_NB_LAUNCH(false, true);
            else PSFEM_NB_LAUNCH(false, false);
#undef PSFEM_NB_LAUNCH
            PSFEM_LAUNCH_CHECK();
            return PSFEM_OK;
        }
    }
    if (!dual && A->n_blocks >= 8 && ((reinterpret_cast<uintptr_t>(A->vals) | reinterpret_cast<uintptr_t>(A->col_idx)) & 15) == 0) {
        // v2: persistent TMA-staged kernel (vals / col_idx base pointers must be 16-byte aligned)
        const int cap = (int)((SPMV_T + A->max_row_nnz + 8 + 3) & ~3ll);
        const size_t stage_bytes = (size_t)cap * 12 + (SPMV2_ROWCAP + 8) * 4;
        const size_t smem2 = SPMV2_STAGES * stage_bytes;
        if (smem2 <= 110 * 1024) {
            int dev = 0, sms = 148;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
            int64_t grid = 2ll * sms;
            if (grid > A->n_blocks) grid = A->n_blocks;
            if (dot) {
                PSFEM_CUDA_CHECK(cudaFuncSetAttribute(spmv2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
                PSFEM_CUDA_CHECK(launch_pdl(spmv2_kernel<true>, (unsigned)grid, SPMV2_THREADS, smem2, stream, A->row_ptr, A->col_idx, A->vals, A->rowblk, (int)A->n_blocks, (int)A->n_rows, cap,
                    x1, alpha, gamma, z, y, xdot, dot_out, partials, counter, done, ps));
            } else {
                PSFEM_CUDA_CHECK(cudaFuncSetAttribute(spmv2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
                PSFEM_CUDA_CHECK(launch_pdl(spmv2_kernel<false>, (unsigned)grid, SPMV2_THREADS, smem2, stream, A->row_ptr, A->col_idx, A->vals, A->rowblk, (int)A->n_blocks, (int)A->n_rows, cap,
                    x1, alpha, gamma, z, y, xdot, dot_out, partials, counter, done, ps));
            }
            PSFEM_LAUNCH_CHECK();
            return PSFEM_OK;
        }
    }
    PSFEM_REQUIRE(!peer, "spmv: the peer-synchronised SpMV needs the TMA-staged kernel (>= 8 row blocks, 16-byte aligned CSR arrays)");
#define PSFEM_SPMV_LAUNCH(D, O)                                                                                      \
    do {                                                                                                             \
        if (smem > 48 * 1024)                                                                                        \
            PSFEM_CUDA_CHECK(cudaFuncSetAttribute(spmv_kernel<D, O>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        spmv_kernel<D, O><<<(unsigned)A->n_blocks, SPMV_THREADS, smem, stream>>>(                                    \
            A->row_ptr, A->col_idx, A->vals, v2, A->rowblk, x1, x2, alpha, beta, gamma, z, y, xdot, dot_out, partials, counter, done); \
    } while (0)
    if (dual && dot) PSFEM_SPMV_LAUNCH(true, true);
    else if (dual) PSFEM_SPMV_LAUNCH(true, false);
    else if (dot) PSFEM_SPMV_LAUNCH(false, true);
    else PSFEM_SPMV_LAUNCH(false, false);
#undef PSFEM_SPMV_LAUNCH
    PSFEM_LAUNCH_CHECK();
    return PSFEM_OK;
}

// ---------------------------------------------------------------------------------------------
// diagonal, element-wise PCG kernels
// ---------------------------------------------------------------------------------------------
__global__ void extract_diag_kernel(int64_t n, const int32_t *__restrict__ row_ptr, const int32_t *__restrict__ col_idx,
                                    const double *__restrict__ vals, int64_t col_offset, double *__restrict__ diag, int invert) {
    int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r >= n) return;
    const int32_t want = (int32_t)(r + col_offset);
    double d = 0.0;
    for (int32_t k = row_ptr[r]; k < row_ptr[r + 1]; ++k)
        if (col_idx[k] == want) { d = vals[k]; break; }
    diag[r] = invert ? 1.0 / d : d;
}

// z = dinv*r, p = z, rz = r.z, rr = r.r, bb = b.b   (PCG start)
__global__ void __launch_bounds__(EW_THREADS)
pcg_init_kernel(int64_t n, double *__restrict__ r, const double *__restrict__ dinv, const double *__restrict__ b,
                double *__restrict__ p, double *__restrict__ out3, double *__restrict__ partials, unsigned int *counter,
                PcgCtl *ctl, double rtol, int masked) {
    pdl_enter();
    __shared__ double red[EW_THREADS / 32];
    double v[3] = {0, 0, 0};
    for (int64_t i = blockIdx.x * (int64_t)EW_THREADS + threadIdx.x; i < n; i += (int64_t)gridDim.x * EW_THREADS) {
        const double di = dinv[i];
        double ri = r[i];
        // masked solve (psfem_pcg_jacobi_masked): rows with D^-1 = 0 are constrained DOFs that stayed in the matrix;
        // their residual, search direction and right-hand side are pinned to zero
        const bool off = masked && di == 0.0;
        if (off) { ri = 0.0; r[i] = 0.0; }
        const double zi = di * ri;
        p[i] = zi;
        v[0] += ri * zi;
        v[1] += ri * ri;
        if (b && !off) { const double bi = b[i]; v[2] += bi * bi; }
    }
    double tot[3];
    if (grid_sum<EW_THREADS, 3>(v, partials, counter, red, tot)) {
        if (out3) { out3[0] = tot[0]; out3[1] = tot[1]; out3[2] = tot[2]; }
        if (ctl) {
            ctl->rz[0] = tot[0]; ctl->rz[1] = 0; ctl->pq = 0; ctl->rr = tot[1]; ctl->bb = tot[2];
            ctl->tol2 = rtol * rtol * tot[2];
            ctl->iters = 0; ctl->parity = 0;
            ctl->done = (tot[2] == 0.0 || tot[1] <= ctl->tol2) ? 1 : 0;
        }
    }
}

// r -= alpha q; rz_new = r.(dinv r); rr = r.r      alpha = rz/pq
// (x += alpha p is done by the p-update kernel, which streams p anyway: 8 B per DOF and iteration less than
//  reading p here as well)
template <bool MASKED>
__global__ void __launch_bounds__(EW_THREADS)
pcg_update_kernel(int64_t n, double *__restrict__ r,
                  const double *__restrict__ q, const double *__restrict__ dinv, const double *__restrict__ rz_in,
                  const double *__restrict__ pq_in, double *__restrict__ out2, double *__restrict__ partials,
                  unsigned int *counter, PcgCtl *ctl) {
    pdl_enter();
    __shared__ double red[EW_THREADS / 32];
    double rz, pq;
    if (ctl) {
        if (ctl->done) return;
        rz = ctl->rz[ctl->parity]; pq = ctl->pq;
    } else {
        rz = *rz_in; pq = *pq_in;
    }
    const double alpha = rz / pq;
    double v[2] = {0, 0};
    for (int64_t i = blockIdx.x * (int64_t)EW_THREADS + threadIdx.x; i < n; i += (int64_t)gridDim.x * EW_THREADS) {
        // MASKED: (K p) at a constrained row (D^-1 = 0) is not part of the system.  The mask is applied as a factor (x 1.0
        // is exact) so that the loads of r and q do not wait for D^-1: written as a select, ptxas predicated them on the
        // loaded value (dependent loads, +30 % kernel time).  The unmasked instance is the plain recurrence.
        const double di = dinv[i];
        const double ri = MASKED ? (r[i] - alpha * q[i]) * (di == 0.0 ? 0.0 : 1.0) : r[i] - alpha * q[i];
        r[i] = ri;
        v[0] += ri * (di * ri);
        v[1] += ri * ri;
    }
    double tot[2];
    if (grid_sum<EW_THREADS, 2>(v, partials, counter, red, tot)) {
        if (ctl) {
            ctl->rz[ctl->parity ^ 1] = tot[0];
            ctl->rr = tot[1];
            ctl->iters += 1;
            // `done` and `parity` are published by pcg_pupdate_kernel so that every kernel of this
            // iteration sees one consistent state
        } else {
            out2[0] = tot[0]; out2[1] = tot[1];
        }
    }
}

// x += alpha p (the step of THIS iteration, with the old p); p = dinv*r + beta p     alpha = rz_old/pq, beta = rz_new/rz_old
__global__ void __launch_bounds__(EW_THREADS)
pcg_pupdate_kernel(int64_t n, double *__restrict__ p, double *__restrict__ x, const double *__restrict__ r,
                   const double *__restrict__ dinv, const double *__restrict__ rz_old_in, const double *__restrict__ rz_new_in,
                   const double *__restrict__ pq_in, PcgCtl *ctl, unsigned int *counter) {
    pdl_enter();
    double rz_old, rz_new, pq;
    if (ctl) {
        if (ctl->done) return;
        rz_old = ctl->rz[ctl->parity]; rz_new = ctl->rz[ctl->parity ^ 1]; pq = ctl->pq;
    } else {
        rz_old = *rz_old_in; rz_new = *rz_new_in; pq = *pq_in;
    }
    const double beta = rz_new / rz_old, alpha = rz_old / pq;
    for (int64_t i = blockIdx.x * (int64_t)EW_THREADS + threadIdx.x; i < n; i += (int64_t)gridDim.x * EW_THREADS) {
        const double pi = p[i];
        x[i] += alpha * pi;
        p[i] = dinv[i] * r[i] + beta * pi;
    }
    if (ctl) {
        // last block flips parity / sets done after every block has read the old state
        __shared__ bool is_last;
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            unsigned int ticket = atomicAdd(counter, 1u);
            is_last = (ticket == gridDim.x - 1);
        }
        __syncthreads();
        if (is_last && threadIdx.x == 0) {
            *counter = 0;
            ctl->parity ^= 1;
            if (ctl->rr <= ctl->tol2 || !(ctl->pq > 0.0)) ctl->done = 1;  // converged or breakdown
            __threadfence();
        }
    }
}

__global__ void pq_store_kernel(PcgCtl *ctl, const double *pq) { ctl->pq = *pq; }

// ---------------------------------------------------------------------------------------------
// single-CTA PCG for small systems (whole solve in one launch; Newmark / modal on small meshes)
// ---------------------------------------------------------------------------------------------
constexpr int SMALL_THREADS = 1024;
constexpr int SMALL_MAX_N = 16384;

__global__ void __launch_bounds__(SMALL_THREADS)
pcg_small_kernel(int n, const int32_t *__restrict__ row_ptr, const int32_t *__restrict__ col_idx,
                 const double *__restrict__ vals, const double *__restrict__ b, double *__restrict__ x,
                 double *__restrict__ r, double *__restrict__ p, double *__restrict__ q, const double *__restrict__ dinv,
                 double rtol, long long maxit, double *__restrict__ info /* iters, relres, bnorm, status */) {
    __shared__ double red[SMALL_THREADS / 32];
    __shared__ double bc[2];
    const int t = threadIdx.x;
    auto bsum = [&](double v) -> double {
        double s = block_sum<SMALL_THREADS>(v, red);
        if (t == 0) bc[0] = s;
        __syncthreads();
        double o = bc[0];
        __syncthreads();
        return o;
    };
    // r = b - A x ; z = dinv r ; p = z
    double lrz = 0, lrr = 0, lbb = 0;
    for (int i = t; i < n; i += SMALL_THREADS) {
        double s = 0;
        for (int k = row_ptr[i]; k < row_ptr[i + 1]; ++k) s += vals[k] * x[col_idx[k]];
        const double bi = b[i], ri = bi - s, zi = dinv[i] * ri;
        r[i] = ri; p[i] = zi;
        lrz += ri * zi; lrr += ri * ri; lbb += bi * bi;
    }
    double rz = bsum(lrz), rr = bsum(lrr);
    const double bb = bsum(lbb);
    const double tol2 = rtol * rtol * bb;
    long long it = 0;
    int status = 0;
    while (bb > 0.0 && rr > tol2 && it < maxit) {
        double lpq = 0;
        for (int i = t; i < n; i += SMALL_THREADS) {
            double s = 0;
            for (int k = row_ptr[i]; k < row_ptr[i + 1]; ++k) s += vals[k] * p[col_idx[k]];
            q[i] = s;
            lpq += p[i] * s;
        }
        const double pq = bsum(lpq);
        if (!(pq > 0.0)) { status = 2; break; }
        const double alpha = rz / pq;
        lrz = 0; lrr = 0;
        for (int i = t; i < n; i += SMALL_THREADS) {
            x[i] += alpha * p[i];
            const double ri = r[i] - alpha * q[i];
            r[i] = ri;
            lrz += ri * (dinv[i] * ri); lrr += ri * ri;
        }
        const double rz_new = bsum(lrz);
        rr = bsum(lrr);
        const double beta = rz_new / rz;
        rz = rz_new;
        for (int i = t; i < n; i += SMALL_THREADS) p[i] = dinv[i] * r[i] + beta * p[i];
        __syncthreads();
        ++it;
    }
    if (t == 0) {
        if (status == 0 && bb > 0.0 && rr > tol2) status = 1;
        info[0] += (double)it;
        const double rel = bb > 0.0 ? sqrt(rr / bb) : 0.0;
        info[1] = fmax(info[1], rel);
        info[2] = sqrt(bb);
        info[3] = fmax(info[3], (double)status);
    }
}

// ---------------------------------------------------------------------------------------------
// BLAS-1
// ---------------------------------------------------------------------------------------------
__global__ void axpby_kernel(int64_t n, double a, const double *__restrict__ x, double b, const double *__restrict__ y, double *__restrict__ out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double v = a * x[i];
        if (y) v += b * y[i];
        out[i] = v;
    }
}

__global__ void __launch_bounds__(EW_THREADS)
dot_kernel(int64_t n, const double *__restrict__ x, const double *__restrict__ y, double *__restrict__ out,
           double *__restrict__ partials, unsigned int *counter) {
    __shared__ double red[EW_THREADS / 32];
    double v[1] = {0};
    for (int64_t i = blockIdx.x * (int64_t)EW_THREADS + threadIdx.x; i < n; i += (int64_t)gridDim.x * EW_THREADS) v[0] += x[i] * y[i];
    double tot[1];
    if (grid_sum<EW_THREADS, 1>(v, partials, counter, red, tot)) *out = tot[0];
}

// out[j] = V[j,:] . w : one CTA per row j, deterministic block tree
__global__ void __launch_bounds__(EW_THREADS)
multi_dot_kernel(int64_t n, const double *__restrict__ V, const double *__restrict__ w, double *__restrict__ out) {
    __shared__ double red[EW_THREADS / 32];
    const double *v = V + (int64_t)blockIdx.x * n;
    double s = 0;
    for (int64_t i = threadIdx.x; i < n; i += EW_THREADS) s += v[i] * w[i];
    s = block_sum<EW_THREADS>(s, red);
    if (threadIdx.x == 0) out[blockIdx.x] = s;
}

// w -= sum_j c[j] V[j,:]
__global__ void multi_axpy_kernel(int64_t m, int64_t n, const double *__restrict__ V, const double *__restrict__ c, double *__restrict__ w) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double s = w[i];
        for (int64_t j = 0; j < m; ++j) s -= c[j] * V[j * n + i];
        w[i] = s;
    }
}

// ---------------------------------------------------------------------------------------------
// strip-partitioned PCG with the collectives fused into the kernels (see PeerBoard above)
// ---------------------------------------------------------------------------------------------
struct HaloPush {   // where the first / last owned entries of p go in the neighbours' extended vectors
    double *left, *right;
    long long left_n, right_n, n;
    // Called by every thread AFTER its grid-stride loop over i = base, base + stride, ...: the thread re-reads the
    // halo entries it wrote itself (same index mapping, so program order suffices) and stores them into the
    // neighbours' vectors.  Kept out of the streaming loop so that loop stays free of possibly aliasing stores.
    __device__ __forceinline__ bool operator()(const double *p, long long base, long long stride) const {
        bool sent = false;
        if (left)
            for (long long i = base; i < left_n; i += stride) { left[i] = p[i]; sent = true; }
        if (right) {
            const long long first = n - right_n;
            long long i = base;
            if (i < first) i += (first - i + stride - 1) / stride * stride;
            for (; i < n; i += stride) { right[i - first] = p[i]; sent = true; }
        }
        return sent;
    }
};
__device__ __forceinline__ void halo_signal(const PeerSync &ps) {   // one thread, after a system-scope fence that follows every block's halo stores
    if (ps.has_left) st_relaxed_sys(&ps.board[ps.rank - 1]->halo_flag[1], ps.seq + 1);
    if (ps.has_right) st_relaxed_sys(&ps.board[ps.rank + 1]->halo_flag[0], ps.seq + 1);
}

// x = 0, r = b, p = dinv r (+ halo push), publishes {r.z, r.r, b.b}
__global__ void __launch_bounds__(EW_THREADS)
dist_init_kernel(int64_t n, const double *__restrict__ b, const double *__restrict__ dinv, double *__restrict__ x,
                 double *__restrict__ r, double *__restrict__ p, HaloPush halo, PeerSync ps, double *__restrict__ partials,
                 unsigned int *counter) {
    pdl_enter();
    __shared__ double red[EW_THREADS / 32];
    double v[3] = {0, 0, 0};
    for (int64_t i = blockIdx.x * (int64_t)EW_THREADS + threadIdx.x; i < n; i += (int64_t)gridDim.x * EW_THREADS) {
        const double bi = b[i], zi = dinv[i] * bi;
        x[i] = 0.0; r[i] = bi; p[i] = zi;
        v[0] += bi * zi;
        v[1] += bi * bi;
    }
    v[2] = v[1];
    halo(p, blockIdx.x * (long long)EW_THREADS + threadIdx.x, (long long)gridDim.x * EW_THREADS);
    // the block barriers inside grid_sum order every thread's halo stores before thread 0's system-scope fence
    // below (cumulativity), so one fence per block suffices
    double tot[3];
    if (grid_sum<EW_THREADS, 3, true>(v, partials, counter, red, tot)) {
        peer_publish<3>(ps, 1, tot);
        __threadfence_system();   // the halo stores of every block (ordered before this thread by the ticket) precede the halo flags
        halo_signal(ps);
        double g[3];
        peer_wait_sum<3>(ps, 1, ps.seq, g);
        ps.scal[0] = g[0]; ps.scal[1] = g[1]; ps.scal[2] = g[2]; ps.scal[5 + (int)(ps.seq & 1)] = g[0];
        PeerSync nxt = ps;
        nxt.seq = ps.seq + 1;
        peer_wait_halo(nxt);   // the neighbours' first p columns are in my extended vector before the first SpMV starts
    }
}

// alpha = (r.z)/(p.q) from the boards; r -= alpha q; publishes {r.z, r.r}  (x += alpha p: see dist_pupdate_kernel)
template <bool MASKED>
__global__ void __launch_bounds__(EW_THREADS)
dist_update_kernel(int64_t n, double *__restrict__ r,
                   const double *__restrict__ q, const double *__restrict__ dinv, PeerSync ps, double *__restrict__ partials,
                   unsigned int *counter) {
    pdl_enter();
    __shared__ double red[EW_THREADS / 32];
    const double alpha = ps.scal[5 + (int)((ps.seq - 1) & 1)] / ps.scal[4];   // global sums left by the previous kernels' last blocks
    double v[2] = {0, 0};
    for (int64_t i = blockIdx.x * (int64_t)EW_THREADS + threadIdx.x; i < n; i += (int64_t)gridDim.x * EW_THREADS) {
        // D^-1 = 0 marks a constrained DOF that stayed in the matrix (masked strips: boundary conditions that remove
        // single DOFs of a node): its residual stays zero, (K p) at that row is not part of the system
        const double di = dinv[i];
        const double ri = MASKED ? (r[i] - alpha * q[i]) * (di == 0.0 ? 0.0 : 1.0) : r[i] - alpha * q[i];   // a factor, not a select: see pcg_update_kernel
        r[i] = ri;
        v[0] += ri * (di * ri);
        v[1] += ri * ri;
    }
    double tot[2];
    if (grid_sum<EW_THREADS, 2>(v, partials, counter, red, tot)) {
        const double t3[3] = {tot[0], tot[1], 0.0};
        peer_publish<3>(ps, 1, t3);
        double g[3];
        peer_wait_sum<3>(ps, 1, ps.seq, g);
        ps.scal[0] = g[0]; ps.scal[1] = g[1]; ps.scal[5 + (int)(ps.seq & 1)] = g[0];
    }
}

// beta = (r.z)_new/(r.z)_old from the boards; p = dinv r + beta p, halo entries stored into the neighbours' vectors
__global__ void __launch_bounds__(EW_THREADS)
dist_pupdate_kernel(int64_t n, double *__restrict__ p, double *__restrict__ x, const double *__restrict__ r,
                    const double *__restrict__ dinv, HaloPush halo, PeerSync ps, unsigned int *counter) {
    pdl_enter();
    __shared__ bool is_last;
    const double rz_old = ps.scal[5 + (int)((ps.seq - 1) & 1)];
    const double beta = ps.scal[5 + (int)(ps.seq & 1)] / rz_old, alpha = rz_old / ps.scal[4];
    for (int64_t i = blockIdx.x * (int64_t)EW_THREADS + threadIdx.x; i < n; i += (int64_t)gridDim.x * EW_THREADS) {
        const double pi = p[i];
        x[i] += alpha * pi;                    // the step of this iteration, with the old p
        p[i] = dinv[i] * r[i] + beta * pi;
    }
    const bool sent = halo(p, blockIdx.x * (long long)EW_THREADS + threadIdx.x, (long long)gridDim.x * EW_THREADS);
    const int block_sent = __syncthreads_or(sent ? 1 : 0);
    if (threadIdx.x == 0) {
        // after the barrier one fence covers the halo stores of the whole block; blocks without halo entries skip it
        if (block_sent) __threadfence_system(); else __threadfence();
        unsigned int ticket = atomicAdd(counter, 1u);
        is_last = (ticket == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last && threadIdx.x == 0) {
        *counter = 0;
        __threadfence_system();
        halo_signal(ps);
        if (!ps.halo_in_spmv) {   // spmv_sl_kernel waits itself, in the two column segments that read the halo
            PeerSync nxt = ps;
            nxt.seq = ps.seq + 1;
            peer_wait_halo(nxt);   // my neighbours' columns for the next SpMV are in before this kernel ends
        }
    }
}

PeerSync make_peer(const psfem_dist_ctx *c, unsigned long long seq, bool halo_in_spmv = false) {
    PeerSync ps;
    memset(&ps, 0, sizeof(ps));
    ps.halo_in_spmv = halo_in_spmv ? 1 : 0;
    for (int w = 0; w < c->world && w < PSFEM_MAX_RANKS; ++w) ps.board[w] = static_cast<PeerBoard *>(c->board[w]);
    ps.world = c->world; ps.rank = c->rank; ps.has_left = c->has_left; ps.has_right = c->has_right;
    ps.seq = seq;
    ps.err = reinterpret_cast<int *>(c->counter + 3);
    ps.scal = c->out3;
    return ps;
}
HaloPush make_halo(const psfem_dist_ctx *c) {
    HaloPush h;
    h.left = c->has_left ? c->left_p_ext + c->left_dst_off : nullptr;
    h.right = c->has_right ? c->right_p_ext + c->right_dst_off : nullptr;
    h.left_n = c->send_left_n; h.right_n = c->send_right_n; h.n = c->n_own;
    return h;
}

struct CgWs {
    double *r[2], *s[2], *w[2], *p, *dinv, *partials, *scal;
    CgCtl *ctl;
    unsigned int *counter;
    size_t total;
};
CgWs carve_cg(void *ws, int64_t n) {
    CgWs w;
    Carver c(ws);
    // every vector is read through 116-row window copies (cg_sl_tma_kernel) that start two rows before a strip group and
    // may run past the last column: 256 bytes in front, 2 KB behind
    auto vec = [&]() { c.take<char>(256); double *p = c.take<double>(n); c.take<char>(2048); return p; };
    for (int q = 0; q < 2; ++q) { w.r[q] = vec(); w.s[q] = vec(); w.w[q] = vec(); }
    w.p = vec();
    w.dinv = vec();
    w.partials = c.take<double>(3 * 148 * 8);
    w.scal = c.take<double>(16);
    w.ctl = c.take<CgCtl>(1);
    w.counter = c.take<unsigned int>(4);
    w.total = c.used();
    return w;
}

bool cg_sl_ok(const psfem_csr *A, bool strip);   // defined with the fused-iteration entry points below

struct PcgWs {
    double *r, *p, *q, *dinv, *partials, *scal;
    PcgCtl *ctl;
    unsigned int *counter;
    size_t total;
};

PcgWs carve_pcg(void *ws, int64_t n, int64_t n_blocks) {
    PcgWs w;
    Carver c(ws);
    w.r = c.take<double>(n);
    w.p = c.take<double>(n);
    w.q = c.take<double>(n);
    w.dinv = c.take<double>(n);
    int64_t np = n_blocks > 148 * 8 ? n_blocks : 148 * 8;
    w.partials = c.take<double>(3 * np);
    w.scal = c.take<double>(16);
    w.ctl = c.take<PcgCtl>(1);
    w.counter = c.take<unsigned int>(4);
    w.total = c.used();
    return w;
}

}  // namespace

// =============================================================================================
// C ABI
// =============================================================================================
extern "C" int psfem_spmv_plan_size(int64_t n, int64_t nnz, int64_t *n_blocks) {
    PSFEM_REQUIRE(n_blocks && n >= 0 && nnz >= 0, "spmv_plan_size: bad arguments");
    int64_t b = ceil_div(nnz, SPMV_T);
    *n_blocks = b < 1 ? 1 : b;
    return PSFEM_OK;
}

extern "C" int psfem_spmv_plan(int64_t n, const int32_t *d_row_ptr, int64_t nnz, int32_t *d_rowblk, int64_t n_blocks,
                               int64_t *h_max_row_nnz, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    int64_t want = 0;
    psfem_spmv_plan_size(n, nnz, &want);
    PSFEM_REQUIRE(n_blocks == want && h_max_row_nnz, "spmv_plan: n_blocks must come from psfem_spmv_plan_size");
    spmv_plan_kernel<<<(unsigned)ceil_div(n_blocks + 1, 256), 256, 0, stream>>>(d_row_ptr, n, n_blocks, d_rowblk);
    PSFEM_LAUNCH_CHECK();
    // the max row length is parked in the slot after the plan (d_rowblk has n_blocks+2 entries)
    int *d_max = reinterpret_cast<int *>(d_rowblk + n_blocks + 1);
    PSFEM_CUDA_CHECK(cudaMemsetAsync(d_max, 0, sizeof(int), stream));
    if (n > 0) {
        max_row_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, stream>>>(d_row_ptr, n, d_max);
        PSFEM_LAUNCH_CHECK();
    }
    int h = 0;
    PSFEM_CUDA_CHECK(cudaMemcpyAsync(&h, d_max, sizeof(int), cudaMemcpyDeviceToHost, stream));
    PSFEM_CUDA_CHECK(cudaStreamSynchronize(stream));
    *h_max_row_nnz = h;
    return PSFEM_OK;
}

extern "C" int psfem_spmv_nb_plan_size(int64_t n, int64_t nnz, int64_t *n_blocks) {
    PSFEM_REQUIRE(n_blocks && n >= 0 && nnz >= 0, "spmv_nb_plan_size: bad arguments");
    int64_t b = ceil_div(nnz / 4, NB_PAIRS);
    *n_blocks = b < 1 ? 1 : b;
    return PSFEM_OK;
}

extern "C" int psfem_spmv_nb_plan(int64_t n, const int32_t *d_row_ptr, const int32_t *d_col_idx, int64_t nnz,
                                  int32_t *d_node_ptr, int32_t *d_pcol, int32_t *d_nodeblk, int64_t n_blocks,
                                  int64_t *h_info, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    int64_t want = 0;
    psfem_spmv_nb_plan_size(n, nnz, &want);
    PSFEM_REQUIRE(n_blocks == want && h_info && d_node_ptr && d_pcol && d_nodeblk, "spmv_nb_plan: bad arguments");
    h_info[0] = 0; h_info[1] = 0;
    if (n <= 0 || (n & 1) || (nnz & 3) || nnz == 0) return PSFEM_OK;   // not a node-block matrix
    const int64_t n_nodes = n / 2;
    int *info = reinterpret_cast<int *>(d_nodeblk + n_blocks + 1);    // three scratch ints after the plan
    PSFEM_CUDA_CHECK(cudaMemsetAsync(info, 0, 3 * sizeof(int), stream));
    PSFEM_CUDA_CHECK(cudaMemsetAsync(d_pcol + nnz / 4, 0, 8 * sizeof(int32_t), stream));            // padding read by widened bulk copies
    PSFEM_CUDA_CHECK(cudaMemsetAsync(d_node_ptr + n_nodes + 1, 0, 8 * sizeof(int32_t), stream));
    nb_build_kernel<<<(unsigned)ceil_div(n_nodes + 1, 256), 256, 0, stream>>>(n_nodes, d_row_ptr, d_col_idx, d_node_ptr, d_pcol, info);
    PSFEM_LAUNCH_CHECK();
    nb_plan_kernel<<<(unsigned)ceil_div(n_blocks + 1, 256), 256, 0, stream>>>(d_node_ptr, n_nodes, n_blocks, d_nodeblk);
    PSFEM_LAUNCH_CHECK();
    int h[3] = {1, 0, 0};
    PSFEM_CUDA_CHECK(cudaMemcpyAsync(h, info, sizeof(h), cudaMemcpyDeviceToHost, stream));
    PSFEM_CUDA_CHECK(cudaStreamSynchronize(stream));
    h_info[0] = h[0] == 0 ? 1 : 0;
    h_info[1] = h[1];
    h_info[2] = h[2];
    return PSFEM_OK;
}

extern "C" int psfem_spmv_nb_compress(int64_t n, const int32_t *d_node_ptr, const int32_t *d_pcol, int64_t nnz,
                                      int16_t *d_pcol16, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    PSFEM_REQUIRE(n > 0 && !(n & 1) && d_node_ptr && d_pcol && d_pcol16, "spmv_nb_compress: bad arguments");
    PSFEM_CUDA_CHECK(cudaMemsetAsync(d_pcol16 + nnz / 4, 0, 16 * sizeof(int16_t), stream));
    nb_compress_kernel<<<(unsigned)ceil_div(n / 2, 256), 256, 0, stream>>>(n / 2, d_node_ptr, d_pcol, d_pcol16);
    PSFEM_LAUNCH_CHECK();
    return PSFEM_OK;
}

// ---- symmetric-lattice storage ----------------------------------------------------------------
extern "C" int psfem_spmv_sl_candidates(const psfem_csr *A, int64_t row_node_offset, int64_t x_nodes, int64_t *h_m, int64_t *h_count,
                                        void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    PSFEM_REQUIRE(A && h_m && h_count, "spmv_sl_candidates: null argument");
    *h_count = 0;
    if (!A->nb_node_ptr || !A->nb_pcol || A->n_rows < 4 || (A->n_rows & 1)) return PSFEM_OK;   // needs the 32-bit node-block view
    int32_t np[2] = {0, 0};
    PSFEM_CUDA_CHECK(cudaMemcpyAsync(np, A->nb_node_ptr, sizeof(np), cudaMemcpyDeviceToHost, stream));
    PSFEM_CUDA_CHECK(cudaStreamSynchronize(stream));
    const int deg = np[1] - np[0];
    if (deg < 1 || deg > 9) return PSFEM_OK;
    int32_t pc[9];
    PSFEM_CUDA_CHECK(cudaMemcpyAsync(pc, A->nb_pcol, deg * sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
    PSFEM_CUDA_CHECK(cudaStreamSynchronize(stream));
    // the first row node sits at the bottom of a lattice column: its neighbours in the adjacent columns are m-1, m or
    // m+1 ids away, so every |distance| > 1 proposes m in {d-1, d, d+1}; psfem_spmv_sl_build verifies a proposal
    const int64_t nodes = A->n_rows / 2;
    int64_t cnt = 0;
    // most likely first: the nearest neighbour of the next column is exactly m ids ahead; a last column only has
    // the previous one, whose farthest neighbour is m ids back
    int64_t fwd = 0, bwd = 0;
    for (int k = 0; k < deg; ++k) {
        const int64_t d = (int64_t)pc[k] - row_node_offset;
        if (d > 1 && (fwd == 0 || d < fwd)) fwd = d;
        if (d < -1 && -d > bwd) bwd = -d;
    }
    const int64_t first = fwd ? fwd : bwd;
    if (first >= 2 && nodes % first == 0 && row_node_offset % first == 0 && x_nodes % first == 0) h_m[cnt++] = first;
    for (int k = 0; k < deg; ++k) {
        const int64_t d = pc[k] > row_node_offset ? pc[k] - row_node_offset : row_node_offset - pc[k];
        if (d <= 1) continue;
        for (int64_t mm = d - 1; mm <= d + 1; ++mm) {
            if (mm < 2 || nodes % mm || row_node_offset % mm || x_nodes % mm) continue;
            bool seen = false;
            for (int64_t q = 0; q < cnt; ++q) seen = seen || h_m[q] == mm;
            if (!seen && cnt < 8) h_m[cnt++] = mm;
        }
    }
    *h_count = cnt;
    return PSFEM_OK;
}

extern "C" int psfem_spmv_sl_size(int64_t n_row_nodes, int64_t row_node_offset, int64_t m, size_t *bytes) {
    PSFEM_REQUIRE(bytes && m >= 2 && n_row_nodes > 0 && n_row_nodes % m == 0 && row_node_offset >= 0 && row_node_offset % m == 0,
                  "spmv_sl_size: bad arguments");
    const int64_t S = n_row_nodes / m + (row_node_offset > 0 ? 1 : 0), mpad = (m + 3) & ~3ll;
    *bytes = (size_t)S * SL_NC * (size_t)mpad * sizeof(double);
    return PSFEM_OK;
}

extern "C" int psfem_spmv_sl_build(const psfem_csr *A, int64_t row_node_offset, int64_t x_nodes, int64_t m, double tol,
                                   double *d_sl_vals, size_t bytes, int64_t *h_ok, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    PSFEM_REQUIRE(A && d_sl_vals && h_ok && A->nb_node_ptr && A->nb_pcol && A->vals, "spmv_sl_build: needs a matrix with its 32-bit node-block view");
    const int64_t nodes = A->n_rows / 2;
    size_t want = 0;
    if (int rc = psfem_spmv_sl_size(nodes, row_node_offset, m, &want)) return rc;
    PSFEM_REQUIRE(bytes >= want + 256 && x_nodes % m == 0 && x_nodes >= row_node_offset + nodes && m < (1ll << 30) && x_nodes < (1ll << 31),
                  "spmv_sl_build: bad sizes (buffer must hold psfem_spmv_sl_size bytes + 256)");
    *h_ok = 0;
    int *bad = reinterpret_cast<int *>(reinterpret_cast<char *>(d_sl_vals) + want);   // flag parked after the storage
    PSFEM_CUDA_CHECK(cudaMemsetAsync(d_sl_vals, 0, want + 256, stream));
    const int pre = row_node_offset > 0 ? 1 : 0;
    sl_build_kernel<<<(unsigned)ceil_div(nodes, 128), 128, 0, stream>>>(nodes, row_node_offset, x_nodes, (int)m, (int)((m + 3) & ~3ll), pre,
                                                                      (int)(row_node_offset / m), A->nb_node_ptr, A->nb_pcol, A->vals, tol,
                                                                      d_sl_vals, bad);
    PSFEM_LAUNCH_CHECK();
    int h = 1;
    PSFEM_CUDA_CHECK(cudaMemcpyAsync(&h, bad, sizeof(int), cudaMemcpyDeviceToHost, stream));
    PSFEM_CUDA_CHECK(cudaStreamSynchronize(stream));
    *h_ok = h == 0 ? 1 : 0;
    return PSFEM_OK;
}

extern "C" int psfem_spmv(const psfem_csr *A, const double *d_x, double alpha, double gamma, const double *d_z,
                          double *d_y, void *stream_) {
    return launch_spmv(A, nullptr, d_x, nullptr, alpha, 0.0, gamma, d_z, d_y, nullptr, nullptr, nullptr, nullptr, nullptr,
                       static_cast<cudaStream_t>(stream_));
}

extern "C" int psfem_spmv2(const psfem_csr *A, const double *d_vals2, const double *d_x1, const double *d_x2, double alpha,
                           double beta, double gamma, const double *d_z, double *d_y, void *stream_) {
    PSFEM_REQUIRE(d_vals2 && d_x2, "spmv2: second operand missing");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (A && A->nb_node_ptr && (A->nb_pcol || A->nb_pcol16) && A->nb_n_blocks >= 8 && d_y != d_x1 && d_y != d_x2 &&
        ((reinterpret_cast<uintptr_t>(A->vals) | reinterpret_cast<uintptr_t>(d_vals2) | reinterpret_cast<uintptr_t>(d_x1) |
          reinterpret_cast<uintptr_t>(d_x2)) & 15) == 0) {
        // node-block operands: two passes of the TMA-staged 9-byte kernel (17 B per non-zero pair + 16 B per row for the
        // round trip of y) beat the one-pass CSR kernel (20 B per non-zero, one CTA per row block, 51 % of peak)
        int rc = launch_spmv(A, nullptr, d_x1, nullptr, alpha, 0.0, gamma, d_z, d_y, nullptr, nullptr, nullptr, nullptr, nullptr, stream);
        if (rc) return rc;
        psfem_csr A2 = *A;
        A2.vals = d_vals2;
        return launch_spmv(&A2, nullptr, d_x2, nullptr, beta, 0.0, 1.0, d_y, d_y, nullptr, nullptr, nullptr, nullptr, nullptr, stream);
    }
    return launch_spmv(A, d_vals2, d_x1, d_x2, alpha, beta, gamma, d_z, d_y, nullptr, nullptr, nullptr, nullptr, nullptr, stream);
}

extern "C" int psfem_extract_diag(const psfem_csr *A, int64_t col_offset, int invert, double *d_diag, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    PSFEM_REQUIRE(A && d_diag, "extract_diag: null argument");
    if (A->n_rows == 0) return PSFEM_OK;
    extract_diag_kernel<<<(unsigned)ceil_div(A->n_rows, 256), 256, 0, stream>>>(A->n_rows, A->row_ptr, A->col_idx, A->vals, col_offset, d_diag, invert);
    PSFEM_LAUNCH_CHECK();
    return PSFEM_OK;
}

extern "C" int psfem_pcg_workspace(int64_t n, int64_t n_blocks, size_t *bytes) {
    PSFEM_REQUIRE(bytes && n >= 0, "pcg_workspace: bad arguments");
    const size_t classic = carve_pcg(nullptr, n, n_blocks).total, fused = carve_cg(nullptr, n).total;
    *bytes = classic > fused ? classic : fused;     // either recurrence runs in a work space of this size (pcg_solve picks)
    return PSFEM_OK;
}

// q = A p_ext ; *d_pq = p_own . q
extern "C" int psfem_pcg_spmv_dot(const psfem_csr *A, const double *d_p_ext, int64_t x_offset, double *d_q, double *d_pq,
                                  double *d_partials, uint32_t *d_counter, void *stream_) {
    return launch_spmv(A, nullptr, d_p_ext, nullptr, 1.0, 0.0, 0.0, nullptr, d_q, d_p_ext + x_offset, d_pq, d_partials,
                       d_counter, nullptr, static_cast<cudaStream_t>(stream_));
}

extern "C" int psfem_pcg_init(int64_t n, const double *d_r, const double *d_dinv, const double *d_b, double *d_p,
                              double *d_out3, double *d_partials, uint32_t *d_counter, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    pcg_init_kernel<<<ew_grid(n), EW_THREADS, 0, stream>>>(n, const_cast<double *>(d_r), d_dinv, d_b, d_p, d_out3, d_partials, d_counter, nullptr, 0.0, 0);
    PSFEM_LAUNCH_CHECK();
    return PSFEM_OK;
}

extern "C" int psfem_pcg_update(int64_t n, double *d_r, const double *d_q,
                                const double *d_dinv, const double *d_rz, const double *d_pq, double *d_out2,
                                double *d_partials, uint32_t *d_counter, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    PSFEM_CUDA_CHECK(launch_pdl(pcg_update_kernel<false>, ew_grid(n), EW_THREADS, 0, stream, n, d_r, d_q, d_dinv, d_rz, d_pq, d_out2, d_partials, d_counter, (PcgCtl *)nullptr));
    PSFEM_LAUNCH_CHECK();
    return PSFEM_OK;
}

// the same step for systems that keep constrained DOFs as masked rows (D^-1 = 0 there): their residual stays zero
extern "C" int psfem_pcg_update_masked(int64_t n, double *d_r, const double *d_q,
                                       const double *d_dinv, const double *d_rz, const double *d_pq, double *d_out2,
                                       double *d_partials, uint32_t *d_counter, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    PSFEM_CUDA_CHECK(launch_pdl(pcg_update_kernel<true>, ew_grid(n), EW_THREADS, 0, stream, n, d_r, d_q, d_dinv, d_rz, d_pq, d_out2, d_partials, d_counter, (PcgCtl *)nullptr));
    PSFEM_LAUNCH_CHECK();
    return PSFEM_OK;
}

extern "C" int psfem_pcg_pupdate(int64_t n, double *d_p, double *d_x, const double *d_r, const double *d_dinv,
                                 const double *d_rz_old, const double *d_rz_new, const double *d_pq, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    PSFEM_CUDA_CHECK(launch_pdl(pcg_pupdate_kernel, ew_grid(n), EW_THREADS, 0, stream, n, d_p, d_x, d_r, d_dinv, d_rz_old, d_rz_new, d_pq, (PcgCtl *)nullptr, (unsigned int *)nullptr));
    PSFEM_LAUNCH_CHECK();
    return PSFEM_OK;
}

// ---- peer memory + fused strip-partitioned PCG ------------------------------------------------
extern "C" int psfem_peer_board_bytes(size_t *bytes) {
    PSFEM_REQUIRE(bytes, "peer_board_bytes: null output");
    *bytes = align_up(sizeof(PeerBoard));
    return PSFEM_OK;
}

extern "C" int psfem_peer_alloc(size_t bytes, void **d_ptr, unsigned char *handle64) {
    PSFEM_REQUIRE(d_ptr && handle64 && bytes > 0, "peer_alloc: bad arguments");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    void *p = nullptr;
    PSFEM_CUDA_CHECK(cudaMalloc(&p, bytes));
    cudaError_t e = cudaMemset(p, 0, bytes);
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        set_error("peer_alloc: %s", cudaGetErrorString(e));
        return PSFEM_ECUDA;
    }
    memcpy(handle64, &h, 64);
    *d_ptr = p;
    return PSFEM_OK;
}

extern "C" int psfem_peer_open(const unsigned char *handle64, void **d_ptr) {
    PSFEM_REQUIRE(d_ptr && handle64, "peer_open: bad arguments");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    PSFEM_CUDA_CHECK(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return PSFEM_OK;
}

extern "C" int psfem_peer_close(void *d_ptr) {
    if (d_ptr) PSFEM_CUDA_CHECK(cudaIpcCloseMemHandle(d_ptr));
    return PSFEM_OK;
}

extern "C" int psfem_peer_free(void *d_ptr) {
    if (d_ptr) PSFEM_CUDA_CHECK(cudaFree(d_ptr));
    return PSFEM_OK;
}

static int dist_check(const psfem_dist_ctx *c) {
    PSFEM_REQUIRE(c && c->world >= 1 && c->world <= PSFEM_MAX_RANKS && c->rank >= 0 && c->rank < c->world, "dist_pcg: bad rank/world");
    PSFEM_REQUIRE(c->p_ext && c->x && c->r && c->q && c->dinv && c->partials && c->counter && c->n_own > 0, "dist_pcg: null buffers");
    PSFEM_REQUIRE(!c->has_left || (c->left_p_ext && c->rank > 0), "dist_pcg: left neighbour not mapped");
    PSFEM_REQUIRE(!c->has_right || (c->right_p_ext && c->rank < c->world - 1), "dist_pcg: right neighbour not mapped");
    for (int w = 0; w < c->world; ++w) PSFEM_REQUIRE(c->board[w], "dist_pcg: board of rank %d not mapped", w);
    return PSFEM_OK;
}

extern "C" int psfem_dist_pcg_start(const psfem_dist_ctx *c, const double *d_b, uint64_t seq, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (int rc = dist_check(c)) return rc;
    PSFEM_REQUIRE(d_b && seq >= 1, "dist_pcg_start: bad arguments");
    dist_init_kernel<<<ew_grid(c->n_own), EW_THREADS, 0, stream>>>(c->n_own, d_b, c->dinv, c->x, c->r, c->p_ext + c->own_lo, make_halo(c),
                                                                  make_peer(c, seq), c->partials, c->counter + 2);
    PSFEM_LAUNCH_CHECK();
    return PSFEM_OK;
}

extern "C" int psfem_dist_pcg_iterations(const psfem_csr *A, const psfem_dist_ctx *c, int64_t k, uint64_t seq_first, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (int rc = dist_check(c)) return rc;
    PSFEM_REQUIRE(A && A->n_rows == c->n_own && seq_first >= 2, "dist_pcg_iterations: bad arguments");
    double *p_own = c->p_ext + c->own_lo;
    // with the symmetric-lattice SpMV only its first / last column segment read the halo columns: those warps wait for
    // the neighbours' flags themselves and the p-update kernel ends without waiting (one cross-GPU wait less on the
    // critical path of an iteration); the other SpMV kernels keep the wait at the end of the p-update
    static const bool halo_late_off = getenv("PSFEM_DIST_HALO_EARLY") != nullptr;   // A/B switch
    const bool halo_in_spmv = !halo_late_off && sl_usable(A, false, c->p_ext, c->q, nullptr, p_own, true);
    for (int64_t j = 0; j < k; ++j) {
        const PeerSync ps = make_peer(c, seq_first + (uint64_t)j, halo_in_spmv);
        // q = A p_ext (waits for the neighbours' halo), p.q partial -> all boards
        int rc = launch_spmv(A, nullptr, c->p_ext, nullptr, 1.0, 0.0, 0.0, nullptr, c->q, p_own, c->out3 + 3 /* local p.q, unused */, c->partials,
                             c->counter, nullptr, stream, &ps);
        if (rc) return rc;
        if (c->masked)
            PSFEM_CUDA_CHECK(launch_pdl(dist_update_kernel<true>, ew_grid(c->n_own), EW_THREADS, 0, stream, c->n_own, c->r, (const double *)c->q, (const double *)c->dinv, ps, c->partials,
                                        (unsigned int *)(c->counter + 1)));
        else
            PSFEM_CUDA_CHECK(launch_pdl(dist_update_kernel<false>, ew_grid(c->n_own), EW_THREADS, 0, stream, c->n_own, c->r, (const double *)c->q, (const double *)c->dinv, ps, c->partials,
                                        (unsigned int *)(c->counter + 1)));
        PSFEM_CUDA_CHECK(launch_pdl(dist_pupdate_kernel, ew_grid(c->n_own), EW_THREADS, 0, stream, c->n_own, p_own, c->x, (const double *)c->r, (const double *)c->dinv, make_halo(c), ps,
                                    (unsigned int *)(c->counter + 2)));
    }
    PSFEM_LAUNCH_CHECK();
    return PSFEM_OK;
}

// diagnostic twin of psfem_dist_pcg_iterations: CUDA events around every kernel, mean milliseconds of
// {SpMV+dot, update, p-update} over the k iterations (synchronises)
extern "C" int psfem_dist_pcg_profile(const psfem_csr *A, const psfem_dist_ctx *c, int64_t k, uint64_t seq_first, double *h_ms3,
                                      void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (int rc = dist_check(c)) return rc;
    PSFEM_REQUIRE(A && A->n_rows == c->n_own && seq_first >= 2 && h_ms3 && k > 0 && k <= 4096, "dist_pcg_profile: bad arguments");
    double *p_own = c->p_ext + c->own_lo;
    std::vector<cudaEvent_t> ev(3 * k + 1);
    for (auto &e : ev) PSFEM_CUDA_CHECK(cudaEventCreate(&e));
    PSFEM_CUDA_CHECK(cudaEventRecord(ev[0], stream));
    for (int64_t j = 0; j < k; ++j) {
        const PeerSync ps = make_peer(c, seq_first + (uint64_t)j);
        int rc = launch_spmv(A, nullptr, c->p_ext, nullptr, 1.0, 0.0, 0.0, nullptr, c->q, p_own, c->out3 + 3, c->partials, c->counter, nullptr, stream, &ps);
        if (rc) return rc;
        cudaEventRecord(ev[3 * j + 1], stream);
        if (c->masked) dist_update_kernel<true><<<ew_grid(c->n_own), EW_THREADS, 0, stream>>>(c->n_own, c->r, c->q, c->dinv, ps, c->partials, c->counter + 1);
        else dist_update_kernel<false><<<ew_grid(c->n_own), EW_THREADS, 0, stream>>>(c->n_own, c->r, c->q, c->dinv, ps, c->partials, c->counter + 1);
        cudaEventRecord(ev[3 * j + 2], stream);
        dist_pupdate_kernel<<<ew_grid(c->n_own), EW_THREADS, 0, stream>>>(c->n_own, p_own, c->x, c->r, c->dinv, make_halo(c), ps, c->counter + 2);
        cudaEventRecord(ev[3 * j + 3], stream);
    }
    PSFEM_LAUNCH_CHECK();
    PSFEM_CUDA_CHECK(cudaStreamSynchronize(stream));
    h_ms3[0] = h_ms3[1] = h_ms3[2] = 0.0;
    for (int64_t j = 0; j < k; ++j)
        for (int q = 0; q < 3; ++q) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, ev[3 * j + q], ev[3 * j + q + 1]);
            h_ms3[q] += ms / (double)k;
        }
    for (auto &e : ev) cudaEventDestroy(e);
    return PSFEM_OK;
}

extern "C" int psfem_dist_pcg_residual(const psfem_dist_ctx *c, uint64_t seq, double *h_out3, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (int rc = dist_check(c)) return rc;
    PSFEM_REQUIRE(h_out3 && c->out3, "dist_pcg_residual: null output");
    (void)seq;   // the sums of the last completed step were left in local memory by its last block
    double h[4] = {0, 0, 0, 0};
    PSFEM_CUDA_CHECK(cudaMemcpyAsync(h, c->out3, 3 * sizeof(double), cudaMemcpyDeviceToHost, stream));
    int err = 0;
    PSFEM_CUDA_CHECK(cudaMemcpyAsync(&err, c->counter + 3, sizeof(int), cudaMemcpyDeviceToHost, stream));
    PSFEM_CUDA_CHECK(cudaStreamSynchronize(stream));
    h_out3[0] = h[0]; h_out3[1] = h[1]; h_out3[2] = h[2];
    if (err) { set_error("dist_pcg: a wait on a peer rank timed out"); return PSFEM_ECUDA; }
    return PSFEM_OK;
}

namespace psfem {
// shared with psfem_newmark.cu: enqueue one complete PCG solve (no host sync unless large).
// Small systems: single-CTA kernel.  Large: batches of `check_every` iterations with a device-side
// done flag; the host only polls the flag between batches.
double *pcg_dinv_ptr(void *d_ws, int64_t n, int64_t n_blocks) { return carve_pcg(d_ws, n, n_blocks).dinv; }

// h_info (optional): iterations, relres, ||b|| of THIS solve (forces a sync on the small path).
// d_info: device accumulators of the single-CTA path (sum of iterations, max relres, ||b||, status).
// h_acc (optional): host accumulators of the multi-kernel path (sum of iterations, max relres, status).
int pcg_solve(const psfem_csr *A, const double *d_b, double *d_x, double rtol, int64_t maxit, int64_t check_every,
              double *h_info, double *d_info, double *h_acc, void *d_ws, size_t ws_bytes, bool have_dinv,
              cudaStream_t stream, bool masked, const int32_t *d_new_id) {
    const int64_t n = A->n_rows;
    {
        // operands that carry their symmetric-lattice storage: one kernel per iteration (psfem_cg_sl.cuh)
        const char *mode = getenv("PSFEM_PCG");
        const bool classic = mode && strcmp(mode, "classic") == 0;
        if (!classic && n > SMALL_MAX_N && cg_sl_ok(A, false) && (!masked || d_new_id) && ws_bytes >= carve_cg(nullptr, n).total) {
            double info3[4] = {0, 0, 0, 0};
            const int rc = psfem_cg_sl_solve(A, d_b, d_x, masked ? d_new_id : nullptr, rtol, maxit, check_every, info3, d_ws, ws_bytes, stream);
            if (h_info) { h_info[0] = info3[0]; h_info[1] = info3[1]; h_info[2] = info3[2]; }
            if (h_acc) { h_acc[0] += info3[0]; if (info3[1] > h_acc[1]) h_acc[1] = info3[1]; if (rc == PSFEM_ENOCONV) h_acc[2] = 1.0; }
            return rc;
        }
    }
    PSFEM_REQUIRE(!masked || n > SMALL_MAX_N, "pcg: the masked solve is for systems above %d rows (reduce smaller ones)", SMALL_MAX_N);
    PcgWs w = carve_pcg(d_ws, n, A->n_blocks);
    PSFEM_REQUIRE(ws_bytes >= w.total, "pcg: workspace too small (%zu < %zu)", ws_bytes, w.total);
    if (!have_dinv) {
        int rc = psfem_extract_diag(A, 0, 1, w.dinv, stream);
        if (rc) return rc;
    }
    if (n <= SMALL_MAX_N) {
        pcg_small_kernel<<<1, SMALL_THREADS, 0, stream>>>((int)n, A->row_ptr, A->col_idx, A->vals, d_b, d_x, w.r, w.p, w.q,
                                                          w.dinv, rtol, (long long)maxit, d_info);
        PSFEM_LAUNCH_CHECK();
        if (h_info) {
            double h[4];
            PSFEM_CUDA_CHECK(cudaMemcpyAsync(h, d_info, sizeof(h), cudaMemcpyDeviceToHost, stream));
            PSFEM_CUDA_CHECK(cudaStreamSynchronize(stream));
            h_info[0] = h[0]; h_info[1] = h[1]; h_info[2] = h[2];
            if (h[3] != 0.0) { set_error("pcg: not converged (relres %.3e after %.0f iterations, status %.0f)", h[1], h[0], h[3]); return PSFEM_ENOCONV; }
        }
        return PSFEM_OK;
    }
    PSFEM_CUDA_CHECK(cudaMemsetAsync(w.counter, 0, 4 * sizeof(unsigned int), stream));
    // r = b - A x
    int rc = launch_spmv(A, nullptr, d_x, nullptr, -1.0, 0.0, 1.0, d_b, w.r, nullptr, nullptr, nullptr, nullptr, nullptr, stream);
    if (rc) return rc;
    const int g = ew_grid(n);
    pcg_init_kernel<<<g, EW_THREADS, 0, stream>>>(n, w.r, w.dinv, d_b, w.p, nullptr, w.partials, w.counter, w.ctl, rtol, masked ? 1 : 0);
    PSFEM_LAUNCH_CHECK();
    if (check_every < 1) check_every = 50;
    PcgCtl h{};
    int64_t launched = 0;
    while (true) {
        PSFEM_CUDA_CHECK(cudaMemcpyAsync(&h, w.ctl, sizeof(h), cudaMemcpyDeviceToHost, stream));
        PSFEM_CUDA_CHECK(cudaStreamSynchronize(stream));
        if (h.done || launched >= maxit) break;
        int64_t batch = check_every;
        if (launched + batch > maxit) batch = maxit - launched;
        for (int64_t it = 0; it < batch; ++it) {
            rc = launch_spmv(A, nullptr, w.p, nullptr, 1.0, 0.0, 0.0, nullptr, w.q, w.p, &w.ctl->pq, w.partials, w.counter, &w.ctl->done, stream);
            if (rc) return rc;
            if (masked)
                PSFEM_CUDA_CHECK(launch_pdl(pcg_update_kernel<true>, (unsigned)g, EW_THREADS, 0, stream, n, w.r, (const double *)w.q, (const double *)w.dinv, (const double *)nullptr,
                                            (const double *)nullptr, (double *)nullptr, w.partials, w.counter + 1, w.ctl));
            else
                PSFEM_CUDA_CHECK(launch_pdl(pcg_update_kernel<false>, (unsigned)g, EW_THREADS, 0, stream, n, w.r, (const double *)w.q, (const double *)w.dinv, (const double *)nullptr,
                                            (const double *)nullptr, (double *)nullptr, w.partials, w.counter + 1, w.ctl));
            PSFEM_CUDA_CHECK(launch_pdl(pcg_pupdate_kernel, (unsigned)g, EW_THREADS, 0, stream, n, w.p, d_x, (const double *)w.r, (const double *)w.dinv, (const double *)nullptr,
                                        (const double *)nullptr, (const double *)nullptr, w.ctl, w.counter + 2));
        }
        PSFEM_LAUNCH_CHECK();
        launched += batch;
    }
    const double rel = h.bb > 0 ? sqrt(h.rr / h.bb) : 0.0;
    if (h_info) { h_info[0] = (double)h.iters; h_info[1] = rel; h_info[2] = sqrt(h.bb); }
    if (h_acc) { h_acc[0] += (double)h.iters; if (rel > h_acc[1]) h_acc[1] = rel; }
    if (!(h.rr <= h.tol2) && h.bb > 0) {
        if (h_acc) h_acc[2] = 1.0;
        set_error("pcg: not converged (relres %.3e after %lld iterations)", rel, h.iters);
        return PSFEM_ENOCONV;
    }
    return PSFEM_OK;
}
}  // namespace psfem

extern "C" int psfem_pcg_jacobi(const psfem_csr *A, const double *d_b, double *d_x, double rtol, int64_t maxit,
                                int64_t check_every, double *h_info, void *d_ws, size_t ws_bytes, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    PSFEM_REQUIRE(A && d_b && d_x && h_info, "pcg_jacobi: null argument");
    PSFEM_REQUIRE(A->n_rows > 0, "pcg_jacobi: empty system");
    PcgWs w = carve_pcg(d_ws, A->n_rows, A->n_blocks);
    PSFEM_REQUIRE(ws_bytes >= w.total, "pcg_jacobi: workspace too small (%zu < %zu)", ws_bytes, w.total);
    PSFEM_CUDA_CHECK(cudaMemsetAsync(w.scal, 0, 16 * sizeof(double), stream));
    return psfem::pcg_solve(A, d_b, d_x, rtol, maxit, check_every, h_info, w.scal, nullptr, d_ws, ws_bytes, false, stream, false, nullptr);
}

__global__ void mask_dinv_kernel(int64_t n, const int32_t *__restrict__ new_id, double *__restrict__ dinv) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n && new_id[i] < 0) dinv[i] = 0.0;
}

extern "C" int psfem_pcg_jacobi_masked(const psfem_csr *A, const double *d_b, double *d_x, const int32_t *d_new_id, double rtol,
                                       int64_t maxit, int64_t check_every, double *h_info, void *d_ws, size_t ws_bytes, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    PSFEM_REQUIRE(A && d_b && d_x && d_new_id && h_info, "pcg_jacobi_masked: null argument");
    const int64_t n = A->n_rows;
    PcgWs w = carve_pcg(d_ws, n, A->n_blocks);
    PSFEM_REQUIRE(ws_bytes >= w.total, "pcg_jacobi_masked: workspace too small (%zu < %zu)", ws_bytes, w.total);
    PSFEM_CUDA_CHECK(cudaMemsetAsync(w.scal, 0, 16 * sizeof(double), stream));
    int rc = psfem_extract_diag(A, 0, 1, w.dinv, stream);
    if (rc) return rc;
    mask_dinv_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, stream>>>(n, d_new_id, w.dinv);
    PSFEM_LAUNCH_CHECK();
    return psfem::pcg_solve(A, d_b, d_x, rtol, maxit, check_every, h_info, w.scal, nullptr, d_ws, ws_bytes, true, stream, true, d_new_id);
}

extern "C" int psfem_axpby(int64_t n, double a, const double *d_x, double b, const double *d_y, double *d_out, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (n == 0) return PSFEM_OK;
    axpby_kernel<<<ew_grid(n), EW_THREADS, 0, stream>>>(n, a, d_x, b, d_y, d_out);
    PSFEM_LAUNCH_CHECK();
    return PSFEM_OK;
}

extern "C" int psfem_dot_workspace(int64_t n, size_t *bytes) {
    PSFEM_REQUIRE(bytes, "dot_workspace: null output");
    *bytes = align_up(148 * 8 * sizeof(double)) + 512;
    return PSFEM_OK;
}

extern "C" int psfem_dot(int64_t n, const double *d_x, const double *d_y, double *h_out, void *d_ws, size_t ws_bytes, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    PSFEM_REQUIRE(h_out && d_ws, "dot: null argument");
    Carver c(d_ws);
    double *partials = c.take<double>(148 * 8);
    double *out = c.take<double>(1);
    unsigned int *counter = c.take<unsigned int>(1);
    PSFEM_REQUIRE(ws_bytes >= c.used(), "dot: workspace too small");
    PSFEM_CUDA_CHECK(cudaMemsetAsync(counter, 0, sizeof(unsigned int), stream));
    dot_kernel<<<ew_grid(n), EW_THREADS, 0, stream>>>(n, d_x, d_y, out, partials, counter);
    PSFEM_LAUNCH_CHECK();
    PSFEM_CUDA_CHECK(cudaMemcpyAsync(h_out, out, sizeof(double), cudaMemcpyDeviceToHost, stream));
    PSFEM_CUDA_CHECK(cudaStreamSynchronize(stream));
    return PSFEM_OK;
}

extern "C" int psfem_multi_dot(int64_t m, int64_t n, const double *d_V, const double *d_w, double *d_out, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (m == 0) return PSFEM_OK;
    multi_dot_kernel<<<(unsigned)m, EW_THREADS, 0, stream>>>(n, d_V, d_w, d_out);
    PSFEM_LAUNCH_CHECK();
    return PSFEM_OK;
}

extern "C" int psfem_multi_axpy(int64_t m, int64_t n, const double *d_V, const double *d_c, double *d_w, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (m == 0 || n == 0) return PSFEM_OK;
    multi_axpy_kernel<<<ew_grid(n), EW_THREADS, 0, stream>>>(m, n, d_V, d_c, d_w);
    PSFEM_LAUNCH_CHECK();
    return PSFEM_OK;
}

// =============================================================================================
// fused single-kernel PCG iteration on symmetric-lattice operands (psfem_cg_sl.cuh)
// =============================================================================================
namespace {
bool cg_sl_ok(const psfem_csr *A, bool strip) {
    static const bool off = getenv("PSFEM_NO_SYMLATTICE") != nullptr;
    SlGeom g;
    return !off && A && A->sl_vals && A->sl_src_vals == A->vals && sl_geometry(A, &g, 148, 2) && (strip || A->sl_roff == 0) &&
           (!strip || A->sl_roff == 0 || A->sl_roff == A->sl_m);
}
int cg_sl_launch(const psfem_csr *A, double *d_x, const CgWs &w, bool masked, int64_t it, const PeerSync *peer, const CgVec *halo, cudaStream_t stream) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    SlGeom g;
    PSFEM_REQUIRE(sl_geometry(A, &g, sms, 2), "cg_sl: operand has no symmetric-lattice geometry");
    const int cur = (int)(it & 1);
    CgVec v;
    memset(&v, 0, sizeof(v));
    if (halo) v = *halo;
    v.x = d_x; v.p = w.p;
    v.r_in = w.r[cur]; v.s_in = w.s[cur]; v.w_in = w.w[cur];
    v.r_out = w.r[cur ^ 1]; v.s_out = w.s[cur ^ 1]; v.w_out = w.w[cur ^ 1];
    v.dinv = w.dinv;
    PeerSync ps;
    memset(&ps, 0, sizeof(ps));
    if (peer) ps = *peer;
    static const int defer_x = getenv("PSFEM_CG_NO_DEFER") == nullptr ? 1 : 0;   // A/B switch: x updated every iteration
    // CTAs of four strips; one resident wave of 2 CTAs per SM
    const int n_groups = (int)ceil_div(g.n_strips, CGT_CONSUMERS);
    int64_t G = (int64_t)sms * 2 / n_groups;
    const int64_t owned_cols = g.S - g.pre;
    if (G > owned_cols / 4) G = owned_cols / 4;
    if (G < 1) G = 1;
    g.G = (int)G;
    const unsigned grid = (unsigned)(n_groups * g.G);
    PSFEM_REQUIRE(grid <= 148 * 8, "cg_sl: grid larger than the partial-sum buffer");
    const size_t smem = sizeof(CgStage) * CGT_STAGES + 128;
#define PSFEM_CG_LAUNCH(D, P_)                                                                                                         \
    do {                                                                                                                              \
        PSFEM_CUDA_CHECK(cudaFuncSetAttribute(cg_sl_tma_kernel<D, P_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));        \
        PSFEM_CUDA_CHECK(launch_pdl(cg_sl_tma_kernel<D, P_>, grid, CGT_THREADS, smem, stream, A->sl_vals, g, v, w.ctl, w.partials, w.counter, ps, n_groups, defer_x)); \
    } while (0)
    if (peer) { if (masked) PSFEM_CG_LAUNCH(true, true); else PSFEM_CG_LAUNCH(false, true); }
    else { if (masked) PSFEM_CG_LAUNCH(true, false); else PSFEM_CG_LAUNCH(false, false); }
#undef PSFEM_CG_LAUNCH
    return PSFEM_OK;
}
}  // namespace

extern "C" int psfem_cg_sl_usable(const psfem_csr *A, int *h_usable) {
    PSFEM_REQUIRE(A && h_usable, "cg_sl_usable: null argument");
    *h_usable = cg_sl_ok(A, false) ? 1 : (cg_sl_ok(A, true) ? 2 : 0);   // 2: as a strip of a partition only (psfem_dist_cg_*)
    return PSFEM_OK;
}

extern "C" int psfem_cg_sl_workspace(int64_t n, size_t *bytes) {
    PSFEM_REQUIRE(bytes && n >= 0, "cg_sl_workspace: bad arguments");
    *bytes = carve_cg(nullptr, n).total;
    return PSFEM_OK;
}

extern "C" int psfem_cg_sl_start(const psfem_csr *A, const double *d_b, double *d_x, const int32_t *d_new_id, double rtol,
                                 void *d_ws, size_t ws_bytes, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    PSFEM_REQUIRE(A && d_b && d_x && d_ws, "cg_sl_start: null argument");
    PSFEM_REQUIRE(cg_sl_ok(A, false), "cg_sl_start: the operand needs its symmetric-lattice storage (psfem_spmv_sl_build, rows = columns)");
    const int64_t n = A->n_rows;
    CgWs w = carve_cg(d_ws, n);
    PSFEM_REQUIRE(ws_bytes >= w.total, "cg_sl_start: workspace too small (%zu < %zu)", ws_bytes, w.total);
    PSFEM_CUDA_CHECK(cudaMemsetAsync(w.counter, 0, 4 * sizeof(unsigned int), stream));
    PSFEM_CUDA_CHECK(cudaMemsetAsync(w.scal, 0, 16 * sizeof(double), stream));
    PSFEM_CUDA_CHECK(cudaMemsetAsync(w.ctl, 0, sizeof(CgCtl), stream));
    {
        SlGeom g;
        PSFEM_REQUIRE(sl_geometry(A, &g, 148, 2), "cg_sl_start: operand has no symmetric-lattice geometry");
        cg_sl_dinv_kernel<<<ew_grid(n / 2), EW_THREADS, 0, stream>>>(A->sl_vals, g, w.dinv);
        PSFEM_LAUNCH_CHECK();
    }
    const bool masked = d_new_id != nullptr;
    if (masked) {
        mask_dinv_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, stream>>>(n, d_new_id, w.dinv);
        PSFEM_LAUNCH_CHECK();
    }
    // r_0 = b - A x_0 ; u_0 = D^-1 r_0 (parked in p) ; w_0 = A u_0 ; gamma_0 = r.u, delta_0 = u.w ; s = 0
    int rc = launch_spmv(A, nullptr, d_x, nullptr, -1.0, 0.0, 1.0, d_b, w.r[0], nullptr, nullptr, nullptr, nullptr, nullptr, stream);
    if (rc) return rc;
    pcg_init_kernel<<<ew_grid(n), EW_THREADS, 0, stream>>>(n, w.r[0], w.dinv, d_b, w.p, w.scal, w.partials, w.counter, nullptr, rtol, masked ? 1 : 0);
    PSFEM_LAUNCH_CHECK();
    rc = launch_spmv(A, nullptr, w.p, nullptr, 1.0, 0.0, 0.0, nullptr, w.w[0], w.p, w.scal + 4, w.partials, w.counter + 1, nullptr, stream);
    if (rc) return rc;
    PSFEM_CUDA_CHECK(cudaMemsetAsync(w.s[0], 0, (size_t)n * sizeof(double), stream));
    cg_start_ctl_kernel<<<1, 32, 0, stream>>>(w.ctl, w.scal, w.scal + 4, rtol);
    PSFEM_LAUNCH_CHECK();
    return PSFEM_OK;
}

extern "C" int psfem_cg_sl_iterations(const psfem_csr *A, double *d_x, int masked, int64_t k, int64_t iters_done, void *d_ws,
                                      size_t ws_bytes, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    PSFEM_REQUIRE(A && d_x && d_ws && k >= 0 && iters_done >= 0, "cg_sl_iterations: bad arguments");
    PSFEM_REQUIRE(cg_sl_ok(A, false), "cg_sl_iterations: the operand needs its symmetric-lattice storage");
    CgWs w = carve_cg(d_ws, A->n_rows);
    PSFEM_REQUIRE(ws_bytes >= w.total, "cg_sl_iterations: workspace too small");
    for (int64_t j = 0; j < k; ++j) {
        int rc = cg_sl_launch(A, d_x, w, masked != 0, iters_done + j, nullptr, nullptr, stream);
        if (rc) return rc;
    }
    PSFEM_LAUNCH_CHECK();
    return PSFEM_OK;
}

extern "C" int psfem_cg_sl_state(void *d_ws, int64_t n, double *h_info, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    PSFEM_REQUIRE(d_ws && h_info, "cg_sl_state: null argument");
    CgWs w = carve_cg(d_ws, n);
    CgCtl h;
    PSFEM_CUDA_CHECK(cudaMemcpyAsync(&h, w.ctl, sizeof(h), cudaMemcpyDeviceToHost, stream));
    PSFEM_CUDA_CHECK(cudaStreamSynchronize(stream));
    h_info[0] = (double)h.iters;
    h_info[1] = h.bb > 0 ? sqrt(h.rr / h.bb) : 0.0;
    h_info[2] = sqrt(h.bb);
    h_info[3] = (double)h.done;
    h_info[4] = (double)h.status;
    return PSFEM_OK;
}

extern "C" int psfem_cg_sl_finish(double *d_x, int64_t n, void *d_ws, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    PSFEM_REQUIRE(d_x && d_ws && n >= 0, "cg_sl_finish: bad arguments");
    CgWs w = carve_cg(d_ws, n);
    if (n > 0) {
        cg_finish_x_kernel<<<ew_grid(n), EW_THREADS, 0, stream>>>(n, d_x, w.p, w.ctl);
        cg_clear_pend_kernel<<<1, 1, 0, stream>>>(w.ctl);
        PSFEM_LAUNCH_CHECK();
    }
    return PSFEM_OK;
}

// One solve, then the TRUE residual: the recurrence residual of a Krylov solver drifts away from b - A x over tens of thousands
// of iterations (measured at the 4096x4096 plate: recurrence 9.4e-9, true 1.9e-8, rounding floor of the problem 1.5e-10), so
// after convergence the solve is started again FROM x -- r = b - A x is evaluated by the SpMV, not recurred -- and continued
// until the true residual passes the same test, stagnates (less than a factor 2 per cycle: the fp64 floor of an ill-conditioned
// operand) or `polish` cycles are spent.  A cycle costs two SpMVs plus the few iterations it needs.
// polish < 0: no evaluation (h_info[3] = -1); polish = 0: evaluate only.
// h_info: {iterations (all cycles), recurrence relres at the first convergence, ||b||, true relres at the last evaluation,
//          cycles continued, iterations spent in them}
extern "C" int psfem_cg_sl_solve_checked(const psfem_csr *A, const double *d_b, double *d_x, const int32_t *d_new_id, double rtol,
                                         int64_t maxit, int64_t check_every, int64_t polish, double *h_info, void *d_ws,
                                         size_t ws_bytes, void *stream_) {
    PSFEM_REQUIRE(h_info, "cg_sl_solve: null argument");
    int rc = psfem_cg_sl_start(A, d_b, d_x, d_new_id, rtol, d_ws, ws_bytes, stream_);
    if (rc) return rc;
    if (check_every < 1) check_every = 50;
    int64_t launched = 0, cycles = 0;
    double st[5], total = 0.0, first_relres = -1.0, true_relres = -1.0, prev_true = 0.0, polish_iters = 0.0;
    while (true) {
        while (true) {
            rc = psfem_cg_sl_state(d_ws, A->n_rows, st, stream_);
            if (rc) return rc;
            if (st[3] != 0.0 || launched >= maxit) break;
            int64_t batch = check_every;
            if (launched + batch > maxit) batch = maxit - launched;
            rc = psfem_cg_sl_iterations(A, d_x, d_new_id != nullptr, batch, launched, d_ws, ws_bytes, stream_);
            if (rc) return rc;
            launched += batch;
        }
        rc = psfem_cg_sl_finish(d_x, A->n_rows, d_ws, stream_);
        if (rc) return rc;
        total += st[0];
        if (cycles > 0) polish_iters += st[0];
        if (first_relres < 0.0) first_relres = st[1];
        h_info[0] = total; h_info[1] = first_relres; h_info[2] = st[2];
        h_info[3] = true_relres; h_info[4] = (double)cycles; h_info[5] = polish_iters;
        const bool failed = st[4] != 0.0 || (!(st[1] <= rtol) && st[2] > 0);
        if (failed && cycles == 0) {
            if (st[4] != 0.0) set_error("cg_sl: breakdown (matrix not positive definite?) after %.0f iterations, relres %.3e", total, st[1]);
            else set_error("pcg: not converged (relres %.3e after %.0f iterations)", st[1], total);
            return PSFEM_ENOCONV;
        }
        if (polish < 0) return PSFEM_OK;
        if (failed) polish = 0;     // a continuation cycle ran out of iterations (or broke down at the rounding level): the solve
                                    // had converged before it; evaluate the residual of what is returned and stop
        // r = b - A x from the SpMV; the start leaves the solve ready to continue from x
        rc = psfem_cg_sl_start(A, d_b, d_x, d_new_id, rtol, d_ws, ws_bytes, stream_);
        if (rc) return rc;
        rc = psfem_cg_sl_state(d_ws, A->n_rows, st, stream_);
        if (rc) return rc;
        true_relres = st[1];
        h_info[3] = true_relres;
        const bool stagnated = cycles > 0 && !(true_relres < 0.5 * prev_true);
        if (st[3] != 0.0 || cycles >= polish || stagnated || launched >= maxit) return PSFEM_OK;
        prev_true = true_relres;
        ++cycles;
        // the iteration index of the next launch decides the double buffers: a fresh start begins at an even index
        launched += launched & 1;
    }
}

extern "C" int psfem_cg_sl_solve(const psfem_csr *A, const double *d_b, double *d_x, const int32_t *d_new_id, double rtol, int64_t maxit,
                                 int64_t check_every, double *h_info, void *d_ws, size_t ws_bytes, void *stream_) {
    PSFEM_REQUIRE(h_info, "cg_sl_solve: null argument");
    double h6[6] = {0, 0, 0, 0, 0, 0};
    const int rc = psfem_cg_sl_solve_checked(A, d_b, d_x, d_new_id, rtol, maxit, check_every, -1, h6, d_ws, ws_bytes, stream_);
    h_info[0] = h6[0]; h_info[1] = h6[1]; h_info[2] = h6[2];
    return rc;
}

// ---- the fused iteration on a strip of a multi-GPU partition: halo columns of u as LL words, one scalar exchange ----
namespace {
int dist_cg_check(const psfem_csr *A, const psfem_dist_cg_ctx *c) {
    PSFEM_REQUIRE(A && c && c->world >= 1 && c->world <= PSFEM_MAX_RANKS && c->rank >= 0 && c->rank < c->world, "dist_cg: bad rank/world");
    PSFEM_REQUIRE(cg_sl_ok(A, true), "dist_cg: the strip operand needs its symmetric-lattice storage");
    PSFEM_REQUIRE(c->x && c->counter && c->ll_mine && c->m == A->sl_m, "dist_cg: null buffers / column height mismatch");
    PSFEM_REQUIRE(!c->has_left || (c->ll_left && c->rank > 0 && A->sl_roff == A->sl_m), "dist_cg: left neighbour not mapped (or no halo column in the storage)");
    PSFEM_REQUIRE(!c->has_right || (c->ll_right && c->rank < c->world - 1), "dist_cg: right neighbour not mapped");
    for (int w = 0; w < c->world; ++w) PSFEM_REQUIRE(c->board[w], "dist_cg: board of rank %d not mapped", w);
    return PSFEM_OK;
}
void dist_cg_setup(const psfem_dist_cg_ctx *c, uint64_t seq, PeerSync *ps, CgVec *halo) {
    memset(ps, 0, sizeof(*ps));
    for (int w = 0; w < c->world && w < PSFEM_MAX_RANKS; ++w) ps->board[w] = static_cast<PeerBoard *>(c->board[w]);
    ps->world = c->world; ps->rank = c->rank; ps->has_left = c->has_left; ps->has_right = c->has_right;
    ps->seq = seq;
    ps->err = reinterpret_cast<int *>(c->counter + 3);
    memset(halo, 0, sizeof(*halo));
    unsigned long long *mine = static_cast<unsigned long long *>(c->ll_mine);
    const size_t side = (size_t)c->m * 4;
    halo->ll_from_left = mine;                  // side 0 of my buffer: written by my left neighbour
    halo->ll_from_right = mine + side;          // side 1: written by my right neighbour
    if (c->has_left) halo->ll_to_left = static_cast<unsigned long long *>(c->ll_left) + side;   // I am its right neighbour
    if (c->has_right) halo->ll_to_right = static_cast<unsigned long long *>(c->ll_right);       // I am its left neighbour
}
}  // namespace

extern "C" int psfem_dist_cg_halo_bytes(int64_t m, size_t *bytes) {
    PSFEM_REQUIRE(bytes && m > 0, "dist_cg_halo_bytes: bad arguments");
    *bytes = align_up((size_t)m * 4 * sizeof(unsigned long long) * 2);
    return PSFEM_OK;
}

// x = 0, r = b, then ONE launch of the iteration kernel in start mode: u_0 = D^-1 b, w_0 = A u_0 (halo exchanged inside the
// kernel), gamma_0, delta_0, b.b, alpha_0.  Launch 0 of the solve, sequence number seq; the iterations follow with
// launches_done = 1, 2, ... and seq + 1, seq + 2, ...
extern "C" int psfem_dist_cg_start(const psfem_csr *A, const psfem_dist_cg_ctx *c, const double *d_b, const int32_t *d_new_id,
                                   double rtol, uint64_t seq, void *d_ws, size_t ws_bytes, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (int rc = dist_cg_check(A, c)) return rc;
    PSFEM_REQUIRE(d_b && d_ws && seq >= 1, "dist_cg_start: bad arguments");
    PSFEM_REQUIRE(!c->masked || d_new_id, "dist_cg_start: masked strips pass the free map of their owned rows (negative = constrained)");
    const int64_t n = A->n_rows;
    CgWs w = carve_cg(d_ws, n);
    PSFEM_REQUIRE(ws_bytes >= w.total, "dist_cg_start: workspace too small (%zu < %zu)", ws_bytes, w.total);
    const size_t nb = (size_t)n * sizeof(double);
    PSFEM_CUDA_CHECK(cudaMemsetAsync(w.counter, 0, 3 * sizeof(unsigned int), stream));
    PSFEM_CUDA_CHECK(cudaMemcpyAsync(w.r[0], d_b, nb, cudaMemcpyDeviceToDevice, stream));
    PSFEM_CUDA_CHECK(cudaMemsetAsync(w.w[0], 0, nb, stream));
    PSFEM_CUDA_CHECK(cudaMemsetAsync(w.s[0], 0, nb, stream));
    PSFEM_CUDA_CHECK(cudaMemsetAsync(w.p, 0, nb, stream));
    PSFEM_CUDA_CHECK(cudaMemsetAsync(c->x, 0, nb, stream));
    if (c->masked) {   // D^-1 from the lattice diagonal, zero at the constrained rows
        SlGeom g;
        PSFEM_REQUIRE(sl_geometry(A, &g, 148, 2), "dist_cg_start: operand has no symmetric-lattice geometry");
        cg_sl_dinv_kernel<<<ew_grid(n / 2), EW_THREADS, 0, stream>>>(A->sl_vals, g, w.dinv);
        mask_dinv_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, stream>>>(n, d_new_id, w.dinv);
        PSFEM_LAUNCH_CHECK();
    }
    cg_start_pass_ctl_kernel<<<1, 32, 0, stream>>>(w.ctl, rtol);
    PSFEM_LAUNCH_CHECK();
    PeerSync ps;
    CgVec halo;
    dist_cg_setup(c, seq, &ps, &halo);
    // the kernel's own counter lives in the workspace; the error flag in the caller's counter block
    return cg_sl_launch(A, c->x, w, c->masked != 0, 0, &ps, &halo, stream);
}

// y = A x over the strips (x, y: owned rows; the halo columns of x travel inside the kernel like the halo of u): one launch
// of the iteration kernel in apply mode, sequence number seq.  For the TRUE residual of a finished strip solve; the vectors
// of the recurrence are overwritten (a psfem_dist_cg_start follows), the control block keeps the state of the solve.
extern "C" int psfem_dist_cg_apply(const psfem_csr *A, const psfem_dist_cg_ctx *c, const double *d_x, double *d_y, uint64_t seq,
                                   void *d_ws, size_t ws_bytes, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (int rc = dist_cg_check(A, c)) return rc;
    PSFEM_REQUIRE(d_x && d_y && d_ws && seq >= 2, "dist_cg_apply: bad arguments (an apply pass follows a solve)");
    const int64_t n = A->n_rows;
    CgWs w = carve_cg(d_ws, n);
    PSFEM_REQUIRE(ws_bytes >= w.total, "dist_cg_apply: workspace too small (%zu < %zu)", ws_bytes, w.total);
    const size_t nb = (size_t)n * sizeof(double);
    PSFEM_CUDA_CHECK(cudaMemsetAsync(w.counter, 0, 3 * sizeof(unsigned int), stream));
    PSFEM_CUDA_CHECK(cudaMemcpyAsync(w.r[0], d_x, nb, cudaMemcpyDeviceToDevice, stream));
    cg_apply_ctl_kernel<<<1, 32, 0, stream>>>(w.ctl);
    PSFEM_LAUNCH_CHECK();
    PeerSync ps;
    CgVec halo;
    dist_cg_setup(c, seq, &ps, &halo);
    if (int rc = cg_sl_launch(A, c->x, w, c->masked != 0, 0, &ps, &halo, stream)) return rc;
    PSFEM_CUDA_CHECK(cudaMemcpyAsync(d_y, w.w[1], nb, cudaMemcpyDeviceToDevice, stream));
    return PSFEM_OK;
}

extern "C" int psfem_dist_cg_iterations(const psfem_csr *A, const psfem_dist_cg_ctx *c, int64_t k, int64_t launches_done, uint64_t seq_first,
                                        void *d_ws, size_t ws_bytes, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (int rc = dist_cg_check(A, c)) return rc;
    PSFEM_REQUIRE(d_ws && k >= 0 && launches_done >= 1 && seq_first >= 2, "dist_cg_iterations: bad arguments");
    CgWs w = carve_cg(d_ws, A->n_rows);
    PSFEM_REQUIRE(ws_bytes >= w.total, "dist_cg_iterations: workspace too small");
    for (int64_t j = 0; j < k; ++j) {
        PeerSync ps;
        CgVec halo;
        dist_cg_setup(c, seq_first + (uint64_t)j, &ps, &halo);
        int rc = cg_sl_launch(A, c->x, w, c->masked != 0, launches_done + j, &ps, &halo, stream);
        if (rc) return rc;
    }
    PSFEM_LAUNCH_CHECK();
    return PSFEM_OK;
}

// {iterations, relative residual, ||b||, done, status} of the strip solve (global values, identical on every rank) and the
// peer time-out flag; synchronises
extern "C" int psfem_dist_cg_state(const psfem_dist_cg_ctx *c, int64_t n, void *d_ws, double *h_info, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    PSFEM_REQUIRE(c && d_ws && h_info && c->counter, "dist_cg_state: null argument");
    int rc = psfem_cg_sl_state(d_ws, n, h_info, stream_);
    if (rc) return rc;
    int err = 0;
    PSFEM_CUDA_CHECK(cudaMemcpyAsync(&err, c->counter + 3, sizeof(int), cudaMemcpyDeviceToHost, stream));
    PSFEM_CUDA_CHECK(cudaStreamSynchronize(stream));
    if (err) {
        static const char *site[4] = {"a wait", "the wait for the left neighbour's halo column", "the wait for the right neighbour's halo column",
                                      "the wait for the scalars"};
        set_error("dist_cg: %s%s timed out (code 0x%x: source rank %d, sequence number %d mod 2^20)", site[err & 3],
                  (err & 3) == 3 ? " of a rank" : "", err, (err >> 4) & 15, (err >> 8) & 0xfffff);
        return PSFEM_ECUDA;
    }
    return PSFEM_OK;
}
