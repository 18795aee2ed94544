// psfem_core.cu -- version, error text, mesh generation (Polygones.py:62-108).
#include "psfem_common.cuh"

namespace psfem {
static thread_local char g_err[1024] = "";
void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
}  // namespace psfem

using namespace psfem;

extern "C" int psfem_version(void) { return 100; }
extern "C" const char *psfem_last_error(void) { return psfem::g_err; }
extern "C" int psfem_set_device(int device) {
    PSFEM_CUDA_CHECK(cudaSetDevice(device));
    return PSFEM_OK;
}
extern "C" int psfem_nodes_per_element(int etype) {
    switch (etype) {
        case PSFEM_T3: return 3;
        case PSFEM_Q4: return 4;
        case PSFEM_T6: return 6;
        case PSFEM_Q9: return 9;
    }
    return -1;
}

// node k = i*m + j at [i*L/(n-1) - offX, j*l/(m-1) - offY]   (Polygones.py:77)
// The product, the IEEE division and the subtraction are three separately rounded operations,
// as in the Python expression; __dmul_rn/__ddiv_rn/__dsub_rn forbid FMA contraction.
__global__ void mesh_nodes_kernel(double2 *__restrict__ nodes, int64_t count, int64_t m, int64_t col0,
                                  double L, double l, double nm1, double mm1, double offX, double offY) {
    int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k >= count) return;
    int64_t i = col0 + k / m, j = k % m;
    double x = __dsub_rn(__ddiv_rn(__dmul_rn((double)i, L), nm1), offX);
    double y = __dsub_rn(__ddiv_rn(__dmul_rn((double)j, l), mm1), offY);
    nodes[k] = make_double2(x, y);
}

// connectivity, Polygones.py:83,85 (T3) :91,93 (T6) :100 (Q4) :106 (Q9); ids local to the strip
template <int ETYPE>
__global__ void mesh_conn_kernel(int32_t *__restrict__ conn, int64_t ncells_total, int64_t M, int64_t m,
                                 int64_t cell0, int64_t node_shift) {
    int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (c >= ncells_total) return;
    int64_t i = cell0 + c / M, j = c % M;
    if (ETYPE == PSFEM_T3) {
        int32_t b = (int32_t)(j + i * m - node_shift);
        int32_t mm = (int32_t)m;
        int32_t *o = conn + c * 6;
        o[0] = b; o[1] = b + mm; o[2] = b + 1;
        o[3] = b + mm + 1; o[4] = b + 1; o[5] = b + mm;
    } else if (ETYPE == PSFEM_Q4) {
        int32_t b = (int32_t)(j + i * m - node_shift);
        int32_t mm = (int32_t)m;
        reinterpret_cast<int4 *>(conn)[c] = make_int4(b, b + mm, b + mm + 1, b + 1);
    } else if (ETYPE == PSFEM_T6) {
        int32_t q = (int32_t)(2 * j + 2 * i * m - node_shift);
        int32_t mm = (int32_t)m;
        int32_t *o = conn + c * 12;
        o[0] = q; o[1] = q + mm; o[2] = q + 2 * mm; o[3] = q + mm + 1; o[4] = q + 2; o[5] = q + 1;
        o[6] = q + 2 * mm + 2; o[7] = q + mm + 2; o[8] = q + 2; o[9] = q + mm + 1; o[10] = q + 2 * mm; o[11] = q + 2 * mm + 1;
    } else {
        int32_t q = (int32_t)(2 * j + 2 * i * m - node_shift);
        int32_t mm = (int32_t)m;
        int32_t *o = conn + c * 9;
        o[0] = q; o[1] = q + mm; o[2] = q + 2 * mm; o[3] = q + 2 * mm + 1; o[4] = q + 2 * mm + 2;
        o[5] = q + mm + 2; o[6] = q + 2; o[7] = q + 1; o[8] = q + mm + 1;
    }
}

extern "C" int psfem_mesh_rect(double L, double l, int64_t N, int64_t M, double offX, double offY, int etype,
                               int64_t col0, int64_t ncols, int64_t cell0, int64_t ncells,
                               double *d_nodes, int32_t *d_conn, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    PSFEM_REQUIRE(N > 0, "mesh: N must be positive");
    if (M == 0) M = N;  // Polygones.py:64-65
    PSFEM_REQUIRE(M > 0, "mesh: M must be positive");
    PSFEM_REQUIRE(etype >= PSFEM_T3 && etype <= PSFEM_Q9, "mesh: unknown element kind %d", etype);
    const bool fine = (etype == PSFEM_T6 || etype == PSFEM_Q9);
    const int64_t n = fine ? 2 * N + 1 : N + 1, m = fine ? 2 * M + 1 : M + 1;  // Polygones.py:66-71
    PSFEM_REQUIRE(col0 >= 0 && ncols >= 0 && col0 + ncols <= n, "mesh: node column range out of bounds");
    PSFEM_REQUIRE(cell0 >= 0 && ncells >= 0 && cell0 + ncells <= N, "mesh: cell column range out of bounds");
    PSFEM_REQUIRE(ncols * m < (int64_t)INT32_MAX, "mesh: strip has too many nodes for int32 ids");
    const int T = 256;
    if (d_nodes && ncols > 0) {
        int64_t count = ncols * m;
        mesh_nodes_kernel<<<(unsigned)ceil_div(count, T), T, 0, stream>>>(
            reinterpret_cast<double2 *>(d_nodes), count, m, col0, L, l, (double)(n - 1), (double)(m - 1), offX, offY);
        PSFEM_LAUNCH_CHECK();
    }
    if (d_conn && ncells > 0) {
        int64_t cells = ncells * M;
        int64_t shift = col0 * m;
        unsigned grid = (unsigned)ceil_div(cells, T);
        switch (etype) {
            case PSFEM_T3: mesh_conn_kernel<PSFEM_T3><<<grid, T, 0, stream>>>(d_conn, cells, M, m, cell0, shift); break;
            case PSFEM_Q4: mesh_conn_kernel<PSFEM_Q4><<<grid, T, 0, stream>>>(d_conn, cells, M, m, cell0, shift); break;
            case PSFEM_T6: mesh_conn_kernel<PSFEM_T6><<<grid, T, 0, stream>>>(d_conn, cells, M, m, cell0, shift); break;
            default: mesh_conn_kernel<PSFEM_Q9><<<grid, T, 0, stream>>>(d_conn, cells, M, m, cell0, shift); break;
        }
        PSFEM_LAUNCH_CHECK();
    }
    return PSFEM_OK;
}
