// psfem_common.cuh -- shared helpers of libpsfem.so (error text, launch checks, reductions).
#pragma once
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include "psfem.h"

namespace psfem {

void set_error(const char *fmt, ...);

#define PSFEM_CUDA_CHECK(expr)                                                              \
    do {                                                                                    \
        cudaError_t e_ = (expr);                                                            \
        if (e_ != cudaSuccess) {                                                            \
            ::psfem::set_error("%s:%d %s: %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e_)); \
            return PSFEM_ECUDA;                                                             \
        }                                                                                   \
    } while (0)

#define PSFEM_LAUNCH_CHECK() PSFEM_CUDA_CHECK(cudaGetLastError())

#define PSFEM_REQUIRE(cond, ...)                                                            \
    do {                                                                                    \
        if (!(cond)) {                                                                      \
            ::psfem::set_error(__VA_ARGS__);                                                \
            return PSFEM_EINVAL;                                                            \
        }                                                                                   \
    } while (0)

static inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }
static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// carve a caller-provided workspace
struct Carver {
    char *base;
    size_t off = 0;
    explicit Carver(void *p) : base(static_cast<char *>(p)) {}
    template <class T> T *take(size_t count) {
        T *p = reinterpret_cast<T *>(base + off);
        off += align_up(count * sizeof(T));
        return p;
    }
    size_t used() const { return off; }
};

// ---- deterministic block reduction (fixed shuffle tree, fixed warp order) ----
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}

template <int THREADS> __device__ __forceinline__ double block_sum(double v, double *smem /* >= THREADS/32 */) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    v = warp_sum(v);
    __syncthreads();  // smem may still be read from a previous call
    if (lane == 0) smem[wid] = v;
    __syncthreads();
    double r = 0.0;
    if (wid == 0) {
        r = (lane < THREADS / 32) ? smem[lane] : 0.0;
        r = warp_sum(r);
    }
    return r;  // valid in thread 0
}

}  // namespace psfem
