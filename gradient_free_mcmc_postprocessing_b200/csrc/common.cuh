// common.cuh - error plumbing and small device helpers shared by the libgfk translation units.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/gfk.h"
#include "stein_math.cuh"

namespace gfk {

// thread-local last-error message (gfk_last_error)
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);

#define GFK_CUDA(call)                                                          \
    do {                                                                        \
        cudaError_t e__ = (call);                                               \
        if (e__ != cudaSuccess) return gfk::cuda_fail(e__, #call, __FILE__, __LINE__); \
    } while (0)

#define GFK_CHECK_LAUNCH() GFK_CUDA(cudaGetLastError())

#define GFK_REQUIRE(cond, ...)            \
    do {                                  \
        if (!(cond)) {                    \
            gfk::set_error(__VA_ARGS__);  \
            return GFK_EINVAL;            \
        }                                 \
    } while (0)

// dynamic shared memory of the running block as `unsigned char* name` (headers are not run through the test suite's
// launch-syntax rewriter, so kernels that live in a .cuh declare it through this)
#if defined(GFK_SIMT_HOST)
#define GFK_DYN_SMEM(name) unsigned char* name = reinterpret_cast<unsigned char*>(simt::dyn_smem())
#else
#define GFK_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#endif

// nanosecond wall clock of the device (%globaltimer); every bounded wait of the library measures against it
__device__ __forceinline__ unsigned long long gfk_timer_ns() {
#if defined(GFK_SIMT_HOST)
    return simt::now_ns();
#else
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t) :: "memory");
    return t;
#endif
}


int sm_count();
// running count of kernel launches issued by this library (gfk_launch_count)
void count_launches(long long n);

// ---- warp / block argmin with numpy.argmin semantics -----------------------------------------
__device__ __forceinline__ void warp_argmin(double& v, unsigned& i) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, v, off);
        const unsigned oi = __shfl_xor_sync(0xffffffffu, i, off);
        if (cand_better(ov, oi, v, i)) { v = ov; i = oi; }
    }
}

// all threads call; result valid in thread 0.  smem: 32 doubles + 32 unsigned.
__device__ __forceinline__ void block_argmin(double& v, unsigned& i, double* sv, unsigned* si) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    warp_argmin(v, i);
    if (lane == 0) { sv[warp] = v; si[warp] = i; }
    __syncthreads();
    if (warp == 0) {
        const int nwarps = (blockDim.x + 31) >> 5;
        v = (lane < nwarps) ? sv[lane] : __longlong_as_double(0x7ff0000000000000LL);
        i = (lane < nwarps) ? si[lane] : 0xffffffffu;
        warp_argmin(v, i);
    }
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}

}  // namespace gfk
