// tma.cuh - the Tensor Memory Accelerator and mbarrier pieces shared by the kernels that use them (thin.cu: TMA-fed
// step kernel; ksd.cu: tensor-core KSD).  Device side: thin wrappers over the PTX (the CPU emulator of the test suite
// models the same operations, tests/host/simt).  Host side: a 2-D SWIZZLE_128B tensor map through
// cuTensorMapEncodeTiled, fetched with cudaGetDriverEntryPoint so the library keeps no link-time dependency on libcuda.
#pragma once

#if !defined(GFK_SIMT_HOST)
#include <cuda.h>      // CUtensorMap and its enums only
#endif
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace gfk {

#if defined(GFK_SIMT_HOST)
using TensorMap = simt::TensorMap2D;
__device__ __forceinline__ void dev_mbar_init(unsigned long long* bar, int count) { simt::mbar_init(bar, count); }
__device__ __forceinline__ void dev_mbar_fence_init() { }
__device__ __forceinline__ void dev_mbar_arrive(unsigned long long* bar) { simt::mbar_arrive(bar); }
__device__ __forceinline__ void dev_mbar_arrive_expect_tx(unsigned long long* bar, unsigned bytes) { simt::mbar_arrive_expect_tx(bar, bytes); }
__device__ __forceinline__ bool dev_mbar_try(unsigned long long* bar, unsigned parity) {
    if (simt::mbar_test(bar, parity)) return true;
    simt::poll_yield();
    return false;
}
// emulated CTAs are slow: 120 s, or SIMT_MBAR_LIMIT_MS for the fault-injection tests
static inline unsigned long long mbar_wait_limit_ns() {
    static unsigned long long v = 0;
    if (!v) {
        const char* e = getenv("SIMT_MBAR_LIMIT_MS");
        v = (e && atof(e) > 0.0) ? (unsigned long long)(atof(e) * 1e6) : 120ull * 1000000000ull;
    }
    return v;
}
#define kMbarWaitLimitNs (gfk::mbar_wait_limit_ns())
__device__ __forceinline__ void dev_tma_load_2d(void* dst, const TensorMap* map, int c0, int c1, unsigned long long* bar) {
    simt::tma_load_2d(dst, map, c0, c1, bar);
}
#else
using TensorMap = CUtensorMap;
__device__ __forceinline__ void dev_mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
// makes the initialised barriers visible to the async proxy (the copy engine) before the first copy names them
__device__ __forceinline__ void dev_mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void dev_mbar_arrive(unsigned long long* bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}"
                 ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void dev_mbar_arrive_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}"
                 ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool dev_mbar_try(unsigned long long* bar, unsigned parity) {
    unsigned ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}" : "=r"(ok) : "r"((unsigned)__cvta_generic_to_shared(bar)), "r"(parity) : "memory");
    return ok != 0u;
}
constexpr unsigned long long kMbarWaitLimitNs = 2ull * 1000000000ull;
__device__ __forceinline__ void dev_tma_load_2d(void* dst, const TensorMap* map, int c0, int c1, unsigned long long* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(reinterpret_cast<unsigned long long>(map)),
                   "r"((unsigned)__cvta_generic_to_shared(bar)), "r"(c0), "r"(c1) : "memory");
}
#endif

// Wait for the phase with this parity, BOUNDED: these barriers only connect warps of one CTA (and its copy / tensor
// engines), so a wait that lasts seconds means a defect, not a slow peer - the waiter then raises *abort_flag (device
// memory; the host entry's check call reports it as GFK_ECUDA) and every later wait of the kernel returns at once, so
// the kernel always terminates instead of hanging the GPU.  Returns false when the launch is aborted: the caller leaves
// its loop (no further arrivals or copies on barriers whose phases will never complete).  Results of an aborted launch
// are garbage, never addresses.
__device__ __forceinline__ bool dev_mbar_wait(unsigned long long* bar, unsigned parity, unsigned* abort_flag) {
    if (dev_mbar_try(bar, parity)) return true;
    unsigned long long t0 = 0;
    unsigned tries = 0;
#pragma unroll 1
    for (;;) {
        if (dev_mbar_try(bar, parity)) return true;
        if (*reinterpret_cast<volatile unsigned*>(abort_flag)) return false;   // slow path only: the first try already failed
        if ((++tries & 255u) == 0u) {
            const unsigned long long now = gfk_timer_ns();
            if (t0 == 0) t0 = now;
            else if (now - t0 > kMbarWaitLimitNs) { *reinterpret_cast<volatile unsigned*>(abort_flag) = 1u; return false; }
        }
    }
}


// 2-D tensor map over a row-major array of `elem`-byte floating-point elements: dim0 (fastest) x dim1, rows dim0 * elem
// bytes apart (a multiple of 16), box = box0 x box1 elements with box0 * elem <= 128, CU_TENSOR_MAP_SWIZZLE_128B,
// out-of-bounds elements read as zero.  false: the driver entry point is missing or refused the layout.
static inline bool make_tensor_map_2d(TensorMap* map, const void* base, unsigned elem, unsigned long long dim0,
                                      unsigned long long dim1, unsigned box0, unsigned box1) {
#if defined(GFK_SIMT_HOST)
    memset(map, 0, sizeof(*map));
    map->magic = 0x7a3du;
    map->elem = elem;
    map->box0 = box0;
    map->box1 = box1;
    map->swizzle = 128;
    map->base = base;
    map->dim0 = dim0;
    map->dim1 = dim1;
    map->stride1 = dim0 * elem;
    return true;
#else
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn encode = nullptr;
    static bool looked = false;
    if (!looked) {
        looked = true;
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            encode = reinterpret_cast<EncodeFn>(fn);
        else
            cudaGetLastError();
    }
    if (!encode) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)dim0, (cuuint64_t)dim1};
    const cuuint64_t strides[1] = {(cuuint64_t)dim0 * elem};
    const cuuint32_t box[2] = {(cuuint32_t)box0, (cuuint32_t)box1};
    const cuuint32_t estr[2] = {1, 1};
    const CUtensorMapDataType dt = elem == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
    return encode(map, dt, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
#endif
}

}  // namespace gfk
