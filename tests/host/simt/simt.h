// simt.h - a small SIMT emulator so the CPU test suite can run the REAL kernel sources of libgfk
// (csrc/*.cu) without a GPU.  TEST INFRASTRUCTURE ONLY: nothing in the product package loads or links it.
//
// The .cu files are compiled as plain C++ (after tests/host/simt/cu2cpp.py rewrites the launch syntax):
//   * every CUDA thread is a fiber with its own stack; a block's fibers run on one OS thread,
//     blocks of a cooperative launch run concurrently on one OS thread each (spin-waits between blocks work);
//   * __syncthreads / __shfl_*_sync / __any_sync are rendezvous points between fibers; a barrier that can never
//     complete (divergent __syncthreads, missing lanes) is reported as a deadlock instead of hanging;
//   * the fiber order between rendezvous points is round-robin or, with SIMT_SCHED=random[:seed], random -
//     a missing __syncthreads then shows up as a wrong result;
//   * dynamic shared memory is poisoned before every block, vector loads and the 16-byte slot accesses check
//     their alignment (the misaligned-address class of faults);
//   * "device memory" is host memory; the handful of CUDA runtime calls the library makes are implemented
//     synchronously in simt.cpp.
#pragma once

#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <functional>

// ---- CUDA keywords -------------------------------------------------------------------------------------
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__ __attribute__((noinline))
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))
#define __shared__ static thread_local
#define __grid_constant__

// ---- vector types ------------------------------------------------------------------------------------------
struct uint3 { unsigned x, y, z; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
struct __attribute__((aligned(16))) double2 { double x, y; };
struct __attribute__((aligned(8))) float2 { float x, y; };
struct __attribute__((aligned(16))) float4 { float x, y, z, w; };
struct __attribute__((aligned(16))) uint4 { unsigned x, y, z, w; };
inline uint4 make_uint4(unsigned a, unsigned b, unsigned c, unsigned d) { uint4 v; v.x = a; v.y = b; v.z = c; v.w = d; return v; }
inline double2 make_double2(double a, double b) { double2 v; v.x = a; v.y = b; return v; }

namespace simt {

struct ThreadCtx {
    uint3 tid, bid, bdim, gdim;
};
ThreadCtx* cur();                       // the running fiber
void* dyn_smem();                       // this block's dynamic shared memory
void syncthreads(const char* file, int line);
// PTX bar.sync / bar.arrive on named barrier `id` (1..15) expecting `count` threads; wait: bar.sync
void named_barrier(int id, int count, bool wait);
// every lane of the warp contributes 64 bits; returns the 32 contributions (valid until the next collective)
const unsigned long long* warp_gather(unsigned long long bits);
unsigned warp_alive_mask();
void poll_yield();                      // inside spin loops on memory another thread writes
unsigned long long now_ns();            // %globaltimer: monotonic wall clock in nanoseconds
// cp.async with its real (deferred) semantics: the bytes move no earlier than the cp_async_wait that covers their
// group, so reading a staged buffer before waiting - or before the barrier that publishes OTHER threads' waits -
// finds the poison pattern instead of data
// tensor memory (sm_100a): 512 columns x 128 lanes x 32 bit per CTA; tcgen05.alloc / dealloc by one full warp,
// tcgen05.st / ld.32x32b.xN by full, converged warps that may only touch lanes 32 (warp % 4) .. + 31
void tmem_alloc(unsigned* dst_smem, int cols);
void tmem_dealloc(unsigned base, int cols);
void tmem_st(unsigned taddr, const unsigned* regs, int n);
void tmem_ld(unsigned taddr, unsigned* regs, int n);
void cp_async(void* dst, const void* src, int bytes);
void cp_async_commit();
void cp_async_wait(int allow_pending_groups);
[[noreturn]] void fail(const char* fmt, ...);
void check_align(const void* p, size_t a, const char* what);

// ---- mbarrier + TMA (cp.async.bulk.tensor.2d) model ------------------------------------------------------------
// What the emulator's stand-in for cuTensorMapEncodeTiled fills in: a 2-D row-major tensor of `elem`-byte elements,
// dim0 fastest, rows stride1 bytes apart; loads move a box0 x box1 box, out-of-bounds elements read as zero,
// swizzle = 128: CU_TENSOR_MAP_SWIZZLE_128B (16-byte chunk index XOR (128-byte row index mod 8)).
struct __attribute__((aligned(64))) TensorMap2D {
    unsigned magic;
    unsigned elem, box0, box1, swizzle;
    const void* base;
    unsigned long long dim0, dim1, stride1;
    char pad[128 - 56];
};
static_assert(sizeof(TensorMap2D) == 128, "same size as CUtensorMap");
// an mbarrier object is the 8 bytes in shared memory it is on the device; a copy issued by tma_load_2d stays pending
// (its destination poisoned) until some thread of the block next tests a barrier, then completes and reports its bytes
void mbar_init(unsigned long long* bar, int count);
void mbar_arrive(unsigned long long* bar);
void mbar_arrive_expect_tx(unsigned long long* bar, unsigned bytes);
bool mbar_test(unsigned long long* bar, unsigned parity);
void tma_load_2d(void* dst, const TensorMap2D* map, int c0, int c1, unsigned long long* bar);
// shared-window address (cvta.to.shared) of a pointer into dynamic shared memory; tcgen05.mma kind::tf32 with both
// operands K-major SWIZZLE_128B in shared memory and the accumulator in tensor memory; tcgen05.commit
unsigned smem_u32(const void* p);
void umma_tf32(unsigned tmem_d, unsigned long long desc_a, unsigned long long desc_b, unsigned idesc, bool accumulate);
void umma_commit(unsigned long long* bar);

// grid x block fibers run fn(); cooperative: all blocks concurrently
void launch(dim3 grid, dim3 block, size_t smem, bool cooperative, const std::function<void()>& fn);

}  // namespace simt

#define threadIdx (simt::cur()->tid)
#define blockIdx (simt::cur()->bid)
#define blockDim (simt::cur()->bdim)
#define gridDim (simt::cur()->gdim)

#define SIMT_LAUNCH(kernel, grid, block, smem, stream, ...) \
    simt::launch(dim3(grid), dim3(block), (size_t)(smem), false, [=]() { kernel(__VA_ARGS__); })

// ---- intrinsics ------------------------------------------------------------------------------------------
// the call site travels with the arrival: threads of one block meeting at DIFFERENT __syncthreads is reported
#define __syncthreads() simt::syncthreads(__FILE__, __LINE__)
inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }

template <typename T>
inline T __ldg(const T* p) {
    simt::check_align(p, alignof(T), "__ldg");
    return *p;
}
template <typename T>
inline T __ldcg(const T* p) {
    simt::check_align(p, alignof(T), "__ldcg");
    return *reinterpret_cast<const volatile T*>(p);
}

inline double __dmul_rn(double a, double b) { return a * b; }   // compiled with -ffp-contract=off
inline double __dadd_rn(double a, double b) { return a + b; }
inline double __dsub_rn(double a, double b) { return a - b; }
inline double __ddiv_rn(double a, double b) { return a / b; }
inline double __dsqrt_rn(double a) { return sqrt(a); }
inline double __fma_rn(double a, double b, double c) { return fma(a, b, c); }
inline double __longlong_as_double(long long v) { double d; memcpy(&d, &v, 8); return d; }
inline long long __double_as_longlong(double d) { long long v; memcpy(&v, &d, 8); return v; }
inline int __double2hiint(double d) { long long v; memcpy(&v, &d, 8); return (int)(v >> 32); }
inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
inline double rsqrt(double a) { return 1.0 / sqrt(a); }
inline float rsqrtf(float a) { return 1.0f / sqrtf(a); }

namespace simt {
template <typename T>
inline unsigned long long to_bits(T v) {
    static_assert(sizeof(T) <= 8, "shuffle of a type wider than 64 bits");
    unsigned long long b = 0;
    memcpy(&b, &v, sizeof(T));
    return b;
}
template <typename T>
inline T from_bits(unsigned long long b) {
    T v;
    memcpy(&v, &b, sizeof(T));
    return v;
}
inline int lane_id() { return (int)(cur()->tid.x & 31u); }
}  // namespace simt

template <typename T>
inline T __shfl_xor_sync(unsigned mask, T v, int lane_mask) {
    (void)mask;
    const unsigned long long* all = simt::warp_gather(simt::to_bits(v));
    const int src = simt::lane_id() ^ lane_mask;
    // CUDA: reading from an inactive / out-of-range lane returns the caller's own value
    if (src > 31 || !((simt::warp_alive_mask() >> src) & 1u)) return v;
    return simt::from_bits<T>(all[src]);
}
template <typename T>
inline T __shfl_sync(unsigned mask, T v, int src_lane) {
    (void)mask;
    const unsigned long long* all = simt::warp_gather(simt::to_bits(v));
    const int src = src_lane & 31;
    if (!((simt::warp_alive_mask() >> src) & 1u)) return v;
    return simt::from_bits<T>(all[src]);
}
// warp barrier: a rendezvous of the warp's live lanes, like every other *_sync collective here
inline void __syncwarp(unsigned mask = 0xffffffffu) { (void)mask; (void)simt::warp_gather(0ull); }
inline int __any_sync(unsigned mask, int pred) {
    (void)mask;
    const unsigned long long* all = simt::warp_gather(pred ? 1ull : 0ull);
    const unsigned alive = simt::warp_alive_mask();
    for (int l = 0; l < 32; ++l)
        if (((alive >> l) & 1u) && all[l]) return 1;
    return 0;
}
inline int __all_sync(unsigned mask, int pred) {
    (void)mask;
    const unsigned long long* all = simt::warp_gather(pred ? 1ull : 0ull);
    const unsigned alive = simt::warp_alive_mask();
    for (int l = 0; l < 32; ++l)
        if (((alive >> l) & 1u) && !all[l]) return 0;
    return 1;
}
inline unsigned __ballot_sync(unsigned mask, int pred) {
    (void)mask;
    const unsigned long long* all = simt::warp_gather(pred ? 1ull : 0ull);
    const unsigned alive = simt::warp_alive_mask();
    unsigned r = 0;
    for (int l = 0; l < 32; ++l)
        if (((alive >> l) & 1u) && all[l]) r |= 1u << l;
    return r;
}

inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) {
    return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST);
}

// CUDA's overloaded min / max for mixed integer widths
inline int min(int a, int b) { return a < b ? a : b; }
inline int max(int a, int b) { return a > b ? a : b; }
inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
inline long long min(long long a, long long b) { return a < b ? a : b; }
inline long long max(long long a, long long b) { return a > b ? a : b; }
inline long long min(long long a, int b) { return a < b ? a : (long long)b; }
inline long long min(int a, long long b) { return a < b ? (long long)a : b; }
inline size_t min(size_t a, size_t b) { return a < b ? a : b; }

inline size_t __cvta_generic_to_shared(const void* p) { return (size_t)p; }

// ---- the slice of the CUDA runtime libgfk uses (synchronous, "device memory" = host memory) -----------------
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorInvalidValue = 1, cudaErrorMemoryAllocation = 2, cudaErrorNotSupported = 801 };
typedef struct SimtStream* cudaStream_t;
struct SimtEvent { double t_ms; };
typedef SimtEvent* cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount = 16, cudaDevAttrCooperativeLaunch = 95, cudaDevAttrMaxSharedMemoryPerBlockOptin = 97 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
struct cudaDeviceProp {
    char name[256];
    int major, minor, multiProcessorCount;
};
struct cudaIpcMemHandle_t { char reserved[64]; };
enum { cudaIpcMemLazyEnablePeerAccess = 1 };

const char* cudaGetErrorString(cudaError_t e);
cudaError_t cudaGetLastError();
cudaError_t cudaGetDevice(int* dev);
cudaError_t cudaDeviceGetAttribute(int* v, cudaDeviceAttr a, int dev);
cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int dev);
cudaError_t cudaMalloc(void** p, size_t bytes);
cudaError_t cudaFree(void* p);
cudaError_t cudaMemcpy(void* dst, const void* src, size_t bytes, cudaMemcpyKind k);
cudaError_t cudaMemcpyAsync(void* dst, const void* src, size_t bytes, cudaMemcpyKind k, cudaStream_t st);
cudaError_t cudaMemset(void* p, int v, size_t bytes);
cudaError_t cudaMemsetAsync(void* p, int v, size_t bytes, cudaStream_t st);
cudaError_t cudaStreamSynchronize(cudaStream_t st);
cudaError_t cudaDeviceSynchronize();
cudaError_t cudaEventCreate(cudaEvent_t* e);
cudaError_t cudaEventDestroy(cudaEvent_t e);
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t st);
cudaError_t cudaEventSynchronize(cudaEvent_t e);
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b);
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t* h, void* p);
cudaError_t cudaIpcOpenMemHandle(void** p, cudaIpcMemHandle_t h, unsigned flags);
cudaError_t cudaIpcCloseMemHandle(void* p);
cudaError_t simt_func_set_attribute(size_t dyn_smem_bytes);
cudaError_t simt_occupancy(int* blocks, int threads, size_t smem);

template <typename F>
inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int value) { return simt_func_set_attribute((size_t)value); }
template <typename F>
inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int* blocks, F, int threads, size_t smem) {
    return simt_occupancy(blocks, threads, smem);
}
