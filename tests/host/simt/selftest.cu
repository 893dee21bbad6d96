// selftest.cu - kernels with deliberate defects, to show the SIMT emulator reports what it is meant to catch.
// TEST INFRASTRUCTURE ONLY (compiled into tests/host/libgfk_sim.so, never into libgfk.so).
#include <cuda_runtime.h>

namespace {

__global__ void divergent_barrier_kernel(int* out) {
    if (threadIdx.x < 48) __syncthreads();       // half a warp never arrives
    out[threadIdx.x] = 1;
    __syncthreads();
}

__global__ void misaligned_kernel(const double* p, double* out) {
    const double2 v = __ldg(reinterpret_cast<const double2*>(p + 1));   // 8-byte aligned address, 16-byte access
    out[0] = v.x + v.y;
}

// neighbour exchange through shared memory; with_barrier = 0 leaves out the __syncthreads between write and read
__global__ void neighbour_kernel(int* out, int with_barrier) {
    __shared__ int buf[128];
    buf[threadIdx.x] = (int)threadIdx.x * 3 + (int)blockIdx.x;
    if (with_barrier) __syncthreads();
    out[blockIdx.x * blockDim.x + threadIdx.x] = buf[(threadIdx.x + 1) % blockDim.x];
}

__global__ void early_exit_kernel(int* out, int n) {
    // threads beyond n exit before the collectives: they no longer take part in ballots or barriers
    if ((int)threadIdx.x >= n) return;
    const unsigned present = __ballot_sync(0xffffffffu, 1);
    __syncthreads();
    out[threadIdx.x] = (int)present;
}

// cp.async staging; with_wait = 0 reads the staged value without cp.async.wait_group
__global__ void async_copy_kernel(const double* src, double* out, int with_wait) {
    extern __shared__ double stage[];
    simt::cp_async(&stage[threadIdx.x], src + threadIdx.x, 8);
    simt::cp_async_commit();
    if (with_wait) simt::cp_async_wait(0);
    __syncthreads();
    out[threadIdx.x] = stage[(threadIdx.x + 1) % blockDim.x];
}

}  // namespace

extern "C" int simt_selftest_async_copy(const double* src, double* out, int with_wait) {
    async_copy_kernel<<<1, 64, 64 * sizeof(double)>>>(src, out, with_wait);
    return (int)cudaGetLastError();
}

extern "C" int simt_selftest_divergent_barrier(void) {
    int storage[64] = {0};
    int* out = storage;
    divergent_barrier_kernel<<<1, 64>>>(out);
    return (int)cudaGetLastError();
}

extern "C" int simt_selftest_misaligned(void) {
    alignas(16) double storage[5] = {1, 2, 3, 4, 0};
    const double* buf = storage;
    double* out = storage + 4;
    misaligned_kernel<<<1, 1>>>(buf, out);
    return (int)cudaGetLastError();
}

extern "C" int simt_selftest_neighbour(int* out, int blocks, int with_barrier) {
    neighbour_kernel<<<blocks, 128>>>(out, with_barrier);
    return (int)cudaGetLastError();
}

extern "C" int simt_selftest_early_exit(int* out, int n) {
    early_exit_kernel<<<1, 64>>>(out, n);
    return (int)cudaGetLastError();
}

extern "C" const char* simt_selftest_error(int code) { return cudaGetErrorString(code); }
