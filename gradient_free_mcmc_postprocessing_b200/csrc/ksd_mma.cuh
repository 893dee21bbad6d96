// ksd_mma.cuh - the KSD pair sums on the 5th-generation tensor cores (included by ksd.cu).
//
// BASELINE.json north_star, config 5: "for large d the x'Ly cross terms are GEMM-shaped and belong on the tensor
// cores".  With a scalar preconditioner the kernel k(x_i, x_j) needs three dot products per pair
// (report/report.tex:853-868):
//     |x_i - x_j|^2          = |x_i|^2 + |x_j|^2 - 2 <x_i, x_j>
//     <s_i - s_j, x_i - x_j> = <s_i, x_i> + <s_j, x_j> - <s_i, x_j> - <x_i, s_j>
//     <s_i, s_j>
// i.e. four tile GEMMs  G1 = X X',  G2 = X S',  G3 = S X',  G4 = S S'  plus per-row scalars; everything after the dot
// products (powers of the quadratic form, weights, triangular mask, row sums) is the epilogue of ksd_rowsum_kernel.
// This is the float32 mode of the path (1e-5 relative on KSD values, BASELINE.json), for d <= 512:
//
//   * operands: tcgen05.mma kind::tf32 keeps 11 significant bits of an operand, far short of 1e-5, so every value
//     is split v = hi + lo (hi = tf32(v), lo = tf32(v - hi): 22 bits together) by a prep kernel and a product is
//     hi hi' + hi lo' + lo hi' ("3xTF32"; the dropped lo lo' term is 2^-22 relative).  x is centred first (both
//     expansions above are invariant under a common shift of x), which keeps the cancellation in the first one small.
//   * a CTA of six warps owns a 128-row tile: warp 0 is the TMA producer (for d <= 64 the four split row operands stay
//     resident in shared memory - beyond that their K blocks stream with the column tiles' -, column tiles of 64 rows stream through a two-buffer ring, one 128-byte K block per stage, all as
//     SWIZZLE_128B K-major tiles - exactly what a UMMA shared-memory descriptor addresses), warp 1 allocates tensor
//     memory and one of its lanes issues the MMAs (M = 128, N = 64, K = 8: 12 per K step), warps 2-5 are the epilogue:
//     thread r reads row r of the four accumulators from tensor memory with tcgen05.ld and finishes its 64 pairs.
//   * the four 128 x 64 float accumulators take 256 of the 512 tensor-memory columns, so they are double-buffered: the
//     MMAs of column tile J + 1 run while the epilogue works on J.  Producer, issuer and epilogue meet only through
//     mbarriers (TMA complete_tx, tcgen05.commit, plain arrivals).
//
// [round 2: opt-in with GFK_KSD_MMA=1; executed on the CPU emulator of the test suite only (which decodes the same
//  descriptors, tests/host/simt), not yet on a GPU.  Descriptor bit layouts follow cute::UMMA::SmemDescriptor /
//  InstrDescriptor of the CUTLASS headers vendored in this image.]
#pragma once

#include "tma.cuh"

namespace gfk {

constexpr int MM_M = 128;            // rows of a row tile = tensor-memory lanes = epilogue threads
constexpr int MM_N = 64;             // rows of a column tile = accumulator columns
constexpr int MM_KB = 32;            // floats per K block = one 128-byte swizzle row
constexpr int MM_THREADS = 192;      // producer warp, MMA warp, four epilogue warps
constexpr int MM_ROW_TILE_BYTES = MM_M * 128;       // one operand array, one K block
constexpr int MM_COL_TILE_BYTES = MM_N * 128;
constexpr int MM_RING = 2;           // column-tile stages in flight
constexpr int MM_ACC_COLS = 4 * MM_N;               // G1 .. G4 of one column tile

struct KsdMmaArgs {
    const float* nx;      // |x_c|^2 per selected row (x_c: centred)
    const float* sxx;     // <s, x_c>
    const float* w;       // exp(lw) per selected row, or nullptr
    long long m;          // selected rows
    int nkb;              // K blocks (dp / 32)
    int part, n_parts;    // this launch computes the 128-row tiles I with I % n_parts == part (ranks split the triangle)
    KParams kp;
    double* low;
    unsigned* status;     // device word, zeroed by the host entry; 1: a pipeline wait timed out (results invalid)
};

// round to tf32 (10 explicit mantissa bits), ties away from zero - by integer arithmetic so that the emulator build and
// the device produce the same bits
__device__ __forceinline__ float tf32_round(float v) {
    unsigned b = __float_as_uint(v);
    b = (b + 0x1000u) & 0xffffe000u;
    return __uint_as_float(b);
}

// mean[k] of column k over the selected rows (double accumulation, fixed order): one block per column
template <typename Dummy = void>
__global__ void __launch_bounds__(256) ksd_mma_colmean_kernel(const float* __restrict__ x, const uint32_t* __restrict__ sel,
                                                              long long m, int d, double* __restrict__ mean) {
    __shared__ double sh[8];
    const int k = blockIdx.x;
    double acc = 0.0;
    for (long long t = threadIdx.x; t < m; t += blockDim.x) {
        const long long row = sel ? (long long)sel[t] : t;
        acc += (double)x[row * d + k];
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double tot = 0.0;
        for (int w = 0; w < 8; ++w) tot += sh[w];
        mean[k] = tot / (double)m;
    }
}

// one warp per padded row t: centred x and s split into tf32 (hi, lo) pairs, zero padding for t >= m and k >= d, and the
// per-row scalars |x_c|^2, <s, x_c>, weight
template <typename Dummy = void>
__global__ void __launch_bounds__(256) ksd_mma_split_kernel(const float* __restrict__ x, const float* __restrict__ s,
                                                            const float* __restrict__ lw, const float* __restrict__ w,
                                                            const uint32_t* __restrict__ sel, long long m, long long m_pad,
                                                            int d, int dp, const double* __restrict__ mean,
                                                            float* __restrict__ xh, float* __restrict__ xl,
                                                            float* __restrict__ sh, float* __restrict__ sl,
                                                            float* __restrict__ nx, float* __restrict__ sxx,
                                                            float* __restrict__ wout) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long t = warp; t < m_pad; t += nwarps) {
        const long long row = (t < m) ? (sel ? (long long)sel[t] : t) : 0;
        double a_nx = 0.0, a_sx = 0.0;
        for (int k = lane; k < dp; k += 32) {
            float xc = 0.f, sv = 0.f;
            if (t < m && k < d) {
                xc = x[row * d + k] - (float)mean[k];
                sv = s[row * d + k];
            }
            const float xhi = tf32_round(xc), shi = tf32_round(sv);
            const float xlo = tf32_round(xc - xhi), slo = tf32_round(sv - shi);
            xh[t * dp + k] = xhi;
            xl[t * dp + k] = xlo;
            sh[t * dp + k] = shi;
            sl[t * dp + k] = slo;
            // the scalars use the values the GEMMs see (hi + lo), so the expansions cancel consistently
            const double xe = (double)xhi + (double)xlo, se = (double)shi + (double)slo;
            a_nx += xe * xe;
            a_sx += se * xe;
        }
        a_nx = warp_sum(a_nx);
        a_sx = warp_sum(a_sx);
        if (lane == 0) {
            nx[t] = (float)a_nx;
            sxx[t] = (float)a_sx;
            float wv = 1.f;
            if (t < m && lw) wv = w ? w[row] : expf(lw[row]);
            wout[t] = (t < m) ? wv : 0.f;
        }
    }
}

// ---- tcgen05 wrappers ---------------------------------------------------------------------------------------------
#if defined(GFK_SIMT_HOST)
__device__ __forceinline__ unsigned dev_smem_u32(const void* p) { return simt::smem_u32(p); }
__device__ __forceinline__ void dev_umma_tf32(unsigned tmem_d, unsigned long long da, unsigned long long db, unsigned idesc,
                                              unsigned accumulate) {
    simt::umma_tf32(tmem_d, da, db, idesc, accumulate != 0);
}
__device__ __forceinline__ void dev_umma_commit(unsigned long long* bar) { simt::umma_commit(bar); }
__device__ __forceinline__ void dev_tmem_alloc512(uint32_t* dst) { simt::tmem_alloc(dst, 512); }
__device__ __forceinline__ void dev_tmem_free512(uint32_t base) { simt::tmem_dealloc(base, 512); }
__device__ __forceinline__ void dev_tc_fence_before() { }
__device__ __forceinline__ void dev_tc_fence_after() { }
__device__ __forceinline__ void dev_tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) { simt::tmem_ld(taddr, r, 16); }
__device__ __forceinline__ void dev_tmem_wait_ld() { }
#else
__device__ __forceinline__ unsigned dev_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
// D[tmem] (+)= A[smem desc] * B[smem desc]'; one thread issues on behalf of the CTA
__device__ __forceinline__ void dev_umma_tf32(unsigned tmem_d, unsigned long long da, unsigned long long db, unsigned idesc,
                                              unsigned accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
// the mbarrier gets one arrival when every MMA issued so far by this thread has completed (implies
// tcgen05.fence::before_thread_sync)
__device__ __forceinline__ void dev_umma_commit(unsigned long long* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                 ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void dev_tmem_alloc512(uint32_t* dst) {       // one full warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"((unsigned)__cvta_generic_to_shared(dst)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void dev_tmem_free512(uint32_t base) {        // the same warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(512u) : "memory");
}
__device__ __forceinline__ void dev_tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void dev_tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void dev_tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void dev_tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
#endif

// cute::UMMA::SmemDescriptor of a K-major SWIZZLE_128B tile whose rows are 128 bytes: start address and offsets in
// 16-byte units; bits [0,14) start, [16,30) leading byte offset (1: unused for swizzled K-major), [32,46) stride byte
// offset (8 rows x 128 bytes = 1024), [46,48) version = 1 (Blackwell), [61,64) layout type = 2 (SWIZZLE_128B).
// K advances inside the swizzle row by moving the start address (32 bytes = one K = 8 tf32 step).
__device__ __forceinline__ unsigned long long umma_desc_sw128(unsigned smem_addr) {
    return (unsigned long long)((smem_addr & 0x3ffffu) >> 4) | (1ull << 16) | ((unsigned long long)(1024 >> 4) << 32) |
           (1ull << 46) | (2ull << 61);
}
// cute::UMMA::InstrDescriptor: [4,6) accumulator format 1 = F32, [7,10) / [10,13) A / B format 2 = TF32, bits 15 / 16
// A / B major 0 = K, [17,23) N >> 3, [24,29) M >> 4
constexpr unsigned kUmmaIdescTf32 = (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(MM_N >> 3) << 17) | ((unsigned)(MM_M >> 4) << 24);

// RESIDENT: the four row operands stay in shared memory for the whole tile (d <= 64: 4 arrays x nkb x 16 KB) and only
// column tiles stream; otherwise (64 < d <= 512) a ring stage carries the K block of BOTH the row and the column
// operands (96 KB per stage) - the row tile is then re-read once per column tile, which only L2 makes affordable.
template <bool RESIDENT>
__global__ void __launch_bounds__(MM_THREADS, 1) ksd_mma_kernel(KsdMmaArgs a, const __grid_constant__ TensorMap map_xh,
                                                                const __grid_constant__ TensorMap map_xl,
                                                                const __grid_constant__ TensorMap map_sh,
                                                                const __grid_constant__ TensorMap map_sl) {
    GFK_DYN_SMEM(mm_smem);
    __shared__ __align__(8) unsigned long long row_full, col_full[MM_RING], col_empty[MM_RING], acc_full[2], acc_empty[2];
    __shared__ uint32_t tmem_base_sh;
    const int nkb = a.nkb;
    // RESIDENT: [1024-byte aligned] row operands: array arr (xh, xl, sh, sl), K block kb -> tile (arr * nkb + kb); then the
    // column ring: stage st, array arr -> tile (st * 4 + arr).  Streaming: stage st = { 4 row tiles, 4 column tiles }.
    unsigned char* smem0 = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(mm_smem) + 1023u) & ~(uintptr_t)1023u);
    constexpr size_t kStageBytes = RESIDENT ? (size_t)4 * MM_COL_TILE_BYTES : (size_t)4 * (MM_ROW_TILE_BYTES + MM_COL_TILE_BYTES);
    unsigned char* ring = RESIDENT ? smem0 + (size_t)4 * nkb * MM_ROW_TILE_BYTES : smem0;
    auto row_tile = [&](int st, int arr, int kb) -> unsigned char* {
        return RESIDENT ? smem0 + (size_t)(arr * nkb + kb) * MM_ROW_TILE_BYTES
                        : ring + (size_t)st * kStageBytes + (size_t)arr * MM_ROW_TILE_BYTES;
    };
    auto col_tile = [&](int st, int arr) -> unsigned char* {
        return ring + (size_t)st * kStageBytes + (RESIDENT ? 0 : (size_t)4 * MM_ROW_TILE_BYTES) + (size_t)arr * MM_COL_TILE_BYTES;
    };

    const long long ntiles = (a.m + MM_M - 1) / MM_M;
    const long long my_tiles = (ntiles - a.part + a.n_parts - 1) / a.n_parts;       // == gridDim.x
    const long long I = a.part + (my_tiles - 1 - (long long)blockIdx.x) * a.n_parts;   // this part's tiles, largest first
    const long long t0 = I * MM_M;
    long long nJ = (t0 + MM_M - 1 + MM_N - 1) / MM_N;                               // column tiles holding some b < t
    const long long nJ_all = (a.m + MM_N - 1) / MM_N;
    if (nJ > nJ_all) nJ = nJ_all;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        dev_mbar_init(&row_full, 1);
        for (int i = 0; i < MM_RING; ++i) { dev_mbar_init(&col_full[i], 1); dev_mbar_init(&col_empty[i], 1); }
        for (int i = 0; i < 2; ++i) { dev_mbar_init(&acc_full[i], 1); dev_mbar_init(&acc_empty[i], 4); }
        dev_mbar_fence_init();
    }
    if (warp == 1) dev_tmem_alloc512(&tmem_base_sh);
    dev_tc_fence_before();
    __syncthreads();
    dev_tc_fence_after();
    const uint32_t tmem_base = tmem_base_sh;
    const TensorMap* maps[4] = {&map_xh, &map_xl, &map_sh, &map_sl};

    if (warp == 0) {
        // ---- TMA producer -----------------------------------------------------------------------------------
        if (lane == 0) {
            if (RESIDENT) {
                dev_mbar_arrive_expect_tx(&row_full, 4u * (unsigned)nkb * MM_ROW_TILE_BYTES);
                for (int arr = 0; arr < 4; ++arr)
                    for (int kb = 0; kb < nkb; ++kb)
                        for (int half = 0; half < MM_M / MM_N; ++half)
                            dev_tma_load_2d(row_tile(0, arr, kb) + (size_t)half * MM_COL_TILE_BYTES, maps[arr], kb * MM_KB,
                                            (int)(t0 + half * MM_N), &row_full);
            }
            for (long long q = 0; q < nJ * nkb; ++q) {
                const int st = (int)(q % MM_RING);
                const long long use = q / MM_RING;
                if (use > 0 && !dev_mbar_wait(&col_empty[st], (unsigned)((use - 1) & 1), a.status)) break;   // the MMAs that read it are done
                const long long J = q / nkb;
                const int kb = (int)(q - J * nkb);
                dev_mbar_arrive_expect_tx(&col_full[st], (unsigned)kStageBytes);
                for (int arr = 0; arr < 4; ++arr) {
                    if (!RESIDENT)
                        for (int half = 0; half < MM_M / MM_N; ++half)
                            dev_tma_load_2d(row_tile(st, arr, kb) + (size_t)half * MM_COL_TILE_BYTES, maps[arr], kb * MM_KB,
                                            (int)(t0 + half * MM_N), &col_full[st]);
                    dev_tma_load_2d(col_tile(st, arr), maps[arr], kb * MM_KB, (int)(J * MM_N), &col_full[st]);
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ---- MMA issuer: ONE lane waits, issues and commits --------------------------------------------------------
        // (a lane that only polled a barrier its warp-mate acts on could miss a phase: the elected lane's commit lets
        //  the producer refill the stage, and two completions later the parity looks unfinished again)
        if (lane == 0) {
            bool alive = RESIDENT ? dev_mbar_wait(&row_full, 0, a.status) : true;
            for (long long J = 0; alive && J < nJ; ++J) {
                const int ab = (int)(J & 1);
                const long long useA = J >> 1;
                if (useA > 0 && !dev_mbar_wait(&acc_empty[ab], (unsigned)((useA - 1) & 1), a.status)) break;   // the epilogue has drained it
                dev_tc_fence_after();
                const uint32_t acc = tmem_base + (uint32_t)(ab * MM_ACC_COLS);
                for (int kb = 0; kb < nkb; ++kb) {
                    const long long q = J * nkb + kb;
                    const int st = (int)(q % MM_RING);
                    if (!dev_mbar_wait(&col_full[st], (unsigned)((q / MM_RING) & 1), a.status)) { alive = false; break; }   // K block landed
                    dev_tc_fence_after();
                    const unsigned r_xh = dev_smem_u32(row_tile(st, 0, kb)), r_xl = dev_smem_u32(row_tile(st, 1, kb));
                    const unsigned r_sh = dev_smem_u32(row_tile(st, 2, kb)), r_sl = dev_smem_u32(row_tile(st, 3, kb));
                    const unsigned c_xh = dev_smem_u32(col_tile(st, 0)), c_xl = dev_smem_u32(col_tile(st, 1));
                    const unsigned c_sh = dev_smem_u32(col_tile(st, 2)), c_sl = dev_smem_u32(col_tile(st, 3));
#pragma unroll
                    for (int ks = 0; ks < MM_KB / 8; ++ks) {
                        const unsigned off = (unsigned)ks * 32u;                   // 8 tf32 = 32 bytes along K
                        const unsigned first = (kb | ks) ? 1u : 0u;                // the first MMA of a tile overwrites
                        const unsigned long long Rxh = umma_desc_sw128(r_xh + off), Rxl = umma_desc_sw128(r_xl + off);
                        const unsigned long long Rsh = umma_desc_sw128(r_sh + off), Rsl = umma_desc_sw128(r_sl + off);
                        const unsigned long long Cxh = umma_desc_sw128(c_xh + off), Cxl = umma_desc_sw128(c_xl + off);
                        const unsigned long long Csh = umma_desc_sw128(c_sh + off), Csl = umma_desc_sw128(c_sl + off);
                        // G1 = <x_i, x_j>
                        dev_umma_tf32(acc + 0 * MM_N, Rxh, Cxh, kUmmaIdescTf32, first);
                        dev_umma_tf32(acc + 0 * MM_N, Rxh, Cxl, kUmmaIdescTf32, 1u);
                        dev_umma_tf32(acc + 0 * MM_N, Rxl, Cxh, kUmmaIdescTf32, 1u);
                        // G2 = <x_i, s_j>
                        dev_umma_tf32(acc + 1 * MM_N, Rxh, Csh, kUmmaIdescTf32, first);
                        dev_umma_tf32(acc + 1 * MM_N, Rxh, Csl, kUmmaIdescTf32, 1u);
                        dev_umma_tf32(acc + 1 * MM_N, Rxl, Csh, kUmmaIdescTf32, 1u);
                        // G3 = <s_i, x_j>
                        dev_umma_tf32(acc + 2 * MM_N, Rsh, Cxh, kUmmaIdescTf32, first);
                        dev_umma_tf32(acc + 2 * MM_N, Rsh, Cxl, kUmmaIdescTf32, 1u);
                        dev_umma_tf32(acc + 2 * MM_N, Rsl, Cxh, kUmmaIdescTf32, 1u);
                        // G4 = <s_i, s_j>
                        dev_umma_tf32(acc + 3 * MM_N, Rsh, Csh, kUmmaIdescTf32, first);
                        dev_umma_tf32(acc + 3 * MM_N, Rsh, Csl, kUmmaIdescTf32, 1u);
                        dev_umma_tf32(acc + 3 * MM_N, Rsl, Csh, kUmmaIdescTf32, 1u);
                    }
                    dev_umma_commit(&col_empty[st]);               // the ring stage may be refilled once these MMAs are done
                    if (kb == nkb - 1) dev_umma_commit(&acc_full[ab]);   // ... and the epilogue may read the accumulators
                }
            }
        }
        __syncwarp();
    } else {
        // ---- epilogue warps: thread r owns row r of the tile = tensor-memory lane r ------------------------------
        const int quarter = warp & 3;                                // the lanes this warp may touch: 32 quarter .. + 31
        const int r = quarter * 32 + lane;
        const long long t = t0 + r;
        const bool live = t < a.m;
        const float nx_i = live ? a.nx[t] : 0.f, sxx_i = live ? a.sxx[t] : 0.f;
        const float w_i = (live && a.w) ? a.w[t] : 1.f;
        double rowacc = 0.0;
        for (long long J = 0; J < nJ; ++J) {
            const int ab = (int)(J & 1);
            // warp-uniform verdict: the tcgen05.ld below are .sync.aligned, a warp must leave the loop together
            if (!__all_sync(0xffffffffu, dev_mbar_wait(&acc_full[ab], (unsigned)((J >> 1) & 1), a.status))) break;
            dev_tc_fence_after();
            const uint32_t acc = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(ab * MM_ACC_COLS);
#pragma unroll 1
            for (int c = 0; c < MM_N / 16; ++c) {
                uint32_t g1[16], g2[16], g3[16], g4[16];
                dev_tmem_ld16(acc + 0 * MM_N + c * 16, g1);
                dev_tmem_ld16(acc + 1 * MM_N + c * 16, g2);
                dev_tmem_ld16(acc + 2 * MM_N + c * 16, g3);
                dev_tmem_ld16(acc + 3 * MM_N + c * 16, g4);
                dev_tmem_wait_ld();
                double part = 0.0;
#pragma unroll
                for (int jj = 0; jj < 16; ++jj) {
                    const long long b = J * MM_N + c * 16 + jj;
                    if (live && b < t) {                              // b < t < m: the column scalars exist
                        const float ny = __ldg(a.nx + b), syy = __ldg(a.sxx + b);
                        const float rr = fmaxf(fmaf(-2.f, __uint_as_float(g1[jj]), nx_i + ny), 0.f);
                        const float su = (sxx_i + syy) - (__uint_as_float(g3[jj]) + __uint_as_float(g2[jj]));
                        float kv = fast_finish<float, false>(rr, 0.f, su, __uint_as_float(g4[jj]), a.kp);
                        if (a.w) kv *= w_i * __ldg(a.w + b);
                        part += (double)kv;
                    }
                }
                rowacc += part;
            }
            dev_tc_fence_before();
            __syncwarp();
            if (lane == 0) dev_mbar_arrive(&acc_empty[ab]);          // one arrival per epilogue warp
        }
        if (live) a.low[t] = rowacc;
    }
    // ---- teardown: every warp is done with tensor memory -------------------------------------------------------------
    dev_tc_fence_before();
    __syncthreads();
    dev_tc_fence_after();
    if (warp == 1) dev_tmem_free512(tmem_base);
}

constexpr int MM_RESIDENT_MAX_KB = 2;                // row operands resident up to d = 64
static inline size_t ksd_mma_smem_bytes(int nkb) {
    if (nkb <= MM_RESIDENT_MAX_KB) return 1024 + (size_t)4 * nkb * MM_ROW_TILE_BYTES + (size_t)MM_RING * 4 * MM_COL_TILE_BYTES;
    return 1024 + (size_t)MM_RING * 4 * (MM_ROW_TILE_BYTES + MM_COL_TILE_BYTES);
}

}  // namespace gfk
