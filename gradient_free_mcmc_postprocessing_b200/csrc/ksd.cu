// ksd.cu - O(m^2) kernel Stein discrepancy: tiled pair sums, cumulative KSD, Gram matrix.
//
// Replaces stein_thinning.stein.ksd / kmat (call sites code/src/utils/ksd.py:27,
// code/tests/test_ksd.py:20; definition report/report.tex:309-313):
//   ks[t] = sqrt( sum_{a,b<=t} k(a,b) ) / (t+1),   row t adds 2 sum_{b<t} k(t,b) + k(t,t).
//
// gfk_ksd_rowsums: one CTA per 64-row tile I walks the column tiles J <= I.  Row and column tiles
// are staged in shared memory transposed ([k][row], conflict-free), in chunks of KC dimensions so any
// d fits; each thread owns MTR rows x MTC columns (columns strided by TILE / MTC, so a warp reads conflict-free
// 128-byte lines per shared-memory load) and keeps the three running dot products
// (<r,r>, <ds,r>, <sx,sy>) of every pair in registers - a GEMM-shaped loop on the FP64 (or FP32)
// FMA pipe with ~0 bytes of HBM traffic per pair.  float64: 4 x 2 pairs per thread, 512 threads per CTA (round 1's
// 4 x 4 / 256-thread shape needed 176 registers, i.e. 8 warps per SM that stalled together at the two barriers of
// every stage: FP64 pipe ~39 % busy; 4 x 2 fits 128 registers, so the SM holds 16 warps - same tile, same shared
// memory, same operation order per pair).  float32 keeps 4 x 4 / 256 threads (two CTAs per SM).  Row sums are reduced in a fixed order (no
// atomics), so results are run-to-run reproducible and independent of the GPU count.
#include <stdlib.h>

#include "common.cuh"

namespace gfk {

constexpr int TILE = 64;     // rows / columns per tile
constexpr int LD = TILE + 2; // padded leading dimension of a staged tile: keeps 16-byte alignment per k
constexpr int MTR = 4;       // rows of the tile per thread
// columns of the tile per thread and the resulting CTA size
template <typename T> struct KsdShape { static constexpr int MTC = 4; static constexpr int THREADS = (TILE / MTR) * (TILE / 4); };
template <> struct KsdShape<double> { static constexpr int MTC = 2; static constexpr int THREADS = (TILE / MTR) * (TILE / 2); };

template <typename T>
struct KsdArgs {
    const T* x;
    const T* s;
    const T* lw;
    const T* w;
    const T* xl;
    const uint32_t* sel;
    long long m;
    KParams kp;
    int wstride;        // elements between consecutive rows' weights in w: 1 (fast modes), 2 (strict: (hi, lo) pairs)
    int part, n_parts;
    int deal_rows;      // rows per dealt tile: (t / deal_rows) % n_parts owns row t (TILE; 128 for the tensor-core kernel)
    double* low;
    double* diag;
};

template <typename T>
struct GView {
    const T* p;
    __device__ __forceinline__ T operator[](int k) const { return p[k]; }
};

template <typename T>
__device__ __forceinline__ long long row_of(const KsdArgs<T>& a, long long t) {
    return a.sel ? (long long)a.sel[t] : t;
}

// ---- asynchronous global -> shared copies (LDGSTS): no register staging, overlap with the FMA loop ----
#if defined(GFK_SIMT_HOST)   // CPU emulation build of the test suite: the emulator defers the copies like the hardware
__device__ __forceinline__ void cp_async_elem(double* dst, const double* src) { simt::cp_async(dst, src, 8); }
__device__ __forceinline__ void cp_async_elem(float* dst, const float* src) { simt::cp_async(dst, src, 4); }
__device__ __forceinline__ void cp_async_commit() { simt::cp_async_commit(); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { simt::cp_async_wait(N); }
#else
__device__ __forceinline__ void cp_async_elem(double* dst, const double* src) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(sa), "l"(src));
}
__device__ __forceinline__ void cp_async_elem(float* dst, const float* src) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(sa), "l"(src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }
#endif

// stage dimensions [k0, k0 + kc) of rows [t0, t0 + TILE) into dst[k][LD] (rows beyond m zeroed), asynchronously.
// Consecutive threads read consecutive k of one row (coalesced); the padded LD spreads the stores.  Only the kc
// dimensions the consumer reads are touched (at d = 10 that is 10 of the chunk's 32 rows of dst: a third of the
// staging work of a full chunk); (r, k) advance incrementally - one division per call, none per element.
template <typename T, int KC>
__device__ __forceinline__ void stage_tile(const KsdArgs<T>& a, const T* __restrict__ src, long long t0, int k0,
                                           int kc, T* __restrict__ dst) {
    const int d = a.kp.d;
    constexpr int NT = KsdShape<T>::THREADS;
    int r = (int)threadIdx.x / kc, k = (int)threadIdx.x - r * kc;
    const int dr = NT / kc, dk = NT - dr * kc;
    for (int e = threadIdx.x; e < TILE * kc; e += NT) {
        const long long t = t0 + r;
        if (t < a.m) cp_async_elem(dst + k * LD + r, src + row_of(a, t) * d + k0 + k);
        else dst[k * LD + r] = T(0);
        r += dr;
        k += dk;
        if (k >= kc) { k -= kc; ++r; }
    }
}

// dynamic shared memory: 2 stage buffers, each { xr sr xc sc [xlr xlc] } of KC x LD, then wr[TILE], wc[2][TILE].
// Stages are the (column tile J, dimension chunk k0) pairs in order; stage s+1 is copied in with cp.async
// while stage s is consumed.  With d <= KC the row tile is invariant and staged once into buffer 0's row slots.
template <typename T, bool DENSE, int KC>
__global__ void __launch_bounds__(KsdShape<T>::THREADS) ksd_rowsum_kernel(KsdArgs<T> a) {
    extern __shared__ __align__(16) unsigned char ksd_smem[];
    constexpr int NARR = DENSE ? 6 : 4;
    constexpr int MTC = KsdShape<T>::MTC;
    constexpr int CG = TILE / MTC;            // column groups = threads that share a row (consecutive lanes)
    T* buf0 = reinterpret_cast<T*>(ksd_smem);
    T* buf1 = buf0 + NARR * KC * LD;
    T* wr = buf1 + NARR * KC * LD;
    T* wcb = wr + TILE;                       // wc[2][TILE]

    const long long ntiles = (a.m + TILE - 1) / TILE;
    // this rank's tiles, largest (most column tiles) first
    const long long my_tiles = (ntiles - a.part + a.n_parts - 1) / a.n_parts;
    if ((long long)blockIdx.x >= my_tiles) return;
    const long long I = a.part + (my_tiles - 1 - blockIdx.x) * (long long)a.n_parts;
    const long long t0 = I * TILE;
    const int ty = threadIdx.x / CG, tx = threadIdx.x % CG;
    const int d = a.kp.d;
    const bool gf = (a.lw != nullptr);
    const bool pre_w = (a.w != nullptr);
    const bool single_chunk = (d <= KC);
    const int nchunks = (d + KC - 1) / KC;
    const long long nstages = (I + 1) * nchunks;

    auto issue = [&](long long stage) {
        T* base = (stage & 1) ? buf1 : buf0;
        const long long J = stage / nchunks;
        const int k0 = (int)(stage - J * nchunks) * KC;
        const int kc = min(KC, d - k0);
        const long long c0 = J * TILE;
        if (!single_chunk) {
            stage_tile<T, KC>(a, a.x, t0, k0, kc, base);
            stage_tile<T, KC>(a, a.s, t0, k0, kc, base + KC * LD);
            if (DENSE) stage_tile<T, KC>(a, a.xl, t0, k0, kc, base + 4 * KC * LD);
        }
        stage_tile<T, KC>(a, a.x, c0, k0, kc, base + 2 * KC * LD);
        stage_tile<T, KC>(a, a.s, c0, k0, kc, base + 3 * KC * LD);
        if (DENSE) stage_tile<T, KC>(a, a.xl, c0, k0, kc, base + 5 * KC * LD);
        if (k0 == 0 && threadIdx.x < TILE) {
            const long long t = c0 + threadIdx.x;
            T v = T(0);
            if (gf && t < a.m) v = pre_w ? __ldg(a.w + row_of(a, t) * a.wstride) : __ldg(a.lw + row_of(a, t));
            wcb[(J & 1) * TILE + threadIdx.x] = v;
        }
        cp_async_commit();
    };

    if (threadIdx.x < TILE) {
        const long long t = t0 + threadIdx.x;
        T v = T(0);
        if (gf && t < a.m) v = pre_w ? __ldg(a.w + row_of(a, t) * a.wstride) : __ldg(a.lw + row_of(a, t));
        wr[threadIdx.x] = v;
    }
    if (single_chunk) {   // the row tile is invariant over J: stage it once (buffer 0's row slots)
        stage_tile<T, KC>(a, a.x, t0, 0, d, buf0);
        stage_tile<T, KC>(a, a.s, t0, 0, d, buf0 + KC * LD);
        if (DENSE) stage_tile<T, KC>(a, a.xl, t0, 0, d, buf0 + 4 * KC * LD);
    }
    issue(0);

    double rowacc[MTR];
#pragma unroll
    for (int i = 0; i < MTR; ++i) rowacc[i] = 0.0;
    T rr[MTR][MTC], su[MTR][MTC], ss[MTR][MTC], uu[DENSE ? MTR : 1][DENSE ? MTC : 1];

    for (long long stage = 0; stage < nstages; ++stage) {
        const long long J = stage / nchunks;
        const int chunk = (int)(stage - J * nchunks);
        const int kc = min(KC, d - chunk * KC);
        const long long c0 = J * TILE;
        if (stage + 1 < nstages) {
            issue(stage + 1);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        T* base = (stage & 1) ? buf1 : buf0;
        const T* xr = single_chunk ? buf0 : base;
        const T* sr = single_chunk ? buf0 + KC * LD : base + KC * LD;
        const T* xlr = single_chunk ? buf0 + 4 * KC * LD : base + 4 * KC * LD;
        const T* xc = base + 2 * KC * LD;
        const T* sc = base + 3 * KC * LD;
        const T* xlc = base + 5 * KC * LD;
        if (chunk == 0) {
#pragma unroll
            for (int i = 0; i < MTR; ++i)
#pragma unroll
                for (int j = 0; j < MTC; ++j) {
                    rr[i][j] = T(0); su[i][j] = T(0); ss[i][j] = T(0);
                    if (DENSE) uu[i][j] = T(0);
                }
        }
#pragma unroll 2
        for (int k = 0; k < kc; ++k) {
            T xi[MTR], si[MTR], xj[MTC], sj[MTC], li[DENSE ? MTR : 1], lj[DENSE ? MTC : 1];
#pragma unroll
            for (int i = 0; i < MTR; ++i) {
                xi[i] = xr[k * LD + ty * MTR + i];
                si[i] = sr[k * LD + ty * MTR + i];
                if (DENSE) li[i] = xlr[k * LD + ty * MTR + i];
            }
#pragma unroll
            for (int j = 0; j < MTC; ++j) {
                xj[j] = xc[k * LD + tx + CG * j];
                sj[j] = sc[k * LD + tx + CG * j];
                if (DENSE) lj[j] = xlc[k * LD + tx + CG * j];
            }
#pragma unroll
            for (int i = 0; i < MTR; ++i)
#pragma unroll
                for (int j = 0; j < MTC; ++j) {
                    const T r = xi[i] - xj[j];
                    const T ds = si[i] - sj[j];
                    if (DENSE) {
                        const T u = li[i] - lj[j];
                        rr[i][j] = fma_t(u, r, rr[i][j]);
                        uu[i][j] = fma_t(u, u, uu[i][j]);
                        su[i][j] = fma_t(ds, u, su[i][j]);
                    } else {
                        rr[i][j] = fma_t(r, r, rr[i][j]);
                        su[i][j] = fma_t(ds, r, su[i][j]);
                    }
                    ss[i][j] = fma_t(si[i], sj[j], ss[i][j]);
                }
        }
        if (chunk == nchunks - 1) {
            // epilogue of column tile J: powers, weights, triangular mask
            const T* wc = wcb + (J & 1) * TILE;
#pragma unroll
            for (int i = 0; i < MTR; ++i) {
                const long long t = t0 + ty * MTR + i;
                double part_sum = 0.0;
#pragma unroll
                for (int j = 0; j < MTC; ++j) {
                    const long long b = c0 + tx + CG * j;
                    T kv = fast_finish<T, DENSE>(rr[i][j], DENSE ? uu[i][j] : T(0), su[i][j], ss[i][j], a.kp);
                    if (gf) {
                        const T wi = wr[ty * MTR + i], wj = wc[tx + CG * j];
                        kv *= pre_w ? (wi * wj) : (T)exp((double)wi + (double)wj);
                    }
                    if (b < t && t < a.m) part_sum += (double)kv;
                }
                rowacc[i] += part_sum;
            }
        }
        __syncthreads();   // buffer (stage & 1) is free for the copy issued at the top of the next iteration
    }
    // reduce the CG threads that share a row (consecutive lanes: tx = lane % CG; CG = 16 or 32)
#pragma unroll
    for (int i = 0; i < MTR; ++i) {
        double v = rowacc[i];
#pragma unroll
        for (int off = CG / 2; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        const long long t = t0 + ty * MTR + i;
        if (tx == 0 && t < a.m) a.low[t] = v;
    }
}

// diag[t] = k(t, t) (with weights); rows of tiles owned by other parts are zeroed in low and diag
template <typename T>
__global__ void __launch_bounds__(256) ksd_diag_kernel(KsdArgs<T> a) {
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < a.m;
         t += (long long)gridDim.x * blockDim.x) {
        const bool mine = ((t / a.deal_rows) % a.n_parts) == a.part;
        if (!mine) { a.low[t] = 0.0; a.diag[t] = 0.0; continue; }
        const long long i = row_of(a, t);
        const int d = a.kp.d;
        GView<T> xv{a.x + i * d}, sv{a.s + i * d};
        T k;
        if (a.xl) {
            GView<T> lv{a.xl + i * d};
            k = pair_fast<T, 0, true>(xv, xv, sv, sv, lv, lv, a.kp);
        } else {
            k = pair_fast<T, 0, false>(xv, xv, sv, sv, xv, xv, a.kp);
        }
        double kd = (double)k;
        if (a.lw) {
            if (a.w) { const T wv = __ldg(a.w + i * a.wstride); kd = (double)(k * (wv * wv)); }
            else { const double l = (double)__ldg(a.lw + i); kd *= exp(l + l); }
        }
        a.diag[t] = kd;
    }
}

// ks[t] = sqrt(running sum of 2 low + diag) / (t + 1): sequential scan in index order (numpy's order)
__global__ void __launch_bounds__(256) ksd_finish_kernel(const double* __restrict__ low,
                                                         const double* __restrict__ diag, long long m,
                                                         double* __restrict__ ks) {
    __shared__ double buf[1024];
    __shared__ double carry;
    if (threadIdx.x == 0) carry = 0.0;
    for (long long base = 0; base < m; base += 1024) {
        const int cnt = (int)min(1024LL, m - base);
        __syncthreads();
        for (int e = threadIdx.x; e < cnt; e += blockDim.x) {
            const double l = low[base + e], dg = diag[base + e];
            buf[e] = (base + e > 0) ? __dadd_rn(__dmul_rn(2.0, l), dg) : dg;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            double c = carry;
            for (int e = 0; e < cnt; ++e) { c = __dadd_rn(c, buf[e]); buf[e] = c; }
            carry = c;
        }
        __syncthreads();
        for (int e = threadIdx.x; e < cnt; e += blockDim.x)
            ks[base + e] = __ddiv_rn(__dsqrt_rn(buf[e]), (double)(base + e + 1));
    }
}

// ---- element-wise evaluation (integrand protocol, Gram matrix) --------------------------------------
template <typename T, bool STRICT>
__device__ __forceinline__ double eval_pair(const KsdArgs<T>& a, long long i, long long j) {
    const int d = a.kp.d;
    GView<T> xi{a.x + i * d}, xj{a.x + j * d}, si{a.s + i * d}, sj{a.s + j * d};
    double k;
    if constexpr (STRICT) {
        k = pair_strict<0>(xi, xj, si, sj, a.kp);
        if (a.lw) {
            double wgt;
            if (a.w) {     // strict mode: double-double exponentials (hi, lo) per row, the weight of the thinning kernels
                wgt = exp_sum_dd((double)a.lw[i], (double)a.lw[j], (double)a.w[2 * i], (double)a.w[2 * i + 1],
                                 (double)a.w[2 * j], (double)a.w[2 * j + 1]);
            } else {
                wgt = exp(__dadd_rn((double)a.lw[i], (double)a.lw[j]));
            }
            k = __dmul_rn(wgt, k);
        }
    } else {
        if (a.xl) {
            GView<T> li{a.xl + i * d}, lj{a.xl + j * d};
            k = (double)pair_fast<T, 0, true>(xi, xj, si, sj, li, lj, a.kp);
        } else {
            k = (double)pair_fast<T, 0, false>(xi, xj, si, sj, xi, xj, a.kp);
        }
        if (a.lw) k *= a.w ? (double)(a.w[i * a.wstride] * a.w[j * a.wstride]) : exp((double)a.lw[i] + (double)a.lw[j]);
    }
    return k;
}

template <typename T, bool STRICT>
__global__ void __launch_bounds__(256) pairs_kernel(KsdArgs<T> a, const long long* __restrict__ ia,
                                                    const long long* __restrict__ ib, long long count,
                                                    double* __restrict__ out) {
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < count;
         e += (long long)gridDim.x * blockDim.x)
        out[e] = eval_pair<T, STRICT>(a, ia[e], ib[e]);
}

template <typename T, bool STRICT>
__global__ void __launch_bounds__(256) kmat_kernel(KsdArgs<T> a, double* __restrict__ out) {
    const long long total = a.m * a.m;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (long long)gridDim.x * blockDim.x) {
        const long long r = e / a.m, c = e - r * a.m;
        if (c < r) continue;
        const double k = eval_pair<T, STRICT>(a, row_of(a, r), row_of(a, c));
        out[r * a.m + c] = k;
        out[c * a.m + r] = k;
    }
}

template <typename T>
static KsdArgs<T> make_args(const gfk_problem* p, const uint32_t* sel, int64_t m) {
    KsdArgs<T> a;
    a.x = static_cast<const T*>(p->x);
    a.s = static_cast<const T*>(p->s);
    a.lw = static_cast<const T*>(p->lw);
    a.w = static_cast<const T*>(p->w);
    a.xl = static_cast<const T*>(p->xl);
    a.sel = sel;
    a.m = m;
    a.kp = KParams{p->lambda, p->lambda_sq, p->trace, p->c, p->d, p->xl != nullptr};
    a.wstride = (p->mode == GFK_F64_STRICT) ? 2 : 1;
    a.part = 0;
    a.n_parts = 1;
    a.deal_rows = TILE;
    a.low = nullptr;
    a.diag = nullptr;
    return a;
}

static int grid1d(long long work) {
    long long b = (work + 255) / 256;
    const long long cap = (long long)sm_count() * 8;
    if (b > cap) b = cap;
    return (int)(b < 1 ? 1 : b);
}

template <typename T>
static int rowsums_typed(const gfk_problem* p, const uint32_t* sel, int64_t m, int part, int n_parts, double* low,
                         double* diag, cudaStream_t st) {
    KsdArgs<T> a = make_args<T>(p, sel, m);
    a.part = part;
    a.n_parts = n_parts;
    a.low = low;
    a.diag = diag;
    ksd_diag_kernel<T><<<grid1d(m), 256, 0, st>>>(a);
    const long long ntiles = (m + TILE - 1) / TILE;
    const long long my_tiles = (ntiles - part + n_parts - 1) / n_parts;
    if (my_tiles > 0) {
        if (p->xl) {
            constexpr int KC = 16;
            const size_t smem = sizeof(T) * (size_t)(2 * 6 * KC * LD + 3 * TILE);
            GFK_CUDA(cudaFuncSetAttribute(ksd_rowsum_kernel<T, true, KC>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)smem));
            ksd_rowsum_kernel<T, true, KC><<<(unsigned)my_tiles, KsdShape<T>::THREADS, smem, st>>>(a);
        } else {
            constexpr int KC = 32;
            const size_t smem = sizeof(T) * (size_t)(2 * 4 * KC * LD + 3 * TILE);
            GFK_CUDA(cudaFuncSetAttribute(ksd_rowsum_kernel<T, false, KC>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)smem));
            ksd_rowsum_kernel<T, false, KC><<<(unsigned)my_tiles, KsdShape<T>::THREADS, smem, st>>>(a);
        }
    }
    count_launches(my_tiles > 0 ? 2 : 1);
    GFK_CHECK_LAUNCH();
    return GFK_OK;
}

}  // namespace gfk
#include "ksd_mma.cuh"
namespace gfk {

static int check_problem(const gfk_problem* p, const char* who) {
    GFK_REQUIRE(p && p->x && p->s, "%s: NULL problem / x / s", who);
    GFK_REQUIRE(p->n >= 1 && p->d >= 1, "%s: empty problem", who);
    GFK_REQUIRE(p->mode >= GFK_F64_STRICT && p->mode <= GFK_F32_FAST, "%s: unknown mode %d", who, p->mode);
    GFK_REQUIRE(!(p->xl && p->mode == GFK_F64_STRICT), "%s: a dense preconditioner needs a fast mode", who);
    return GFK_OK;
}

}  // namespace gfk

using namespace gfk;

extern "C" int gfk_ksd_rowsums(const gfk_problem* p, const uint32_t* sel, int64_t m, int32_t part, int32_t n_parts,
                               double* low, double* diag, void* stream) {
    int rc = check_problem(p, "gfk_ksd_rowsums");
    if (rc) return rc;
    GFK_REQUIRE(low && diag && m >= 1, "gfk_ksd_rowsums: bad arguments");
    GFK_REQUIRE(sel || m <= p->n, "gfk_ksd_rowsums: m = %lld exceeds n = %lld", (long long)m, (long long)p->n);
    GFK_REQUIRE(n_parts >= 1 && part >= 0 && part < n_parts, "gfk_ksd_rowsums: bad part %d of %d", part, n_parts);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // the sum order of numpy's ksd() is not reproducible on a GPU, so the strict mode shares the
    // FMA formulation here (KSD values are checked to 1e-10 relative, not bit for bit)
    if (p->mode == GFK_F32_FAST) return rowsums_typed<float>(p, sel, m, part, n_parts, low, diag, st);
    return rowsums_typed<double>(p, sel, m, part, n_parts, low, diag, st);
}

// ---- tensor-core row sums (ksd_mma.cuh): float32 mode, scalar preconditioner, d <= 512 (padded to 32s) -------------
static inline long long mma_pad_rows(long long m) { return (m + MM_M - 1) / MM_M * MM_M; }
static inline int mma_pad_dims(int d) { return (d + MM_KB - 1) / MM_KB * MM_KB; }

extern "C" int gfk_ksd_mma_supported(const gfk_problem* p) {
    return (p && p->mode == GFK_F32_FAST && !p->xl && p->d >= 1 && p->d <= 512) ? 1 : 0;
}

extern "C" size_t gfk_ksd_mma_ws_bytes(int64_t m, int32_t d) {
    if (m < 1 || d < 1) return 0;
    const size_t mp = (size_t)mma_pad_rows(m), dp = (size_t)mma_pad_dims(d);
    return 512 + sizeof(double) * dp + 4 * sizeof(float) * mp * dp + 3 * sizeof(float) * mp + 1024;
}

extern "C" int gfk_ksd_rowsums_mma(const gfk_problem* p, const uint32_t* sel, int64_t m, int32_t part, int32_t n_parts,
                                   double* low, double* diag, void* ws, void* stream) {
    int rc = check_problem(p, "gfk_ksd_rowsums_mma");
    if (rc) return rc;
    GFK_REQUIRE(low && diag && ws && m >= 1, "gfk_ksd_rowsums_mma: bad arguments");
    GFK_REQUIRE(sel || m <= p->n, "gfk_ksd_rowsums_mma: m = %lld exceeds n = %lld", (long long)m, (long long)p->n);
    GFK_REQUIRE(m < 0x7fffff00LL, "gfk_ksd_rowsums_mma: m too large for 32-bit tensor-map coordinates");
    GFK_REQUIRE(n_parts >= 1 && part >= 0 && part < n_parts, "gfk_ksd_rowsums_mma: bad part %d of %d", part, n_parts);
    if (!gfk_ksd_mma_supported(p)) {
        set_error("gfk_ksd_rowsums_mma: needs the float32 mode, a scalar preconditioner and d <= 512");
        return GFK_EUNSUPPORTED;
    }
    GFK_REQUIRE(reinterpret_cast<uintptr_t>(ws) % 256 == 0, "gfk_ksd_rowsums_mma: ws must be 256-byte aligned");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int d = p->d, dp = mma_pad_dims(d), nkb = dp / MM_KB;
    const long long mp = mma_pad_rows(m);
    char* w8 = static_cast<char*>(ws);
    unsigned* status = reinterpret_cast<unsigned*>(w8);                 // first 256 bytes: the status word
    double* mean = reinterpret_cast<double*>(w8 + 256);
    float* xh = reinterpret_cast<float*>(w8 + 512 + sizeof(double) * dp);
    xh = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(xh) + 255) & ~(uintptr_t)255);
    float* xl = xh + (size_t)mp * dp;
    float* sh = xl + (size_t)mp * dp;
    float* sl = sh + (size_t)mp * dp;
    float* nx = sl + (size_t)mp * dp;
    float* sxx = nx + mp;
    float* wout = sxx + mp;
    const float* x = static_cast<const float*>(p->x);
    const float* s = static_cast<const float*>(p->s);
    // the diagonal (and the zeroing convention of the row sums) is the tile kernel's
    KsdArgs<float> a = make_args<float>(p, sel, m);
    a.low = low;
    a.diag = diag;
    a.part = part;
    a.n_parts = n_parts;
    a.deal_rows = MM_M;                                  // 128-row tiles are dealt round-robin over the parts
    GFK_CUDA(cudaMemsetAsync(status, 0, 256, st));
    ksd_diag_kernel<float><<<grid1d(m), 256, 0, st>>>(a);
    ksd_mma_colmean_kernel<><<<d, 256, 0, st>>>(x, sel, m, d, mean);
    ksd_mma_split_kernel<><<<grid1d(mp * 32), 256, 0, st>>>(x, s, static_cast<const float*>(p->lw), static_cast<const float*>(p->w),
                                                           sel, m, mp, d, dp, mean, xh, xl, sh, sl, nx, sxx, wout);
    TensorMap maps[4];
    float* arrays[4] = {xh, xl, sh, sl};
    for (int i = 0; i < 4; ++i) {
        if (!make_tensor_map_2d(&maps[i], arrays[i], 4, (unsigned long long)dp, (unsigned long long)mp, MM_KB, MM_N)) {
            set_error("gfk_ksd_rowsums_mma: cuTensorMapEncodeTiled is unavailable or refused the layout");
            return GFK_EUNSUPPORTED;
        }
    }
    KsdMmaArgs ma;
    ma.nx = nx;
    ma.sxx = sxx;
    ma.w = p->lw ? wout : nullptr;
    ma.m = m;
    ma.nkb = nkb;
    ma.kp = a.kp;
    ma.low = low;
    ma.status = status;
    ma.part = part;
    ma.n_parts = n_parts;
    const size_t smem = ksd_mma_smem_bytes(nkb);
    const long long ntiles = mp / MM_M;
    const long long my_tiles = (ntiles - part + n_parts - 1) / n_parts;
    if (my_tiles > 0) {
        if (nkb <= MM_RESIDENT_MAX_KB) {
            GFK_CUDA(cudaFuncSetAttribute(ksd_mma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            ksd_mma_kernel<true><<<(unsigned)my_tiles, MM_THREADS, smem, st>>>(ma, maps[0], maps[1], maps[2], maps[3]);
        } else {
            GFK_CUDA(cudaFuncSetAttribute(ksd_mma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            ksd_mma_kernel<false><<<(unsigned)my_tiles, MM_THREADS, smem, st>>>(ma, maps[0], maps[1], maps[2], maps[3]);
        }
    }
    count_launches(my_tiles > 0 ? 4 : 3);
    GFK_CHECK_LAUNCH();
    return GFK_OK;
}

extern "C" int gfk_ksd_mma_check(const void* ws, void* stream) {
    GFK_REQUIRE(ws, "gfk_ksd_mma_check: NULL workspace");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    unsigned flag = 0;
    GFK_CUDA(cudaMemcpyAsync(&flag, ws, sizeof(flag), cudaMemcpyDeviceToHost, st));
    GFK_CUDA(cudaStreamSynchronize(st));
    if (flag) {
        set_error("gfk_ksd_rowsums_mma: a pipeline wait inside the kernel ran past its deadline; the row sums are invalid");
        return GFK_ECUDA;
    }
    return GFK_OK;
}

extern "C" int gfk_ksd_finish(const double* low, const double* diag, int64_t m, double* ks, void* stream) {
    GFK_REQUIRE(low && diag && ks && m >= 1, "gfk_ksd_finish: bad arguments");
    ksd_finish_kernel<<<1, 256, 0, static_cast<cudaStream_t>(stream)>>>(low, diag, m, ks);
    count_launches(1);
    GFK_CHECK_LAUNCH();
    return GFK_OK;
}

extern "C" int gfk_kmat(const gfk_problem* p, const uint32_t* sel, int64_t m, double* out, void* stream) {
    int rc = check_problem(p, "gfk_kmat");
    if (rc) return rc;
    GFK_REQUIRE(out && m >= 1, "gfk_kmat: bad arguments");
    GFK_REQUIRE(sel || m <= p->n, "gfk_kmat: m exceeds n");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int g = grid1d(m * m);
    if (p->mode == GFK_F64_STRICT) kmat_kernel<double, true><<<g, 256, 0, st>>>(make_args<double>(p, sel, m), out);
    else if (p->mode == GFK_F64_FAST) kmat_kernel<double, false><<<g, 256, 0, st>>>(make_args<double>(p, sel, m), out);
    else kmat_kernel<float, false><<<g, 256, 0, st>>>(make_args<float>(p, sel, m), out);
    count_launches(1);
    GFK_CHECK_LAUNCH();
    return GFK_OK;
}

extern "C" int gfk_pairs(const gfk_problem* p, const int64_t* ia, const int64_t* ib, int64_t count, double* out,
                         void* stream) {
    int rc = check_problem(p, "gfk_pairs");
    if (rc) return rc;
    GFK_REQUIRE(ia && ib && out && count >= 0, "gfk_pairs: bad arguments");
    if (count == 0) return GFK_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int g = grid1d(count);
    const long long* a64 = reinterpret_cast<const long long*>(ia);
    const long long* b64 = reinterpret_cast<const long long*>(ib);
    if (p->mode == GFK_F64_STRICT)
        pairs_kernel<double, true><<<g, 256, 0, st>>>(make_args<double>(p, nullptr, 0), a64, b64, count, out);
    else if (p->mode == GFK_F64_FAST)
        pairs_kernel<double, false><<<g, 256, 0, st>>>(make_args<double>(p, nullptr, 0), a64, b64, count, out);
    else
        pairs_kernel<float, false><<<g, 256, 0, st>>>(make_args<float>(p, nullptr, 0), a64, b64, count, out);
    count_launches(1);
    GFK_CHECK_LAUNCH();
    return GFK_OK;
}
