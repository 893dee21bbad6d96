// prep.cu - the O(n d) passes in front of the pair kernels.
//
//   gfk_standardize   : stein_thinning.thinning._validate_and_standardize
//                       (call site code/notebooks/examples/JAX_Stein_Thinning.ipynb#c16;
//                        scale = mean absolute deviation)
//   gfk_median_pdist  : the 'med' rule of stein_thinning.kernel.make_precon
//                       (report/report.tex:323,435; call site JAX_Stein_Thinning.ipynb#c28)
//   gfk_exp_weights / gfk_to_f32 / gfk_check_finite : small element-wise helpers
#include <math.h>
#include <string.h>

#include <vector>

#include "common.cuh"

namespace gfk {

constexpr int kChunkRows = 2048;
constexpr int kRedThreads = 256;

// standardisation workspace: [nonfinite counter | 64 B] loc[d] scl[d] partial[nchunks][d]

__device__ __forceinline__ double block_sum_256(double v, double* sh) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    double r = 0.0;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int w = 0; w < kRedThreads / 32; ++w) r += sh[w];
    }
    return r;  // valid in thread 0
}

// pass A (absdev == 0): partial[c][k] = sum over the chunk of x[r][k]; counts non-finite x
// pass B (absdev == 1): partial[c][k] = sum over the chunk of |x[r][k] - loc[k]|
__global__ void __launch_bounds__(kRedThreads) col_chunk_sum_kernel(const double* __restrict__ x, long long n,
                                                                    int d, int absdev,
                                                                    const double* __restrict__ loc,
                                                                    double* __restrict__ partial,
                                                                    unsigned long long* nonfinite) {
    __shared__ double sh[kRedThreads / 32];
    const long long r0 = (long long)blockIdx.x * kChunkRows;
    const long long r1 = min(n, r0 + kChunkRows);
    unsigned bad = 0;
    for (int k = 0; k < d; ++k) {
        const double lk = absdev ? loc[k] : 0.0;
        double acc = 0.0;
        for (long long r = r0 + threadIdx.x; r < r1; r += kRedThreads) {
            const double v = x[r * d + k];
            if (!absdev && !isfinite(v)) ++bad;
            acc += absdev ? fabs(v - lk) : v;
        }
        const double tot = block_sum_256(acc, sh);
        if (threadIdx.x == 0) partial[(long long)blockIdx.x * d + k] = tot;
    }
    if (bad) atomicAdd(nonfinite, (unsigned long long)bad);
}

// out[k] = (sum of the chunk partials, fixed order) / n
// (one warp per column: lane l adds partials l, l+32, ... in index order, then a fixed shuffle tree)
__global__ void col_finish_kernel(const double* __restrict__ partial, int nchunks, int d, long long n,
                                  double* __restrict__ out) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    for (int k = warp; k < d; k += nwarps) {
        double acc = 0.0;
        for (int c = lane; c < nchunks; c += 32) acc += partial[(long long)c * d + k];
        acc = warp_sum(acc);
        if (lane == 0) out[k] = acc / (double)n;
    }
}

// x_out = x / scl, g_out = g * scl (IEEE division / multiplication, as numpy); counts non-finite g
__global__ void __launch_bounds__(256) scale_kernel(const double* __restrict__ x, const double* __restrict__ g,
                                                    long long count, int d, const double* __restrict__ scl,
                                                    int standardize, double* __restrict__ x_out,
                                                    double* __restrict__ g_out, unsigned long long* nonfinite) {
    unsigned bad = 0;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < count;
         e += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(e % d);
        const double xv = x[e], gv = g[e];
        if (!isfinite(gv)) ++bad;
        if (standardize) {
            const double sc = scl[k];
            x_out[e] = __ddiv_rn(xv, sc);
            g_out[e] = __dmul_rn(gv, sc);
        } else {
            x_out[e] = xv;
            g_out[e] = gv;
        }
    }
    if (bad) atomicAdd(nonfinite, (unsigned long long)bad);
}

// ---- median of pairwise distances ----------------------------------------------------------------
// dist[p] for the p-th pair (i < j) in condensed order, scipy's euclidean: s += (u-v)*(u-v); sqrt(s)
__global__ void __launch_bounds__(256) pdist_kernel(const double* __restrict__ x, const long long* __restrict__ rows,
                                                    int R, int d, double* __restrict__ dist) {
    const long long npairs = (long long)R * (R - 1) / 2;
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < npairs;
         p += (long long)gridDim.x * blockDim.x) {
        // invert p -> (i, j): pairs before row i: i*R - i*(i+1)/2
        long long i = (long long)floor(((2.0 * R - 1.0) - sqrt((2.0 * R - 1.0) * (2.0 * R - 1.0) - 8.0 * (double)p)) * 0.5);
        while (i > 0 && i * R - i * (i + 1) / 2 > p) --i;
        while ((i + 1) * R - (i + 1) * (i + 2) / 2 <= p) ++i;
        const long long j = p - (i * R - i * (i + 1) / 2) + i + 1;
        const double* u = x + rows[i] * d;
        const double* v = x + rows[j] * d;
        double s = 0.0;
        for (int k = 0; k < d; ++k) {
            const double t = __dsub_rn(u[k], v[k]);
            s = __dadd_rn(s, __dmul_rn(t, t));
        }
        dist[p] = __dsqrt_rn(s);
    }
}

// ---- exact order statistics by radix select on the bit patterns of non-negative doubles ---------------
// Six passes of 11 bits (the last 9): every pass histograms the keys that match the prefix found so far
// (per-block shared-memory histograms merged with global atomics, all SMs busy), then one block scans the
// 2048 bins to extend the prefix.  Two ranks (the two middle order statistics) are tracked together.
constexpr int kSelBits = 11;
constexpr int kSelBins = 1 << kSelBits;
struct SelState {
    unsigned long long prefix[2];
    long long rank[2];
};

__global__ void __launch_bounds__(256) select_hist_kernel(const double* __restrict__ dist, long long count, int shift,
                                                          int width, const SelState* __restrict__ state,
                                                          unsigned* __restrict__ hist /* [2][kSelBins] */) {
    __shared__ unsigned sh[2][kSelBins];
    for (int b = threadIdx.x; b < 2 * kSelBins; b += blockDim.x) (&sh[0][0])[b] = 0;
    __syncthreads();
    const unsigned long long* keys = reinterpret_cast<const unsigned long long*>(dist);
    const unsigned long long himask = (shift + width >= 64) ? 0ull : (~0ull << (shift + width));
    const unsigned long long p0 = state->prefix[0], p1 = state->prefix[1];
    const unsigned mask = (1u << width) - 1u;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < count;
         e += (long long)gridDim.x * blockDim.x) {
        const unsigned long long key = keys[e];
        const unsigned bin = (unsigned)(key >> shift) & mask;
        if ((key & himask) == p0) atomicAdd(&sh[0][bin], 1u);
        if ((key & himask) == p1) atomicAdd(&sh[1][bin], 1u);
    }
    __syncthreads();
    for (int b = threadIdx.x; b < 2 * kSelBins; b += blockDim.x) {
        const unsigned v = (&sh[0][0])[b];
        if (v) atomicAdd(&hist[b], v);
    }
}

// one block: for each rank, the bin whose cumulative count first exceeds the rank; clears hist for the next pass
__global__ void __launch_bounds__(1024) select_pick_kernel(unsigned* __restrict__ hist, int shift, SelState* state,
                                                           double* __restrict__ out, int last) {
    __shared__ unsigned long long partial[1024];
    for (int r = 0; r < 2; ++r) {
        unsigned* h = hist + r * kSelBins;
        // thread t owns bins 2t, 2t+1
        const unsigned c0 = h[2 * threadIdx.x], c1 = h[2 * threadIdx.x + 1];
        partial[threadIdx.x] = (unsigned long long)c0 + c1;
        __syncthreads();
        for (int off = 1; off < 1024; off <<= 1) {          // inclusive scan
            const unsigned long long add = (threadIdx.x >= off) ? partial[threadIdx.x - off] : 0ull;
            __syncthreads();
            partial[threadIdx.x] += add;
            __syncthreads();
        }
        const long long before = (long long)(partial[threadIdx.x] - c0 - c1);   // keys in bins < 2t
        const long long rank = state->rank[r];
        __syncthreads();
        if (rank >= before && rank < before + (long long)c0 + c1) {
            const int bin = (rank < before + (long long)c0) ? 2 * threadIdx.x : 2 * threadIdx.x + 1;
            state->rank[r] = rank - before - (bin & 1 ? (long long)c0 : 0);
            state->prefix[r] |= (unsigned long long)bin << shift;
            if (last) out[r] = __longlong_as_double((long long)state->prefix[r]);
        }
        h[2 * threadIdx.x] = 0;
        h[2 * threadIdx.x + 1] = 0;
        __syncthreads();
    }
}

__global__ void select_init_kernel(SelState* state, unsigned* hist, long long rank0, long long rank1) {
    for (int b = threadIdx.x; b < 2 * kSelBins; b += blockDim.x) hist[b] = 0;
    if (threadIdx.x == 0) {
        state->prefix[0] = state->prefix[1] = 0ull;
        state->rank[0] = rank0;
        state->rank[1] = rank1;
    }
}

__global__ void to_f32_kernel(const double* __restrict__ src, float* __restrict__ dst, long long count) {
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < count;
         e += (long long)gridDim.x * blockDim.x)
        dst[e] = (float)src[e];
}

template <typename T>
__global__ void exp_weights_kernel(const T* __restrict__ lw, T* __restrict__ w, long long count) {
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < count;
         e += (long long)gridDim.x * blockDim.x)
        w[e] = (T)exp((double)lw[e]);
}

// ew[2 i], ew[2 i + 1] = double-double exp(lw_i) for the strict-mode weights (stein_math.cuh exp_dd); rows with
// |lw| beyond kExpDdLimit are never used this way (the host checks the range first) and get NaN markers
__global__ void exp_weights_dd_kernel(const double* __restrict__ lw, double* __restrict__ ew, long long count) {
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < count;
         e += (long long)gridDim.x * blockDim.x) {
        const double l = lw[e];
        double hi, lo;
        if (fabs(l) <= kExpDdLimit) {
            exp_dd(l, hi, lo);
        } else {
            hi = __longlong_as_double(0x7ff8000000000000LL);
            lo = hi;
        }
        ew[2 * e] = hi;
        ew[2 * e + 1] = lo;
    }
}

// lw = log_q - log_p; optional range cap: lw = max(lw - max(lw), -cap).  Two launches: (min, max, NaN) then apply.
__global__ void __launch_bounds__(256) lw_minmax_kernel(const double* __restrict__ log_p,
                                                        const double* __restrict__ log_q, long long n,
                                                        double* __restrict__ stats /* [gridDim][3] */) {
    __shared__ double sh_min[8], sh_max[8], sh_bad[8];
    double mn = INFINITY, mx = -INFINITY, bad = 0.0;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
        const double a = log_p ? log_p[e] : 0.0, b = log_q[e];     // log_p == nullptr: statistics of log_q itself
        if (isnan(a) || isnan(b)) bad += 1.0;
        const double v = __dsub_rn(b, a);
        mn = fmin(mn, v);
        mx = fmax(mx, v);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, off));
        mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, off));
        bad += __shfl_xor_sync(0xffffffffu, bad, off);
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { sh_min[warp] = mn; sh_max[warp] = mx; sh_bad[warp] = bad; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) { mn = fmin(mn, sh_min[w]); mx = fmax(mx, sh_max[w]); bad += sh_bad[w]; }
        stats[blockIdx.x * 3 + 0] = mn;
        stats[blockIdx.x * 3 + 1] = mx;
        stats[blockIdx.x * 3 + 2] = bad;
    }
}

__global__ void lw_apply_kernel(const double* __restrict__ log_p, const double* __restrict__ log_q, long long n,
                                int use_cap, double mx, double cap, double* __restrict__ lw) {
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
        double v = __dsub_rn(log_q[e], log_p[e]);
        if (use_cap) v = fmax(__dsub_rn(v, mx), -cap);
        lw[e] = v;
    }
}

// ---- Gaussian proxy of gaussian_thin (reference: code/src/thinning.py:15-16) -----------------
// log_q[i] = -0.5 (cst + (x_i - mean)' P (x_i - mean)),  grad_q[i] = -P (x_i - mean),  P = cov^-1.
// Thread per row; mean and P are staged in shared memory; sums run over j in index order without contraction.
__global__ void __launch_bounds__(256) gaussian_proxy_kernel(const double* __restrict__ x, long long n, int d,
                                                             const double* __restrict__ mean,
                                                             const double* __restrict__ prec, double cst,
                                                             double* __restrict__ log_q, double* __restrict__ grad_q) {
    extern __shared__ double gp_sm[];
    double* mean_s = gp_sm;
    double* prec_s = gp_sm + d;
    for (int e = threadIdx.x; e < d; e += blockDim.x) mean_s[e] = mean[e];
    for (int e = threadIdx.x; e < d * d; e += blockDim.x) prec_s[e] = prec[e];
    __syncthreads();
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double* xi = x + i * d;
        double maha = 0.0;
        for (int k = 0; k < d; ++k) {
            double acc = 0.0;
            for (int j = 0; j < d; ++j) acc = __dadd_rn(acc, __dmul_rn(prec_s[k * d + j], __dsub_rn(xi[j], mean_s[j])));
            grad_q[i * d + k] = -acc;
            maha = __dadd_rn(maha, __dmul_rn(__dsub_rn(xi[k], mean_s[k]), acc));
        }
        log_q[i] = __dmul_rn(-0.5, __dadd_rn(cst, maha));
    }
}

// ---- 'smpcov' preconditioner (stein_thinning.kernel.make_precon; parity unpinned) ----
// np.cov(sample, rowvar=False): c = x - mean;  cov = c' c / (n - 1).  One block per 2048-row chunk and entry (a, b),
// a <= b, of the chunk's partial; the chunk partials are added by col_finish_kernel in its fixed order.
__global__ void __launch_bounds__(kRedThreads) cov_chunk_kernel(const double* __restrict__ x, long long n, int d,
                                                                const double* __restrict__ loc,
                                                                double* __restrict__ partial /* [nchunks][d * d] */) {
    __shared__ double sh[kRedThreads / 32];
    const long long r0 = (long long)blockIdx.x * kChunkRows;
    const long long r1 = min(n, r0 + kChunkRows);
    for (int a = 0; a < d; ++a) {
        const double la = loc[a];
        for (int b = a; b < d; ++b) {
            const double lb = loc[b];
            double acc = 0.0;
            for (long long r = r0 + threadIdx.x; r < r1; r += kRedThreads) acc = fma(x[r * d + a] - la, x[r * d + b] - lb, acc);
            const double tot = block_sum_256(acc, sh);
            if (threadIdx.x == 0) {
                partial[(long long)blockIdx.x * d * d + a * d + b] = tot;
                partial[(long long)blockIdx.x * d * d + b * d + a] = tot;
            }
        }
    }
}

// out = x @ L for a dense d x d preconditioner L (row-major, device): thread per row, L in shared memory
__global__ void __launch_bounds__(256) rows_times_matrix_kernel(const double* __restrict__ x, long long n, int d,
                                                                const double* __restrict__ L, double* __restrict__ out) {
    extern __shared__ double rm_sm[];
    for (int e = threadIdx.x; e < d * d; e += blockDim.x) rm_sm[e] = L[e];
    __syncthreads();
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double* xi = x + i * d;
        for (int k = 0; k < d; ++k) {
            double acc = 0.0;
            for (int j = 0; j < d; ++j) acc = fma(xi[j], rm_sm[j * d + k], acc);
            out[i * d + k] = acc;
        }
    }
}

static int grid_for(long long work, int threads) {
    long long b = (work + threads - 1) / threads;
    const long long cap = (long long)sm_count() * 8;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (int)b;
}

}  // namespace gfk

using namespace gfk;

extern "C" size_t gfk_standardize_ws_bytes(int64_t n, int32_t d) {
    const int64_t nchunks = (n + kChunkRows - 1) / kChunkRows;
    return 64 + sizeof(double) * (size_t)(2 * (int64_t)d + nchunks * d);
}

extern "C" int gfk_standardize(const double* x, const double* g, int64_t n, int32_t d, int standardize,
                               double* x_out, double* g_out, double* scl_host, void* ws, void* stream) {
    GFK_REQUIRE(x && g && x_out && g_out && scl_host && ws, "gfk_standardize: NULL pointer argument");
    GFK_REQUIRE(n >= 1 && d >= 1, "gfk_standardize: empty sample (n = %lld, d = %d)", (long long)n, d);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    unsigned long long* nonfinite = reinterpret_cast<unsigned long long*>(ws);
    double* loc = reinterpret_cast<double*>(reinterpret_cast<char*>(ws) + 64);
    double* scl = loc + d;
    double* partial = scl + d;
    const int nchunks = (int)((n + kChunkRows - 1) / kChunkRows);
    GFK_CUDA(cudaMemsetAsync(nonfinite, 0, 64, st));
    col_chunk_sum_kernel<<<nchunks, kRedThreads, 0, st>>>(x, n, d, 0, nullptr, partial, nonfinite);
    col_finish_kernel<<<1, 128, 0, st>>>(partial, nchunks, d, n, loc);
    if (standardize) {
        col_chunk_sum_kernel<<<nchunks, kRedThreads, 0, st>>>(x, n, d, 1, loc, partial, nonfinite);
        col_finish_kernel<<<1, 128, 0, st>>>(partial, nchunks, d, n, scl);
    }
    scale_kernel<<<grid_for(n * d, 256), 256, 0, st>>>(x, g, n * d, d, scl, standardize, x_out, g_out, nonfinite);
    count_launches(standardize ? 5 : 3);
    GFK_CHECK_LAUNCH();
    unsigned long long bad = 0;
    GFK_CUDA(cudaMemcpyAsync(&bad, nonfinite, sizeof(bad), cudaMemcpyDeviceToHost, st));
    if (standardize) {
        GFK_CUDA(cudaMemcpyAsync(scl_host, scl, sizeof(double) * d, cudaMemcpyDeviceToHost, st));
    }
    GFK_CUDA(cudaStreamSynchronize(st));
    if (!standardize) for (int k = 0; k < d; ++k) scl_host[k] = 1.0;
    if (bad) {
        set_error("sample or gradient contains NaNs or infs (%llu entries)", bad);
        return GFK_ENUMERIC;
    }
    if (standardize) {
        for (int k = 0; k < d; ++k) {
            if (!(scl_host[k] > 0.0)) {
                set_error("Too few unique samples in sample (column %d has zero mean absolute deviation).", k);
                return GFK_ENUMERIC;
            }
        }
    }
    return GFK_OK;
}

extern "C" size_t gfk_covariance_ws_bytes(int64_t n, int32_t d) {
    const int64_t nchunks = (n + kChunkRows - 1) / kChunkRows;
    return 64 + sizeof(double) * (size_t)((int64_t)d + (int64_t)d * d + nchunks * d * d);
}

extern "C" int gfk_covariance(const double* x, int64_t n, int32_t d, double* cov_host, void* ws, void* stream) {
    GFK_REQUIRE(x && cov_host && ws, "gfk_covariance: NULL pointer argument");
    GFK_REQUIRE(n >= 2 && d >= 1, "gfk_covariance: need n >= 2 and d >= 1");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    double* loc = reinterpret_cast<double*>(reinterpret_cast<char*>(ws) + 64);
    double* cov = loc + d;
    double* partial = cov + (size_t)d * d;
    const int nchunks = (int)((n + kChunkRows - 1) / kChunkRows);
    unsigned long long* nonfinite = reinterpret_cast<unsigned long long*>(ws);
    GFK_CUDA(cudaMemsetAsync(nonfinite, 0, 64, st));
    col_chunk_sum_kernel<<<nchunks, kRedThreads, 0, st>>>(x, n, d, 0, nullptr, partial, nonfinite);
    col_finish_kernel<<<1, 128, 0, st>>>(partial, nchunks, d, n, loc);
    cov_chunk_kernel<<<nchunks, kRedThreads, 0, st>>>(x, n, d, loc, partial);
    // the same fixed-order finish over d * d "columns", dividing by n - 1 (np.cov's default ddof = 1)
    col_finish_kernel<<<1, 128, 0, st>>>(partial, nchunks, d * d, n - 1, cov);
    count_launches(4);
    GFK_CHECK_LAUNCH();
    GFK_CUDA(cudaMemcpyAsync(cov_host, cov, sizeof(double) * (size_t)d * d, cudaMemcpyDeviceToHost, st));
    GFK_CUDA(cudaStreamSynchronize(st));
    return GFK_OK;
}

extern "C" int gfk_rows_times_matrix(const double* x, int64_t n, int32_t d, const double* L, double* out, void* stream) {
    GFK_REQUIRE(x && L && out, "gfk_rows_times_matrix: NULL pointer argument");
    GFK_REQUIRE(n >= 1 && d >= 1 && d <= 160, "gfk_rows_times_matrix: need n >= 1 and 1 <= d <= 160");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t smem = sizeof(double) * (size_t)d * d;
    if (smem > 48 * 1024)
        GFK_CUDA(cudaFuncSetAttribute(rows_times_matrix_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    rows_times_matrix_kernel<<<grid_for(n, 256), 256, smem, st>>>(x, n, d, L, out);
    count_launches(1);
    GFK_CHECK_LAUNCH();
    return GFK_OK;
}

extern "C" size_t gfk_median_ws_bytes(int64_t n, int32_t cap) {
    const int64_t R = n > cap ? cap : n;
    return 64 + sizeof(SelState) + sizeof(unsigned) * 2 * kSelBins + sizeof(long long) * (size_t)R +
           sizeof(double) * (size_t)(R * (R - 1) / 2 + 2);
}

extern "C" int gfk_median_pdist(const double* x, int64_t n, int32_t d, int32_t cap, double* mid_host, void* ws,
                                void* stream) {
    GFK_REQUIRE(x && mid_host && ws, "gfk_median_pdist: NULL pointer argument");
    GFK_REQUIRE(n >= 2 && d >= 1 && cap >= 2, "gfk_median_pdist: need n >= 2, d >= 1, cap >= 2");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int R = (int)(n > cap ? cap : n);
    // rows = np.linspace(0, n - 1, cap, dtype=int) when n > cap (numpy: arange(num) * step, last = stop)
    std::vector<long long> rows(R);
    if (n > cap) {
        const double step = (double)(n - 1) / (double)(cap - 1);
        for (int a = 0; a < R; ++a) rows[a] = (long long)((double)a * step);
        rows[R - 1] = n - 1;
    } else {
        for (int a = 0; a < R; ++a) rows[a] = a;
    }
    double* out2 = reinterpret_cast<double*>(ws);  // two order statistics
    SelState* state = reinterpret_cast<SelState*>(reinterpret_cast<char*>(ws) + 64);
    unsigned* hist = reinterpret_cast<unsigned*>(state + 1);
    long long* rows_dev = reinterpret_cast<long long*>(hist + 2 * kSelBins);
    double* dist = reinterpret_cast<double*>(rows_dev + R);
    // keep dist 8-byte aligned (rows_dev is 8-byte elements at a 64-byte offset)
    GFK_CUDA(cudaMemcpyAsync(rows_dev, rows.data(), sizeof(long long) * R, cudaMemcpyHostToDevice, st));
    const long long npairs = (long long)R * (R - 1) / 2;
    pdist_kernel<<<grid_for(npairs, 256), 256, 0, st>>>(x, rows_dev, R, d, dist);
    // np.median: mean of the two middle order statistics (equal ranks when the count is odd)
    const long long k1 = npairs / 2;
    const long long k0 = (npairs % 2 == 0) ? k1 - 1 : k1;
    select_init_kernel<<<1, 1024, 0, st>>>(state, hist, k0, k1);
    int launches = 2;
    for (int hi = 64; hi > 0; hi -= kSelBits) {
        const int shift = hi - kSelBits > 0 ? hi - kSelBits : 0;
        const int width = hi - shift;
        select_hist_kernel<<<grid_for(npairs, 256), 256, 0, st>>>(dist, npairs, shift, width, state, hist);
        select_pick_kernel<<<1, 1024, 0, st>>>(hist, shift, state, out2, shift == 0);
        launches += 2;
    }
    count_launches(launches);
    GFK_CHECK_LAUNCH();
    GFK_CUDA(cudaMemcpyAsync(mid_host, out2, 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
    GFK_CUDA(cudaStreamSynchronize(st));  // also keeps `rows` alive until the copy has been consumed
    return GFK_OK;
}

extern "C" int gfk_to_f32(const double* src, float* dst, int64_t count, void* stream) {
    GFK_REQUIRE(src && dst && count >= 0, "gfk_to_f32: bad arguments");
    if (count == 0) return GFK_OK;
    to_f32_kernel<<<grid_for(count, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(src, dst, count);
    count_launches(1);
    GFK_CHECK_LAUNCH();
    return GFK_OK;
}

extern "C" int gfk_exp_weights(const void* lw, void* w, int64_t count, int is_f32, void* stream) {
    GFK_REQUIRE(lw && w && count >= 0, "gfk_exp_weights: bad arguments");
    if (count == 0) return GFK_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (is_f32)
        exp_weights_kernel<float><<<grid_for(count, 256), 256, 0, st>>>(static_cast<const float*>(lw),
                                                                        static_cast<float*>(w), count);
    else
        exp_weights_kernel<double><<<grid_for(count, 256), 256, 0, st>>>(static_cast<const double*>(lw),
                                                                         static_cast<double*>(w), count);
    count_launches(1);
    GFK_CHECK_LAUNCH();
    return GFK_OK;
}

extern "C" int gfk_exp_weights_dd(const double* lw, double* ew, int64_t count, void* stream) {
    GFK_REQUIRE(lw && ew && count >= 0, "gfk_exp_weights_dd: bad arguments");
    GFK_REQUIRE(reinterpret_cast<uintptr_t>(ew) % 16 == 0, "gfk_exp_weights_dd: ew must be 16-byte aligned");
    if (count == 0) return GFK_OK;
    exp_weights_dd_kernel<<<grid_for(count, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(lw, ew, count);
    count_launches(1);
    GFK_CHECK_LAUNCH();
    return GFK_OK;
}

// range of the log-weights written by this thread's most recent gfk_log_weights call (host bookkeeping only)
static thread_local double g_last_lw_min = 0.0, g_last_lw_max = 0.0;
static thread_local bool g_last_lw_valid = false;

extern "C" int gfk_log_weights_last_range(double* min_max_host) {
    GFK_REQUIRE(min_max_host, "gfk_log_weights_last_range: NULL output");
    GFK_REQUIRE(g_last_lw_valid, "gfk_log_weights_last_range: no gfk_log_weights call on this thread yet");
    min_max_host[0] = g_last_lw_min;
    min_max_host[1] = g_last_lw_max;
    return GFK_OK;
}

extern "C" size_t gfk_log_weights_ws_bytes(void) { return sizeof(double) * 3 * 1024; }

extern "C" int gfk_log_weights(const double* log_p, const double* log_q, int64_t n, int use_cap, double range_cap,
                               double* lw, double* spread_host, void* ws, void* stream) {
    GFK_REQUIRE(log_p && log_q && lw && spread_host && ws && n >= 1, "gfk_log_weights: bad arguments");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int grid = grid_for(n, 256);
    if (grid > 1024) grid = 1024;
    double* stats = static_cast<double*>(ws);
    lw_minmax_kernel<<<grid, 256, 0, st>>>(log_p, log_q, n, stats);
    GFK_CHECK_LAUNCH();
    std::vector<double> h(3 * (size_t)grid);
    GFK_CUDA(cudaMemcpyAsync(h.data(), stats, sizeof(double) * h.size(), cudaMemcpyDeviceToHost, st));
    GFK_CUDA(cudaStreamSynchronize(st));
    double mn = INFINITY, mx = -INFINITY, bad = 0.0;
    for (int b = 0; b < grid; ++b) {
        mn = fmin(mn, h[3 * b]);
        mx = fmax(mx, h[3 * b + 1]);
        bad += h[3 * b + 2];
    }
    if (bad > 0.0) {
        set_error("log-densities contain NaNs (%g entries)", bad);
        return GFK_ENUMERIC;
    }
    *spread_host = mx - mn;
    // what lw_apply_kernel is about to write: lw itself, or max(lw - mx, -cap) (monotone, so the extremes map to extremes)
    g_last_lw_min = use_cap ? fmax(mn - mx, -range_cap) : mn;
    g_last_lw_max = use_cap ? fmax(0.0, -range_cap) : mx;
    g_last_lw_valid = true;
    lw_apply_kernel<<<grid_for(n, 256), 256, 0, st>>>(log_p, log_q, n, use_cap, mx, range_cap, lw);
    count_launches(2);
    GFK_CHECK_LAUNCH();
    return GFK_OK;
}

extern "C" int gfk_log_weights_range(const double* lw, int64_t n, double* min_max_host, void* ws, void* stream) {
    GFK_REQUIRE(lw && min_max_host && ws && n >= 1, "gfk_log_weights_range: bad arguments");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int grid = grid_for(n, 256);
    if (grid > 1024) grid = 1024;
    double* stats = static_cast<double*>(ws);
    lw_minmax_kernel<<<grid, 256, 0, st>>>(nullptr, lw, n, stats);
    count_launches(1);
    GFK_CHECK_LAUNCH();
    std::vector<double> h(3 * (size_t)grid);
    GFK_CUDA(cudaMemcpyAsync(h.data(), stats, sizeof(double) * h.size(), cudaMemcpyDeviceToHost, st));
    GFK_CUDA(cudaStreamSynchronize(st));
    double mn = INFINITY, mx = -INFINITY;
    for (int b = 0; b < grid; ++b) {
        mn = fmin(mn, h[3 * b]);
        mx = fmax(mx, h[3 * b + 1]);
    }
    min_max_host[0] = mn;
    min_max_host[1] = mx;
    return GFK_OK;
}

extern "C" int gfk_gaussian_proxy(const double* x, int64_t n, int32_t d, const double* mean, const double* prec,
                                  double cst, double* log_q, double* grad_q, void* stream) {
    GFK_REQUIRE(x && mean && prec && log_q && grad_q, "gfk_gaussian_proxy: NULL pointer argument");
    GFK_REQUIRE(n >= 1 && d >= 1 && d <= 160, "gfk_gaussian_proxy: need n >= 1 and 1 <= d <= 160");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t smem = sizeof(double) * (size_t)(d + d * d);
    if (smem > 48 * 1024)
        GFK_CUDA(cudaFuncSetAttribute(gaussian_proxy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    gaussian_proxy_kernel<<<grid_for(n, 256), 256, smem, st>>>(x, n, d, mean, prec, cst, log_q, grad_q);
    count_launches(1);
    GFK_CHECK_LAUNCH();
    return GFK_OK;
}
