// proxy.cu - proxy densities q for gradient-free Stein thinning, evaluated on the device ("next").
//
//   gfk_student_t_proxy : log q and grad log q of a multivariate Student-t proxy
//                         (reference: code/notebooks/lotka_volterra/Gradient_free_Student_t.ipynb#c26-#c31,#c40)
//   gfk_kde_proxy       : log q and grad log q of a (weighted) Gaussian kernel density estimate
//                         (reference: code/notebooks/gaussian_mixture/Gaussian_mixture.ipynb#c42-#c54,
//                          code/notebooks/lotka_volterra/KDE.ipynb#c7-#c11; the report names its O(n^2) cost as the
//                          blocker for large samples, report/report.tex:746,783)
//
// The KDE is an O(n m) pair sum with the tiling of the KSD kernel: thread per evaluation point, data rows staged
// in shared memory (whitened on the way in, so a pair costs d subtractions + d FMAs + one exp), a running
// log-sum-exp maximum per point, and the data range split over gridDim.y so small m still fills the GPU; the
// splits are merged in a fixed order (run-to-run reproducible, independent of the GPU count).
#include "common.cuh"

namespace gfk {

constexpr int kKdeThreads = 128;     // evaluation points per CTA = data rows per staged tile
constexpr int kKdeMaxSplits = 64;

struct KdeArgs {
    const double* data;      // (n, d)
    const double* weights;   // (n) normalised, or nullptr for 1/n
    long long n;
    const double* points;    // (m, d)
    long long m;
    const double* whiten;    // (d, d) row-major W with W' W = H^-1: z = W x
    double log_norm;         // -0.5 log det(2 pi H)
    double* partial;         // [splits][m][D + 2]: running max, sum, weighted offset sum (whitened space)
    double* log_q;           // (m)
    double* grad_q;          // (m, d)
    int splits;
};

template <int D>
__device__ __forceinline__ void whiten_row(const double* __restrict__ W, const double* __restrict__ x, double (&z)[D]) {
#pragma unroll
    for (int k = 0; k < D; ++k) {
        double acc = 0.0;
#pragma unroll
        for (int j = 0; j < D; ++j) acc = fma(W[k * D + j], x[j], acc);
        z[k] = acc;
    }
}

template <int D>
__global__ void __launch_bounds__(kKdeThreads) kde_partial_kernel(KdeArgs a) {
    __shared__ double zt[kKdeThreads * D];
    __shared__ double wt[kKdeThreads];
    __shared__ double Wsh[D * D];
    const int tid = threadIdx.x;
    for (int e = tid; e < D * D; e += kKdeThreads) Wsh[e] = a.whiten[e];
    __syncthreads();
    const long long i = (long long)blockIdx.x * kKdeThreads + tid;
    const bool active = i < a.m;
    double zi[D];
    {
        double xi[D];
#pragma unroll
        for (int k = 0; k < D; ++k) xi[k] = active ? a.points[i * D + k] : 0.0;
        whiten_row<D>(Wsh, xi, zi);
    }
    // this CTA's share of the data rows: whole tiles, dealt contiguously
    const long long ntiles = (a.n + kKdeThreads - 1) / kKdeThreads;
    const long long per = (ntiles + a.splits - 1) / a.splits;
    const long long tile0 = (long long)blockIdx.y * per;
    const long long tile1 = min(ntiles, tile0 + per);
    const double uniform = 1.0 / (double)a.n;
    double mx = -INFINITY, s = 0.0, v[D];
#pragma unroll
    for (int k = 0; k < D; ++k) v[k] = 0.0;
    for (long long tile = tile0; tile < tile1; ++tile) {
        const long long j0 = tile * kKdeThreads;
        const int cnt = (int)min((long long)kKdeThreads, a.n - j0);
        __syncthreads();                                     // previous tile consumed
        if (tid < cnt) {
            double xj[D], zj[D];
#pragma unroll
            for (int k = 0; k < D; ++k) xj[k] = a.data[(j0 + tid) * D + k];
            whiten_row<D>(Wsh, xj, zj);
#pragma unroll
            for (int k = 0; k < D; ++k) zt[tid * D + k] = zj[k];
            wt[tid] = a.weights ? a.weights[j0 + tid] : uniform;
        }
        __syncthreads();
        if (!active) continue;
        for (int j = 0; j < cnt; ++j) {
            double off[D], q = 0.0;
#pragma unroll
            for (int k = 0; k < D; ++k) {
                off[k] = zt[j * D + k] - zi[k];
                q = fma(off[k], off[k], q);
            }
            const double e = -0.5 * q;
            if (e > mx) {                                    // rare after the first few rows
                const double sc = exp(mx - e);               // exp(-inf) = 0 on the first row
                s *= sc;
#pragma unroll
                for (int k = 0; k < D; ++k) v[k] *= sc;
                mx = e;
            }
            const double p = wt[j] * exp(e - mx);
            s += p;
#pragma unroll
            for (int k = 0; k < D; ++k) v[k] = fma(p, off[k], v[k]);
        }
    }
    if (active) {
        double* out = a.partial + ((long long)blockIdx.y * a.m + i) * (D + 2);
        out[0] = mx;
        out[1] = s;
#pragma unroll
        for (int k = 0; k < D; ++k) out[2 + k] = v[k];
    }
}

// merge the splits in index order; log q = log_norm + max + log(sum); grad log q = W' (offset sum / sum)
template <int D>
__global__ void __launch_bounds__(kKdeThreads) kde_merge_kernel(KdeArgs a) {
    const long long i = (long long)blockIdx.x * kKdeThreads + threadIdx.x;
    if (i >= a.m) return;
    double mx = -INFINITY, s = 0.0, v[D];
#pragma unroll
    for (int k = 0; k < D; ++k) v[k] = 0.0;
    for (int sp = 0; sp < a.splits; ++sp) {
        const double* in = a.partial + ((long long)sp * a.m + i) * (D + 2);
        const double pm = in[0];
        if (pm == -INFINITY) continue;                       // an empty split
        const double nm = fmax(mx, pm);
        const double sa = exp(mx - nm), sb = exp(pm - nm);
        s = s * sa + in[1] * sb;
#pragma unroll
        for (int k = 0; k < D; ++k) v[k] = v[k] * sa + in[2 + k] * sb;
        mx = nm;
    }
    a.log_q[i] = a.log_norm + mx + log(s);
#pragma unroll
    for (int k = 0; k < D; ++k) {
        double acc = 0.0;
#pragma unroll
        for (int j = 0; j < D; ++j) acc = fma(a.whiten[j * D + k], v[j], acc);    // W' v
        a.grad_q[i * D + k] = acc / s;
    }
}

static int kde_splits(long long n, long long m) {
    const long long blocks_x = (m + kKdeThreads - 1) / kKdeThreads;
    const long long ntiles = (n + kKdeThreads - 1) / kKdeThreads;
    long long want = ((long long)sm_count() * 4 + blocks_x - 1) / blocks_x;
    if (want > ntiles) want = ntiles;
    if (want > kKdeMaxSplits) want = kKdeMaxSplits;
    if (want < 1) want = 1;
    return (int)want;
}

template <int D>
static int kde_typed(KdeArgs a, cudaStream_t st) {
    const unsigned bx = (unsigned)((a.m + kKdeThreads - 1) / kKdeThreads);
    kde_partial_kernel<D><<<dim3(bx, (unsigned)a.splits), kKdeThreads, 0, st>>>(a);
    kde_merge_kernel<D><<<bx, kKdeThreads, 0, st>>>(a);
    count_launches(2);
    GFK_CHECK_LAUNCH();
    return GFK_OK;
}

// ---- multivariate Student-t: scipy.stats.multivariate_t.logpdf's expression and the notebook's score ------------
//   maha = (x - mu)' P (x - mu);  log q = cst - (df + d)/2 log(1 + maha / df);  grad = -(df + d)/df / (1 + maha/df) P (x - mu)
__global__ void __launch_bounds__(256) student_t_proxy_kernel(const double* __restrict__ x, long long n, int d,
                                                              const double* __restrict__ mean,
                                                              const double* __restrict__ prec, double df, double cst,
                                                              double* __restrict__ log_q, double* __restrict__ grad_q) {
    extern __shared__ double tp_sm[];
    double* mean_s = tp_sm;
    double* prec_s = tp_sm + d;
    for (int e = threadIdx.x; e < d; e += blockDim.x) mean_s[e] = mean[e];
    for (int e = threadIdx.x; e < d * d; e += blockDim.x) prec_s[e] = prec[e];
    __syncthreads();
    const double half_t = 0.5 * (df + (double)d);
    const double lead = -(df + (double)d) / df;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double* xi = x + i * d;
        double maha = 0.0;
        for (int k = 0; k < d; ++k) {
            double acc = 0.0;
            for (int j = 0; j < d; ++j) acc = __dadd_rn(acc, __dmul_rn(prec_s[k * d + j], __dsub_rn(xi[j], mean_s[j])));
            grad_q[i * d + k] = acc;                          // P (x - mu), scaled below
            maha = __dadd_rn(maha, __dmul_rn(__dsub_rn(xi[k], mean_s[k]), acc));
        }
        const double u = __dadd_rn(1.0, __ddiv_rn(maha, df));
        log_q[i] = __dadd_rn(cst, __dmul_rn(-half_t, log(u)));
        const double sc = __ddiv_rn(lead, u);
        for (int k = 0; k < d; ++k) grad_q[i * d + k] = __dmul_rn(sc, grad_q[i * d + k]);
    }
}

}  // namespace gfk

using namespace gfk;

extern "C" size_t gfk_kde_ws_bytes(int64_t n, int64_t m, int32_t d) {
    if (n < 1 || m < 1 || d < 1) return 0;
    return sizeof(double) * (size_t)kde_splits(n, m) * (size_t)m * (size_t)(d + 2);
}

extern "C" int gfk_kde_proxy(const double* data, const double* weights, int64_t n, const double* points, int64_t m,
                             int32_t d, const double* whiten, double log_norm, double* log_q, double* grad_q,
                             void* ws, void* stream) {
    GFK_REQUIRE(data && points && whiten && log_q && grad_q && ws, "gfk_kde_proxy: NULL pointer argument");
    GFK_REQUIRE(n >= 1 && m >= 1, "gfk_kde_proxy: need n >= 1 data rows and m >= 1 evaluation points");
    if (d < 1 || d > 8) {
        set_error("gfk_kde_proxy: d = %d is outside the supported range [1, 8]", d);
        return GFK_EUNSUPPORTED;
    }
    KdeArgs a;
    a.data = data;
    a.weights = weights;
    a.n = n;
    a.points = points;
    a.m = m;
    a.whiten = whiten;
    a.log_norm = log_norm;
    a.partial = static_cast<double*>(ws);
    a.log_q = log_q;
    a.grad_q = grad_q;
    a.splits = kde_splits(n, m);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    switch (d) {
        case 1: return kde_typed<1>(a, st);
        case 2: return kde_typed<2>(a, st);
        case 3: return kde_typed<3>(a, st);
        case 4: return kde_typed<4>(a, st);
        case 5: return kde_typed<5>(a, st);
        case 6: return kde_typed<6>(a, st);
        case 7: return kde_typed<7>(a, st);
        default: return kde_typed<8>(a, st);
    }
}

extern "C" int gfk_student_t_proxy(const double* x, int64_t n, int32_t d, const double* mean, const double* prec,
                                   double df, double cst, double* log_q, double* grad_q, void* stream) {
    GFK_REQUIRE(x && mean && prec && log_q && grad_q, "gfk_student_t_proxy: NULL pointer argument");
    GFK_REQUIRE(n >= 1 && d >= 1 && d <= 160, "gfk_student_t_proxy: need n >= 1 and 1 <= d <= 160");
    GFK_REQUIRE(df > 0.0, "gfk_student_t_proxy: df must be positive");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t smem = sizeof(double) * (size_t)(d + d * d);
    if (smem > 48 * 1024)
        GFK_CUDA(cudaFuncSetAttribute(student_t_proxy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    long long blocks = (n + 255) / 256;
    const long long cap = (long long)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    student_t_proxy_kernel<<<(unsigned)blocks, 256, smem, st>>>(x, n, d, mean, prec, df, cst, log_q, grad_q);
    count_launches(1);
    GFK_CHECK_LAUNCH();
    return GFK_OK;
}
