// energy.cu - energy distance between a validation sample and (every prefix of) a thinned sample
// ("next").
//
// Reference: code/notebooks/lotka_volterra/Comparison.ipynb#c19-#c23 -
//   fit_quality(subsample) = sqrt(dcor.energy_distance(validation_sample[::10], subsample)), evaluated from scratch
//   for each of 250 prefix lengths of the selected indices; definition report/report.tex:437-455:
//   D^2(x, y) = 2 mean_{a,t} |x_a - y_t| - mean_{a,b} |x_a - x_b| - mean_{t,u} |y_t - y_u|     (V-statistic, as dcor).
//
// All prefixes share their pair distances, so ONE pass gives the whole curve: three O(n m) / O(n^2) / O(m^2) row-sum
// launches of the same tiled kernel (thread per row, the other sample staged in shared memory, column range split
// over gridDim.y and merged in a fixed order), then a sequential scan:
//   cross[t] = sum_a |y_t - x_a|,  low[t] = sum_{u<t} |y_t - y_u|,  xx[a] = sum_{b<a} |x_a - x_b|
//   D^2[t]   = 2 C_t / (n (t+1)) - 2 sum(xx) / n^2 - 2 L_t / (t+1)^2,   C_t, L_t = running sums of cross, low.
#include "common.cuh"

namespace gfk {

constexpr int kEnThreads = 128;
constexpr int kEnMaxSplits = 64;

struct EnArgs {
    const double* p;   // rows (np, d)
    long long np;
    const double* q;   // columns (nq, d)
    long long nq;
    int triangular;    // 1: only columns j < i (p and q are the same sample)
    int splits;
    double* partial;   // [splits][np]
};

template <int D>
__global__ void __launch_bounds__(kEnThreads) dist_rowsum_kernel(EnArgs a) {
    __shared__ double qt[kEnThreads * D];
    const int tid = threadIdx.x;
    // largest row tiles first when triangular (they have the most columns)
    const long long nrt = (a.np + kEnThreads - 1) / kEnThreads;
    const long long rt = a.triangular ? (nrt - 1 - blockIdx.x) : blockIdx.x;
    const long long i = rt * kEnThreads + tid;
    const bool active = i < a.np;
    double pi[D];
#pragma unroll
    for (int k = 0; k < D; ++k) pi[k] = active ? a.p[i * D + k] : 0.0;
    const long long col_end = a.triangular ? min(a.nq, (rt + 1) * kEnThreads) : a.nq;   // columns any row of this tile needs
    const long long ntiles = (col_end + kEnThreads - 1) / kEnThreads;
    const long long per = (ntiles + a.splits - 1) / a.splits;
    const long long tile0 = (long long)blockIdx.y * per;
    const long long tile1 = min(ntiles, tile0 + per);
    const long long my_end = a.triangular ? i : a.nq;
    double acc = 0.0;
    for (long long tile = tile0; tile < tile1; ++tile) {
        const long long j0 = tile * kEnThreads;
        const int cnt = (int)min((long long)kEnThreads, col_end - j0);
        __syncthreads();
        if (tid < cnt) {
#pragma unroll
            for (int k = 0; k < D; ++k) qt[tid * D + k] = a.q[(j0 + tid) * D + k];
        }
        __syncthreads();
        if (!active) continue;
        const int lim = (int)max(0LL, min((long long)cnt, my_end - j0));
        double part = 0.0;                                   // per-tile partial keeps the long sum accurate
        for (int j = 0; j < lim; ++j) {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < D; ++k) {
                const double r = pi[k] - qt[j * D + k];
                s = fma(r, r, s);
            }
            part += sqrt(s);
        }
        acc += part;
    }
    if (active) a.partial[(long long)blockIdx.y * a.np + i] = acc;
}

// out[i] = sum over splits (index order)
__global__ void __launch_bounds__(256) rowsum_merge_kernel(const double* __restrict__ partial, long long np, int splits,
                                                           double* __restrict__ out) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < np; i += (long long)gridDim.x * blockDim.x) {
        double acc = 0.0;
        for (int s = 0; s < splits; ++s) acc += partial[(long long)s * np + i];
        out[i] = acc;
    }
}

// one block: sxx = 2 * sum(xx) in a fixed order (thread-strided partials, then a fixed tree), then the scan
__global__ void __launch_bounds__(256) energy_finish_kernel(const double* __restrict__ xx, long long n,
                                                            const double* __restrict__ cross,
                                                            const double* __restrict__ low, long long m,
                                                            double* __restrict__ e2) {
    __shared__ double red[256];
    __shared__ double buf_c[1024], buf_l[1024];
    __shared__ double carry_c, carry_l;
    double acc = 0.0;
    for (long long a = threadIdx.x; a < n; a += 256) acc += xx[a];
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
        if ((int)threadIdx.x < off) red[threadIdx.x] += red[threadIdx.x + off];
        __syncthreads();
    }
    const double mean_xx = 2.0 * red[0] / ((double)n * (double)n);
    if (threadIdx.x == 0) { carry_c = 0.0; carry_l = 0.0; }
    for (long long base = 0; base < m; base += 1024) {
        const int cnt = (int)min(1024LL, m - base);
        __syncthreads();
        for (int e = threadIdx.x; e < cnt; e += 256) { buf_c[e] = cross[base + e]; buf_l[e] = low[base + e]; }
        __syncthreads();
        if (threadIdx.x == 0) {
            double c = carry_c, l = carry_l;
            for (int e = 0; e < cnt; ++e) { c += buf_c[e]; l += buf_l[e]; buf_c[e] = c; buf_l[e] = l; }
            carry_c = c;
            carry_l = l;
        }
        __syncthreads();
        for (int e = threadIdx.x; e < cnt; e += 256) {
            const double t1 = (double)(base + e + 1);
            e2[base + e] = 2.0 * buf_c[e] / ((double)n * t1) - mean_xx - 2.0 * buf_l[e] / (t1 * t1);
        }
    }
}

static int en_splits(long long np, long long nq) {
    const long long blocks_x = (np + kEnThreads - 1) / kEnThreads;
    const long long ntiles = (nq + kEnThreads - 1) / kEnThreads;
    long long want = ((long long)sm_count() * 4 + blocks_x - 1) / blocks_x;
    if (want > ntiles) want = ntiles;
    if (want > kEnMaxSplits) want = kEnMaxSplits;
    if (want < 1) want = 1;
    return (int)want;
}

template <int D>
static void rowsums_launch(const double* p, long long np, const double* q, long long nq, int triangular,
                           double* partial, double* out, cudaStream_t st) {
    EnArgs a;
    a.p = p; a.np = np; a.q = q; a.nq = nq; a.triangular = triangular;
    a.splits = en_splits(np, nq);
    a.partial = partial;
    const unsigned bx = (unsigned)((np + kEnThreads - 1) / kEnThreads);
    dist_rowsum_kernel<D><<<dim3(bx, (unsigned)a.splits), kEnThreads, 0, st>>>(a);
    long long mb = (np + 255) / 256;
    const long long cap = (long long)sm_count() * 8;
    if (mb > cap) mb = cap;
    rowsum_merge_kernel<<<(unsigned)mb, 256, 0, st>>>(partial, np, a.splits, out);
}

template <int D>
static int energy_typed(const double* x, long long n, const double* y, long long m, double* e2, double* ws,
                        cudaStream_t st) {
    double* xx = ws;
    double* cross = xx + n;
    double* low = cross + m;
    double* partial = low + m;
    rowsums_launch<D>(x, n, x, n, 1, partial, xx, st);
    rowsums_launch<D>(y, m, x, n, 0, partial, cross, st);
    rowsums_launch<D>(y, m, y, m, 1, partial, low, st);
    energy_finish_kernel<<<1, 256, 0, st>>>(xx, n, cross, low, m, e2);
    count_launches(7);
    GFK_CHECK_LAUNCH();
    return GFK_OK;
}

}  // namespace gfk

using namespace gfk;

extern "C" size_t gfk_energy_ws_bytes(int64_t n, int64_t m) {
    if (n < 1 || m < 1) return 0;
    // xx[n], cross[m], low[m], then the split partials of the largest of the three row-sum launches
    const int64_t p_xx = (int64_t)en_splits(n, n) * n;
    const int64_t p_cross = (int64_t)en_splits(m, n) * m;
    const int64_t p_low = (int64_t)en_splits(m, m) * m;
    int64_t partial = p_xx > p_cross ? p_xx : p_cross;
    if (p_low > partial) partial = p_low;
    return sizeof(double) * (size_t)(n + 2 * m + partial);
}

extern "C" int gfk_energy_distance_cumulative(const double* x, int64_t n, const double* y, int64_t m, int32_t d,
                                              double* e2, void* ws, void* stream) {
    GFK_REQUIRE(x && y && e2 && ws, "gfk_energy_distance_cumulative: NULL pointer argument");
    GFK_REQUIRE(n >= 1 && m >= 1, "gfk_energy_distance_cumulative: need n >= 1 and m >= 1");
    if (d < 1 || d > 8) {
        set_error("gfk_energy_distance_cumulative: d = %d is outside the supported range [1, 8]", d);
        return GFK_EUNSUPPORTED;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    double* w = static_cast<double*>(ws);
    switch (d) {
        case 1: return energy_typed<1>(x, n, y, m, e2, w, st);
        case 2: return energy_typed<2>(x, n, y, m, e2, w, st);
        case 3: return energy_typed<3>(x, n, y, m, e2, w, st);
        case 4: return energy_typed<4>(x, n, y, m, e2, w, st);
        case 5: return energy_typed<5>(x, n, y, m, e2, w, st);
        case 6: return energy_typed<6>(x, n, y, m, e2, w, st);
        case 7: return energy_typed<7>(x, n, y, m, e2, w, st);
        default: return energy_typed<8>(x, n, y, m, e2, w, st);
    }
}
