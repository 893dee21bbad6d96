// thin.cu - greedy Stein thinning: fused kernel-column update + running objective + argmin.
//
// Replaces stein_thinning.thinning._greedy_search driven by the integrand closures of
// _make_stein_integrand / _make_stein_gf_integrand (reference restatement:
// code/notebooks/examples/JAX_Stein_Thinning.ipynb#c17,#c22; algorithm report/report.tex:413-426).
//
// Data in HBM (caller owned): the standardised replica x, s (n x d row-major), optional
// log-weights lw (n), and this rank's slice of the running objective A (rows x 1, float64).
// One launch per greedy step: every thread first reduces the <= 16 (value, index) candidates
// of the previous exchange to the selected row y, then streams its candidate rows once
// (algorithmic traffic 16d + 16 (+8) bytes per row in float64), adds 2 k(x_i, y) to A_i and
// keeps a running (min, first index).  Warp shuffle -> block -> last-block-done grid reduction
// produces the rank's candidate for the next exchange.  No host round trip inside the loop.
#include "common.cuh"
#include "comm.h"

namespace gfk {

constexpr int kMaxRanks = 16;
constexpr int kMaxBlocks = 148 * 16;
constexpr int kThreads = 256;

struct __align__(16) ThinWs {
    Cand send;               // this rank's (value, global index)
    Cand recv[kMaxRanks];    // all ranks' candidates after the exchange
    Cand partial[kMaxBlocks];
    unsigned counter;        // last-block-done ticket
    unsigned step;           // greedy step whose winner is in recv[]
};

template <typename T>
struct ThinArgs {
    const T* x;
    const T* s;
    const T* lw;    // log-weights (nullptr: score-based kernel)
    const T* w;     // exp(lw) (fast modes; nullptr: exp(lw_i + lw_y) on the fly)
    const T* xl;    // x @ L for a dense preconditioner (nullptr: scalar)
    double* A;
    long long row_begin, row_end;
    KParams kp;
    ThinWs* ws;
    uint32_t* idx_out;
    int n_ranks;
};

// ---- row access -------------------------------------------------------------------------------
// One row of D values with the widest aligned vector loads the row size allows (rows of a
// torch allocation start 16-byte aligned whenever D * sizeof(T) is a multiple of 16).
template <int D>
__device__ __forceinline__ void load_row(const double* __restrict__ base, long long row, double (&out)[D]) {
    const double* p = base + row * D;
    if constexpr (D % 2 == 0) {
        const double2* q = reinterpret_cast<const double2*>(p);
#pragma unroll
        for (int v = 0; v < D / 2; ++v) {
            const double2 u = __ldg(q + v);
            out[2 * v] = u.x;
            out[2 * v + 1] = u.y;
        }
    } else {
#pragma unroll
        for (int k = 0; k < D; ++k) out[k] = __ldg(p + k);
    }
}

template <int D>
__device__ __forceinline__ void load_row(const float* __restrict__ base, long long row, float (&out)[D]) {
    const float* p = base + row * D;
    if constexpr (D % 4 == 0) {
        const float4* q = reinterpret_cast<const float4*>(p);
#pragma unroll
        for (int v = 0; v < D / 4; ++v) {
            const float4 u = __ldg(q + v);
            out[4 * v] = u.x; out[4 * v + 1] = u.y; out[4 * v + 2] = u.z; out[4 * v + 3] = u.w;
        }
    } else if constexpr (D % 2 == 0) {
        const float2* q = reinterpret_cast<const float2*>(p);
#pragma unroll
        for (int v = 0; v < D / 2; ++v) {
            const float2 u = __ldg(q + v);
            out[2 * v] = u.x;
            out[2 * v + 1] = u.y;
        }
    } else {
#pragma unroll
        for (int k = 0; k < D; ++k) out[k] = __ldg(p + k);
    }
}

// strided view over shared / global memory for the runtime-d paths
template <typename T>
struct View {
    const T* p;
    __device__ __forceinline__ T operator[](int k) const { return p[k]; }
};

__device__ __forceinline__ Cand pick_winner(const ThinWs* ws, int n_ranks) {
    Cand best = ws->recv[0];
    for (int r = 1; r < n_ranks; ++r) {
        const Cand c = ws->recv[r];
        if (cand_better(c.v, c.i, best.v, best.i)) best = c;
    }
    return best;
}

// grid-level argmin: block result -> partial[], last block reduces partial[] -> ws->send
// (single rank: the exchange is the identity, so the result goes straight into recv[0])
__device__ __forceinline__ void grid_argmin_publish(double v, unsigned i, ThinWs* ws, int n_ranks) {
    __shared__ double sv[32];
    __shared__ unsigned si[32];
    __shared__ bool is_last;
    block_argmin(v, i, sv, si);
    if (threadIdx.x == 0) {
        ws->partial[blockIdx.x].v = v;
        ws->partial[blockIdx.x].i = i;
        __threadfence();
        const unsigned ticket = atomicAdd(&ws->counter, 1u);
        is_last = (ticket == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    double bv = __longlong_as_double(0x7ff0000000000000LL);
    unsigned bi = 0xffffffffu;
    for (unsigned b = threadIdx.x; b < gridDim.x; b += blockDim.x) {
        const double pv = __ldcg(&ws->partial[b].v);
        const unsigned pi = (unsigned)__ldcg(&ws->partial[b].i);
        if (cand_better(pv, pi, bv, bi)) { bv = pv; bi = pi; }
    }
    __syncthreads();
    block_argmin(bv, bi, sv, si);
    if (threadIdx.x == 0) {
        ws->send.v = bv;
        ws->send.i = bi;
        if (n_ranks == 1) {
            ws->recv[0].v = bv;
            ws->recv[0].i = bi;
        }
        ws->counter = 0;
    }
}

template <typename T, bool STRICT>
__device__ __forceinline__ double weight_pair(const ThinArgs<T>& a, long long i, T lwy, T wy) {
    if (STRICT || a.w == nullptr) return exp((double)__ldg(a.lw + i) + (double)lwy);
    return (double)(__ldg(a.w + i) * wy);
}

// ---- step 0: A_i = k(x_i, x_i) (generic runtime-d, not on the critical path) -----------------------
template <typename T, bool STRICT>
__global__ void __launch_bounds__(kThreads) thin_init_kernel(ThinArgs<T> a) {
    const int d = a.kp.d;
    double bv = __longlong_as_double(0x7ff0000000000000LL);
    unsigned bi = 0xffffffffu;
    const long long rows = a.row_end - a.row_begin;
    for (long long li = (long long)blockIdx.x * blockDim.x + threadIdx.x; li < rows;
         li += (long long)gridDim.x * blockDim.x) {
        const long long i = a.row_begin + li;
        View<T> xv{a.x + i * d}, sv{a.s + i * d};
        double k;
        if constexpr (STRICT) {
            k = pair_strict<0>(xv, xv, sv, sv, a.kp);
        } else {
            if (a.xl) {
                View<T> xlv{a.xl + i * d};
                k = (double)pair_fast<T, 0, true>(xv, xv, sv, sv, xlv, xlv, a.kp);
            } else {
                k = (double)pair_fast<T, 0, false>(xv, xv, sv, sv, xv, xv, a.kp);
            }
        }
        if (a.lw) {
            const T lwi = __ldg(a.lw + i);
            const T wi = (!STRICT && a.w) ? __ldg(a.w + i) : T(0);
            k = STRICT ? __dmul_rn(weight_pair<T, STRICT>(a, i, lwi, wi), k)
                       : weight_pair<T, STRICT>(a, i, lwi, wi) * k;
        }
        a.A[li] = k;
        if (cand_better(k, (unsigned)i, bv, bi)) { bv = k; bi = (unsigned)i; }
    }
    grid_argmin_publish(bv, bi, a.ws, a.n_ranks);
}

// ---- one greedy step, compile-time small D: rows live in registers ----------------------------------
template <typename T, int D, bool STRICT, int ITEMS>
__global__ void __launch_bounds__(kThreads) thin_step_small_kernel(ThinArgs<T> a) {
    const Cand win = pick_winner(a.ws, a.n_ranks);
    const long long sel = (long long)win.i;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        a.idx_out[a.ws->step] = (uint32_t)sel;
        a.ws->step += 1;
    }
    T y[D], sy[D];
    load_row<D>(a.x, sel, y);
    load_row<D>(a.s, sel, sy);
    T lwy = T(0), wy = T(0);
    if (a.lw) {
        lwy = __ldg(a.lw + sel);
        if (!STRICT && a.w) wy = __ldg(a.w + sel);
    }
    double bv = __longlong_as_double(0x7ff0000000000000LL);
    unsigned bi = 0xffffffffu;
    const long long rows = a.row_end - a.row_begin;
    constexpr int TILE = kThreads * ITEMS;
    for (long long base = (long long)blockIdx.x * TILE; base < rows; base += (long long)gridDim.x * TILE) {
        T xr[ITEMS][D], sr[ITEMS][D];
        double av[ITEMS];
        double wt[ITEMS];
#pragma unroll
        for (int j = 0; j < ITEMS; ++j) {
            const long long li = base + j * kThreads + threadIdx.x;
            if (li < rows) {
                const long long i = a.row_begin + li;
                load_row<D>(a.x, i, xr[j]);
                load_row<D>(a.s, i, sr[j]);
                av[j] = a.A[li];
                wt[j] = a.lw ? weight_pair<T, STRICT>(a, i, lwy, wy) : 1.0;
            }
        }
#pragma unroll
        for (int j = 0; j < ITEMS; ++j) {
            const long long li = base + j * kThreads + threadIdx.x;
            if (li < rows) {
                double k;
                if constexpr (STRICT) k = pair_strict<D>(xr[j], y, sr[j], sy, a.kp);
                else k = (double)pair_fast<T, D, false>(xr[j], y, sr[j], sy, xr[j], y, a.kp);
                double v;
                if constexpr (STRICT) {
                    if (a.lw) k = __dmul_rn(wt[j], k);
                    v = __dadd_rn(av[j], __dmul_rn(2.0, k));
                } else {
                    v = a.lw ? fma(2.0 * wt[j], k, av[j]) : fma(2.0, k, av[j]);
                }
                a.A[li] = v;
                const unsigned gi = (unsigned)(a.row_begin + li);
                if (cand_better(v, gi, bv, bi)) { bv = v; bi = gi; }
            }
        }
    }
    grid_argmin_publish(bv, bi, a.ws, a.n_ranks);
}

// ---- one greedy step, runtime d: coalesced tile -> padded shared memory -> thread per row -----------
// shared: y[d] sy[d] yl[d] | x tile [TR][d+1] | s tile [TR][d+1] | (dense) xl tile [TR][d+1]
template <typename T, bool STRICT>
__global__ void thin_step_tiled_kernel(ThinArgs<T> a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int d = a.kp.d;
    const int ld = d + 1;
    const int TR = blockDim.x;
    const bool dense = (a.xl != nullptr);
    T* ysh = reinterpret_cast<T*>(smem_raw);
    T* sysh = ysh + d;
    T* ylsh = sysh + d;
    T* xt = ylsh + d;
    T* st = xt + (size_t)TR * ld;
    T* xlt = st + (size_t)TR * ld;

    const Cand win = pick_winner(a.ws, a.n_ranks);
    const long long sel = (long long)win.i;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        a.idx_out[a.ws->step] = (uint32_t)sel;
        a.ws->step += 1;
    }
    for (int k = threadIdx.x; k < d; k += TR) {
        ysh[k] = __ldg(a.x + sel * d + k);
        sysh[k] = __ldg(a.s + sel * d + k);
        ylsh[k] = dense ? __ldg(a.xl + sel * d + k) : T(0);
    }
    T lwy = T(0), wy = T(0);
    if (a.lw) {
        lwy = __ldg(a.lw + sel);
        if (!STRICT && a.w) wy = __ldg(a.w + sel);
    }
    double bv = __longlong_as_double(0x7ff0000000000000LL);
    unsigned bi = 0xffffffffu;
    const long long rows = a.row_end - a.row_begin;
    for (long long base = (long long)blockIdx.x * TR; base < rows; base += (long long)gridDim.x * TR) {
        const int tile_rows = (int)min((long long)TR, rows - base);
        const long long g0 = (a.row_begin + base) * d;
        const int count = tile_rows * d;
        __syncthreads();  // previous tile fully consumed (also orders the y row on the first pass)
        for (int e = threadIdx.x; e < count; e += TR) {
            const int r = e / d, k = e - r * d;
            xt[r * ld + k] = __ldg(a.x + g0 + e);
            st[r * ld + k] = __ldg(a.s + g0 + e);
            if (dense) xlt[r * ld + k] = __ldg(a.xl + g0 + e);
        }
        __syncthreads();
        if ((int)threadIdx.x < tile_rows) {
            const long long li = base + threadIdx.x;
            const long long i = a.row_begin + li;
            View<T> xv{xt + threadIdx.x * ld}, sv{st + threadIdx.x * ld}, yv{ysh}, syv{sysh};
            double k;
            if constexpr (STRICT) {
                k = pair_strict<0>(xv, yv, sv, syv, a.kp);
            } else {
                if (dense) {
                    View<T> xlv{xlt + threadIdx.x * ld}, ylv{ylsh};
                    k = (double)pair_fast<T, 0, true>(xv, yv, sv, syv, xlv, ylv, a.kp);
                } else {
                    k = (double)pair_fast<T, 0, false>(xv, yv, sv, syv, xv, yv, a.kp);
                }
            }
            double v;
            const double av = a.A[li];
            if constexpr (STRICT) {
                if (a.lw) k = __dmul_rn(weight_pair<T, STRICT>(a, i, lwy, wy), k);
                v = __dadd_rn(av, __dmul_rn(2.0, k));
            } else {
                v = a.lw ? fma(2.0 * weight_pair<T, STRICT>(a, i, lwy, wy), k, av) : fma(2.0, k, av);
            }
            a.A[li] = v;
            if (cand_better(v, (unsigned)i, bv, bi)) { bv = v; bi = (unsigned)i; }
        }
    }
    grid_argmin_publish(bv, bi, a.ws, a.n_ranks);
}

// winner of the last exchange -> idx_out[m - 1]
__global__ void thin_finish_kernel(ThinWs* ws, uint32_t* idx_out, int n_ranks) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        const Cand win = pick_winner(ws, n_ranks);
        idx_out[ws->step] = (uint32_t)win.i;
        ws->step += 1;
    }
}

__global__ void thin_reset_kernel(ThinWs* ws) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        ws->counter = 0;
        ws->step = 0;
    }
}

// ---- host side -----------------------------------------------------------------------------------
template <typename T>
static int tiled_config(int d, bool dense, int& threads, size_t& smem) {
    const int arrays = dense ? 3 : 2;
    for (int t = 256; t >= 32; t >>= 1) {
        const size_t need = (size_t)(3 * d + (size_t)arrays * t * (d + 1)) * sizeof(T);
        if (need <= 100 * 1024 || (t == 32 && need <= 220 * 1024)) {
            threads = t;
            smem = need;
            return GFK_OK;
        }
    }
    set_error("d = %d is too large for the tiled thinning kernel (shared memory)", d);
    return GFK_EUNSUPPORTED;
}

template <typename T, bool STRICT>
static int launch_step(const ThinArgs<T>& a, int d, cudaStream_t st, int grid_small, int grid_tiled,
                       int tiled_threads, size_t tiled_smem) {
    const bool dense = a.xl != nullptr;
    if (!dense && d == 2) {
        thin_step_small_kernel<T, 2, STRICT, 4><<<grid_small, kThreads, 0, st>>>(a);
    } else if (!dense && d == 4) {
        thin_step_small_kernel<T, 4, STRICT, 2><<<grid_small, kThreads, 0, st>>>(a);
    } else if (!dense && d == 1) {
        thin_step_small_kernel<T, 1, STRICT, 4><<<grid_small, kThreads, 0, st>>>(a);
    } else if (!dense && d == 3) {
        thin_step_small_kernel<T, 3, STRICT, 2><<<grid_small, kThreads, 0, st>>>(a);
    } else {
        thin_step_tiled_kernel<T, STRICT><<<grid_tiled, tiled_threads, tiled_smem, st>>>(a);
    }
    return GFK_OK;
}

template <typename T, bool STRICT>
static int thin_typed(const gfk_problem* p, int64_t row_begin, int64_t row_end, int32_t m,
                      double* objective, uint32_t* idx_out, void* ws_raw, cudaStream_t st) {
    ThinArgs<T> a;
    a.x = static_cast<const T*>(p->x);
    a.s = static_cast<const T*>(p->s);
    a.lw = static_cast<const T*>(p->lw);
    a.w = static_cast<const T*>(p->w);
    a.xl = static_cast<const T*>(p->xl);
    a.A = objective;
    a.row_begin = row_begin;
    a.row_end = row_end;
    a.kp = KParams{p->lambda, p->lambda_sq, p->trace, p->c, p->d, p->xl != nullptr};
    a.ws = static_cast<ThinWs*>(ws_raw);
    a.idx_out = idx_out;
    a.n_ranks = comm_size();
    const int d = p->d;
    const long long rows = row_end - row_begin;
    const int sms = sm_count();

    int tiled_threads = 0;
    size_t tiled_smem = 0;
    const bool uses_tiled = (a.xl != nullptr) || !(d >= 1 && d <= 4);
    if (uses_tiled) {
        int rc = tiled_config<T>(d, a.xl != nullptr, tiled_threads, tiled_smem);
        if (rc != GFK_OK) return rc;
        GFK_CUDA(cudaFuncSetAttribute(thin_step_tiled_kernel<T, STRICT>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tiled_smem));
    }
    auto grid_for = [&](long long per_block, int max_per_sm) {
        long long blocks = (rows + per_block - 1) / per_block;
        long long cap = (long long)sms * max_per_sm;
        if (cap > kMaxBlocks) cap = kMaxBlocks;
        if (blocks > cap) blocks = cap;
        if (blocks < 1) blocks = 1;
        return (int)blocks;
    };
    const int grid_init = grid_for(kThreads, 8);
    const int grid_small = grid_for(kThreads * 2, 4);
    const int grid_tiled = uses_tiled ? grid_for(tiled_threads, tiled_smem > 50 * 1024 ? 2 : 4) : 1;

    thin_reset_kernel<<<1, 32, 0, st>>>(a.ws);
    thin_init_kernel<T, STRICT><<<grid_init, kThreads, 0, st>>>(a);
    GFK_CHECK_LAUNCH();
    for (int step = 1; step <= m; ++step) {
        // exchange the candidates of step-1, then either continue or finish
        if (a.n_ranks > 1) {
            int rc = comm_allgather16(&a.ws->send, a.ws->recv, st);
            if (rc != GFK_OK) return rc;
        }
        if (step == m) break;
        launch_step<T, STRICT>(a, d, st, grid_small, grid_tiled, tiled_threads, tiled_smem);
    }
    thin_finish_kernel<<<1, 32, 0, st>>>(a.ws, idx_out, a.n_ranks);
    count_launches(2 + (long long)m);   // reset + init + (m - 1) steps + finish
    GFK_CHECK_LAUNCH();
    return GFK_OK;
}

}  // namespace gfk

extern "C" size_t gfk_thin_ws_bytes(void) { return sizeof(gfk::ThinWs); }

extern "C" int gfk_thin(const gfk_problem* p, int64_t row_begin, int64_t row_end, int32_t m,
                        double* objective, uint32_t* idx_out, void* ws, void* stream) {
    using namespace gfk;
    GFK_REQUIRE(p && p->x && p->s && objective && idx_out && ws, "gfk_thin: NULL pointer argument");
    GFK_REQUIRE(p->n >= 1 && p->n <= 0xfffffffeLL, "gfk_thin: n = %lld out of range [1, 2^32-2]", (long long)p->n);
    GFK_REQUIRE(p->d >= 1, "gfk_thin: d = %d must be >= 1", p->d);
    GFK_REQUIRE(m >= 1, "gfk_thin: m = %d must be >= 1", m);
    GFK_REQUIRE(row_begin >= 0 && row_begin < row_end && row_end <= p->n,
                "gfk_thin: row range [%lld, %lld) invalid for n = %lld", (long long)row_begin, (long long)row_end,
                (long long)p->n);
    GFK_REQUIRE(!(p->xl && p->mode == GFK_F64_STRICT),
                "gfk_thin: a dense preconditioner needs a fast mode (numpy's BLAS order is not reproducible)");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    switch (p->mode) {
        case GFK_F64_STRICT: return thin_typed<double, true>(p, row_begin, row_end, m, objective, idx_out, ws, st);
        case GFK_F64_FAST: return thin_typed<double, false>(p, row_begin, row_end, m, objective, idx_out, ws, st);
        case GFK_F32_FAST: return thin_typed<float, false>(p, row_begin, row_end, m, objective, idx_out, ws, st);
        default: set_error("gfk_thin: unknown mode %d", p->mode); return GFK_EINVAL;
    }
}
