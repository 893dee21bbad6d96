// thin.cu - greedy Stein thinning: fused kernel-column update + running objective + argmin.
//
// Replaces stein_thinning.thinning._greedy_search driven by the integrand closures of
// _make_stein_integrand / _make_stein_gf_integrand (reference restatement:
// code/notebooks/examples/JAX_Stein_Thinning.ipynb#c17,#c22; algorithm report/report.tex:413-426).
//
// Data in HBM (caller owned): the standardised replica x, s (n x d row-major), optional
// log-weights lw (n), and this rank's slice of the running objective A (rows x 1, float64).
//
// Kernels in this file:
//   thin_persistent_kernel          d <= 4: ALL m greedy steps of a job in one cooperative launch; A, the log-weights
//                                   and as many rows as fit stay on chip; per-step flag-in-data grid exchange
//   thin_persistent_batched_kernel  the same loop for several independent chains, one group of CTAs per chain
//   thin_step_small / chunked / tiled / tma kernels
//                                   one launch per greedy step: every thread first reduces the <= 16 (value, index)
//                                   candidates of the previous exchange to the selected row y, then streams its
//                                   candidate rows once (16d + 16 (+8) bytes per row in float64), adds 2 k(x_i, y) to
//                                   A_i and keeps a running (min, first index); warp shuffle -> block ->
//                                   last-block-done grid reduction produces the rank's candidate for the next exchange
// plus the cross-rank exchange (NCCL all-gather between launches, or peer mailboxes inside the kernels) and its
// failure protocol.  No host round trip inside the greedy loop of either strategy.
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

#include <type_traits>
#include <vector>

#include "common.cuh"
#include "tma.cuh"
#include "comm.h"

namespace gfk {

constexpr int kMaxRanks = 16;
constexpr int kMaxBlocks = 148 * 16;
constexpr int kThreads = 256;

struct __align__(16) ThinWs {
    Cand send;               // this rank's (value, global index)
    Cand recv[kMaxRanks];    // all ranks' candidates after the exchange
    Cand partial[kMaxBlocks];
    unsigned counter;        // last-block-done ticket / persistent-kernel barrier arrivals
    unsigned step;           // greedy step whose winner is in recv[]
    unsigned abort;          // kAbort*: an exchange timed out / a peer aborted / an exchanged index was out of range
    unsigned pad_;
    unsigned long long spin_ns;        // deadline of every wait on another CTA / rank (GFK_SPIN_MS, default 200 ms)
    unsigned long long spin_first_ns;  // ... of the first cross-rank exchange of a job (GFK_SPIN_FIRST_MS): host skew
    // persistent kernel, flag-in-data slots: every 8-byte word carries (tag << 32 | payload), so a reader
    // that sees the expected tag in all three words has the whole record - no fences, no atomics
    __align__(16) unsigned long long ll[2][kMaxBlocks / 2][4];
    MailArgs mail;           // peer mailboxes of this job (written by thin_reset_kernel)
    __align__(16) unsigned long long bcast[2][4];   // persistent multi-rank kernel: CTA 0 re-publishes the global winner here
#ifdef GFK_TRACE
    long long trace[128][8];          // development builds: globaltimer (ns) of CTA 0 / thread 0 per step and phase
    long long trace_cta[64][160];     // globaltimer at which each CTA published its block minimum, steps 32..95
#endif
};

static_assert(offsetof(ThinWs, ll) % 16 == 0 && offsetof(ThinWs, bcast) % 16 == 0 && sizeof(ThinWs::ll[0][0]) == 32,
              "flag-in-data slots are moved with 16-byte vector accesses");
static_assert(offsetof(ThinWs, spin_ns) % 8 == 0, "64-bit fields are naturally aligned");
static_assert(offsetof(ThinWs, send) % 16 == 0 && offsetof(ThinWs, recv) % 16 == 0, "Cand records are 16-byte aligned");

#ifdef GFK_TRACE
__device__ __forceinline__ long long gfk_now() {
#if defined(GFK_SIMT_HOST)
    return 0;
#else
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t) :: "memory");
    return t;
#endif
}
#define GFK_STAMP(step, phase) do { if (blockIdx.x == 0 && threadIdx.x == 0 && (step) < 128) ws->trace[(step)][(phase)] = gfk_now(); } while (0)
#define GFK_STAMP_CTA(step) do { if (threadIdx.x == 0 && (step) >= 32 && (step) < 96 && blockIdx.x < 160) ws->trace_cta[(step) - 32][blockIdx.x] = gfk_now(); } while (0)
#else
#define GFK_STAMP(step, phase) do { } while (0)
#define GFK_STAMP_CTA(step) do { } while (0)
#endif

template <typename T>
struct ThinArgs {
    const T* x;
    const T* s;
    const T* lw;    // log-weights (nullptr: score-based kernel)
    const T* w;     // exp(lw) (fast modes; nullptr: exp(lw_i + lw_y) on the fly)
    const T* xl;    // x @ L for a dense preconditioner (nullptr: scalar)
    double* A;
    long long row_begin, row_end;
    long long n;     // rows of the replica: every exchanged index is checked against it before it is used as an address
    KParams kp;
    ThinWs* ws;
    uint32_t* idx_out;
    int n_ranks;
    int step;        // multi-launch kernels: index of the exchange this launch produces (consumes step - 1)
    int mail_enabled; // exchange through the peer mailboxes in ws->mail instead of NCCL
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

// ---- failure protocol ---------------------------------------------------------------------------------------
// Every wait on memory that another CTA or another rank writes is bounded by a TIME deadline (globaltimer), never
// by an iteration count: a polling round is an L2 or NVLink round trip of unknown length.  When a wait runs out
// the waiter sets ws->abort (all local CTAs poll it in their own waits and leave), and - multi-rank - stores this
// job's tag into the poison word of every peer mailbox, which the peers' waits poll as well, so every rank leaves
// within one deadline.  A record that does arrive is still range-checked before its index is used as an address.
// gfk_thin_check() reports the abort code as GFK_ECUDA.
constexpr unsigned kAbortTimeout = 1u;     // a wait ran out
constexpr unsigned kAbortPeer = 2u;        // a peer rank posted its poison word
constexpr unsigned kAbortIndex = 3u;       // an exchanged index was >= n
constexpr unsigned kSpinCheckEvery = 64u;  // read the timer every 64th unsuccessful polling round only

__device__ __forceinline__ unsigned ld_volatile_u32(const unsigned* p) {
    return *reinterpret_cast<const volatile unsigned*>(p);
}

// one per wait loop; expired() is called after every unsuccessful polling round
struct SpinDeadline {
    unsigned rounds = 0;
    unsigned long long t0 = 0;
    __device__ __forceinline__ bool expired(unsigned long long limit_ns) {
        if ((++rounds % kSpinCheckEvery) != 0) return false;
        const unsigned long long now = gfk_timer_ns();
        if (t0 == 0) { t0 = now; return false; }
        return now - t0 > limit_ns;
    }
};

// inside every spin loop on memory another CTA / rank writes: nothing on the device; the CPU emulation build of
// the test suite (tests/host/simt) lets the other emulated threads run
#if defined(GFK_SIMT_HOST)
#define GFK_SPIN_HINT() simt::poll_yield()
#else
#define GFK_SPIN_HINT() do { } while (0)
#endif

// Intra-GPU record: (value 64 bits, index 32 bits) in TWO self-validating words, each a 16-bit tag over a
// 48-bit payload, moved with one 16-byte vector access (each 8-byte half is single-copy atomic and carries
// its own tag, so a torn pair is simply seen as "not there yet").  One request per slot instead of three.
__device__ __forceinline__ void slot_store(unsigned long long* p, unsigned tag16, double v, unsigned i) {
    const unsigned long long vb = (unsigned long long)__double_as_longlong(v);
    const unsigned long long t = (unsigned long long)(tag16 & 0xffffu) << 48;
    const unsigned long long a = t | (vb >> 16);
    const unsigned long long b = t | ((vb & 0xffffull) << 32) | (unsigned long long)i;
#if defined(GFK_SIMT_HOST)
    simt::check_align(p, 16, "slot_store");
    __atomic_store_n(p, a, __ATOMIC_RELAXED);
    __atomic_store_n(p + 1, b, __ATOMIC_RELAXED);
#else
    asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(a), "l"(b) : "memory");
#endif
}
__device__ __forceinline__ void slot_load(const unsigned long long* p, unsigned long long& a, unsigned long long& b) {
#if defined(GFK_SIMT_HOST)
    simt::check_align(p, 16, "slot_load");
    a = __atomic_load_n(p, __ATOMIC_RELAXED);
    b = __atomic_load_n(p + 1, __ATOMIC_RELAXED);
#else
    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
#endif
}
__device__ __forceinline__ bool slot_ready(unsigned long long a, unsigned long long b, unsigned tag16) {
    return (unsigned)(a >> 48) == (tag16 & 0xffffu) && (unsigned)(b >> 48) == (tag16 & 0xffffu);
}
__device__ __forceinline__ void slot_decode(unsigned long long a, unsigned long long b, double& v, unsigned& i) {
    const unsigned long long vb = (a << 16) | ((b >> 32) & 0xffffull);
    v = __longlong_as_double((long long)vb);
    i = (unsigned)(b & 0xffffffffull);
}

__device__ __forceinline__ int mail_offset(const MailArgs& m, int exchange, int rank) {
    return (((m.job_parity * 2 + (exchange & 1)) * kMailRanks) + rank) * 4;
}

// lanes 0 .. n_ranks-1 of ONE warp: lane r stores this rank's candidate into rank r's mailbox (NVLink store
// for r != rank).  Three tagged 8-byte words; no fence is needed because every word validates itself.
__device__ __forceinline__ void mail_post(const MailArgs& m, int n_ranks, int exchange, double v, unsigned i,
                                          int lane) {
    if (lane < n_ranks) {
        const unsigned long long tag = (unsigned long long)(m.tag_base + (unsigned)exchange + 1u) << 32;
        const unsigned long long vb = (unsigned long long)__double_as_longlong(v);
        volatile unsigned long long* w = m.peer[lane] + mail_offset(m, exchange, m.rank);
        w[0] = tag | (vb & 0xffffffffull);
        w[1] = tag | (vb >> 32);
        w[2] = tag | (unsigned long long)i;
    }
}

// lanes 0 .. n_ranks-1 of one warp: tell every rank (this one included) that this job is over
__device__ __forceinline__ void mail_poison(const MailArgs& m, int n_ranks, int lane) {
    if (lane < n_ranks) {
        volatile unsigned long long* w = m.peer[lane] + kMailPoison;
        *w = (unsigned long long)m.tag_base;
    }
}
__device__ __forceinline__ bool mail_poisoned(const MailArgs& m) {
    const volatile unsigned long long* w = m.local + kMailPoison;
    return *w == (unsigned long long)m.tag_base;
}

// a full warp: lane r polls slot r of the local mailbox until its three tags match, then the warp reduces
// the candidates with numpy.argmin semantics.  Every lane returns the winner.  dead != 0 (a kAbort* code, the same
// in every lane): the wait ran out, the local abort flag went up or a peer posted its poison word - the returned
// record is then (+inf, 0xffffffff), never a stale slot.
__device__ __forceinline__ Cand mail_wait(const MailArgs& m, int n_ranks, int exchange, int lane, unsigned& dead,
                                          const unsigned* abort_flag, unsigned long long limit_ns) {
    double v = __longlong_as_double(0x7ff0000000000000LL);
    unsigned i = 0xffffffffu;
    unsigned bad = 0;
    if (lane < n_ranks) {
        const unsigned tag = m.tag_base + (unsigned)exchange + 1u;
        const volatile unsigned long long* w = m.local + mail_offset(m, exchange, lane);
        unsigned long long w0, w1, w2;
        SpinDeadline dl;
#pragma unroll 1
        for (;;) {
            w0 = w[0]; w1 = w[1]; w2 = w[2];
            if ((unsigned)(w0 >> 32) == tag && (unsigned)(w1 >> 32) == tag && (unsigned)(w2 >> 32) == tag) break;
            if (abort_flag && ld_volatile_u32(abort_flag)) { bad = kAbortTimeout; break; }
            if (mail_poisoned(m)) { bad = kAbortPeer; break; }
            if (dl.expired(limit_ns)) { bad = kAbortTimeout; break; }
            GFK_SPIN_HINT();
        }
        if (!bad) {
            v = __longlong_as_double((long long)(((w1 & 0xffffffffull) << 32) | (w0 & 0xffffffffull)));
            i = (unsigned)(w2 & 0xffffffffull);
        }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const unsigned o = __shfl_xor_sync(0xffffffffu, bad, off);
        bad = o > bad ? o : bad;
    }
    dead = bad;
    warp_argmin(v, i);
    Cand c;
    c.v = dead ? __longlong_as_double(0x7ff0000000000000LL) : v;
    c.i = dead ? 0xffffffffu : i;
    return c;
}

// post + wait as one out-of-line call: keeps the single-GPU step loop of the persistent kernel small enough
// for the instruction cache (one warp walks this code once per step, so every cold line is exposed latency)
__device__ __noinline__ Cand mail_exchange(const MailArgs& m, int n_ranks, int exchange, double v, unsigned i,
                                            int lane, bool post, unsigned& dead, const unsigned* abort_flag,
                                            unsigned long long limit_ns) {
    if (post) mail_post(m, n_ranks, exchange, v, i, lane);
    return mail_wait(m, n_ranks, exchange, lane, dead, abort_flag, limit_ns);
}

// winner of exchange (a.step - 1): from the peer mailboxes or from recv[]; warp 0 decides, the block gets it through
// shared memory.  Returns index 0xffffffff when the job is over - aborted by an earlier launch, another CTA or a peer,
// the exchange timed out, or the record is out of range (then ws->abort goes up and, with mailboxes, the peers are
// told).  The callers leave at once (winner_ok), so a stale or foreign record never becomes an address.
template <typename T>
__device__ __forceinline__ Cand pick_winner(const ThinArgs<T>& a) {
    __shared__ Cand win_bc;
    if (threadIdx.x < 32) {
        const bool mail = a.mail_enabled && a.n_ranks > 1;
        unsigned dead = __shfl_sync(0xffffffffu, ld_volatile_u32(&a.ws->abort), 0);
        Cand c;
        c.v = __longlong_as_double(0x7ff0000000000000LL);
        c.i = 0xffffffffull;
        if (!dead) {
            if (mail) {
                const unsigned long long limit = (a.step <= 1) ? a.ws->spin_first_ns : a.ws->spin_ns;
                c = mail_wait(a.ws->mail, a.n_ranks, a.step - 1, threadIdx.x, dead, &a.ws->abort, limit);
            } else {
                c = a.ws->recv[0];
                for (int r = 1; r < a.n_ranks; ++r) {
                    const Cand o = a.ws->recv[r];
                    if (cand_better(o.v, o.i, c.v, c.i)) c = o;
                }
            }
            if (!dead && c.i >= (unsigned long long)a.n) dead = kAbortIndex;
            if (dead) {
                c.i = 0xffffffffull;
                if (threadIdx.x == 0 && ld_volatile_u32(&a.ws->abort) == 0u) a.ws->abort = dead;
                if (mail && dead != kAbortPeer) mail_poison(a.ws->mail, a.n_ranks, threadIdx.x);
            }
        }
        if (threadIdx.x == 0) win_bc = c;
    }
    __syncthreads();
    return win_bc;
}

__device__ __forceinline__ bool winner_ok(const Cand& win) { return win.i != 0xffffffffull; }

// grid-level argmin: block result -> partial[], last block reduces partial[] -> ws->send
// (single rank: the exchange is the identity, so the result goes straight into recv[0])
template <typename T>
__device__ __forceinline__ void grid_argmin_publish(double v, unsigned i, const ThinArgs<T>& a) {
    ThinWs* ws = a.ws;
    const int n_ranks = a.n_ranks;
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
    block_argmin(bv, bi, sv, si);      // every lane of warp 0 holds the result
    if (a.mail_enabled && n_ranks > 1 && threadIdx.x < 32) mail_post(ws->mail, n_ranks, a.step, bv, bi, threadIdx.x);
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

// gradient-free weight state of the selected row y.  Strict mode: a.w, when present, holds double-double
// exponentials (hi, lo) per row (gfk_exp_weights_dd) and the weight is exp_sum_dd (stein_math.cuh) - IEEE operations
// only; without it the device exp() is used.  Fast modes: a.w holds plain exp(lw) and the weight is a product.
template <typename T>
struct YWeight {
    T lw, w;
    double eh, el;
};

template <typename T, bool STRICT>
__device__ __forceinline__ YWeight<T> load_yweight(const ThinArgs<T>& a, long long sel) {
    YWeight<T> y;
    y.lw = T(0); y.w = T(0); y.eh = 0.0; y.el = 0.0;
    if (a.lw) {
        y.lw = __ldg(a.lw + sel);
        if constexpr (STRICT) {
            if (a.w) {
                const double2 e = __ldg(reinterpret_cast<const double2*>(a.w) + sel);
                y.eh = e.x;
                y.el = e.y;
            }
        } else {
            if (a.w) y.w = __ldg(a.w + sel);
        }
    }
    return y;
}

template <typename T, bool STRICT>
__device__ __forceinline__ double weight_pair(const ThinArgs<T>& a, long long i, const YWeight<T>& y) {
    if constexpr (STRICT) {
        const double lwi = (double)__ldg(a.lw + i);
        if (a.w) {
            const double2 e = __ldg(reinterpret_cast<const double2*>(a.w) + i);
            return exp_sum_dd(lwi, (double)y.lw, e.x, e.y, y.eh, y.el);
        }
        return exp(lwi + (double)y.lw);
    } else {
        if (a.w == nullptr) return exp((double)__ldg(a.lw + i) + (double)y.lw);
        return (double)(__ldg(a.w + i) * y.w);
    }
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
            const YWeight<T> self = load_yweight<T, STRICT>(a, i);
            k = STRICT ? __dmul_rn(weight_pair<T, STRICT>(a, i, self), k)
                       : weight_pair<T, STRICT>(a, i, self) * k;
        }
        a.A[li] = k;
        if (cand_better(k, (unsigned)i, bv, bi)) { bv = k; bi = (unsigned)i; }
    }
    grid_argmin_publish(bv, bi, a);
}

// ---- one greedy step, compile-time small D: rows live in registers ----------------------------------
template <typename T, int D, bool STRICT, int ITEMS>
__global__ void __launch_bounds__(kThreads) thin_step_small_kernel(ThinArgs<T> a) {
    const Cand win = pick_winner(a);
    if (!winner_ok(win)) return;                 // job aborted (block-uniform): nothing is read, nothing is published
    const long long sel = (long long)win.i;
    if (blockIdx.x == 0 && threadIdx.x == 0) a.idx_out[a.step - 1] = (uint32_t)sel;
    T y[D], sy[D];
    load_row<D>(a.x, sel, y);
    load_row<D>(a.s, sel, sy);
    const YWeight<T> yw = load_yweight<T, STRICT>(a, sel);
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
                wt[j] = a.lw ? weight_pair<T, STRICT>(a, i, yw) : 1.0;
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
    grid_argmin_publish(bv, bi, a);
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

    const Cand win = pick_winner(a);
    if (!winner_ok(win)) return;                 // job aborted (block-uniform): nothing is read, nothing is published
    const long long sel = (long long)win.i;
    if (blockIdx.x == 0 && threadIdx.x == 0) a.idx_out[a.step - 1] = (uint32_t)sel;
    for (int k = threadIdx.x; k < d; k += TR) {
        ysh[k] = __ldg(a.x + sel * d + k);
        sysh[k] = __ldg(a.s + sel * d + k);
        ylsh[k] = dense ? __ldg(a.xl + sel * d + k) : T(0);
    }
    const YWeight<T> yw = load_yweight<T, STRICT>(a, sel);
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
                if (a.lw) k = __dmul_rn(weight_pair<T, STRICT>(a, i, yw), k);
                v = __dadd_rn(av, __dmul_rn(2.0, k));
            } else {
                v = a.lw ? fma(2.0 * weight_pair<T, STRICT>(a, i, yw), k, av) : fma(2.0, k, av);
            }
            a.A[li] = v;
            if (cand_better(v, (unsigned)i, bv, bi)) { bv = v; bi = (unsigned)i; }
        }
    }
    grid_argmin_publish(bv, bi, a);
}

// ---- one greedy step, runtime d, chunked: [256 rows] x [KC dims] slabs through shared memory ---------
// Each row contributes KC consecutive values (one 128-byte line in float64) to a slab, so the global
// loads are fully coalesced for any d; thread t owns row t of the tile and carries its running sums
// across the slabs in registers, in numpy's order (sequential s1..s3, eight strided accumulators for s4
// when d >= 8), so the strict mode stays bit-faithful for d <= 128.
constexpr int kChunkK = 16;
constexpr int kChunkLd = kChunkK + 1;

template <typename T, bool STRICT, bool DENSE>
__global__ void __launch_bounds__(kThreads) thin_step_chunked_kernel(ThinArgs<T> a) {
    extern __shared__ __align__(16) unsigned char csm[];
    const int d = a.kp.d;
    T* ysh = reinterpret_cast<T*>(csm);
    T* sysh = ysh + d;
    T* ylsh = sysh + d;
    T* xt = ylsh + d;
    T* st = xt + kThreads * kChunkLd;
    T* xlt = st + kThreads * kChunkLd;

    const Cand win = pick_winner(a);
    if (!winner_ok(win)) return;                 // job aborted (block-uniform): nothing is read, nothing is published
    const long long sel = (long long)win.i;
    if (blockIdx.x == 0 && threadIdx.x == 0) a.idx_out[a.step - 1] = (uint32_t)sel;
    for (int k = threadIdx.x; k < d; k += kThreads) {
        ysh[k] = __ldg(a.x + sel * d + k);
        sysh[k] = __ldg(a.s + sel * d + k);
        ylsh[k] = DENSE ? __ldg(a.xl + sel * d + k) : T(0);
    }
    const YWeight<T> yw = load_yweight<T, STRICT>(a, sel);
    double bv = __longlong_as_double(0x7ff0000000000000LL);
    unsigned bi = 0xffffffffu;
    const long long rows = a.row_end - a.row_begin;
    const int d8 = d - (d % 8);                     // numpy's pairwise main region
    for (long long base = (long long)blockIdx.x * kThreads; base < rows; base += (long long)gridDim.x * kThreads) {
        const int tile_rows = (int)min((long long)kThreads, rows - base);
        const bool active = (int)threadIdx.x < tile_rows;
        const long long li = base + threadIdx.x;
        // running sums of this thread's row
        double s1 = 0.0, s2 = 0.0, s3 = 0.0, s4 = 0.0;          // strict
        double r0 = 0, r1 = 0, r2 = 0, r3 = 0, r4 = 0, r5 = 0, r6 = 0, r7 = 0;
        T rr = T(0), uu = T(0), su = T(0), ss = T(0);             // fast
        for (int k0 = 0; k0 < d; k0 += kChunkK) {
            const int kc = min(kChunkK, d - k0);
            __syncthreads();                                      // previous slab consumed; y row visible
            {
                // thread t stages column (t % 16) of rows (t / 16) + 16 j: all 2 x 16 loads are independent
                const int kk = threadIdx.x % kChunkK, r0 = threadIdx.x / kChunkK;
                T xv[kChunkK], sv[kChunkK], lv[DENSE ? kChunkK : 1];
                const long long g0 = (a.row_begin + base + r0) * d + k0 + kk;
#pragma unroll
                for (int j = 0; j < kChunkK; ++j) {
                    const bool ok = (kk < kc) && (r0 + j * (kThreads / kChunkK) < tile_rows);
                    const long long g = g0 + (long long)j * (kThreads / kChunkK) * d;
                    xv[j] = ok ? __ldg(a.x + g) : T(0);
                    sv[j] = ok ? __ldg(a.s + g) : T(0);
                    if (DENSE) lv[j] = ok ? __ldg(a.xl + g) : T(0);
                }
#pragma unroll
                for (int j = 0; j < kChunkK; ++j) {
                    const int r = r0 + j * (kThreads / kChunkK);
                    xt[r * kChunkLd + kk] = xv[j];
                    st[r * kChunkLd + kk] = sv[j];
                    if (DENSE) xlt[r * kChunkLd + kk] = lv[j];
                }
            }
            __syncthreads();
            if (active) {
                const T* xr = xt + threadIdx.x * kChunkLd;
                const T* sr = st + threadIdx.x * kChunkLd;
                const T* xlr = xlt + threadIdx.x * kChunkLd;
                if constexpr (STRICT) {
                    for (int k = 0; k < kc; ++k) {
                        const double r = nf_sub(xr[k], ysh[k0 + k]);
                        const double pa = nf_mul(nf_mul(a.kp.lambda, r), r);
                        const double pb = nf_mul(nf_mul(a.kp.lambda_sq, r), r);
                        const double pc = nf_mul(nf_mul(a.kp.lambda, nf_sub(sr[k], sysh[k0 + k])), r);
                        if (k0 + k == 0) { s1 = pa; s2 = pb; s3 = pc; }
                        else { s1 = nf_add(s1, pa); s2 = nf_add(s2, pb); s3 = nf_add(s3, pc); }
                    }
                    if (d < 8) {
                        for (int k = 0; k < kc; ++k) {
                            const double p = nf_mul(sr[k], sysh[k0 + k]);
                            s4 = (k0 + k == 0) ? p : nf_add(s4, p);
                        }
                    } else {
                        // groups of 8 inside the main region [0, d8); the tail [d8, d) is added after the tree
                        int k = 0;
                        for (; k + 8 <= kc && k0 + k + 8 <= d8; k += 8) {
                            const double p0 = nf_mul(sr[k + 0], sysh[k0 + k + 0]), p1 = nf_mul(sr[k + 1], sysh[k0 + k + 1]);
                            const double p2 = nf_mul(sr[k + 2], sysh[k0 + k + 2]), p3 = nf_mul(sr[k + 3], sysh[k0 + k + 3]);
                            const double p4 = nf_mul(sr[k + 4], sysh[k0 + k + 4]), p5 = nf_mul(sr[k + 5], sysh[k0 + k + 5]);
                            const double p6 = nf_mul(sr[k + 6], sysh[k0 + k + 6]), p7 = nf_mul(sr[k + 7], sysh[k0 + k + 7]);
                            if (k0 + k == 0) { r0 = p0; r1 = p1; r2 = p2; r3 = p3; r4 = p4; r5 = p5; r6 = p6; r7 = p7; }
                            else {
                                r0 = nf_add(r0, p0); r1 = nf_add(r1, p1); r2 = nf_add(r2, p2); r3 = nf_add(r3, p3);
                                r4 = nf_add(r4, p4); r5 = nf_add(r5, p5); r6 = nf_add(r6, p6); r7 = nf_add(r7, p7);
                            }
                        }
                        if (k0 + k >= d8 && k < kc) {
                            if (k0 + k == d8)
                                s4 = nf_add(nf_add(nf_add(r0, r1), nf_add(r2, r3)), nf_add(nf_add(r4, r5), nf_add(r6, r7)));
                            for (; k < kc; ++k) s4 = nf_add(s4, nf_mul(sr[k], sysh[k0 + k]));
                        }
                    }
                } else {
#pragma unroll 4
                    for (int k = 0; k < kc; ++k) {
                        const T r = xr[k] - ysh[k0 + k];
                        const T ds = sr[k] - sysh[k0 + k];
                        if (DENSE) {
                            const T u = xlr[k] - ylsh[k0 + k];
                            rr = fma_t(u, r, rr);
                            uu = fma_t(u, u, uu);
                            su = fma_t(ds, u, su);
                        } else {
                            rr = fma_t(r, r, rr);
                            su = fma_t(ds, r, su);
                        }
                        ss = fma_t(sr[k], sysh[k0 + k], ss);
                    }
                }
            }
        }
        if (active) {
            const long long i = a.row_begin + li;
            double k;
            if constexpr (STRICT) {
                if (d >= 8) {
                    if (d8 == d) s4 = nf_add(nf_add(nf_add(r0, r1), nf_add(r2, r3)), nf_add(nf_add(r4, r5), nf_add(r6, r7)));
                    s4 = nf_add(0.0, s4);
                }
                k = strict_finish(s1, s2, s3, s4, a.kp);
            } else {
                k = (double)fast_finish<T, DENSE>(rr, uu, su, ss, a.kp);
            }
            double v;
            const double av = a.A[li];
            if constexpr (STRICT) {
                if (a.lw) k = __dmul_rn(weight_pair<T, STRICT>(a, i, yw), k);
                v = __dadd_rn(av, __dmul_rn(2.0, k));
            } else {
                v = a.lw ? fma(2.0 * weight_pair<T, STRICT>(a, i, yw), k, av) : fma(2.0, k, av);
            }
            a.A[li] = v;
            if (cand_better(v, (unsigned)i, bv, bi)) { bv = v; bi = (unsigned)i; }
        }
    }
    grid_argmin_publish(bv, bi, a);
}

// ---- the same step fed by the Tensor Memory Accelerator: one CTA per SM, slabs arrive by cp.async.bulk.tensor -------
// thin_step_chunked_kernel loads a slab with LDG, waits, consumes it, and relies on the three CTAs of an SM to overlap
// one another (C5, d = 64: 56 % of the HBM roofline).  Here warp 8 of a 288-thread CTA is a producer: its elected
// lane keeps kTmaStages slabs in flight with TWO instructions per slab - one 2-D bulk tensor copy each for x and s
// ([16 dims] x [256 rows] of float64 = 32 KB per array) - and the eight consumer warps only compute.  Producer and
// consumers meet through mbarriers: full[stage] (expect_tx bytes of both copies), empty[stage] (one arrival per
// consumer warp); nobody executes a __syncthreads inside the loop.  The tensor maps use CU_TENSOR_MAP_SWIZZLE_128B:
// a row of the tile is 128 bytes and its 16-byte chunks are XOR-ed with (row mod 8), so thread-per-row reads of one
// dimension take 4 shared-memory wavefronts per warp (2 is the floor for 32 x 8 bytes; the unswizzled layout needs
// 32).  Per-row arithmetic and summation order are those of thin_step_chunked_kernel (bit-identical results): the
// row state lives in ChunkRow, the slab loop is flattened over (tile, slab) stages.  Rows past the end of the sample
// and dimensions past d are zero-filled by the copy engine and never read.
// [round 2: opt-in with GFK_STEP_TMA=1; executed on the CPU emulator of the test suite only, not yet on a GPU]
constexpr int kTmaRows = 256;                        // rows of a tile = consumer threads
constexpr int kTmaStages = 3;
constexpr int kTmaThreads = kTmaRows + 32;           // + the producer warp
constexpr int kTmaRowBytes = 128;                    // bytes of one tile row = the swizzle span
template <typename T> struct TmaSlab { static constexpr int K = kTmaRowBytes / (int)sizeof(T); };   // dims per slab
constexpr int kTmaTileBytes = kTmaRows * kTmaRowBytes;                                                // one array

// row r of a SWIZZLE_128B tile: element k lives in 16-byte chunk (k * sizeof(T) / 16) XOR (r mod 8)
template <typename T>
struct SwizzledRow {
    const unsigned char* row;
    unsigned r7;
    __device__ __forceinline__ T operator[](int k) const {
        const unsigned byte = (unsigned)k * (unsigned)sizeof(T);
        return *reinterpret_cast<const T*>(row + ((((byte >> 4) ^ r7) << 4) | (byte & 15u)));
    }
};

// running sums of one row across the slabs, numpy's order (see thin_step_chunked_kernel)
template <typename T, bool STRICT, bool DENSE>
struct ChunkRow {
    double s1, s2, s3, s4, r0, r1, r2, r3, r4, r5, r6, r7;   // strict
    T rr, uu, su, ss;                                         // fast

    __device__ __forceinline__ void reset() {
        s1 = s2 = s3 = s4 = 0.0;
        r0 = r1 = r2 = r3 = r4 = r5 = r6 = r7 = 0.0;
        rr = uu = su = ss = T(0);
    }

    // xr / sr / xlr: anything indexable by k in [0, kc) - plain pointers or the swizzled rows of a TMA tile
    template <typename AX, typename AS, typename AL>
    __device__ __forceinline__ void slab(const AX& xr, const AS& sr, const AL& xlr, const T* ysh, const T* sysh,
                                         const T* ylsh, int k0, int kc, int d, int d8, const KParams& kp) {
        if constexpr (STRICT) {
            for (int k = 0; k < kc; ++k) {
                const double r = nf_sub(xr[k], ysh[k0 + k]);
                const double pa = nf_mul(nf_mul(kp.lambda, r), r);
                const double pb = nf_mul(nf_mul(kp.lambda_sq, r), r);
                const double pc = nf_mul(nf_mul(kp.lambda, nf_sub(sr[k], sysh[k0 + k])), r);
                if (k0 + k == 0) { s1 = pa; s2 = pb; s3 = pc; }
                else { s1 = nf_add(s1, pa); s2 = nf_add(s2, pb); s3 = nf_add(s3, pc); }
            }
            if (d < 8) {
                for (int k = 0; k < kc; ++k) {
                    const double p = nf_mul(sr[k], sysh[k0 + k]);
                    s4 = (k0 + k == 0) ? p : nf_add(s4, p);
                }
            } else {
                int k = 0;
                for (; k + 8 <= kc && k0 + k + 8 <= d8; k += 8) {
                    const double p0 = nf_mul(sr[k + 0], sysh[k0 + k + 0]), p1 = nf_mul(sr[k + 1], sysh[k0 + k + 1]);
                    const double p2 = nf_mul(sr[k + 2], sysh[k0 + k + 2]), p3 = nf_mul(sr[k + 3], sysh[k0 + k + 3]);
                    const double p4 = nf_mul(sr[k + 4], sysh[k0 + k + 4]), p5 = nf_mul(sr[k + 5], sysh[k0 + k + 5]);
                    const double p6 = nf_mul(sr[k + 6], sysh[k0 + k + 6]), p7 = nf_mul(sr[k + 7], sysh[k0 + k + 7]);
                    if (k0 + k == 0) { r0 = p0; r1 = p1; r2 = p2; r3 = p3; r4 = p4; r5 = p5; r6 = p6; r7 = p7; }
                    else {
                        r0 = nf_add(r0, p0); r1 = nf_add(r1, p1); r2 = nf_add(r2, p2); r3 = nf_add(r3, p3);
                        r4 = nf_add(r4, p4); r5 = nf_add(r5, p5); r6 = nf_add(r6, p6); r7 = nf_add(r7, p7);
                    }
                }
                if (k0 + k >= d8 && k < kc) {
                    if (k0 + k == d8)
                        s4 = nf_add(nf_add(nf_add(r0, r1), nf_add(r2, r3)), nf_add(nf_add(r4, r5), nf_add(r6, r7)));
                    for (; k < kc; ++k) s4 = nf_add(s4, nf_mul(sr[k], sysh[k0 + k]));
                }
            }
        } else {
#pragma unroll 4
            for (int k = 0; k < kc; ++k) {
                const T r = xr[k] - ysh[k0 + k];
                const T ds = sr[k] - sysh[k0 + k];
                if (DENSE) {
                    const T u = xlr[k] - ylsh[k0 + k];
                    rr = fma_t(u, r, rr);
                    uu = fma_t(u, u, uu);
                    su = fma_t(ds, u, su);
                } else {
                    rr = fma_t(r, r, rr);
                    su = fma_t(ds, r, su);
                }
                ss = fma_t(sr[k], sysh[k0 + k], ss);
            }
        }
    }

    __device__ __forceinline__ double value(int d, int d8, const KParams& kp) {
        if constexpr (STRICT) {
            if (d >= 8) {
                if (d8 == d) s4 = nf_add(nf_add(nf_add(r0, r1), nf_add(r2, r3)), nf_add(nf_add(r4, r5), nf_add(r6, r7)));
                s4 = nf_add(0.0, s4);
            }
            return strict_finish(s1, s2, s3, s4, kp);
        } else {
            return (double)fast_finish<T, DENSE>(rr, uu, su, ss, kp);
        }
    }
};

template <typename T, bool STRICT>
__global__ void __launch_bounds__(kTmaThreads, 1) thin_step_tma_kernel(ThinArgs<T> a, const __grid_constant__ TensorMap map_x,
                                                                       const __grid_constant__ TensorMap map_s) {
    extern __shared__ __align__(16) unsigned char tsm[];
    __shared__ __align__(8) unsigned long long full_bar[kTmaStages], empty_bar[kTmaStages];
    constexpr int KS = TmaSlab<T>::K;
    const int d = a.kp.d;
    // [1024-byte aligned] kTmaStages x { x tile, s tile } | y[d] sy[d]
    unsigned char* tiles = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(tsm) + 1023u) & ~(uintptr_t)1023u);
    T* ysh = reinterpret_cast<T*>(tiles + (size_t)kTmaStages * 2 * kTmaTileBytes);
    T* sysh = ysh + d;

    const Cand win = pick_winner(a);
    if (!winner_ok(win)) return;                 // job aborted (block-uniform): nothing is read, nothing is published
    const long long sel = (long long)win.i;
    if (blockIdx.x == 0 && threadIdx.x == 0) a.idx_out[a.step - 1] = (uint32_t)sel;
    for (int k = threadIdx.x; k < d; k += kTmaThreads) {
        ysh[k] = __ldg(a.x + sel * d + k);
        sysh[k] = __ldg(a.s + sel * d + k);
    }
    if (threadIdx.x == 0) {
        for (int st = 0; st < kTmaStages; ++st) {
            dev_mbar_init(&full_bar[st], 1);                     // the producer's arrive.expect_tx
            dev_mbar_init(&empty_bar[st], kTmaRows / 32);        // one arrival per consumer warp
        }
        dev_mbar_fence_init();
    }
    __syncthreads();                                         // barriers initialised, y row visible

    const YWeight<T> yw = load_yweight<T, STRICT>(a, sel);
    const long long rows = a.row_end - a.row_begin;
    const long long stride = (long long)gridDim.x * kTmaRows;
    const long long first = (long long)blockIdx.x * kTmaRows;
    const long long ntiles = first < rows ? (rows - first + stride - 1) / stride : 0;
    const int nslabs = (d + KS - 1) / KS;
    const long long nstages = ntiles * nslabs;
    const int d8 = d - (d % 8);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double bv = __longlong_as_double(0x7ff0000000000000LL);
    unsigned bi = 0xffffffffu;

    if (warp == kTmaRows / 32) {
        // ---- producer warp: one elected lane keeps the ring full ----------------------------------------------
        if (lane == 0) {
            for (long long q = 0; q < nstages; ++q) {
                const int st = (int)(q % kTmaStages);
                const long long use = q / kTmaStages;                       // how often this buffer was filled before
                if (use > 0 && !dev_mbar_wait(&empty_bar[st], (unsigned)((use - 1) & 1), &a.ws->abort)) break;   // previous contents consumed
                const long long it = q / nslabs;
                const int k0 = (int)(q - it * nslabs) * KS;
                const long long row0 = a.row_begin + first + it * stride;       // global row of the tile's first row
                unsigned char* xt = tiles + (size_t)st * 2 * kTmaTileBytes;
                dev_mbar_arrive_expect_tx(&full_bar[st], 2u * (unsigned)kTmaTileBytes);
                dev_tma_load_2d(xt, &map_x, k0, (int)row0, &full_bar[st]);
                dev_tma_load_2d(xt + kTmaTileBytes, &map_s, k0, (int)row0, &full_bar[st]);
            }
        }
        __syncwarp();
    } else {
        // ---- consumer warps: thread per row -----------------------------------------------------------------
        ChunkRow<T, STRICT, false> row;
        row.reset();
        const unsigned r7 = threadIdx.x & 7u;
        for (long long q = 0; q < nstages; ++q) {
            const int st = (int)(q % kTmaStages);
            // both tiles of slab q have landed; the verdict is made warp-uniform so that a warp leaves the loop together
            if (!__all_sync(0xffffffffu, dev_mbar_wait(&full_bar[st], (unsigned)((q / kTmaStages) & 1), &a.ws->abort))) break;
            const long long it = q / nslabs;
            const int slab = (int)(q - it * nslabs);
            const int k0 = slab * KS;
            const int kc = min(KS, d - k0);
            const long long base = first + it * stride;
            const int tile_rows = (int)min((long long)kTmaRows, rows - base);
            const unsigned char* xt = tiles + (size_t)st * 2 * kTmaTileBytes;
            if ((int)threadIdx.x < tile_rows) {
                const SwizzledRow<T> xr{xt + (size_t)threadIdx.x * kTmaRowBytes, r7};
                const SwizzledRow<T> sr{xt + kTmaTileBytes + (size_t)threadIdx.x * kTmaRowBytes, r7};
                if (slab == 0) row.reset();
                row.slab(xr, sr, xr, ysh, sysh, ysh, k0, kc, d, d8, a.kp);
                if (slab == nslabs - 1) {
                    const long long li = base + threadIdx.x;
                    const long long i = a.row_begin + li;
                    double k = row.value(d, d8, a.kp);
                    double v;
                    const double av = a.A[li];
                    if constexpr (STRICT) {
                        if (a.lw) k = __dmul_rn(weight_pair<T, STRICT>(a, i, yw), k);
                        v = __dadd_rn(av, __dmul_rn(2.0, k));
                    } else {
                        v = a.lw ? fma(2.0 * weight_pair<T, STRICT>(a, i, yw), k, av) : fma(2.0, k, av);
                    }
                    a.A[li] = v;
                    if (cand_better(v, (unsigned)i, bv, bi)) { bv = v; bi = (unsigned)i; }
                }
            }
            __syncwarp();                                                    // every lane is done with buffer st
            if (lane == 0) dev_mbar_arrive(&empty_bar[st]);
        }
    }
    grid_argmin_publish(bv, bi, a);
}

// ---- tensor memory as a row store -----------------------------------------------------------------------------
// sm_100a has 256 KB of tensor memory (TMEM) per SM - 512 columns x 128 lanes x 32 bit - next to the 227 KB of shared
// memory; it is private to the CTA, read and written with tcgen05.ld / tcgen05.st (12-cycle latency, 64 B / clock
// per SM).  The single-launch kernel is bound by re-reading from L2, every step, the rows that do not fit in shared
// memory; parking another 8 rows per thread (d = 4; 16 for d = 2) in TMEM takes C3 from 3 to 11 of 14 rows on chip.
// Layout: warp w may touch lanes 32 (w % 4) .. + 31 only, so the four warps sharing a lane quarter each get 128 of
// the 512 columns; with the .32x32b shape thread i of a warp owns lane 32 (w % 4) + i, and one row (x and s, 2 D
// doubles = 4 D columns) is one tcgen05.ld.  All TMEM instructions are .sync.aligned: they sit in passes where every
// thread of the grid owns a row, outside any divergent branch.
// [round 1: opt-in with GFK_TMEM=1; executed on the CPU emulator of the test suite only, not yet on a GPU]
constexpr int kTmemCols = 512;

__device__ __forceinline__ void tmem_alloc_all(uint32_t* dst_smem) {      // one full warp
#if defined(GFK_SIMT_HOST)
    simt::tmem_alloc(dst_smem, kTmemCols);
#else
    const unsigned sa = (unsigned)__cvta_generic_to_shared(dst_smem);
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sa), "r"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
#endif
}
__device__ __forceinline__ void tmem_free_all(uint32_t base) {            // the same warp
#if defined(GFK_SIMT_HOST)
    simt::tmem_dealloc(base, kTmemCols);
#else
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(kTmemCols) : "memory");
#endif
}
__device__ __forceinline__ void tmem_fence_before_sync() {
#if !defined(GFK_SIMT_HOST)
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
#endif
}
__device__ __forceinline__ void tmem_fence_after_sync() {
#if !defined(GFK_SIMT_HOST)
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#endif
}
__device__ __forceinline__ void tmem_wait_st() {
#if !defined(GFK_SIMT_HOST)
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
#endif
}
__device__ __forceinline__ void tmem_wait_ld() {
#if !defined(GFK_SIMT_HOST)
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#endif
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
#if defined(GFK_SIMT_HOST)
    simt::tmem_st(taddr, r, 16);
#else
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
                   "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
#endif
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
#if defined(GFK_SIMT_HOST)
    simt::tmem_ld(taddr, r, 16);
#else
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr) : "memory");
#endif
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
#if defined(GFK_SIMT_HOST)
    simt::tmem_st(taddr, r, 8);
#else
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
#endif
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
#if defined(GFK_SIMT_HOST)
    simt::tmem_ld(taddr, r, 8);
#else
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr) : "memory");
#endif
}

// one row (x[D], s[D] in float64; D = 2 or 4) <-> 4 D tensor-memory columns of this thread's lane
template <int D>
__device__ __forceinline__ void tmem_put_row(uint32_t taddr, const double (&x)[D], const double (&sx)[D]) {
    static_assert(D == 2 || D == 4, "tensor-memory row store: d = 2 or 4");
    uint32_t r[4 * D];
#pragma unroll
    for (int k = 0; k < D; ++k) {
        const unsigned long long xb = (unsigned long long)__double_as_longlong(x[k]);
        const unsigned long long sb = (unsigned long long)__double_as_longlong(sx[k]);
        r[2 * k] = (uint32_t)xb;
        r[2 * k + 1] = (uint32_t)(xb >> 32);
        r[2 * D + 2 * k] = (uint32_t)sb;
        r[2 * D + 2 * k + 1] = (uint32_t)(sb >> 32);
    }
    if constexpr (D == 4) tmem_st16(taddr, r);
    else tmem_st8(taddr, r);
}
template <int D>
__device__ __forceinline__ void tmem_get_row(uint32_t taddr, double (&x)[D], double (&sx)[D]) {
    static_assert(D == 2 || D == 4, "tensor-memory row store: d = 2 or 4");
    uint32_t r[4 * D];
    if constexpr (D == 4) tmem_ld16(taddr, r);
    else tmem_ld8(taddr, r);
    tmem_wait_ld();
#pragma unroll
    for (int k = 0; k < D; ++k) {
        x[k] = __longlong_as_double((long long)(((unsigned long long)r[2 * k + 1] << 32) | r[2 * k]));
        sx[k] = __longlong_as_double((long long)(((unsigned long long)r[2 * D + 2 * k + 1] << 32) | r[2 * D + 2 * k]));
    }
}
// float32 rows: 2 D columns; only d = 4 (8 columns, one .x8 access) is instantiated
template <int D>
__device__ __forceinline__ void tmem_put_row(uint32_t taddr, const float (&x)[D], const float (&sx)[D]) {
    static_assert(D == 4, "tensor-memory row store (float32): d = 4");
    uint32_t r[8];
#pragma unroll
    for (int k = 0; k < D; ++k) {
        r[k] = __float_as_uint(x[k]);
        r[D + k] = __float_as_uint(sx[k]);
    }
    tmem_st8(taddr, r);
}
template <int D>
__device__ __forceinline__ void tmem_get_row(uint32_t taddr, float (&x)[D], float (&sx)[D]) {
    static_assert(D == 4, "tensor-memory row store (float32): d = 4");
    uint32_t r[8];
    tmem_ld8(taddr, r);
    tmem_wait_ld();
#pragma unroll
    for (int k = 0; k < D; ++k) {
        x[k] = __uint_as_float(r[k]);
        sx[k] = __uint_as_float(r[D + k]);
    }
}

// element types / dimensions with a tensor-memory instantiation, and the columns one row takes
template <typename T, int D>
struct TmemRow {
    static constexpr bool ok = (std::is_same<T, double>::value && (D == 2 || D == 4)) || (std::is_same<T, float>::value && D == 4);
    static constexpr int cols = 2 * D * (int)sizeof(T) / 4;
};

// ---- all m greedy steps in ONE cooperative launch (single rank, small D) -------------------------------
// One CTA per SM (or fewer for small shards).  Thread g of the grid owns rows g, g + T, g + 2T, ... for the
// whole job (T = threads in the grid).  The running objective A and the log-weights of the owned rows live
// in shared memory for all m steps; if the shard is small enough the owned x / s rows are copied to shared
// memory too (SoA, conflict-free) and the step loop touches no global memory at all except the selected
// row - otherwise x / s are re-read each step through L2 (no HBM traffic once the 126 MB L2 holds them).
// Per step: local update + argmin -> block argmin -> partial[parity][cta] -> grid barrier (arrival counter)
// -> every CTA reduces the <= 148 partials itself (no second barrier) -> next selected row.
constexpr int kPersistThreads = 512;
constexpr int kPollPerLane = 5;                       // persistent grid <= 160 CTAs (one per SM)

// WDD: a.w is present - strict mode: per-row double-double exponentials (exp_sum_dd); fast modes: plain exp(lw), the
// weight is a product.  Without it the weight is exp(lw_i + lw_y) per pair; as a template flag so that the loop of
// either variant does not carry the other's code.
// TM: passes [0, rpt_tm) of every thread's rows are parked in tensor memory (float64, d = 2 or 4), the next rpt_res
// passes in shared memory, the rest streams through L2
// cta / ncta: this CTA's index within, and the size of, the group of CTAs that runs job `a` - the whole grid for one
// job per launch, one slice of it per chain in the batched launch (every chain has its own workspace and slots)
template <typename T, int D, bool STRICT, bool MULTI, bool WDD, bool TM>
__device__ __forceinline__ void thin_persistent_body(const ThinArgs<T>& a, int m, int rpt, int rpt_res, int rpt_tm,
                                                     const unsigned cta, const unsigned ncta) {
    extern __shared__ __align__(16) unsigned char psm[];
    __shared__ double sv[32];
    __shared__ unsigned si[32];
    __shared__ unsigned long long win_sh;
    __shared__ int abort_sh;
    const int tid = threadIdx.x;
    const int nthr = blockDim.x;
    const long long TT = (long long)ncta * nthr;
    const long long gt = (long long)cta * nthr + tid;
    const long long rows = a.row_end - a.row_begin;
    // Row ownership: pass j < rpt - 1 gives thread g of the grid row j * TT + g (coalesced, every thread busy).  The
    // LAST pass is usually partial; dealing its rows to the first threads of the grid would hand whole CTAs one more
    // row per thread than others (measured: CTAs 0-29 of 148 finish C3's update ~1.9 us after the rest, and every step
    // waits for the slowest CTA).  Its rows are dealt round-robin over the CTAs instead, so all SMs carry the same load.
    const long long last_base = (long long)(rpt - 1) * TT;
    const long long last_e = (long long)tid * ncta + cta;
    const long long i_last = a.row_begin + last_base + last_e;     // this thread's global row of the last pass
    const int slots = rpt * nthr;
    double* A_sh = reinterpret_cast<double*>(psm);
    T* lw_sh = reinterpret_cast<T*>(A_sh + slots);              // lw (strict / no w) or w (fast)
    const int slots_res = rpt_res * nthr;                        // rpt_res passes of every thread's rows
    const int j0_res = TM ? rpt_tm : 0;                          // passes [0, rpt_tm): tensor memory; then shared memory
    T* xs = lw_sh + (a.lw ? slots : 0);                          // resident x as [D][slots_res], then s
    T* ss = xs + (size_t)D * slots_res;
    constexpr bool use_w = !STRICT && WDD;                       // fast modes: weight = w_i w_y (no exp code in the loop)
    ThinWs* ws = a.ws;
    if (tid == 0) abort_sh = 0;
    __shared__ uint32_t tmem_base_sh;
    uint32_t my_taddr = 0;
    if constexpr (TM) {
        if (tid < 32) tmem_alloc_all(&tmem_base_sh);
        tmem_fence_before_sync();
        __syncthreads();
        tmem_fence_after_sync();
        // lanes 32 (warp % 4) .. + 31, columns 128 (warp / 4) .. + 127 belong to this warp
        my_taddr = tmem_base_sh + ((uint32_t)((tid >> 5) & 3) * 32u << 16) + (uint32_t)(tid >> 7) * 128u;
    }

    // ---- step 0: diagonal, and stage the owned rows ----------------------------------------------------
    double bv = __longlong_as_double(0x7ff0000000000000LL);
    unsigned bi = 0xffffffffu;
    for (int j = 0; j < rpt; ++j) {
        const long long li = (j == rpt - 1) ? last_base + last_e : (long long)j * TT + gt;
        const int slot = j * nthr + tid;
        if (li < rows) {
            const long long i = a.row_begin + li;
            T xr[D], sr[D];
            load_row<D>(a.x, i, xr);
            load_row<D>(a.s, i, sr);
            if (j >= j0_res && j < j0_res + rpt_res) {
                const int rs = slot - j0_res * nthr;
#pragma unroll
                for (int k = 0; k < D; ++k) { xs[k * slots_res + rs] = xr[k]; ss[k * slots_res + rs] = sr[k]; }
            }
            double k0;
            if constexpr (STRICT) k0 = pair_strict<D>(xr, xr, sr, sr, a.kp);
            else k0 = (double)pair_fast<T, D, false>(xr, xr, sr, sr, xr, xr, a.kp);
            if (a.lw) {
                const T lwi = __ldg(a.lw + i);
                if (use_w) {
                    const T wi = __ldg(a.w + i);
                    lw_sh[slot] = wi;
                    k0 *= (double)(wi * wi);
                } else {
                    lw_sh[slot] = lwi;
                    double wgt;
                    if constexpr (STRICT && WDD) {
                        const double2 e = __ldg(reinterpret_cast<const double2*>(a.w) + i);
                        wgt = exp_sum_dd((double)lwi, (double)lwi, e.x, e.y, e.x, e.y);
                    } else {
                        wgt = exp((double)lwi + (double)lwi);
                    }
                    k0 = STRICT ? __dmul_rn(wgt, k0) : wgt * k0;
                }
            }
            A_sh[slot] = k0;
            if (cand_better_in_order(k0, bv, bi == 0xffffffffu)) { bv = k0; bi = (unsigned)i; }   // rows in increasing order
        }
    }

    if constexpr (TM) {
        // park the rows of the full passes [0, rpt_tm) in tensor memory (every thread owns a row there)
        for (int jt = 0; jt < rpt_tm; ++jt) {
            const long long i = a.row_begin + (long long)jt * TT + gt;
            T xr[D], sr[D];
            load_row<D>(a.x, i, xr);
            load_row<D>(a.s, i, sr);
            tmem_put_row<D>(my_taddr + (uint32_t)(jt * TmemRow<T, D>::cols), xr, sr);
        }
        tmem_wait_st();
    }

    for (int step = 0; step < m; ++step) {
        // ---- grid-wide argmin of the running objective ----------------------------------------------
        // flag-in-data exchange: each CTA publishes its block minimum as 3 tagged 8-byte words; warp 0 of
        // every CTA polls all slots and reduces them itself, so there is no separate barrier, no atomic and
        // no fence on the critical path.  Slots alternate by step parity (a CTA can be at most one step ahead).
        GFK_STAMP(step, 1);
        block_argmin(bv, bi, sv, si);
        GFK_STAMP(step, 2);
        GFK_STAMP_CTA(step);
        {
            unsigned long long (*slots_ll)[4] = ws->ll[step & 1];
            const unsigned tag16 = (unsigned)(step + 1);
            if (tid == 0) slot_store(slots_ll[cta], tag16, bv, bi);
            if (tid < 32) {
                double wv = __longlong_as_double(0x7ff0000000000000LL);
                unsigned wi = 0xffffffffu;
                unsigned dead = 0;                                  // kAbort* code, warp-uniform after each stage
                unsigned long long* bc = ws->bcast[step & 1];
                const unsigned long long limit = ws->spin_ns;
                // the first cross-rank exchange absorbs the skew between the ranks' host threads
                const unsigned long long limit_x = (MULTI && step == 0) ? ws->spin_first_ns : limit;
                if (!MULTI || cta == 0) {
                    // gather.  Single rank: warp 0 of EVERY CTA polls all per-CTA slots and reduces them itself
                    // (one store->load hop per step).  Multi rank: only CTA 0 gathers, then talks to the peers
                    // and re-publishes the global winner for the other CTAs (measured: fewer pollers on lines a
                    // peer writes, 16.0 -> 12.4 us per step at 2 x 0.5 M rows).  Lane l owns slots l, l+32, ... and
                    // requests all their words together, so a polling round costs one L2 latency.
                    unsigned long long w0[kPollPerLane], w1[kPollPerLane];
                    unsigned pending = 0;
#pragma unroll
                    for (int q = 0; q < kPollPerLane; ++q)
                        if (tid + 32u * q < ncta) pending |= 1u << q;
                    SpinDeadline dl;
#pragma unroll 1
                    while (pending) {
#pragma unroll
                        for (int q = 0; q < kPollPerLane; ++q) {
                            if (pending & (1u << q)) {
                                slot_load(slots_ll[tid + 32 * q], w0[q], w1[q]);
                            }
                        }
#pragma unroll
                        for (int q = 0; q < kPollPerLane; ++q) {
                            if ((pending & (1u << q)) && slot_ready(w0[q], w1[q], tag16)) {
                                pending &= ~(1u << q);
                                double pv;
                                unsigned pi;
                                slot_decode(w0[q], w1[q], pv, pi);
                                if (cand_better(pv, pi, wv, wi)) { wv = pv; wi = pi; }
                            }
                        }
                        if (pending) {
                            if (ld_volatile_u32(&ws->abort) || dl.expired(limit)) { dead = kAbortTimeout; break; }
                            GFK_SPIN_HINT();
                        }
                    }
                    dead = __any_sync(0xffffffffu, dead != 0u) ? kAbortTimeout : 0u;
                    warp_argmin(wv, wi);
                    GFK_STAMP(step, 3);
                    if (MULTI && !dead) {
                        // cross-GPU stage: store this rank's winner into every peer's mailbox over NVLink, wait
                        // for all ranks' winners in the local mailbox, reduce
                        const Cand g = mail_exchange(ws->mail, a.n_ranks, step, wv, wi, tid, true, dead, &ws->abort, limit_x);
                        wv = g.v;
                        wi = (unsigned)g.i;
                    }
                    if (!dead && (unsigned long long)wi >= (unsigned long long)a.n) dead = kAbortIndex;
                    if (MULTI && dead && dead != kAbortPeer) mail_poison(ws->mail, a.n_ranks, tid);   // peers leave too
                    // broadcast: one flag-in-data slot that the other CTAs poll (nothing on failure: they see ws->abort)
                    if (MULTI && !dead && tid == 0) slot_store(bc, tag16, wv, wi);
                } else {
                    unsigned long long w0 = 0, w1 = 0;
                    if (tid == 0) {
                        SpinDeadline dl;
#pragma unroll 1
                        for (;;) {
                            slot_load(bc, w0, w1);
                            if (slot_ready(w0, w1, tag16)) break;
                            if (ld_volatile_u32(&ws->abort) || dl.expired(limit + 2 * limit_x)) { dead = kAbortTimeout; break; }
                            GFK_SPIN_HINT();
                        }
                    }
                    dead = __shfl_sync(0xffffffffu, dead, 0);
                    w0 = __shfl_sync(0xffffffffu, w0, 0);
                    w1 = __shfl_sync(0xffffffffu, w1, 0);
                    slot_decode(w0, w1, wv, wi);
                    if (!dead && (unsigned long long)wi >= (unsigned long long)a.n) dead = kAbortIndex;
                }
                GFK_STAMP(step, 4);
                if (tid == 0) {
                    if (dead) {
                        abort_sh = 1;
                        if (ld_volatile_u32(&ws->abort) == 0u) ws->abort = dead;
                        wi = 0u;                                   // never used: every thread leaves below
                    }
                    win_sh = wi;
                    if (cta == 0 && !dead) {
                        a.idx_out[step] = wi;
                        if (step == m - 1) { ws->send.v = wv; ws->send.i = wi; ws->recv[0] = ws->send; ws->step = (unsigned)m; }
                    }
                }
            }
            __syncthreads();
            if (abort_sh) break;
        }
        GFK_STAMP(step, 5);
        if (step == m - 1) break;
        const long long sel = (long long)win_sh;

        // ---- add 2 k(x_i, y) for the new selection y ---------------------------------------------------
        T y[D], sy[D];
        load_row<D>(a.x, sel, y);
        load_row<D>(a.s, sel, sy);
        T lwy = T(0);
        if (a.lw) lwy = use_w ? __ldg(a.w + sel) : __ldg(a.lw + sel);
        double ehy = 0.0, ely = 0.0;
        if constexpr (STRICT && WDD) {
            if (a.lw) {
                const double2 e = __ldg(reinterpret_cast<const double2*>(a.w) + sel);
                ehy = e.x;
                ely = e.y;
            }
        }
        bv = __longlong_as_double(0x7ff0000000000000LL);
        bi = 0xffffffffu;
        GFK_STAMP(step + 1, 0);
        // one row of the update: kernel value against y, weight, running objective, in-order argmin
        auto update_row = [&](long long i, int slot, const T (&xr)[D], const T (&sr)[D]) {
            double2 ei = make_double2(0.0, 0.0);
            if constexpr (STRICT && WDD) {
                if (a.lw) ei = __ldg(reinterpret_cast<const double2*>(a.w) + i);   // in flight under the pair kernel
            }
            double k;
            if constexpr (STRICT) k = pair_strict<D>(xr, y, sr, sy, a.kp);
            else k = (double)pair_fast<T, D, false>(xr, y, sr, sy, xr, y, a.kp);
            double v;
            const double av = A_sh[slot];
            if constexpr (STRICT) {
                if (a.lw) {
                    double wgt;
                    if constexpr (STRICT && WDD) wgt = exp_sum_dd((double)lw_sh[slot], (double)lwy, ei.x, ei.y, ehy, ely);
                    else wgt = exp((double)lw_sh[slot] + (double)lwy);
                    k = __dmul_rn(wgt, k);
                }
                v = __dadd_rn(av, __dmul_rn(2.0, k));
            } else {
                if (a.lw) {
                    const double wgt = use_w ? (double)(lw_sh[slot] * lwy) : exp((double)lw_sh[slot] + (double)lwy);
                    v = fma(2.0 * wgt, k, av);
                } else {
                    v = fma(2.0, k, av);
                }
            }
            A_sh[slot] = v;
            if (cand_better_in_order(v, bv, bi == 0xffffffffu)) { bv = v; bi = (unsigned)i; }
        };
        // Two loops so that neither carries the other's address arithmetic: rows resident in shared memory first,
        // then the rows streamed through L2.  The global row index advances by TT per pass; the last pass uses the
        // round-robin dealing (i_last).  Rows are visited in increasing index order (cand_better_in_order).
        long long i_cur = a.row_begin + gt;
        int slot = tid;
        if constexpr (TM) {
            // rows parked in tensor memory: full passes only, so no thread sits out and the warp stays converged
#pragma unroll 2
            for (int jt = 0; jt < rpt_tm; ++jt, i_cur += TT, slot += nthr) {
                T xr[D], sr[D];
                tmem_get_row<D>(my_taddr + (uint32_t)(jt * TmemRow<T, D>::cols), xr, sr);
                update_row(i_cur, slot, xr, sr);
            }
        }
#pragma unroll 2
        for (int j = j0_res; j < j0_res + rpt_res; ++j, i_cur += TT, slot += nthr) {
            const long long i = (j == rpt - 1) ? i_last : i_cur;
            if (i < a.row_end) {
                const int rs = slot - j0_res * nthr;
                T xr[D], sr[D];
#pragma unroll
                for (int k = 0; k < D; ++k) { xr[k] = xs[k * slots_res + rs]; sr[k] = ss[k * slots_res + rs]; }
                update_row(i, slot, xr, sr);
            }
        }
#pragma unroll 2
        for (int j = j0_res + rpt_res; j < rpt; ++j, i_cur += TT, slot += nthr) {
            const long long i = (j == rpt - 1) ? i_last : i_cur;
            if (i < a.row_end) {
                T xr[D], sr[D];
                load_row<D>(a.x, i, xr);
                load_row<D>(a.s, i, sr);
                update_row(i, slot, xr, sr);
            }
        }
    }
    // ---- write the running objective back -----------------------------------------------------------
    for (int j = 0; j < rpt; ++j) {
        const long long li = (j == rpt - 1) ? last_base + last_e : (long long)j * TT + gt;
        if (li < rows) a.A[li] = A_sh[j * nthr + tid];
    }
    if constexpr (TM) {
        tmem_fence_before_sync();
        __syncthreads();                       // every warp is done with its tensor-memory rows
        tmem_fence_after_sync();
        if (tid < 32) tmem_free_all(tmem_base_sh);
    }
}

template <typename T, int D, bool STRICT, bool MULTI, bool WDD, bool TM>
__global__ void __launch_bounds__(kPersistThreads, 1) thin_persistent_kernel(ThinArgs<T> a, int m, int rpt, int rpt_res,
                                                                             int rpt_tm) {
    thin_persistent_body<T, D, STRICT, MULTI, WDD, TM>(a, m, rpt, rpt_res, rpt_tm, blockIdx.x, gridDim.x);
}

// ---- several independent chains in ONE cooperative launch --------------------------
// The reference thins the five chains of an experiment with five calls (joblib over chains,
// code/notebooks/lotka_volterra/Stein_thinning.ipynb#c15-#c17).  A greedy step of a small chain is latency, not work
// (C2-sized: ~3 us of exchange per 2.6 us of update), so five chains run one after the other pay the floor five
// times.  Here the grid is split into one group of CTAs per chain; each group runs the single-launch loop above on its
// own job, workspace and flag-in-data slots, and the groups never talk to each other.
template <typename T>
struct ThinBatchJob {
    ThinArgs<T> a;
    int rpt, rpt_res;
};

template <typename T, int D, bool STRICT, bool WDD>
__global__ void __launch_bounds__(kPersistThreads, 1) thin_persistent_batched_kernel(const ThinBatchJob<T>* __restrict__ jobs,
                                                                                     int ncta_per_chain, int m) {
    const unsigned chain = blockIdx.x / (unsigned)ncta_per_chain;
    const unsigned cta = blockIdx.x - chain * (unsigned)ncta_per_chain;
    const ThinBatchJob<T> job = jobs[chain];
    thin_persistent_body<T, D, STRICT, false, WDD, false>(job.a, m, job.rpt, job.rpt_res, 0, cta, (unsigned)ncta_per_chain);
}

// winner of the last exchange -> idx_out[m - 1]
template <typename T>
__global__ void thin_finish_kernel(ThinArgs<T> a) {
    const Cand win = pick_winner(a);
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        a.idx_out[a.step - 1] = (uint32_t)win.i;
        a.ws->step = (unsigned)a.step;
    }
}

// GFK_SPIN_MS / GFK_SPIN_FIRST_MS (milliseconds; defaults 200 and 5000) -> nanoseconds
static unsigned long long env_ms_to_ns(const char* name, double def_ms) {
    double ms = def_ms;
    if (const char* e = getenv(name)) {
        char* end = nullptr;
        const double v = strtod(e, &end);
        if (end != e && v > 0.0 && v < 3.6e6) ms = v;
    }
    return (unsigned long long)(ms * 1e6);
}
#if defined(GFK_SIMT_HOST)
constexpr double kSpinMsDefault = 120000.0;   // emulated CTAs are OS threads sharing a few cores
#else
constexpr double kSpinMsDefault = 200.0;
#endif
static unsigned long long spin_limit_ns() { return env_ms_to_ns("GFK_SPIN_MS", kSpinMsDefault); }
static unsigned long long spin_first_limit_ns() {
    const unsigned long long a = env_ms_to_ns("GFK_SPIN_FIRST_MS", 5000.0), b = spin_limit_ns();
    return a > b ? a : b;
}

__global__ void thin_reset_kernel(ThinWs* ws, MailArgs mail, unsigned long long spin_ns, unsigned long long spin_first_ns) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        ws->counter = 0;
        ws->step = 0;
        ws->abort = 0;
        ws->mail = mail;
        ws->spin_ns = spin_ns;
        ws->spin_first_ns = spin_first_ns;
    }
    unsigned long long* ll = &ws->ll[0][0][0];
    for (int e = threadIdx.x; e < 2 * (kMaxBlocks / 2) * 4; e += blockDim.x) ll[e] = 0ull;
    if (threadIdx.x < 8) (&ws->bcast[0][0])[threadIdx.x] = 0ull;
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

template <typename T>
static size_t step_tma_smem(int d) {
    return 1024 + (size_t)kTmaStages * 2 * kTmaTileBytes + sizeof(T) * 2 * (size_t)d;
}

// tensor map of one array for the TMA-fed step kernel: box = [128 bytes of dims] x [kTmaRows rows]
template <typename T>
static bool make_step_tensor_map(TensorMap* map, const T* base, long long n, int d) {
    return make_tensor_map_2d(map, base, (unsigned)sizeof(T), (unsigned long long)d, (unsigned long long)n,
                              (unsigned)TmaSlab<T>::K, (unsigned)kTmaRows);
}

template <typename T, bool STRICT>
static int launch_step(const ThinArgs<T>& a, int d, cudaStream_t st, int grid_small, int grid_tiled,
                       int tiled_threads, size_t tiled_smem, const TensorMap* maps) {
    const bool dense = a.xl != nullptr;
    if (!dense && d == 2) {
        thin_step_small_kernel<T, 2, STRICT, 4><<<grid_small, kThreads, 0, st>>>(a);
    } else if (!dense && d == 4) {
        thin_step_small_kernel<T, 4, STRICT, 2><<<grid_small, kThreads, 0, st>>>(a);
    } else if (!dense && d == 1) {
        thin_step_small_kernel<T, 1, STRICT, 4><<<grid_small, kThreads, 0, st>>>(a);
    } else if (!dense && d == 3) {
        thin_step_small_kernel<T, 3, STRICT, 2><<<grid_small, kThreads, 0, st>>>(a);
    } else if (tiled_threads == -1) {
        thin_step_tma_kernel<T, STRICT><<<grid_tiled, kTmaThreads, step_tma_smem<T>(d), st>>>(a, maps[0], maps[1]);
    } else if (tiled_threads == 0) {
        const size_t smem = sizeof(T) * (size_t)(3 * d + (dense ? 3 : 2) * kThreads * kChunkLd);
        if (dense) thin_step_chunked_kernel<T, STRICT, !STRICT><<<grid_tiled, kThreads, smem, st>>>(a);
        else thin_step_chunked_kernel<T, STRICT, false><<<grid_tiled, kThreads, smem, st>>>(a);
    } else {
        thin_step_tiled_kernel<T, STRICT><<<grid_tiled, tiled_threads, tiled_smem, st>>>(a);
    }
    return GFK_OK;
}


// shared-memory opt-in + "does one CTA fit an SM" for a persistent kernel; cudaSuccess if it can be launched
template <typename K>
static cudaError_t persistent_prepare(K kernel, int threads, size_t smem) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { cudaGetLastError(); return e; }
    int occ = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, threads, smem);
    if (e != cudaSuccess || occ < 1) { cudaGetLastError(); return e != cudaSuccess ? e : cudaErrorInvalidValue; }
    return cudaSuccess;
}

// cooperative launch (all CTAs co-resident: the kernels spin on each other's flags) with typed arguments
template <typename... KArgs, typename... Args>
static cudaError_t launch_cooperative(void (*kernel)(KArgs...), unsigned grid, int threads, size_t smem, cudaStream_t st,
                                      Args... args) {
    // cudaLaunchCooperativeKernel reads each argument through an untyped pointer: the types must be the kernel's own
    static_assert(sizeof...(KArgs) == sizeof...(Args), "argument count differs from the kernel's parameter list");
    static_assert((std::is_same<KArgs, Args>::value && ...), "argument types differ from the kernel's parameter list");
#if defined(GFK_SIMT_HOST)
    (void)st;
    simt::launch(dim3(grid), dim3((unsigned)threads), smem, true, [=]() { kernel(args...); });
    return cudaGetLastError();
#else
    void* params[] = {(void*)&args...};
    return cudaLaunchCooperativeKernel((const void*)kernel, dim3(grid), dim3((unsigned)threads), params, smem, st);
#endif
}

// one instantiation of the single-launch kernel: opt-in shared memory, reset, cooperative launch
template <typename T, int D, bool STRICT, bool MULTI, bool WDD, bool TM>
static int persistent_go2(const ThinArgs<T>& a, const MailArgs& mail, int m, int rpt, int rpt_res, int rpt_tm,
                          unsigned grid, int threads, size_t smem, cudaStream_t st) {
    cudaError_t e = persistent_prepare(thin_persistent_kernel<T, D, STRICT, MULTI, WDD, TM>, threads, smem);
    if (e != cudaSuccess) return 0;
    thin_reset_kernel<<<1, 1024, 0, st>>>(a.ws, mail, spin_limit_ns(), spin_first_limit_ns());
    e = launch_cooperative(thin_persistent_kernel<T, D, STRICT, MULTI, WDD, TM>, grid, threads, smem, st, a, m, rpt,
                           rpt_res, rpt_tm);
    if (e != cudaSuccess) return -cuda_fail(e, "cudaLaunchCooperativeKernel(thin_persistent_kernel)", __FILE__, __LINE__);
    count_launches(2);
    return 1;
}

// rpt_tm > 0 selects the tensor-memory instantiation (float64, d = 2 or 4 only)
template <typename T, int D, bool STRICT, bool MULTI, bool WDD>
static int persistent_go(const ThinArgs<T>& a, const MailArgs& mail, int m, int rpt, int rpt_res, int rpt_tm,
                         unsigned grid, int threads, size_t smem, cudaStream_t st) {
    if constexpr (TmemRow<T, D>::ok) {
        if (rpt_tm > 0)
            return persistent_go2<T, D, STRICT, MULTI, WDD, true>(a, mail, m, rpt, rpt_res, rpt_tm, grid, threads, smem, st);
    }
    return persistent_go2<T, D, STRICT, MULTI, WDD, false>(a, mail, m, rpt, rpt_res, 0, grid, threads, smem, st);
}

// persistent path: returns 1 if it ran, 0 if not eligible, < 0 on error
template <typename T, int D, bool STRICT>
static int try_persistent_d(const ThinArgs<T>& a, const MailArgs& mail, int m, cudaStream_t st) {
    const long long rows = a.row_end - a.row_begin;
    int dev = 0, coop = 0, smem_optin = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
    cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (!coop) return 0;
    const int sms = sm_count();
    const int threads = kPersistThreads;
    long long grid = (rows + threads - 1) / threads;
    if (grid > sms) grid = sms;
    if (grid > 32 * kPollPerLane) grid = 32 * kPollPerLane;
    if (grid < 1) grid = 1;                              // an empty shard still takes part in every exchange
    const long long TT = grid * threads;
    int rpt = (int)((rows + TT - 1) / TT);
    if (rpt < 1) rpt = 1;
    const size_t slots = (size_t)rpt * threads;
    const size_t base = slots * (sizeof(double) + (a.lw ? sizeof(T) : 0));
    const size_t budget = (size_t)smem_optin - 1024;   // minus the kernel's static shared memory
    if (base > budget) return 0;
    // Tensor memory first (opt-in GFK_TMEM=1; float64 d = 2 or 4, float32 d = 4): full passes only - every thread owns a row there.
    int rpt_tm = 0;
    if (TmemRow<T, D>::ok) {
        const char* tenv = getenv("GFK_TMEM");
        if (tenv && tenv[0] == '1') {
            rpt_tm = (kTmemCols / 4) / TmemRow<T, D>::cols;       // 128 columns per warp / columns per row
            if (rpt_tm > rpt - 1) rpt_tm = rpt - 1;
        }
    }
    // then as many of the remaining passes as fit stay in shared memory for the whole job; the rest stream via L2
    const size_t per_row_slot = 2 * (size_t)D * threads * sizeof(T);
    int rpt_res = (int)((budget - base) / per_row_slot);
    if (rpt_res > rpt - rpt_tm) rpt_res = rpt - rpt_tm;
    const char* env = getenv("GFK_RESIDENT");
    if (env && env[0] == '0') rpt_res = 0;
    const size_t smem = base + (size_t)rpt_res * per_row_slot;
    const unsigned g = (unsigned)grid;
    {
        if (a.lw && a.w) {       // pre-exponentiated weights present: no device exp in the loop
            return a.n_ranks > 1 ? persistent_go<T, D, STRICT, true, true>(a, mail, m, rpt, rpt_res, rpt_tm, g, threads, smem, st)
                                 : persistent_go<T, D, STRICT, false, true>(a, mail, m, rpt, rpt_res, rpt_tm, g, threads, smem, st);
        }
    }
    return a.n_ranks > 1 ? persistent_go<T, D, STRICT, true, false>(a, mail, m, rpt, rpt_res, rpt_tm, g, threads, smem, st)
                         : persistent_go<T, D, STRICT, false, false>(a, mail, m, rpt, rpt_res, rpt_tm, g, threads, smem, st);
}


template <typename T, bool STRICT>
static int try_persistent(const ThinArgs<T>& a, const MailArgs& mail, int d, int m, cudaStream_t st) {
    const char* env = getenv("GFK_PERSISTENT");
    if (env && env[0] == '0') return 0;
    if ((a.n_ranks != 1 && !a.mail_enabled) || a.xl != nullptr) return 0;
    switch (d) {
        case 1: return try_persistent_d<T, 1, STRICT>(a, mail, m, st);
        case 2: return try_persistent_d<T, 2, STRICT>(a, mail, m, st);
        case 3: return try_persistent_d<T, 3, STRICT>(a, mail, m, st);
        case 4: return try_persistent_d<T, 4, STRICT>(a, mail, m, st);
        default: return 0;
    }
}

template <typename T, bool STRICT>
static int thin_typed(const gfk_problem* p, int64_t row_begin, int64_t row_end, int32_t m,
                      double* objective, uint32_t* idx_out, void* ws_raw, cudaStream_t st, MailArgs mail) {
    ThinArgs<T> a;
    a.x = static_cast<const T*>(p->x);
    a.s = static_cast<const T*>(p->s);
    a.lw = static_cast<const T*>(p->lw);
    a.w = static_cast<const T*>(p->w);
    a.xl = static_cast<const T*>(p->xl);
    a.A = objective;
    a.row_begin = row_begin;
    a.row_end = row_end;
    a.n = p->n;
    a.kp = KParams{p->lambda, p->lambda_sq, p->trace, p->c, p->d, p->xl != nullptr};
    a.ws = static_cast<ThinWs*>(ws_raw);
    a.idx_out = idx_out;
    a.n_ranks = comm_size();
    a.step = 0;
    a.mail_enabled = mail.enabled;
    const int d = p->d;
    const long long rows = row_end - row_begin;
    const int sms = sm_count();

    int tiled_threads = 0;
    size_t tiled_smem = 0;
    const bool uses_tiled = (a.xl != nullptr) || !(d >= 1 && d <= 4);
    // generic d: the chunked kernel, except strict d > 128 (numpy's recursive pairwise order needs the
    // whole row at once) or y rows too long for its shared memory -> the row-tiled kernel
    const size_t chunked_smem = sizeof(T) * (size_t)(3 * d + (a.xl ? 3 : 2) * kThreads * kChunkLd);
    const bool use_chunked = uses_tiled && !(STRICT && d > 128) && chunked_smem <= 200 * 1024;
    // TMA-fed step kernel (opt-in): scalar preconditioner, rows a multiple of 16 bytes (the tensor-map stride rule)
    bool use_pipe = false;
    TensorMap maps[2];
    if (const char* penv = getenv("GFK_STEP_TMA")) {
        use_pipe = penv[0] == '1' && uses_tiled && use_chunked && a.xl == nullptr && ((size_t)d * sizeof(T)) % 16 == 0 &&
                   reinterpret_cast<uintptr_t>(a.x) % 16 == 0 && reinterpret_cast<uintptr_t>(a.s) % 16 == 0 &&
                   a.n < 0x7fffffffLL && step_tma_smem<T>(d) <= 220 * 1024 &&
                   make_step_tensor_map<T>(&maps[0], a.x, a.n, d) && make_step_tensor_map<T>(&maps[1], a.s, a.n, d);
    }
    if (use_pipe) {
        GFK_CUDA(cudaFuncSetAttribute(thin_step_tma_kernel<T, STRICT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)step_tma_smem<T>(d)));
        tiled_threads = -1;                          // launch_step: the TMA-fed kernel
    } else if (uses_tiled && use_chunked) {
        if (a.xl) GFK_CUDA(cudaFuncSetAttribute(thin_step_chunked_kernel<T, STRICT, !STRICT>,
                                                cudaFuncAttributeMaxDynamicSharedMemorySize, (int)chunked_smem));
        else GFK_CUDA(cudaFuncSetAttribute(thin_step_chunked_kernel<T, STRICT, false>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)chunked_smem));
    } else if (uses_tiled) {
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
    {
        const int pr = try_persistent<T, STRICT>(a, mail, d, m, st);
        if (pr < 0) return -pr;
        if (pr == 1) return GFK_OK;
    }
    const int grid_init = grid_for(kThreads, 8);
    const int grid_small = grid_for(kThreads * 2, 4);
    const int grid_tiled = !uses_tiled ? 1
                           : use_pipe ? grid_for(kTmaRows, 1)
                           : use_chunked ? grid_for(kThreads, chunked_smem > 72 * 1024 ? 2 : 3)
                                         : grid_for(tiled_threads, tiled_smem > 50 * 1024 ? 2 : 4);

    thin_reset_kernel<<<1, 1024, 0, st>>>(a.ws, mail, spin_limit_ns(), spin_first_limit_ns());
    a.step = 0;                                  // produces exchange 0
    thin_init_kernel<T, STRICT><<<grid_init, kThreads, 0, st>>>(a);
    GFK_CHECK_LAUNCH();
    const bool nccl_exchange = a.n_ranks > 1 && !a.mail_enabled;
    for (int step = 1; step <= m; ++step) {
        // exchange the candidates of step-1 (peer mailboxes: inside the kernels), then continue or finish
        if (nccl_exchange) {
            int rc = comm_allgather16(&a.ws->send, a.ws->recv, st);
            if (rc != GFK_OK) return rc;
        }
        a.step = step;                           // consumes exchange step-1, produces exchange step
        if (step == m) break;
        launch_step<T, STRICT>(a, d, st, grid_small, grid_tiled, tiled_threads, tiled_smem, maps);
    }
    thin_finish_kernel<T><<<1, 32, 0, st>>>(a);
    count_launches(2 + (long long)m);   // reset + init + (m - 1) steps + finish
    GFK_CHECK_LAUNCH();
    return GFK_OK;
}

// ---- batched chains: host side -------------------------------------------------------------------------------
constexpr size_t kBatchWsStride = (sizeof(ThinWs) + 255) / 256 * 256;

template <typename T, int D, bool STRICT, bool WDD>
static int thin_batched_go(const std::vector<ThinBatchJob<T>>& jobs, int ncta, int m, size_t smem, void* jobs_dev,
                           cudaStream_t st) {
    auto kernel = thin_persistent_batched_kernel<T, D, STRICT, WDD>;
    if (persistent_prepare(kernel, kPersistThreads, smem) != cudaSuccess) {
        set_error("gfk_thin_batched: the batched kernel does not fit this device (shared memory %zu bytes)", smem);
        return GFK_EUNSUPPORTED;
    }
    GFK_CUDA(cudaMemcpyAsync(jobs_dev, jobs.data(), sizeof(ThinBatchJob<T>) * jobs.size(), cudaMemcpyHostToDevice, st));
    MailArgs no_mail;
    memset(&no_mail, 0, sizeof(no_mail));
    for (const ThinBatchJob<T>& j : jobs)
        thin_reset_kernel<<<1, 1024, 0, st>>>(j.a.ws, no_mail, spin_limit_ns(), spin_first_limit_ns());
    const cudaError_t e = launch_cooperative(kernel, (unsigned)(jobs.size() * (size_t)ncta), kPersistThreads, smem, st,
                                             static_cast<const ThinBatchJob<T>*>(jobs_dev), ncta, m);
    if (e != cudaSuccess) return cuda_fail(e, "cudaLaunchCooperativeKernel(thin_persistent_batched_kernel)", __FILE__, __LINE__);
    count_launches(1 + (long long)jobs.size());
    return GFK_OK;
}

template <typename T, bool STRICT>
static int thin_batched_typed(const gfk_problem* probs, int n_chains, int m, double* const* objectives,
                              uint32_t* const* idx_outs, void* ws_raw, cudaStream_t st) {
    int dev = 0, coop = 0, smem_optin = 0;
    GFK_CUDA(cudaGetDevice(&dev));
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
    cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (!coop) { set_error("gfk_thin_batched: the device has no cooperative launch"); return GFK_EUNSUPPORTED; }
    const int d = probs[0].d;
    const bool wdd = probs[0].lw && probs[0].w;
    long long max_rows = 0;
    for (int c = 0; c < n_chains; ++c) max_rows = probs[c].n > max_rows ? probs[c].n : max_rows;
    // one group of CTAs per chain, all groups the same size, one CTA per SM at most
    long long ncta = sm_count() / n_chains;
    const long long want = (max_rows + kPersistThreads - 1) / kPersistThreads;
    if (ncta > want) ncta = want;
    if (ncta > 32 * kPollPerLane) ncta = 32 * kPollPerLane;
    if (ncta < 1) { set_error("gfk_thin_batched: more chains (%d) than SMs", n_chains); return GFK_EUNSUPPORTED; }
    const long long TT = ncta * kPersistThreads;
    const size_t budget = (size_t)smem_optin - 1024;
    const size_t per_row_slot = 2 * (size_t)d * kPersistThreads * sizeof(T);
    std::vector<ThinBatchJob<T>> jobs((size_t)n_chains);
    size_t smem = 0;
    char* ws_bytes = static_cast<char*>(ws_raw);
    for (int c = 0; c < n_chains; ++c) {
        const gfk_problem& p = probs[c];
        ThinBatchJob<T>& j = jobs[(size_t)c];
        j.a.x = static_cast<const T*>(p.x);
        j.a.s = static_cast<const T*>(p.s);
        j.a.lw = static_cast<const T*>(p.lw);
        j.a.w = static_cast<const T*>(p.w);
        j.a.xl = nullptr;
        j.a.A = objectives[c];
        j.a.row_begin = 0;
        j.a.row_end = p.n;
        j.a.n = p.n;
        j.a.kp = KParams{p.lambda, p.lambda_sq, p.trace, p.c, p.d, false};
        j.a.ws = reinterpret_cast<ThinWs*>(ws_bytes + (size_t)c * kBatchWsStride);
        j.a.idx_out = idx_outs[c];
        j.a.n_ranks = 1;
        j.a.step = 0;
        j.a.mail_enabled = 0;
        int rpt = (int)((p.n + TT - 1) / TT);
        if (rpt < 1) rpt = 1;
        const size_t base = (size_t)rpt * kPersistThreads * (sizeof(double) + (p.lw ? sizeof(T) : 0));
        if (base > budget) {
            set_error("gfk_thin_batched: chain %d (n = %lld) does not fit the shared memory of its %lld CTAs", c,
                      (long long)p.n, ncta);
            return GFK_EUNSUPPORTED;
        }
        int rpt_res = (int)((budget - base) / per_row_slot);
        if (rpt_res > rpt) rpt_res = rpt;
        j.rpt = rpt;
        j.rpt_res = rpt_res;
        const size_t need = base + (size_t)rpt_res * per_row_slot;
        smem = need > smem ? need : smem;
    }
    void* jobs_dev = ws_bytes + (size_t)n_chains * kBatchWsStride;
    switch (d) {
#define GFK_BATCH_CASE(DD)                                                                                              \
        case DD:                                                                                                        \
            return wdd ? thin_batched_go<T, DD, STRICT, true>(jobs, (int)ncta, m, smem, jobs_dev, st)                   \
                       : thin_batched_go<T, DD, STRICT, false>(jobs, (int)ncta, m, smem, jobs_dev, st);
        GFK_BATCH_CASE(1)
        GFK_BATCH_CASE(2)
        GFK_BATCH_CASE(3)
        GFK_BATCH_CASE(4)
#undef GFK_BATCH_CASE
        default: break;
    }
    set_error("gfk_thin_batched: d = %d (the single-launch kernel covers d <= 4)", d);
    return GFK_EUNSUPPORTED;
}

// exchange-only loop over the peer mailboxes: measures the per-step cost of the cross-GPU stage alone.
// failed: 0 ok, otherwise a kAbort* code (a rank that gives up tells its peers, so all leave within one deadline)
// or 0x100 for a record that arrived with the wrong content
__global__ void mail_selftest_kernel(MailArgs m, int n_ranks, int iters, unsigned long long first_ns,
                                     unsigned long long spin_ns, unsigned* failed) {
    const int lane = threadIdx.x;
    unsigned status = 0;
    for (int it = 0; it < iters; ++it) {
        unsigned dead;
        const Cand c = mail_exchange(m, n_ranks, it, (double)(it + 1), (unsigned)(m.rank + 1), lane, true, dead, nullptr,
                                     it == 0 ? first_ns : spin_ns);
        if (dead) {
            if (dead != kAbortPeer) mail_poison(m, n_ranks, lane);
            status = dead;
            break;
        }
        if (c.v != (double)(it + 1) || c.i != 1u) status = 0x100u;
    }
    if (lane == 0) *failed = status;
}

}  // namespace gfk

extern "C" int gfk_p2p_selftest(int iters, double* us_per_exchange_host, void* stream) {
    using namespace gfk;
    GFK_REQUIRE(iters >= 1 && us_per_exchange_host, "gfk_p2p_selftest: bad arguments");
    MailArgs mail;
    p2p_next_job(&mail, iters);
    GFK_REQUIRE(mail.enabled && comm_size() > 1, "gfk_p2p_selftest: peer mailboxes are not open");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    unsigned* failed = nullptr;
    GFK_CUDA(cudaMalloc(reinterpret_cast<void**>(&failed), sizeof(unsigned)));
    cudaEvent_t e0, e1;
    GFK_CUDA(cudaEventCreate(&e0));
    GFK_CUDA(cudaEventCreate(&e1));
    GFK_CUDA(cudaEventRecord(e0, st));
    mail_selftest_kernel<<<1, 32, 0, st>>>(mail, comm_size(), iters, spin_first_limit_ns(), spin_limit_ns(), failed);
    GFK_CUDA(cudaEventRecord(e1, st));
    GFK_CUDA(cudaEventSynchronize(e1));
    float ms = 0.f;
    GFK_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    unsigned bad = 0;
    GFK_CUDA(cudaMemcpy(&bad, failed, sizeof(bad), cudaMemcpyDeviceToHost));
    cudaFree(failed);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    count_launches(1);
    if (bad) {
        set_error("gfk_p2p_selftest: %s", bad == kAbortTimeout ? "a peer did not answer within the deadline (GFK_SPIN_MS)"
                                          : bad == kAbortPeer  ? "a peer rank gave up and said so"
                                                               : "a record arrived with the wrong content");
        return GFK_ECUDA;
    }
    *us_per_exchange_host = (double)ms * 1e3 / iters;
    return GFK_OK;
}

extern "C" size_t gfk_thin_ws_bytes(void) { return sizeof(gfk::ThinWs); }

extern "C" size_t gfk_thin_batched_ws_bytes(int32_t n_chains) {
    if (n_chains < 1) return 0;
    return (size_t)n_chains * (gfk::kBatchWsStride + (sizeof(gfk::ThinBatchJob<double>) + 255) / 256 * 256);
}

extern "C" int gfk_thin_batched(const gfk_problem* problems, int32_t n_chains, int32_t m, double* const* objectives,
                                uint32_t* const* idx_outs, void* ws, void* stream) {
    using namespace gfk;
    GFK_REQUIRE(problems && objectives && idx_outs && ws, "gfk_thin_batched: NULL pointer argument");
    GFK_REQUIRE(n_chains >= 1 && m >= 1, "gfk_thin_batched: need n_chains >= 1 and m >= 1");
    GFK_REQUIRE(reinterpret_cast<uintptr_t>(ws) % 256 == 0, "gfk_thin_batched: ws must be 256-byte aligned");
    for (int c = 0; c < n_chains; ++c) {
        const gfk_problem& p = problems[c];
        GFK_REQUIRE(p.x && p.s && objectives[c] && idx_outs[c], "gfk_thin_batched: NULL pointer in chain %d", c);
        GFK_REQUIRE(p.n >= 1 && p.n <= 0xfffffffeLL, "gfk_thin_batched: chain %d: n out of range", c);
        if (p.d != problems[0].d || p.mode != problems[0].mode || (p.lw != nullptr) != (problems[0].lw != nullptr) ||
            (p.w != nullptr) != (problems[0].w != nullptr) || p.xl != nullptr) {
            set_error("gfk_thin_batched: chains must share d, mode and kernel kind, with a scalar preconditioner");
            return GFK_EUNSUPPORTED;
        }
        const size_t row_bytes = (size_t)p.d * (p.mode == GFK_F32_FAST ? sizeof(float) : sizeof(double));
        const size_t al = row_bytes % 16 == 0 ? 16 : (row_bytes % 8 == 0 ? 8 : 4);
        GFK_REQUIRE(reinterpret_cast<uintptr_t>(p.x) % al == 0 && reinterpret_cast<uintptr_t>(p.s) % al == 0,
                    "gfk_thin_batched: chain %d: x and s must be %zu-byte aligned", c, al);
        if (p.mode == GFK_F64_STRICT && p.lw && p.w)
            GFK_REQUIRE(reinterpret_cast<uintptr_t>(p.w) % 16 == 0, "gfk_thin_batched: chain %d: w must be 16-byte aligned", c);
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    switch (problems[0].mode) {
        case GFK_F64_STRICT: return thin_batched_typed<double, true>(problems, n_chains, m, objectives, idx_outs, ws, st);
        case GFK_F64_FAST: return thin_batched_typed<double, false>(problems, n_chains, m, objectives, idx_outs, ws, st);
        case GFK_F32_FAST: return thin_batched_typed<float, false>(problems, n_chains, m, objectives, idx_outs, ws, st);
        default: set_error("gfk_thin_batched: unknown mode %d", problems[0].mode); return GFK_EINVAL;
    }
}

extern "C" int gfk_thin_batched_check(const void* ws, int32_t n_chains, void* stream) {
    using namespace gfk;
    GFK_REQUIRE(ws && n_chains >= 1, "gfk_thin_batched_check: bad arguments");
    for (int c = 0; c < n_chains; ++c) {
        const int rc = gfk_thin_check(static_cast<const char*>(ws) + (size_t)c * kBatchWsStride, stream);
        if (rc != GFK_OK) return rc;
    }
    return GFK_OK;
}

extern "C" int gfk_thin_check(const void* ws, void* stream) {
    using namespace gfk;
    GFK_REQUIRE(ws, "gfk_thin_check: NULL workspace");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    unsigned flag = 0;
    GFK_CUDA(cudaMemcpyAsync(&flag, static_cast<const char*>(ws) + offsetof(ThinWs, abort), sizeof(flag),
                             cudaMemcpyDeviceToHost, st));
    GFK_CUDA(cudaStreamSynchronize(st));
    if (flag) {
        set_error("gfk_thin: job aborted on the device (%s); the index vector is invalid",
                  flag == kAbortTimeout ? "a wait for another block or a peer rank ran past its deadline, GFK_SPIN_MS"
                  : flag == kAbortPeer  ? "a peer rank aborted the job"
                  : flag == kAbortIndex ? "an exchanged row index was out of range"
                                        : "unknown code");
        return GFK_ECUDA;
    }
    return GFK_OK;
}

extern "C" int gfk_thin(const gfk_problem* p, int64_t row_begin, int64_t row_end, int32_t m,
                        double* objective, uint32_t* idx_out, void* ws, void* stream) {
    using namespace gfk;
    // The mailbox epoch advances on EVERY call, before anything can fail on this rank alone: ranks make the same
    // sequence of gfk_thin calls, so a rank-local argument error must not leave the tags of later jobs out of step.
    MailArgs mail;
    p2p_next_job(&mail, m > 0 ? m : 1);
    {
        // Peer mailboxes are opt-in (GFK_P2P=1): the multi-rank default is the NCCL all-gather exchange.
        const char* penv = getenv("GFK_P2P");
        if (comm_size() == 1 || !(penv && penv[0] == '1')) mail.enabled = 0;
    }
    GFK_REQUIRE(p && p->x && p->s && idx_out && ws, "gfk_thin: NULL pointer argument");
    GFK_REQUIRE(objective || row_begin == row_end, "gfk_thin: NULL objective for a non-empty shard");
    GFK_REQUIRE(p->n >= 1 && p->n <= 0xfffffffeLL, "gfk_thin: n = %lld out of range [1, 2^32-2]", (long long)p->n);
    GFK_REQUIRE(p->d >= 1, "gfk_thin: d = %d must be >= 1", p->d);
    GFK_REQUIRE(m >= 1, "gfk_thin: m = %d must be >= 1", m);
    // an empty shard (more ranks than rows) is legal: the rank offers (+inf, no index) in every exchange
    GFK_REQUIRE(row_begin >= 0 && row_begin <= row_end && row_end <= p->n && (row_begin < row_end || comm_size() > 1),
                "gfk_thin: row range [%lld, %lld) invalid for n = %lld", (long long)row_begin, (long long)row_end,
                (long long)p->n);
    GFK_REQUIRE(!(p->xl && p->mode == GFK_F64_STRICT),
                "gfk_thin: a dense preconditioner needs a fast mode (numpy's BLAS order is not reproducible)");
    GFK_REQUIRE(reinterpret_cast<uintptr_t>(ws) % 16 == 0, "gfk_thin: ws must be 16-byte aligned");
    {
        const size_t esz = p->mode == GFK_F32_FAST ? sizeof(float) : sizeof(double);
        const size_t row_bytes = (size_t)p->d * esz;
        // small-d kernels read rows with 16-byte vector loads whenever the row size allows it
        if (p->d <= 4 && row_bytes % 16 == 0)
            GFK_REQUIRE(reinterpret_cast<uintptr_t>(p->x) % 16 == 0 && reinterpret_cast<uintptr_t>(p->s) % 16 == 0,
                        "gfk_thin: x and s must be 16-byte aligned for d = %d", p->d);
        else if (p->d <= 4 && row_bytes % 8 == 0)
            GFK_REQUIRE(reinterpret_cast<uintptr_t>(p->x) % 8 == 0 && reinterpret_cast<uintptr_t>(p->s) % 8 == 0,
                        "gfk_thin: x and s must be 8-byte aligned for d = %d", p->d);
    }
    if (p->mode == GFK_F64_STRICT && p->lw && p->w)
        GFK_REQUIRE(reinterpret_cast<uintptr_t>(p->w) % 16 == 0,
                    "gfk_thin: in the strict mode w holds (hi, lo) pairs and must be 16-byte aligned");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    switch (p->mode) {
        case GFK_F64_STRICT: return thin_typed<double, true>(p, row_begin, row_end, m, objective, idx_out, ws, st, mail);
        case GFK_F64_FAST: return thin_typed<double, false>(p, row_begin, row_end, m, objective, idx_out, ws, st, mail);
        case GFK_F32_FAST: return thin_typed<float, false>(p, row_begin, row_end, m, objective, idx_out, ws, st, mail);
        default: set_error("gfk_thin: unknown mode %d", p->mode); return GFK_EINVAL;
    }
}
