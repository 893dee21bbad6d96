/*
 * gfk.h - C ABI of libgfk: B200 (sm_100a) Stein thinning / gradient-free KSD.
 *
 * The reference (aglebov/gradient-free-mcmc-postprocessing) has no FFI of its
 * own: its hot path is the Python call surface of the third-party package
 * `stein_thinning` (reference: code/src/thinning.py:5,17; code/src/utils/ksd.py:5-6,25,27).
 * This header is therefore the boundary a maintainer binds with ctypes
 * (INTEGRATION.md shows the stub).  Each entry point names the reference
 * function whose arithmetic it replaces.
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / C++ types in any signature
 *   - every data pointer is a DEVICE pointer owned by the caller unless the
 *     parameter is documented as "host"; the library allocates nothing that
 *     outlives a call except the communicator created by gfk_comm_init
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream)
 *   - calls are stream-ordered and asynchronous unless documented otherwise
 *   - return value: 0 on success, a GFK_E* code otherwise; the message is in
 *     gfk_last_error() (thread local).  The library never aborts.
 *   - matrices are row-major (n, d), exactly the C-contiguous float64 arrays
 *     the reference passes around
 */
#ifndef GFK_H_
#define GFK_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GFK_VERSION 100

/* error codes */
#define GFK_OK 0
#define GFK_EINVAL 1    /* bad argument (shape, dtype, mode, NULL pointer)        */
#define GFK_ECUDA 2     /* a CUDA runtime call or kernel launch failed            */
#define GFK_ENCCL 3     /* NCCL missing, failed to load, or a collective failed   */
#define GFK_EUNSUPPORTED 4
#define GFK_ENUMERIC 5  /* non-finite input, zero scale, zero median              */

/* arithmetic of the pair kernel */
#define GFK_F64_STRICT 0 /* float64, numpy's operation order, no FMA contraction, correctly
                            rounded powers: index-exact parity mode                        */
#define GFK_F64_FAST 1   /* float64, FMA + rsqrt formulation (KSD within 1e-10 relative)   */
#define GFK_F32_FAST 2   /* float32 pair kernel, float64 running objective (1e-5 relative) */

/* Device-resident description of one thinning / KSD problem.  `x` and `s` are the
 * STANDARDISED sample and score (gfk_standardize output), `lw` the log-weights
 * log q - log p (NULL selects the score-based kernel k_P, non-NULL the gradient-free
 * kernel k_PQ = exp(lw_i + lw_j) * k_Q).  For GFK_F32_FAST x/s/lw are float, else double.
 * Scalar preconditioner Gamma^-1 = lambda * I (the only kind the reference uses);
 * `xl` != NULL selects a dense symmetric Gamma^-1 = L: xl = x @ L (n, d) and `trace` = tr(L). */
typedef struct gfk_problem {
    const void* x;
    const void* s;
    const void* lw;   /* log-weights log q - log p (n), or NULL                             */
    const void* w;    /* optional pre-exponentiated weights; NULL: exp(lw_i + lw_j) per pair with the device exp.
                         Fast modes: exp(lw) (n).  GFK_F64_STRICT: double-double exp(lw) as (hi, lo) pairs (2n
                         doubles, 16-byte aligned, gfk_exp_weights_dd) - the pair weight exp(fl(lw_i + lw_j)) is then
                         formed with 12 IEEE operations, correctly rounded in all but ~2^-30 of the cases and
                         identical on any machine; requires every |lw| <= 350 (gfk_log_weights_range)          */
    const void* xl;   /* optional x @ L (n, d) for a dense symmetric Gamma^-1 = L, or NULL       */
    int64_t n;        /* rows in x / s / lw (the full replica)                          */
    int32_t d;        /* columns                                                        */
    int32_t mode;     /* GFK_F64_STRICT | GFK_F64_FAST | GFK_F32_FAST                   */
    double lambda;    /* Gamma^-1 = lambda * I                                          */
    double lambda_sq; /* (Gamma^-1 Gamma^-1)[0][0] as numpy computes it (np.dot(linv, linv)) */
    double trace;     /* np.trace(linv), computed on the host                           */
    double c;         /* IMQ constant c (reference: 1.0).  beta is fixed at -1/2.       */
} gfk_problem;

int gfk_version(void);
/* kernel launches issued by this library so far (process-wide); reset != 0 zeroes the counter */
long long gfk_launch_count(int reset);
const char* gfk_last_error(void);
/* compute capability (major*10+minor), SM count and name of the current device; host outputs */
int gfk_device_info(int* cc, int* sm_count, char* name, int name_len);

/* ---- a3: stein_thinning.thinning._validate_and_standardize ------------------------
 * loc = mean(x), scl = mean(|x - loc|) per column; x_out = x / scl, g_out = g * scl.
 * With standardize = 0 the data are copied unchanged (scl = 1).  Also counts
 * non-finite entries of x and g.  Reduction order: fixed 2048-row chunks, then a fixed
 * tree over the chunk partials - deterministic and independent of the GPU count, but
 * not numpy's strictly sequential order ("standardisation").
 * SYNCHRONOUS: scl_host (d doubles, host) is valid on return.  Returns GFK_ENUMERIC if
 * any input is non-finite or (standardize and min(scl) == 0).
 * ws: device scratch of gfk_standardize_ws_bytes(n, d) bytes. */
size_t gfk_standardize_ws_bytes(int64_t n, int32_t d);
int gfk_standardize(const double* x, const double* g, int64_t n, int32_t d, int standardize,
                    double* x_out, double* g_out, double* scl_host, void* ws, void* stream);

/* ---- a2: stein_thinning.kernel.make_precon, 'med' / 'sclmed' -----------------------
 * Median of the pairwise Euclidean distances (scipy pdist order of operations) of the
 * rows x[linspace(0, n-1, cap, dtype=int)] if n > cap else all rows.  Writes the two
 * middle order statistics to mid_host[2] (host; equal when the count is odd), so the
 * caller finishes np.median / **2 / inv on the host exactly as numpy does.
 * SYNCHRONOUS.  ws: device scratch of gfk_median_ws_bytes(n, cap) bytes. */
size_t gfk_median_ws_bytes(int64_t n, int32_t cap);
int gfk_median_pdist(const double* x, int64_t n, int32_t d, int32_t cap, double* mid_host,
                     void* ws, void* stream);

/* ---- a2: stein_thinning.kernel.make_precon, 'smpcov' (never called by the reference; parity unpinned) ----------
 * gfk_covariance: cov_host (d x d doubles, host) <- np.cov(x, rowvar=False) (ddof = 1) of the device sample, summed
 * in fixed 2048-row chunks.  SYNCHRONOUS.  ws: gfk_covariance_ws_bytes(n, d) bytes of device scratch.
 * gfk_rows_times_matrix: out (n, d) <- x (n, d) @ L (d, d row-major, device), the `xl` array of gfk_problem for a
 * dense symmetric preconditioner L.  Asynchronous.  1 <= d <= 160. */
size_t gfk_covariance_ws_bytes(int64_t n, int32_t d);
int gfk_covariance(const double* x, int64_t n, int32_t d, double* cov_host, void* ws, void* stream);
int gfk_rows_times_matrix(const double* x, int64_t n, int32_t d, const double* L, double* out, void* stream);

/* float64 -> float32 copy of a device array (count elements), for GFK_F32_FAST */
int gfk_to_f32(const double* src, float* dst, int64_t count, void* stream);

/* ---- a4: the log-weights of stein_thinning.thinning._make_stein_gf_integrand ----------
 * lw = log_q - log_p; with use_cap: lw = max(lw - max(lw), -range_cap) (PARITY UNPINNED, see
 * oracle/stein_thinning_oracle/thinning.py).  spread_host <- max(lw) - min(lw) of the raw
 * difference (the caller warns when it exceeds WEIGHT_SCALE_THRESHOLD = 10, as the reference
 * does: Gaussian_mixture.ipynb#c62).  GFK_ENUMERIC if a log-density is NaN.  SYNCHRONOUS.
 * gfk_exp_weights: w = exp(lw) for the fast modes (is_f32: float in / float out). */
size_t gfk_log_weights_ws_bytes(void);
int gfk_log_weights(const double* log_p, const double* log_q, int64_t n, int use_cap, double range_cap,
                    double* lw, double* spread_host, void* ws, void* stream);
int gfk_exp_weights(const void* lw, void* w, int64_t count, int is_f32, void* stream);
/* Strict mode: ew[2i], ew[2i+1] = (hi, lo) with hi + lo = exp(lw_i) to ~2^-96 (double-double arithmetic, IEEE
 * operations only).  ew: 2 * count doubles, 16-byte aligned.  Entries with |lw| > 350 are NaN markers: pass the
 * array as gfk_problem.w only when gfk_log_weights_range shows every |lw| <= 350.  Asynchronous. */
int gfk_exp_weights_dd(const double* lw, double* ew, int64_t count, void* stream);
/* min and max of a log-weight vector (host out, 2 doubles).  SYNCHRONOUS.  The KSD row sums may use the
 * pre-exponentiated weights (one multiplication per pair instead of one exp) whenever every |lw| is far from the
 * overflow / underflow range of exp; the caller decides with this.  ws: gfk_log_weights_ws_bytes(). */
int gfk_log_weights_range(const double* lw, int64_t n, double* min_max_host, void* ws, void* stream);
/* the same two numbers for the lw written by this thread's most recent gfk_log_weights call, from its own host-side
 * statistics: no launch, no synchronisation */
int gfk_log_weights_last_range(double* min_max_host);

/* ---- next: the Gaussian proxy of gaussian_thin, code/src/thinning.py:15-16 ----------------
 * log_q[i] = -0.5 (cst + (x_i - mean)' P (x_i - mean)) with cst = d log(2 pi) + log det(cov) (what scipy's
 * mvn.logpdf evaluates), grad_q[i] = -P (x_i - mean), P = cov^-1.  x (n, d) RAW sample, mean (d) and prec (d, d
 * row-major) on the device.  Asynchronous.  [added in round 1 without GPU access: verified on the host only] */
int gfk_gaussian_proxy(const double* x, int64_t n, int32_t d, const double* mean, const double* prec, double cst,
                       double* log_q, double* grad_q, void* stream);

/* Multivariate Student-t proxy (code/notebooks/lotka_volterra/Gradient_free_Student_t.ipynb#c26-#c31,#c40):
 * maha = (x - mean)' P (x - mean);  log_q = cst - (df + d)/2 log(1 + maha/df)  with
 * cst = lgamma((df+d)/2) - lgamma(df/2) - d/2 log(df pi) - 1/2 log det(shape)  (scipy's multivariate_t.logpdf),
 * grad_q = -(df + d)/df / (1 + maha/df) P (x - mean),  P = shape^-1.  Same layout as gfk_gaussian_proxy.
 * Asynchronous.  [round 1: executed on the CPU emulator of the test suite only, not yet on a GPU] */
int gfk_student_t_proxy(const double* x, int64_t n, int32_t d, const double* mean, const double* prec, double df,
                        double cst, double* log_q, double* grad_q, void* stream);

/* ---- next: (weighted) Gaussian KDE proxy -------------------------------------------------------
 * Reference: jax/scipy gaussian_kde(sample.T, bw_method='silverman'[, weights=w]).logpdf and its gradient
 * (code/notebooks/gaussian_mixture/Gaussian_mixture.ipynb#c42-#c54; lotka_volterra/KDE.ipynb#c7-#c11).
 *   q(p) = sum_j w_j N(p; data_j, H);   log_q[i] = log q(points_i),   grad_q[i] = grad log q(points_i)
 * data (n, d), points (m, d), weights (n, normalised to sum 1) or NULL for 1/n, all on the device;
 * whiten (d, d row-major, device): any W with W' W = H^-1 (e.g. the inverse Cholesky factor of H);
 * log_norm = -0.5 log det(2 pi H).  1 <= d <= 8.  ws: gfk_kde_ws_bytes(n, m, d) bytes of device scratch.
 * Asynchronous; O(n m) pairs, each d subtractions + 2d FMAs + one exp.
 * [round 1: executed on the CPU emulator of the test suite only, not yet on a GPU] */
size_t gfk_kde_ws_bytes(int64_t n, int64_t m, int32_t d);
int gfk_kde_proxy(const double* data, const double* weights, int64_t n, const double* points, int64_t m, int32_t d,
                  const double* whiten, double log_norm, double* log_q, double* grad_q, void* ws, void* stream);

/* ---- next: energy distance of every prefix of a thinned sample -----------------------------------
 * Reference: np.sqrt(dcor.energy_distance(validation, sample[idx[:t]])) for a series of t
 * (code/notebooks/lotka_volterra/Comparison.ipynb#c19-#c23; definition report/report.tex:437-455).
 *   e2[t] = 2 mean_{a, u<=t} |x_a - y_u| - mean_{a,b} |x_a - x_b| - mean_{u,v<=t} |y_u - y_v|,  t = 0 .. m-1
 * (the SQUARED energy distance of x against the first t+1 rows of y; V-statistic means as dcor's default).
 * x (n, d) validation sample, y (m, d) selected points in selection order, e2 (m), all on the device.  1 <= d <= 8.
 * ws: gfk_energy_ws_bytes(n, m) bytes of device scratch.  Asynchronous.
 * [round 1: executed on the CPU emulator of the test suite only, not yet on a GPU] */
size_t gfk_energy_ws_bytes(int64_t n, int64_t m);
int gfk_energy_distance_cumulative(const double* x, int64_t n, const double* y, int64_t m, int32_t d, double* e2,
                                   void* ws, void* stream);

/* ---- a5 + a4: stein_thinning.thinning._greedy_search over the fused integrand -------
 * Greedy selection of m indices (with replacement, first-index tie-break = np.argmin):
 *   A_i = k(i, i);  idx[0] = argmin A;  repeat  A_i += 2 k(i, idx[t-1]);  idx[t] = argmin A.
 * This rank owns candidate rows [row_begin, row_end) of the replica described by `p`
 * (row_begin = 0, row_end = n on one GPU).  With a communicator (gfk_comm_init) the ranks
 * exchange their (value, index) pair every step; all ranks return the same idx.
 *   objective : device, (row_end - row_begin) doubles, the running objective A (out)
 *   idx_out   : device, m uint32 (global row indices)
 *   ws        : device scratch of gfk_thin_ws_bytes() bytes
 * Asynchronous on `stream`. */
size_t gfk_thin_ws_bytes(void);
/* SYNCHRONOUS: after gfk_thin on the same stream and workspace - GFK_ECUDA if the job aborted on the device: a wait
 * for another block or a peer rank ran past its deadline (GFK_SPIN_MS, default 200 ms; first cross-rank exchange
 * GFK_SPIN_FIRST_MS, default 5000 ms), a peer rank aborted and said so, or an exchanged row index was out of range.
 * Every rank of the job gets the error within one deadline; idx_out is then invalid.  An empty shard
 * (row_begin == row_end, more ranks than rows) is legal on a multi-rank communicator. */
int gfk_thin_check(const void* ws, void* stream);
int gfk_thin(const gfk_problem* p, int64_t row_begin, int64_t row_end, int32_t m,
             double* objective, uint32_t* idx_out, void* ws, void* stream);

/* ---- next: several independent chains in one launch ------------------------------------------------
 * Reference: thin() mapped over the five chains of an experiment with joblib
 * (code/notebooks/lotka_volterra/Stein_thinning.ipynb#c15-#c17; code/src/utils/parallel.py:48-52).
 * gfk_thin_batched runs gfk_thin's single-launch greedy search for n_chains problems in ONE cooperative launch: the grid
 * is split into one group of CTAs per chain, every group with its own workspace and exchange slots, so a set of
 * small chains pays the per-step latency floor once instead of n_chains times.  Single rank (chains are dealt over
 * ranks by the caller, no data-path collective).
 *   problems   : HOST array of n_chains structs (device pointers inside), all with the same d <= 4, mode and kernel kind
 *                (lw / w both present or both absent) and a scalar preconditioner
 *   objectives : HOST array of n_chains device pointers, problems[c].n doubles each (out)
 *   idx_outs   : HOST array of n_chains device pointers, m uint32 each (out)
 *   ws         : device scratch of gfk_thin_batched_ws_bytes(n_chains) bytes, 256-byte aligned
 * Returns GFK_EUNSUPPORTED (nothing launched) when the set is not eligible - d > 4, mixed kinds, a chain too long for
 * the shared memory of its CTA group, more chains than SMs: the caller then loops over gfk_thin.  Asynchronous;
 * gfk_thin_batched_check is the synchronous counterpart of gfk_thin_check for the whole set. */
size_t gfk_thin_batched_ws_bytes(int32_t n_chains);
int gfk_thin_batched(const gfk_problem* problems, int32_t n_chains, int32_t m, double* const* objectives,
                     uint32_t* const* idx_outs, void* ws, void* stream);
int gfk_thin_batched_check(const void* ws, int32_t n_chains, void* stream);

/* ---- a7: stein_thinning.stein.ksd / kmat ----------------------------------------------
 * gfk_ksd_rowsums: for the rows sel[0..m) of the replica (sel == NULL: rows 0..m-1),
 *   low[t] = sum_{b<t} k(sel[t], sel[b]),  diag[t] = k(sel[t], sel[t]).
 * Only the row tiles t with (t / 64) % n_parts == part are computed (others are written
 * as 0), so ranks can split the triangle and all-reduce the two vectors.
 * gfk_ksd_finish: ks[t] = sqrt(sum_{u<=t} (2 low[u] + diag[u])) / (t + 1)   (cumulative KSD).
 * gfk_kmat: dense symmetric m x m Gram matrix (row-major, device). */
int gfk_ksd_rowsums(const gfk_problem* p, const uint32_t* sel, int64_t m, int32_t part,
                    int32_t n_parts, double* low, double* diag, void* stream);
int gfk_ksd_finish(const double* low, const double* diag, int64_t m, double* ks, void* stream);
/* The same row sums (low, diag as gfk_ksd_rowsums, except that the tiles dealt over the parts are 128 rows) on the tensor cores - BASELINE.json config 5, "for
 * large d the cross terms are GEMM-shaped": the four Gram tiles X X', X S', S X', S S' (report/report.tex:853-868) by
 * tcgen05.mma kind::tf32 with 3xTF32 operand splits (22 significant bits), operands fed by TMA, accumulators in
 * tensor memory, the pair epilogue of the tile kernel on top.  float32 mode only (1e-5 relative on KSD values), scalar
 * preconditioner, d <= 512 (row operands resident in shared memory up to d = 64, streamed beyond): gfk_ksd_mma_supported(p) != 0; otherwise GFK_EUNSUPPORTED and nothing is launched.
 * ws: gfk_ksd_mma_ws_bytes(m, d) bytes of device scratch, 256-byte aligned.  Asynchronous.
 * [round 2: executed on the CPU emulator of the test suite only, not yet on a GPU; opt-in from Python, GFK_KSD_MMA=1] */
int gfk_ksd_mma_supported(const gfk_problem* p);
size_t gfk_ksd_mma_ws_bytes(int64_t m, int32_t d);
int gfk_ksd_rowsums_mma(const gfk_problem* p, const uint32_t* sel, int64_t m, int32_t part, int32_t n_parts, double* low,
                        double* diag, void* ws, void* stream);
/* SYNCHRONOUS, after gfk_ksd_rowsums_mma on the same stream and workspace: GFK_ECUDA if a wait inside the kernel's
 * producer / MMA / epilogue pipeline ran past its 2 s deadline (the kernel then terminated with invalid sums) */
int gfk_ksd_mma_check(const void* ws, void* stream);
int gfk_kmat(const gfk_problem* p, const uint32_t* sel, int64_t m, double* out, void* stream);
/* element-wise paired kernel, the integrand protocol: out[t] = k(ia[t], ib[t]), count pairs */
int gfk_pairs(const gfk_problem* p, const int64_t* ia, const int64_t* ib, int64_t count,
              double* out, void* stream);

/* ---- multi-GPU: one process per GPU, NCCL over NVLink -----------------------------------
 * gfk_comm_unique_id: host, fills 128 bytes (ncclUniqueId) - call on rank 0, broadcast out of band.
 * gfk_comm_init: collective; afterwards gfk_thin exchanges candidates and gfk_allreduce_sum works.
 * nccl_path: path of libnccl.so.2 to dlopen, or NULL to use the one already loaded / default. */
int gfk_comm_unique_id(void* id128_host, const char* nccl_path);
int gfk_comm_init(int n_ranks, int rank, const void* id128_host, const char* nccl_path);
/* Optional: peer-memory mailboxes for the per-step exchange (CUDA IPC, NVLink P2P).  Each rank calls
 * gfk_p2p_local_handle (64-byte cudaIpcMemHandle_t out), the handles are all-gathered out of band into
 * handles_host[n_ranks * 64], then every rank calls gfk_p2p_open.  Afterwards gfk_thin exchanges its
 * (value, index) pairs with flag-in-data stores into the peers' mailboxes from inside the thinning kernels
 * (and the single-launch persistent kernel becomes available on every rank).  OPT-IN PER JOB: gfk_thin uses the
 * mailboxes only when they are open AND the environment of the calling process has GFK_P2P=1 (the same on every
 * rank); the default multi-rank exchange is a 16-byte ncclAllGather per greedy step. */
int gfk_p2p_local_handle(void* handle64_host);
int gfk_p2p_open(const void* handles_host, int n_ranks, int rank);
int gfk_p2p_enabled(void);
/* collective: `iters` back-to-back mailbox exchanges with no other work; time per exchange in us (host out) */
int gfk_p2p_selftest(int iters, double* us_per_exchange_host, void* stream);
int gfk_comm_size(void);
int gfk_comm_rank(void);
int gfk_allreduce_sum(double* buf, int64_t count, void* stream);
int gfk_comm_finalize(void);

/* ---- measurement helpers (bench.py) -----------------------------------------------------
 * gfk_fp64_fma_peak: runs a register-resident DFMA (is_f32: FFMA) chain on every SM and
 * writes the achieved TFLOP/s (FMA = 2 flops) to tflops_host.  SYNCHRONOUS. */
int gfk_fma_peak(int is_f32, double* tflops_host, void* stream);
/* writes `bytes` of device scratch to evict L2 between timed iterations */
int gfk_flush_l2(void* scratch, size_t bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GFK_H_ */
