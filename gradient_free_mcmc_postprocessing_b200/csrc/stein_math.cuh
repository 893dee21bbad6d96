// stein_math.cuh - the pair kernel k(x, y) of the Stein thinning / GF-KSD path.
//
// Replaces stein_thinning.kernel.vfk0_imq (restated in the reference at
// code/notebooks/examples/JAX_Stein_Thinning.ipynb#c27 and, term by term, at
// code/notebooks/Kernel_Stein_discrepancy.ipynb#c7; derivation report/report.tex:853-868)
// with c = 1 and beta = -1/2 (report/report.tex:435), for the scalar preconditioner
// Gamma^-1 = lambda * I and, through a pre-multiplied copy xl = x @ L, a dense symmetric one.
//
//   r = x - y,  qf = c + lambda |r|^2
//   k = -3 lambda^2 |r|^2 / qf^2.5 + (tr + lambda <sx - sy, r>) / qf^1.5 + <sx, sy> / qf^0.5
//
// Two formulations, both plain scalar code so the same header also compiles for the host
// (tests/host/pair_check.cpp checks the strict path against numpy bit for bit on the CPU):
//
//   strict : numpy's operation order, no FMA contraction, correctly rounded sqrt / powers /
//            divisions.  Every operation except pow() and exp() is reproduced bit for bit;
//            pow(q, 1.5 | 2.5) is computed correctly rounded from a double-double sqrt
//            (numpy's own SVML pow is correctly rounded in ~95 % of cases, glibc's in 99.9 %).
//   fast   : FMA sums, one rsqrt, r, r^3, r^5 by multiplication.  <= a few ulp from strict.
#pragma once

#include <math.h>
#include <string.h>

#if defined(__CUDACC__)
#define GFK_HD __host__ __device__ __forceinline__
#define GFK_HD_COLD static __host__ __device__ __noinline__   // rare paths: kept out of the hot loops' instruction footprint
#else
#define GFK_HD inline
#define GFK_HD_COLD inline
#endif

namespace gfk {

struct KParams {
    double lambda;     // Gamma^-1 = lambda * I
    double lambda_sq;  // fl(lambda * lambda), numpy's np.dot(linv, linv) diagonal
    double trace;      // np.trace(linv)
    double c;          // IMQ constant (1.0)
    int d;             // runtime dimension (used when the template D is 0)
    int dense;         // 1: dense L through the pre-multiplied copy xl (fast formulation only)
};

// ---- non-contracted IEEE operations ------------------------------------------------------
#if defined(__CUDA_ARCH__)
GFK_HD double nf_mul(double a, double b) { return __dmul_rn(a, b); }
GFK_HD double nf_add(double a, double b) { return __dadd_rn(a, b); }
GFK_HD double nf_sub(double a, double b) { return __dsub_rn(a, b); }
GFK_HD double nf_sqrt(double a) { return __dsqrt_rn(a); }
GFK_HD double nf_div(double a, double b) { return __ddiv_rn(a, b); }
GFK_HD double nf_fma(double a, double b, double c) { return __fma_rn(a, b, c); }
GFK_HD double approx_rsqrt(double a) { return rsqrt(a); }
GFK_HD float approx_rsqrt(float a) { return rsqrtf(a); }
#else
// host build: compiled with -ffp-contract=off so a * b + c stays two roundings
GFK_HD double nf_mul(double a, double b) { return a * b; }
GFK_HD double nf_add(double a, double b) { return a + b; }
GFK_HD double nf_sub(double a, double b) { return a - b; }
GFK_HD double nf_sqrt(double a) { return sqrt(a); }
GFK_HD double nf_div(double a, double b) { return a / b; }
GFK_HD double nf_fma(double a, double b, double c) { return fma(a, b, c); }
#if defined(GFK_TEST_RSQRT_PERTURB)
// test hook (tests/host): emulate a reciprocal square root that is several ulp off, as a device
// approximation may be, to show the correction steps below do not depend on its last bits
GFK_HD double approx_rsqrt(double a) {
    const double r = 1.0 / sqrt(a);
    unsigned long long bits;
    memcpy(&bits, &a, sizeof(bits));
    const int k = (int)((bits * 0x9E3779B97F4A7C15ull) >> 61) - 3;   // -3 .. 4 ulp
    return r * (1.0 + k * 2.220446049250313e-16);
}
#else
GFK_HD double approx_rsqrt(double a) { return 1.0 / sqrt(a); }
#endif
GFK_HD float approx_rsqrt(float a) { return 1.0f / sqrtf(a); }
#endif

GFK_HD double fma_t(double a, double b, double c) { return nf_fma(a, b, c); }
GFK_HD float fma_t(float a, float b, float c) { return fmaf(a, b, c); }

// ---- exponent-range tests (integer pipe; keep the FP64 pipe and the control flow free of guards) ----
GFK_HD unsigned exponent_field(double v) {
#if defined(__CUDA_ARCH__)
    return ((unsigned)__double2hiint(v) >> 20) & 0x7ffu;
#else
    unsigned long long bits;
    memcpy(&bits, &v, sizeof(bits));
    return (unsigned)(bits >> 52) & 0x7ffu;
#endif
}
// |v| within [2^-span, 2^span]
GFK_HD bool exponent_within(double v, unsigned span) { return (exponent_field(v) - (0x3ffu - span)) <= 2u * span; }
// v == +-0, on the integer pipe
GFK_HD bool is_zero(double v) {
#if defined(__CUDA_ARCH__)
    return (((unsigned)__double2hiint(v) & 0x7fffffffu) | (unsigned)__double2loint(v)) == 0u;
#else
    unsigned long long bits;
    memcpy(&bits, &v, sizeof(bits));
    return (bits << 1) == 0ull;
#endif
}

// Correctly rounded quotient n / p from an approximate reciprocal rinv ~ 1/p (a few ulp): one
// Markstein correction step.  q0 = n * rinv is within a few ulp of n / p, the residual n - q0 * p is
// exact in an FMA, and q0 + resid * rinv is then accurate to ~2^-100 relative before its single final
// rounding (a quotient of doubles is never closer than ~2^-106 relative to a rounding midpoint, so the
// result equals IEEE division except in astronomically rare cases).  The caller guarantees the operands
// are far from overflow / underflow (strict_finish checks the exponents once per pair).
GFK_HD double div_markstein(double n, double p, double rinv) {
    const double q0 = nf_mul(n, rinv);
    const double resid = nf_fma(-q0, p, n);
    return nf_fma(resid, rinv, q0);
}

// q^0.5, q^1.5 and q^2.5, each correctly rounded up to ~2^-100, plus approximate reciprocals (a few
// ulp) of the three powers for div_markstein.  Requires 2^-150 <= q <= 2^150 (checked by the caller).
// r = rsqrt(q) approximately; s0 = q * r; sqrt(q) = s0 + (q - s0^2) * r / 2 to ~2^-100: the rounded sum is
// RN(sqrt q) (sqrt of a double is never closer than ~2^-105 relative to a midpoint), and the unrounded pair
// (s, sl) feeds the double-double products q * sqrt(q) and q^2 * sqrt(q), each rounded once.
GFK_HD void strict_powers_r(double q, double& p05, double& p15, double& p25, double& r05, double& r15,
                            double& r25) {
    const double r = approx_rsqrt(q);
    const double hr = nf_mul(0.5, r);
    const double s0 = nf_mul(q, r);
    const double e0 = nf_fma(-s0, s0, q);                   // exact: q - s0*s0
    const double s = nf_fma(e0, hr, s0);                    // RN(sqrt q)
    p05 = s;
    const double resid = nf_fma(-s, s, q);                  // exact: q - s*s
    const double sl = nf_mul(resid, hr);                    // sqrt(q) = s + sl (+ O(2^-105))
    // q^1.5
    const double h1 = nf_mul(q, s);
    double l1 = nf_fma(q, s, -h1);                          // exact low part of q*s
    l1 = nf_fma(q, sl, l1);
    p15 = nf_add(h1, l1);
    // q^2.5 = q * (h1 + l1): the double-double q^1.5 times q, rounded once
    const double h2 = nf_mul(q, h1);
    double l2 = nf_fma(q, h1, -h2);                         // exact low part of q*h1
    l2 = nf_fma(q, l1, l2);
    p25 = nf_add(h2, l2);
    const double r2 = nf_mul(r, r);
    r05 = r;
    r15 = nf_mul(r, r2);
    r25 = nf_mul(r15, r2);
}

// the same three powers for ANY q >= 0 (including 0, inf, nan and the far exponent ranges): IEEE sqrt
// and plain products where the double-double construction could overflow or lose its low parts
GFK_HD void strict_powers(double q, double& p05, double& p15, double& p25) {
    if (exponent_within(q, 150u)) {
        double r05, r15, r25;
        strict_powers_r(q, p05, p15, p25, r05, r15, r25);
        return;
    }
    const double s = nf_sqrt(q);
    p05 = s;
    p15 = nf_mul(q, s);
    p25 = nf_mul(nf_mul(q, q), s);
}

// numpy's pairwise summation of the d products sx_k * sy_k (np.sum over the contiguous
// axis of an F-ordered (d, n) array; numpy/_core/src/umath/loops_utils.h.src, pairwise_sum):
// plain loop for d < 8, eight strided accumulators combined as a tree for d <= 128,
// recursive halving (multiple of 8) above.  STRIDE lets the caller pass SoA / padded rows.
template <typename AX, typename AY>
GFK_HD double pairwise_dot_block(const AX& sx, const AY& sy, int lo, int n) {
    if (n < 8) {
        double res = 0.0;
        for (int i = 0; i < n; ++i) res = nf_add(res, nf_mul(sx[lo + i], sy[lo + i]));
        return res;
    }
    double r0 = nf_mul(sx[lo + 0], sy[lo + 0]), r1 = nf_mul(sx[lo + 1], sy[lo + 1]);
    double r2 = nf_mul(sx[lo + 2], sy[lo + 2]), r3 = nf_mul(sx[lo + 3], sy[lo + 3]);
    double r4 = nf_mul(sx[lo + 4], sy[lo + 4]), r5 = nf_mul(sx[lo + 5], sy[lo + 5]);
    double r6 = nf_mul(sx[lo + 6], sy[lo + 6]), r7 = nf_mul(sx[lo + 7], sy[lo + 7]);
    int i = 8;
    for (; i < n - (n % 8); i += 8) {
        r0 = nf_add(r0, nf_mul(sx[lo + i + 0], sy[lo + i + 0]));
        r1 = nf_add(r1, nf_mul(sx[lo + i + 1], sy[lo + i + 1]));
        r2 = nf_add(r2, nf_mul(sx[lo + i + 2], sy[lo + i + 2]));
        r3 = nf_add(r3, nf_mul(sx[lo + i + 3], sy[lo + i + 3]));
        r4 = nf_add(r4, nf_mul(sx[lo + i + 4], sy[lo + i + 4]));
        r5 = nf_add(r5, nf_mul(sx[lo + i + 5], sy[lo + i + 5]));
        r6 = nf_add(r6, nf_mul(sx[lo + i + 6], sy[lo + i + 6]));
        r7 = nf_add(r7, nf_mul(sx[lo + i + 7], sy[lo + i + 7]));
    }
    double res = nf_add(nf_add(nf_add(r0, r1), nf_add(r2, r3)), nf_add(nf_add(r4, r5), nf_add(r6, r7)));
    for (; i < n; ++i) res = nf_add(res, nf_mul(sx[lo + i], sy[lo + i]));
    return res;
}

template <typename AX, typename AY>
GFK_HD double pairwise_dot(const AX& sx, const AY& sy, int d) {
    if (d <= 128) return nf_add(0.0, pairwise_dot_block(sx, sy, 0, d));
    // iterative form of numpy's recursion: split [lo, lo + n) at n2 = (n / 2) rounded down to 8
    // until blocks are <= 128, summing leaves in recursion order with an explicit stack
    int lo_stack[24], n_stack[24];
    double val_stack[24];
    int state[24];
    int sp = 0;
    lo_stack[0] = 0; n_stack[0] = d; state[0] = 0;
    double ret = 0.0;
    while (sp >= 0) {
        const int n = n_stack[sp], lo = lo_stack[sp];
        if (n <= 128) { ret = pairwise_dot_block(sx, sy, lo, n); --sp; continue; }
        int n2 = n / 2; n2 -= n2 % 8;
        if (state[sp] == 0) {          // descend left
            state[sp] = 1; ++sp; lo_stack[sp] = lo; n_stack[sp] = n2; state[sp] = 0;
        } else if (state[sp] == 1) {   // left done, descend right
            val_stack[sp] = ret; state[sp] = 2; ++sp; lo_stack[sp] = lo + n2; n_stack[sp] = n - n2; state[sp] = 0;
        } else {                       // both done
            ret = nf_add(val_stack[sp], ret); --sp;
        }
    }
    return nf_add(0.0, ret);
}

// out-of-line: IEEE sqrt / division, valid for any operands
GFK_HD_COLD double strict_finish_rare(double qf, double n1, double n2, double n3) {
    double p05, p15, p25;
    strict_powers(qf, p05, p15, p25);
    return nf_add(nf_add(nf_div(n1, p25), nf_div(n2, p15)), nf_div(n3, p05));
}

// From the four sums to the kernel value, numpy's order:
//   s1 = sum (lambda r) r, s2 = sum (lambda^2 r) r, s3 = sum (lambda (sx - sy)) r, s4 = sum sx sy
GFK_HD double strict_finish(double s1, double s2, double s3, double s4, const KParams& kp) {
    const double qf = nf_add(kp.c, s1);
    const double n1 = nf_mul(-3.0, s2);
    const double n2 = nf_add(kp.trace, s3);
    const double n3 = s4;
    // one exponent test per pair instead of guards around every operation
    const bool regular = exponent_within(qf, 150u) && (exponent_within(n1, 400u) || is_zero(n1)) &&
                         (exponent_within(n2, 400u) || is_zero(n2)) && (exponent_within(n3, 400u) || is_zero(n3));
    if (regular) {
        double p05, p15, p25, r05, r15, r25;
        strict_powers_r(qf, p05, p15, p25, r05, r15, r25);
        const double t1 = div_markstein(n1, p25, r25);
        const double t2 = div_markstein(n2, p15, r15);
        const double t3 = div_markstein(n3, p05, r05);
        return nf_add(nf_add(t1, t2), t3);
    }
    return strict_finish_rare(qf, n1, n2, n3);              // extreme magnitudes, zeros, inf, nan
}

// ---- strict pair kernel (float64) ------------------------------------------------------------
// x, y, sx, sy: anything indexable by [k] (register arrays, strided shared-memory views).
// Returns k(x, y) WITHOUT the gradient-free weight.
template <int D, typename AX, typename AY, typename ASX, typename ASY>
GFK_HD double pair_strict(const AX& x, const AY& y, const ASX& sx, const ASY& sy, const KParams& kp) {
    const int d = D > 0 ? D : kp.d;
    double s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
    for (int k = 0; k < d; ++k) {
        const double r = nf_sub(x[k], y[k]);
        const double a = nf_mul(nf_mul(kp.lambda, r), r);
        const double b = nf_mul(nf_mul(kp.lambda_sq, r), r);
        const double cc = nf_mul(nf_mul(kp.lambda, nf_sub(sx[k], sy[k])), r);
        if (k == 0) { s1 = a; s2 = b; s3 = cc; }
        else { s1 = nf_add(s1, a); s2 = nf_add(s2, b); s3 = nf_add(s3, cc); }
    }
    double s4;
    if (d < 8) {
        s4 = 0.0;
#pragma unroll
        for (int k = 0; k < d; ++k) {
            const double p = nf_mul(sx[k], sy[k]);
            s4 = (k == 0) ? p : nf_add(s4, p);
        }
    } else {
        s4 = pairwise_dot(sx, sy, d);
    }
    return strict_finish(s1, s2, s3, s4, kp);
}

// From the dot products to the kernel value, FMA formulation:
//   scalar: rr = <r, r>, su = <sx - sy, r>;  dense: rr = <L r, r>, uu = <L r, L r>, su = <sx - sy, L r>;  ss = <sx, sy>
template <typename T, bool DENSE>
GFK_HD T fast_finish(T rr, T uu, T su, T ss, const KParams& kp) {
    const T lam = T(kp.lambda);
    const T qf = DENSE ? (T(kp.c) + rr) : fma_t(lam, rr, T(kp.c));
    const T n1 = DENSE ? (T(-3) * uu) : (T(-3) * T(kp.lambda_sq) * rr);
    const T n2 = DENSE ? (T(kp.trace) + su) : fma_t(lam, su, T(kp.trace));
    const T i1 = approx_rsqrt(qf);
    const T i2 = i1 * i1;
    const T i3 = i1 * i2;
    const T i5 = i3 * i2;
    return fma_t(n1, i5, fma_t(n2, i3, ss * i1));
}

// ---- fast pair kernel (float64 or float32) ------------------------------------------------------
// xl == x for the scalar preconditioner; for a dense symmetric L the caller passes the
// pre-multiplied rows xl = x @ L, yl = y @ L and kp.lambda = kp.lambda_sq = 1:
//   u = xl - yl = L r;  qf = c + <u, r>;  |L r|^2 = <u, u>;  <L (sx - sy), r> = <sx - sy, u>.
template <typename T, int D, bool DENSE, typename AX, typename AY, typename ASX, typename ASY, typename AXL,
          typename AYL>
GFK_HD T pair_fast(const AX& x, const AY& y, const ASX& sx, const ASY& sy, const AXL& xl, const AYL& yl,
                   const KParams& kp) {
    const int d = D > 0 ? D : kp.d;
    T rr = T(0), uu = T(0), su = T(0), ss = T(0);
#pragma unroll
    for (int k = 0; k < d; ++k) {
        const T r = T(x[k]) - T(y[k]);
        const T ds = T(sx[k]) - T(sy[k]);
        if (DENSE) {
            const T u = T(xl[k]) - T(yl[k]);
            rr = fma_t(u, r, rr);
            uu = fma_t(u, u, uu);
            su = fma_t(ds, u, su);
        } else {
            rr = fma_t(r, r, rr);
            su = fma_t(ds, r, su);
        }
        ss = fma_t(T(sx[k]), T(sy[k]), ss);
    }
    return fast_finish<T, DENSE>(rr, uu, su, ss, kp);
}

// ---- gradient-free weights in the strict mode: exp(lw_i + lw_j) without a device exp ---------------------------
// numpy evaluates exp(fl(lw_i + lw_j)) per pair.  The device's exp() is accurate to 1 ulp, not correctly rounded,
// costs ~30 FP64 instructions plus as many constant moves per pair, and differs from the host's.  Instead every row
// gets E_i = exp(lw_i) ONCE per job as a double-double (exp_dd: argument reduction, Taylor polynomial and eight
// squarings in double-double arithmetic - IEEE operations only, so the result is the same on any machine, accurate
// to ~2^-96), and a pair costs 12 IEEE operations (exp_sum_dd):
//   a + b = s + err exactly (TwoSum),  exp(s) = E_a E_b exp(-err) = (E_a E_b)(1 - err) up to err^2 ~ 2^-86,
// i.e. exp(fl(a + b)) correctly rounded in all but ~2^-30 of the cases.  Valid while |a|, |b| <= kExpDdLimit (no
// overflow, no subnormal parts); outside that range the caller keeps the device exp.
constexpr double kExpDdLimit = 350.0;

GFK_HD void two_sum(double a, double b, double& s, double& e) {
    s = nf_add(a, b);
    const double bb = nf_sub(s, a);
    e = nf_add(nf_sub(a, nf_sub(s, bb)), nf_sub(b, bb));
}
GFK_HD void quick_two_sum(double a, double b, double& s, double& e) {      // |a| >= |b|
    s = nf_add(a, b);
    e = nf_sub(b, nf_sub(s, a));
}
GFK_HD void dd_mul(double ah, double al, double bh, double bl, double& ch, double& cl) {
    const double p = nf_mul(ah, bh);
    double e = nf_fma(ah, bh, -p);
    e = nf_fma(ah, bl, e);
    e = nf_fma(al, bh, e);
    quick_two_sum(p, e, ch, cl);
}
GFK_HD void dd_add(double ah, double al, double bh, double bl, double& ch, double& cl) {
    double s, e;
    two_sum(ah, bh, s, e);
    e = nf_add(e, nf_add(al, bl));
    quick_two_sum(s, e, ch, cl);
}

// (hi, lo) with hi + lo = exp(l) to ~2^-96 relative, |l| <= kExpDdLimit
GFK_HD void exp_dd(double l, double& hi, double& lo) {
    const double LN2_H = 0.6931471805599453, LN2_L = 2.3190468138462996e-17, LOG2E = 1.4426950408889634;
    // 1 / n! as double-doubles, n = 0 .. 11
    const double CH[12] = {1.0, 1.0, 0.5, 0.16666666666666666, 0.041666666666666664, 0.008333333333333333,
                           0.001388888888888889, 0.0001984126984126984, 2.48015873015873e-05,
                           2.7557319223985893e-06, 2.755731922398589e-07, 2.505210838544172e-08};
    const double CL[12] = {0.0, 0.0, 0.0, 9.25185853854297e-18, 2.3129646346357427e-18, 1.1564823173178714e-19,
                           -5.300543954373577e-20, 1.7209558293420705e-22, 2.1511947866775882e-23,
                           -1.858393274046472e-22, 2.3767714622250297e-23, -1.448814070935912e-24};
    const double kf = rint(nf_mul(l, LOG2E));
    // r = l - k ln2 as a double-double (k ln2_h is formed exactly, |k| < 2^10)
    const double ph = nf_mul(kf, LN2_H);
    double pl = nf_fma(kf, LN2_H, -ph);
    pl = nf_fma(kf, LN2_L, pl);
    double rh, rl;
    two_sum(l, -ph, rh, rl);
    rl = nf_sub(rl, pl);
    quick_two_sum(rh, rl, rh, rl);
    rh = nf_mul(rh, 0.00390625);                                          // r / 256: |r'| <= 1.36e-3
    rl = nf_mul(rl, 0.00390625);
    double sh = CH[11], sl = CL[11];
    for (int n = 10; n >= 0; --n) {
        dd_mul(sh, sl, rh, rl, sh, sl);
        dd_add(sh, sl, CH[n], CL[n], sh, sl);
    }
    for (int j = 0; j < 8; ++j) dd_mul(sh, sl, sh, sl, sh, sl);           // exp(r) = exp(r / 256)^256
    const int k = (int)kf;
    hi = ldexp(sh, k);
    lo = ldexp(sl, k);
}

// exp(fl(a + b)) from E_a = (eah, eal) ~ exp(a) and E_b = (ebh, ebl) ~ exp(b)
GFK_HD double exp_sum_dd(double a, double b, double eah, double eal, double ebh, double ebl) {
    double s, err;
    two_sum(a, b, s, err);                                               // numpy's exponent is s; a + b = s + err
    const double p = nf_mul(eah, ebh);
    double pl = nf_fma(eah, ebh, -p);
    pl = nf_fma(eah, ebl, pl);
    pl = nf_fma(eal, ebh, pl);
    return nf_add(p, nf_fma(-p, err, pl));
}

// ---- (value, index) with numpy.argmin semantics ---------------------------------------------------
// smaller value wins; NaN beats every number (np.argmin returns the first NaN); ties -> lower index
struct Cand {
    double v;
    unsigned long long i;
};

// the same decision for a scan in INCREASING index order (a tie keeps the earlier row, the first NaN sticks):
// `none` = nothing taken yet.  Two floating-point compares instead of five.
GFK_HD bool cand_better_in_order(double v, double best, bool none) {
    return none || ((best == best) && !(v >= best));
}

GFK_HD bool cand_better(double va, unsigned long long ia, double vb, unsigned long long ib) {
    const bool na = (va != va), nb = (vb != vb);
    if (na || nb) return (na && !nb) || (na && nb && ia < ib);
    return (va < vb) || (va == vb && ia < ib);
}

}  // namespace gfk
