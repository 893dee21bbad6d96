// Host build of the pair-kernel header (gfk::pair_strict / pair_fast) so the CPU test suite
// can compare the exact arithmetic the CUDA kernels run against numpy, bit for bit, without a
// GPU.  Test infrastructure only: compiled by tests/test_host_math.py with
//   g++ -O2 -ffp-contract=off -shared -fPIC
#include <stdint.h>
#include "stein_math.cuh"

struct RowView {
    const double* p;
    double operator[](int k) const { return p[k]; }
};

extern "C" {

// out[i] = k(x_i, y_(i or 0)); y_rows == 1 broadcasts row 0 of y / sy
void host_pair_strict(const double* x, const double* y, const double* sx, const double* sy, int64_t n,
                      int64_t y_rows, int d, double lambda, double lambda_sq, double trace, double c,
                      double* out) {
    gfk::KParams kp{lambda, lambda_sq, trace, c, d, 0};
    for (int64_t i = 0; i < n; ++i) {
        const int64_t j = (y_rows == 1) ? 0 : i;
        RowView xv{x + i * d}, yv{y + j * d}, sxv{sx + i * d}, syv{sy + j * d};
        out[i] = gfk::pair_strict<0>(xv, yv, sxv, syv, kp);
    }
}

void host_pair_fast(const double* x, const double* y, const double* sx, const double* sy, int64_t n,
                    int64_t y_rows, int d, double lambda, double lambda_sq, double trace, double c,
                    double* out) {
    gfk::KParams kp{lambda, lambda_sq, trace, c, d, 0};
    for (int64_t i = 0; i < n; ++i) {
        const int64_t j = (y_rows == 1) ? 0 : i;
        RowView xv{x + i * d}, yv{y + j * d}, sxv{sx + i * d}, syv{sy + j * d};
        out[i] = gfk::pair_fast<double, 0, false>(xv, yv, sxv, syv, xv, yv, kp);
    }
}

void host_strict_powers(const double* q, int64_t n, double* p05, double* p15, double* p25) {
    for (int64_t i = 0; i < n; ++i) gfk::strict_powers(q[i], p05[i], p15[i], p25[i]);
}

void host_exp_dd(const double* l, int64_t n, double* hi, double* lo) {
    for (int64_t i = 0; i < n; ++i) gfk::exp_dd(l[i], hi[i], lo[i]);
}

// out[i] = exp(fl(a_i + b_i)) through the double-double route the strict kernels take
void host_exp_sum_dd(const double* a, const double* b, int64_t n, double* out) {
    for (int64_t i = 0; i < n; ++i) {
        double ah, al, bh, bl;
        gfk::exp_dd(a[i], ah, al);
        gfk::exp_dd(b[i], bh, bl);
        out[i] = gfk::exp_sum_dd(a[i], b[i], ah, al, bh, bl);
    }
}

void host_pairwise_dot(const double* a, const double* b, int64_t n, int d, double* out) {
    for (int64_t i = 0; i < n; ++i) {
        RowView av{a + i * d}, bv{b + i * d};
        out[i] = gfk::pairwise_dot(av, bv, d);
    }
}
}
