// TEST INFRASTRUCTURE ONLY — CPU oracle, "exact" mode of the MonoSLAM EKF stacked update (EKF.cpp:977-1125).
//
// The reference forms S^-1 explicitly (Eigen .inverse(), EKF.cpp:1019) and P - (K S) K^T in plain double; at n = 3013 / 2m = 2000 that
// chain is itself ~1e-8 away from exact arithmetic (cond(S) * eps grows with the point count), which is more than the parity budget.
// Like the BA oracle's `exact` mode (DESIGN.md "Parity budget") this file evaluates the SAME update in long double through the
// textbook form   S = H P H^T + R = L L^T,  Y = L^-1 (H P),  x += Y^T L^-1 (z - h),  P -= Y^T Y
// followed by the reference's quaternion normalisation (:1652-1711, as the congruence J P J^T of its 4x4 block), the symmetrisation
// (:1120-1121) and EnsureNonnegativeStateVariance (:1739-1750).  It is the parity TARGET at the sizes BASELINE.json names; the
// faithful restatement (srk_oracle_ekf.hpp) stays the one that is timed and whose own distance from this evaluation is reported.
// Threads (srk_oracle_parallel.hpp) work on independent columns / rows only (no reduction across threads): the result does not depend on
// the thread count.
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <vector>
#include "srk_oracle_geom.hpp"
#include "srk_oracle_parallel.hpp"

namespace srk_oracle {

// H in the sparse form of the C ABI: Hcam [2m x 13], Hpt [2m x s] row-major per observation row, pt_off[m].  P col-major [n x n] in/out.
inline bool EkfStackedUpdateExact(int64_t n, int64_t m, double* P, double* x, const double* Hcam, const double* Hpt, const int64_t* pt_off, int s,
                                  const double* z, const double* hpred, double meas_var, bool fix_symmetry) {
    typedef long double W;
    const int64_t m2 = 2 * m;
    // PHt row-major by STATE index: A[i][r] = sum_c P(i, c) H(r, c)     (13 + s non-zeros per row of H)
    std::vector<W> A((size_t)n * m2);
    ParallelFor(0, n, 16, [&](int64_t i) {
        for (int64_t r = 0; r < m2; ++r) {
            W acc = 0;
            for (int c = 0; c < 13; ++c) acc += (W)P[(size_t)c * n + i] * (W)Hcam[r * 13 + c];
            const int64_t off = pt_off[r / 2];
            for (int c = 0; c < s; ++c) acc += (W)P[(size_t)(off + c) * n + i] * (W)Hpt[r * s + c];
            A[(size_t)i * m2 + r] = acc;
        }
    });
    // S = H (P H^T) + R, lower triangle, row-major
    std::vector<W> L((size_t)m2 * m2, (W)0);
    ParallelFor(0, m2, 16, [&](int64_t r) {
        for (int64_t c2 = 0; c2 <= r; ++c2) {
            W acc = 0;
            for (int c = 0; c < 13; ++c) acc += (W)Hcam[r * 13 + c] * A[(size_t)c * m2 + c2];
            const int64_t off = pt_off[r / 2];
            for (int c = 0; c < s; ++c) acc += (W)Hpt[r * s + c] * A[(size_t)(off + c) * m2 + c2];
            if (r == c2) acc += (W)meas_var;
            L[(size_t)r * m2 + c2] = acc;
        }
    });
    // Cholesky, left-looking by columns, rows of L contiguous
    for (int64_t j = 0; j < m2; ++j) {
        W d = L[(size_t)j * m2 + j];
        const W* lj = &L[(size_t)j * m2];
        for (int64_t k = 0; k < j; ++k) d -= lj[k] * lj[k];
        if (!(d > 0)) return false;
        d = std::sqrt(d);
        L[(size_t)j * m2 + j] = d;
        ParallelFor(j + 1, m2, 64, [&](int64_t i) {
            W* li = &L[(size_t)i * m2];
            W v = li[j];
            for (int64_t k = 0; k < j; ++k) v -= li[k] * lj[k];
            li[j] = v / d;
        });
    }
    // Y^T rows: A[i][:] <- L^-1 A[i][:]   (forward substitution per state row)
    ParallelFor(0, n, 8, [&](int64_t i) {
        W* a = &A[(size_t)i * m2];
        for (int64_t r = 0; r < m2; ++r) {
            const W* lr = &L[(size_t)r * m2];
            W v = a[r];
            for (int64_t k = 0; k < r; ++k) v -= lr[k] * a[k];
            a[r] = v / lr[r];
        }
    });
    std::vector<W> w(m2);
    for (int64_t r = 0; r < m2; ++r) {
        const W* lr = &L[(size_t)r * m2];
        W v = (W)z[r] - (W)hpred[r];
        for (int64_t k = 0; k < r; ++k) v -= lr[k] * w[k];
        w[r] = v / lr[r];
    }
    std::vector<W> xs(n);
    ParallelFor(0, n, 64, [&](int64_t i) {
        const W* a = &A[(size_t)i * m2];
        W acc = 0;
        for (int64_t r = 0; r < m2; ++r) acc += a[r] * w[r];
        xs[i] = (W)x[i] + acc;
    });
    // P1 = P - Y^T Y, lower triangle then mirrored (the exact update is symmetric)
    std::vector<W> P1((size_t)n * n);
    ParallelFor(0, n, 8, [&](int64_t i) {
        const W* ai = &A[(size_t)i * m2];
        for (int64_t j = 0; j <= i; ++j) {
            const W* aj = &A[(size_t)j * m2];
            W a0 = 0, a1 = 0, a2 = 0, a3 = 0;
            int64_t r = 0;
            for (; r + 3 < m2; r += 4) { a0 += ai[r] * aj[r]; a1 += ai[r + 1] * aj[r + 1]; a2 += ai[r + 2] * aj[r + 2]; a3 += ai[r + 3] * aj[r + 3]; }
            for (; r < m2; ++r) a0 += ai[r] * aj[r];
            // P is symmetric on input up to rounding; the reference reads P(i, j) as stored.  Take the mean of the stored pair.
            W pij = ((W)P[(size_t)j * n + i] + (W)P[(size_t)i * n + j]) / 2;
            W v = pij - ((a0 + a1) + (a2 + a3));
            P1[(size_t)j * n + i] = v; P1[(size_t)i * n + j] = v;
        }
    });
    // quaternion normalisation (EKF.cpp:1652-1711): x[3:7] /= |q|,  P <- J P J^T,  J = I except J[3:7, 3:7] = (|q|^2 I - q q^T) / |q|^3
    W q[4] = {xs[3], xs[4], xs[5], xs[6]};
    W qq = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
    W q_len = std::sqrt(qq);
    if (!IsClose<double>(1.0, (double)q_len)) {
        for (int i = 0; i < 4; ++i) xs[3 + i] = q[i] / q_len;
        W dq[4][4];
        W mult = 1 / (qq * q_len);
        for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) dq[i][j] = ((i == j ? qq : (W)0) - q[i] * q[j]) * mult;
        std::vector<W> rows(4 * (size_t)n);
        for (int i = 0; i < 4; ++i) for (int64_t c = 0; c < n; ++c) { W acc = 0; for (int k = 0; k < 4; ++k) acc += dq[i][k] * P1[(size_t)c * n + 3 + k]; rows[(size_t)i * n + c] = acc; }
        for (int i = 0; i < 4; ++i) for (int64_t c = 0; c < n; ++c) P1[(size_t)c * n + 3 + i] = rows[(size_t)i * n + c];
        std::vector<W> cols(4 * (size_t)n);
        for (int j = 0; j < 4; ++j) for (int64_t r = 0; r < n; ++r) { W acc = 0; for (int k = 0; k < 4; ++k) acc += P1[(size_t)(3 + k) * n + r] * dq[j][k]; cols[(size_t)j * n + r] = acc; }
        for (int j = 0; j < 4; ++j) for (int64_t r = 0; r < n; ++r) P1[(size_t)(3 + j) * n + r] = cols[(size_t)j * n + r];
    }
    if (fix_symmetry)
        for (int64_t i = 0; i < n; ++i) for (int64_t j = i + 1; j < n; ++j) { W v = (P1[(size_t)j * n + i] + P1[(size_t)i * n + j]) / 2; P1[(size_t)j * n + i] = v; P1[(size_t)i * n + j] = v; }
    for (int64_t i = 0; i < n; ++i) {
        if (P1[(size_t)i * n + i] >= 0) continue;
        for (int64_t j = 0; j < n; ++j) { P1[(size_t)j * n + i] = 0; P1[(size_t)i * n + j] = 0; }
    }
    for (size_t i = 0; i < (size_t)n * n; ++i) P[i] = (double)P1[i];
    for (int64_t i = 0; i < n; ++i) x[i] = (double)xs[i];
    return true;
}

}  // namespace srk_oracle
