// TEST INFRASTRUCTURE ONLY (CPU oracle) -- restatement of the input front end of the BA demos:
//   Triangulate3DPointByLeastSquares   /root/reference/cpp_impl/suriko-engine/src/obs-geom.cpp:679-727
//   DecomposeProjMat                   /root/reference/cpp_impl/suriko-engine/src/obs-geom.cpp:606-677
// The reference solves the 2k x 3 system with Eigen's colPivHouseholderQr (obs-geom.cpp:714); Eigen is not vendored in
// /root/reference, so the published algorithm is restated: Householder reflections with column pivoting on the largest
// remaining column norm, then back substitution in the pivoted order.  Pinned by round trips (project known points through
// known cameras, triangulate, compare; compose P from K, R, t, decompose, compare) in tests/test_cpu_frontend.py.
#pragma once
#include <cmath>
#include <cstdint>
#include <vector>

namespace srk_oracle {

// A: [rows x 3] row-major, b: [rows]; least-squares solution of A x = b by column-pivoted Householder QR
inline void ColPivHouseholderSolve3(std::vector<double>& A, std::vector<double>& b, size_t rows, double x[3]) {
    int perm[3] = {0, 1, 2};
    for (int k = 0; k < 3; ++k) {
        // pivot: remaining column with the largest norm below row k
        int best = k; double best_n = -1.0;
        for (int c = k; c < 3; ++c) {
            double n2 = 0.0;
            for (size_t r = k; r < rows; ++r) n2 += A[r * 3 + c] * A[r * 3 + c];
            if (n2 > best_n) { best_n = n2; best = c; }
        }
        if (best != k) {
            for (size_t r = 0; r < rows; ++r) std::swap(A[r * 3 + k], A[r * 3 + best]);
            std::swap(perm[k], perm[best]);
        }
        // Householder vector for column k
        double norm = std::sqrt(best_n);
        if (norm == 0.0) continue;
        const double alpha = A[k * 3 + k] > 0 ? -norm : norm;
        std::vector<double> v(rows - k);
        for (size_t r = k; r < rows; ++r) v[r - k] = A[r * 3 + k];
        v[0] -= alpha;
        double vn2 = 0.0;
        for (double e : v) vn2 += e * e;
        if (vn2 == 0.0) continue;
        for (int c = k; c < 3; ++c) {
            double dot = 0.0;
            for (size_t r = k; r < rows; ++r) dot += v[r - k] * A[r * 3 + c];
            const double f = 2.0 * dot / vn2;
            for (size_t r = k; r < rows; ++r) A[r * 3 + c] -= f * v[r - k];
        }
        double dot = 0.0;
        for (size_t r = k; r < rows; ++r) dot += v[r - k] * b[r];
        const double f = 2.0 * dot / vn2;
        for (size_t r = k; r < rows; ++r) b[r] -= f * v[r - k];
    }
    double y[3];
    for (int k = 2; k >= 0; --k) {
        double s = b[k];
        for (int c = k + 1; c < 3; ++c) s -= A[k * 3 + c] * y[c];
        y[k] = s / A[k * 3 + k];
    }
    for (int k = 0; k < 3; ++k) x[perm[k]] = y[k];
}

// obs-geom.cpp:679-727; proj: 3x4 per frame, column-major
inline void Triangulate(int64_t n_tracks, const int64_t* track_begin, const int32_t* obs_frame, const double* obs_xy, const double* proj, double f0,
                        double* out) {
    for (int64_t t = 0; t < n_tracks; ++t) {
        const int64_t b0 = track_begin[t], e0 = track_begin[t + 1];
        const size_t k = (size_t)(e0 - b0);
        std::vector<double> A(2 * k * 3), B(2 * k);
        for (size_t i = 0; i < k; ++i) {
            const int64_t o = b0 + (int64_t)i;
            const double x = obs_xy[2 * o], y = obs_xy[2 * o + 1];
            const double* P = proj + (size_t)obs_frame[o] * 12;
            auto Pm = [&](int r, int c) { return P[c * 3 + r]; };
            for (int c = 0; c < 3; ++c) {
                A[(2 * i) * 3 + c] = x * Pm(2, c) - f0 * Pm(0, c);        // :699-701
                A[(2 * i + 1) * 3 + c] = y * Pm(2, c) - f0 * Pm(1, c);    // :702-704
            }
            B[2 * i] = -(x * Pm(2, 3) - f0 * Pm(0, 3));                   // :706
            B[2 * i + 1] = -(y * Pm(2, 3) - f0 * Pm(1, 3));               // :707
        }
        ColPivHouseholderSolve3(A, B, 2 * k, out + 3 * t);
    }
}

}  // namespace srk_oracle
