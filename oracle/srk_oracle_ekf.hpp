// TEST INFRASTRUCTURE ONLY — CPU oracle: plain C++17 restatement of the dense covariance chain of suriko-engine's
// Davison/Civera MonoSLAM EKF ("EKF.cpp" = /root/reference/cpp_impl/suriko-engine/src/davison-mono-slam.cpp).
//   EkfPredictCovariance   : PredictEstimVars, covariance part                       EKF.cpp:669-693
//   EkfStackedUpdate       : ProcessFrame_StackedObservationsPerUpdateCore           EKF.cpp:977-1125
//                            (H*P, S = HP*H^T + R, S^-1 by partial-pivot LU as Eigen's .inverse() on a dynamic matrix,
//                             K = HP^T*S^-1, x += K(z-h), P -= (K*S)*K^T, quaternion normalisation :1652-1711,
//                             FixSymmetricMat :4308 / eigen-helpers.hpp:99-104, EnsureNonnegativeStateVariance :1739-1750)
// The reference multiplies the DENSE H ([2m x n], zeros included, EKF.cpp:1004-1007); so does this restatement.
// Parity unpinned: the reference has no test for this path; the restatement is anchored on the call sites.
// Matrices are column-major like Eigen's default.  Third-party arithmetic not in /root/reference: Eigen3 (version unpinned).
#pragma once
#include <cmath>
#include <cstddef>
#include <vector>
#include "srk_oracle_geom.hpp"

namespace srk_oracle {

struct EkfMat {  // column-major dynamic matrix
    size_t rows = 0, cols = 0;
    std::vector<double> d;
    EkfMat() = default;
    EkfMat(size_t r, size_t c) : rows(r), cols(c), d(r * c, 0.0) {}
    double& operator()(size_t r, size_t c) { return d[c * rows + r]; }
    double operator()(size_t r, size_t c) const { return d[c * rows + r]; }
};

inline EkfMat MatMul(const EkfMat& A, const EkfMat& B, bool transB = false) {  // A * B  or  A * B^T
    size_t m = A.rows, k = A.cols, n = transB ? B.rows : B.cols;
    EkfMat C(m, n);
    for (size_t j = 0; j < n; ++j)
        for (size_t p = 0; p < k; ++p) {
            double b = transB ? B(j, p) : B(p, j);
            if (b == 0.0) continue;   // skipping exact zeros does not change the sum
            const double* a = &A.d[p * m];
            double* c = &C.d[j * m];
            for (size_t i = 0; i < m; ++i) c[i] += a[i] * b;
        }
    return C;
}

// Eigen's inverse() of a dynamic matrix = PartialPivLU().inverse(): LU with row pivoting, then solve against the identity.
inline bool InversePartialPivLU(const EkfMat& A, EkfMat* inv) {
    size_t n = A.rows;
    EkfMat lu = A;
    std::vector<size_t> perm(n);
    for (size_t i = 0; i < n; ++i) perm[i] = i;
    for (size_t k = 0; k < n; ++k) {
        size_t piv = k; double best = std::fabs(lu(k, k));
        for (size_t i = k + 1; i < n; ++i) if (std::fabs(lu(i, k)) > best) { best = std::fabs(lu(i, k)); piv = i; }
        if (best == 0.0) return false;
        if (piv != k) { for (size_t j = 0; j < n; ++j) std::swap(lu(k, j), lu(piv, j)); std::swap(perm[k], perm[piv]); }
        for (size_t i = k + 1; i < n; ++i) lu(i, k) /= lu(k, k);
        for (size_t j = k + 1; j < n; ++j) {
            double u = lu(k, j);
            if (u == 0.0) continue;
            for (size_t i = k + 1; i < n; ++i) lu(i, j) -= lu(i, k) * u;
        }
    }
    *inv = EkfMat(n, n);
    std::vector<double> y(n);
    std::vector<double> rowmaj(n * n);   // the same factors, rows contiguous: the substitutions below walk rows (same operations, same order)
    for (size_t i = 0; i < n; ++i) for (size_t j = 0; j < n; ++j) rowmaj[i * n + j] = lu(i, j);
    for (size_t c = 0; c < n; ++c) {
        for (size_t i = 0; i < n; ++i) y[i] = perm[i] == c ? 1.0 : 0.0;
        for (size_t i = 0; i < n; ++i) { const double* l = &rowmaj[i * n]; double s = y[i]; for (size_t j = 0; j < i; ++j) s -= l[j] * y[j]; y[i] = s; }
        for (size_t ii = n; ii-- > 0;) { const double* l = &rowmaj[ii * n]; double s = y[ii]; for (size_t j = ii + 1; j < n; ++j) s -= l[j] * y[j]; y[ii] = s / l[ii]; }
        for (size_t i = 0; i < n; ++i) (*inv)(i, c) = y[i];
    }
    return true;
}

// EKF.cpp:669-693.  F, GQGt: 13x13 column-major.  P: n x n.
inline void EkfPredictCovariance(EkfMat* P, const double* F13, const double* GQGt13, bool fix_symmetry) {
    const size_t n = P->rows, c = 13;
    EkfMat F(c, c), Q(c, c), Pvv(c, c), Pvm(c, n - c);
    for (size_t i = 0; i < c * c; ++i) { F.d[i] = F13[i]; Q.d[i] = GQGt13[i]; }
    for (size_t i = 0; i < c; ++i) for (size_t j = 0; j < c; ++j) Pvv(i, j) = (*P)(i, j);
    for (size_t i = 0; i < c; ++i) for (size_t j = 0; j < n - c; ++j) Pvm(i, j) = (*P)(i, c + j);
    EkfMat Pvv_new = MatMul(MatMul(F, Pvv), F, true);
    for (size_t i = 0; i < c * c; ++i) Pvv_new.d[i] += Q.d[i];
    EkfMat Pvm_new = MatMul(F, Pvm);
    for (size_t i = 0; i < c; ++i) for (size_t j = 0; j < c; ++j) (*P)(i, j) = Pvv_new(i, j);
    for (size_t i = 0; i < c; ++i) for (size_t j = 0; j < n - c; ++j) { (*P)(i, c + j) = Pvm_new(i, j); (*P)(c + j, i) = Pvm_new(i, j); }
    if (fix_symmetry) for (size_t i = 0; i < n; ++i) for (size_t j = i + 1; j < n; ++j) { double v = ((*P)(i, j) + (*P)(j, i)) / 2; (*P)(i, j) = v; (*P)(j, i) = v; }
}

// EKF.cpp:1652-1711
inline void EkfNormalizeQuaternion(std::vector<double>* x, EkfMat* P) {
    const size_t n = P->rows;
    double q[4] = {(*x)[3], (*x)[4], (*x)[5], (*x)[6]};
    double q_len = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    if (IsClose<double>(1.0, q_len)) return;
    for (int i = 0; i < 4; ++i) (*x)[3 + i] = q[i] / q_len;
    double dq[4][4];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            if (i == j) { double s = 0; for (int k = 0; k < 4; ++k) if (k != i) s += q[k] * q[k]; dq[i][j] = s; }
            else dq[i][j] = -q[i] * q[j];
        }
    double q_mult = std::pow(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3], (double)-1.5f);
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) dq[i][j] *= q_mult;
    // q_up = P[0:3, 3:7] * dq^T ; q_down = P[7:, 3:7] * dq^T ; centre = dq * P[3:7, 3:7] * dq^T   (all from the un-mutated P)
    std::vector<double> up(3 * 4), down((n - 7) * 4), cen(16);
    for (size_t r = 0; r < 3; ++r) for (int j = 0; j < 4; ++j) { double s = 0; for (int k = 0; k < 4; ++k) s += (*P)(r, 3 + k) * dq[j][k]; up[r * 4 + j] = s; }
    for (size_t r = 0; r < n - 7; ++r) for (int j = 0; j < 4; ++j) { double s = 0; for (int k = 0; k < 4; ++k) s += (*P)(7 + r, 3 + k) * dq[j][k]; down[r * 4 + j] = s; }
    for (size_t r = 0; r < 3; ++r) for (int j = 0; j < 4; ++j) { (*P)(r, 3 + j) = up[r * 4 + j]; (*P)(3 + j, r) = up[r * 4 + j]; }
    double tmp[4][4];
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) { double s = 0; for (int k = 0; k < 4; ++k) s += dq[i][k] * (*P)(3 + k, 3 + j); tmp[i][j] = s; }
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) { double s = 0; for (int k = 0; k < 4; ++k) s += tmp[i][k] * dq[j][k]; cen[i * 4 + j] = s; }
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) (*P)(3 + i, 3 + j) = cen[i * 4 + j];
    for (size_t r = 0; r < n - 7; ++r) for (int j = 0; j < 4; ++j) { (*P)(7 + r, 3 + j) = down[r * 4 + j]; (*P)(3 + j, 7 + r) = down[r * 4 + j]; }
}

// EKF.cpp:977-1125.  H dense [2m x n]; z, h: [2m]; R = meas_var * I.
inline bool EkfStackedUpdate(std::vector<double>* x, EkfMat* P, const EkfMat& H, const std::vector<double>& z, const std::vector<double>& hpred,
                             double meas_var, bool fix_symmetry) {
    const size_t n = P->rows, m2 = H.rows;
    EkfMat HP = MatMul(H, *P);                 // :1004
    EkfMat S = MatMul(HP, H, true);            // :1006
    for (size_t i = 0; i < m2; ++i) S(i, i) += meas_var;   // :1007
    EkfMat Sinv;
    if (!InversePartialPivLU(S, &Sinv)) return false;       // :1019
    EkfMat HPt(n, m2);
    for (size_t i = 0; i < m2; ++i) for (size_t j = 0; j < n; ++j) HPt(j, i) = HP(i, j);
    EkfMat K = MatMul(HPt, Sinv);              // :1029
    for (size_t r = 0; r < n; ++r) { double s = 0; for (size_t i = 0; i < m2; ++i) s += K(r, i) * (z[i] - hpred[i]); (*x)[r] += s; }   // :1081
    EkfMat KS = MatMul(K, S);                  // :1109
    EkfMat KSKt = MatMul(KS, K, true);         // :1114
    for (size_t i = 0; i < n * n; ++i) P->d[i] -= KSKt.d[i];
    EkfNormalizeQuaternion(x, P);              // :1118
    if (fix_symmetry) for (size_t i = 0; i < n; ++i) for (size_t j = i + 1; j < n; ++j) { double v = ((*P)(i, j) + (*P)(j, i)) / 2; (*P)(i, j) = v; (*P)(j, i) = v; }   // :1120-1121
    for (size_t i = 0; i < n; ++i) {           // :1739-1750
        if ((*P)(i, i) >= 0) continue;
        for (size_t j = 0; j < n; ++j) { (*P)(i, j) = 0; (*P)(j, i) = 0; }
    }
    return true;
}

}  // namespace srk_oracle
