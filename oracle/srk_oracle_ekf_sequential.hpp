// TEST INFRASTRUCTURE ONLY — CPU oracle: the per-observation update variants of suriko-engine's MonoSLAM
// ("EKF.cpp" = /root/reference/cpp_impl/suriko-engine/src/davison-mono-slam.cpp):
//   per_component == 0   ProcessFrame_OneObservationPerUpdate                  EKF.cpp:1153-1269  (2x2 innovation per observed point)
//   per_component == 1   ProcessFrame_OneComponentOfOneObservationPerUpdate    EKF.cpp:1525-1650  (scalar innovation per pixel component)
// Every update re-derives the measurement Jacobian at the LATEST state, applies x += K (z - h), P -= (K S) K^T, the quaternion
// normalisation (:1652-1711) and FixSymmetricMat (:4308).  Parity unpinned (no reference test for this path).
#pragma once
#include <vector>
#include "srk_oracle_ekf.hpp"
#include "srk_oracle_ekf_ransac.hpp"

namespace srk_oracle {

inline void EkfSequentialUpdate(std::vector<double>* x, EkfMat* P, int64_t m, const int64_t* pt_off, int s, const double* z, double meas_var, const EkfCamera& cam,
                                bool per_component) {
    const size_t n = x->size();
    std::vector<double> PH(n * 2), K(n * 2), KS(n * 2);
    auto fix_sym = [&]() { for (size_t i = 0; i < n; ++i) for (size_t j = i + 1; j < n; ++j) { double v = ((*P)(i, j) + (*P)(j, i)) / 2; (*P)(i, j) = v; (*P)(j, i) = v; } };
    for (int64_t i = 0; i < m; ++i) {
        const size_t off = (size_t)pt_off[i];
        const int ncomp = per_component ? 2 : 1;
        for (int pass = 0; pass < ncomp; ++pass) {
            double Hx[26], Hy[12], hd[2];
            EkfMeasurementJacobian(cam, x->data(), x->data() + off, s, Hx, Hy, hd);     // at the latest state (:1177-1197, :1559-1581)
            const int k0 = per_component ? pass : 0, k1 = per_component ? pass + 1 : 2;
            const int d = k1 - k0;
            // innovation covariance S = Hx Pxx Hx^T + mid + mid^T + Hy Pyy Hy^T + R  (:1208-1216; scalar form :1591-1597)
            double S[2][2] = {{0, 0}, {0, 0}};
            for (int a = k0; a < k1; ++a)
                for (int b = k0; b < k1; ++b) {
                    double xx = 0, mid = 0, midt = 0, yy = 0;
                    for (int c = 0; c < 13; ++c) for (int e = 0; e < 13; ++e) xx += Hx[a * 13 + c] * (*P)((size_t)c, (size_t)e) * Hx[b * 13 + e];
                    for (int c = 0; c < 13; ++c) for (int e = 0; e < s; ++e) { mid += Hx[a * 13 + c] * (*P)((size_t)c, off + e) * Hy[b * s + e]; midt += Hx[b * 13 + c] * (*P)((size_t)c, off + e) * Hy[a * s + e]; }
                    for (int c = 0; c < s; ++c) for (int e = 0; e < s; ++e) yy += Hy[a * s + c] * (*P)(off + c, off + e) * Hy[b * s + e];
                    S[a - k0][b - k0] = xx + mid + midt + yy + (a == b ? meas_var : 0.0);
                }
            double Si[2][2] = {{0, 0}, {0, 0}};
            if (d == 2) { const double det = S[0][0] * S[1][1] - S[0][1] * S[1][0]; Si[0][0] = S[1][1] / det; Si[0][1] = -S[0][1] / det; Si[1][0] = -S[1][0] / det; Si[1][1] = S[0][0] / det; }
            else Si[0][0] = 1 / S[0][0];
            // gain K = (P[:, cam] Hx^T + P[:, pnt] Hy^T) S^-1  (:1220-1226, :1603-1605)
            for (size_t r = 0; r < n; ++r)
                for (int a = 0; a < d; ++a) {
                    double v = 0;
                    for (int c = 0; c < 13; ++c) v += (*P)(r, (size_t)c) * Hx[(k0 + a) * 13 + c];
                    for (int c = 0; c < s; ++c) v += (*P)(r, off + c) * Hy[(k0 + a) * s + c];
                    PH[r * 2 + a] = v;
                }
            double delta[2] = {0, 0};
            for (int a = 0; a < d; ++a) delta[a] = z[2 * i + k0 + a] - hd[k0 + a];
            for (size_t r = 0; r < n; ++r) {
                for (int a = 0; a < d; ++a) { double v = 0; for (int b = 0; b < d; ++b) v += PH[r * 2 + b] * Si[b][a]; K[r * 2 + a] = v; }
                for (int a = 0; a < d; ++a) { double v = 0; for (int b = 0; b < d; ++b) v += K[r * 2 + b] * S[b][a]; KS[r * 2 + a] = v; }
            }
            for (size_t r = 0; r < n; ++r) { double v = 0; for (int a = 0; a < d; ++a) v += K[r * 2 + a] * delta[a]; (*x)[r] += v; }      // :1247, :1635
            for (size_t c2 = 0; c2 < n; ++c2) for (size_t r = 0; r < n; ++r) { double v = 0; for (int a = 0; a < d; ++a) v += KS[r * 2 + a] * K[c2 * 2 + a]; (*P)(r, c2) -= v; }   // :1248, :1636
            EkfNormalizeQuaternion(x, P);      // :1250, :1638
            fix_sym();                          // :1252-1253, :1640-1641
        }
    }
}

}  // namespace srk_oracle
