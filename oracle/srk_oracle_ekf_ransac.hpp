// TEST INFRASTRUCTURE ONLY — CPU oracle: plain C++17 restatement of the hypothesis scoring of the 1-point RANSAC EKF update of
// suriko-engine's MonoSLAM ("EKF.cpp" = /root/reference/cpp_impl/suriko-engine/src/davison-mono-slam.cpp):
//   EkfRotMatFromQuat          RotMatFromQuat                                     quat.cpp:75-91
//   EkfProjectSalientPoint     ProjectInternalSalientPoint -> InternalSalientPointToCamera (scaled_by_inv_dist = true) ->
//                              ProjectCameraSalientPoint -> DistortPixel         EKF.cpp:2887-3033, polar direction :469-484
//   EkfRansacConsensus         OnePointRansac_GetConsensusMatches                 EKF.cpp:1271-1391
// State layout: camera [pos_w 3 | orientation_wfc quaternion 4 | velocity 3 | angular velocity 3] (EKF.cpp:3687-3704); a salient point
// is pos_w[3] (XYZ, s = 3) or [first_cam_pos_w 3 | azimuth theta | elevation phi | inverse distance rho] (s = 6) (EKF.cpp:3423-3445).
// Quirks kept: the closed-form root of the cubic distortion takes its cube roots with the FLOAT exponent 1.0f/3 widened to double
// (EKF.cpp:2990-2991); the hypothesis state x + K (z - h) is used as it is (its quaternion is not re-normalised, :1349-1367);
// a later hypothesis replaces the best one only with strictly more support (:1383).
// Third-party arithmetic not in /root/reference: Eigen's PolynomialSolver (unsupported module, version unpinned) finds the real root
// of rd + k1 rd^3 + k2 rd^5 = ru through the eigenvalues of the companion matrix; the polynomial is strictly increasing for
// k1, k2 >= 0, so its single real root is restated with Newton's iteration from rd = ru (agrees to the last bits).
// Parity unpinned: the reference has no test for this path.
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <vector>
#include "srk_oracle_ekf.hpp"

namespace srk_oracle {

struct EkfCamera { double fx_pix, fy_pix, cx, cy, dx_mm, dy_mm, k1, k2; int enable_distortion; };

inline void EkfRotMatFromQuat(const double* q, double R[3][3]) {   // quat.cpp:75-91
    R[0][0] = q[0] * q[0] + q[1] * q[1] - q[2] * q[2] - q[3] * q[3];
    R[0][1] = 2 * (q[1] * q[2] - q[0] * q[3]);
    R[0][2] = 2 * (q[1] * q[3] + q[0] * q[2]);
    R[1][0] = 2 * (q[1] * q[2] + q[0] * q[3]);
    R[1][1] = q[0] * q[0] - q[1] * q[1] + q[2] * q[2] - q[3] * q[3];
    R[1][2] = 2 * (q[2] * q[3] - q[0] * q[1]);
    R[2][0] = 2 * (q[1] * q[3] - q[0] * q[2]);
    R[2][1] = 2 * (q[2] * q[3] + q[0] * q[1]);
    R[2][2] = q[0] * q[0] - q[1] * q[1] - q[2] * q[2] + q[3] * q[3];
}

inline void EkfDistortPixel(const EkfCamera& c, const double hu[2], double hd[2]) {   // EKF.cpp:2960-3005
    const double ru = std::sqrt((c.dx_mm * (hu[0] - c.cx)) * (c.dx_mm * (hu[0] - c.cx)) + (c.dy_mm * (hu[1] - c.cy)) * (c.dy_mm * (hu[1] - c.cy)));
    double rd;
    if (c.k2 != 0) {
        rd = ru;
        for (int it = 0; it < 60; ++it) {
            const double r2 = rd * rd;
            const double f = rd + c.k1 * r2 * rd + c.k2 * r2 * r2 * rd - ru;
            const double df = 1 + 3 * c.k1 * r2 + 5 * c.k2 * r2 * r2;
            const double step = f / df;
            rd -= step;
            if (std::fabs(step) <= 1e-17 * std::fabs(rd)) break;
        }
    } else if (c.k1 == 0) {
        rd = ru;
    } else {
        const double third = (double)(1.0f / 3);   // the reference writes 1.0f / 3
        const double e = std::pow(9 * c.k1 * c.k1 * ru + std::sqrt(3 * c.k1 * c.k1 * c.k1 * (4 + 27 * c.k1 * ru * ru)), third);
        rd = (-2 * std::pow(3.0, third) * c.k1 + std::pow(2.0, third) * e * e) / (std::pow(6.0, 2.0 / 3) * c.k1 * e);
    }
    const double stretch = 1 + c.k1 * (rd * rd) + c.k2 * (rd * rd) * (rd * rd);
    hd[0] = c.cx + (hu[0] - c.cx) / stretch;
    hd[1] = c.cy + (hu[1] - c.cy) / stretch;
}

// cam13 = camera state, sp = salient point variables (s = 3 or 6) -> distorted pixel
inline void EkfProjectSalientPoint(const EkfCamera& c, const double* cam13, const double* sp, int s, double hd[2]) {
    double Rwfc[3][3];
    EkfRotMatFromQuat(cam13 + 3, Rwfc);
    double v[3];
    if (s == 3) {                                   // A.22 (EKF.cpp:2905): Rcw (pos_w - cam_pos)
        for (int k = 0; k < 3; ++k) v[k] = sp[k] - cam13[k];
    } else {                                        // A.21 (EKF.cpp:2935): Rcw (rho (first_cam_pos - cam_pos) + m(theta, phi))
        const double cos_th = std::cos(sp[3]), sin_th = std::sin(sp[3]), cos_ph = std::cos(sp[4]), sin_ph = std::sin(sp[4]);
        const double m[3] = {cos_ph * sin_th, -sin_ph, cos_ph * cos_th};
        for (int k = 0; k < 3; ++k) v[k] = sp[5] * (sp[k] - cam13[k]) + m[k];
    }
    double pc[3];                                   // Rcw = Rwfc^T
    for (int r = 0; r < 3; ++r) pc[r] = Rwfc[0][r] * v[0] + Rwfc[1][r] * v[1] + Rwfc[2][r] * v[2];
    double hu[2] = {c.cx - c.fx_pix * pc[0] / pc[2], c.cy - c.fy_pix * pc[1] / pc[2]};   // EKF.cpp:3021-3022
    if (c.enable_distortion) EkfDistortPixel(c, hu, hd); else { hd[0] = hu[0]; hd[1] = hu[1]; }
}

// OnePointRansac_GetConsensusMatches.  x [n], P [n x n] column-major; per matched point i: Hcam [2 x 13], Hpt [2 x s] (row-major
// per observation row), pt_off[i], measured corner z[2i..2i+1].  support[i] = size of the consensus set of hypothesis i;
// returns the winning hypothesis (-1 when every support is zero) and its inlier mask.
inline int EkfRansacConsensus(const std::vector<double>& x, const EkfMat& P, int64_t m, const double* Hcam, const double* Hpt, const int64_t* pt_off, int s,
                              const double* z, double meas_var, const EkfCamera& cam, double max_divergence_pix, int32_t* support, unsigned char* best_inliers) {
    const size_t n = x.size();
    int best = -1; int64_t best_count = 0;
    std::vector<double> PHxy(n * 2), Knew(n * 2), xn(n);
    std::vector<unsigned char> mask((size_t)m);
    for (int64_t i = 0; i < m; ++i) {
        const double* Hx = Hcam + (size_t)(2 * i) * 13;      // Hx[k*13 + c]
        const double* Hy = Hpt + (size_t)(2 * i) * s;        // Hy[k*s + c]
        const size_t off = (size_t)pt_off[i];
        // 1. innovation variance S = Hx Pxx Hx^T + mid + mid^T + Hy Pyy Hy^T + Rk   (:1321-1326)
        double S[2][2];
        for (int a = 0; a < 2; ++a)
            for (int b = 0; b < 2; ++b) {
                double t1 = 0, mid_ab = 0, mid_ba = 0, t3 = 0;
                for (int p = 0; p < 13; ++p) for (int q = 0; q < 13; ++q) t1 += Hx[a * 13 + p] * P((size_t)p, (size_t)q) * Hx[b * 13 + q];
                for (int p = 0; p < 13; ++p) for (int q = 0; q < s; ++q) {
                    mid_ab += Hx[a * 13 + p] * P((size_t)p, off + q) * Hy[b * s + q];
                    mid_ba += Hx[b * 13 + p] * P((size_t)p, off + q) * Hy[a * s + q];
                }
                for (int p = 0; p < s; ++p) for (int q = 0; q < s; ++q) t3 += Hy[a * s + p] * P(off + p, off + q) * Hy[b * s + q];
                S[a][b] = t1 + mid_ab + mid_ba + t3 + (a == b ? meas_var : 0.0);
            }
        const double det = S[0][0] * S[1][1] - S[0][1] * S[1][0];
        const double idet = 1.0 / det;                       // Eigen's fixed 2x2 inverse: cofactors times 1/det
        const double Si[2][2] = {{S[1][1] * idet, -S[0][1] * idet}, {-S[1][0] * idet, S[0][0] * idet}};
        // 2. gain K = (P[:, cam] Hx^T + P[:, pt] Hy^T) S^-1   (:1331-1335)
        for (size_t r = 0; r < n; ++r)
            for (int a = 0; a < 2; ++a) {
                double t = 0;
                for (int p = 0; p < 13; ++p) t += P(r, (size_t)p) * Hx[a * 13 + p];
                double t2 = 0;
                for (int q = 0; q < s; ++q) t2 += P(r, off + q) * Hy[a * s + q];
                PHxy[a * n + r] = t + t2;
            }
        for (size_t r = 0; r < n; ++r)
            for (int a = 0; a < 2; ++a) Knew[a * n + r] = PHxy[0 * n + r] * Si[0][a] + PHxy[1 * n + r] * Si[1][a];
        // 3. hypothesis state x + K (z_i - h_i)   (:1343-1347)
        double hd[2];
        EkfProjectSalientPoint(cam, x.data(), x.data() + off, s, hd);
        const double r0 = z[2 * i] - hd[0], r1 = z[2 * i + 1] - hd[1];
        for (size_t r = 0; r < n; ++r) xn[r] = x[r] + (Knew[0 * n + r] * r0 + Knew[1 * n + r] * r1);
        // support of the hypothesis (:1349-1381)
        int64_t cnt = 0;
        for (int64_t j = 0; j < m; ++j) {
            double a_hd[2];
            EkfProjectSalientPoint(cam, xn.data(), xn.data() + (size_t)pt_off[j], s, a_hd);
            const double d0 = z[2 * j] - a_hd[0], d1 = z[2 * j + 1] - a_hd[1];
            const double dist = std::sqrt(d0 * d0 + d1 * d1);
            mask[(size_t)j] = dist < max_divergence_pix ? 1 : 0;
            cnt += mask[(size_t)j];
        }
        if (support != nullptr) support[i] = (int32_t)cnt;
        if (cnt > best_count) {                              // strictly more (:1383)
            best_count = cnt; best = (int)i;
            if (best_inliers != nullptr) for (int64_t j = 0; j < m; ++j) best_inliers[j] = mask[(size_t)j];
        }
    }
    if (best < 0 && best_inliers != nullptr) for (int64_t j = 0; j < m; ++j) best_inliers[j] = 0;
    return best;
}

}  // namespace srk_oracle
