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

// Deriv_hd_by_cam_state_and_sal_pnt (EKF.cpp:3067-3113): hd and its derivatives by the 13 camera variables (Hx[k*13 + c]) and by the
// salient point's own s variables (Hy[k*s + c]), as the chain  hd <- hu (A.32/A.33, :2651-2687) <- hc (A.34, :2689-2704) <- state:
// camera position A.35/A.36, quaternion A.37-A.49 through q_cw = conj(q_wc) (:2747-2817; velocity columns stay zero), point A.51-A.55 (:2819-2865).
inline void EkfMeasurementJacobian(const EkfCamera& c, const double* cam13, const double* sp, int s, double* Hx, double* Hy, double hd[2]) {
    EkfProjectSalientPoint(c, cam13, sp, s, hd);
    double Rwfc[3][3];
    EkfRotMatFromQuat(cam13 + 3, Rwfc);
    double Rcw[3][3];
    for (int r = 0; r < 3; ++r) for (int k = 0; k < 3; ++k) Rcw[r][k] = Rwfc[k][r];
    // part2 = what Rcw multiplies (the scaled camera-frame point hc = Rcw part2), m = first-camera unity direction
    double part2[3], mdir[3] = {0, 0, 0}, rho = 1.0;
    if (s == 3) {
        for (int k = 0; k < 3; ++k) part2[k] = sp[k] - cam13[k];
    } else {
        const double cos_th = std::cos(sp[3]), sin_th = std::sin(sp[3]), cos_ph = std::cos(sp[4]), sin_ph = std::sin(sp[4]);
        mdir[0] = cos_ph * sin_th; mdir[1] = -sin_ph; mdir[2] = cos_ph * cos_th;
        rho = sp[5];
        for (int k = 0; k < 3; ++k) part2[k] = rho * (sp[k] - cam13[k]) + mdir[k];
    }
    double hc[3];
    for (int r = 0; r < 3; ++r) hc[r] = Rcw[r][0] * part2[0] + Rcw[r][1] * part2[1] + Rcw[r][2] * part2[2];
    // hd_by_hu = inverse of hu_by_hd (A.32, A.33)
    double hd_by_hu[2][2] = {{1, 0}, {0, 1}};
    if (c.enable_distortion) {
        const double ax = hd[0] - c.cx, ay = hd[1] - c.cy;
        const double rd = std::sqrt((c.dx_mm * ax) * (c.dx_mm * ax) + (c.dy_mm * ay) * (c.dy_mm * ay));   // Calc_rd, A.24 (:48-59)
        const double stretch = 1 + c.k1 * (rd * rd) + c.k2 * (rd * rd) * (rd * rd);
        const double kk = c.k1 + 2 * c.k2 * (rd * rd);
        const double side = 2 * kk * ay * ax;
        const double r00 = stretch + 2 * kk * ((c.dx_mm * ax) * (c.dx_mm * ax)), r11 = stretch + 2 * kk * ((c.dy_mm * ay) * (c.dy_mm * ay));
        const double r10 = side * (c.dx_mm * c.dx_mm), r01 = side * (c.dy_mm * c.dy_mm);
        const double idet = 1.0 / (r00 * r11 - r01 * r10);
        hd_by_hu[0][0] = r11 * idet; hd_by_hu[0][1] = -r01 * idet; hd_by_hu[1][0] = -r10 * idet; hd_by_hu[1][1] = r00 * idet;
    }
    const double hu_by_hc[2][3] = {{-c.fx_pix / hc[2], 0.0, c.fx_pix * hc[0] / (hc[2] * hc[2])}, {0.0, -c.fy_pix / hc[2], c.fy_pix * hc[1] / (hc[2] * hc[2])}};
    double D[2][3];   // hd_by_hc = hd_by_hu * hu_by_hc
    for (int a = 0; a < 2; ++a) for (int k = 0; k < 3; ++k) D[a][k] = hd_by_hu[a][0] * hu_by_hc[0][k] + hd_by_hu[a][1] * hu_by_hc[1][k];
    for (int e = 0; e < 2 * 13; ++e) Hx[e] = 0.0;
    // camera position: hc_by_rwc = -Rcw (A.36) or -rho Rcw (A.35)
    for (int a = 0; a < 2; ++a)
        for (int k = 0; k < 3; ++k) {
            double t = 0;
            for (int r = 0; r < 3; ++r) t += D[a][r] * (-(s == 3 ? 1.0 : rho) * Rcw[r][k]);
            Hx[a * 13 + k] = t;
        }
    // quaternion: q_cw = conj(q_wc); dRcw/dq_cw (A.46-A.49) applied to part2; d q_cw / d q_wc = diag(1, -1, -1, -1) (A.39)
    const double q[4] = {cam13[3], -cam13[4], -cam13[5], -cam13[6]};
    const double dR[4][3][3] = {
        {{2 * q[0], -2 * q[3], 2 * q[2]}, {2 * q[3], 2 * q[0], -2 * q[1]}, {-2 * q[2], 2 * q[1], 2 * q[0]}},
        {{2 * q[1], 2 * q[2], 2 * q[3]}, {2 * q[2], -2 * q[1], -2 * q[0]}, {2 * q[3], 2 * q[0], -2 * q[1]}},
        {{-2 * q[2], 2 * q[1], 2 * q[0]}, {2 * q[1], 2 * q[2], 2 * q[3]}, {-2 * q[0], 2 * q[3], -2 * q[2]}},
        {{-2 * q[3], -2 * q[0], 2 * q[1]}, {2 * q[0], -2 * q[3], 2 * q[2]}, {2 * q[1], 2 * q[2], 2 * q[3]}}};
    for (int qi = 0; qi < 4; ++qi) {
        double col[3];
        for (int r = 0; r < 3; ++r) col[r] = dR[qi][r][0] * part2[0] + dR[qi][r][1] * part2[1] + dR[qi][r][2] * part2[2];
        const double sign = qi == 0 ? 1.0 : -1.0;
        for (int a = 0; a < 2; ++a) Hx[a * 13 + 3 + qi] = (D[a][0] * col[0] + D[a][1] * col[1] + D[a][2] * col[2]) * sign;
    }
    // salient point: dhc_by_dy = Rcw (A.55) or [rho Rcw | Rcw dm/dtheta | Rcw dm/dphi | Rcw (first_cam_pos - cam_pos)] (A.52-A.54)
    double dy[3][6];
    if (s == 3) {
        for (int r = 0; r < 3; ++r) for (int k = 0; k < 3; ++k) dy[r][k] = Rcw[r][k];
    } else {
        const double cos_th = std::cos(sp[3]), sin_th = std::sin(sp[3]), cos_ph = std::cos(sp[4]), sin_ph = std::sin(sp[4]);
        const double dth[3] = {cos_ph * cos_th, 0.0, -cos_ph * sin_th}, dph[3] = {-sin_ph * sin_th, -cos_ph, -sin_ph * cos_th};
        const double dp[3] = {sp[0] - cam13[0], sp[1] - cam13[1], sp[2] - cam13[2]};
        for (int r = 0; r < 3; ++r) {
            for (int k = 0; k < 3; ++k) dy[r][k] = rho * Rcw[r][k];
            dy[r][3] = Rcw[r][0] * dth[0] + Rcw[r][1] * dth[1] + Rcw[r][2] * dth[2];
            dy[r][4] = Rcw[r][0] * dph[0] + Rcw[r][1] * dph[1] + Rcw[r][2] * dph[2];
            dy[r][5] = Rcw[r][0] * dp[0] + Rcw[r][1] * dp[1] + Rcw[r][2] * dp[2];
        }
    }
    for (int a = 0; a < 2; ++a)
        for (int k = 0; k < s; ++k) Hy[a * s + k] = D[a][0] * dy[0][k] + D[a][1] * dy[1][k] + D[a][2] * dy[2][k];
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
