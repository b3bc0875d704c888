// TEST INFRASTRUCTURE ONLY — CPU oracle: state and covariance of a NEW salient point of suriko-engine's MonoSLAM
// ("EKF.cpp" = /root/reference/cpp_impl/suriko-engine/src/davison-mono-slam.cpp).
//   EkfNewSalientPoint      GetNewSphericalSalientPointState :2398-2455 (undistortion A.58 back-projection, azimuth / elevation),
//                           the small Jacobians of GetNewSphericalSalientPointCovar :2457-2527 (A.67-A.79: sal_pnt_by_cam [6 x 7],
//                           sal_pnt_by_h_rho [6 x 3]) and its P-independent auto-covariance term :2535-2539,
//                           ConvertXyzFromSphericalSalientPoint :405-416, DerivSalPnt_xyz_by_spher :3790-3828 for the XYZ representation
//   EkfAddSalientPoint      AllocateAndInitStateForNewSalientPoint :2322-2396 -- ONE point appended: the P-dependent products
//                           (:2528-2534), the XYZ conversion (:2579-2592), conservativeResize + the three block assignments (:2373-2395)
// Parity unpinned (the reference has no test for this path); the analytic Jacobians are checked against central differences of the
// state function in tests/test_cpu_ekf_newpoint.py.
#pragma once
#include <cmath>
#include <vector>
#include "srk_oracle_ekf.hpp"
#include "srk_oracle_ekf_ransac.hpp"

namespace srk_oracle {

struct EkfNewPoint {
    double spher[6];        // first camera position (3), azimuth theta, elevation phi, inverse distance rho
    double Jy6[6 * 7];      // d spher / d (camera position, quaternion), row-major
    double Q6[6 * 6];       // sal_pnt_by_h_rho [R 0; 0 rho_var] sal_pnt_by_h_rho^T, row-major
    double xyz[3];
    double Jy3[3 * 7], Q3[3 * 3];
    int xyz_ok;
};

// EKF.cpp:2398-2527 (+ :405-416, :3790-3828).  cam13 = camera state (position, quaternion wfc, velocities).
inline void EkfNewSalientPoint(const EkfCamera& c, const double* cam13, const double corner_pix[2], double inv_dist, double inv_dist_std, double meas_std_pix,
                               EkfNewPoint* out) {
    const double* q = cam13 + 3;
    // undistort (:2404-2417)
    double hu[2] = {corner_pix[0], corner_pix[1]};
    double hu_by_hd[2][2] = {{1, 0}, {0, 1}};
    if (c.enable_distortion) {
        const double rd = std::sqrt((c.dx_mm * (corner_pix[0] - c.cx)) * (c.dx_mm * (corner_pix[0] - c.cx)) + (c.dy_mm * (corner_pix[1] - c.cy)) * (c.dy_mm * (corner_pix[1] - c.cy)));   // A.24
        const double stretch = 1 + c.k1 * (rd * rd) + c.k2 * ((rd * rd) * (rd * rd));
        hu[0] = c.cx + (corner_pix[0] - c.cx) * stretch;
        hu[1] = c.cy + (corner_pix[1] - c.cy) * stretch;
        // Deriv_hu_by_hd (:2650-2679), A.32
        const double kk = c.k1 + 2 * c.k2 * (rd * rd);
        const double side = 2 * kk * (corner_pix[1] - c.cy) * (corner_pix[0] - c.cx);
        hu_by_hd[0][0] = stretch + 2 * kk * ((c.dx_mm * (corner_pix[0] - c.cx)) * (c.dx_mm * (corner_pix[0] - c.cx)));
        hu_by_hd[1][1] = stretch + 2 * kk * ((c.dy_mm * (corner_pix[1] - c.cy)) * (c.dy_mm * (corner_pix[1] - c.cy)));
        hu_by_hd[1][0] = side * (c.dx_mm * c.dx_mm);
        hu_by_hd[0][1] = side * (c.dy_mm * c.dy_mm);
    }
    // A.58 (:2309-2320)
    const double hc[3] = {-(hu[0] - c.cx) / c.fx_pix, -(hu[1] - c.cy) / c.fy_pix, 1.0};
    double R[3][3];
    EkfRotMatFromQuat(q, R);
    double hw[3];
    for (int i = 0; i < 3; ++i) hw[i] = R[i][0] * hc[0] + R[i][1] * hc[1] + R[i][2] * hc[2];
    // AzimElevFromEuclidCoords (:399-403)
    const double theta = std::atan2(hw[0], hw[2]);
    const double phi = std::atan2(-hw[1], std::sqrt(hw[0] * hw[0] + hw[2] * hw[2]));
    out->spher[0] = cam13[0]; out->spher[1] = cam13[1]; out->spher[2] = cam13[2];
    out->spher[3] = theta; out->spher[4] = phi; out->spher[5] = inv_dist;
    // Deriv_R_by_q (:2716-2747), A.46-A.49
    const double dR[4][3][3] = {
        {{2 * q[0], -2 * q[3], 2 * q[2]}, {2 * q[3], 2 * q[0], -2 * q[1]}, {-2 * q[2], 2 * q[1], 2 * q[0]}},
        {{2 * q[1], 2 * q[2], 2 * q[3]}, {2 * q[2], -2 * q[1], -2 * q[0]}, {2 * q[3], 2 * q[0], -2 * q[1]}},
        {{-2 * q[2], 2 * q[1], 2 * q[0]}, {2 * q[1], 2 * q[2], 2 * q[3]}, {-2 * q[0], 2 * q[3], -2 * q[2]}},
        {{-2 * q[3], -2 * q[0], 2 * q[1]}, {2 * q[0], -2 * q[3], 2 * q[2]}, {2 * q[1], 2 * q[2], 2 * q[3]}}};
    double hw_by_q[3][4];   // A.73
    for (int k = 0; k < 4; ++k) for (int i = 0; i < 3; ++i) hw_by_q[i][k] = dR[k][i][0] * hc[0] + dR[k][i][1] * hc[1] + dR[k][i][2] * hc[2];
    // Deriv_azim_theta_elev_phi_by_hw (:2867-2885)
    const double dxz2 = hw[0] * hw[0] + hw[2] * hw[2];
    const double th_by_hw[3] = {hw[2] / dxz2, 0.0, -hw[0] / dxz2};
    const double d2 = dxz2 + hw[1] * hw[1], dxz = std::sqrt(dxz2), sf = hw[1] / (d2 * dxz);
    const double ph_by_hw[3] = {hw[0] * sf, -dxz / d2, hw[2] * sf};
    // sal_pnt_by_cam (:2483-2512): [I3 | 0] ; rows 3, 4 = angles by quaternion ; row 5 = 0
    for (int i = 0; i < 42; ++i) out->Jy6[i] = 0.0;
    for (int i = 0; i < 3; ++i) out->Jy6[i * 7 + i] = 1.0;
    for (int k = 0; k < 4; ++k) {
        out->Jy6[3 * 7 + 3 + k] = th_by_hw[0] * hw_by_q[0][k] + th_by_hw[1] * hw_by_q[1][k] + th_by_hw[2] * hw_by_q[2][k];
        out->Jy6[4 * 7 + 3 + k] = ph_by_hw[0] * hw_by_q[0][k] + ph_by_hw[1] * hw_by_q[1][k] + ph_by_hw[2] * hw_by_q[2][k];
    }
    // sal_pnt_by_h_rho (:2536-2546): top-left [5 x 2] = sal_pnt_by_hw * hw_by_hc (= Rwfc) * hc_by_hu (A.79) * hu_by_hd ; (5, 2) = 1
    double hc_by_hd[3][2];   // hc_by_hu * hu_by_hd, hc_by_hu = diag(-1/fx, -1/fy) over a zero row
    for (int j = 0; j < 2; ++j) { hc_by_hd[0][j] = (-1 / c.fx_pix) * hu_by_hd[0][j]; hc_by_hd[1][j] = (-1 / c.fy_pix) * hu_by_hd[1][j]; hc_by_hd[2][j] = 0.0; }
    double hw_by_hd[3][2];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 2; ++j) hw_by_hd[i][j] = R[i][0] * hc_by_hd[0][j] + R[i][1] * hc_by_hd[1][j] + R[i][2] * hc_by_hd[2][j];
    double A[6][3];
    for (int i = 0; i < 6; ++i) for (int j = 0; j < 3; ++j) A[i][j] = 0.0;
    for (int j = 0; j < 2; ++j) {
        A[3][j] = th_by_hw[0] * hw_by_hd[0][j] + th_by_hw[1] * hw_by_hd[1][j] + th_by_hw[2] * hw_by_hd[2][j];
        A[4][j] = ph_by_hw[0] * hw_by_hd[0][j] + ph_by_hw[1] * hw_by_hd[1][j] + ph_by_hw[2] * hw_by_hd[2][j];
    }
    A[5][2] = 1.0;
    const double meas_var = meas_std_pix * meas_std_pix, rho_var = inv_dist_std * inv_dist_std;
    for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j)
        out->Q6[i * 6 + j] = (A[i][0] * meas_var * A[j][0] + A[i][1] * meas_var * A[j][1]) + A[i][2] * rho_var * A[j][2];
    // XYZ representation: ConvertXyzFromSphericalSalientPoint (:405-416), DerivSalPnt_xyz_by_spher (:3790-3828)
    out->xyz_ok = IsClose<double>(0.0, inv_dist) ? 0 : 1;
    const double ct = std::cos(theta), st = std::sin(theta), cp = std::cos(phi), sp = std::sin(phi);
    const double dist = 1 / inv_dist, dist2 = 1 / (inv_dist * inv_dist);
    const double m[3] = {cp * st, -sp, cp * ct};
    for (int i = 0; i < 3; ++i) out->xyz[i] = cam13[i] + (1 / inv_dist) * m[i];
    double D[3][6] = {{1, 0, 0, dist * cp * ct, -dist * sp * st, -dist2 * cp * st},
                      {0, 1, 0, 0.0, -dist * cp, dist2 * sp},
                      {0, 0, 1, -dist * cp * st, -dist * sp * ct, -dist2 * cp * ct}};
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 7; ++j) { double s = 0; for (int k = 0; k < 6; ++k) s += D[i][k] * out->Jy6[k * 7 + j]; out->Jy3[i * 7 + j] = s; }
    double DQ[3][6];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 6; ++j) { double s = 0; for (int k = 0; k < 6; ++k) s += D[i][k] * out->Q6[k * 6 + j]; DQ[i][j] = s; }
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { double s = 0; for (int k = 0; k < 6; ++k) s += DQ[i][k] * D[j][k]; out->Q3[i * 3 + j] = s; }
}

// EKF.cpp:2322-2396 for ONE new point with s components: the state and the covariance grow by s (conservativeResize keeps the old block),
//   bottom-left = Jy * P[0:7, 0:n]  (:2528-2529; times deriv_xyz_by_spher for s = 3, already folded into Jy),  top-right = its transpose,
//   bottom-right = bottom-left[:, 0:7] * Jy^T + Qnew  (:2548-2551).   diag_only: force_xyz_sal_pnt_pos_diagonal_uncert_ (:2579-2584).
inline void EkfAddSalientPoint(std::vector<double>* x, EkfMat* P, int s, const double* x_new, const double* Jy, const double* Qnew, bool diag_only) {
    const size_t n = P->rows, n2 = n + (size_t)s;
    EkfMat Pn(n2, n2);
    for (size_t c2 = 0; c2 < n; ++c2) for (size_t r = 0; r < n; ++r) Pn(r, c2) = (*P)(r, c2);
    EkfMat BL((size_t)s, n);
    if (!diag_only)
        for (int a = 0; a < s; ++a) for (size_t c2 = 0; c2 < n; ++c2) { double v = 0; for (int q = 0; q < 7; ++q) v += Jy[a * 7 + q] * (*P)((size_t)q, c2); BL((size_t)a, c2) = v; }
    for (int a = 0; a < s; ++a) for (size_t c2 = 0; c2 < n; ++c2) { Pn(n + (size_t)a, c2) = BL((size_t)a, c2); Pn(c2, n + (size_t)a) = BL((size_t)a, c2); }
    for (int a = 0; a < s; ++a) for (int b = 0; b < s; ++b) {
        double v = 0;
        if (!diag_only) for (int q = 0; q < 7; ++q) v += BL((size_t)a, (size_t)q) * Jy[b * 7 + q];
        Pn(n + (size_t)a, n + (size_t)b) = v + Qnew[a * s + b];
    }
    *P = Pn;
    for (int a = 0; a < s; ++a) x->push_back(x_new[a]);
}

}  // namespace srk_oracle
