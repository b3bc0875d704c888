/*
 * suriko-b200 — C ABI of the MonoSLAM EKF dense covariance chain (north_star kernel 4, second half).
 *
 * Replaces the Eigen dense algebra of suriko-engine's Davison/Civera MonoSLAM ("EKF.cpp" =
 * /root/reference/cpp_impl/suriko-engine/src/davison-mono-slam.cpp):
 *   srk_ekf_predict*  : PredictEstimVars, covariance part                             EKF.cpp:669-693
 *   srk_ekf_update*   : ProcessFrame_StackedObservationsPerUpdateCore                 EKF.cpp:977-1125
 *                       (+ NormalizeCameraOrientationQuaternionAndCovariances :1652-1711, FixSymmetricMat :4308,
 *                          EnsureNonnegativeStateVariance :1739-1750)
 * The measurement Jacobian arrives in its sparse form: every observation row touches the 13 camera-state columns and the s
 * columns of its own salient point (EKF.cpp:3115-3159); the reference multiplies the dense [2m x n] matrix instead.
 * State layout (EKF.h): x = [cam pos 3 | cam quaternion 4 | velocity 3 | angular velocity 3 | s components per salient point ...],
 * P is n x n column-major (Eigen default), n = 13 + s * salient points.
 *
 * Algebra used on the device (equal to the reference's in exact arithmetic):
 *   PHt = P*H^T (sparse H) ; S = H*PHt + meas_var*I ; S = L*L^T ; Z = PHt*L^-T ; x += Z*(L^-1 (z - h)) ; P -= Z*Z^T.
 * All functions return 0 on success or a negative SRK_E_* code (ba_c_api.h); srk_last_error() has the text.
 */
#ifndef SRK_EKF_C_API_H
#define SRK_EKF_C_API_H

#include <stdint.h>
#include "ba_c_api.h"

#ifdef __cplusplus
extern "C" {
#endif

SRK_API int srk_ekf_create(void** h, int device);
SRK_API void srk_ekf_destroy(void* h);
SRK_API int srk_ekf_set_stream(void* h, void* cuda_stream);

/* Resident state: covariance P [n*n] column-major and state x [n] live in HBM between calls. */
SRK_API int srk_ekf_set_state(void* h, int64_t n, const double* P, const double* x);
SRK_API int srk_ekf_get_state(void* h, double* P, double* x);

/* Predict on the resident state: Pvv <- F*Pvv*F^T + GQGt, Pvm <- F*Pvm, Pmv <- Pvm^T (EKF.cpp:669-690); x[0..13) <- cam_state_new
 * (the kinematic model itself, EKF.cpp:583-637, is 13 scalars of host code and stays with the caller).  F13, GQGt13: 13x13 column-major. */
SRK_API int srk_ekf_predict_resident(void* h, const double* F13, const double* GQGt13, const double* cam_state_new);

/* Stacked update on the resident state.  Hcam [2m x 13] and Hpt [2m x s] row-major per observation row, pt_off[m] = state index of
 * the observed salient point, z / h_pred [2m] measured and projected pixels, R = meas_var * I (EKF.cpp:563-581).
 * info (optional): 0 ok, >0 = 1-based index of the first non-positive pivot of the innovation covariance. */
SRK_API int srk_ekf_update_resident(void* h, int64_t m, const double* Hcam, const double* Hpt, const int64_t* pt_off, int32_t s, const double* z,
                                    const double* h_pred, double meas_var, int32_t* info);

/* What ProjectCameraSalientPoint / DistortPixel read (EKF.cpp:3007-3033, :2960-3005): CameraIntrinsicParams::FocalLengthPix(),
 * principal_point_pix, pixel_size_mm, MikhailRadialDistortionParams k1 / k2, cam_enable_distortion_. */
typedef struct srk_ekf_camera {
    double fx_pix, fy_pix, cx, cy, dx_mm, dy_mm, k1, k2;
    int32_t enable_distortion;
} srk_ekf_camera;

/* Measurement Jacobians of the matched points at the resident state: Deriv_hd_by_cam_state_and_sal_pnt batched over points
 * (EKF.cpp:3067-3159; the reference fills a dense, zero-initialised [2m x n] H point by point before every update).  Outputs in the
 * layout srk_ekf_update_resident / srk_ekf_ransac_consensus_resident take: Hcam [2m x 13] (velocity columns zero), Hpt [2m x s],
 * h_pred [2m] = ProjectInternalSalientPoint (EKF.cpp:2947-2958). */
SRK_API int srk_ekf_measurement_jacobians_resident(void* h, int64_t m, const int64_t* pt_off, int32_t s, const srk_ekf_camera* camera, double* Hcam, double* Hpt,
                                                   double* h_pred);

/* Hypothesis scoring of the 1-point RANSAC update on the resident state: OnePointRansac_GetConsensusMatches (EKF.cpp:1271-1391), every
 * matched point as a hypothesis, all at once.  Inputs as srk_ekf_update_resident (the projections are formed on the device from the
 * state: no h_pred).  support[m] (optional) = size of each hypothesis' consensus set (matched points whose projection under the
 * hypothesis state lies within max_divergence_pix of their corner), *best = the winning hypothesis (the earliest maximum, -1 when no
 * support at all: the reference replaces its best only on strictly more support, :1383), best_inliers[m] (optional) its 0/1 mask =
 * low_innov_inliers.  The state is not modified. */
SRK_API int srk_ekf_ransac_consensus_resident(void* h, int64_t m, const double* Hcam, const double* Hpt, const int64_t* pt_off, int32_t s, const double* z,
                                              double meas_var, const srk_ekf_camera* camera, double max_divergence_pix, int32_t* support, int32_t* best,
                                              unsigned char* best_inliers);

/* Projected 2-D covariance of every listed point at the resident state, J P_in J^T over the camera position / quaternion and the point's own
 * variables (GetSalientPointProjected2DPosWithUncertainty, EKF.cpp:3901-4025; no measurement noise is added there).  cov [m][4] row-major,
 * symmetrised like FixAlmostSymmetricMat.  The second stage of the 1-point RANSAC update (EKF.cpp:1466-1498) is this + a chi^2 test; the whole
 * update -- consensus, stacked update of the low-innovation inliers, rescue, second stacked update -- is `surikatoko_b200/ekf.py::
 * one_point_ransac_update` on top of the resident entry points (ProcessFrame_OnePointRansacUpdateCore :1393-1513 is host control flow). */
SRK_API int srk_ekf_projected_covariances_resident(void* h, int64_t m, const double* Hcam, const double* Hpt, const int64_t* pt_off, int32_t s, double* cov);

/* The per-observation update variants on the resident state: ProcessFrame_OneObservationPerUpdate (EKF.cpp:1153-1269; per_component = 0) and
 * ProcessFrame_OneComponentOfOneObservationPerUpdate (:1525-1650; per_component = 1).  The measurement Jacobian of every point is re-derived on
 * the device at the latest state, so the call is a dependent chain of m (2m) rank-2 (rank-1) updates of the whole covariance: 16 n^2 bytes of
 * HBM traffic each -- use the stacked update for throughput; this is the reference's alternative semantics (relinearisation per observation). */
SRK_API int srk_ekf_sequential_update_resident(void* h, int64_t m, const int64_t* pt_off, int32_t s, const double* z, const srk_ekf_camera* camera, double meas_var,
                                               int32_t per_component);

/* Covariance growth for k new salient points on the resident state: the dense part of AllocateAndInitStateForNewSalientPoint
 * (EKF.cpp:2322-2396; the reference calls conservativeResize -- a temporary and a full copy of P -- once per point):
 *     x <- [x ; x_new],   P[new_i, old] = Jy_i P[0:7, old],   P[new_i, new_l] = Jy_i P[0:7, 0:7] Jy_l^T (+ Qnew_i when i == l)
 * x_new [k*s]; Jy [k][s][7] row-major = d(new point) / d(camera position, quaternion) (sal_pnt_by_cam, EKF.cpp:2483-2512; for s = 3 times
 * deriv_sal_pnt_xyz_by_spher, :2586-2592); Qnew [k][s][s] row-major = the part of the auto-covariance that does not come from P (pixel
 * noise and initial inverse-distance variance, :2535-2539, conjugated likewise).  diag_only != 0: force_xyz_sal_pnt_pos_diagonal_uncert_
 * (:2579-2584) -- zero cross covariance, auto-covariance = Qnew.  surikatoko_b200/ekf.py::new_salient_point forms x_new / Jy / Qnew from a
 * corner pixel the way GetNewSphericalSalientPointState / ...Covar do (13 scalars of host code per point, like the kinematic model). */
SRK_API int srk_ekf_add_points_resident(void* h, int64_t k, int32_t s, const double* x_new, const double* Jy, const double* Qnew, int32_t diag_only);
SRK_API int srk_ekf_state_size(void* h, int64_t* n);

/* One-shot forms with host buffers (upload, operate, download). */
SRK_API int srk_ekf_update(void* h, int64_t n, int64_t m, double* P, double* x, const double* Hcam, const double* Hpt, const int64_t* pt_off, int32_t s,
                           const double* z, const double* h_pred, double meas_var);
SRK_API int srk_ekf_predict(void* h, int64_t n, double* P, const double* F13, const double* GQGt13);

/* Kernel-family timing (CUDA events on the handle's stream): "pht", "innov", "chol", "trsm", "syrk", "state", "predict", "ransac". */
SRK_API int srk_ekf_set_timing(void* h, int enabled);
SRK_API int srk_ekf_get_timing(void* h, const char* name, double* ms_total, int64_t* count);
SRK_API int64_t srk_ekf_launches(void* h);

#ifdef __cplusplus
}
#endif
#endif /* SRK_EKF_C_API_H */
