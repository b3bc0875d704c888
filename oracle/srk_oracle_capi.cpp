// TEST INFRASTRUCTURE ONLY — C entry points of the CPU oracle for ctypes (tests/, bench.py cpu_baseline,
// __graft_entry__.smoke()).  The product path never loads this library.
//
// The flat problem layout is the one of include/srk/ba_c_api.h (observations sorted by (pnt_ind, frame_ind));
// here it is expanded back into FragmentMap / CornerTrackRepository mirrors so that the oracle walks them with
// the reference's own GetCorner probing loops (BA.cpp:430-484, :1160-1412).
#include <chrono>
#include <cstring>
#include "srk_oracle_ba.hpp"
#include "srk_oracle_scene.hpp"
#include "srk_oracle_ekf.hpp"
#include "srk_oracle_ekf_exact.hpp"
#include "srk_oracle_ekf_ransac.hpp"
#include "srk_oracle_ekf_newpoint.hpp"
#include "srk_oracle_ekf_ransac_update.hpp"
#include "srk_oracle_ekf_sequential.hpp"

using namespace srk_oracle;

namespace {

struct Mirrors {
    FragmentMap<double> map;
    std::vector<SE3<double>> cams;
    std::vector<Mat33<double>> Ks;
    Mat33<double> sharedK;
    bool shared = false;
    CornerTrackRepository<double> tracks;
    std::vector<size_t> ids;  // pnt_ind -> SalientPointId
};

void Expand(int64_t n_cams, int64_t n_points, int64_t n_obs, const int32_t* obs_cam, const int32_t* obs_point, const double* obs_xy,
            const double* points, const double* cams, const double* K, int shared_K, Mirrors* m) {
    for (int64_t p = 0; p < n_points; ++p) {
        size_t id = 0;
        m->map.AddSalientPointTempl(Vec3<double>(points[3 * p], points[3 * p + 1], points[3 * p + 2]), &id);
        m->ids.push_back(id);
        auto& t = m->tracks.AddCornerTrackObj();
        t.SalientPointId = id;
    }
    for (int64_t o = 0; o < n_obs; ++o) {
        auto& cd = m->tracks.GetPointTrackById((size_t)obs_point[o]).AddCorner((size_t)obs_cam[o]);
        cd.pixel_coord = Point2<double>(obs_xy[2 * o], obs_xy[2 * o + 1]);
    }
    m->cams.resize((size_t)n_cams);
    static_assert(sizeof(SE3<double>) == 12 * sizeof(double), "SE3 must be 12 doubles: T[3], R col-major[9]");
    std::memcpy(m->cams.data(), cams, sizeof(double) * 12 * (size_t)n_cams);
    m->shared = shared_K != 0;
    if (m->shared) std::memcpy(m->sharedK.a, K, sizeof(double) * 9);
    else { m->Ks.resize((size_t)n_cams); std::memcpy(m->Ks.data(), K, sizeof(double) * 9 * (size_t)n_cams); }
}

void Collapse(const Mirrors& m, double* points, double* cams) {
    for (size_t p = 0; p < m.ids.size(); ++p) {
        const auto& x = m.map.GetSalientPoint(m.ids[p]);
        points[3 * p] = x[0]; points[3 * p + 1] = x[1]; points[3 * p + 2] = x[2];
    }
    std::memcpy(cams, m.cams.data(), sizeof(double) * 12 * m.cams.size());
}

template <class BA>
void Configure(BA& ba, int flow, int solve_impl, int unity_ind, double unity_val, int max_outer_iters) {
    ba.schur_flow = flow == 0 ? SchurFlow::DenseReference : flow == 2 ? SchurFlow::SparseThreaded : SchurFlow::SparseEquivalent;
    ba.solve_impl = solve_impl == 0 ? SolveImpl::HouseholderQR : solve_impl == 2 ? SolveImpl::None : SolveImpl::CholeskyRefined;
    if (solve_impl >= 10) { ba.chol_in_double = true; ba.refine_steps = solve_impl - 10; }  // 10+k: double LL^T with k refinement steps
    ba.unity_t1_comp_ind_ = (size_t)unity_ind;
    ba.unity_t1_comp_value_ = unity_val;
    ba.max_outer_iters = (size_t)max_outer_iters;
}

}  // namespace

#include "srk_oracle_frontend.hpp"

extern "C" {

struct srk_oracle_report {
    int32_t converged, stop_reason, outer_iters, n_attempts;
    double err_initial, err_final, hessian_factor_final;
    int64_t seen_points;
    double world_scale;
    double seconds;
};

// Full ComputeInplace.  err_trace[cap] receives the accepted error of each outer iteration; attempts[4*cap_att]
// receives (hessian_factor, err_new, accepted, skipped_points) per attempt.  acc: 0 = double (faithful), 1 = long double Schur ("exact").
int srk_oracle_ba_solve(int64_t n_cams, int64_t n_points, int64_t n_obs, const int32_t* obs_cam, const int32_t* obs_point,
                        const double* obs_xy, double* points, double* cams, const double* K, int shared_K, double f0,
                        int has_err_change, double err_change, int has_max_hf, double max_hf, int unity_ind, double unity_val,
                        int max_outer_iters, int flow, int solve_impl, int acc, srk_oracle_report* rep, double* err_trace, int err_trace_cap,
                        double* attempts, int attempts_cap) {
    try {
        Mirrors m;
        Expand(n_cams, n_points, n_obs, obs_cam, obs_point, obs_xy, points, cams, K, shared_K, &m);
        TermCriteria<double> tc;
        if (has_err_change) tc.allowed_reproj_err_rel_change = err_change;
        if (has_max_hf) tc.max_hessian_factor = max_hf;
        auto run = [&](auto& ba) {
            Configure(ba, flow, solve_impl, unity_ind, unity_val, max_outer_iters);
            auto t0 = std::chrono::steady_clock::now();
            bool ok = ba.ComputeInplace(f0, m.map, m.cams, m.tracks, m.shared ? &m.sharedK : nullptr, m.shared ? nullptr : &m.Ks, tc);
            auto t1 = std::chrono::steady_clock::now();
            const auto& tr = ba.trace;
            rep->converged = ok ? 1 : 0;
            rep->stop_reason = tr.stop_reason;
            rep->outer_iters = (int32_t)tr.outer_iters;
            rep->n_attempts = (int32_t)tr.attempts.size();
            rep->err_initial = tr.err_initial;
            rep->err_final = tr.err_per_iter.empty() ? tr.err_initial : tr.err_per_iter.back();
            rep->hessian_factor_final = tr.attempts.empty() ? 0.0001f : tr.attempts.back().hessian_factor;
            rep->seen_points = (int64_t)tr.seen_points;
            rep->world_scale = ba.WorldScale();
            rep->seconds = std::chrono::duration<double>(t1 - t0).count();
            for (size_t i = 0; i < tr.err_per_iter.size() && (int)i < err_trace_cap; ++i) err_trace[i] = tr.err_per_iter[i];
            for (size_t i = 0; i < tr.attempts.size() && (int)i < attempts_cap; ++i) {
                attempts[4 * i] = tr.attempts[i].hessian_factor; attempts[4 * i + 1] = tr.attempts[i].err_new;
                attempts[4 * i + 2] = tr.attempts[i].accepted; attempts[4 * i + 3] = (double)tr.attempts[i].skipped_points;
            }
        };
        if (acc == 0) { BundleAdjustmentKanatani<double, double> ba; run(ba); }
        else { BundleAdjustmentKanatani<double, long double> ba; run(ba); }
        Collapse(m, points, cams);
        return 0;
    } catch (const std::exception&) { return -1; }
}

// BundleAdjustmentKanatani::ReprojError (static, no normalisation) — BA.cpp:589-600.
int srk_oracle_reproj_error(int64_t n_cams, int64_t n_points, int64_t n_obs, const int32_t* obs_cam, const int32_t* obs_point,
                            const double* obs_xy, const double* points, const double* cams, const double* K, int shared_K, double f0,
                            double* err, int64_t* seen_points) {
    try {
        Mirrors m;
        Expand(n_cams, n_points, n_obs, obs_cam, obs_point, obs_xy, points, cams, K, shared_K, &m);
        size_t seen = 0;
        *err = BundleAdjustmentKanatani<double>::ReprojError(f0, m.map, m.cams, m.tracks, m.shared ? &m.sharedK : nullptr, m.shared ? nullptr : &m.Ks, &seen);
        *seen_points = (int64_t)seen;
        return 0;
    } catch (const std::exception&) { return -1; }
}

// NormalizeSceneInplace / RevertNormalization (BA.cpp:203-286) on flat buffers.  revert != 0 applies the inverse
// using (cam0_prenorm[12], world_scale) returned by a previous normalise call.
int srk_oracle_normalize(int64_t n_cams, int64_t n_points, double* points, double* cams, int unity_ind, double unity_val,
                         int revert, double* cam0_prenorm, double* world_scale) {
    FragmentMap<double> map;
    for (int64_t p = 0; p < n_points; ++p) map.AddSalientPointTempl(Vec3<double>(points[3 * p], points[3 * p + 1], points[3 * p + 2]));
    std::vector<SE3<double>> cs((size_t)n_cams);
    std::memcpy(cs.data(), cams, sizeof(double) * 12 * (size_t)n_cams);
    CornerTrackRepository<double> tr;
    BundleAdjustmentKanatani<double> ba;
    ba.map_ = &map; ba.inverse_orient_cams_ = &cs; ba.track_rep_ = &tr;
    ba.unity_t1_comp_ind_ = (size_t)unity_ind; ba.unity_t1_comp_value_ = unity_val;
    int rc = 0;
    if (!revert) {
        if (!ba.NormalizeWorldInplace()) rc = 1;
        else { std::memcpy(cam0_prenorm, &ba.prenorm_cam0_from_world_, sizeof(double) * 12); *world_scale = ba.world_scale_; }
    } else {
        std::memcpy(&ba.prenorm_cam0_from_world_, cam0_prenorm, sizeof(double) * 12);
        ba.world_scale_ = *world_scale;
        ba.RevertNormalization();
    }
    if (rc == 0) {
        for (int64_t p = 0; p < n_points; ++p) { const auto& x = map.SalientPoints()[(size_t)p].coord.value(); points[3 * p] = x[0]; points[3 * p + 1] = x[1]; points[3 * p + 2] = x[2]; }
        std::memcpy(cams, cs.data(), sizeof(double) * 12 * (size_t)n_cams);
    }
    return rc;
}

// One derivative pass + one two-phase solve at damping c on the scene AS GIVEN (no normalisation is applied; pass an
// already-normalised scene).  Outputs (any may be null): gradE[3N+10M], E[9N] (per point row-major 3x3), G[100M] (per
// frame row-major 10x10), Fblk[30*n_obs] (per observation row-major 3x10, observation order of the input),
// S[n_f*n_f] col-major, rhs[n_f], skipped[N], corrections[3N+10M].  Returns 0, or 1 if the solve reported failure.
int srk_oracle_derivs_and_solve(int64_t n_cams, int64_t n_points, int64_t n_obs, const int32_t* obs_cam, const int32_t* obs_point,
                                const double* obs_xy, const double* points, const double* cams, const double* K, int shared_K, double f0,
                                int unity_ind, double c, int flow, int solve_impl, int acc,
                                double* gradE, double* E, double* G, double* Fblk, double* S, double* rhs, unsigned char* skipped,
                                double* corrections) {
    try {
        Mirrors m;
        Expand(n_cams, n_points, n_obs, obs_cam, obs_point, obs_xy, points, cams, K, shared_K, &m);
        int rc = 0;
        auto run = [&](auto& ba) {
            Configure(ba, flow, solve_impl, unity_ind, 1.0, 0);
            ba.Bind(f0, m.map, m.cams, m.tracks, m.shared ? &m.sharedK : nullptr, m.shared ? nullptr : &m.Ks);
            ba.EnsureMemoryAllocated();
            ba.ComputeCloseFormReprErrorDerivatives();
            size_t N = (size_t)n_points, M = (size_t)n_cams;
            if (gradE) for (size_t i = 0; i < ba.gradE_.size(); ++i) gradE[i] = ba.gradE_[i];
            if (E) for (size_t p = 0; p < N; ++p) for (size_t r = 0; r < 3; ++r) for (size_t cc = 0; cc < 3; ++cc) E[p * 9 + r * 3 + cc] = ba.E_(p * 3 + r, cc);
            if (G) for (size_t f = 0; f < M; ++f) for (size_t r = 0; r < 10; ++r) for (size_t cc = 0; cc < 10; ++cc) G[f * 100 + r * 10 + cc] = ba.G_(f * 10 + r, cc);
            if (Fblk) {
                for (size_t o = 0; o < ba.obs_.size(); ++o)
                    for (size_t r = 0; r < 3; ++r) for (size_t cc = 0; cc < 10; ++cc)
                        Fblk[o * 30 + r * 10 + cc] = flow == 0 ? ba.Fdense_(ba.obs_[o].pnt * 3 + r, ba.obs_[o].frame * 10 + cc) : ba.Fblk_[o * 30 + r * 10 + cc];
            }
            if (c >= 0) {
                bool ok = ba.EstimateCorrectionsDecomposedInTwoPhases(c);
                if (!ok) rc = 1;
                size_t nf = ba.S_.rows;
                if (S) for (size_t i = 0; i < nf * nf; ++i) S[i] = (double)ba.S_.d[i];
                if (rhs) for (size_t i = 0; i < nf; ++i) rhs[i] = (double)ba.rhs_[i];
                if (skipped) for (size_t p = 0; p < N; ++p) skipped[p] = ba.skipped_mask_[p];
                if (corrections) for (size_t i = 0; i < ba.corrections_.size(); ++i) corrections[i] = ba.corrections_[i];
            }
        };
        if (acc == 0) { BundleAdjustmentKanatani<double, double> ba; run(ba); }
        else { BundleAdjustmentKanatani<double, long double> ba; run(ba); }
        return rc;
    } catch (const std::exception&) { return -1; }
}

// EstimateCorrectionsNaive (BA.cpp:1700-1769 with FillHessian :1551-1598): the full (3N+10M)^2 damped system with the gauge rows and
// columns removed, solved by Householder QR -- the reference's own cross-check of the two-phase solve (compare_with_naive, :788-797).
// Dense flow, tiny scenes only.  corrections[3N+10M] with the gaps re-inserted.
int srk_oracle_naive_solve(int64_t n_cams, int64_t n_points, int64_t n_obs, const int32_t* obs_cam, const int32_t* obs_point,
                           const double* obs_xy, const double* points, const double* cams, const double* K, int shared_K, double f0,
                           int unity_ind, double c, double* corrections) {
    try {
        Mirrors m;
        Expand(n_cams, n_points, n_obs, obs_cam, obs_point, obs_xy, points, cams, K, shared_K, &m);
        BundleAdjustmentKanatani<double, double> ba;
        Configure(ba, 0, 0, unity_ind, 1.0, 0);
        ba.Bind(f0, m.map, m.cams, m.tracks, m.shared ? &m.sharedK : nullptr, m.shared ? nullptr : &m.Ks);
        ba.EnsureMemoryAllocated();
        ba.ComputeCloseFormReprErrorDerivatives();
        std::vector<double> out;
        if (!ba.EstimateCorrectionsNaive(c, &out)) return 1;
        for (size_t i = 0; i < out.size(); ++i) corrections[i] = out[i];
        return 0;
    } catch (const std::exception&) { return -1; }
}

// ApplyCorrections (BA.cpp:1997-2063) on flat buffers.
int srk_oracle_apply_corrections(int64_t n_cams, int64_t n_points, double* points, double* cams, const double* corrections) {
    FragmentMap<double> map;
    CornerTrackRepository<double> tr;
    for (int64_t p = 0; p < n_points; ++p) {
        size_t id = 0;
        map.AddSalientPointTempl(Vec3<double>(points[3 * p], points[3 * p + 1], points[3 * p + 2]), &id);
        tr.AddCornerTrackObj().SalientPointId = id;
    }
    std::vector<SE3<double>> cs((size_t)n_cams);
    std::memcpy(cs.data(), cams, sizeof(double) * 12 * (size_t)n_cams);
    BundleAdjustmentKanatani<double> ba;
    ba.map_ = &map; ba.inverse_orient_cams_ = &cs; ba.track_rep_ = &tr;
    ba.corrections_.assign(corrections, corrections + 3 * n_points + 10 * n_cams);
    ba.ApplyCorrections();
    for (int64_t p = 0; p < n_points; ++p) { const auto& x = map.SalientPoints()[(size_t)p].coord.value(); points[3 * p] = x[0]; points[3 * p + 1] = x[1]; points[3 * p + 2] = x[2]; }
    std::memcpy(cams, cs.data(), sizeof(double) * 12 * (size_t)n_cams);
    return 0;
}

// Circle-grid demo scene (demo-bundle-adj-circle-grid.cpp:64-257).  params[16]:
//  f0, xmin, xmax, ymin, ymax, zmin, zmax, cell_x, cell_y, ang_start, ang_end, ang_step, noise_R_hi, noise_x3D_hi, rot_radius, ascentZ
// Call with null outputs to query sizes.  All arrays are flat problem layout; K is per frame [9*M] col-major.
int srk_oracle_circle_grid_scene(const double* params, unsigned seed, int64_t* n_cams, int64_t* n_points, int64_t* n_obs,
                                 int32_t* obs_cam, int32_t* obs_point, double* obs_xy, double* points, double* cams, double* K,
                                 double* gt_points, double* gt_cams) {
    CircleGridParams<double> p;
    p.f0 = params[0]; p.world_xmin = params[1]; p.world_xmax = params[2]; p.world_ymin = params[3]; p.world_ymax = params[4];
    p.world_zmin = params[5]; p.world_zmax = params[6]; p.cell_x = params[7]; p.cell_y = params[8];
    p.ang_start = params[9]; p.ang_end = params[10]; p.ang_step = params[11]; p.noise_R_hi = params[12]; p.noise_x3D_hi = params[13];
    p.rot_radius = params[14]; p.ascentZ = params[15]; p.seed = seed;
    Scene<double> sc;
    CircleGridScene(p, &sc);
    size_t M = sc.cams.size(), N = sc.map.SalientPointsCount();
    *n_cams = (int64_t)M; *n_points = (int64_t)N;
    size_t o = 0;
    for (size_t t = 0; t < sc.tracks.CornerTracksCount(); ++t) {
        sc.tracks.GetPointTrackById(t).EachCorner([&](size_t frame, const std::optional<CornerData<double>>& cd) {
            if (!cd.has_value()) return;
            if (obs_cam) { obs_cam[o] = (int32_t)frame; obs_point[o] = (int32_t)t; obs_xy[2 * o] = cd.value().pixel_coord[0]; obs_xy[2 * o + 1] = cd.value().pixel_coord[1]; }
            ++o;
        });
    }
    *n_obs = (int64_t)o;
    if (points) for (size_t i = 0; i < N; ++i) for (int k = 0; k < 3; ++k) points[3 * i + k] = sc.map.SalientPoints()[i].coord.value()[k];
    if (gt_points) for (size_t i = 0; i < N; ++i) for (int k = 0; k < 3; ++k) gt_points[3 * i + k] = sc.gt_points[i][k];
    if (cams) std::memcpy(cams, sc.cams.data(), sizeof(double) * 12 * M);
    if (gt_cams) std::memcpy(gt_cams, sc.gt_cams.data(), sizeof(double) * 12 * M);
    if (K) std::memcpy(K, sc.Ks.data(), sizeof(double) * 9 * M);
    return 0;
}

// GenerateCircleCameraShots (scene-generator.cpp:9-55)
int srk_oracle_circle_camera_shots(const double* center, double radius, double ascentZ, const double* angles, int n, double* cams_out) {
    std::vector<double> a(angles, angles + n);
    std::vector<SE3<double>> cs;
    GenerateCircleCameraShots(Vec3<double>(center[0], center[1], center[2]), radius, ascentZ, a, &cs);
    std::memcpy(cams_out, cs.data(), sizeof(double) * 12 * cs.size());
    return 0;
}

// Rodrigues / Log helpers (obs-geom.cpp:520-604) for the known-answer tests of test-obs-geom.cpp.
int srk_oracle_rotmat_from_axis_angle(const double* w, double* R) {
    Mat33<double> r;
    bool ok = RotMatFromAxisAngle(Vec3<double>(w[0], w[1], w[2]), &r);
    if (ok) std::memcpy(R, r.a, sizeof(double) * 9);
    return ok ? 1 : 0;
}
int srk_oracle_rotmat_from_unity_dir_and_angle(const double* dir, double ang, double* R) {
    Mat33<double> r;
    bool ok = RotMatFromUnityDirAndAngle(Vec3<double>(dir[0], dir[1], dir[2]), ang, &r, true);
    if (ok) std::memcpy(R, r.a, sizeof(double) * 9);
    return ok ? 1 : 0;
}
int srk_oracle_axis_angle_from_rotmat(const double* R, double* w) {
    Mat33<double> r; std::memcpy(r.a, R, sizeof(double) * 9);
    Vec3<double> d;
    bool ok = AxisAngleFromRotMat(r, &d);
    if (ok) { w[0] = d[0]; w[1] = d[1]; w[2] = d[2]; }
    return ok ? 1 : 0;
}

// CornerTrack::AddCorner(frame, value) push_back semantics (quirk Q11): feeds (frame, x, y) triples through the
// push_back variant and reports what GetCorner returns for frames [0, n_frames).
int srk_oracle_track_pushback_probe(const int32_t* frames, const double* xy, int n, int n_frames, int32_t* has, double* out_xy) {
    CornerTrack<double> t;
    for (int i = 0; i < n; ++i) t.AddCorner((size_t)frames[i], Point2<double>(xy[2 * i], xy[2 * i + 1]));
    for (int f = 0; f < n_frames; ++f) {
        auto c = t.GetCorner((size_t)f);
        has[f] = c.has_value() ? 1 : 0;
        if (c.has_value()) { out_xy[2 * f] = c.value()[0]; out_xy[2 * f + 1] = c.value()[1]; }
    }
    return 0;
}

// MonoSLAM EKF dense chain.  H is given in the sparse form of the C ABI (Hcam [2m x 13] and Hpt [2m x s], both row-major per
// observation row, pt_off[m] = first state index of the observed point) and expanded into the dense [2m x n] matrix the reference
// multiplies.  P [n x n] column-major and x [n] are updated in place.  Returns 0, or 1 if S is singular.
int srk_oracle_ekf_update(int64_t n, int64_t m, double* P, double* x, const double* Hcam, const double* Hpt, const int64_t* pt_off, int s,
                          const double* z, const double* hpred, double meas_var, int fix_symmetry, double* seconds) {
    EkfMat Pm((size_t)n, (size_t)n);
    std::memcpy(Pm.d.data(), P, sizeof(double) * (size_t)n * (size_t)n);
    std::vector<double> xs(x, x + n), zs(z, z + 2 * m), hs(hpred, hpred + 2 * m);
    EkfMat H((size_t)(2 * m), (size_t)n);
    for (int64_t i = 0; i < m; ++i)
        for (int k = 0; k < 2; ++k) {
            size_t row = (size_t)(2 * i + k);
            for (int c = 0; c < 13; ++c) H(row, (size_t)c) = Hcam[row * 13 + c];
            for (int c = 0; c < s; ++c) H(row, (size_t)(pt_off[i] + c)) = Hpt[row * s + c];
        }
    auto t0 = std::chrono::steady_clock::now();
    bool ok = EkfStackedUpdate(&xs, &Pm, H, zs, hs, meas_var, fix_symmetry != 0);
    auto t1 = std::chrono::steady_clock::now();
    if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
    std::memcpy(P, Pm.d.data(), sizeof(double) * (size_t)n * (size_t)n);
    std::memcpy(x, xs.data(), sizeof(double) * (size_t)n);
    return ok ? 0 : 1;
}
// The same update evaluated in long double through the Cholesky form (srk_oracle_ekf_exact.hpp): the parity target at large n.
int srk_oracle_ekf_update_exact(int64_t n, int64_t m, double* P, double* x, const double* Hcam, const double* Hpt, const int64_t* pt_off, int s,
                                const double* z, const double* hpred, double meas_var, int fix_symmetry) {
    return EkfStackedUpdateExact(n, m, P, x, Hcam, Hpt, pt_off, s, z, hpred, meas_var, fix_symmetry != 0) ? 0 : 1;
}
// OnePointRansac_GetConsensusMatches (EKF.cpp:1271-1391).  cam9 = {fx_pix, fy_pix, cx, cy, dx_mm, dy_mm, k1, k2, enable_distortion}.
// Returns the winning hypothesis or -1.  project_only != 0: just fills hd_out[2m] with the projections at the given state.
int srk_oracle_ekf_ransac(int64_t n, int64_t m, const double* P, const double* x, const double* Hcam, const double* Hpt, const int64_t* pt_off, int s, const double* z,
                          double meas_var, const double* cam9, double max_divergence_pix, int32_t* support, unsigned char* best_inliers, double* hd_out) {
    EkfCamera cam{cam9[0], cam9[1], cam9[2], cam9[3], cam9[4], cam9[5], cam9[6], cam9[7], cam9[8] != 0.0 ? 1 : 0};
    std::vector<double> xs(x, x + n);
    if (hd_out != nullptr)
        for (int64_t i = 0; i < m; ++i) EkfProjectSalientPoint(cam, xs.data(), xs.data() + pt_off[i], s, hd_out + 2 * i);
    if (P == nullptr) return -1;
    EkfMat Pm((size_t)n, (size_t)n);
    std::memcpy(Pm.d.data(), P, sizeof(double) * (size_t)n * (size_t)n);
    return EkfRansacConsensus(xs, Pm, m, Hcam, Hpt, pt_off, s, z, meas_var, cam, max_divergence_pix, support, best_inliers);
}
// Deriv_H_by_estim_vars in the sparse form of the C ABI (EKF.cpp:3115-3159): Hcam [2m x 13], Hpt [2m x s], hd [2m]
int srk_oracle_ekf_jacobians(int64_t n, int64_t m, const double* x, const int64_t* pt_off, int s, const double* cam9, double* Hcam, double* Hpt, double* hd) {
    (void)n;
    EkfCamera cam{cam9[0], cam9[1], cam9[2], cam9[3], cam9[4], cam9[5], cam9[6], cam9[7], cam9[8] != 0.0 ? 1 : 0};
    for (int64_t i = 0; i < m; ++i) EkfMeasurementJacobian(cam, x, x + pt_off[i], s, Hcam + (size_t)(2 * i) * 13, Hpt + (size_t)(2 * i) * s, hd + 2 * i);
    return 0;
}
// ProcessFrame_OneObservationPerUpdate (EKF.cpp:1153-1269) / ...OneComponentOfOneObservationPerUpdate (:1525-1650) on copies.
int srk_oracle_ekf_sequential_update(int64_t n, int64_t m, double* P, double* x, const int64_t* pt_off, int s, const double* z, double meas_var, const double* cam9,
                                     int per_component) {
    EkfCamera cam{cam9[0], cam9[1], cam9[2], cam9[3], cam9[4], cam9[5], cam9[6], cam9[7], cam9[8] != 0.0 ? 1 : 0};
    EkfMat Pm((size_t)n, (size_t)n);
    std::memcpy(Pm.d.data(), P, sizeof(double) * (size_t)n * (size_t)n);
    std::vector<double> xs(x, x + n);
    EkfSequentialUpdate(&xs, &Pm, m, pt_off, s, z, meas_var, cam, per_component != 0);
    std::memcpy(P, Pm.d.data(), sizeof(double) * (size_t)n * (size_t)n);
    std::memcpy(x, xs.data(), sizeof(double) * (size_t)n);
    return 0;
}
// ProcessFrame_OnePointRansacUpdateCore (EKF.cpp:1393-1513) on copies: P [n x n] col-major and x [n] in / out, low[m] / high[m] masks, counts[2].
int srk_oracle_ekf_ransac_update(int64_t n, int64_t m, double* P, double* x, const int64_t* pt_off, int s, const double* z, double meas_var, const double* cam9,
                                 double max_divergence_pix, double chi2_thr, unsigned char* low, unsigned char* high, int64_t* counts) {
    EkfCamera cam{cam9[0], cam9[1], cam9[2], cam9[3], cam9[4], cam9[5], cam9[6], cam9[7], cam9[8] != 0.0 ? 1 : 0};
    EkfMat Pm((size_t)n, (size_t)n);
    std::memcpy(Pm.d.data(), P, sizeof(double) * (size_t)n * (size_t)n);
    std::vector<double> xs(x, x + n);
    EkfOnePointRansacUpdate(&xs, &Pm, m, pt_off, s, z, meas_var, cam, max_divergence_pix, chi2_thr, low, high, counts);
    std::memcpy(P, Pm.d.data(), sizeof(double) * (size_t)n * (size_t)n);
    std::memcpy(x, xs.data(), sizeof(double) * (size_t)n);
    return 0;
}
// State / small Jacobians of a new salient point (EKF.cpp:2398-2527).  out55x = spher[6], Jy6[42], Q6[36], xyz[3], Jy3[21], Q3[9] in that order; returns xyz_ok.
int srk_oracle_ekf_new_point(const double* cam9, const double* cam13, const double* corner_pix, double inv_dist, double inv_dist_std, double meas_std_pix, double* out117) {
    EkfCamera cam{cam9[0], cam9[1], cam9[2], cam9[3], cam9[4], cam9[5], cam9[6], cam9[7], cam9[8] != 0.0 ? 1 : 0};
    EkfNewPoint np;
    EkfNewSalientPoint(cam, cam13, corner_pix, inv_dist, inv_dist_std, meas_std_pix, &np);
    double* o = out117;
    std::memcpy(o, np.spher, sizeof(np.spher)); o += 6;
    std::memcpy(o, np.Jy6, sizeof(np.Jy6)); o += 42;
    std::memcpy(o, np.Q6, sizeof(np.Q6)); o += 36;
    std::memcpy(o, np.xyz, sizeof(np.xyz)); o += 3;
    std::memcpy(o, np.Jy3, sizeof(np.Jy3)); o += 21;
    std::memcpy(o, np.Q3, sizeof(np.Q3));
    return np.xyz_ok;
}
// AllocateAndInitStateForNewSalientPoint (EKF.cpp:2322-2396), k points appended ONE AFTER THE OTHER as the reference does.
// P [n x n] col-major in, Pout [(n + k s) x (n + k s)] col-major out, xout [n + k s].
int srk_oracle_ekf_add_points(int64_t n, const double* P, const double* x, int64_t k, int s, const double* x_new, const double* Jy, const double* Qnew, int diag_only,
                              double* Pout, double* xout) {
    EkfMat Pm((size_t)n, (size_t)n);
    std::memcpy(Pm.d.data(), P, sizeof(double) * (size_t)n * (size_t)n);
    std::vector<double> xs(x, x + n);
    for (int64_t i = 0; i < k; ++i) EkfAddSalientPoint(&xs, &Pm, s, x_new + i * s, Jy + i * s * 7, Qnew + i * s * s, diag_only != 0);
    std::memcpy(Pout, Pm.d.data(), sizeof(double) * Pm.d.size());
    std::memcpy(xout, xs.data(), sizeof(double) * xs.size());
    return 0;
}
int srk_oracle_ekf_predict(int64_t n, double* P, const double* F13, const double* GQGt13, int fix_symmetry) {
    EkfMat Pm((size_t)n, (size_t)n);
    std::memcpy(Pm.d.data(), P, sizeof(double) * (size_t)n * (size_t)n);
    EkfPredictCovariance(&Pm, F13, GQGt13, fix_symmetry != 0);
    std::memcpy(P, Pm.d.data(), sizeof(double) * (size_t)n * (size_t)n);
    return 0;
}

int srk_oracle_triangulate(int64_t n_tracks, const int64_t* track_begin, const int32_t* obs_frame, const double* obs_xy, const double* proj, double f0,
                           double* out) {
    srk_oracle::Triangulate(n_tracks, track_begin, obs_frame, obs_xy, proj, f0, out);
    return 0;
}

}  // extern "C"
