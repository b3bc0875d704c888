// suriko-b200 — drop-in adapter with the surface of suriko::BundleAdjustmentKanatani
// (/root/reference/cpp_impl/suriko-engine/include/suriko/bundle-adj-kanatani.h:68-261) on top of the C ABI
// (include/srk/ba_c_api.h).  Header-only C++17; link with libsrk_ba.so.
//
// The adapter only (a) flattens the containers into the SoA problem through GetCorner/EachCorner semantics (quirks Q10/Q11
// of SURVEY.md section 8: pnt_ind is the running index over tracks that own a SalientPointId; frames beyond the pose vector
// are never probed), (b) calls srk_ba_solve, (c) scatters the refined points and poses back in place.  K is never written
// (quirk Q2).  Every numerical step runs on the GPU; there is no CPU fallback: construction throws when no B200 is usable.
#pragma once
#include <cmath>
#include <stdexcept>
#include <string>
#include <vector>

#include "../srk/ba_c_api.h"
#include "obs-geom.h"

namespace suriko_compat {

class BundleAdjustmentKanataniTermCriteria {  // bundle-adj-kanatani.h:68-92
    std::optional<Scalar> allowed_reproj_err_rel_change_;
    std::optional<Scalar> max_hessian_factor_;
public:
    void AllowedReprojErrRelativeChange(Scalar v) { allowed_reproj_err_rel_change_ = v; }
    std::optional<Scalar> AllowedReprojErrRelativeChange() const { return allowed_reproj_err_rel_change_; }
    void MaxHessianFactor(Scalar v) { max_hessian_factor_ = v; }
    std::optional<Scalar> MaxHessianFactor() const { return max_hessian_factor_; }
};

struct FlatScene {
    std::vector<int32_t> obs_cam, obs_point;
    std::vector<double> obs_xy, points, K;
    std::vector<size_t> salient_point_ids;  // pnt_ind -> SalientPointId
    bool shared_K = false;
    srk_ba_problem Problem(std::vector<SE3Transform>& cams, Scalar f0) {
        srk_ba_problem p{};
        p.n_cams = (int64_t)cams.size(); p.n_points = (int64_t)salient_point_ids.size(); p.n_obs = (int64_t)obs_cam.size();
        p.obs_cam = obs_cam.data(); p.obs_point = obs_point.data(); p.obs_xy = obs_xy.data();
        p.points = points.data(); p.cams = reinterpret_cast<double*>(cams.data()); p.K = K.data();
        p.shared_K = shared_K ? 1 : 0; p.f0 = f0;
        return p;
    }
};

inline FlatScene FlattenScene(const FragmentMap& map, size_t frames_count, const CornerTrackRepository& track_rep, const Mat33* shared_K,
                              const std::vector<Mat33>* Ks) {
    if (!((shared_K != nullptr) ^ (Ks != nullptr))) throw std::invalid_argument("Provide either shared K or separate K for each camera frame");  // BA.cpp:421
    FlatScene fs;
    fs.shared_K = shared_K != nullptr;
    for (const CornerTrack& track : track_rep.CornerTracks) {
        if (!track.SalientPointId.has_value()) continue;
        const int32_t pnt_ind = (int32_t)fs.salient_point_ids.size();
        fs.salient_point_ids.push_back(track.SalientPointId.value());
        const Point3& x = map.GetSalientPoint(track.SalientPointId.value());
        fs.points.insert(fs.points.end(), {x[0], x[1], x[2]});
        if (!track.HasCorners()) continue;
        track.EachCorner([&](size_t frame_ind, const std::optional<CornerData>& cd) {
            if (!cd.has_value() || frame_ind >= frames_count) return;
            fs.obs_cam.push_back((int32_t)frame_ind); fs.obs_point.push_back(pnt_ind);
            fs.obs_xy.push_back(cd.value().pixel_coord[0]); fs.obs_xy.push_back(cd.value().pixel_coord[1]);
        });
    }
    if (shared_K != nullptr) fs.K.assign(shared_K->a, shared_K->a + 9);
    else for (const Mat33& k : *Ks) fs.K.insert(fs.K.end(), k.a, k.a + 9);
    return fs;
}

class BundleAdjustmentKanatani {
public:
    static constexpr size_t kPointVarsCount = 3, kIntrinsicVarsCount = 4, kTVarsCount = 3, kWVarsCount = 3;

    explicit BundleAdjustmentKanatani(int device = 0) {
        int dev = device;
        if (srk_ba_create(&handle_, &dev, 1) != SRK_OK) throw std::runtime_error(std::string("srk_ba_create: ") + srk_last_error());
    }
    ~BundleAdjustmentKanatani() { srk_ba_destroy(handle_); }
    BundleAdjustmentKanatani(const BundleAdjustmentKanatani&) = delete;
    BundleAdjustmentKanatani& operator=(const BundleAdjustmentKanatani&) = delete;

    // bundle-adj-kanatani.h:167-172 / BA.cpp:589-600 — sum of squared residuals in (pix/f0)^2, no normalisation.
    Scalar ReprojError(Scalar f0, const FragmentMap& map, const std::vector<SE3Transform>& inverse_orient_cams, const CornerTrackRepository& track_rep,
                       const Mat33* shared_intrinsic_cam_mat, const std::vector<Mat33>* intrinsic_cam_mats, size_t* seen_points_count = nullptr) {
        FlatScene fs = FlattenScene(map, inverse_orient_cams.size(), track_rep, shared_intrinsic_cam_mat, intrinsic_cam_mats);
        std::vector<SE3Transform> cams = inverse_orient_cams;
        srk_ba_problem p = fs.Problem(cams, f0);
        double err = 0; int64_t seen = 0;
        if (srk_ba_reproj_error(handle_, &p, &err, &seen) != SRK_OK) throw std::runtime_error(std::string("srk_ba_reproj_error: ") + srk_last_error());
        if (seen_points_count != nullptr) *seen_points_count = (size_t)seen;
        return err;
    }
    // BA.cpp:602-615 (quirk Q15: no -7 dof correction)
    Scalar ReprojErrorPixPerPoint(Scalar reproj_err, size_t seen_points_count) const { return f0_ * std::sqrt(reproj_err / (Scalar)seen_points_count); }

    // bundle-adj-kanatani.h:179-184 / BA.cpp:617-718 — refines map and inverse_orient_cams in place, returns is_optimized.
    bool ComputeInplace(Scalar f0, FragmentMap& map, std::vector<SE3Transform>& inverse_orient_cams, const CornerTrackRepository& track_rep,
                        const Mat33* shared_intrinsic_cam_mat, std::vector<Mat33>* intrinsic_cam_mats, const BundleAdjustmentKanataniTermCriteria& term_crit) {
        if (!(unity_t1_comp_ind_ < kTVarsCount)) throw std::invalid_argument("Can normalize only one of [T1x, T1y, Tz] components");  // BA.cpp:628
        f0_ = f0;
        FlatScene fs = FlattenScene(map, inverse_orient_cams.size(), track_rep, shared_intrinsic_cam_mat, intrinsic_cam_mats);
        points_count_ = fs.salient_point_ids.size(); frames_count_ = inverse_orient_cams.size();
        srk_ba_problem p = fs.Problem(inverse_orient_cams, f0);
        srk_ba_options opt; srk_ba_default_options(&opt);
        if (term_crit.AllowedReprojErrRelativeChange().has_value()) { opt.has_err_change = 1; opt.err_change = term_crit.AllowedReprojErrRelativeChange().value(); }
        if (term_crit.MaxHessianFactor().has_value()) { opt.has_max_hessian_factor = 1; opt.max_hessian_factor = term_crit.MaxHessianFactor().value(); }
        opt.unity_comp_ind = (int32_t)unity_t1_comp_ind_; opt.unity_comp_value = unity_t1_comp_value_;
        opt.max_outer_iters = max_outer_iters; opt.solver = solver;
        srk_ba_report rep{};
        if (srk_ba_solve(handle_, &p, &opt, &rep) != SRK_OK) throw std::runtime_error(std::string("srk_ba_solve: ") + srk_last_error());
        report_ = rep;
        optimization_stop_reason_ = srk_stop_reason_string(rep.stop_reason);
        for (size_t i = 0; i < fs.salient_point_ids.size(); ++i) {   // poses were refined in place through p.cams
            Point3& x = map.GetSalientPoint(fs.salient_point_ids[i]);
            x[0] = fs.points[3 * i]; x[1] = fs.points[3 * i + 1]; x[2] = fs.points[3 * i + 2];
        }
        return rep.converged != 0;
    }

    size_t PointsCount() const { return points_count_; }
    size_t FramesCount() const { return frames_count_; }
    size_t VarsCount() const { return kPointVarsCount * points_count_ + 10 * frames_count_; }   // quirk Q14: always 10 per frame
    size_t NormalizedVarsCount() const { return VarsCount() - 7; }
    const std::string& OptimizationStatusString() const { return optimization_stop_reason_; }
    const srk_ba_report& LastReport() const { return report_; }

    Scalar unity_t1_comp_value_ = 1.0;   // bundle-adj-kanatani.h:133
    size_t unity_t1_comp_ind_ = 1;       // bundle-adj-kanatani.h:134
    int32_t max_outer_iters = 0;         // 0 = unlimited (the reference's behaviour, quirk Q9)
    int32_t solver = SRK_SOLVER_AUTO;

private:
    void* handle_ = nullptr;
    Scalar f0_ = 0;
    size_t points_count_ = 0, frames_count_ = 0;
    std::string optimization_stop_reason_;
    srk_ba_report report_{};
};

}  // namespace suriko_compat
