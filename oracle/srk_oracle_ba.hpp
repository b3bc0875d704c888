// TEST INFRASTRUCTURE ONLY — CPU oracle: a plain C++17 restatement of suriko-engine's Kanatani bundle
// adjustment (BA.cpp = /root/reference/cpp_impl/suriko-engine/src/bundle-adj-kanatani.cpp,
// BA.h = .../include/suriko/bundle-adj-kanatani.h).  Function-for-function, single-threaded, FP64 by default.
// Never used by the product path; see oracle/README.md.
//
// Two data flows share every per-observation formula:
//   * SchurFlow::DenseReference   — the reference's own flow: O(N*M) GetCorner probes, dense F [3N x 10M],
//                                   dense per-point  S -= (A^T E^-1) A  over all n_f columns (BA.cpp:1859-1900).
//   * SchurFlow::SparseEquivalent — identical arithmetic restricted to the non-zero 3x10 blocks; needed for
//                                   scenes where the dense F (O(N*M) memory) cannot exist.  On the same scene it is
//                                   bit-identical to DenseReference (the skipped terms are exact zeros).
// Two solvers for S*df = rhs:
//   * SolveImpl::HouseholderQR    — unpivoted Householder QR, the algorithm behind Eigen's householderQr().solve
//                                   (BA.cpp:1911; Eigen is a third-party dependency, version unpinned).
//   * SolveImpl::CholeskyRefined  — LL^T + iterative refinement with a wider-precision residual ("exact" mode).
// `Acc` is the type in which the Schur complement is formed (double = faithful; long double = "exact" mode used
// as the parity target, see DESIGN.md "Parity budget").
#pragma once
#include <array>
#include <limits>
#include <string>
#include <type_traits>
#include <unordered_map>
#include <vector>
#include "srk_oracle_geom.hpp"
#include "srk_oracle_parallel.hpp"

namespace srk_oracle {

template <class F>
struct DynMat {  // column-major dense matrix (Eigen default layout)
    size_t rows = 0, cols = 0;
    std::vector<F> d;
    void resize(size_t r, size_t c) { rows = r; cols = c; d.assign(r * c, F(0)); }
    void fill(F v) { std::fill(d.begin(), d.end(), v); }
    F& operator()(size_t r, size_t c) { return d[c * rows + r]; }
    F operator()(size_t r, size_t c) const { return d[c * rows + r]; }
};

// SparseThreaded: the SparseEquivalent arithmetic spread over host threads for the timed CPU arm of the full-size configurations
// (bench.py --impl reference): derivative passes per point / per frame (bit-identical to the serial order), per-observation errors summed
// serially in the reference's frame-major order (bit-identical), the Schur complement accumulated per thread in 10x10 camera-pair blocks
// (lower block triangle) merged in thread order, a skyline Cholesky of the block-banded system instead of the reference's dense
// Householder QR (which needs 1.3e12 flop at n_f = 9993), back substitution per point.  NOT the reference's single-threaded data flow:
// it exists so that the CPU arm can run the SAME configuration as the GPU arm in minutes, and says so wherever it is reported.
enum class SchurFlow { DenseReference = 0, SparseEquivalent = 1, SparseThreaded = 2 };
enum class SolveImpl { HouseholderQR = 0, CholeskyRefined = 1, None = 2 };   // None: assemble S / rhs only (parity checks of large systems)
enum StopReason { kStopNone = 0, kStopAbsErrThreshold = 1, kStopSmallErrChange = 2, kStopHessianOverflow = 3, kStopErrConverged = 4,
                  kStopNormalizationFailed = 5, kStopMaxIters = 6 };

inline const char* StopReasonString(int r) {
    switch (r) {  // BA.cpp:751, :866, :868, :882
    case kStopAbsErrThreshold: return "abs err threshold";
    case kStopSmallErrChange: return "small relative err change";
    case kStopHessianOverflow: return "hessian overflow";
    case kStopErrConverged: return "err converged to limit value";
    case kStopMaxIters: return "max iterations (not a reference stop reason)";
    default: return "";
    }
}

template <class F>
struct TermCriteria {  // BA.h:68-92
    std::optional<F> allowed_reproj_err_rel_change;
    std::optional<F> max_hessian_factor;
};

template <class F>
struct AttemptRecord { F hessian_factor; F err_new; int accepted; long skipped_points; };

template <class F>
struct Trace {
    F err_initial = 0;
    std::vector<F> err_per_iter;        // accepted error after each successful outer iteration
    std::vector<F> factor_per_iter;     // hessian factor that produced it
    std::vector<AttemptRecord<F>> attempts;
    size_t seen_points = 0;
    int stop_reason = kStopNone;
    bool converged = false;
    size_t outer_iters = 0;
};

// ---------------------------------------------------------------------------------------------------------
// Dense linear algebra used by the solve.

// Householder QR solve: the published algorithm of Eigen::HouseholderQR (unblocked makeHouseholder /
// applyHouseholderOnTheLeft), A is destroyed.  Returns x with A x = b (square A).
template <class F>
inline void HouseholderQrSolve(DynMat<F>& A, std::vector<F>& b, std::vector<F>* x) {
    using std::sqrt; using std::abs;
    size_t n = A.rows;
    std::vector<F> tau(n, F(0));
    for (size_t k = 0; k < n; ++k) {
        F c0 = A(k, k);
        F tail_sq = 0;
        for (size_t i = k + 1; i < n; ++i) tail_sq += A(i, k) * A(i, k);
        F beta;
        if (tail_sq <= std::numeric_limits<F>::min()) {
            tau[k] = 0; beta = c0;
            for (size_t i = k + 1; i < n; ++i) A(i, k) = 0;
        } else {
            beta = sqrt(c0 * c0 + tail_sq);
            if (c0 >= 0) beta = -beta;
            for (size_t i = k + 1; i < n; ++i) A(i, k) /= (c0 - beta);
            tau[k] = (beta - c0) / beta;
        }
        A(k, k) = beta;
        if (tau[k] != F(0)) {
            // apply H = I - tau v v^T (v = [1; essential]) to trailing columns and to b
            for (size_t j = k + 1; j < n; ++j) {
                F s = A(k, j);
                for (size_t i = k + 1; i < n; ++i) s += A(i, k) * A(i, j);
                s *= tau[k];
                A(k, j) -= s;
                for (size_t i = k + 1; i < n; ++i) A(i, j) -= s * A(i, k);
            }
            F s = b[k];
            for (size_t i = k + 1; i < n; ++i) s += A(i, k) * b[i];
            s *= tau[k];
            b[k] -= s;
            for (size_t i = k + 1; i < n; ++i) b[i] -= s * A(i, k);
        }
    }
    x->assign(n, F(0));
    for (size_t ii = n; ii-- > 0;) {
        F s = b[ii];
        for (size_t j = ii + 1; j < n; ++j) s -= A(ii, j) * (*x)[j];
        (*x)[ii] = s / A(ii, ii);
    }
}

// LL^T in place (lower).  Returns false if a pivot is not positive/finite.
template <class F>
inline bool CholeskyInplace(DynMat<F>& A) {
    using std::sqrt;
    size_t n = A.rows;
    for (size_t j = 0; j < n; ++j) {
        F d = A(j, j);
        for (size_t k = 0; k < j; ++k) d -= A(j, k) * A(j, k);
        if (!(d > F(0)) || !std::isfinite((double)d)) return false;
        d = sqrt(d);
        A(j, j) = d;
        for (size_t i = j + 1; i < n; ++i) {
            F s = A(i, j);
            for (size_t k = 0; k < j; ++k) s -= A(i, k) * A(j, k);
            A(i, j) = s / d;
        }
    }
    return true;
}
template <class F>
inline void CholeskySolve(const DynMat<F>& L, const std::vector<F>& b, std::vector<F>* x) {
    size_t n = L.rows;
    std::vector<F>& y = *x; y = b;
    for (size_t i = 0; i < n; ++i) { F s = y[i]; for (size_t k = 0; k < i; ++k) s -= L(i, k) * y[k]; y[i] = s / L(i, i); }
    for (size_t ii = n; ii-- > 0;) { F s = y[ii]; for (size_t k = ii + 1; k < n; ++k) s -= L(k, ii) * y[k]; y[ii] = s / L(ii, ii); }
}

// ---------------------------------------------------------------------------------------------------------

template <class F, class Acc = F>
class BundleAdjustmentKanatani {
public:
    static constexpr size_t kPointVarsCount = 3, kIntrinsicVarsCount = 4, kTVarsCount = 3, kWVarsCount = 3, kV = 10;

    SchurFlow schur_flow = SchurFlow::DenseReference;
    SolveImpl solve_impl = SolveImpl::HouseholderQR;
    size_t max_outer_iters = 0;   // 0 = unlimited, as the reference (quirk Q9)
    bool apply_intrinsics = false; // false = reference behaviour: K corrections solved but dropped (quirk Q2)
    int refine_steps = 3;
    bool chol_in_double = false;
    Trace<F> trace;

    // state (public so tests can inspect a single derivative pass)
    F f0_ = 0;
    FragmentMap<F>* map_ = nullptr;
    std::vector<SE3<F>>* inverse_orient_cams_ = nullptr;
    const CornerTrackRepository<F>* track_rep_ = nullptr;
    const Mat33<F>* shared_K_ = nullptr;
    std::vector<Mat33<F>>* Ks_ = nullptr;
    F unity_t1_comp_value_ = 1.0;
    size_t unity_t1_comp_ind_ = 1;
    size_t vars_count_per_frame_ = kV;
    std::array<size_t, 7> normalized_var_indices_{};
    size_t normalized_var_indices_count_ = 0;
    std::string optimization_stop_reason_;

    std::vector<F> gradE_;         // [3N + 10M]
    DynMat<F> E_;                  // deriv_second_pointpoint [3N x 3]
    DynMat<F> G_;                  // deriv_second_frameframe [10M x 10]
    DynMat<F> Fdense_;             // deriv_second_pointframe [3N x 10M]   (DenseReference only)
    std::vector<F> corrections_;   // [3N + 10M]
    DynMat<Acc> S_;                // left_side [n_f x n_f]
    std::vector<Acc> rhs_;         // right_side [n_f]
    std::vector<unsigned char> skipped_mask_;  // per pnt_ind, last attempt (quirk Q5)

    // flattened view (SparseEquivalent): observations in (pnt_ind, frame_ind) order == reference track order
    struct Obs { uint32_t pnt, frame; F x, y; };
    std::vector<Obs> obs_;
    std::vector<size_t> pt_begin_;           // [N+1]
    std::vector<size_t> pnt_to_map_id_;      // pnt_ind -> SalientPointId (quirk Q10)
    std::vector<std::vector<uint32_t>> by_frame_;  // per frame: obs indices in track order (frame-major sum, quirk Q12)
    std::vector<F> Fblk_;                    // per obs 3x10 block, row-major [3][10]
    std::vector<int> red_index_;             // [10M] -> reduced index or -1

    // BA.cpp:539-563
    void InitializeNormalizedVarIndices() {
        size_t off = 0, out = 0;
        off += kIntrinsicVarsCount;
        for (size_t i = 0; i < 6; ++i) normalized_var_indices_[out++] = off + i;
        off += kTVarsCount + kWVarsCount;
        off += kIntrinsicVarsCount;
        normalized_var_indices_[out++] = off + unity_t1_comp_ind_;
        normalized_var_indices_count_ = out;
    }

    size_t PointsCount() const { return map_->SalientPointsCount(); }
    size_t FramesCount() const { return inverse_orient_cams_->size(); }
    size_t VarsCount() const { return kPointVarsCount * PointsCount() + vars_count_per_frame_ * FramesCount(); }
    size_t NormalizedVarsCount() const { return VarsCount() - normalized_var_indices_count_; }
    size_t NormalizedFrameVarsCount() const { return vars_count_per_frame_ * FramesCount() - normalized_var_indices_count_; }
    const std::string& OptimizationStatusString() const { return optimization_stop_reason_; }

    const Mat33<F>& GetK(size_t frame_ind) const { return shared_K_ != nullptr ? *shared_K_ : (*Ks_)[frame_ind]; }

    // BA.cpp:410-490 (ReprojErrorWithOverlap without patches) — frame-major, tracks inner, plain left-to-right sum.
    static F ReprojError(F f0, const FragmentMap<F>& map, const std::vector<SE3<F>>& inverse_orient_cams,
                         const CornerTrackRepository<F>& track_rep, const Mat33<F>* shared_K,
                         const std::vector<Mat33<F>>* Ks, size_t* seen_points_count = nullptr) {
        if (IsClose<F>(F(0), f0)) throw std::invalid_argument("f0 != 0");
        if (!((shared_K != nullptr) ^ (Ks != nullptr))) throw std::invalid_argument("Provide either shared K or separate K for each camera frame");
        F err_sum = 0;
        size_t seen = 0;
        size_t frames_count = inverse_orient_cams.size();
        for (size_t frame_ind = 0; frame_ind < frames_count; ++frame_ind) {
            const SE3<F>& rt = inverse_orient_cams[frame_ind];
            const Mat33<F>& K = shared_K != nullptr ? *shared_K : (*Ks)[frame_ind];
            for (const CornerTrack<F>& track : track_rep.CornerTracks) {
                if (!track.SalientPointId.has_value()) continue;
                std::optional<Point2<F>> corner = track.GetCorner(frame_ind);
                if (!corner.has_value()) continue;
                err_sum += OneReprojErr(f0, K, rt, map.GetSalientPoint(track.SalientPointId.value()), corner.value());
                seen += 1;
            }
        }
        if (seen_points_count != nullptr) *seen_points_count = seen;
        return err_sum;
    }
    // BA.cpp:460-479
    static F OneReprojErr(F f0, const Mat33<F>& K, const SE3<F>& rt, const Vec3<F>& x3D, const Point2<F>& pix) {
        F c0 = pix[0] / f0, c1 = pix[1] / f0;
        Vec3<F> x3D_cam = SE3Apply(rt, x3D);
        Vec3<F> h = K * x3D_cam;
        F x = h[0] / h[2], y = h[1] / h[2];
        F dx = x - c0, dy = y - c1;
        return dx * dx + dy * dy;
    }
    // BA.cpp:602-615 (quirk Q15)
    F ReprojErrorPixPerPoint(F reproj_err, size_t seen_points_count) const { using std::sqrt; return f0_ * sqrt(reproj_err / F(seen_points_count)); }

    F CurrentReprojError(size_t* seen = nullptr) const {
        if (schur_flow == SchurFlow::DenseReference)
            return ReprojError(f0_, *map_, *inverse_orient_cams_, *track_rep_, shared_K_, Ks_, seen);
        if (schur_flow == SchurFlow::SparseThreaded) {   // per-observation terms in parallel, summed serially in the same frame-major order
            std::vector<F> term(obs_.size());
            ParallelFor(0, (int64_t)obs_.size(), 4096, [&](int64_t oi) {
                const Obs& o = obs_[(size_t)oi];
                term[(size_t)oi] = OneReprojErr(f0_, GetK(o.frame), (*inverse_orient_cams_)[o.frame], map_->GetSalientPoint(pnt_to_map_id_[o.pnt]), Point2<F>(o.x, o.y));
            });
            F acc = 0; size_t n = 0;
            for (size_t frame_ind = 0; frame_ind < by_frame_.size(); ++frame_ind)
                for (uint32_t oi : by_frame_[frame_ind]) { acc += term[oi]; n += 1; }
            if (seen != nullptr) *seen = n;
            return acc;
        }
        // SparseEquivalent: same per-observation arithmetic, same frame-major / track-inner summation order.
        F err_sum = 0; size_t cnt = 0;
        for (size_t frame_ind = 0; frame_ind < by_frame_.size(); ++frame_ind) {
            const SE3<F>& rt = (*inverse_orient_cams_)[frame_ind];
            const Mat33<F>& K = GetK(frame_ind);
            for (uint32_t oi : by_frame_[frame_ind]) {
                const Obs& o = obs_[oi];
                err_sum += OneReprojErr(f0_, K, rt, map_->GetSalientPoint(pnt_to_map_id_[o.pnt]), Point2<F>(o.x, o.y));
                cnt += 1;
            }
        }
        if (seen != nullptr) *seen = cnt;
        return err_sum;
    }

    // Flatten through GetCorner/EachCorner semantics (quirks Q10, Q11); pnt_ind = running index over tracks with a SalientPointId.
    void BuildFlatView() {
        size_t M = FramesCount();
        obs_.clear(); pt_begin_.clear(); pnt_to_map_id_.clear();
        by_frame_.assign(M, {});
        pt_begin_.push_back(0);
        for (size_t tid = 0; tid < track_rep_->CornerTracksCount(); ++tid) {
            const CornerTrack<F>& track = track_rep_->GetPointTrackById(tid);
            if (!track.SalientPointId.has_value()) continue;
            uint32_t pnt_ind = (uint32_t)pnt_to_map_id_.size();
            pnt_to_map_id_.push_back(track.SalientPointId.value());
            track.EachCorner([&](size_t frame_ind, const std::optional<CornerData<F>>& cd) {
                if (!cd.has_value() || frame_ind >= M) return;  // frames >= M are never probed by the reference loops
                by_frame_[frame_ind].push_back((uint32_t)obs_.size());
                obs_.push_back(Obs{pnt_ind, (uint32_t)frame_ind, cd.value().pixel_coord[0], cd.value().pixel_coord[1]});
            });
            pt_begin_.push_back(obs_.size());
        }
        red_index_.assign(kV * M, -1);
        int out = 0;
        for (size_t v = 0; v < kV * M; ++v) {
            bool removed = false;
            for (size_t i = 0; i < normalized_var_indices_count_; ++i) if (normalized_var_indices_[i] == v) removed = true;
            if (!removed) red_index_[v] = out++;
        }
    }

    // BA.cpp:572-587
    void EnsureMemoryAllocated() {
        size_t N = PointsCount(), M = FramesCount();
        gradE_.assign(N * kPointVarsCount + M * vars_count_per_frame_, F(0));
        E_.resize(N * kPointVarsCount, kPointVarsCount);
        G_.resize(M * vars_count_per_frame_, vars_count_per_frame_);
        if (schur_flow == SchurFlow::DenseReference) Fdense_.resize(N * kPointVarsCount, M * vars_count_per_frame_);
        else Fblk_.assign(obs_.size() * 30, F(0));
        corrections_.assign(VarsCount(), F(0));
        size_t nf = NormalizedFrameVarsCount();
        if (schur_flow == SchurFlow::SparseThreaded) S_.resize(0, 0); else S_.resize(nf, nf);   // the threaded flow keeps S in skyline form
        rhs_.assign(nf, Acc(0));
        skipped_mask_.assign(N, 0);
    }

    // BA.cpp:1450-1455 — row v = d(p,q,r)/d(X|Y|Z) = column v of P.
    static void ComputePointPqrDerivatives(const Mat33<F>& K, const SE3<F>& rt, F out[3][3]) {
        Mat33<F> KR = K * rt.R;  // BA.cpp:1193-1194  P << K*R, K*T
        for (size_t v = 0; v < 3; ++v) for (size_t c = 0; c < 3; ++c) out[v][c] = KR(c, v);
    }

    // BA.cpp:1457-1525 (quirk Q3: f0 appears verbatim)
    void ComputeFramePqrDerivatives(const Mat33<F>& K, const SE3<F>& inverse_orient_cam, const Vec3<F>& salient_point, F out[10][3]) const {
        F fx = K(0, 0), fy = K(1, 1), u0 = K(0, 2), v0 = K(1, 2);
        Vec3<F> x3D_cam = SE3Apply(inverse_orient_cam, salient_point);
        Vec3<F> pqr = K * x3D_cam;
        size_t i = 0;
        out[i][0] = (F(1) / fx) * pqr[0] - u0 / (f0_ * fx) * pqr[2]; out[i][1] = 0; out[i][2] = 0; ++i;  // fx
        out[i][0] = 0; out[i][1] = (F(1) / fy) * pqr[1] - v0 / (f0_ * fy) * pqr[2]; out[i][2] = 0; ++i;  // fy
        out[i][0] = (F(1) / f0_) * pqr[2]; out[i][1] = 0; out[i][2] = 0; ++i;                             // u0
        out[i][0] = 0; out[i][1] = (F(1) / f0_) * pqr[2]; out[i][2] = 0; ++i;                             // v0
        SE3<F> direct = SE3Inv(inverse_orient_cam);
        Vec3<F> c0 = direct.R.col(0), c1 = direct.R.col(1), c2 = direct.R.col(2);
        Vec3<F> tp = -(fx * c0 + u0 * c2), tq = -(fy * c1 + v0 * c2), tr = -(f0_ * c2);
        for (size_t t = 0; t < 3; ++t) { out[i][0] = tp[t]; out[i][1] = tq[t]; out[i][2] = tr[t]; ++i; }
        Vec3<F> rot1 = fx * c0 + u0 * c2, rot2 = fy * c1 + v0 * c2, rot3 = f0_ * c2;
        Vec3<F> t_to_pnt = salient_point - direct.T;
        Vec3<F> wp = Cross(rot1, t_to_pnt), wq = Cross(rot2, t_to_pnt), wr = Cross(rot3, t_to_pnt);
        for (size_t t = 0; t < 3; ++t) { out[i][0] = wp[t]; out[i][1] = wq[t]; out[i][2] = wr[t]; ++i; }
    }

    // BA.cpp:1528-1537 — formula 8
    static F FirstDerivFromPqrDerivative(F f0, const Vec3<F>& pqr, const Point2<F>& pix, F gp, F gq, F gr) {
        F result = (pqr[0] / pqr[2] - pix[0] / f0) * (pqr[2] * gp - pqr[0] * gr) +
                   (pqr[1] / pqr[2] - pix[1] / f0) * (pqr[2] * gq - pqr[1] * gr);
        result *= F(2) / (pqr[2] * pqr[2]);
        return result;
    }
    // BA.cpp:1540-1549 — formula 9
    static F SecondDerivFromPqrDerivative(const Vec3<F>& pqr, F gp1, F gq1, F gr1, F gp2, F gq2, F gr2) {
        F s = (pqr[2] * gp1 - pqr[0] * gr1) * (pqr[2] * gp2 - pqr[0] * gr2) +
              (pqr[2] * gq1 - pqr[1] * gr1) * (pqr[2] * gq2 - pqr[1] * gr2);
        s *= F(2) / (pqr[2] * pqr[2] * pqr[2] * pqr[2]);
        return s;
    }

    // One observation's contribution to the point pass (BA.cpp:1184-1220).
    void AccumulatePointPass(size_t pnt_ind, size_t frame_ind, const Vec3<F>& X, const Point2<F>& pix) {
        const SE3<F>& rt = (*inverse_orient_cams_)[frame_ind];
        const Mat33<F>& K = GetK(frame_ind);
        Vec3<F> pqr = K * SE3Apply(rt, X);
        F pd[3][3];
        ComputePointPqrDerivatives(K, rt, pd);
        for (size_t v = 0; v < 3; ++v)
            gradE_[pnt_ind * 3 + v] += FirstDerivFromPqrDerivative(f0_, pqr, pix, pd[v][0], pd[v][1], pd[v][2]);
        for (size_t v1 = 0; v1 < 3; ++v1)
            for (size_t v2 = 0; v2 < 3; ++v2)
                E_(pnt_ind * 3 + v1, v2) += SecondDerivFromPqrDerivative(pqr, pd[v1][0], pd[v1][1], pd[v1][2], pd[v2][0], pd[v2][1], pd[v2][2]);
    }
    // One observation's contribution to the frame pass (BA.cpp:1295-1358); Fout = 3x10 row-major destination.
    template <class StoreF>
    void AccumulateFramePass(size_t frame_ind, const Vec3<F>& X, const Point2<F>& pix, StoreF&& store_F) {
        const SE3<F>& rt = (*inverse_orient_cams_)[frame_ind];
        const Mat33<F>& K = GetK(frame_ind);
        Vec3<F> pqr = K * SE3Apply(rt, X);
        F fd[10][3];
        ComputeFramePqrDerivatives(K, rt, X, fd);
        size_t goff = PointsCount() * 3 + frame_ind * kV;
        for (size_t v = 0; v < kV; ++v) gradE_[goff + v] += FirstDerivFromPqrDerivative(f0_, pqr, pix, fd[v][0], fd[v][1], fd[v][2]);
        for (size_t v1 = 0; v1 < kV; ++v1)
            for (size_t v2 = 0; v2 < kV; ++v2)
                G_(frame_ind * kV + v1, v2) += SecondDerivFromPqrDerivative(pqr, fd[v1][0], fd[v1][1], fd[v1][2], fd[v2][0], fd[v2][1], fd[v2][2]);
        F pd[3][3];
        ComputePointPqrDerivatives(K, rt, pd);
        for (size_t pv = 0; pv < 3; ++pv)
            for (size_t fv = 0; fv < kV; ++fv)
                store_F(pv, fv, SecondDerivFromPqrDerivative(pqr, pd[pv][0], pd[pv][1], pd[pv][2], fd[fv][0], fd[fv][1], fd[fv][2]));
    }

    // BA.cpp:1140-1448
    void ComputeCloseFormReprErrorDerivatives() {
        size_t M = FramesCount();
        std::fill(gradE_.begin(), gradE_.end(), F(0));
        E_.fill(0); G_.fill(0);
        if (schur_flow == SchurFlow::DenseReference) {
            Fdense_.fill(0);
            // point pass: tracks outer, frames inner
            size_t pnt_ind = (size_t)-1;
            for (size_t tid = 0; tid < track_rep_->CornerTracksCount(); ++tid) {
                const CornerTrack<F>& track = track_rep_->GetPointTrackById(tid);
                if (!track.SalientPointId.has_value()) continue;
                pnt_ind += 1;
                const Vec3<F>& X = map_->GetSalientPoint(track.SalientPointId.value());
                for (size_t frame_ind = 0; frame_ind < M; ++frame_ind) {
                    auto corner = track.GetCorner(frame_ind);
                    if (!corner.has_value()) continue;
                    AccumulatePointPass(pnt_ind, frame_ind, X, corner.value());
                }
            }
            // frame pass: frames outer, tracks inner
            for (size_t frame_ind = 0; frame_ind < M; ++frame_ind) {
                size_t pi = (size_t)-1;
                for (size_t tid = 0; tid < track_rep_->CornerTracksCount(); ++tid) {
                    const CornerTrack<F>& track = track_rep_->GetPointTrackById(tid);
                    if (!track.SalientPointId.has_value()) continue;
                    pi += 1;
                    auto corner = track.GetCorner(frame_ind);
                    if (!corner.has_value()) continue;
                    const Vec3<F>& X = map_->GetSalientPoint(track.SalientPointId.value());
                    AccumulateFramePass(frame_ind, X, corner.value(), [&](size_t pv, size_t fv, F s) {
                        Fdense_(pi * 3 + pv, frame_ind * kV + fv) += s;
                    });
                }
            }
        } else if (schur_flow == SchurFlow::SparseThreaded) {
            std::fill(Fblk_.begin(), Fblk_.end(), F(0));
            ParallelFor(0, (int64_t)pt_begin_.size() - 1, 1024, [&](int64_t pnt_ind) {      // a point's gradE / E entries are its own
                const Vec3<F>& X = map_->GetSalientPoint(pnt_to_map_id_[(size_t)pnt_ind]);
                for (size_t oi = pt_begin_[(size_t)pnt_ind]; oi < pt_begin_[(size_t)pnt_ind + 1]; ++oi)
                    AccumulatePointPass((size_t)pnt_ind, obs_[oi].frame, X, Point2<F>(obs_[oi].x, obs_[oi].y));
            });
            ParallelFor(0, (int64_t)M, 1, [&](int64_t frame_ind) {                            // a frame's gradE / G entries are its own
                for (uint32_t oi : by_frame_[(size_t)frame_ind]) {
                    const Obs& o = obs_[oi];
                    const Vec3<F>& X = map_->GetSalientPoint(pnt_to_map_id_[o.pnt]);
                    F* blk = &Fblk_[(size_t)oi * 30];
                    AccumulateFramePass((size_t)frame_ind, X, Point2<F>(o.x, o.y), [&](size_t pv, size_t fv, F sv) { blk[pv * 10 + fv] += sv; });
                }
            });
        } else {
            std::fill(Fblk_.begin(), Fblk_.end(), F(0));
            for (size_t pnt_ind = 0; pnt_ind + 1 < pt_begin_.size(); ++pnt_ind) {
                const Vec3<F>& X = map_->GetSalientPoint(pnt_to_map_id_[pnt_ind]);
                for (size_t oi = pt_begin_[pnt_ind]; oi < pt_begin_[pnt_ind + 1]; ++oi)
                    AccumulatePointPass(pnt_ind, obs_[oi].frame, X, Point2<F>(obs_[oi].x, obs_[oi].y));
            }
            for (size_t frame_ind = 0; frame_ind < M; ++frame_ind)
                for (uint32_t oi : by_frame_[frame_ind]) {
                    const Obs& o = obs_[oi];
                    const Vec3<F>& X = map_->GetSalientPoint(pnt_to_map_id_[o.pnt]);
                    F* blk = &Fblk_[(size_t)oi * 30];
                    AccumulateFramePass(frame_ind, X, Point2<F>(o.x, o.y), [&](size_t pv, size_t fv, F s) { blk[pv * 10 + fv] += s; });
                }
        }
    }

    // BA.cpp:1780-1823 (fill_matG): gauge rows/cols dropped, diagonal *(1+c)
    void FillMatG(F hessian_factor) {
        S_.fill(Acc(0));
        size_t M = FramesCount();
        for (size_t frame_ind = 0; frame_ind < M; ++frame_ind)
            for (size_t v1 = 0; v1 < kV; ++v1) {
                int r1 = ReducedIndex(frame_ind * kV + v1);
                if (r1 < 0) continue;
                for (size_t v2 = 0; v2 < kV; ++v2) {
                    int r2 = ReducedIndex(frame_ind * kV + v2);
                    if (r2 < 0) continue;
                    F g = G_(frame_ind * kV + v1, v2);
                    if (v1 == v2) g *= F(1) + hessian_factor;
                    S_((size_t)r1, (size_t)r2) = Acc(g);
                }
            }
    }
    int ReducedIndex(size_t frame_var) const {
        if (!red_index_.empty()) return red_index_[frame_var];
        int out = 0;
        for (size_t v = 0; v <= frame_var; ++v) {
            bool removed = false;
            for (size_t i = 0; i < normalized_var_indices_count_; ++i) if (normalized_var_indices_[i] == v) removed = true;
            if (v == frame_var) return removed ? -1 : out;
            if (!removed) ++out;
        }
        return -1;
    }

    // BA.cpp:1825-1834 + :1873-1881 — damped point block, cofactor inverse, |det| > 1e-12 (quirk Q5)
    // The invertibility decision is always taken in F exactly as the reference does; when Acc is wider than F
    // ("exact" mode) the inverse itself is re-evaluated in Acc from the same damped F-valued block.
    bool ScaledPointHessianInverse(size_t pnt_ind, F hessian_factor, Mat33<Acc>* inv) const {
        Mat33<F> H;
        for (size_t r = 0; r < 3; ++r) for (size_t c = 0; c < 3; ++c) H(r, c) = E_(pnt_ind * 3 + r, c);
        for (size_t i = 0; i < 3; ++i) H(i, i) *= F(1) + hessian_factor;
        F det = 0; bool ok = false;
        Mat33<F> invF;
        Inverse3x3WithCheck<F>(H, F(1e-12), &invF, &det, &ok);
        if (!ok) return false;
        if constexpr (std::is_same<F, Acc>::value) { *inv = invF; }
        else {
            Mat33<Acc> HA; for (int i = 0; i < 9; ++i) HA.a[i] = Acc(H.a[i]);
            Acc detA = 0; bool okA = false;
            Inverse3x3WithCheck<Acc>(HA, Acc(0), inv, &detA, &okA);
        }
        return true;
    }

    // Columns (reduced index, 3 values) of the point's gauge-reduced [3 x n_f] row block, ascending reduced index.
    struct Col { int r; F a[3]; };
    void PointColumns(size_t pnt_ind, std::vector<Col>* cols) const {
        cols->clear();
        if (schur_flow == SchurFlow::DenseReference) {
            // get_normalized_point_allframes (BA.cpp:1836-1845): a dense copy with the 7 gauge columns removed
            size_t nv = kV * FramesCount();
            for (size_t v = 0; v < nv; ++v) {
                int r = ReducedIndex(v);
                if (r < 0) continue;
                cols->push_back(Col{r, {Fdense_(pnt_ind * 3 + 0, v), Fdense_(pnt_ind * 3 + 1, v), Fdense_(pnt_ind * 3 + 2, v)}});
            }
        } else {
            for (size_t oi = pt_begin_[pnt_ind]; oi < pt_begin_[pnt_ind + 1]; ++oi) {
                const F* blk = &Fblk_[oi * 30];
                for (size_t fv = 0; fv < kV; ++fv) {
                    int r = ReducedIndex(obs_[oi].frame * kV + fv);
                    if (r < 0) continue;
                    cols->push_back(Col{r, {blk[0 * 10 + fv], blk[1 * 10 + fv], blk[2 * 10 + fv]}});
                }
            }
        }
    }

    // SparseThreaded variant of the function below: same per-point arithmetic (damped block, |det| rule, A^T E^-1 A, A^T E^-1 g),
    // accumulated per thread into the lower block triangle of camera-pair 10x10 blocks, merged in thread order into a skyline
    // (row-profile) matrix, factored by a skyline Cholesky in double, no refinement.
    bool EstimateCorrectionsThreaded(F hessian_factor, long* skipped_points) {
        const size_t N = PointsCount(), M = FramesCount(), nf = NormalizedFrameVarsCount();
        const int nt = OracleThreads();
        typedef std::array<double, 100> Blk;
        struct Local { std::unordered_map<uint64_t, Blk> blocks; std::unordered_map<uint32_t, std::array<double, 10>> rhs; long skipped = 0; };
        std::vector<Local> loc((size_t)nt);
        std::vector<double> einv(N * 9, 0.0);
        ParallelFor(0, nt, 1, [&](int64_t t) {
            Local& L = loc[(size_t)t];
            const size_t p0 = N * (size_t)t / (size_t)nt, p1 = N * ((size_t)t + 1) / (size_t)nt;
            std::vector<double> tmp;
            for (size_t p = p0; p < p1; ++p) {
                Mat33<Acc> Einv;
                const bool ok = ScaledPointHessianInverse(p, hessian_factor, &Einv);
                skipped_mask_[p] = ok ? 0 : 1;
                if (!ok) { ++L.skipped; continue; }
                for (int i = 0; i < 9; ++i) einv[p * 9 + i] = (double)Einv.a[i];
                const size_t ob = pt_begin_[p], oe = pt_begin_[p + 1];
                tmp.assign((oe - ob) * 30, 0.0);
                for (size_t oi = ob; oi < oe; ++oi) {          // tmp_i = F_i^T E^-1   [10 x 3]
                    const F* b = &Fblk_[oi * 30];
                    double* ti = &tmp[(oi - ob) * 30];
                    for (int a = 0; a < 10; ++a)
                        for (int k = 0; k < 3; ++k) ti[a * 3 + k] = (b[a] * (double)Einv(0, k) + b[10 + a] * (double)Einv(1, k)) + b[20 + a] * (double)Einv(2, k);
                }
                const double g0 = gradE_[p * 3], g1 = gradE_[p * 3 + 1], g2 = gradE_[p * 3 + 2];
                for (size_t oi = ob; oi < oe; ++oi) {
                    const double* ti = &tmp[(oi - ob) * 30];
                    auto& r = L.rhs[obs_[oi].frame];
                    for (int a = 0; a < 10; ++a) r[(size_t)a] += (ti[a * 3] * g0 + ti[a * 3 + 1] * g1) + ti[a * 3 + 2] * g2;
                    for (size_t oj = ob; oj <= oi; ++oj) {    // observations are frame-ascending inside a track: frame(oj) <= frame(oi)
                        const F* bj = &Fblk_[oj * 30];
                        Blk& blk = L.blocks[((uint64_t)obs_[oi].frame << 32) | obs_[oj].frame];
                        for (int a = 0; a < 10; ++a)
                            for (int b2 = 0; b2 < 10; ++b2) blk[(size_t)(a * 10 + b2)] += (ti[a * 3] * bj[b2] + ti[a * 3 + 1] * bj[10 + b2]) + ti[a * 3 + 2] * bj[20 + b2];
                    }
                }
            }
        });
        long skipped = 0;
        for (const Local& L : loc) skipped += L.skipped;
        if (skipped_points != nullptr) *skipped_points = skipped;
        // row profile of the reduced system from the union of the block structure (+ the diagonal blocks)
        std::vector<int> first_cam(M);
        for (size_t f = 0; f < M; ++f) first_cam[f] = (int)f;
        for (const Local& L : loc) for (const auto& kv : L.blocks) { int i = (int)(kv.first >> 32), j = (int)(kv.first & 0xffffffffu); if (j < first_cam[(size_t)i]) first_cam[(size_t)i] = j; }
        auto first_red = [&](size_t cam) { for (size_t v = 0; v < kV; ++v) { int r = ReducedIndex(cam * kV + v); if (r >= 0) return r; } return 0; };
        std::vector<size_t> rowptr(nf + 1, 0);
        std::vector<int> first(nf, 0);
        for (size_t cam = 0; cam < M; ++cam)
            for (size_t v = 0; v < kV; ++v) { int r = ReducedIndex(cam * kV + v); if (r >= 0) first[(size_t)r] = first_red((size_t)first_cam[cam]); }
        for (size_t r = 0; r < nf; ++r) rowptr[r + 1] = rowptr[r] + (size_t)((int)r - first[r] + 1);
        std::vector<double> A(rowptr[nf], 0.0);                      // row r holds columns first[r] .. r
        auto at = [&](int r, int c) -> double& { return A[rowptr[(size_t)r] + (size_t)(c - first[(size_t)r])]; };
        for (size_t cam = 0; cam < M; ++cam)                          // fill_matG (BA.cpp:1780-1823)
            for (size_t v1 = 0; v1 < kV; ++v1) {
                int r1 = ReducedIndex(cam * kV + v1); if (r1 < 0) continue;
                for (size_t v2 = 0; v2 <= v1; ++v2) {
                    int r2 = ReducedIndex(cam * kV + v2); if (r2 < 0) continue;
                    F g = G_(cam * kV + v1, v2);
                    if (v1 == v2) g *= F(1) + hessian_factor;
                    at(r1, r2) = (double)g;
                }
            }
        std::vector<double> b(nf, 0.0);
        for (const Local& L : loc) {                                  // thread order: deterministic for a given thread count
            for (const auto& kv : L.blocks) {
                const size_t ci = (size_t)(kv.first >> 32), cj = (size_t)(kv.first & 0xffffffffu);
                for (size_t a = 0; a < kV; ++a) {
                    int r = ReducedIndex(ci * kV + a); if (r < 0) continue;
                    for (size_t b2 = 0; b2 < kV; ++b2) {
                        int c2 = ReducedIndex(cj * kV + b2); if (c2 < 0 || c2 > r) continue;
                        at(r, c2) -= kv.second[a * 10 + b2];
                    }
                }
            }
            for (const auto& kv : L.rhs) for (size_t a = 0; a < kV; ++a) { int r = ReducedIndex((size_t)kv.first * kV + a); if (r >= 0) b[(size_t)r] += kv.second[a]; }
        }
        for (size_t v = 0; v < kV * M; ++v) { int r = ReducedIndex(v); if (r >= 0) b[(size_t)r] -= (double)gradE_[N * 3 + v]; }
        for (size_t i = 0; i < nf; ++i) rhs_[i] = Acc(b[i]);
        // skyline Cholesky, row by row
        for (int i = 0; i < (int)nf; ++i) {
            double* li = &A[rowptr[(size_t)i]]; const int fi = first[(size_t)i];
            for (int j = fi; j <= i; ++j) {
                const double* lj = &A[rowptr[(size_t)j]]; const int fj = first[(size_t)j];
                const int k0 = fi > fj ? fi : fj;
                double sacc = li[j - fi];
                for (int k = k0; k < j; ++k) sacc -= li[k - fi] * lj[k - fj];
                if (j < i) li[j - fi] = sacc / lj[j - fj];
                else { if (!(sacc > 0)) return false; li[j - fi] = std::sqrt(sacc); }
            }
        }
        std::vector<double> y(b);
        for (int i = 0; i < (int)nf; ++i) { const double* li = &A[rowptr[(size_t)i]]; const int fi = first[(size_t)i]; double sacc = y[(size_t)i]; for (int k = fi; k < i; ++k) sacc -= li[k - fi] * y[(size_t)k]; y[(size_t)i] = sacc / li[i - fi]; }
        for (int i = (int)nf - 1; i >= 0; --i) { const double* li = &A[rowptr[(size_t)i]]; const int fi = first[(size_t)i]; const double xi = y[(size_t)i] / li[i - fi]; y[(size_t)i] = xi; for (int k = fi; k < i; ++k) y[(size_t)k] -= li[k - fi] * xi; }
        for (double v : y) if (!std::isfinite(v)) return false;
        // back substitution (BA.cpp:1919-1960), per point
        std::vector<F> normalized(NormalizedVarsCount(), F(0));
        std::vector<unsigned char> bad((size_t)nt, 0);
        ParallelFor(0, nt, 1, [&](int64_t t) {
            const size_t p0 = N * (size_t)t / (size_t)nt, p1 = N * ((size_t)t + 1) / (size_t)nt;
            for (size_t p = p0; p < p1; ++p) {
                if (skipped_mask_[p]) continue;
                double tt[3] = {0, 0, 0};
                for (size_t oi = pt_begin_[p]; oi < pt_begin_[p + 1]; ++oi) {
                    const F* blk = &Fblk_[oi * 30];
                    for (size_t fv = 0; fv < kV; ++fv) { int r = ReducedIndex(obs_[oi].frame * kV + fv); if (r < 0) continue; for (int k = 0; k < 3; ++k) tt[k] += blk[(size_t)k * 10 + fv] * y[(size_t)r]; }
                }
                for (int k = 0; k < 3; ++k) tt[k] += gradE_[p * 3 + (size_t)k];
                const double* ei = &einv[p * 9];
                for (int r = 0; r < 3; ++r) {   // Mat33 is column-major: (r, c) -> a[c * 3 + r]
                    double sacc = (ei[0 * 3 + r] * tt[0] + ei[1 * 3 + r] * tt[1]) + ei[2 * 3 + r] * tt[2];
                    normalized[p * 3 + (size_t)r] = F(-sacc);
                    if (!std::isfinite(sacc)) bad[(size_t)t] = 1;
                }
            }
        });
        for (unsigned char f : bad) if (f) return false;
        for (size_t i = 0; i < nf; ++i) normalized[N * 3 + i] = F(y[i]);
        FillCorrectionsGapsFromNormalized(normalized);
        return true;
    }

    // BA.cpp:1771-1995
    bool EstimateCorrectionsDecomposedInTwoPhases(F hessian_factor, long* skipped_points = nullptr) {
        if (schur_flow == SchurFlow::SparseThreaded) return EstimateCorrectionsThreaded(hessian_factor, skipped_points);
        size_t N = PointsCount();
        size_t nf = NormalizedFrameVarsCount();
        FillMatG(hessian_factor);
        std::fill(rhs_.begin(), rhs_.end(), Acc(0));
        std::vector<Col> cols;
        std::vector<Acc> tmp;  // (A^T Einv) rows: per column 3 values
        long skipped = 0;
        for (size_t pnt_ind = 0; pnt_ind < N; ++pnt_ind) {
            Mat33<Acc> Einv;
            bool ok = ScaledPointHessianInverse(pnt_ind, hessian_factor, &Einv);
            skipped_mask_[pnt_ind] = ok ? 0 : 1;
            if (!ok) { ++skipped; continue; }
            PointColumns(pnt_ind, &cols);
            size_t nc = cols.size();
            tmp.resize(nc * 3);
            // tmp = A^T * Einv   [n_f x 3]
            for (size_t i = 0; i < nc; ++i)
                for (size_t k = 0; k < 3; ++k)
                    tmp[i * 3 + k] = (Acc(cols[i].a[0]) * Acc(Einv(0, k)) + Acc(cols[i].a[1]) * Acc(Einv(1, k))) + Acc(cols[i].a[2]) * Acc(Einv(2, k));
            // left_side -= tmp * A   (BA.cpp:1891-1892; zero columns contribute exact zeros)
            for (size_t j = 0; j < nc; ++j) {
                Acc a0 = cols[j].a[0], a1 = cols[j].a[1], a2 = cols[j].a[2];
                size_t cj = (size_t)cols[j].r;
                for (size_t i = 0; i < nc; ++i) {
                    Acc v = (tmp[i * 3 + 0] * a0 + tmp[i * 3 + 1] * a1) + tmp[i * 3 + 2] * a2;
                    S_((size_t)cols[i].r, cj) -= v;
                }
            }
            // right_side += tmp * gradE_point   (BA.cpp:1895-1897)
            Acc g0 = gradE_[pnt_ind * 3], g1 = gradE_[pnt_ind * 3 + 1], g2 = gradE_[pnt_ind * 3 + 2];
            for (size_t i = 0; i < nc; ++i) rhs_[(size_t)cols[i].r] += (tmp[i * 3 + 0] * g0 + tmp[i * 3 + 1] * g1) + tmp[i * 3 + 2] * g2;
        }
        if (skipped_points != nullptr) *skipped_points = skipped;
        // right_side -= normalized frame derivatives (BA.cpp:1902-1908)
        size_t nv = kV * FramesCount();
        for (size_t v = 0; v < nv; ++v) {
            int r = ReducedIndex(v);
            if (r >= 0) rhs_[(size_t)r] -= Acc(gradE_[N * 3 + v]);
        }
        // solve (BA.cpp:1911-1913)
        std::vector<F> corrections_frame;
        if (!SolveReduced(&corrections_frame)) return false;
        for (F v : corrections_frame) if (!std::isfinite((double)v)) return false;

        // back substitution (BA.cpp:1919-1960)
        std::vector<F> normalized(NormalizedVarsCount(), F(0));
        for (size_t pnt_ind = 0; pnt_ind < N; ++pnt_ind) {
            Mat33<Acc> Einv;
            bool ok = ScaledPointHessianInverse(pnt_ind, hessian_factor, &Einv);
            F d[3] = {0, 0, 0};
            if (ok) {
                PointColumns(pnt_ind, &cols);
                Acc t[3] = {Acc(0), Acc(0), Acc(0)};
                for (const Col& c : cols) for (size_t k = 0; k < 3; ++k) t[k] += Acc(c.a[k]) * Acc(corrections_frame[(size_t)c.r]);
                for (size_t k = 0; k < 3; ++k) t[k] += Acc(gradE_[pnt_ind * 3 + k]);
                for (size_t r = 0; r < 3; ++r) {
                    Acc s = (Acc(Einv(r, 0)) * t[0] + Acc(Einv(r, 1)) * t[1]) + Acc(Einv(r, 2)) * t[2];
                    d[r] = F(-s);
                    if (!std::isfinite((double)d[r])) return false;
                }
            }
            for (size_t k = 0; k < 3; ++k) normalized[pnt_ind * 3 + k] = d[k];
        }
        for (size_t i = 0; i < nf; ++i) normalized[N * 3 + i] = corrections_frame[i];
        FillCorrectionsGapsFromNormalized(normalized);
        return true;
    }

    bool SolveReduced(std::vector<F>* x) {
        size_t nf = S_.rows;
        if (solve_impl == SolveImpl::None) { x->assign(nf, F(0)); return true; }
        if (solve_impl == SolveImpl::HouseholderQR) {
            DynMat<F> A; A.resize(nf, nf);
            for (size_t i = 0; i < nf * nf; ++i) A.d[i] = F(S_.d[i]);
            std::vector<F> b(nf);
            for (size_t i = 0; i < nf; ++i) b[i] = F(rhs_[i]);
            HouseholderQrSolve(A, b, x);
            return true;
        }
        // CholeskyRefined: factor in long double (or plain double when chol_in_double, the GPU engine's scheme),
        // refine with the Acc-precision S and a W-precision residual.
        if (chol_in_double) return SolveReducedCholesky<double>(x);
        return SolveReducedCholesky<long double>(x);
    }
    template <class W>
    bool SolveReducedCholesky(std::vector<F>* x) {
        size_t nf = S_.rows;
        DynMat<W> L; L.resize(nf, nf);
        for (size_t i = 0; i < nf * nf; ++i) L.d[i] = W(S_.d[i]);
        if (!CholeskyInplace(L)) { x->assign(nf, std::numeric_limits<F>::quiet_NaN()); return true; }
        std::vector<W> b(nf), xs, r(nf), dx;
        for (size_t i = 0; i < nf; ++i) b[i] = W(rhs_[i]);
        CholeskySolve(L, b, &xs);
        for (int it = 0; it < refine_steps; ++it) {
            for (size_t i = 0; i < nf; ++i) { W s = b[i]; for (size_t j = 0; j < nf; ++j) s -= W(S_(i, j)) * xs[j]; r[i] = s; }
            CholeskySolve(L, r, &dx);
            for (size_t i = 0; i < nf; ++i) xs[i] += dx[i];
        }
        x->resize(nf);
        for (size_t i = 0; i < nf; ++i) (*x)[i] = F(xs[i]);
        return true;
    }

    // BA.cpp:1600-1679 — re-insert the 7 fixed zeros (quirk Q13)
    void FillCorrectionsGapsFromNormalized(const std::vector<F>& normalized) {
        size_t N3 = PointsCount() * 3;
        for (size_t i = 0; i < N3; ++i) corrections_[i] = normalized[i];
        size_t nv = kV * FramesCount();
        for (size_t v = 0; v < nv; ++v) {
            int r = ReducedIndex(v);
            corrections_[N3 + v] = r < 0 ? F(0) : normalized[N3 + (size_t)r];
        }
    }

    // BA.cpp:59-92
    static void IncrementRotMat(const Mat33<F>& R, const Vec3<F>& w_delta, Mat33<F>* Rnew) {
        Mat33<F> rot_w;
        if (RotMatFromAxisAngle(w_delta, &rot_w)) *Rnew = rot_w * R;
        else *Rnew = R;  // quirk Q6
    }

    // BA.cpp:1997-2063
    void ApplyCorrections() {
        size_t N = PointsCount(), M = FramesCount();
        size_t pnt_ind = (size_t)-1;
        for (size_t tid = 0; tid < track_rep_->CornerTracksCount(); ++tid) {
            const CornerTrack<F>& track = track_rep_->GetPointTrackById(tid);
            if (!track.SalientPointId.has_value()) continue;
            pnt_ind += 1;
            Vec3<F>& X = map_->GetSalientPoint(track.SalientPointId.value());
            for (size_t k = 0; k < 3; ++k) X[k] += corrections_[pnt_ind * 3 + k];
        }
        for (size_t frame_ind = 0; frame_ind < M; ++frame_ind) {
            size_t off = N * 3 + frame_ind * kV;
            if (apply_intrinsics && Ks_ != nullptr) {  // NOT the reference behaviour: BA.cpp:2027 edits a dropped copy (quirk Q2)
                Mat33<F>& K = (*Ks_)[frame_ind];
                K(0, 0) += corrections_[off]; K(1, 1) += corrections_[off + 1]; K(0, 2) += corrections_[off + 2]; K(1, 2) += corrections_[off + 3];
            }
            off += kIntrinsicVarsCount;
            SE3<F>& inv_cam = (*inverse_orient_cams_)[frame_ind];
            SE3<F> direct = SE3Inv(inv_cam);
            for (size_t k = 0; k < 3; ++k) direct.T[k] += corrections_[off + k];
            off += kTVarsCount;
            Vec3<F> dW(corrections_[off], corrections_[off + 1], corrections_[off + 2]);
            Mat33<F> newR;
            IncrementRotMat(direct.R, dW, &newR);
            direct.R = newR;
            inv_cam = SE3Inv(direct);
        }
    }

    // BA.cpp:143-162, :179-199, :203-247 (SceneNormalizer)
    SE3<F> prenorm_cam0_from_world_;
    F world_scale_ = 0;
    bool NormalizeWorldInplace() {
        using std::abs;
        const SE3<F>& cam0 = (*inverse_orient_cams_)[0];
        const SE3<F>& cam1 = (*inverse_orient_cams_)[1];
        SE3<F> cam0_from1 = SE3AFromB(cam0, cam1);
        F shift = cam0_from1.T[unity_t1_comp_ind_];
        F atol = F(1e-5);
        if (IsClose<F>(F(0), shift, atol)) return false;  // quirk Q4: 1e-5 lands in the rtol slot
        world_scale_ = unity_t1_comp_value_ / abs(shift);
        prenorm_cam0_from_world_ = cam0;
        const Mat33<F> R0t = Transpose(prenorm_cam0_from_world_.R);
        for (SE3<F>& rt : *inverse_orient_cams_) {
            SE3<F> n;
            n.R = rt.R * R0t;
            n.T = (rt.T - (rt.R * R0t) * prenorm_cam0_from_world_.T) * world_scale_;
            rt = n;
        }
        for (auto& sp : map_->SalientPoints()) {
            Vec3<F> x = SE3Apply(prenorm_cam0_from_world_, sp.coord.value());
            sp.coord = x * world_scale_;
        }
        return true;
    }
    // BA.cpp:249-270, :164-177, :187-191
    void RevertNormalization() {
        for (auto& sp : map_->SalientPoints()) {
            Vec3<F> tmp = sp.coord.value() * (F(1) / world_scale_);
            sp.coord = Transpose(prenorm_cam0_from_world_.R) * (tmp - prenorm_cam0_from_world_.T);
        }
        for (SE3<F>& rt : *inverse_orient_cams_) {
            SE3<F> r;
            r.R = rt.R * prenorm_cam0_from_world_.R;
            r.T = rt.T / world_scale_ + rt.R * prenorm_cam0_from_world_.T;
            rt = r;
        }
    }
    F WorldScale() const { return world_scale_; }

    void Bind(F f0, FragmentMap<F>& map, std::vector<SE3<F>>& cams, const CornerTrackRepository<F>& track_rep,
              const Mat33<F>* shared_K, std::vector<Mat33<F>>* Ks) {
        if (!((shared_K != nullptr) ^ (Ks != nullptr))) throw std::invalid_argument("Provide either shared K or separate K for each camera frame");
        f0_ = f0; map_ = &map; inverse_orient_cams_ = &cams; track_rep_ = &track_rep; shared_K_ = shared_K; Ks_ = Ks;
        vars_count_per_frame_ = kIntrinsicVarsCount + kTVarsCount + kWVarsCount;  // quirk Q14
        InitializeNormalizedVarIndices();
        BuildFlatView();
    }

    // BA.cpp:617-718
    bool ComputeInplace(F f0, FragmentMap<F>& map, std::vector<SE3<F>>& cams, const CornerTrackRepository<F>& track_rep,
                        const Mat33<F>* shared_K, std::vector<Mat33<F>>* Ks, const TermCriteria<F>& term_crit) {
        if (!(unity_t1_comp_ind_ < kTVarsCount)) throw std::invalid_argument("Can normalize only one of [T1x, T1y, Tz] components");
        Bind(f0, map, cams, track_rep, shared_K, Ks);
        optimization_stop_reason_.clear();
        trace = Trace<F>();
        if (!NormalizeWorldInplace()) { trace.stop_reason = kStopNormalizationFailed; return false; }
        EnsureMemoryAllocated();
        bool ok = ComputeOnNormalizedWorld(term_crit);
        RevertNormalization();
        return ok;
    }

    // BA.cpp:720-893
    bool ComputeOnNormalizedWorld(const TermCriteria<F>& term_crit) {
        using std::abs;
        F hessian_factor = F(0.0001f);  // quirk Q1: float literal widened
        size_t seen = 0;
        F err_initial = CurrentReprojError(&seen);
        trace.err_initial = err_initial; trace.seen_points = seen;
        std::optional<F> err_thresh = term_crit.allowed_reproj_err_rel_change;
        if (err_thresh.has_value() && err_initial < err_thresh.value()) {
            optimization_stop_reason_ = StopReasonString(kStopAbsErrThreshold); trace.stop_reason = kStopAbsErrThreshold; trace.converged = true;
            return true;
        }
        F err_value = err_initial;
        size_t it = 1;
        while (true) {
            if (max_outer_iters != 0 && it > max_outer_iters) {
                optimization_stop_reason_ = StopReasonString(kStopMaxIters); trace.stop_reason = kStopMaxIters;
                return false;
            }
            ComputeCloseFormReprErrorDerivatives();  // once per outer iteration (quirk Q7)

            // try_decrease_targ_fun (BA.cpp:764-852)
            enum { Success, FailedHessianOverflow, FailedButConverged } result;
            F err_new = std::numeric_limits<F>::quiet_NaN();
            {
                FragmentMap<F> map_copy = *map_;
                std::optional<std::vector<Mat33<F>>> Ks_copy;
                if (Ks_ != nullptr) Ks_copy = *Ks_;
                std::vector<SE3<F>> cams_copy = *inverse_orient_cams_;
                std::optional<F> err_new_prev;
                while (true) {
                    long skipped = 0;
                    bool suc = EstimateCorrectionsDecomposedInTwoPhases(hessian_factor, &skipped);
                    if (!suc) { result = FailedHessianOverflow; break; }
                    ApplyCorrections();
                    err_new = CurrentReprojError();
                    F change = err_new - err_value;
                    bool decreased = change < 0;
                    trace.attempts.push_back(AttemptRecord<F>{hessian_factor, err_new, decreased ? 1 : 0, skipped});
                    if (decreased) { result = Success; break; }
                    *map_ = map_copy;
                    if (Ks_copy.has_value()) *Ks_ = Ks_copy.value();
                    *inverse_orient_cams_ = cams_copy;
                    if (err_new_prev.has_value() && err_thresh.has_value()) {
                        F ch = err_new - err_new_prev.value();
                        if (abs(ch) < err_thresh.value()) { result = FailedButConverged; break; }
                    }
                    hessian_factor *= 10;
                    if (term_crit.max_hessian_factor.has_value() && hessian_factor > term_crit.max_hessian_factor.value()) { result = FailedHessianOverflow; break; }
                    err_new_prev = err_new;
                }
            }
            trace.outer_iters = it;
            if (result != Success) {
                int r = result == FailedHessianOverflow ? kStopHessianOverflow : kStopErrConverged;
                optimization_stop_reason_ = StopReasonString(r); trace.stop_reason = r;
                return false;
            }
            trace.err_per_iter.push_back(err_new);
            trace.factor_per_iter.push_back(hessian_factor);
            F err_value_change = err_new - err_value;
            if (err_thresh.has_value() && abs(err_value_change) < err_thresh.value()) {
                optimization_stop_reason_ = StopReasonString(kStopSmallErrChange); trace.stop_reason = kStopSmallErrChange; trace.converged = true;
                return true;
            }
            err_value = err_new;
            hessian_factor /= 10;
            it += 1;
        }
    }

    // BA.cpp:1551-1598 + :1700-1769 — naive full solve, kept as a cross-check only (dense flow, tiny scenes).
    bool EstimateCorrectionsNaive(F hessian_factor, std::vector<F>* corrections_with_gaps) {
        size_t N = PointsCount(), M = FramesCount();
        size_t n = VarsCount();
        DynMat<F> H; H.resize(n, n);
        for (size_t p = 0; p < N; ++p)
            for (size_t r = 0; r < 3; ++r) {
                for (size_t c = 0; c < 3; ++c) H(p * 3 + r, p * 3 + c) = E_(p * 3 + r, c);
                for (size_t v = 0; v < kV * M; ++v) { H(p * 3 + r, N * 3 + v) = Fdense_(p * 3 + r, v); H(N * 3 + v, p * 3 + r) = Fdense_(p * 3 + r, v); }
            }
        for (size_t f = 0; f < M; ++f)
            for (size_t r = 0; r < kV; ++r) for (size_t c = 0; c < kV; ++c) H(N * 3 + f * kV + r, N * 3 + f * kV + c) = G_(f * kV + r, c);
        for (size_t i = 0; i < n; ++i) H(i, i) *= F(1) + hessian_factor;
        // remove gauge rows/cols
        std::vector<size_t> keep;
        for (size_t i = 0; i < n; ++i) {
            bool removed = false;
            for (size_t k = 0; k < normalized_var_indices_count_; ++k) if (N * 3 + normalized_var_indices_[k] == i) removed = true;
            if (!removed) keep.push_back(i);
        }
        size_t m = keep.size();
        DynMat<F> A; A.resize(m, m);
        std::vector<F> b(m);
        for (size_t i = 0; i < m; ++i) { b[i] = -gradE_[keep[i]]; for (size_t j = 0; j < m; ++j) A(i, j) = H(keep[i], keep[j]); }
        std::vector<F> x;
        HouseholderQrSolve(A, b, &x);
        corrections_with_gaps->assign(n, F(0));
        for (size_t i = 0; i < m; ++i) (*corrections_with_gaps)[keep[i]] = x[i];
        return true;
    }
};

}  // namespace srk_oracle
