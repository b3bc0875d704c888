// suriko-b200 — Eigen-free mirrors of the reference's data model at the BA boundary, for the drop-in adapter in
// bundle-adj-kanatani.h.  Same names, member functions and semantics as
//   /root/reference/cpp_impl/suriko-engine/include/suriko/obs-geom.h:177-304  and  src/obs-geom.cpp:117-416
// so that code written against the reference's FragmentMap / CornerTrackRepository / SE3Transform compiles against these
// when Eigen is not available.  In a build that HAS the reference headers, use those and only the adapter's flattening.
#pragma once
#include <array>
#include <cstddef>
#include <cstdint>
#include <functional>
#include <optional>
#include <stdexcept>
#include <vector>

namespace suriko_compat {

using Scalar = double;  // rt-config.h:41-48 (the f32 build option is out of scope)

struct Point2f { Scalar v[2] = {0, 0}; Point2f() = default; Point2f(Scalar x, Scalar y) : v{x, y} {} Scalar operator[](size_t i) const { return v[i]; } Scalar& operator[](size_t i) { return v[i]; } };
struct Point3 { Scalar v[3] = {0, 0, 0}; Point3() = default; Point3(Scalar x, Scalar y, Scalar z) : v{x, y, z} {} Scalar operator[](size_t i) const { return v[i]; } Scalar& operator[](size_t i) { return v[i]; } };

// Column-major 3x3 like Eigen::Matrix<Scalar,3,3>: (r,c) -> a[c*3+r].
struct Mat33 {
    Scalar a[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    Scalar operator()(size_t r, size_t c) const { return a[c * 3 + r]; }
    Scalar& operator()(size_t r, size_t c) { return a[c * 3 + r]; }
    static Mat33 Identity() { Mat33 m; m(0, 0) = m(1, 1) = m(2, 2) = 1; return m; }
};

// obs-geom.h:177-190: T first, then R — 12 contiguous doubles, byte-compatible with srk_ba_problem::cams.
struct SE3Transform {
    Point3 T;
    Mat33 R = Mat33::Identity();
    SE3Transform() = default;
    SE3Transform(const Mat33& R_, const Point3& T_) : T(T_), R(R_) {}
    static SE3Transform NoTransform() { return SE3Transform(); }
};
static_assert(sizeof(SE3Transform) == 12 * sizeof(Scalar), "SE3Transform must be T[3] + R[9]");

struct SalientPointFragment {  // obs-geom.h:199-204
    std::optional<size_t> synthetic_virtual_point_id;
    std::optional<Point3> coord;
    void* user_obj = nullptr;
};

class FragmentMap {  // obs-geom.h:207-243, obs-geom.cpp:152-256
    std::vector<SalientPointFragment> salient_points_;
    size_t fragment_id_offset_;
    size_t next_salient_point_id_;
public:
    explicit FragmentMap(size_t fragment_id_offset = 1000'000) : fragment_id_offset_(fragment_id_offset), next_salient_point_id_(fragment_id_offset + 1) {}
    SalientPointFragment& AddSalientPointTempl(const std::optional<Point3>& coord, size_t* salient_point_id = nullptr) {
        size_t new_id = next_salient_point_id_++;
        if (salient_point_id != nullptr) *salient_point_id = new_id;
        salient_points_.resize(salient_points_.size() + 1);
        salient_points_.back().coord = coord;
        return salient_points_.back();
    }
    void SetSalientPoint(size_t point_track_id, const Point3& coord) { salient_points_.at(point_track_id).coord = coord; }
    const SalientPointFragment& GetSalientPointNew(size_t id) const { return salient_points_.at(SalientPointIdToInd(id)); }
    SalientPointFragment& GetSalientPointNew(size_t id) { return salient_points_.at(SalientPointIdToInd(id)); }
    const Point3& GetSalientPoint(size_t id) const { return salient_points_.at(SalientPointIdToInd(id)).coord.value(); }
    Point3& GetSalientPoint(size_t id) { return salient_points_.at(SalientPointIdToInd(id)).coord.value(); }
    size_t SalientPointsCount() const { return salient_points_.size(); }
    const std::vector<SalientPointFragment>& SalientPoints() const { return salient_points_; }
    std::vector<SalientPointFragment>& SalientPoints() { return salient_points_; }
    void GetSalientPointsIds(std::vector<size_t>* ids) const { for (size_t i = 0; i < salient_points_.size(); ++i) ids->push_back(SalientPointIndToId(i)); }
    size_t SalientPointIdToInd(size_t id) const { return id - fragment_id_offset_ - 1; }
    size_t SalientPointIndToId(size_t ind) const { return ind + fragment_id_offset_ + 1; }
};

struct CornerData { Point2f pixel_coord; Point3 image_coord; };

class CornerTrack {  // obs-geom.h:251-283, obs-geom.cpp:258-355
public:
    size_t TrackId = 0;
private:
    ptrdiff_t StartFrameInd = -1;
    std::vector<std::optional<CornerData>> CoordPerFramePixels;
public:
    std::optional<size_t> SalientPointId;
    std::optional<size_t> SyntheticVirtualPointId;
    bool HasCorners() const { return StartFrameInd != -1; }
    size_t CornersCount() const { return CoordPerFramePixels.size(); }
    // push_back variant: the k-th added corner is reported at frame Start+k (gaps collapse) — obs-geom.cpp:277-292
    void AddCorner(size_t frame_ind, const Point2f& value) {
        CheckStart(frame_ind);
        CornerData cd; cd.pixel_coord = value;
        CoordPerFramePixels.push_back(std::optional<CornerData>(cd));
    }
    // resize variant: gaps stay empty — obs-geom.cpp:294-314
    CornerData& AddCorner(size_t frame_ind) {
        CheckStart(frame_ind);
        ptrdiff_t local_ind = (ptrdiff_t)frame_ind - StartFrameInd;
        CoordPerFramePixels.resize((size_t)local_ind + 1);
        CoordPerFramePixels.back() = std::optional<CornerData>(CornerData{});
        return CoordPerFramePixels.back().value();
    }
    std::optional<CornerData> GetCornerData(size_t frame_ind) const {
        if (StartFrameInd == -1) throw std::logic_error("CHECK(StartFrameInd != -1)");
        ptrdiff_t local_ind = (ptrdiff_t)frame_ind - StartFrameInd;
        if (local_ind < 0 || (size_t)local_ind >= CoordPerFramePixels.size()) return std::nullopt;
        return CoordPerFramePixels[(size_t)local_ind];
    }
    std::optional<Point2f> GetCorner(size_t frame_ind) const {
        auto cd = GetCornerData(frame_ind);
        if (!cd.has_value()) return std::nullopt;
        return cd.value().pixel_coord;
    }
    void EachCorner(const std::function<void(size_t, const std::optional<CornerData>&)>& on_item) const {
        for (size_t i = 0; i < CoordPerFramePixels.size(); ++i) on_item((size_t)StartFrameInd + i, CoordPerFramePixels[i]);
    }
private:
    void CheckStart(size_t frame_ind) {
        if (StartFrameInd == -1) StartFrameInd = (ptrdiff_t)frame_ind;
        else if (!((size_t)StartFrameInd <= frame_ind)) throw std::logic_error("Can insert points later than the initial (start) frame");
    }
};

class CornerTrackRepository {  // obs-geom.h:285-304, obs-geom.cpp:357-416
public:
    std::vector<CornerTrack> CornerTracks;
    CornerTrack& AddCornerTrackObj() { CornerTrack t; t.TrackId = CornerTracks.size(); CornerTracks.push_back(t); return CornerTracks.back(); }
    size_t CornerTracksCount() const { return CornerTracks.size(); }
    size_t ReconstructedCornerTracksCount() const { size_t n = 0; for (const auto& t : CornerTracks) if (t.SalientPointId.has_value()) ++n; return n; }
    size_t FramesCount() const { return CornerTracks.empty() ? 0 : CornerTracks[0].CornersCount(); }
    const CornerTrack& GetPointTrackById(size_t id) const { return CornerTracks[id]; }
    CornerTrack& GetPointTrackById(size_t id) { return CornerTracks[id]; }
};

}  // namespace suriko_compat
