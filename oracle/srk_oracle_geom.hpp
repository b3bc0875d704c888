// TEST INFRASTRUCTURE ONLY — CPU oracle for the suriko-engine bundle-adjustment hot path.
// Nothing under oracle/ is linked, imported or executed by the product (surikatoko_b200/, include/);
// only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it.
//
// This header restates, Eigen-free, the geometry layer the BA path depends on.  All paths are relative
// to /root/reference/cpp_impl/suriko-engine.  Matrices are column-major like Eigen's default so that
// SE3<double> is byte-compatible with suriko::SE3Transform (include/suriko/obs-geom.h:177-190).
//
// Third-party arithmetic not present under /root/reference: Eigen3 (version unpinned,
// cpp_impl/CMakeLists.txt:47).  Fixed-size products are restated with the natural left-to-right
// coefficient order; the result is not guaranteed bit-identical to any particular Eigen build.
#pragma once
#include <algorithm>
#include <array>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <functional>
#include <optional>
#include <stdexcept>
#include <string>
#include <vector>

namespace srk_oracle {

// include/suriko/approx-alg.h:7-16 — numpy-like isclose, but with max(a,b) instead of max(|a|,|b|).
template <class F>
inline bool IsClose(F a, F b, F rtol = F(1.0e-5), F atol = F(1.0e-8)) {
    using std::abs;
    return abs(a - b) <= (atol + rtol * abs(std::max<F>(a, b)));
}

template <class F>
struct Vec3 {
    F v[3] = {F(0), F(0), F(0)};
    Vec3() = default;
    Vec3(F x, F y, F z) : v{x, y, z} {}
    F operator[](size_t i) const { return v[i]; }
    F& operator[](size_t i) { return v[i]; }
};

template <class F> inline Vec3<F> operator+(const Vec3<F>& a, const Vec3<F>& b) { return {a[0] + b[0], a[1] + b[1], a[2] + b[2]}; }
template <class F> inline Vec3<F> operator-(const Vec3<F>& a, const Vec3<F>& b) { return {a[0] - b[0], a[1] - b[1], a[2] - b[2]}; }
template <class F> inline Vec3<F> operator-(const Vec3<F>& a) { return {-a[0], -a[1], -a[2]}; }
template <class F> inline Vec3<F> operator*(const Vec3<F>& a, F s) { return {a[0] * s, a[1] * s, a[2] * s}; }
template <class F> inline Vec3<F> operator*(F s, const Vec3<F>& a) { return {s * a[0], s * a[1], s * a[2]}; }
template <class F> inline Vec3<F> operator/(const Vec3<F>& a, F s) { return {a[0] / s, a[1] / s, a[2] / s}; }
template <class F> inline F Dot(const Vec3<F>& a, const Vec3<F>& b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
template <class F> inline Vec3<F> Cross(const Vec3<F>& a, const Vec3<F>& b) {
    return {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
}
template <class F> inline F Norm(const Vec3<F>& a) { using std::sqrt; return sqrt(Dot(a, a)); }

// 3x3, column-major: (r,c) -> a[c*3+r]
template <class F>
struct Mat33 {
    F a[9] = {F(0), F(0), F(0), F(0), F(0), F(0), F(0), F(0), F(0)};
    F operator()(size_t r, size_t c) const { return a[c * 3 + r]; }
    F& operator()(size_t r, size_t c) { return a[c * 3 + r]; }
    static Mat33 Identity() { Mat33 m; m(0, 0) = m(1, 1) = m(2, 2) = F(1); return m; }
    Vec3<F> col(size_t c) const { return {a[c * 3], a[c * 3 + 1], a[c * 3 + 2]}; }
};

template <class F> inline Mat33<F> operator*(const Mat33<F>& A, const Mat33<F>& B) {
    Mat33<F> C;
    for (size_t r = 0; r < 3; ++r)
        for (size_t c = 0; c < 3; ++c) C(r, c) = A(r, 0) * B(0, c) + A(r, 1) * B(1, c) + A(r, 2) * B(2, c);
    return C;
}
template <class F> inline Vec3<F> operator*(const Mat33<F>& A, const Vec3<F>& x) {
    return {A(0, 0) * x[0] + A(0, 1) * x[1] + A(0, 2) * x[2],
            A(1, 0) * x[0] + A(1, 1) * x[1] + A(1, 2) * x[2],
            A(2, 0) * x[0] + A(2, 1) * x[1] + A(2, 2) * x[2]};
}
template <class F> inline Mat33<F> operator*(const Mat33<F>& A, F s) { Mat33<F> C; for (int i = 0; i < 9; ++i) C.a[i] = A.a[i] * s; return C; }
template <class F> inline Mat33<F> operator*(F s, const Mat33<F>& A) { Mat33<F> C; for (int i = 0; i < 9; ++i) C.a[i] = s * A.a[i]; return C; }
template <class F> inline Mat33<F> operator+(const Mat33<F>& A, const Mat33<F>& B) { Mat33<F> C; for (int i = 0; i < 9; ++i) C.a[i] = A.a[i] + B.a[i]; return C; }
template <class F> inline Mat33<F> operator-(const Mat33<F>& A, const Mat33<F>& B) { Mat33<F> C; for (int i = 0; i < 9; ++i) C.a[i] = A.a[i] - B.a[i]; return C; }
template <class F> inline Mat33<F> Transpose(const Mat33<F>& A) {
    Mat33<F> C;
    for (size_t r = 0; r < 3; ++r) for (size_t c = 0; c < 3; ++c) C(r, c) = A(c, r);
    return C;
}
template <class F> inline F Det(const Mat33<F>& m) {
    return m(0, 0) * (m(1, 1) * m(2, 2) - m(1, 2) * m(2, 1)) - m(0, 1) * (m(1, 0) * m(2, 2) - m(1, 2) * m(2, 0)) +
           m(0, 2) * (m(1, 0) * m(2, 1) - m(1, 1) * m(2, 0));
}
template <class F> inline F FrobNorm(const Mat33<F>& m) { using std::sqrt; F s = 0; for (int i = 0; i < 9; ++i) s += m.a[i] * m.a[i]; return sqrt(s); }

// Eigen LU/InverseImpl.h (third-party, not in /root/reference): computeInverseAndDetWithCheck for 3x3 —
// cofactor inverse, determinant = cofactors(col 0) . matrix(col 0), invertible <=> |det| > threshold
// (default NumTraits<double>::dummy_precision() = 1e-12).  Call sites: src/bundle-adj-kanatani.cpp:1876, :1936.
template <class F>
inline void Inverse3x3WithCheck(const Mat33<F>& m, F abs_det_threshold, Mat33<F>* inv, F* det, bool* invertible) {
    using std::abs;
    auto cof = [&m](size_t i, size_t j) {
        size_t i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
        return m(i1, j1) * m(i2, j2) - m(i1, j2) * m(i2, j1);
    };
    F c00 = cof(0, 0), c10 = cof(1, 0), c20 = cof(2, 0);
    *det = (c00 * m(0, 0) + c10 * m(1, 0)) + c20 * m(2, 0);
    *invertible = abs(*det) > abs_det_threshold;
    if (!*invertible) return;
    F invdet = F(1) / *det;
    Mat33<F>& r = *inv;
    r(0, 0) = c00 * invdet; r(0, 1) = c10 * invdet; r(0, 2) = c20 * invdet;
    r(1, 0) = cof(0, 1) * invdet; r(1, 1) = cof(1, 1) * invdet; r(1, 2) = cof(2, 1) * invdet;
    r(2, 0) = cof(0, 2) * invdet; r(2, 1) = cof(1, 2) * invdet; r(2, 2) = cof(2, 2) * invdet;
}

// include/suriko/obs-geom.h:177-190 — T first, then R (column-major).
template <class F>
struct SE3 {
    Vec3<F> T;
    Mat33<F> R = Mat33<F>::Identity();
};

// src/obs-geom.cpp:117-150
template <class F> inline SE3<F> SE3Inv(const SE3<F>& rt) { SE3<F> r; r.R = Transpose(rt.R); r.T = -(r.R * rt.T); return r; }
template <class F> inline Vec3<F> SE3Apply(const SE3<F>& rt, const Vec3<F>& x) { return rt.R * x + rt.T; }
template <class F> inline SE3<F> SE3Compose(const SE3<F>& a, const SE3<F>& b) { SE3<F> r; r.R = a.R * b.R; r.T = a.R * b.T + a.T; return r; }
template <class F> inline SE3<F> SE3AFromB(const SE3<F>& a_from_world, const SE3<F>& b_from_world) { return SE3Compose(a_from_world, SE3Inv(b_from_world)); }

// src/obs-geom.cpp:512-518
template <class F> inline Mat33<F> SkewSymmetricMat(const Vec3<F>& v) {
    Mat33<F> s;
    s(0, 0) = 0; s(0, 1) = -v[2]; s(0, 2) = v[1];
    s(1, 0) = v[2]; s(1, 1) = 0; s(1, 2) = -v[0];
    s(2, 0) = -v[1]; s(2, 1) = v[0]; s(2, 2) = 0;
    return s;
}

// src/obs-geom.cpp:520-551 (release build: kSurikoDebug=false)
template <class F>
inline bool RotMatFromUnityDirAndAngle(const Vec3<F>& unity_dir, F ang, Mat33<F>* rot_mat, bool check_input = true) {
    using std::sin; using std::cos;
    if (check_input) {
        F dir_len = Norm(unity_dir);
        if (!IsClose<F>(F(1), dir_len)) return false;
        if (IsClose<F>(F(0), ang)) return false;
    }
    F s = sin(ang), c = cos(ang);
    Mat33<F> skew1 = SkewSymmetricMat(unity_dir);
    *rot_mat = Mat33<F>::Identity() + s * skew1 + (F(1) - c) * skew1 * skew1;  // ((1-c)*skew)*skew, left to right
    return true;
}

// src/obs-geom.cpp:553-561
template <class F>
inline bool RotMatFromAxisAngle(const Vec3<F>& axis_angle, Mat33<F>* rot_mat) {
    F ang = Norm(axis_angle);
    if (IsClose<F>(F(0), ang)) return false;
    Vec3<F> unity_dir = axis_angle / ang;
    return RotMatFromUnityDirAndAngle(unity_dir, ang, rot_mat, false);
}

// src/obs-geom.cpp:563-596 (input check skipped as in release with check_input=false; the float literals
// 0.5f / 1.0f / 1e-3f of the source widen exactly except atol=1e-3f).
template <class F>
inline bool LogSO3(const Mat33<F>& rot_mat, Vec3<F>* unity_dir, F* ang) {
    using std::sqrt; using std::acos;
    F cos_ang = F(0.5f) * ((rot_mat(0, 0) + rot_mat(1, 1) + rot_mat(2, 2)) - F(1));
    cos_ang = std::clamp<F>(cos_ang, F(-1), F(1));
    F sin_ang = sqrt(F(1.0f) - cos_ang * cos_ang);
    F atol = F(1e-3f);
    if (IsClose<F>(F(0), sin_ang, F(0), atol)) return false;
    Vec3<F>& u = *unity_dir;
    u[0] = rot_mat(2, 1) - rot_mat(1, 2);
    u[1] = rot_mat(0, 2) - rot_mat(2, 0);
    u[2] = rot_mat(1, 0) - rot_mat(0, 1);
    u = u * (F(0.5f) / sin_ang);
    F dirlen = Norm(u);
    u = u * (F(1) / dirlen);
    *ang = acos(cos_ang);
    return true;
}

template <class F>
inline bool AxisAngleFromRotMat(const Mat33<F>& rot_mat, Vec3<F>* dir) {
    Vec3<F> u; F ang;
    if (!LogSO3(rot_mat, &u, &ang)) return false;
    *dir = u * ang;
    return true;
}

// src/obs-geom.cpp:418-441
template <class F>
inline bool IsIdentity(const Mat33<F>& M, F rtol, F atol) {
    for (size_t r = 0; r < 3; ++r)
        for (size_t c = 0; c < 3; ++c)
            if (!IsClose<F>(r == c ? F(1) : F(0), M(r, c), rtol, atol)) return false;
    return true;
}

// src/obs-geom.cpp:443-487.  Eigen's isIdentity(prec) (third-party): off-diagonals negligible w.r.t. the
// diagonal and diagonal approx 1, both at precision prec.
template <class F>
inline bool IsSpecialOrthogonal(const Mat33<F>& R) {
    using std::abs;
    Mat33<F> rtr = Transpose(R) * R;
    F prec = F(1.0e-3f);
    for (size_t c = 0; c < 3; ++c)
        for (size_t r = 0; r < 3; ++r) {
            if (r == c) { if (!(abs(rtr(r, c) - F(1)) <= prec * std::min<F>(abs(rtr(r, c)), F(1)))) return false; }
            else { if (!(abs(rtr(r, c)) <= prec)) return false; }
        }
    return IsClose<F>(F(1.0f), Det(R), F(1.0e-3f), F(1.0e-3f));
}

// ------------------------------------------------------------------------------------------------------
// Data model at the BA boundary (include/suriko/obs-geom.h:199-304, src/obs-geom.cpp:152-416).

template <class F>
struct Point2 { F v[2] = {F(0), F(0)}; Point2() = default; Point2(F x, F y) : v{x, y} {} F operator[](size_t i) const { return v[i]; } F& operator[](size_t i) { return v[i]; } };

template <class F>
struct SalientPointFragment {
    std::optional<size_t> synthetic_virtual_point_id;
    std::optional<Vec3<F>> coord;
};

template <class F>
class FragmentMap {
    std::vector<SalientPointFragment<F>> salient_points_;
    size_t fragment_id_offset_;
    size_t next_salient_point_id_;
public:
    explicit FragmentMap(size_t fragment_id_offset = 1000'000)
        : fragment_id_offset_(fragment_id_offset), next_salient_point_id_(fragment_id_offset + 1) {}
    // src/obs-geom.cpp:158-174
    SalientPointFragment<F>& AddSalientPointTempl(const std::optional<Vec3<F>>& coord, size_t* salient_point_id = nullptr) {
        size_t new_id = next_salient_point_id_++;
        if (salient_point_id != nullptr) *salient_point_id = new_id;
        salient_points_.resize(salient_points_.size() + 1);
        salient_points_.back().coord = coord;
        return salient_points_.back();
    }
    // src/obs-geom.cpp:247-256
    size_t SalientPointIdToInd(size_t id) const { return id - fragment_id_offset_ - 1; }
    size_t SalientPointIndToId(size_t ind) const { return ind + fragment_id_offset_ + 1; }
    const Vec3<F>& GetSalientPoint(size_t id) const {
        size_t ind = SalientPointIdToInd(id);
        if (ind >= salient_points_.size()) throw std::out_of_range("CHECK(ind < salient_points_.size())");
        return salient_points_[ind].coord.value();
    }
    Vec3<F>& GetSalientPoint(size_t id) {
        size_t ind = SalientPointIdToInd(id);
        if (ind >= salient_points_.size()) throw std::out_of_range("CHECK(ind < salient_points_.size())");
        return salient_points_[ind].coord.value();
    }
    const SalientPointFragment<F>& GetSalientPointNew(size_t id) const { return salient_points_.at(SalientPointIdToInd(id)); }
    size_t SalientPointsCount() const { return salient_points_.size(); }
    const std::vector<SalientPointFragment<F>>& SalientPoints() const { return salient_points_; }
    std::vector<SalientPointFragment<F>>& SalientPoints() { return salient_points_; }
    void GetSalientPointsIds(std::vector<size_t>* ids) const { for (size_t i = 0; i < salient_points_.size(); ++i) ids->push_back(SalientPointIndToId(i)); }
};

template <class F>
struct CornerData { Point2<F> pixel_coord; Vec3<F> image_coord; };

template <class F>
class CornerTrack {
public:
    size_t TrackId = 0;
private:
    ptrdiff_t StartFrameInd = -1;
    std::vector<std::optional<CornerData<F>>> CoordPerFramePixels;
public:
    std::optional<size_t> SalientPointId;
    std::optional<size_t> SyntheticVirtualPointId;

    bool HasCorners() const { return StartFrameInd != -1; }
    size_t CornersCount() const { return CoordPerFramePixels.size(); }

    // src/obs-geom.cpp:277-292 — push_back: the k-th added corner is reported at frame Start+k (quirk Q11).
    void AddCorner(size_t frame_ind, const Point2<F>& value) {
        if (StartFrameInd == -1) StartFrameInd = (ptrdiff_t)frame_ind;
        else if (!((size_t)StartFrameInd <= frame_ind)) throw std::logic_error("Can insert points later than the initial (start) frame");
        CornerData<F> cd; cd.pixel_coord = value;
        CoordPerFramePixels.push_back(std::optional<CornerData<F>>(cd));
    }
    // src/obs-geom.cpp:294-314 — resize variant, leaves gaps as nullopt.
    CornerData<F>& AddCorner(size_t frame_ind) {
        if (StartFrameInd == -1) StartFrameInd = (ptrdiff_t)frame_ind;
        else if (!((size_t)StartFrameInd <= frame_ind)) throw std::logic_error("Can insert points later than the initial (start) frame");
        ptrdiff_t local_ind = (ptrdiff_t)frame_ind - StartFrameInd;
        CoordPerFramePixels.resize(local_ind + 1);
        CoordPerFramePixels.back() = std::optional<CornerData<F>>(CornerData<F>{});
        return CoordPerFramePixels.back().value();
    }
    // src/obs-geom.cpp:316-328
    std::optional<Point2<F>> GetCorner(size_t frame_ind) const {
        if (StartFrameInd == -1) throw std::logic_error("CHECK(StartFrameInd != -1)");
        ptrdiff_t local_ind = (ptrdiff_t)frame_ind - StartFrameInd;
        if (local_ind < 0 || (size_t)local_ind >= CoordPerFramePixels.size()) return std::nullopt;
        const auto& cd = CoordPerFramePixels[local_ind];
        if (!cd.has_value()) return std::nullopt;
        return cd.value().pixel_coord;
    }
    // src/obs-geom.cpp:348-355
    void EachCorner(const std::function<void(size_t, const std::optional<CornerData<F>>&)>& on_item) const {
        for (size_t i = 0; i < CoordPerFramePixels.size(); ++i) on_item((size_t)StartFrameInd + i, CoordPerFramePixels[i]);
    }
};

template <class F>
class CornerTrackRepository {
public:
    std::vector<CornerTrack<F>> CornerTracks;
    CornerTrack<F>& AddCornerTrackObj() {  // src/obs-geom.cpp:398-405
        CornerTrack<F> t; t.TrackId = CornerTracks.size();
        CornerTracks.push_back(t);
        return CornerTracks.back();
    }
    size_t CornerTracksCount() const { return CornerTracks.size(); }
    size_t ReconstructedCornerTracksCount() const { size_t n = 0; for (const auto& t : CornerTracks) if (t.SalientPointId.has_value()) ++n; return n; }
    const CornerTrack<F>& GetPointTrackById(size_t id) const { return CornerTracks[id]; }
    CornerTrack<F>& GetPointTrackById(size_t id) { return CornerTracks[id]; }
};

}  // namespace srk_oracle
