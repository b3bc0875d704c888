// Drop-in check of include/suriko_compat: a tiny circle scene built with the reference's container API (AddSalientPointTempl,
// AddCornerTrackObj, AddCorner), refined through the adapter.  Prints "err_initial err_final converged reason launches".
#include <cmath>
#include <cstdio>
#include <random>
#include "../../include/suriko_compat/bundle-adj-kanatani.h"
using namespace suriko_compat;

int main() {
    const double f0 = 600.0;
    FragmentMap map;
    CornerTrackRepository tracks;
    std::vector<Point3> gt;
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 8; ++j) gt.push_back(Point3(-1.0 + 0.28 * i, -1.0 + 0.28 * j, 0.3 * std::cos(0.7 * i) + 0.2 * std::sin(0.5 * j)));
    std::mt19937 gen(1234);
    std::uniform_real_distribution<double> dis(-0.004, 0.004);
    for (const Point3& p : gt) {
        size_t id = 0;
        map.AddSalientPointTempl(Point3(p[0] + dis(gen), p[1] + dis(gen), p[2] + dis(gen)), &id);
        tracks.AddCornerTrackObj().SalientPointId = id;
    }
    std::vector<SE3Transform> cams;
    std::vector<Mat33> Ks;
    Mat33 K; K(0, 0) = 880 / f0; K(1, 1) = 660 / f0; K(0, 2) = 400 / f0; K(1, 2) = 300 / f0; K(2, 2) = 1;
    const int M = 12;
    for (int m = 0; m < M; ++m) {
        const double th = 0.3 * m, R = 7.0;
        const double pos[3] = {R * std::cos(th), R * std::sin(th), 3.0 + 0.4 * (m % 3)};
        double f[3] = {-pos[0], -pos[1], -pos[2]};
        const double fn = std::sqrt(f[0] * f[0] + f[1] * f[1] + f[2] * f[2]);
        for (double& v : f) v /= fn;
        double x[3] = {-f[1], f[0], 0.0};   // up x f
        const double xn = std::sqrt(x[0] * x[0] + x[1] * x[1]);
        for (double& v : x) v /= xn;
        const double y[3] = {f[1] * x[2] - f[2] * x[1], f[2] * x[0] - f[0] * x[2], f[0] * x[1] - f[1] * x[0]};
        SE3Transform c;
        for (int k = 0; k < 3; ++k) { c.R(0, k) = x[k]; c.R(1, k) = y[k]; c.R(2, k) = f[k]; }
        for (int r = 0; r < 3; ++r) c.T[r] = -(c.R(r, 0) * pos[0] + c.R(r, 1) * pos[1] + c.R(r, 2) * pos[2]);
        cams.push_back(c); Ks.push_back(K);
        for (size_t p = 0; p < gt.size(); ++p) {
            double xc[3];
            for (int r = 0; r < 3; ++r) xc[r] = c.R(r, 0) * gt[p][0] + c.R(r, 1) * gt[p][1] + c.R(r, 2) * gt[p][2] + c.T[r];
            tracks.GetPointTrackById(p).AddCorner((size_t)m, Point2f(f0 * (K(0, 0) * xc[0] / xc[2] + K(0, 2)), f0 * (K(1, 1) * xc[1] / xc[2] + K(1, 2))));
        }
    }
    BundleAdjustmentKanatani ba;
    size_t seen = 0;
    const double e0 = ba.ReprojError(f0, map, cams, tracks, nullptr, &Ks, &seen);
    BundleAdjustmentKanataniTermCriteria tc; tc.AllowedReprojErrRelativeChange(1e-10);
    ba.max_outer_iters = 20;
    const bool ok = ba.ComputeInplace(f0, map, cams, tracks, nullptr, &Ks, tc);
    const double e1 = ba.ReprojError(f0, map, cams, tracks, nullptr, &Ks);
    std::printf("%.17g %.17g %d \"%s\" %lld %zu %.6f\n", e0, e1, ok ? 1 : 0, ba.OptimizationStatusString().c_str(), (long long)ba.LastReport().gpu_launches, seen,
                ba.ReprojErrorPixPerPoint(e1, seen));
    // e1 is re-evaluated on the un-normalised scene: at convergence on exact pixels it is a sum of ~1e-14 squares whose own rounding
    // noise (2 |rho| eps per term) is ~1e-9 relative, hence the 1e-6
    return (e1 < e0 && std::fabs(e1 - ba.LastReport().err_final) <= 1e-6 * e1) ? 0 : 1;
}
