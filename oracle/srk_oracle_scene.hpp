// TEST INFRASTRUCTURE ONLY — CPU oracle: synthetic input generators restated from the reference's demos.
//   GenerateCircleCameraShots : /root/reference/cpp_impl/suriko-engine/src/virt-world/scene-generator.cpp:9-55
//   CircleGridScene           : /root/reference/cpp_impl/demos/demo-bundle-adj-circle-grid.cpp:64-257
// std::mt19937 + std::uniform_real_distribution<double> (libstdc++) are used exactly as the demo does
// (seed 1234, draw order :109-128 then :224-257).
#pragma once
#include <random>
#include "srk_oracle_geom.hpp"

namespace srk_oracle {

// scene-generator.cpp:9-55.  The reference multiplies 4x4 matrices SE3Mat(R?,t?) * cam_from_world; for
// [R t; 0 1] operands that product is SE3Compose (the extra +t*0 terms are exact zeros).
template <class F>
inline void GenerateCircleCameraShots(const Vec3<F>& circle_center, F circle_radius, F ascentZ, const std::vector<F>& rot_angles,
                                      std::vector<SE3<F>>* inverse_orient_cams) {
    using std::cos; using std::sin; using std::acos; using std::atan2; using std::sqrt;
    for (F ang : rot_angles) {
        SE3<F> cam_from_world;  // identity
        Vec3<F> shift = circle_center;
        Vec3<F> center_to_cam_pos(circle_radius * cos(ang), circle_radius * sin(ang), ascentZ);
        shift = shift + center_to_cam_pos;
        SE3<F> tr; tr.T = -shift;
        cam_from_world = SE3Compose(tr, cam_from_world);

        Vec3<F> to_center(-shift[0], -shift[1], F(0));
        to_center = to_center / Norm(to_center);  // Eigen normalize(): v /= norm
        Vec3<F> oy(0, 1, 0), oz(0, 0, 1);
        F ang_yawOY = acos(Dot(oy, to_center));
        int sign = Dot(Cross(oy, to_center), oz) >= 0 ? 1 : -1;  // approx-alg.h Sign
        ang_yawOY *= F(sign);

        auto rot_or_identity = [](const Vec3<F>& dir, F a) {  // obs-geom.cpp internals::RotMat
            Mat33<F> r;
            if (!RotMatFromUnityDirAndAngle(dir, a, &r)) r = Mat33<F>::Identity();
            return r;
        };
        SE3<F> yaw; yaw.R = rot_or_identity(oz, -ang_yawOY);
        cam_from_world = SE3Compose(yaw, cam_from_world);

        F look_down_ang = atan2(center_to_cam_pos[2], sqrt(center_to_cam_pos[0] * center_to_cam_pos[0] + center_to_cam_pos[1] * center_to_cam_pos[1] + F(0)));
        SE3<F> pitch; pitch.R = rot_or_identity(Vec3<F>(1, 0, 0), look_down_ang + F(M_PI) / 2);
        cam_from_world = SE3Compose(pitch, cam_from_world);
        inverse_orient_cams->push_back(cam_from_world);
    }
}

template <class F>
struct CircleGridParams {  // gflags defaults of demo-bundle-adj-circle-grid.cpp:46-62
    F f0 = 600;
    F world_xmin = -1, world_xmax = 1, world_ymin = -1, world_ymax = 1, world_zmin = 0, world_zmax = 1;
    F cell_x = 0.5, cell_y = 0.5;
    F ang_start = -M_PI / 2 + M_PI / 6, ang_end = 2 * M_PI / 3, ang_step = M_PI / 180 * 5;
    F noise_R_hi = 0.005, noise_x3D_hi = 0.005;
    F rot_radius = -1;  // < 0: 15*cell_x as in the demo (:89)
    F ascentZ = -1;     // < 0: 10*cell_x as in the demo (:90)
    unsigned seed = 1234;
};

template <class F>
struct Scene {
    F f0 = 600;
    FragmentMap<F> map;                  // noisy points handed to BA
    std::vector<SE3<F>> cams;            // noisy inverse poses handed to BA
    std::vector<Mat33<F>> Ks;            // per-frame intrinsics (rows 0,1 divided by f0)
    CornerTrackRepository<F> tracks;
    std::vector<Vec3<F>> gt_points;
    std::vector<SE3<F>> gt_cams;
};

template <class F>
inline void CircleGridScene(const CircleGridParams<F>& p, Scene<F>* out) {
    using std::cos;
    Scene<F>& sc = *out;
    sc = Scene<F>();
    sc.f0 = p.f0;
    F rot_radius = p.rot_radius < 0 ? 15 * p.cell_x : p.rot_radius;
    F ascentZ = p.ascentZ < 0 ? 10 * p.cell_x : p.ascentZ;
    Vec3<F> circle_center(1, 0.5, 0);
    const F inclusive_gap = 1e-8;

    FragmentMap<F> map;
    size_t next_virtual_id = 1000001;
    F xmid = (p.world_xmin + p.world_xmax) / 2, xlen = p.world_xmax - p.world_xmin, zlen = p.world_zmax - p.world_zmin;
    for (F x = p.world_xmin; x < p.world_xmax + inclusive_gap; x += p.cell_x)
        for (F y = p.world_ymin; y < p.world_ymax + inclusive_gap; y += p.cell_y) {
            F val_z = cos((x - xmid) / xlen * F(M_PI));
            F z = p.world_zmin + val_z * zlen;
            auto& sp = map.AddSalientPointTempl(Vec3<F>(x, y, z));
            sp.synthetic_virtual_point_id = next_virtual_id++;
        }
    for (const auto& sp : map.SalientPoints()) sc.gt_points.push_back(sp.coord.value());

    std::mt19937 gen;
    gen.seed(p.seed);

    sc.map = map;
    if (p.noise_x3D_hi > 0) {
        std::uniform_real_distribution<F> dis(p.noise_x3D_hi / 2, p.noise_x3D_hi);
        for (auto& frag : sc.map.SalientPoints()) {
            Vec3<F>& pnt = frag.coord.value();
            F d1 = dis(gen), d2 = dis(gen), d3 = dis(gen);
            pnt[0] += d1; pnt[1] += d2; pnt[2] += d3;
        }
    }
    std::vector<size_t> ids;
    sc.map.GetSalientPointsIds(&ids);
    for (size_t id : ids) {
        auto& track = sc.tracks.AddCornerTrackObj();
        track.SalientPointId = id;
        track.SyntheticVirtualPointId = sc.map.GetSalientPointNew(id).synthetic_virtual_point_id;
    }

    // K = diag(1/f0,1/f0,1) * [[880,0,W/2],[0,660,H/2],[0,0,1]], W x H = 800 x 600   (:149-163)
    Mat33<F> K;
    K(0, 0) = (1 / p.f0) * 880; K(0, 2) = (1 / p.f0) * (800 / 2.0);
    K(1, 1) = (1 / p.f0) * 660; K(1, 2) = (1 / p.f0) * (600 / 2.0);
    K(2, 2) = 1;

    std::vector<F> rot_angles;
    for (F ang = p.ang_start;; ang += p.ang_step) {
        if ((p.ang_start < p.ang_end && ang >= p.ang_end) || (p.ang_start > p.ang_end && ang <= p.ang_end)) break;
        rot_angles.push_back(ang);
    }
    GenerateCircleCameraShots(circle_center, rot_radius, ascentZ, rot_angles, &sc.gt_cams);

    for (size_t ang_ind = 0; ang_ind < sc.gt_cams.size(); ++ang_ind) {
        const SE3<F>& rt = sc.gt_cams[ang_ind];
        sc.Ks.push_back(K);
        for (size_t frag_ind = 0; frag_ind < map.SalientPoints().size(); ++frag_ind) {
            // ProjectPnt (:35-44): K * (X_cam / X_cam.z); pixel = f0 * (h.xy / h.z)   — no visibility clipping
            Vec3<F> pc = SE3Apply(rt, map.SalientPoints()[frag_ind].coord.value());
            Vec3<F> img = pc / pc[2];
            Vec3<F> h = K * img;
            Point2<F> pix((h[0] / h[2]) * p.f0, (h[1] / h[2]) * p.f0);
            sc.tracks.GetPointTrackById(frag_ind).AddCorner(ang_ind, pix);
        }
    }

    sc.cams = sc.gt_cams;
    if (p.noise_R_hi > 0) {
        std::uniform_real_distribution<F> dis(0, 1);
        for (SE3<F>& rt : sc.cams) {
            Vec3<F> dir; F ang;
            if (!LogSO3(rt.R, &dir, &ang)) continue;
            F da = dis(gen) * p.noise_R_hi;
            ang += da;
            F dw1 = dis(gen) * p.noise_R_hi, dw2 = dis(gen) * p.noise_R_hi, dw3 = dis(gen) * p.noise_R_hi;
            dir[0] += dw1; dir[1] += dw2; dir[2] += dw3;
            dir = dir / Norm(dir);
            Mat33<F> R;
            if (RotMatFromUnityDirAndAngle(dir, ang, &R)) rt.R = R;
        }
    }
}

}  // namespace srk_oracle
