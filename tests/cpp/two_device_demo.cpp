// C++17 host, ONE process, TWO devices, no Python anywhere: one thread per device, each with its own engine handle; the library binds
// NCCL itself (srk_nccl_unique_id + srk_ba_nccl_init, include/srk/ba_c_api.h).  A small ring scene is sharded over the two handles
// (contiguous point ranges, cameras replicated) and solved; the result must match ONE handle on the whole scene and both
// ranks must hold bit-identical cameras.  Also exercises the per-device kernel attributes (PerDeviceOnce): the second device launches
// every > 48 KB shared-memory kernel of the library.
// Prints "err_initial err_final_sharded err_final_whole rel_dev identical"; exit code 0 = pass, 77 = fewer than two sm_100 devices.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>
#include <cuda_runtime_api.h>
#include "../../include/srk/ba_c_api.h"

struct Scene { int M, N, k; std::vector<int32_t> cam, pt; std::vector<double> xy, points, cams, K; double f0; };

static Scene ring(int M, int N, int k) {
    Scene s; s.M = M; s.N = N; s.k = k; s.f0 = 600.0;
    const double Kn[9] = {880 / 600.0, 0, 0, 0, 660 / 600.0, 0, 400 / 600.0, 300 / 600.0, 1};   // column-major
    unsigned long long st = 88172645463325252ull;
    auto rnd = [&]() { st ^= st << 13; st ^= st >> 7; st ^= st << 17; return (double)(st >> 11) / 9007199254740992.0; };
    std::vector<double> R(9 * M), T(3 * M);
    for (int m = 0; m < M; ++m) {
        const double th = 2 * M_PI * m / M, pos[3] = {10 * std::cos(th), 10 * std::sin(th), 2.0 + 0.6 * (m % 4)};
        double f[3] = {-pos[0], -pos[1], -pos[2]}; const double fn = std::sqrt(f[0] * f[0] + f[1] * f[1] + f[2] * f[2]);
        for (double& v : f) v /= fn;
        double x[3] = {-f[1], f[0], 0}; const double xn = std::sqrt(x[0] * x[0] + x[1] * x[1]);
        for (double& v : x) v /= xn;
        const double y[3] = {f[1] * x[2] - f[2] * x[1], f[2] * x[0] - f[0] * x[2], f[0] * x[1] - f[1] * x[0]};
        for (int c = 0; c < 3; ++c) { R[9 * m + 3 * c + 0] = x[c]; R[9 * m + 3 * c + 1] = y[c]; R[9 * m + 3 * c + 2] = f[c]; }   // column-major rows x, y, f
        for (int r = 0; r < 3; ++r) T[3 * m + r] = -(R[9 * m + r] * pos[0] + R[9 * m + 3 + r] * pos[1] + R[9 * m + 6 + r] * pos[2]);
    }
    std::vector<double> gt(3 * N);
    for (int j = 0; j < N; ++j) {
        const double phi = 2 * M_PI * (j + 0.5) / N, rho = 6 * std::sqrt(0.05 + 0.95 * rnd());
        gt[3 * j] = rho * std::cos(phi); gt[3 * j + 1] = rho * std::sin(phi); gt[3 * j + 2] = 2 * rnd() - 1;
        const int centre = (int)std::floor(phi / (2 * M_PI) * M + 0.5);
        std::vector<int> cs;
        for (int i = 0; i < k; ++i) cs.push_back(((centre + i - k / 2) % M + M) % M);
        std::sort(cs.begin(), cs.end());
        for (int c : cs) {
            double xc[3];
            for (int r = 0; r < 3; ++r) xc[r] = R[9 * c + r] * gt[3 * j] + R[9 * c + 3 + r] * gt[3 * j + 1] + R[9 * c + 6 + r] * gt[3 * j + 2] + T[3 * c + r];
            s.cam.push_back(c); s.pt.push_back(j);
            s.xy.push_back(600.0 * (Kn[0] * xc[0] / xc[2] + Kn[6]) + 0.5 * (rnd() - 0.5));
            s.xy.push_back(600.0 * (Kn[4] * xc[1] / xc[2] + Kn[7]) + 0.5 * (rnd() - 0.5));
        }
    }
    s.points.resize(3 * N);
    for (int i = 0; i < 3 * N; ++i) s.points[i] = gt[i] + 0.03 * (rnd() - 0.5);
    s.cams.resize(12 * M);
    for (int m = 0; m < M; ++m) { std::memcpy(&s.cams[12 * m], &T[3 * m], 24); std::memcpy(&s.cams[12 * m + 3], &R[9 * m], 72); }
    s.K.resize(9 * M);
    for (int m = 0; m < M; ++m) std::memcpy(&s.K[9 * m], Kn, 72);
    return s;
}

struct Result { int rc = 0; srk_ba_report rep{}; std::vector<double> cams, points; };

static void solve(int device, const Scene& s, int p0, int p1, const unsigned char* nccl_id, int rank, int world, Result* out) {
    void* h = nullptr;
    out->rc = srk_ba_create(&h, &device, 1);
    if (out->rc != 0) return;
    if (world > 1) out->rc = srk_ba_nccl_init(h, nccl_id, rank, world);
    if (out->rc == 0) {
        const int64_t o0 = (int64_t)p0 * s.k, o1 = (int64_t)p1 * s.k;
        std::vector<int32_t> pt(s.pt.begin() + o0, s.pt.begin() + o1);
        for (int32_t& v : pt) v -= p0;
        out->points.assign(s.points.begin() + 3 * (size_t)p0, s.points.begin() + 3 * (size_t)p1);
        out->cams = s.cams;
        srk_ba_problem pr{};
        pr.n_cams = s.M; pr.n_points = p1 - p0; pr.n_obs = o1 - o0;
        pr.obs_cam = s.cam.data() + o0; pr.obs_point = pt.data(); pr.obs_xy = s.xy.data() + 2 * o0;
        pr.points = out->points.data(); pr.cams = out->cams.data(); pr.K = s.K.data(); pr.shared_K = 0; pr.f0 = s.f0;
        srk_ba_options opt; srk_ba_default_options(&opt);
        opt.has_err_change = 1; opt.err_change = 1e-10; opt.max_outer_iters = 4;
        out->rc = srk_ba_solve(h, &pr, &opt, &out->rep);
        if (out->rc != 0) std::fprintf(stderr, "rank %d: %s\n", rank, srk_last_error());
    } else std::fprintf(stderr, "rank %d: %s\n", rank, srk_last_error());
    srk_ba_destroy(h);
}

int main() {
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count < 2) { std::printf("skip: %d device(s)\n", count); return 77; }
    const Scene s = ring(170, 4000, 6);    // n_f = 1693: the nested-dissection order (camera graph = union over ranks) is on the path
    unsigned char id[SRK_NCCL_UNIQUE_ID_BYTES];
    if (srk_nccl_unique_id(id) != 0) { std::fprintf(stderr, "%s\n", srk_last_error()); return 1; }
    Result r[2], whole;
    std::thread t0(solve, 0, std::cref(s), 0, s.N / 2, id, 0, 2, &r[0]);
    std::thread t1(solve, 1, std::cref(s), s.N / 2, s.N, id, 1, 2, &r[1]);
    t0.join(); t1.join();
    if (r[0].rc != 0 || r[1].rc != 0) return 1;
    solve(1, s, 0, s.N, nullptr, 0, 1, &whole);     // the whole scene on the SECOND device, same process
    if (whole.rc != 0) return 1;
    const bool identical = std::memcmp(r[0].cams.data(), r[1].cams.data(), sizeof(double) * r[0].cams.size()) == 0 && r[0].rep.err_final == r[1].rep.err_final;
    const double dev = std::fabs(r[0].rep.err_final - whole.rep.err_final) / whole.rep.err_final;
    std::printf("%.17g %.17g %.17g %.3e %d\n", r[0].rep.err_initial, r[0].rep.err_final, whole.rep.err_final, dev, identical ? 1 : 0);
    return (identical && dev < 1e-9 && r[0].rep.outer_iters == whole.rep.outer_iters && r[0].rep.err_final < r[0].rep.err_initial) ? 0 : 1;
}
