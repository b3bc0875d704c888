// suriko-b200 — host side of the bundle-adjustment engine behind the C ABI (include/srk/ba_c_api.h).
//
// Mirrors BundleAdjustmentKanatani::ComputeInplace / ComputeOnNormalizedWorld of the reference
// ("BA.cpp" = /root/reference/cpp_impl/suriko-engine/src/bundle-adj-kanatani.cpp): gauge normalisation (BA.cpp:203-247),
// the Levenberg-Marquardt control loop with its accept/reject and stop rules (BA.cpp:720-893), and the revert
// (BA.cpp:249-270).  All per-observation / per-point / per-camera arithmetic runs in the CUDA kernels of ba_kernels.cu,
// chol_kernels.cu and pcg_kernels.cu; the host only owns buffers, launch order and the scalar control flow.
// There is no CPU fallback: without a CUDA device srk_ba_create fails.
#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/srk/ba_c_api.h"
#include "kernels.h"
#include "solve_order.h"
#include "pcg.h"
#include "prep.h"

namespace {

thread_local std::string g_last_error;

void set_error(const std::string& s) { g_last_error = s; }

#define SRK_CUDA(call)                                                                                         \
    do {                                                                                                       \
        cudaError_t e__ = (call);                                                                              \
        if (e__ != cudaSuccess) {                                                                              \
            set_error(std::string(#call) + ": " + cudaGetErrorString(e__));                                    \
            return SRK_E_CUDA;                                                                                 \
        }                                                                                                      \
    } while (0)

// Grow-only device buffer: work buffers are cached on the handle between calls (BA.h:136-163).
struct Buf {
    void* p = nullptr;
    size_t cap = 0;
    Buf() = default;
    Buf(const Buf&) = delete;
    Buf& operator=(const Buf&) = delete;
    ~Buf() { release(); }   // `delete engine` frees every buffer, listed anywhere or not (the destroy call selects the device first)
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap && p != nullptr) return cudaSuccess;
        if (p != nullptr) { cudaFree(p); p = nullptr; cap = 0; }
        size_t want = bytes < 256 ? 256 : bytes;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        // development aid: SRK_POISON=1 fills fresh buffers with NaN patterns, so that a read of memory nobody wrote shows up at once
        // (the fill runs on the legacy default stream, which the engine's non-blocking streams do not wait for: finish it before anybody writes the buffer)
        if (e == cudaSuccess && std::getenv("SRK_POISON") != nullptr) { cudaMemset(p, 0xFF, want); cudaDeviceSynchronize(); }
        return e;
    }
    void release() { if (p != nullptr) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

enum Family { F_JACOBIAN = 0, F_FRAME, F_SCHUR, F_SOLVE, F_BACKSUB, F_UPDATE, F_RESIDUAL, F_ALLREDUCE, F_FACTOR, F_TRSV, F_COUNT };
const char* kFamilyNames[F_COUNT] = {"jacobian", "frame_blocks", "schur", "solve", "backsub", "update", "residual", "allreduce", "solve_factor", "solve_trsv"};

struct FamilyTimer {
    std::vector<cudaEvent_t> pending;  // start, stop, start, stop, ...
    std::vector<cudaEvent_t> pool;
    double total_ms = 0.0, last_ms = 0.0;
    int64_t count = 0;
};

// ---- host 3x3 / SE3 helpers (column-major like Eigen; products left to right as the oracle restates them) ----------
struct M3 { double a[9]; double& operator()(int r, int c) { return a[c * 3 + r]; } double operator()(int r, int c) const { return a[c * 3 + r]; } };
struct Pose { double T[3]; M3 R; };  // byte-compatible with suriko::SE3Transform (obs-geom.h:177-190)
static_assert(sizeof(Pose) == 12 * sizeof(double), "pose must be T[3] + R[9]");

M3 mul(const M3& A, const M3& B) {
    M3 C;
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) C(r, c) = A(r, 0) * B(0, c) + A(r, 1) * B(1, c) + A(r, 2) * B(2, c);
    return C;
}
void mulv(const M3& A, const double* x, double* y) {
    for (int r = 0; r < 3; ++r) y[r] = A(r, 0) * x[0] + A(r, 1) * x[1] + A(r, 2) * x[2];
}
M3 transpose(const M3& A) { M3 C; for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) C(r, c) = A(c, r); return C; }
Pose pose_inv(const Pose& p) {  // obs-geom.cpp:117-122
    Pose r; r.R = transpose(p.R);
    double t[3]; mulv(r.R, p.T, t);
    for (int i = 0; i < 3; ++i) r.T[i] = -t[i];
    return r;
}
Pose pose_compose(const Pose& a, const Pose& b) {  // a * b
    Pose r; r.R = mul(a.R, b.R);
    double t[3]; mulv(a.R, b.T, t);
    for (int i = 0; i < 3; ++i) r.T[i] = t[i] + a.T[i];
    return r;
}
// approx-alg.h:7-16 (max(a,b), not max(|a|,|b|))
bool is_close(double a, double b, double rtol = 1.0e-5, double atol = 1.0e-8) { return std::fabs(a - b) <= (atol + rtol * std::fabs(std::fmax(a, b))); }

struct Engine {
    int device = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t side_stream = nullptr;   // structure pass of a bind, concurrent with the H2D copies on `stream`
    cudaEvent_t ev_idx = nullptr;
    bool G_partial = false;               // multi-GPU dense path: Ggf holds this rank's partial sums (they travel inside S)
    cudaStream_t stream = nullptr;

    bool bound = false, norm_failed = false, normalized = false;
    int64_t N = 0, O = 0;
    int M = 0, shared_K = 0, unity = 1;
    double f0 = 0.0, unity_val = 1.0;
    int nf = 0;
    int64_t ld = 0;
    Pose cam0_prenorm{};
    double world_scale = 1.0;

    // observations (point-major) and the camera-major copy
    Buf obs_cam, obs_pt, obs_xy, ox, oy, pt_begin, cam_cnt, cam_cursor, cam_begin, c_pt, c_x, c_y;
    Buf chunk_cams, chunk_cnt, chunk_pts, obs_slot, obs_pos;
    Buf rows_F, rows_W, rows_D;   // dense-rows form of K2 (schur_dense_rows)
    Buf k2_slot, k2_mask, k2_tab, k2_ntab, k2_gi, k2_u, k2_exc, k2_eacc;   // K2 third form (schur_v3.cu): bind-time tile tables, per-attempt point factors, exception list
    static constexpr int kExcCap = 16384;
    bool k2_eacc_valid = false;   // K1 left the per-point sums of this outer iteration in k2_eacc
    int k2_fuse_k1 = 1;           // SRK_K2_FUSE_K1=0: k_point_factor re-reads rho / Jp instead (cross-check)
    bool schur_dense_rows = false;
    Buf tacc;                 // K2' per-point accumulators [3N]; zero between attempts (k_backsub_finish re-zeroes what it reads)
    bool tacc_zero = false;
    int schur_rows_enabled = 1;  // SRK_SCHUR_ROWS=0: never use the dense-rows form of K2
    int backsub_impl = 1;     // SRK_BACKSUB_IMPL=0: the point-per-half-warp kernel   // K1' structure: per-chunk camera lists, per-observation table slots (k_chunk_tables)
    // state: current / trial / as bound
    Buf Xa, Xb, Xbound, pts_stage, cams_a, cams_b, cams_bound, Kd, camd_a, camd_b;
    double *X_cur = nullptr, *X_try = nullptr, *cams_cur = nullptr, *cams_try = nullptr, *camd_cur = nullptr, *camd_try = nullptr;
    // derivative pass and solve
    Buf J, Ggf, pinv, skipped, deferred, Srhs, Lfac, dinv, xsol, resid, dfull, partial, errsum, slots, flags, skipped_cnt, dbg;
    Buf tmask, tlist, tpacked;   // multi-GPU tile exchange of S
    // nested-dissection order of the reduced camera system (solve_order.h): built lazily at the first dense solve after a bind
    srk::SolveOrder order;
    bool order_ready = false;
    int solve_order_enabled = 1; // SRK_SOLVE_ORDER=0: always factor in capture order (cross-check)
    int solve_tiles_enabled = 1; // SRK_SOLVE_TILES=0: dense passes over S / L even when the tile structure is known (cross-check)
    Buf adjbuf, order_src, xperm, order_ints, order_pattern;
    bool order_tiles = false;    // the tile lists of `order` are on the device and the factorisation will take the cluster path
    int ot_s = 0, ot_resptr = 0, ot_resent = 0, ot_lin = 0, ot_lall = 0, ot_scn = 0;   // offsets into order_ints
    bool S_clean = false;        // S is zero outside order.s_tiles (so the next attempt only clears those)
    std::vector<unsigned char> order_adj;   // the graph `order` was built from
    unsigned char* h_adj = nullptr;         // pinned
    size_t h_adj_cap = 0;
    int tile_exchange = 1;       // SRK_TILE_EXCHANGE=0: always all-reduce the whole dense system
    int64_t n_deferred = 0;   // points the tiled Schur kernel leaves to the per-point kernel (structure only, known at bind time)
    int schur_tile_points = 256;
    int schur_tile_fixed = 0; // SRK_SCHUR_TILE: force the tile size (0 = choose at bind time)
    int schur_impl = 0;       // 0 = single-operand DMMA tile kernel (schur_v3.cu), 2 = two-operand DMMA tile kernel (schur_mma.cu), 1 = vector-FMA tile kernel (ba_kernels.cu)
    srk::PcgWorkspace pcg;
    int residual_blocks = 0;
    double* h_slots = nullptr;  // pinned
    size_t h_slots_cap = 0;

    // multi-GPU plumbing
    srk_allreduce_fn ar = nullptr;
    void* ar_user = nullptr;
    int rank = 0, world = 1;
    void* nccl_comm = nullptr;   // ncclComm_t when the library drives NCCL itself (srk_ba_nccl_init / srk_ba_set_nccl_comm)
    bool nccl_owned = false;

    // accounting
    bool timing = false;
    FamilyTimer timers[F_COUNT];
    int64_t launches = 0;
    int32_t pcg_iters_last = 0;
    int64_t pcg_iters_total = 0;     // since srk_ba_set_timing(h, 1)
    double pcg_rel_res_last = 0.0;
    int solver_used = 0;
    int64_t factor_failures = 0;     // attempts whose Cholesky factorisation met a non-positive pivot (retried with more damping)
    int debug_fail_factor = 0;       // SRK_DEBUG_FAIL_FACTOR=k: treat the first k factorisations of every run as failed (test hook of that path)
};

struct Scope {  // CUDA-event bracket of one kernel family on the engine's stream
    Engine& e; int fam; cudaEvent_t stop = nullptr;
    Scope(Engine& eng, int f) : e(eng), fam(f) {
        if (!e.timing) return;
        FamilyTimer& t = e.timers[fam];
        cudaEvent_t a, b;
        if (t.pool.size() >= 2) { a = t.pool.back(); t.pool.pop_back(); b = t.pool.back(); t.pool.pop_back(); }
        else { cudaEventCreate(&a); cudaEventCreate(&b); }
        cudaEventRecord(a, e.stream);
        t.pending.push_back(a); t.pending.push_back(b);
        stop = b;
    }
    ~Scope() { if (stop != nullptr) cudaEventRecord(stop, e.stream); }
};

void resolve_timers(Engine& e) {
    cudaStreamSynchronize(e.stream);
    for (int f = 0; f < F_COUNT; ++f) {
        FamilyTimer& t = e.timers[f];
        for (size_t i = 0; i + 1 < t.pending.size(); i += 2) {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, t.pending[i], t.pending[i + 1]) == cudaSuccess) { t.total_ms += ms; t.last_ms = ms; t.count += 1; }
            t.pool.push_back(t.pending[i]); t.pool.push_back(t.pending[i + 1]);
        }
        t.pending.clear();
    }
}

// ---- NCCL through dlopen: the library has no link-time dependency on it (single-GPU hosts need none) -------------------------------
struct NcclId { char internal[128]; };                                      // == ncclUniqueId (nccl.h: NCCL_UNIQUE_ID_BYTES 128)
struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(void*) = nullptr;                                    // ncclGetUniqueId(ncclUniqueId*)
    int (*CommInitRank)(void**, int, NcclId, int) = nullptr;         // ncclCommInitRank(ncclComm_t*, nranks, ncclUniqueId, rank)
    int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
NcclApi g_nccl;
bool nccl_load() {
    if (g_nccl.lib != nullptr) return true;
    const char* names[] = {std::getenv("SRK_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    void* lib = nullptr;
    for (const char* nm : names) { if (nm != nullptr && nm[0] != 0 && (lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL)) != nullptr) break; }
    if (lib == nullptr) { set_error("libnccl.so.2 not found (set SRK_NCCL_LIB to its path)"); return false; }
    NcclApi a; a.lib = lib;
    a.GetUniqueId = (int (*)(void*))dlsym(lib, "ncclGetUniqueId");
    a.CommInitRank = (int (*)(void**, int, NcclId, int))dlsym(lib, "ncclCommInitRank");
    a.AllReduce = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t))dlsym(lib, "ncclAllReduce");
    a.CommDestroy = (int (*)(void*))dlsym(lib, "ncclCommDestroy");
    a.GetErrorString = (const char* (*)(int))dlsym(lib, "ncclGetErrorString");
    if (!a.GetUniqueId || !a.CommInitRank || !a.AllReduce || !a.CommDestroy) { set_error("libnccl lacks the expected entry points"); return false; }
    g_nccl = a;
    return true;
}
int nccl_allreduce_cb(void* user, double* dev, int64_t count, void* stream) {
    Engine* e = (Engine*)user;
    if (e == nullptr || e->nccl_comm == nullptr) return 1;
    const int rc = g_nccl.AllReduce(dev, dev, (size_t)count, /*ncclDouble*/ 8, /*ncclSum*/ 0, e->nccl_comm, (cudaStream_t)stream);
    if (rc != 0) { set_error(std::string("ncclAllReduce: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "error")); return 1; }
    return 0;
}

int do_allreduce(Engine& e, double* dev, int64_t count) {
    if (e.ar == nullptr || e.world <= 1) return SRK_OK;
    Scope s(e, F_ALLREDUCE);
    int rc = e.ar(e.ar_user, dev, count, (void*)e.stream);
    if (rc != 0) { set_error("all-reduce callback failed"); return SRK_E_CUDA; }
    return SRK_OK;
}

inline int residual_grid(int64_t O) {
    int64_t b = (O + 255) / 256;
    if (b < 1) b = 1;
    if (b > 148 * 8) b = 148 * 8;  // 8 resident CTAs of 256 threads per SM
    return (int)b;
}

// Upload + index construction (+ NormalizeSceneInplace when normalize != 0).
int pick_solver(const Engine& e, const srk_ba_options* opt);
int ensure_solve_order(Engine& e, cudaStream_t on_stream = nullptr);

int bind_impl(Engine& e, const srk_ba_problem* p, const srk_ba_options* opt, bool normalize) {
    if (p == nullptr || p->n_cams < 0 || p->n_points < 0 || p->n_obs < 0) { set_error("null or negative-sized problem"); return SRK_E_INVALID_ARG; }
    if ((p->n_obs > 0 && (p->obs_cam == nullptr || p->obs_point == nullptr || p->obs_xy == nullptr)) || (p->n_points > 0 && p->points == nullptr) ||
        p->cams == nullptr || p->K == nullptr) { set_error("null array in problem"); return SRK_E_INVALID_ARG; }
    if (is_close(0.0, p->f0)) { set_error("f0 != 0 (BA.cpp:420)"); return SRK_E_INVALID_ARG; }
    if (p->n_cams > (int64_t)200000000 / 10 || p->n_points > (int64_t)2000000000 || p->n_obs > ((int64_t)1 << 40)) { set_error("problem too large for int32 indices"); return SRK_E_TOO_LARGE; }
    int unity = opt != nullptr ? opt->unity_comp_ind : 1;
    double unity_val = opt != nullptr ? opt->unity_comp_value : 1.0;
    if (normalize) {
        if (!(unity >= 0 && unity < 3)) { set_error("Can normalize only one of [T1x, T1y, Tz] components (BA.cpp:628)"); return SRK_E_INVALID_ARG; }
        if (p->n_cams < 2) { set_error("normalisation needs at least two camera frames"); return SRK_E_INVALID_ARG; }
    }
    SRK_CUDA(cudaSetDevice(e.device));
    e.bound = false;
    e.order_ready = false;
    e.pcg.structure_valid = false;
    e.N = p->n_points; e.O = p->n_obs; e.M = (int)p->n_cams; e.shared_K = p->shared_K ? 1 : 0; e.f0 = p->f0;
    e.unity = unity; e.unity_val = unity_val;
    e.nf = e.M * 10 - 7;
    e.ld = ((int64_t)e.nf + 7) & ~(int64_t)7;
    const int64_t N = e.N, O = e.O; const int M = e.M;
    cudaStream_t st = e.stream;

    SRK_CUDA(e.obs_cam.ensure(sizeof(int32_t) * O)); SRK_CUDA(e.obs_pt.ensure(sizeof(int32_t) * O)); SRK_CUDA(e.obs_xy.ensure(sizeof(double) * 2 * O));
    SRK_CUDA(e.ox.ensure(sizeof(double) * O)); SRK_CUDA(e.oy.ensure(sizeof(double) * O));
    SRK_CUDA(e.pt_begin.ensure(sizeof(int64_t) * (N + 1)));
    SRK_CUDA(e.cam_cnt.ensure(sizeof(unsigned long long) * (M + 1))); SRK_CUDA(e.cam_cursor.ensure(sizeof(unsigned long long) * (M + 1)));
    SRK_CUDA(e.cam_begin.ensure(sizeof(int64_t) * (M + 1)));
    SRK_CUDA(e.c_pt.ensure(sizeof(int32_t) * O)); SRK_CUDA(e.c_x.ensure(sizeof(double) * O)); SRK_CUDA(e.c_y.ensure(sizeof(double) * O));
    SRK_CUDA(e.Xa.ensure(sizeof(double) * 3 * N)); SRK_CUDA(e.Xb.ensure(sizeof(double) * 3 * N)); SRK_CUDA(e.Xbound.ensure(sizeof(double) * 3 * N));
    SRK_CUDA(e.pts_stage.ensure(sizeof(double) * 3 * N));
    SRK_CUDA(e.cams_a.ensure(sizeof(double) * 12 * M)); SRK_CUDA(e.cams_b.ensure(sizeof(double) * 12 * M)); SRK_CUDA(e.cams_bound.ensure(sizeof(double) * 12 * M));
    SRK_CUDA(e.Kd.ensure(sizeof(double) * 9 * (e.shared_K ? 1 : M)));
    SRK_CUDA(e.camd_a.ensure(sizeof(double) * 48 * M)); SRK_CUDA(e.camd_b.ensure(sizeof(double) * 48 * M));
    SRK_CUDA(e.flags.ensure(sizeof(int) * 8)); SRK_CUDA(e.skipped_cnt.ensure(sizeof(unsigned long long) * 2));
    SRK_CUDA(e.errsum.ensure(sizeof(double) * 2));
    e.residual_blocks = residual_grid(O);
    SRK_CUDA(e.partial.ensure(sizeof(double) * e.residual_blocks));
    SRK_CUDA(e.slots.ensure(sizeof(double) * (e.world + 3)));
    if (e.h_slots_cap < (size_t)(e.world + 3)) {
        if (e.h_slots != nullptr) cudaFreeHost(e.h_slots);
        SRK_CUDA(cudaMallocHost((void**)&e.h_slots, sizeof(double) * (e.world + 3)));
        e.h_slots_cap = (size_t)(e.world + 3);
    }
    e.X_cur = e.Xa.as<double>(); e.X_try = e.Xb.as<double>();
    e.cams_cur = e.cams_a.as<double>(); e.cams_try = e.cams_b.as<double>();
    e.camd_cur = e.camd_a.as<double>(); e.camd_try = e.camd_b.as<double>();

    // ---- observations: H2D, validation, point CSR, camera-major copy.  The indices (8 B per observation) travel first; the whole
    // structure pass (validation, CSR, camera histogram + scatter positions, K1' chunk tables, and further down the Schur plan) runs on
    // the side stream while the pixels (16 B per observation), the points and K are still crossing PCIe on the main stream.
    cudaStream_t ss = e.side_stream != nullptr ? e.side_stream : st;
    SRK_CUDA(e.obs_pos.ensure(sizeof(unsigned) * (size_t)(O > 0 ? O : 1)));
    {
        const int64_t nch = srk::residual_chunks(O);
        SRK_CUDA(e.chunk_cams.ensure(sizeof(int) * (size_t)(nch > 0 ? nch : 1) * srk::residual_chunk_slots()));
        SRK_CUDA(e.chunk_cnt.ensure(sizeof(int) * (size_t)(nch > 0 ? nch : 1)));
        SRK_CUDA(e.chunk_pts.ensure(sizeof(int) * 2 * (size_t)(nch > 0 ? nch : 1)));
        SRK_CUDA(e.obs_slot.ensure((size_t)(O > 0 ? O : 1)));
    }
    SRK_CUDA(cudaMemsetAsync(e.flags.p, 0, sizeof(int) * 8, st));
    SRK_CUDA(cudaMemsetAsync(e.cam_cnt.p, 0, sizeof(unsigned long long) * (M + 1), st));
    if (O == 0) SRK_CUDA(cudaMemsetAsync(e.pt_begin.p, 0, sizeof(int64_t) * (N + 1), st));
    if (O > 0) {
        SRK_CUDA(cudaMemcpyAsync(e.obs_cam.p, p->obs_cam, sizeof(int32_t) * O, cudaMemcpyHostToDevice, st));
        SRK_CUDA(cudaMemcpyAsync(e.obs_pt.p, p->obs_point, sizeof(int32_t) * O, cudaMemcpyHostToDevice, st));
    }
    if (ss != st) { SRK_CUDA(cudaEventRecord(e.ev_idx, st)); SRK_CUDA(cudaStreamWaitEvent(ss, e.ev_idx, 0)); }
    if (O > 0) SRK_CUDA(cudaMemcpyAsync(e.obs_xy.p, p->obs_xy, sizeof(double) * 2 * O, cudaMemcpyHostToDevice, st));
    if (N > 0) SRK_CUDA(cudaMemcpyAsync(e.pts_stage.p, p->points, sizeof(double) * 3 * N, cudaMemcpyHostToDevice, st));
    SRK_CUDA(cudaMemcpyAsync(e.Kd.p, p->K, sizeof(double) * 9 * (e.shared_K ? 1 : M), cudaMemcpyHostToDevice, st));
    srk::launch_prep_index(ss, O, N, M, e.obs_cam.as<int32_t>(), e.obs_pt.as<int32_t>(), e.pt_begin.as<int64_t>(), e.cam_cnt.as<unsigned long long>(), e.flags.as<int>());
    srk::launch_scan_counts(ss, M, e.cam_cnt.as<unsigned long long>(), e.cam_begin.as<int64_t>(), e.cam_cursor.as<unsigned long long>());
    srk::launch_scatter_index(ss, O, N, M, e.obs_cam.as<int32_t>(), e.obs_pt.as<int32_t>(), e.cam_cursor.as<unsigned long long>(), e.c_pt.as<int32_t>(),
                              e.obs_pos.as<unsigned>());
    srk::launch_chunk_tables(ss, O, e.obs_cam.as<int32_t>(), e.obs_pt.as<int32_t>(), e.chunk_cams.as<int>(), e.chunk_cnt.as<int>(), e.chunk_pts.as<int>(),
                             e.obs_slot.as<unsigned char>());
    e.launches += 4;
    int h_flag = 0;
    SRK_CUDA(cudaMemcpyAsync(&h_flag, e.flags.p, sizeof(int), cudaMemcpyDeviceToHost, ss));
    SRK_CUDA(cudaStreamSynchronize(ss));
    if (h_flag != 0) {
        cudaStreamSynchronize(st);     // the caller's buffers are still being read by the copies in flight
        set_error(h_flag & 1 ? "observation index out of range" : "observations must be sorted by (pnt_ind, frame_ind) with at most one per pair");
        return SRK_E_INVALID_ARG;
    }

    // ---- plan of the tiled Schur kernel: which points fall back to the per-point kernel (long tracks, scattered cameras).
    // Larger tiles amortise the per-tile table build and the flush of the accumulators, but a tile may only touch 12 cameras:
    // take the largest tile size that defers (almost) no more points than the smallest one.  Structure only, once per bind.
    SRK_CUDA(e.deferred.ensure((size_t)(N > 0 ? N : 1)));
    e.n_deferred = 0;
    if (N > 0) {
        auto plan = [&](int tile, int64_t* out) -> int {
            SRK_CUDA(cudaMemsetAsync(e.skipped_cnt.p, 0, sizeof(unsigned long long), ss));
            srk::launch_schur_plan(ss, N, tile, e.pt_begin.as<int64_t>(), e.obs_cam.as<int32_t>(), e.deferred.as<unsigned char>(),
                                   e.skipped_cnt.as<unsigned long long>());
            e.launches += 1;
            unsigned long long nd = 0;
            SRK_CUDA(cudaMemcpyAsync(&nd, e.skipped_cnt.p, sizeof(nd), cudaMemcpyDeviceToHost, ss));
            SRK_CUDA(cudaStreamSynchronize(ss));
            *out = (int64_t)nd;
            return SRK_OK;
        };
        int64_t nd_small = 0;
        int rcp = plan(256, &nd_small);
        if (rcp != SRK_OK) return rcp;
        e.schur_tile_points = 256; e.n_deferred = nd_small;
        if (e.schur_tile_fixed > 0) {
            if (e.schur_tile_fixed != 256) { rcp = plan(e.schur_tile_fixed, &e.n_deferred); if (rcp != SRK_OK) return rcp; e.schur_tile_points = e.schur_tile_fixed; }
        } else if (N >= 148 * 4 * 1024) {   // enough tiles to keep every SM busy for several waves
            // measured with the third form of K2 at configs[2] (K2 family, ms): 128 points 1.89, 256 1.78, 384 1.77, 512 1.80, 640 1.82, 768 1.82,
            // 1024 1.96, 1536 2.28 -- smaller tiles touch fewer cameras (fewer fragment rows: 105 -> 91 DMMAs per k-step at 10 cameras) and
            // the per-tile table work is gone (bind-time tables), larger ones amortise the flush of the accumulators
            int cands[3] = {384, 512, 768};
            bool chosen = false;
            for (int t : cands) {
                int64_t nd = 0;
                rcp = plan(t, &nd);
                if (rcp != SRK_OK) return rcp;
                if (nd <= nd_small + N / 1000) { e.schur_tile_points = t; e.n_deferred = nd; chosen = true; break; }
            }
            if (!chosen) { rcp = plan(256, &e.n_deferred); if (rcp != SRK_OK) return rcp; }   // restore the flags of the small tiling
        }
    }
    // tables of the tile kernel for the chosen tiling (table slot of every observation, slot mask of every point, camera table of every tile)
    if (N > 0 && e.schur_impl == 0) {
        const int64_t ntile = (N + e.schur_tile_points - 1) / e.schur_tile_points;
        SRK_CUDA(e.k2_slot.ensure((size_t)(O > 0 ? O : 1))); SRK_CUDA(e.k2_mask.ensure(sizeof(unsigned short) * (size_t)N));
        SRK_CUDA(e.k2_tab.ensure(sizeof(int) * 12 * (size_t)ntile)); SRK_CUDA(e.k2_ntab.ensure(sizeof(int) * (size_t)ntile));
        srk::launch_schur_tables(ss, N, e.schur_tile_points, e.pt_begin.as<int64_t>(), e.obs_cam.as<int32_t>(), e.k2_tab.as<int>(), e.k2_ntab.as<int>(),
                                 e.k2_slot.as<unsigned char>(), e.k2_mask.as<unsigned short>(), e.deferred.as<unsigned char>());
        e.launches += 1;
    }
    // dense-rows K2 when most points are left to the per-point kernel and the rows fit comfortably (2 x 3N x 10M doubles)
    e.schur_dense_rows = e.schur_rows_enabled && N > 0 && e.n_deferred * 2 > N && e.nf <= 16384 &&
                         (double)N * 3.0 * (double)M * 10.0 * 16.0 <= 16.0e9 && 3 * N < (int64_t)2000000000;
    // ---- elimination order of the dense solve (camera co-visibility graph -> nested dissection, solve_order.cu): structure only, so it
    // also runs under the copies.  Multi-GPU keeps it at the first solve (the graph is a union over ranks: an all-reduce on the main stream).
    if (normalize && opt != nullptr && e.world <= 1 && pick_solver(e, opt) == SRK_SOLVER_DENSE_CHOLESKY) {
        int rco = ensure_solve_order(e, ss);
        if (rco != SRK_OK) { cudaStreamSynchronize(st); return rco; }
    }
    // ---- gauge normalisation (BA.cpp:203-247): cameras on the host (M records), points on the device
    std::vector<Pose> cams((size_t)M);
    std::memcpy(cams.data(), p->cams, sizeof(double) * 12 * (size_t)M);
    e.norm_failed = false; e.normalized = false; e.world_scale = 1.0;
    srk::launch_points_to_planes(st, N, e.pts_stage.as<double>(), e.X_cur); e.launches += 1;
    if (normalize) {
        Pose cam0_from1 = pose_compose(cams[0], pose_inv(cams[1]));  // SE3AFromB(cam0, cam1)
        double shift = cam0_from1.T[unity];
        if (is_close(0.0, shift, 1e-5)) {  // quirk Q4: the 1e-5 "atol" lands in the rtol slot
            e.norm_failed = true;
        } else {
            e.world_scale = unity_val / std::fabs(shift);
            e.cam0_prenorm = cams[0];
            M3 R0t = transpose(e.cam0_prenorm.R);
            for (int i = 0; i < M; ++i) {  // BA.cpp:143-162
                Pose n;
                n.R = mul(cams[i].R, R0t);
                double t[3]; mulv(mul(cams[i].R, R0t), e.cam0_prenorm.T, t);
                for (int k = 0; k < 3; ++k) n.T[k] = (cams[i].T[k] - t[k]) * e.world_scale;
                cams[i] = n;
            }
            // the 12 doubles of cam0 go through the (not yet used) trial pose buffer
            SRK_CUDA(cudaMemcpyAsync(e.cams_try, &e.cam0_prenorm, sizeof(double) * 12, cudaMemcpyHostToDevice, st));
            srk::launch_normalize_points(st, N, e.X_cur, e.cams_try, e.world_scale, 0); e.launches += 1;
            SRK_CUDA(cudaStreamSynchronize(st));
            e.normalized = true;
        }
    }
    SRK_CUDA(cudaMemcpyAsync(e.cams_cur, cams.data(), sizeof(double) * 12 * M, cudaMemcpyHostToDevice, st));
    SRK_CUDA(cudaMemcpyAsync(e.cams_bound.p, e.cams_cur, sizeof(double) * 12 * M, cudaMemcpyDeviceToDevice, st));
    if (N > 0) SRK_CUDA(cudaMemcpyAsync(e.Xbound.p, e.X_cur, sizeof(double) * 3 * N, cudaMemcpyDeviceToDevice, st));
    srk::launch_cam_prep(st, M, e.cams_cur, e.Kd.as<double>(), e.shared_K, e.f0, e.camd_cur); e.launches += 1;
    // the pixels have arrived by now (or arrive while the main stream waits): value half of the observation prep
    srk::launch_prep_xy(st, O, e.obs_xy.as<double>(), e.f0, e.obs_pos.as<unsigned>(), e.ox.as<double>(), e.oy.as<double>(), e.c_x.as<double>(), e.c_y.as<double>());
    e.launches += O > 0 ? 1 : 0;
    SRK_CUDA(cudaStreamSynchronize(ss));
    SRK_CUDA(cudaStreamSynchronize(st));
    SRK_CUDA(cudaGetLastError());
    e.bound = true;
    e.S_clean = false;
    e.tacc_zero = false;
    return e.norm_failed ? 1 : SRK_OK;
}

int ensure_solver_buffers(Engine& e, int solver) {
    const int64_t N = e.N, O = e.O; const int M = e.M;
    SRK_CUDA(e.J.ensure(sizeof(double) * 28 * (size_t)(O > 0 ? O : 1)));
    SRK_CUDA(e.Ggf.ensure(sizeof(double) * 110 * (size_t)M));
    SRK_CUDA(e.pinv.ensure(sizeof(double) * 9 * (size_t)(N > 0 ? N : 1)));
    SRK_CUDA(e.skipped.ensure((size_t)(N > 0 ? N : 1)));
    SRK_CUDA(e.dfull.ensure(sizeof(double) * 10 * (size_t)M));
    SRK_CUDA(e.xsol.ensure(sizeof(double) * (size_t)e.ld)); SRK_CUDA(e.resid.ensure(sizeof(double) * (size_t)e.ld));
    if (solver == SRK_SOLVER_DENSE_CHOLESKY) {
        size_t sbytes = sizeof(double) * ((size_t)e.ld * (size_t)e.nf + (size_t)e.ld);
        SRK_CUDA(e.Srhs.ensure(sbytes));
        SRK_CUDA(e.Lfac.ensure(sizeof(double) * (size_t)e.ld * (size_t)e.nf));
        SRK_CUDA(e.dinv.ensure(sizeof(double) * srk::dense_cholesky_dinv_doubles(e.nf)));
    }
    return SRK_OK;
}

int pick_solver(const Engine& e, const srk_ba_options* opt) {
    int s = opt != nullptr ? opt->solver : SRK_SOLVER_AUTO;
    if (s == SRK_SOLVER_DENSE_CHOLESKY || s == SRK_SOLVER_BLOCK_PCG) return s;
    // AUTO: a dense FP64 factorisation while the reduced system is a real dense contraction that fits comfortably
    // (n_f <= 16384: 2 GiB for S, 1.5e12 flop); beyond that the block-sparse PCG.
    return e.nf <= 16384 ? SRK_SOLVER_DENSE_CHOLESKY : SRK_SOLVER_BLOCK_PCG;
}

// ReprojError of a state (BA.cpp:410-490) -> e.errsum[0] on the device.
void residual_of(Engine& e, const double* X, const double* camd) {
    Scope s(e, F_RESIDUAL);
    srk::launch_residual(e.stream, e.O, e.obs_cam.as<int32_t>(), e.obs_pt.as<int32_t>(), e.ox.as<double>(), e.oy.as<double>(), X, e.N, camd,
                         e.chunk_cams.as<int>(), e.chunk_cnt.as<int>(), e.chunk_pts.as<int>(), e.obs_slot.as<unsigned char>(), e.partial.as<double>(), e.residual_blocks,
                         e.errsum.as<double>());
    e.launches += 2;
}

// Brings (global error, non-finite flag, skipped count, Cholesky info) of the current attempt to the host.
int fetch_attempt_scalars(Engine& e, bool with_flags, double* err, bool* nonfinite, int64_t* skipped, bool* factor_failed = nullptr) {
    srk::launch_pack_attempt(e.stream, e.errsum.as<double>(), with_flags ? e.flags.as<int>() + 1 : nullptr,
                             with_flags ? e.skipped_cnt.as<unsigned long long>() : nullptr, e.rank, e.world, e.slots.as<double>());
    e.launches += 1;
    int rc = do_allreduce(e, e.slots.as<double>(), e.world + 3);
    if (rc != SRK_OK) return rc;
    SRK_CUDA(cudaMemcpyAsync(e.h_slots, e.slots.p, sizeof(double) * (e.world + 3), cudaMemcpyDeviceToHost, e.stream));
    SRK_CUDA(cudaStreamSynchronize(e.stream));
    double s = 0.0;
    for (int r = 0; r < e.world; ++r) s += e.h_slots[r];  // rank order: identical on every rank
    *err = s;
    if (nonfinite != nullptr) *nonfinite = e.h_slots[e.world] != 0.0;
    if (skipped != nullptr) *skipped = (int64_t)e.h_slots[e.world + 1];
    if (factor_failed != nullptr) *factor_failed = e.h_slots[e.world + 2] != 0.0;   // the solve is replicated: every rank reports the same pivot
    return SRK_OK;
}

// ComputeCloseFormReprErrorDerivatives (BA.cpp:1140-1448) on the current state.
int derivative_pass(Engine& e, bool keep_partial_G = false) {
    cudaStream_t st = e.stream;
    e.G_partial = false;
    {
        Scope s(e, F_JACOBIAN);
        // K2's third form consumes per-point sums of Jp^T Jp / Jp^T rho: K1 has every term in registers and leaves them in k2_eacc
        double* eacc = nullptr;
        e.k2_eacc_valid = false;
        if (e.schur_impl == 0 && e.k2_fuse_k1 && !e.schur_dense_rows && e.N > 0 && e.k2_eacc.ensure(sizeof(double) * 9 * (size_t)e.N) == cudaSuccess) {
            eacc = e.k2_eacc.as<double>();
            SRK_CUDA(cudaMemsetAsync(eacc, 0, sizeof(double) * 9 * (size_t)e.N, st));
            e.k2_eacc_valid = true;
        }
        srk::launch_jacobian(st, e.O, e.obs_cam.as<int32_t>(), e.obs_pt.as<int32_t>(), e.ox.as<double>(), e.oy.as<double>(), e.X_cur, e.N, e.camd_cur,
                             e.J.as<double>(), eacc);
        e.launches += e.O > 0 ? 1 : 0;
    }
    {
        Scope s(e, F_FRAME);
        // split a camera's observation list over several CTAs when there are few cameras (grid >= 2 waves of 148 SMs)
        // (two slices add commutatively, so up to 2 the blocks stay bit-reproducible; more only for few cameras)
        int splits = 1;
        { const int want = e.M < 296 ? 296 : 1184; splits = (want + e.M - 1) / e.M; int64_t per = e.M > 0 ? e.O / e.M : 0;
          const int64_t min_per = e.M < 296 ? 256 : 2048; while (splits > 1 && per / splits < min_per) --splits; }
        srk::launch_frame_blocks(st, e.M, e.cam_begin.as<int64_t>(), e.c_pt.as<int32_t>(), e.c_x.as<double>(), e.c_y.as<double>(), e.X_cur, e.N, e.camd_cur,
                                 e.Ggf.as<double>(), e.Ggf.as<double>() + 100 * (size_t)e.M, splits);
        e.launches += 1;
    }
    // Dense path of a multi-GPU run: G and g_f enter the reduced system linearly (damping multiplies the diagonal of the SUM), so every rank
    // adds its own partial blocks to its partial S and the one all-reduce of S carries them: one latency-bound collective less per iteration.
    if (keep_partial_G && e.world > 1) { e.G_partial = true; return SRK_OK; }
    return do_allreduce(e, e.Ggf.as<double>(), 110 * (int64_t)e.M);
}

// Order of the reduced camera system for the sparse-factor Cholesky: camera co-visibility graph of the bound observations (union over
// ranks, so that every rank factors in the same order and takes bit-identical decisions) -> host nested dissection (solve_order.cu).
int ensure_solve_order(Engine& e, cudaStream_t on_stream) {
    if (e.order_ready) return SRK_OK;
    e.order_ready = true;
    const int M = e.M;
    if (!e.solve_order_enabled || e.nf < 24 * 64 || M < 8) { e.order = srk::SolveOrder{}; return SRK_OK; }
    cudaStream_t st = on_stream != nullptr ? on_stream : e.stream;
    const size_t cells = (size_t)M * (size_t)M;
    SRK_CUDA(e.adjbuf.ensure(cells));
    SRK_CUDA(cudaMemsetAsync(e.adjbuf.p, 0, cells, st));
    srk::launch_cam_adjacency(st, e.N, e.pt_begin.as<int64_t>(), e.obs_cam.as<int32_t>(), M, e.adjbuf.as<unsigned char>()); e.launches += e.N > 0 ? 1 : 0;
    if (e.ar != nullptr && e.world > 1) {
        SRK_CUDA(e.tmask.ensure(sizeof(double) * cells));
        srk::launch_bytes_to_doubles(st, (int64_t)cells, e.adjbuf.as<unsigned char>(), e.tmask.as<double>()); e.launches += 1;
        int rc = do_allreduce(e, e.tmask.as<double>(), (int64_t)cells);
        if (rc != SRK_OK) return rc;
        srk::launch_doubles_to_bytes(st, (int64_t)cells, e.tmask.as<double>(), e.adjbuf.as<unsigned char>()); e.launches += 1;
    }
    if (e.h_adj_cap < cells) {
        if (e.h_adj != nullptr) cudaFreeHost(e.h_adj);
        e.h_adj = nullptr; e.h_adj_cap = 0;
        SRK_CUDA(cudaMallocHost((void**)&e.h_adj, cells));
        e.h_adj_cap = cells;
    }
    SRK_CUDA(cudaMemcpyAsync(e.h_adj, e.adjbuf.p, cells, cudaMemcpyDeviceToHost, st));
    SRK_CUDA(cudaStreamSynchronize(st));
    // the co-visibility graph of a re-bound scene is usually the one already ordered: keep the order (and its device copy) then
    if (e.order_adj.size() == cells && e.order.n == e.nf && std::memcmp(e.order_adj.data(), e.h_adj, cells) == 0) return SRK_OK;
    e.order_adj.assign(e.h_adj, e.h_adj + cells);
    std::vector<int> gsize(M, 10);
    gsize[0] = 4; gsize[1] = 9;                          // quirk Q13: frame 0 keeps its intrinsics, frame 1 loses T[unity]
    e.order = srk::build_solve_order(M, gsize.data(), e.order_adj.data());
    e.order_tiles = false;
    if (!e.order.active) return SRK_OK;
    if (e.order.n != e.nf) { e.order = srk::SolveOrder{}; e.order_adj.clear(); return SRK_OK; }
    const int np = e.order.np; const size_t ldp = ((size_t)np + 7) & ~(size_t)7;
    if (e.solve_tiles_enabled && srk::dense_cholesky_pattern_ok(np, e.order.l_pattern_count, true)) {
        const srk::SolveOrder& o = e.order;
        std::vector<int> ints;
        auto put = [&](const std::vector<int>& v) { const int off = (int)ints.size(); ints.insert(ints.end(), v.begin(), v.end()); return off; };
        e.ot_s = put(o.s_tiles); e.ot_resptr = put(o.res_ptr); e.ot_resent = put(o.res_ent); e.ot_lin = put(o.l_in_tiles); e.ot_lall = put(o.l_all_tiles);
        { std::vector<int> cn(o.s_tiles.size()); const int nb0 = (e.nf + 63) / 64;       // the same tiles in the exchange kernels' c * nblk + r form
          for (size_t i = 0; i < cn.size(); ++i) cn[i] = (o.s_tiles[i] & 0xffff) * nb0 + (o.s_tiles[i] >> 16);
          e.ot_scn = put(cn); }
        SRK_CUDA(e.order_ints.ensure(sizeof(int) * ints.size()));
        SRK_CUDA(cudaMemcpyAsync(e.order_ints.p, ints.data(), sizeof(int) * ints.size(), cudaMemcpyHostToDevice, st));
        SRK_CUDA(e.order_pattern.ensure(o.l_pattern.size()));
        SRK_CUDA(cudaMemcpyAsync(e.order_pattern.p, o.l_pattern.data(), o.l_pattern.size(), cudaMemcpyHostToDevice, st));
        SRK_CUDA(cudaStreamSynchronize(st));
        e.order_tiles = true;
    }
    SRK_CUDA(e.order_src.ensure(sizeof(int) * (size_t)np));
    SRK_CUDA(cudaMemcpyAsync(e.order_src.p, e.order.src.data(), sizeof(int) * (size_t)np, cudaMemcpyHostToDevice, st));
    SRK_CUDA(cudaStreamSynchronize(st));
    SRK_CUDA(e.xperm.ensure(sizeof(double) * ldp));
    SRK_CUDA(e.Lfac.ensure(sizeof(double) * ldp * (size_t)np));
    SRK_CUDA(e.dinv.ensure(sizeof(double) * srk::dense_cholesky_dinv_doubles(np)));
    return SRK_OK;
}

// Sum of the dense reduced camera system over ranks.  The union of the ranks' non-zero 64x64 tiles is found first (a small mask
// all-reduce); when it is sparse only those tiles (+ rhs) travel: 16 MB instead of 0.8 GB at configs[2].
int allreduce_system(Engine& e, double* S, double* rhs) {
    if (e.ar == nullptr || e.world <= 1) return SRK_OK;
    cudaStream_t st = e.stream;
    const int nf = e.nf; const int64_t ld = e.ld;
    const int nblk = (nf + 63) / 64;
    if (nblk < 8 || e.tile_exchange == 0) return do_allreduce(e, S, ld * nf + ld);
    if (e.order_ready && e.order.active && e.order_tiles) {   // the tile structure is known (union over ranks): no mask pass, no host round trip
        const int cnt = (int)e.order.s_tiles.size();
        const int* lst = e.order_ints.as<int>() + e.ot_scn;
        const int64_t npk = (int64_t)cnt * 4096 + ld;
        SRK_CUDA(e.tpacked.ensure(sizeof(double) * (size_t)npk));
        srk::launch_tile_pack(st, nf, S, ld, lst, cnt, rhs, ld, e.tpacked.as<double>(), 0); e.launches += 1;
        int rc2 = do_allreduce(e, e.tpacked.as<double>(), npk);
        if (rc2 != SRK_OK) return rc2;
        srk::launch_tile_pack(st, nf, S, ld, lst, cnt, rhs, ld, e.tpacked.as<double>(), 1); e.launches += 1;
        return SRK_OK;
    }
    SRK_CUDA(e.tmask.ensure(sizeof(double) * (size_t)nblk * nblk));
    SRK_CUDA(e.tlist.ensure(sizeof(int) * ((size_t)nblk * nblk + 4)));
    double* mask = e.tmask.as<double>(); int* list = e.tlist.as<int>() + 4; int* cnt = e.tlist.as<int>();
    SRK_CUDA(cudaMemsetAsync(mask, 0, sizeof(double) * (size_t)nblk * nblk, st));
    srk::launch_tile_mask(st, nf, S, ld, mask); e.launches += 1;
    int rc = do_allreduce(e, mask, (int64_t)nblk * nblk);
    if (rc != SRK_OK) return rc;
    srk::launch_tile_list(st, nf, mask, list, cnt); e.launches += 1;
    int h_cnt = 0;
    SRK_CUDA(cudaMemcpyAsync(&h_cnt, cnt, sizeof(int), cudaMemcpyDeviceToHost, st));
    SRK_CUDA(cudaStreamSynchronize(st));
    if (h_cnt <= 0 || (int64_t)h_cnt * 4096 > (ld * nf) / 4) return do_allreduce(e, S, ld * nf + ld);   // dense: send everything
    const int64_t npk = (int64_t)h_cnt * 4096 + ld;
    SRK_CUDA(e.tpacked.ensure(sizeof(double) * (size_t)npk));
    srk::launch_tile_pack(st, nf, S, ld, list, h_cnt, rhs, ld, e.tpacked.as<double>(), 0); e.launches += 1;
    rc = do_allreduce(e, e.tpacked.as<double>(), npk);
    if (rc != SRK_OK) return rc;
    srk::launch_tile_pack(st, nf, S, ld, list, h_cnt, rhs, ld, e.tpacked.as<double>(), 1); e.launches += 1;
    return SRK_OK;
}

// K2: per-point blocks + Schur accumulation into `sink` (dense S or block-sparse blocks).
void schur_accumulate(Engine& e, const srk::SchurSink& sink, double c) {
    cudaStream_t st = e.stream;
    // Most points fall outside the tile kernel (long tracks: every point in every frame) and the system is dense: one DMMA contraction
    // over all points instead of per-point pair enumeration with atomics.
    if (e.schur_dense_rows && sink.blocks == nullptr && e.N > 0) {
        const int nfull = e.M * 10;
        const size_t rows = 3 * (size_t)e.N;
        if (e.rows_F.ensure(sizeof(double) * rows * nfull) == cudaSuccess && e.rows_W.ensure(sizeof(double) * rows * nfull) == cudaSuccess &&
            e.rows_D.ensure(sizeof(double) * (size_t)nfull * nfull) == cudaSuccess) {
            cudaMemsetAsync(e.rows_F.p, 0, sizeof(double) * rows * nfull, st);
            cudaMemsetAsync(e.rows_W.p, 0, sizeof(double) * rows * nfull, st);
            cudaMemsetAsync(e.rows_D.p, 0, sizeof(double) * (size_t)nfull * nfull, st);
            srk::launch_schur_rows(st, e.N, e.O, e.pt_begin.as<int64_t>(), e.obs_cam.as<int32_t>(), e.J.as<double>(), c, sink, e.pinv.as<double>(),
                                   e.skipped.as<unsigned char>(), e.M, e.rows_F.as<double>(), e.rows_W.as<double>());
            // rows_D starts at zero and the product is SUBTRACTED from it: D = -(Fall^T Wall); the scatter then adds it with the same sign convention
            srk::launch_gemm_nt_dmma(st, nfull, nfull, (int)rows, e.rows_F.as<double>(), nfull, e.rows_W.as<double>(), nfull, e.rows_D.as<double>(), nfull, 1, 1);
            srk::launch_scatter_dense_schur(st, nfull, e.rows_D.as<double>(), sink.unity, sink.S, sink.ld);
            e.launches += 3;
            return;
        }
        cudaGetLastError();   // not enough memory for the rows: the per-point kernels below
    }
    bool v3 = false;
    if (e.schur_impl == 0 && e.N > 0 && e.k2_gi.ensure(sizeof(double) * 6 * (size_t)e.N) == cudaSuccess && e.k2_u.ensure(sizeof(double) * 3 * (size_t)e.N) == cudaSuccess &&
        e.k2_exc.ensure(sizeof(int) * (Engine::kExcCap + 4)) == cudaSuccess) {
        v3 = true;
        int* exc_count = e.k2_exc.as<int>(); int* exc_list = exc_count + 4;
        cudaMemsetAsync(exc_count, 0, sizeof(int), st);
        if (e.k2_eacc_valid)
            srk::launch_point_finish(st, e.N, e.k2_eacc.as<double>(), c, e.pinv.as<double>(), e.skipped.as<unsigned char>(), e.k2_gi.as<double>(), e.k2_u.as<double>(),
                                     exc_list, exc_count, Engine::kExcCap);
        else
            srk::launch_point_factor(st, e.N, e.O, e.pt_begin.as<int64_t>(), e.J.as<double>(), c, e.pinv.as<double>(), e.skipped.as<unsigned char>(), e.k2_gi.as<double>(),
                                     e.k2_u.as<double>(), exc_list, exc_count, Engine::kExcCap);
        srk::launch_schur_v3(st, e.N, e.O, e.schur_tile_points, e.pt_begin.as<int64_t>(), e.obs_pt.as<int32_t>(), e.J.as<double>(), sink, e.k2_gi.as<double>(),
                             e.k2_u.as<double>(), e.skipped.as<unsigned char>(), e.k2_tab.as<int>(), e.k2_ntab.as<int>(), e.k2_slot.as<unsigned char>(),
                             e.k2_mask.as<unsigned short>());
        srk::launch_schur_list(st, e.N, e.O, e.pt_begin.as<int64_t>(), e.obs_cam.as<int32_t>(), e.J.as<double>(), c, sink, e.pinv.as<double>(),
                               e.skipped.as<unsigned char>(), exc_list, exc_count, 1024);
        e.launches += 2;
    } else if (e.schur_impl == 0 || e.schur_impl == 2)
        srk::launch_schur_mma(st, e.N, e.O, e.schur_tile_points, e.pt_begin.as<int64_t>(), e.obs_cam.as<int32_t>(), e.J.as<double>(), c, sink,
                              e.pinv.as<double>(), e.skipped.as<unsigned char>(), e.deferred.as<unsigned char>());
    else   // SRK_SCHUR_IMPL=1: the vector-FMA tile kernel (kept as a cross-check of the DMMA path)
        srk::launch_schur_tile(st, e.N, e.O, e.schur_tile_points, e.pt_begin.as<int64_t>(), e.obs_cam.as<int32_t>(), e.J.as<double>(), c, sink,
                               e.pinv.as<double>(), e.skipped.as<unsigned char>(), e.deferred.as<unsigned char>(), 0, nullptr, 0, nullptr);
    e.launches += e.N > 0 ? 1 : 0;
    if (e.n_deferred > 0) {
        srk::launch_schur(st, e.N, e.O, e.pt_begin.as<int64_t>(), e.obs_cam.as<int32_t>(), e.J.as<double>(), c, sink, e.pinv.as<double>(),
                          e.skipped.as<unsigned char>(), e.deferred.as<unsigned char>());
        e.launches += 1;
    }
}

// EstimateCorrectionsDecomposedInTwoPhases (BA.cpp:1771-1995) + ApplyCorrections (BA.cpp:1997-2063) into the trial state,
// then ReprojError of the trial state.  dp_out (optional) receives the point corrections [3N].
int attempt(Engine& e, int solver, const srk_ba_options* opt, double c, double* dp_out) {
    cudaStream_t st = e.stream;
    const int M = e.M; const int nf = e.nf; const int64_t ld = e.ld;
    double* G = e.Ggf.as<double>(); double* gf = G + 100 * (size_t)M;
    SRK_CUDA(cudaMemsetAsync(e.flags.as<int>() + 1, 0, sizeof(int) * 2, st));
    SRK_CUDA(cudaMemsetAsync(e.skipped_cnt.p, 0, sizeof(unsigned long long), st));
    double* x = e.xsol.as<double>();
    if (solver == SRK_SOLVER_DENSE_CHOLESKY) {
        double* S = e.Srhs.as<double>(); double* rhs = S + (size_t)ld * nf;
        {
            Scope s(e, F_SCHUR);
            // the order (and with it the tile structure of S) is needed before S is cleared
            { int rco = ensure_solve_order(e); if (rco != SRK_OK) return rco; }
            if (e.order.active && e.order_tiles && e.S_clean) {
                srk::launch_zero_tiles(st, nf, S, ld, e.order_ints.as<int>() + e.ot_s, (int)e.order.s_tiles.size()); e.launches += 1;
                SRK_CUDA(cudaMemsetAsync(rhs, 0, sizeof(double) * (size_t)ld, st));
            } else {
                SRK_CUDA(cudaMemsetAsync(S, 0, sizeof(double) * ((size_t)ld * nf + ld), st));
                e.S_clean = e.order.active && e.order_tiles;
            }
            if (e.rank == 0 || e.G_partial) { srk::launch_fill_reduced(st, M, G, gf, c, e.unity, S, ld, rhs); e.launches += 1; }
            srk::SchurSink sink{S, ld, rhs, e.unity, nullptr, nullptr, 0, nullptr};
            schur_accumulate(e, sink, c);
        }
        int rc = allreduce_system(e, S, rhs);
        if (rc != SRK_OK) return rc;
        {
            Scope s(e, F_SOLVE);
            int refine = opt != nullptr ? opt->refine_steps : 1;
            double* L = e.Lfac.as<double>();
            double* di = e.dinv.as<double>();
            double* r = e.resid.as<double>();
            const bool tiles = e.order.active && e.order_tiles;
            if (refine > 0 && !tiles) { srk::launch_mirror_lower(st, nf, S, ld); e.launches += 1; e.S_clean = false; }
            if (e.order.active) {
                // factor P S P^T (parts concurrently, separator last); vectors travel natural -> ordered -> natural
                const int np = e.order.np; const int64_t ldp = ((int64_t)np + 7) & ~(int64_t)7;
                const int* src = e.order_src.as<int>(); double* xp = e.xperm.as<double>();
                const srk::CholPartition* part = &e.order.part;
                const int* oi = e.order_ints.as<int>();
                const bool sc = tiles && srk::dense_cholesky_trsv_is_sparse(np, (int64_t)e.order.l_all_tiles.size());
                if (tiles) {   // only the tiles that can be non-zero: the factor's (symbolic) tiles are cleared, the system's are copied
                    srk::launch_zero_tiles(st, np, L, ldp, oi + e.ot_lall, (int)e.order.l_all_tiles.size());
                    srk::launch_permute_tiles(st, nf, S, ld, np, src, L, ldp, oi + e.ot_lin, (int)e.order.l_in_tiles.size()); e.launches += 2;
                } else {
                    srk::launch_permute_sym(st, nf, S, ld, refine > 0 ? 1 : 0, np, src, L, ldp); e.launches += 1;
                }
                srk::launch_gather_vec(st, np, src, rhs, xp); e.launches += 1;
                { Scope s2(e, F_FACTOR);
                  e.launches += srk::dense_cholesky_factor(st, np, L, ldp, di, e.flags.as<int>() + 2, part, tiles ? e.order_pattern.as<unsigned char>() : nullptr,
                                                           tiles ? e.order.l_pattern_count : 0); }
                { Scope s2(e, F_TRSV);
                  e.launches += srk::dense_cholesky_forward(st, np, L, ldp, di, xp, part, sc);
                  e.launches += srk::dense_cholesky_backward(st, np, L, ldp, di, xp, part, sc); }
                srk::launch_scatter_vec(st, np, src, xp, x); e.launches += 1;
                for (int it = 0; it < refine; ++it) {
                    if (tiles) srk::launch_residual_dd_tiles(st, nf, S, ld, x, rhs, r, oi + e.ot_resptr, oi + e.ot_resent);
                    else srk::launch_residual_dd(st, nf, S, ld, x, rhs, r);
                    e.launches += 1;
                    srk::launch_gather_vec(st, np, src, r, xp); e.launches += 1;
                    { Scope s2(e, F_TRSV);
                      e.launches += srk::dense_cholesky_forward(st, np, L, ldp, di, xp, part, sc);
                      e.launches += srk::dense_cholesky_backward(st, np, L, ldp, di, xp, part, sc); }
                    srk::launch_scatter_vec(st, np, src, xp, r); e.launches += 1;
                    srk::launch_axpy1(st, nf, r, x); e.launches += 1;
                }
            } else {
                SRK_CUDA(cudaMemcpyAsync(L, S, sizeof(double) * (size_t)ld * nf, cudaMemcpyDeviceToDevice, st));
                SRK_CUDA(cudaMemcpyAsync(x, rhs, sizeof(double) * nf, cudaMemcpyDeviceToDevice, st));
                { Scope s2(e, F_FACTOR); e.launches += srk::dense_cholesky_factor(st, nf, L, ld, di, e.flags.as<int>() + 2); }
                { Scope s2(e, F_TRSV);
                  e.launches += srk::dense_cholesky_forward(st, nf, L, ld, di, x);
                  e.launches += srk::dense_cholesky_backward(st, nf, L, ld, di, x); }
                for (int it = 0; it < refine; ++it) {
                    srk::launch_residual_dd(st, nf, S, ld, x, rhs, r);
                    e.launches += 1;
                    { Scope s2(e, F_TRSV);
                      e.launches += srk::dense_cholesky_forward(st, nf, L, ld, di, r);
                      e.launches += srk::dense_cholesky_backward(st, nf, L, ld, di, r); }
                    srk::launch_axpy1(st, nf, r, x); e.launches += 1;
                }
            }
        }
        e.solver_used = SRK_SOLVER_DENSE_CHOLESKY;
    } else {
        int rc = SRK_OK;
        if (!e.pcg.structure_valid) {
            rc = srk::pcg_build_structure(e.pcg, st, e.N, e.O, M, e.schur_tile_points, e.pt_begin.as<int64_t>(), e.obs_cam.as<int32_t>(),
                                          e.deferred.as<unsigned char>(), &e.launches);
            if (rc != SRK_OK) { set_error("block-sparse structure of the reduced camera system could not be built (hash table overflow?)"); return rc; }
        }
        srk::SchurSink sink{};
        {
            Scope s(e, F_SCHUR);
            rc = srk::pcg_begin(e.pcg, st, M, G, gf, c, e.unity, e.rank, &sink, &e.launches);
            if (rc != SRK_OK) { set_error("pcg_begin failed"); return rc; }
            schur_accumulate(e, sink, c);
        }
        {
            Scope s(e, F_SOLVE);
            double rel = 0.0;
            rc = srk::pcg_solve(e.pcg, st, M, e.dfull.as<double>(), opt != nullptr ? opt->pcg_max_iters : 0, opt != nullptr ? opt->pcg_rel_tol : 0.0, e.world,
                                e.ar, e.ar_user, &e.launches, &e.pcg_iters_last, &rel);
            if (rc != SRK_OK) { set_error("pcg solve failed: " + std::string(cudaGetErrorString(cudaGetLastError()))); return rc; }
            e.pcg_rel_res_last = rel;
            e.pcg_iters_total += e.pcg_iters_last;
            // The reference solves the reduced system exactly (BA.cpp:1911).  An iterative solve that stopped far from its tolerance (iteration
            // cap, breakdown) must not pass for one: like a failed factorisation it counts as a failed attempt and is retried with more
            // damping (flags[2] is the factorisation-info slot fetch_attempt_scalars reports).  Every rank sees the same residual.
            const double tol = (opt != nullptr && opt->pcg_rel_tol > 0.0) ? opt->pcg_rel_tol : 1e-13;
            if (!(rel <= 1e4 * tol) && !(rel <= 1e-8)) {
                const int one = 1;
                SRK_CUDA(cudaMemcpyAsync(e.flags.as<int>() + 2, &one, sizeof(int), cudaMemcpyHostToDevice, st));
                SRK_CUDA(cudaStreamSynchronize(st));
            }
        }
        e.solver_used = SRK_SOLVER_BLOCK_PCG;
    }
    // allFinite(corrections_frame) (BA.cpp:1912-1913); the PCG path solves directly in full frame-variable space
    if (solver == SRK_SOLVER_DENSE_CHOLESKY) {
        srk::launch_finite_flag(st, nf, x, e.flags.as<int>() + 1); e.launches += 1;
        srk::launch_expand_df(st, M, x, e.unity, e.dfull.as<double>()); e.launches += 1;
    } else {
        srk::launch_finite_flag(st, 10 * (int64_t)M, e.dfull.as<double>(), e.flags.as<int>() + 1); e.launches += 1;
    }
    {
        Scope s(e, F_BACKSUB);
        int64_t avg = e.N > 0 ? e.O / e.N : 0;
        if (e.backsub_impl == 1 && e.N > 0) {
            if (!e.tacc_zero) { SRK_CUDA(e.tacc.ensure(sizeof(double) * 3 * (size_t)e.N)); SRK_CUDA(cudaMemsetAsync(e.tacc.p, 0, sizeof(double) * 3 * (size_t)e.N, st)); e.tacc_zero = true; }
            srk::launch_backsub_obs(st, e.N, e.O, e.obs_pt.as<int32_t>(), e.obs_cam.as<int32_t>(), e.J.as<double>(), e.dfull.as<double>(), e.pinv.as<double>(),
                                    e.skipped.as<unsigned char>(), e.X_cur, e.X_try, dp_out, e.tacc.as<double>());
            e.launches += 2;
        } else {
            srk::launch_backsub(st, e.N, e.O, e.pt_begin.as<int64_t>(), e.obs_cam.as<int32_t>(), e.J.as<double>(), e.dfull.as<double>(), e.pinv.as<double>(),
                                e.skipped.as<unsigned char>(), e.X_cur, e.X_try, dp_out, (int)(avg < 4 ? 4 : avg));
            e.launches += e.N > 0 ? 1 : 0;
        }
    }
    srk::launch_finite_flag(st, 3 * e.N, e.X_try, e.flags.as<int>() + 1); e.launches += e.N > 0 ? 1 : 0;  // BA.cpp:1953-1954
    srk::launch_count_skipped(st, e.N, e.skipped.as<unsigned char>(), e.skipped_cnt.as<unsigned long long>()); e.launches += e.N > 0 ? 1 : 0;
    {
        Scope s(e, F_UPDATE);
        srk::launch_cam_update(st, M, e.cams_cur, e.dfull.as<double>(), e.cams_try);
        srk::launch_cam_prep(st, M, e.cams_try, e.Kd.as<double>(), e.shared_K, e.f0, e.camd_try);
        e.launches += 2;
    }
    residual_of(e, e.X_try, e.camd_try);
    return SRK_OK;
}

// allFinite over the dense reduced system (S and rhs) of the last attempt -- only on the rare failed-factorisation path.
int system_is_finite(Engine& e, bool* finite) {
    *finite = true;
    if (e.Srhs.p == nullptr) return SRK_OK;
    SRK_CUDA(cudaMemsetAsync(e.flags.as<int>() + 3, 0, sizeof(int), e.stream));
    srk::launch_finite_flag(e.stream, e.ld * (int64_t)e.nf + e.ld, e.Srhs.as<double>(), e.flags.as<int>() + 3); e.launches += 1;
    int h = 0;
    SRK_CUDA(cudaMemcpyAsync(&h, e.flags.as<int>() + 3, sizeof(int), cudaMemcpyDeviceToHost, e.stream));
    SRK_CUDA(cudaStreamSynchronize(e.stream));
    *finite = h == 0;
    return SRK_OK;
}

void accept_trial(Engine& e) {
    std::swap(e.X_cur, e.X_try); std::swap(e.cams_cur, e.cams_try); std::swap(e.camd_cur, e.camd_try);
}

// ComputeOnNormalizedWorld (BA.cpp:720-893).
int run_impl(Engine& e, const srk_ba_options* opt, srk_ba_report* rep) {
    if (!e.bound) { set_error("srk_ba_run before srk_ba_bind"); return SRK_E_NOT_BOUND; }
    srk_ba_options defopt; srk_ba_default_options(&defopt);
    if (opt == nullptr) opt = &defopt;
    const int64_t launches0 = e.launches, failures0 = e.factor_failures;
    if (const char* v = std::getenv("SRK_DEBUG_FAIL_FACTOR")) e.debug_fail_factor = std::atoi(v);
    auto finish = [&](int converged, int reason, int iters, int attempts, double e0, double e1, double c, int64_t seen) {
        if (rep == nullptr) return;
        rep->converged = converged; rep->stop_reason = reason; rep->outer_iters = iters; rep->attempts = attempts;
        rep->err_initial = e0; rep->err_final = e1; rep->hessian_factor_final = c; rep->seen_points = seen;
        rep->gpu_launches = e.launches - launches0; rep->solver_used = e.solver_used; rep->pcg_iters_last = e.pcg_iters_last;
        rep->pcg_rel_res_last = e.pcg_rel_res_last; rep->factor_failures = (int32_t)(e.factor_failures - failures0);
    };
    if (rep != nullptr) { rep->err_trace_len = 0; rep->attempt_trace_len = 0; }
    if (e.norm_failed) { finish(0, SRK_STOP_NORMALIZATION_FAILED, 0, 0, 0.0, 0.0, 0.0, 0); return SRK_OK; }
    SRK_CUDA(cudaSetDevice(e.device));
    const int solver = pick_solver(e, opt);
    e.solver_used = solver;
    int rc = ensure_solver_buffers(e, solver);
    if (rc != SRK_OK) return rc;

    // seen_points_count (BA.cpp:483): observations over all ranks
    int64_t seen = e.O;
    if (e.ar != nullptr && e.world > 1) {
        double v = (double)e.O;
        SRK_CUDA(cudaMemcpyAsync(e.errsum.p, &v, sizeof(double), cudaMemcpyHostToDevice, e.stream));
        double tot = 0.0;
        rc = fetch_attempt_scalars(e, false, &tot, nullptr, nullptr);
        if (rc != SRK_OK) return rc;
        seen = (int64_t)tot;
    }

    double hessian_factor = (double)0.0001f;  // quirk Q1: float literal widened (BA.cpp:723)
    double err_initial = 0.0;
    residual_of(e, e.X_cur, e.camd_cur);
    rc = fetch_attempt_scalars(e, false, &err_initial, nullptr, nullptr);
    if (rc != SRK_OK) return rc;
    if (opt->has_err_change && err_initial < opt->err_change) {  // BA.cpp:749-753
        finish(1, SRK_STOP_ABS_ERR_THRESHOLD, 0, 0, err_initial, err_initial, hessian_factor, seen);
        return SRK_OK;
    }
    double err_value = err_initial;
    int it = 1, attempts = 0;
    while (true) {
        if (opt->max_outer_iters != 0 && it > opt->max_outer_iters) {
            finish(0, SRK_STOP_MAX_ITERS, it - 1, attempts, err_initial, err_value, hessian_factor, seen);
            return SRK_OK;
        }
        rc = derivative_pass(e, solver == SRK_SOLVER_DENSE_CHOLESKY);  // once per outer iteration (quirk Q7)
        if (rc != SRK_OK) return rc;
        enum { Success, FailedHessianOverflow, FailedButConverged } result;
        double err_new = std::nan("");
        bool have_prev = false; double err_new_prev = 0.0;
        while (true) {  // try_decrease_targ_fun (BA.cpp:764-852); the trial buffers replace the reference's backup/restore
            rc = attempt(e, solver, opt, hessian_factor, nullptr);
            if (rc != SRK_OK) return rc;
            bool nonfinite = false, factor_failed = false; int64_t skipped = 0;
            rc = fetch_attempt_scalars(e, true, &err_new, &nonfinite, &skipped, &factor_failed);
            if (rc != SRK_OK) return rc;
            ++attempts;
            if (e.debug_fail_factor > 0 && solver == SRK_SOLVER_DENSE_CHOLESKY) { --e.debug_fail_factor; factor_failed = true; }
            if (factor_failed) {
                // The reduced system is not numerically positive definite at this damping.  The reference solves with Householder QR
                // (BA.cpp:1911), which returns finite corrections for any finite S: the step then fails the decrease test and is retried
                // with hessian_factor * 10 (:841).  Cholesky has no such answer, so a failed factorisation IS the failed attempt -- unless
                // S itself holds a non-finite entry, where the reference's allFinite test (:1912) ends the run.
                e.factor_failures += 1;
                bool finite = true;
                rc = system_is_finite(e, &finite);
                if (rc != SRK_OK) return rc;
                if (!finite) { result = FailedHessianOverflow; break; }
                if (rep != nullptr && rep->attempt_trace != nullptr && rep->attempt_trace_len < rep->attempt_trace_cap) {
                    double* a = rep->attempt_trace + 4 * (size_t)rep->attempt_trace_len++;
                    a[0] = hessian_factor; a[1] = HUGE_VAL; a[2] = 0.0; a[3] = (double)skipped;
                }
                hessian_factor *= 10;
                if (!std::isfinite(hessian_factor) || (opt->has_max_hessian_factor && hessian_factor > opt->max_hessian_factor)) { result = FailedHessianOverflow; break; }
                continue;
            }
            if (nonfinite) { result = FailedHessianOverflow; break; }
            bool decreased = (err_new - err_value) < 0;
            if (rep != nullptr && rep->attempt_trace != nullptr && rep->attempt_trace_len < rep->attempt_trace_cap) {
                double* a = rep->attempt_trace + 4 * (size_t)rep->attempt_trace_len++;
                a[0] = hessian_factor; a[1] = err_new; a[2] = decreased ? 1.0 : 0.0; a[3] = (double)skipped;
            }
            if (decreased) { accept_trial(e); result = Success; break; }
            if (have_prev && opt->has_err_change && std::fabs(err_new - err_new_prev) < opt->err_change) { result = FailedButConverged; break; }
            hessian_factor *= 10;
            if (opt->has_max_hessian_factor && hessian_factor > opt->max_hessian_factor) { result = FailedHessianOverflow; break; }
            err_new_prev = err_new; have_prev = true;
        }
        if (result != Success) {
            finish(0, result == FailedHessianOverflow ? SRK_STOP_HESSIAN_OVERFLOW : SRK_STOP_ERR_CONVERGED, it, attempts, err_initial, err_value,
                   hessian_factor, seen);
            return SRK_OK;
        }
        if (rep != nullptr && rep->err_trace != nullptr && rep->err_trace_len < rep->err_trace_cap) rep->err_trace[rep->err_trace_len++] = err_new;
        double change = err_new - err_value;
        if (opt->has_err_change && std::fabs(change) < opt->err_change) {  // BA.cpp:880-884
            finish(1, SRK_STOP_SMALL_ERR_CHANGE, it, attempts, err_initial, err_new, hessian_factor, seen);
            return SRK_OK;
        }
        err_value = err_new;
        hessian_factor /= 10;
        it += 1;
    }
}

int fetch_impl(Engine& e, srk_ba_problem* p) {
    if (!e.bound) { set_error("srk_ba_fetch before srk_ba_bind"); return SRK_E_NOT_BOUND; }
    if (p == nullptr || p->cams == nullptr || (e.N > 0 && p->points == nullptr)) { set_error("null output"); return SRK_E_INVALID_ARG; }
    SRK_CUDA(cudaSetDevice(e.device));
    cudaStream_t st = e.stream;
    const int64_t N = e.N; const int M = e.M;
    // points: the trial buffer is dead between runs and serves as scratch
    if (N > 0) {
        SRK_CUDA(cudaMemcpyAsync(e.X_try, e.X_cur, sizeof(double) * 3 * N, cudaMemcpyDeviceToDevice, st));
        if (e.normalized) {
            SRK_CUDA(cudaMemcpyAsync(e.cams_try, &e.cam0_prenorm, sizeof(double) * 12, cudaMemcpyHostToDevice, st));
            srk::launch_normalize_points(st, N, e.X_try, e.cams_try, e.world_scale, 1); e.launches += 1;
        }
        srk::launch_planes_to_points(st, N, e.X_try, e.pts_stage.as<double>()); e.launches += 1;
        SRK_CUDA(cudaMemcpyAsync(p->points, e.pts_stage.p, sizeof(double) * 3 * N, cudaMemcpyDeviceToHost, st));
    }
    std::vector<Pose> cams((size_t)M);
    SRK_CUDA(cudaMemcpyAsync(cams.data(), e.cams_cur, sizeof(double) * 12 * M, cudaMemcpyDeviceToHost, st));
    SRK_CUDA(cudaStreamSynchronize(st));
    if (e.normalized) {  // BA.cpp:164-177
        for (int i = 0; i < M; ++i) {
            Pose r;
            r.R = mul(cams[i].R, e.cam0_prenorm.R);
            double t[3]; mulv(cams[i].R, e.cam0_prenorm.T, t);
            for (int k = 0; k < 3; ++k) r.T[k] = cams[i].T[k] / e.world_scale + t[k];
            cams[i] = r;
        }
    }
    std::memcpy(p->cams, cams.data(), sizeof(double) * 12 * (size_t)M);
    return SRK_OK;
}

}  // namespace

// =====================================================================================================================
extern "C" {

int srk_abi_version(void) { return 2; }   // 2: srk_ba_report gained pcg_rel_res_last / factor_failures
void srk_internal_set_error(const char* s) { g_last_error = s != nullptr ? s : ""; }
const char* srk_last_error(void) { return g_last_error.c_str(); }

void srk_ba_default_options(srk_ba_options* o) {
    if (o == nullptr) return;
    std::memset(o, 0, sizeof(*o));
    o->unity_comp_ind = 1;       // BA.h:134
    o->unity_comp_value = 1.0;   // BA.h:133
    o->solver = SRK_SOLVER_AUTO;
    o->refine_steps = 1;
}

const char* srk_stop_reason_string(int32_t r) {
    switch (r) {  // BA.cpp:751, :866, :868, :882
    case SRK_STOP_ABS_ERR_THRESHOLD: return "abs err threshold";
    case SRK_STOP_SMALL_ERR_CHANGE: return "small relative err change";
    case SRK_STOP_HESSIAN_OVERFLOW: return "hessian overflow";
    case SRK_STOP_ERR_CONVERGED: return "err converged to limit value";
    case SRK_STOP_MAX_ITERS: return "max iterations";
    default: return "";
    }
}

int srk_ba_create(void** h, const int* device_ids, int n_devices) {
    if (h == nullptr) { set_error("null handle pointer"); return SRK_E_INVALID_ARG; }
    *h = nullptr;
    if (n_devices > 1) { set_error("one handle drives one device; use one process per GPU"); return SRK_E_INVALID_ARG; }
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) { cudaGetLastError(); set_error("no CUDA device (there is no CPU fallback)"); return SRK_E_NO_DEVICE; }
    int dev = (device_ids != nullptr && n_devices == 1) ? device_ids[0] : 0;
    if (dev < 0 || dev >= count) { set_error("device id out of range"); return SRK_E_NO_DEVICE; }
    cudaDeviceProp prop;
    SRK_CUDA(cudaGetDeviceProperties(&prop, dev));
    if (prop.major != 10) { set_error(std::string("device is sm_") + std::to_string(prop.major * 10 + prop.minor) + ", this library holds sm_100a code only"); return SRK_E_NO_DEVICE; }
    SRK_CUDA(cudaSetDevice(dev));
    Engine* e = new Engine();
    e->device = dev;
    if (const char* v = std::getenv("SRK_SCHUR_IMPL")) { const int q = std::atoi(v); e->schur_impl = (q == 1 || q == 2) ? q : 0; }
    if (const char* v = std::getenv("SRK_K2_FUSE_K1")) e->k2_fuse_k1 = std::atoi(v) != 0 ? 1 : 0;
    if (const char* v = std::getenv("SRK_TILE_EXCHANGE")) e->tile_exchange = std::atoi(v) != 0 ? 1 : 0;
    if (const char* v = std::getenv("SRK_SOLVE_ORDER")) e->solve_order_enabled = std::atoi(v) != 0 ? 1 : 0;
    if (const char* v = std::getenv("SRK_SCHUR_ROWS")) e->schur_rows_enabled = std::atoi(v) != 0 ? 1 : 0;
    if (const char* v = std::getenv("SRK_BACKSUB_IMPL")) e->backsub_impl = std::atoi(v) != 0 ? 1 : 0;
    if (const char* v = std::getenv("SRK_SOLVE_TILES")) e->solve_tiles_enabled = std::atoi(v) != 0 ? 1 : 0;
    if (const char* v = std::getenv("SRK_SCHUR_TILE")) { const int t = std::atoi(v); if (t >= 16 && t <= 4096 && t % 16 == 0) e->schur_tile_fixed = t; }
    if (cudaStreamCreateWithFlags(&e->own_stream, cudaStreamNonBlocking) != cudaSuccess) { delete e; set_error("cudaStreamCreate failed"); return SRK_E_CUDA; }
    e->stream = e->own_stream;
    if (cudaStreamCreateWithFlags(&e->side_stream, cudaStreamNonBlocking) != cudaSuccess || cudaEventCreateWithFlags(&e->ev_idx, cudaEventDisableTiming) != cudaSuccess) {
        cudaGetLastError();
        if (e->side_stream != nullptr) { cudaStreamDestroy(e->side_stream); e->side_stream = nullptr; }   // everything on the main stream then
    }
    *h = e;
    return SRK_OK;
}

void srk_ba_destroy(void* h) {
    if (h == nullptr) return;
    Engine* e = (Engine*)h;
    cudaSetDevice(e->device);
    cudaStreamSynchronize(e->stream);
    srk::pcg_release(e->pcg);
    if (e->nccl_comm != nullptr && e->nccl_owned && g_nccl.CommDestroy != nullptr) g_nccl.CommDestroy(e->nccl_comm);
    if (e->h_slots != nullptr) cudaFreeHost(e->h_slots);
    if (e->h_adj != nullptr) cudaFreeHost(e->h_adj);
    for (int f = 0; f < F_COUNT; ++f) {
        for (cudaEvent_t ev : e->timers[f].pending) cudaEventDestroy(ev);
        for (cudaEvent_t ev : e->timers[f].pool) cudaEventDestroy(ev);
    }
    if (e->side_stream != nullptr) cudaStreamDestroy(e->side_stream);
    if (e->ev_idx != nullptr) cudaEventDestroy(e->ev_idx);
    if (e->own_stream != nullptr) cudaStreamDestroy(e->own_stream);
    delete e;
}

int srk_ba_set_stream(void* h, void* cuda_stream) {
    if (h == nullptr) return SRK_E_INVALID_ARG;
    Engine* e = (Engine*)h;
    cudaStreamSynchronize(e->stream);
    e->stream = cuda_stream != nullptr ? (cudaStream_t)cuda_stream : e->own_stream;
    return SRK_OK;
}

int srk_ba_set_allreduce(void* h, srk_allreduce_fn fn, void* user, int rank, int world_size) {
    if (h == nullptr || world_size < 1 || rank < 0 || rank >= world_size || world_size > 1024) { set_error("bad all-reduce arguments"); return SRK_E_INVALID_ARG; }
    Engine* e = (Engine*)h;
    e->ar = fn; e->ar_user = user; e->rank = rank; e->world = world_size;
    e->bound = false;  // slot buffers are sized at bind time
    return SRK_OK;
}

int srk_nccl_unique_id(unsigned char* id128) {
    if (id128 == nullptr) { set_error("null id buffer"); return SRK_E_INVALID_ARG; }
    if (!nccl_load()) return SRK_E_INVALID_ARG;
    NcclId id; std::memset(&id, 0, sizeof(id));
    const int rc = g_nccl.GetUniqueId(&id);
    if (rc != 0) { set_error(std::string("ncclGetUniqueId: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "error")); return SRK_E_CUDA; }
    std::memcpy(id128, id.internal, 128);
    return SRK_OK;
}

int srk_ba_set_nccl_comm(void* h, void* nccl_comm, int rank, int world_size) {
    if (h == nullptr || nccl_comm == nullptr) { set_error("null handle or communicator"); return SRK_E_INVALID_ARG; }
    if (!nccl_load()) return SRK_E_INVALID_ARG;
    Engine* e = (Engine*)h;
    int rc = srk_ba_set_allreduce(h, nccl_allreduce_cb, e, rank, world_size);
    if (rc != SRK_OK) return rc;
    if (e->nccl_comm != nullptr && e->nccl_owned) g_nccl.CommDestroy(e->nccl_comm);
    e->nccl_comm = nccl_comm; e->nccl_owned = false;
    return SRK_OK;
}

int srk_ba_nccl_init(void* h, const unsigned char* id128, int rank, int world_size) {
    if (h == nullptr || id128 == nullptr) { set_error("null handle or id"); return SRK_E_INVALID_ARG; }
    if (!nccl_load()) return SRK_E_INVALID_ARG;
    Engine* e = (Engine*)h;
    SRK_CUDA(cudaSetDevice(e->device));
    NcclId id; std::memcpy(id.internal, id128, 128);
    void* comm = nullptr;
    const int rc = g_nccl.CommInitRank(&comm, world_size, id, rank);
    if (rc != 0 || comm == nullptr) { set_error(std::string("ncclCommInitRank: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "error")); return SRK_E_CUDA; }
    int rc2 = srk_ba_set_nccl_comm(h, comm, rank, world_size);
    if (rc2 != SRK_OK) { g_nccl.CommDestroy(comm); return rc2; }
    e->nccl_owned = true;
    return SRK_OK;
}

int srk_ba_bind(void* h, const srk_ba_problem* problem, const srk_ba_options* opt) {
    if (h == nullptr) { set_error("null handle"); return SRK_E_INVALID_ARG; }
    return bind_impl(*(Engine*)h, problem, opt, true);
}

int srk_ba_run(void* h, const srk_ba_options* opt, srk_ba_report* rep) {
    if (h == nullptr) { set_error("null handle"); return SRK_E_INVALID_ARG; }
    return run_impl(*(Engine*)h, opt, rep);
}

int srk_ba_reset(void* h) {
    if (h == nullptr) { set_error("null handle"); return SRK_E_INVALID_ARG; }
    Engine& e = *(Engine*)h;
    if (!e.bound) { set_error("srk_ba_reset before srk_ba_bind"); return SRK_E_NOT_BOUND; }
    SRK_CUDA(cudaSetDevice(e.device));
    if (e.N > 0) SRK_CUDA(cudaMemcpyAsync(e.X_cur, e.Xbound.p, sizeof(double) * 3 * e.N, cudaMemcpyDeviceToDevice, e.stream));
    SRK_CUDA(cudaMemcpyAsync(e.cams_cur, e.cams_bound.p, sizeof(double) * 12 * e.M, cudaMemcpyDeviceToDevice, e.stream));
    srk::launch_cam_prep(e.stream, e.M, e.cams_cur, e.Kd.as<double>(), e.shared_K, e.f0, e.camd_cur); e.launches += 1;
    return SRK_OK;
}

int srk_ba_fetch(void* h, srk_ba_problem* problem) {
    if (h == nullptr) { set_error("null handle"); return SRK_E_INVALID_ARG; }
    return fetch_impl(*(Engine*)h, problem);
}

int srk_ba_solve(void* h, srk_ba_problem* problem, const srk_ba_options* opt, srk_ba_report* rep) {
    if (h == nullptr) { set_error("null handle"); return SRK_E_INVALID_ARG; }
    Engine& e = *(Engine*)h;
    srk_ba_options defopt; srk_ba_default_options(&defopt);
    if (opt == nullptr) opt = &defopt;
    int rc = bind_impl(e, problem, opt, true);
    if (rc < 0) return rc;
    int rc2 = run_impl(e, opt, rep);
    if (rc2 != SRK_OK) return rc2;
    if (rc == 1) return SRK_OK;  // normalisation failed: the reference returns false and leaves the scene as it was (BA.cpp:681-682)
    return fetch_impl(e, problem);
}

int srk_ba_reproj_error(void* h, const srk_ba_problem* problem, double* err, int64_t* seen_points) {
    if (h == nullptr || err == nullptr) { set_error("null argument"); return SRK_E_INVALID_ARG; }
    Engine& e = *(Engine*)h;
    int rc = bind_impl(e, problem, nullptr, false);
    if (rc < 0) return rc;
    residual_of(e, e.X_cur, e.camd_cur);
    double v = 0.0;
    rc = fetch_attempt_scalars(e, false, &v, nullptr, nullptr);
    if (rc != SRK_OK) return rc;
    *err = v;
    if (seen_points != nullptr) *seen_points = e.O;
    return SRK_OK;
}

int srk_ba_debug_derivs_and_solve(void* h, double c, double* gradE, double* E, double* G, double* Fblk, double* S, double* rhs,
                                  unsigned char* skipped, double* corrections) {
    return srk_ba_debug_derivs_and_solve_ex(h, c, SRK_SOLVER_DENSE_CHOLESKY, gradE, E, G, Fblk, S, rhs, skipped, corrections, nullptr);
}

int srk_ba_debug_derivs_and_solve_ex(void* h, double c, int32_t solver, double* gradE, double* E, double* G, double* Fblk, double* S, double* rhs,
                                     unsigned char* skipped, double* corrections, int32_t* pcg_iters) {
    if (h == nullptr) { set_error("null handle"); return SRK_E_INVALID_ARG; }
    Engine& e = *(Engine*)h;
    if (!e.bound) { set_error("debug call before srk_ba_bind"); return SRK_E_NOT_BOUND; }
    SRK_CUDA(cudaSetDevice(e.device));
    srk_ba_options opt; srk_ba_default_options(&opt);
    if (solver != SRK_SOLVER_BLOCK_PCG) solver = SRK_SOLVER_DENSE_CHOLESKY;
    opt.solver = solver;
    int rc = ensure_solver_buffers(e, SRK_SOLVER_DENSE_CHOLESKY);   // the dense buffers also receive the PCG system for inspection
    if (rc != SRK_OK) return rc;
    rc = derivative_pass(e);
    if (rc != SRK_OK) return rc;
    cudaStream_t st = e.stream;
    const int64_t N = e.N, O = e.O; const int M = e.M;
    size_t need = sizeof(double) * (size_t)(12 * N + 30 * O + 3 * N + 16);
    SRK_CUDA(e.dbg.ensure(need));
    double* dE = e.dbg.as<double>(); double* dgp = dE + 9 * N; double* dF = dgp + 3 * N; double* ddp = dF + 30 * O;
    srk::launch_debug_point_blocks(st, N, O, e.pt_begin.as<int64_t>(), e.J.as<double>(), dE, dgp);
    srk::launch_debug_F_blocks(st, O, e.J.as<double>(), dF);
    e.launches += 2;
    if (gradE != nullptr) {
        if (N > 0) SRK_CUDA(cudaMemcpyAsync(gradE, dgp, sizeof(double) * 3 * N, cudaMemcpyDeviceToHost, st));
        SRK_CUDA(cudaMemcpyAsync(gradE + 3 * N, e.Ggf.as<double>() + 100 * (size_t)M, sizeof(double) * 10 * M, cudaMemcpyDeviceToHost, st));
    }
    if (E != nullptr && N > 0) SRK_CUDA(cudaMemcpyAsync(E, dE, sizeof(double) * 9 * N, cudaMemcpyDeviceToHost, st));
    if (G != nullptr) SRK_CUDA(cudaMemcpyAsync(G, e.Ggf.p, sizeof(double) * 100 * M, cudaMemcpyDeviceToHost, st));
    if (Fblk != nullptr && O > 0) SRK_CUDA(cudaMemcpyAsync(Fblk, dF, sizeof(double) * 30 * O, cudaMemcpyDeviceToHost, st));
    SRK_CUDA(cudaStreamSynchronize(st));
    if (c >= 0) {
        rc = attempt(e, solver, &opt, c, ddp);
        if (rc != SRK_OK) return rc;
        const int nf = e.nf; const int64_t ld = e.ld;
        double* dS = e.Srhs.as<double>();
        if (solver == SRK_SOLVER_BLOCK_PCG) {
            rc = srk::pcg_debug_to_dense(e.pcg, st, M, e.unity, dS, ld, dS + (size_t)ld * nf, &e.launches);
            if (rc != SRK_OK) { set_error("pcg_debug_to_dense failed"); return rc; }
            if (pcg_iters != nullptr) *pcg_iters = e.pcg_iters_last;
        } else if (opt.refine_steps <= 0 || (e.order.active && e.order_tiles)) { srk::launch_mirror_lower(st, nf, dS, ld); e.launches += 1; }
        e.S_clean = false;   // the inspection copy fills tiles the solve never clears
        if (S != nullptr) SRK_CUDA(cudaMemcpy2DAsync(S, sizeof(double) * nf, dS, sizeof(double) * ld, sizeof(double) * nf, nf, cudaMemcpyDeviceToHost, st));
        if (rhs != nullptr) SRK_CUDA(cudaMemcpyAsync(rhs, dS + (size_t)ld * nf, sizeof(double) * nf, cudaMemcpyDeviceToHost, st));
        if (skipped != nullptr && N > 0) SRK_CUDA(cudaMemcpyAsync(skipped, e.skipped.p, (size_t)N, cudaMemcpyDeviceToHost, st));
        if (corrections != nullptr) {
            if (N > 0) SRK_CUDA(cudaMemcpyAsync(corrections, ddp, sizeof(double) * 3 * N, cudaMemcpyDeviceToHost, st));
            SRK_CUDA(cudaMemcpyAsync(corrections + 3 * N, e.dfull.p, sizeof(double) * 10 * M, cudaMemcpyDeviceToHost, st));
        }
        SRK_CUDA(cudaStreamSynchronize(st));
    }
    SRK_CUDA(cudaGetLastError());
    return SRK_OK;
}

int srk_ba_debug_get_state(void* h, double* points, double* cams) {
    if (h == nullptr) { set_error("null handle"); return SRK_E_INVALID_ARG; }
    Engine& e = *(Engine*)h;
    if (!e.bound) { set_error("debug call before srk_ba_bind"); return SRK_E_NOT_BOUND; }
    SRK_CUDA(cudaSetDevice(e.device));
    if (points != nullptr && e.N > 0) {
        srk::launch_planes_to_points(e.stream, e.N, e.X_cur, e.pts_stage.as<double>()); e.launches += 1;
        SRK_CUDA(cudaMemcpyAsync(points, e.pts_stage.p, sizeof(double) * 3 * e.N, cudaMemcpyDeviceToHost, e.stream));
    }
    if (cams != nullptr) SRK_CUDA(cudaMemcpyAsync(cams, e.cams_cur, sizeof(double) * 12 * e.M, cudaMemcpyDeviceToHost, e.stream));
    SRK_CUDA(cudaStreamSynchronize(e.stream));
    return SRK_OK;
}

int srk_ba_debug_apply(void* h, const double* corrections, double* err_new) {
    if (h == nullptr || corrections == nullptr) { set_error("null argument"); return SRK_E_INVALID_ARG; }
    Engine& e = *(Engine*)h;
    if (!e.bound) { set_error("debug call before srk_ba_bind"); return SRK_E_NOT_BOUND; }
    SRK_CUDA(cudaSetDevice(e.device));
    cudaStream_t st = e.stream;
    SRK_CUDA(e.dfull.ensure(sizeof(double) * 10 * (size_t)e.M));
    if (e.N > 0) {
        SRK_CUDA(cudaMemcpyAsync(e.pts_stage.p, corrections, sizeof(double) * 3 * e.N, cudaMemcpyHostToDevice, st));
        srk::launch_add_points_aos(st, e.N, e.X_cur, e.pts_stage.as<double>(), e.X_try); e.launches += 1;
    }
    SRK_CUDA(cudaMemcpyAsync(e.dfull.p, corrections + 3 * e.N, sizeof(double) * 10 * e.M, cudaMemcpyHostToDevice, st));
    srk::launch_cam_update(st, e.M, e.cams_cur, e.dfull.as<double>(), e.cams_try);
    srk::launch_cam_prep(st, e.M, e.cams_try, e.Kd.as<double>(), e.shared_K, e.f0, e.camd_try);
    e.launches += 2;
    accept_trial(e);
    residual_of(e, e.X_cur, e.camd_cur);
    double v = 0.0;
    int rc = fetch_attempt_scalars(e, false, &v, nullptr, nullptr);
    if (rc != SRK_OK) return rc;
    if (err_new != nullptr) *err_new = v;
    return SRK_OK;
}

int srk_ba_solve_stats(void* h, int64_t* n_f, int64_t* block_rows, int64_t* nonzero_tiles, double* factor_flops) {
    if (h == nullptr) { set_error("null handle"); return SRK_E_INVALID_ARG; }
    Engine& e = *(Engine*)h;
    if (!e.bound || e.solver_used != SRK_SOLVER_DENSE_CHOLESKY || e.dinv.p == nullptr) { set_error("no dense factor on this handle"); return SRK_E_NOT_BOUND; }
    SRK_CUDA(cudaSetDevice(e.device));
    if (n_f != nullptr) *n_f = e.nf;
    if (srk::dense_cholesky_stats(e.stream, (e.order_ready && e.order.active) ? e.order.np : e.nf, e.dinv.as<double>(), block_rows, nonzero_tiles, factor_flops) != 0) { set_error("could not read the factor structure"); return SRK_E_CUDA; }
    return SRK_OK;
}

int srk_ba_debug_build_order(int32_t n_groups, const int32_t* group_size, const unsigned char* adj, int32_t* pos, int64_t* ordered_n, int32_t* part_k0,
                             int32_t* part_k1, int32_t* ksep) {
    if (n_groups <= 0 || group_size == nullptr || adj == nullptr) { set_error("null argument"); return SRK_E_INVALID_ARG; }
    std::vector<int> gs(group_size, group_size + n_groups);
    const srk::SolveOrder o = srk::build_solve_order(n_groups, gs.data(), adj);
    if (ordered_n != nullptr) *ordered_n = o.active ? o.np : o.n;
    if (!o.active) {
        if (pos != nullptr) for (int i = 0; i < o.n; ++i) pos[i] = i;
        if (ksep != nullptr) *ksep = 0;
        return 0;
    }
    if (pos != nullptr) for (int i = 0; i < o.n; ++i) pos[i] = o.pos[(size_t)i];
    for (int p = 0; p < o.part.nparts; ++p) { if (part_k0 != nullptr) part_k0[p] = o.part.k0[p]; if (part_k1 != nullptr) part_k1[p] = o.part.k1[p]; }
    if (ksep != nullptr) *ksep = o.part.ksep;
    return o.part.nparts;
}

int srk_ba_solve_order(void* h, int64_t* ordered_n, int64_t* parts, int64_t* max_part_blocks, int64_t* separator_blocks) {
    if (h == nullptr) { set_error("null handle"); return SRK_E_INVALID_ARG; }
    Engine& e = *(Engine*)h;
    if (!e.bound) { set_error("no problem bound"); return SRK_E_NOT_BOUND; }
    const bool on = e.order_ready && e.order.active;
    if (ordered_n != nullptr) *ordered_n = on ? e.order.np : e.nf;
    if (parts != nullptr) *parts = on ? e.order.part.nparts : 0;
    if (max_part_blocks != nullptr) *max_part_blocks = on ? e.order.max_part_blocks : 0;
    if (separator_blocks != nullptr) *separator_blocks = on ? e.order.sep_blocks : 0;
    return SRK_OK;
}

int srk_ba_solve_levels(void* h, int64_t* mid_separators, int64_t* max_mid_blocks) {
    if (h == nullptr) { set_error("null handle"); return SRK_E_INVALID_ARG; }
    Engine& e = *(Engine*)h;
    if (!e.bound) { set_error("no problem bound"); return SRK_E_NOT_BOUND; }
    const bool on = e.order_ready && e.order.active;
    if (mid_separators != nullptr) *mid_separators = on ? e.order.part.nmids : 0;
    if (max_mid_blocks != nullptr) *max_mid_blocks = on ? e.order.max_mid_blocks : 0;
    return SRK_OK;
}

int srk_ba_debug_order_levels(int32_t n_groups, const int32_t* group_size, const unsigned char* adj, int32_t* mid_k0, int32_t* mid_k1, int32_t* msep) {
    if (n_groups <= 0 || group_size == nullptr || adj == nullptr) { set_error("null argument"); return SRK_E_INVALID_ARG; }
    std::vector<int> gs(group_size, group_size + n_groups);
    const srk::SolveOrder o = srk::build_solve_order(n_groups, gs.data(), adj);
    if (!o.active) { if (msep != nullptr) *msep = 0; return 0; }
    for (int p = 0; p < o.part.nmids; ++p) { if (mid_k0 != nullptr) mid_k0[p] = o.part.m0[p]; if (mid_k1 != nullptr) mid_k1[p] = o.part.m1[p]; }
    if (msep != nullptr) *msep = o.part.msep;
    return o.part.nmids;
}

int srk_ba_pcg_stats(void* h, int64_t* nnz_blocks, int64_t* iters_since_timing) {
    if (h == nullptr) { set_error("null handle"); return SRK_E_INVALID_ARG; }
    Engine& e = *(Engine*)h;
    if (nnz_blocks != nullptr) *nnz_blocks = e.pcg.structure_valid ? e.pcg.nnzb : 0;
    if (iters_since_timing != nullptr) *iters_since_timing = e.pcg_iters_total;
    return SRK_OK;
}

int srk_ba_set_timing(void* h, int enabled) {
    if (h == nullptr) return SRK_E_INVALID_ARG;
    Engine& e = *(Engine*)h;
    resolve_timers(e);
    e.timing = enabled != 0;
    if (enabled) e.pcg_iters_total = 0;
    for (int f = 0; f < F_COUNT; ++f) { e.timers[f].total_ms = 0.0; e.timers[f].last_ms = 0.0; e.timers[f].count = 0; }
    return SRK_OK;
}

int srk_ba_get_timing(void* h, const char* name, double* ms_last, double* ms_total, int64_t* launches) {
    if (h == nullptr || name == nullptr) return SRK_E_INVALID_ARG;
    Engine& e = *(Engine*)h;
    resolve_timers(e);
    for (int f = 0; f < F_COUNT; ++f)
        if (std::strcmp(name, kFamilyNames[f]) == 0) {
            if (ms_last != nullptr) *ms_last = e.timers[f].last_ms;
            if (ms_total != nullptr) *ms_total = e.timers[f].total_ms;
            if (launches != nullptr) *launches = e.timers[f].count;
            return SRK_OK;
        }
    set_error("unknown kernel family");
    return SRK_E_INVALID_ARG;
}

}  // extern "C"
