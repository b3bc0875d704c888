// suriko-b200 — hand-written sm_100a kernels of the bundle-adjustment hot path.
//
//   k_cam_prep        per-camera derived record (P = K[R|T], direct pose, rot1..3)      BA.cpp:1193-1197, :1499-1520
//   k_jacobian   (K1) per-observation residual + analytic Jacobian rows, SoA planes     BA.cpp:1160-1266 / :1268-1412 / :1450-1549
//   k_frame_blocks    per-camera 10x10 block G and gradient g_f (the reference's "frame pass")  BA.cpp:1268-1334
//   k_residual  (K1') per-observation squared residual + deterministic reduction         BA.cpp:410-490
//   k_fill_reduced    S <- gauge-reduced, damped G;  rhs <- -g_f                           BA.cpp:1780-1823, :1902-1908
//   k_schur      (K2) per-point 3x3 block, cofactor inverse, Schur accumulation          BA.cpp:1859-1900
//   k_backsub   (K2') per-point back-substitution + point update into the trial state   BA.cpp:1919-1960, :2003-2017
//   k_cam_update (K4) per-camera pose update (direct T += dT, R <- Rodrigues(dW) R)       BA.cpp:2021-2062, :59-92
//   k_normalize_points / k_revert_points   gauge normalisation of the points               BA.cpp:179-199
//
// Data layout in HBM (all FP64 values, int32 indices): observations point-major (pnt_ind, frame_ind) — the reference's
// own track order — as SoA arrays; Jacobian rows as 28 planes of n_obs doubles so that every warp access is one
// contiguous 256-B segment; points as 3 planes.
#include "common.cuh"
#include "kernels.h"

namespace srk {

// ---------------------------------------------------------------------------------------------------------------------
__global__ void k_cam_prep(int M, const double* __restrict__ cams, const double* __restrict__ K, int shared_K, double f0,
                           double* __restrict__ camd) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M) return;
    const double* c = cams + (size_t)i * 12;  // T[3], R col-major[9]
    const double* k = K + (shared_K ? 0 : (size_t)i * 9);
    double* d = camd + (size_t)i * kCamStride;
    double R[9], T[3], Km[9];
#pragma unroll
    for (int j = 0; j < 3; ++j) T[j] = c[j];
#pragma unroll
    for (int j = 0; j < 9; ++j) { R[j] = c[3 + j]; Km[j] = k[j]; }
#pragma unroll
    for (int j = 0; j < 9; ++j) { d[CD_R + j] = R[j]; d[CD_K + j] = Km[j]; }
#pragma unroll
    for (int j = 0; j < 3; ++j) d[CD_T + j] = T[j];
    // KR(r,c) = sum_m K(r,m) R(m,c)
#pragma unroll
    for (int cc = 0; cc < 3; ++cc)
#pragma unroll
        for (int r = 0; r < 3; ++r) d[CD_KR + cc * 3 + r] = Km[0 * 3 + r] * R[cc * 3 + 0] + Km[1 * 3 + r] * R[cc * 3 + 1] + Km[2 * 3 + r] * R[cc * 3 + 2];
    // direct pose: Rd = R^T, Td = -Rd*T   (obs-geom.cpp:117-122);  Rd(i,c) = R(c,i)
    double Td[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) Td[r] = -(R[r * 3 + 0] * T[0] + R[r * 3 + 1] * T[1] + R[r * 3 + 2] * T[2]);
    const double fx = Km[0], fy = Km[4], u0 = Km[6], v0 = Km[7];
#pragma unroll
    for (int t = 0; t < 3; ++t) {
        double rd0 = R[t * 3 + 0], rd1 = R[t * 3 + 1], rd2 = R[t * 3 + 2];  // Rd(t,0..2) = R(0..2,t)
        d[CD_TD + t] = Td[t];
        d[CD_ROT1 + t] = fx * rd0 + u0 * rd2;
        d[CD_ROT2 + t] = fy * rd1 + v0 * rd2;
        d[CD_ROT3 + t] = f0 * rd2;
    }
    d[CD_IFX] = 1.0 / fx; d[CD_CU] = u0 / (f0 * fx); d[CD_IFY] = 1.0 / fy; d[CD_CV] = v0 / (f0 * fy); d[CD_IF0] = 1.0 / f0; d[47] = 0.0;
}

// ---------------------------------------------------------------------------------------------------------------------
// K1: a CTA of 256 threads owns a chunk of kObsRows*256 consecutive observations (thread t: observations base + t + 256*row).
// Reads 24 B/obs of observation data + the point (gathered, L1-served: a point's observations are consecutive) + the camera
// record through the per-CTA camera table (shared memory), writes 224 B/obs as 28 coalesced planes: rho[2], Jp[6], Jc[20].
constexpr int kObsRows = 4;
__global__ void __launch_bounds__(256) k_jacobian(int64_t O, const int32_t* __restrict__ obs_cam, const int32_t* __restrict__ obs_pt,
                                                  const double* __restrict__ obs_xs, const double* __restrict__ obs_ys,
                                                  const double* __restrict__ X, int64_t N, const double* __restrict__ camd,
                                                  double* __restrict__ J, double* __restrict__ Eacc) {
    __shared__ CamTable tab;
    __shared__ double s_red[8][9][33];     // per warp: the nine per-observation terms of a row of 32 observations (Eacc path)
    __shared__ int s_seg[8][33], s_segpt[8][33];
    const int64_t base = (int64_t)blockIdx.x * (256 * kObsRows) + threadIdx.x;
    const int lane = threadIdx.x & 31;
    cam_table_reset(tab);
    int cam[kObsRows], hp[kObsRows];
#pragma unroll
    for (int rr = 0; rr < kObsRows; ++rr) { const int64_t o = base + 256 * rr; cam[rr] = o < O ? obs_cam[o] : -1; }
    __syncthreads();
#pragma unroll
    for (int rr = 0; rr < kObsRows; ++rr) hp[rr] = cam[rr] >= 0 ? cam_table_insert(tab, cam[rr]) : -1;
    __syncthreads();
    cam_table_stage(tab, camd);
    __syncthreads();
#pragma unroll
    for (int rr = 0; rr < kObsRows; ++rr) {
        const int64_t o = base + 256 * rr;
        const bool valid = o < O;
        int pt = -1;
        double a9[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) a9[i] = 0.0;
        if (valid) {
            pt = obs_pt[o];
            const double xs = obs_xs[o], ys = obs_ys[o];
            const double X0 = X[pt], X1 = X[N + pt], X2 = X[2 * N + pt];
            const double* cd = cam_table_record(tab, hp[rr], cam[rr], camd);
            double rx, ry, jp[6], jc[20];
            obs_jacobian(cd, X0, X1, X2, xs, ys, rx, ry, jp, jc);
            J[o] = rx;
            J[O + o] = ry;
#pragma unroll
            for (int i = 0; i < 6; ++i) J[(int64_t)(2 + i) * O + o] = jp[i];
#pragma unroll
            for (int i = 0; i < 20; ++i) J[(int64_t)(8 + i) * O + o] = jc[i];
            if (Eacc != nullptr) {     // this observation's terms of E_j / 2 = sum Jp^T Jp and g_pj / 2 = sum Jp^T rho (BA.cpp:1184-1220)
                a9[0] = jp[0] * jp[0] + jp[1] * jp[1];
                a9[1] = jp[0] * jp[2] + jp[1] * jp[3];
                a9[2] = jp[0] * jp[4] + jp[1] * jp[5];
                a9[3] = jp[2] * jp[2] + jp[3] * jp[3];
                a9[4] = jp[2] * jp[4] + jp[3] * jp[5];
                a9[5] = jp[4] * jp[4] + jp[5] * jp[5];
                a9[6] = jp[0] * rx + jp[1] * ry;
                a9[7] = jp[2] * rx + jp[3] * ry;
                a9[8] = jp[4] * rx + jp[5] * ry;
            }
        }
        if (Eacc != nullptr) {
            // Observations are point-major: the lanes of one point are contiguous.  The warp parks its 32 x 9 terms in shared memory; one
            // lane per (point segment, value) then adds the segment's terms in lane order and issues ONE red.global.add.f64 -- ~10 shared
            // loads per task instead of a 5-step shuffle scan of all nine values on every lane (measured: K1 0.42 -> 0.61 ms with the scan).
            // A point's observations meet in at most a few warps; the partial sums land on exact zeros and floating-point addition
            // commutes, so the result does not depend on the arrival order for up to two partials (tracks of at most 32 observations).
            // The per-point blocks are consumed by k_point_finish (schur_v3.cu).
            const int wq = threadIdx.x >> 5;
            const int ptu = __shfl_up_sync(0xffffffffu, pt, 1);
            const bool head = valid && (lane == 0 || ptu != pt);
            const unsigned heads = __ballot_sync(0xffffffffu, head);
            const int nvalid = __popc(__ballot_sync(0xffffffffu, valid));
            const int nseg = __popc(heads);
#pragma unroll
            for (int i = 0; i < 9; ++i) s_red[wq][i][lane] = a9[i];
            if (head) { const int idx = __popc(heads & ((1u << lane) - 1u)); s_seg[wq][idx] = lane; s_segpt[wq][idx] = pt; }
            __syncwarp();
            for (int t0 = 0; t0 < 9 * nseg; t0 += 32) {
                const int t = t0 + lane;
                if (t < 9 * nseg) {
                    const int seg = t / 9, i = t - 9 * seg;
                    const int start = s_seg[wq][seg];
                    const int end = seg + 1 < nseg ? s_seg[wq][seg + 1] : nvalid;
                    double sum = 0.0;
                    for (int l = start; l < end; ++l) sum += s_red[wq][i][l];
                    atomicAdd(&Eacc[(int64_t)i * N + s_segpt[wq][seg]], sum);
                }
            }
            __syncwarp();
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Frame pass: camera-major.  grid = (M, splits); each CTA walks a slice of the camera's observations (camera-major copy
// of the observation list: point index + pixel), recomputes the 2x10 row block and accumulates the 55 unique entries of
// 2*Jc^T*Jc and the 10 of 2*Jc^T*rho in registers; one block reduction, one store (splits == 1) or 65 atomics.
__global__ void __launch_bounds__(128) k_frame_blocks(int M, const int64_t* __restrict__ cam_begin, const int32_t* __restrict__ c_pt,
                                                      const double* __restrict__ c_x, const double* __restrict__ c_y,
                                                      const double* __restrict__ X, int64_t N, const double* __restrict__ camd,
                                                      double* __restrict__ G, double* __restrict__ gf, int splits) {
    int cam = blockIdx.x;
    const double* cd = camd + (size_t)cam * kCamStride;
    int64_t b = cam_begin[cam], e = cam_begin[cam + 1];
    int64_t len = e - b;
    int64_t sb = b + (len * blockIdx.y) / splits, se = b + (len * (blockIdx.y + 1)) / splits;
    double acc[65];
#pragma unroll
    for (int i = 0; i < 65; ++i) acc[i] = 0.0;
    // Two-deep software pipeline: the gather X[pt] of observation i+1 and the index / pixel loads of observation i+2 are in flight
    // while observation i is evaluated (one CTA keeps 4 warps per camera slice: without it every iteration waits for two dependent
    // L2 round trips -- ncu: long-scoreboard 4.2 stalls per issue).
    const int stride = blockDim.x;
    int64_t o = sb + threadIdx.x;
    bool v1 = o < se;
    int pt1 = 0; double x1 = 0.0, y1 = 0.0;
    if (v1) { pt1 = c_pt[o]; x1 = c_x[o]; y1 = c_y[o]; }
    bool vc = v1;
    double X0 = 0.0, X1 = 0.0, X2 = 0.0, xc = x1, yc = y1;
    if (vc) { X0 = X[pt1]; X1 = X[N + pt1]; X2 = X[2 * N + pt1]; }
    o += stride; v1 = o < se;
    if (v1) { pt1 = c_pt[o]; x1 = c_x[o]; y1 = c_y[o]; }
    while (vc) {
        double nX0 = 0.0, nX1 = 0.0, nX2 = 0.0;
        const bool nv = v1; const double nx = x1, ny = y1;
        if (nv) { nX0 = X[pt1]; nX1 = X[N + pt1]; nX2 = X[2 * N + pt1]; }
        o += stride; v1 = o < se;
        if (v1) { pt1 = c_pt[o]; x1 = c_x[o]; y1 = c_y[o]; }
        double rx, ry, jp[6], jc[20];
        obs_jacobian(cd, X0, X1, X2, xc, yc, rx, ry, jp, jc);
        int idx = 0;
#pragma unroll
        for (int a = 0; a < 10; ++a)
#pragma unroll
            for (int bb = a; bb < 10; ++bb) { acc[idx] += jc[a * 2] * jc[bb * 2] + jc[a * 2 + 1] * jc[bb * 2 + 1]; ++idx; }
#pragma unroll
        for (int a = 0; a < 10; ++a) acc[55 + a] += jc[a * 2] * rx + jc[a * 2 + 1] * ry;
        X0 = nX0; X1 = nX1; X2 = nX2; xc = nx; yc = ny; vc = nv;
    }
    __shared__ double red[4][65];
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < 65; ++i) {
        double v = acc[i];
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
        if (lane == 0) red[w][i] = v;
    }
    __syncthreads();
    if (threadIdx.x < 65) {
        double v = 2.0 * ((red[0][threadIdx.x] + red[1][threadIdx.x]) + (red[2][threadIdx.x] + red[3][threadIdx.x]));
        int i = threadIdx.x;
        if (i < 55) {
            int a = 0, rem = i;
            while (rem >= 10 - a) { rem -= 10 - a; ++a; }
            int bb = a + rem;
            double* g = G + (size_t)cam * 100;
            if (splits == 1) { g[a * 10 + bb] = v; g[bb * 10 + a] = v; }
            else { atomicAdd(&g[a * 10 + bb], v); if (a != bb) atomicAdd(&g[bb * 10 + a], v); }
        } else {
            if (splits == 1) gf[(size_t)cam * 10 + (i - 55)] = v;
            else atomicAdd(&gf[(size_t)cam * 10 + (i - 55)], v);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// K1': squared residuals.  Fixed grid; a CTA walks chunks of kObsRows*256 observations (camera table per chunk), keeps one
// partial per thread, then a per-block partial and a fixed-order final sum (deterministic for a given grid).
constexpr int kResRows = 4;    // observations per thread and chunk in K1' (8 was measured: 104 registers, 2 CTAs per SM, 0.19 ms instead of 0.12 ms)
constexpr int kResChunk = 256 * kResRows;

// Bind-time structure pass for K1': the distinct cameras of every chunk of kResChunk observations (at most kCamTabSlots of them get a
// shared-memory slot) and, per observation, the slot of its camera (255 = read the record from global memory).  The observation order
// never changes during a solve, so K1' itself needs no hash set, no insert phase and reads 1 byte instead of the 4-byte camera id.
__global__ void __launch_bounds__(256) k_chunk_tables(int64_t O, const int32_t* __restrict__ obs_cam, const int32_t* __restrict__ obs_pt, int* __restrict__ chunk_cams,
                                                      int* __restrict__ chunk_cnt, int2* __restrict__ chunk_pts, unsigned char* __restrict__ obs_slot) {
    __shared__ CamTable tab;
    const int64_t ch = blockIdx.x;
    const int64_t base = ch * kResChunk + threadIdx.x;
    cam_table_reset(tab);
    int cam[kResRows], hp[kResRows];
#pragma unroll
    for (int rr = 0; rr < kResRows; ++rr) { const int64_t o = base + 256 * rr; cam[rr] = o < O ? obs_cam[o] : -1; }
    __syncthreads();
#pragma unroll
    for (int rr = 0; rr < kResRows; ++rr) hp[rr] = cam[rr] >= 0 ? cam_table_insert(tab, cam[rr]) : -1;
    __syncthreads();
    const int n = min(tab.count, kCamTabSlots);
    if (threadIdx.x < kCamTabSlots) chunk_cams[ch * kCamTabSlots + threadIdx.x] = (int)threadIdx.x < n ? tab.cam_of_slot[threadIdx.x] : -1;
    if (threadIdx.x == 0) {
        chunk_cnt[ch] = n;
        // observations are point-major: the chunk's points are the contiguous id range [first, last]
        const int64_t o0 = ch * kResChunk, o1 = min(O, o0 + kResChunk) - 1;
        const int p0 = obs_pt[o0], p1 = obs_pt[o1];
        chunk_pts[ch] = make_int2(p0, p1 >= p0 ? p1 - p0 + 1 : 0);
    }
#pragma unroll
    for (int rr = 0; rr < kResRows; ++rr) {
        const int64_t o = base + 256 * rr;
        if (o >= O) continue;
        const int sl = hp[rr] >= 0 ? tab.slot[hp[rr]] : -1;
        obs_slot[o] = sl >= 0 ? (unsigned char)sl : (unsigned char)255;
    }
}

// K1': squared residuals.  Fixed grid; a CTA walks chunks of kResChunk observations.  Every global load of a chunk is issued up
// front -- slot / point id / pixel of each observation, the chunk's camera list -> camera records into shared memory (odd stride), the
// point gather as soon as the ids arrive -- ONE barrier, then arithmetic.  One partial per thread, then a per-block partial and a
// fixed-order final sum (deterministic for a given grid; same summation order as before the tables existed).
__global__ void __launch_bounds__(256) k_residual(int64_t O, const int32_t* __restrict__ obs_cam, const int32_t* __restrict__ obs_pt,
                                                  const double* __restrict__ obs_xs, const double* __restrict__ obs_ys,
                                                  const double* __restrict__ X, int64_t N, const double* __restrict__ camd,
                                                  const int* __restrict__ chunk_cams, const int* __restrict__ chunk_cnt, const int2* __restrict__ chunk_pts,
                                                  const unsigned char* __restrict__ obs_slot, double* __restrict__ partial) {
    __shared__ double rec[kCamTabSlots * kCamRecPad];
    __shared__ double xsl[3 * kResChunk];      // the chunk's points (a chunk of kResChunk observations spans at most as many point ids)
    double s = 0.0;
    const int64_t nchunks = (O + kResChunk - 1) / kResChunk;
    for (int64_t ch = blockIdx.x; ch < nchunks; ch += gridDim.x) {
        const int64_t base = ch * kResChunk + threadIdx.x;
        int pt[kResRows], sl[kResRows]; double xs[kResRows], ys[kResRows];
#pragma unroll
        for (int rr = 0; rr < kResRows; ++rr) {
            const int64_t o = base + 256 * rr;
            const bool in = o < O;
            sl[rr] = in ? (int)obs_slot[o] : -1; pt[rr] = in ? obs_pt[o] : 0; xs[rr] = in ? obs_xs[o] : 0.0; ys[rr] = in ? obs_ys[o] : 0.0;
        }
        const int n = chunk_cnt[ch] * kCamStride;
        const int2 pr = chunk_pts[ch];
        const bool staged = pr.y <= kResChunk;   // false only when point ids without observations stretch the range: gather from global then
        __syncthreads();                      // the previous chunk's records and points are no longer read
        for (int e = threadIdx.x; e < n; e += 256) {
            const int sidx = e / kCamStride, f = e - sidx * kCamStride;
            rec[sidx * kCamRecPad + f] = camd[(size_t)chunk_cams[ch * kCamTabSlots + sidx] * kCamStride + f];
        }
        if (staged) {                         // coalesced, and independent of the obs_pt loads above: one round trip for everything
            for (int e = threadIdx.x; e < pr.y; e += 256) {
                xsl[e] = X[pr.x + e]; xsl[kResChunk + e] = X[N + pr.x + e]; xsl[2 * kResChunk + e] = X[2 * N + pr.x + e];
            }
        }
        __syncthreads();
#pragma unroll
        for (int rr = 0; rr < kResRows; ++rr) {
            if (sl[rr] < 0) continue;
            const double* cd = sl[rr] < 255 ? rec + sl[rr] * kCamRecPad : camd + (size_t)obs_cam[base + 256 * rr] * kCamStride;
            double X0r, X1r, X2r;
            if (staged) { const int q = pt[rr] - pr.x; X0r = xsl[q]; X1r = xsl[kResChunk + q]; X2r = xsl[2 * kResChunk + q]; }
            else { X0r = X[pt[rr]]; X1r = X[N + pt[rr]]; X2r = X[2 * N + pt[rr]]; }
            double rx, ry;
            obs_residual(cd, X0r, X1r, X2r, xs[rr], ys[rr], rx, ry);
            s += rx * rx + ry * ry;
        }
    }
    __shared__ double red[8];
#pragma unroll
    for (int sft = 16; sft > 0; sft >>= 1) s += __shfl_xor_sync(0xffffffffu, s, sft);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
        partial[blockIdx.x] = t;
    }
}
__global__ void k_sum_partials(int n, const double* __restrict__ partial, double* __restrict__ out) {
    __shared__ double sm[256];
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) s += partial[i];
    sm[threadIdx.x] = s;
    __syncthreads();
    for (int st = 128; st > 0; st >>= 1) { if ((int)threadIdx.x < st) sm[threadIdx.x] += sm[threadIdx.x + st]; __syncthreads(); }
    if (threadIdx.x == 0) out[0] = sm[0];
}

// ---------------------------------------------------------------------------------------------------------------------
// S <- gauge-reduced G with the diagonal multiplied by (1+c) (fill_matG, BA.cpp:1780-1823); rhs <- -g_f (BA.cpp:1908).
// One CTA per camera.  S is dense column-major n_f x n_f (already zeroed).  with_G == 0 (ranks > 0 in a multi-GPU run)
// leaves both untouched so that the all-reduced sum contains G exactly once.
__global__ void k_fill_reduced(int M, const double* __restrict__ G, const double* __restrict__ gf, double c, int unity,
                               double* __restrict__ S, int64_t ld, double* __restrict__ rhs) {
    int cam = blockIdx.x;
    int t = threadIdx.x;
    if (t < 100) {
        int a = t / 10, b = t % 10;
        int ra = red_index(cam, a, unity), rb = red_index(cam, b, unity);
        if (ra >= 0 && rb >= 0) {
            double g = G[(size_t)cam * 100 + t];
            if (a == b) g *= 1.0 + c;
            S[(size_t)rb * ld + ra] = g;
        }
    } else if (t < 110) {
        int a = t - 100;
        int ra = red_index(cam, a, unity);
        if (ra >= 0) rhs[ra] = -gf[(size_t)cam * 10 + a];
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// K2: one warp per point.
//  phase 1: E = 2*sum Jp^T Jp, g_p = 2*sum Jp^T rho over the point's observations (lanes over observations, shuffle
//           reduction), damping, cofactor inverse with the |det| > 1e-12 rule; Einv/g_p/flag are stored for K2'.
//  phase 2: for every observation i: F_i = 2*Jp_i^T*Jc_i (3x10), W_i = Einv*F_i, staged in shared memory in chunks of
//           CH observations; every pair (i >= l) contributes the 10x10 block  -F_i^T W_l  to S[cam_i, cam_l] (lower
//           triangle, cam_i >= cam_l because observations are frame-sorted within a point) as 5x5 register tiles,
//           accumulated with red.global.add.f64; rhs[cam_i] += F_i^T (Einv g_p).
constexpr int kSchurCH = 16;
constexpr int kSchurWarps = 4;

__global__ void __launch_bounds__(kSchurWarps * 32) k_schur(int64_t N, int64_t O, const int64_t* __restrict__ pt_begin,
                                                            const int32_t* __restrict__ obs_cam, const double* __restrict__ J, double c,
                                                            SchurSink sink, double* __restrict__ pinv, unsigned char* __restrict__ skipped,
                                                            const unsigned char* __restrict__ only_flagged, const int* __restrict__ list,
                                                            const int* __restrict__ list_count) {
    __shared__ double sF[kSchurWarps][kSchurCH][30];   // slot 0 only: F of the row chunk
    __shared__ double sW[kSchurWarps][2][kSchurCH][30];
    __shared__ int sCam[kSchurWarps][2][kSchurCH];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int64_t j = (int64_t)blockIdx.x * kSchurWarps + w;
    if (list != nullptr) {   // exception list of the tile kernel (schur_v3.cu): a device-side count, normally zero
        if (j >= (int64_t)*list_count) return;
        j = list[j];
    }
    if (j >= N) return;
    if (only_flagged != nullptr && only_flagged[j] == 0) return;
    const int64_t b = pt_begin[j], e = pt_begin[j + 1];
    const int k = (int)(e - b);

    // ---- phase 1
    double a9[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) a9[i] = 0.0;
    for (int64_t o = b + lane; o < e; o += 32) {
        double rx = J[o], ry = J[O + o];
        double jp[6];
#pragma unroll
        for (int i = 0; i < 6; ++i) jp[i] = J[(int64_t)(2 + i) * O + o];
        a9[0] += jp[0] * jp[0] + jp[1] * jp[1];
        a9[1] += jp[0] * jp[2] + jp[1] * jp[3];
        a9[2] += jp[0] * jp[4] + jp[1] * jp[5];
        a9[3] += jp[2] * jp[2] + jp[3] * jp[3];
        a9[4] += jp[2] * jp[4] + jp[3] * jp[5];
        a9[5] += jp[4] * jp[4] + jp[5] * jp[5];
        a9[6] += jp[0] * rx + jp[1] * ry;
        a9[7] += jp[2] * rx + jp[3] * ry;
        a9[8] += jp[4] * rx + jp[5] * ry;
    }
#pragma unroll
    for (int i = 0; i < 9; ++i) {
        double v = a9[i];
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
        a9[i] = 2.0 * v;
    }
    double inv[6];
    bool ok = point_block_inverse(a9, c, inv);
    if (lane == 0) {
        skipped[j] = ok ? 0 : 1;
#pragma unroll
        for (int i = 0; i < 6; ++i) pinv[(int64_t)i * N + j] = ok ? inv[i] : 0.0;
#pragma unroll
        for (int i = 0; i < 3; ++i) pinv[(int64_t)(6 + i) * N + j] = a9[6 + i];
    }
    if (!ok) return;  // non-invertible point block: no Schur contribution (BA.cpp:1877-1881)
    const double gp0 = a9[6], gp1 = a9[7], gp2 = a9[8];
    const double t0 = inv[0] * gp0 + inv[1] * gp1 + inv[2] * gp2;
    const double t1 = inv[1] * gp0 + inv[3] * gp1 + inv[4] * gp2;
    const double t2 = inv[2] * gp0 + inv[4] * gp1 + inv[5] * gp2;

    // ---- phase 2
    const int nch = (k + kSchurCH - 1) / kSchurCH;
    auto stage = [&](int chunk, int slot) {
        int n = min(kSchurCH, k - chunk * kSchurCH);
        if (lane < n) {
            int64_t o = b + (int64_t)chunk * kSchurCH + lane;
            double jp[6], jc[20];
#pragma unroll
            for (int i = 0; i < 6; ++i) jp[i] = J[(int64_t)(2 + i) * O + o];
#pragma unroll
            for (int i = 0; i < 20; ++i) jc[i] = J[(int64_t)(8 + i) * O + o];
            sCam[w][slot][lane] = obs_cam[o];
#pragma unroll
            for (int a = 0; a < 10; ++a) {
                double f0 = 2.0 * (jp[0] * jc[a * 2] + jp[1] * jc[a * 2 + 1]);
                double f1 = 2.0 * (jp[2] * jc[a * 2] + jp[3] * jc[a * 2 + 1]);
                double f2 = 2.0 * (jp[4] * jc[a * 2] + jp[5] * jc[a * 2 + 1]);
                if (slot == 0) { sF[w][lane][a] = f0; sF[w][lane][10 + a] = f1; sF[w][lane][20 + a] = f2; }
                sW[w][slot][lane][a] = inv[0] * f0 + inv[1] * f1 + inv[2] * f2;
                sW[w][slot][lane][10 + a] = inv[1] * f0 + inv[3] * f1 + inv[4] * f2;
                sW[w][slot][lane][20 + a] = inv[2] * f0 + inv[4] * f1 + inv[5] * f2;
            }
        }
        __syncwarp();
        return n;
    };
    for (int ci = 0; ci < nch; ++ci) {
        int nA = stage(ci, 0);
        // rhs += F_i^T (Einv g_p)
        for (int t = lane; t < nA * 10; t += 32) {
            int i = t / 10, a = t % 10;
            sink_add_rhs(sink, sCam[w][0][i], a, sF[w][i][a] * t0 + sF[w][i][10 + a] * t1 + sF[w][i][20 + a] * t2);
        }
        for (int cl = 0; cl <= ci; ++cl) {
            int slotB = 0, nB = nA;
            if (cl != ci) { nB = stage(cl, 1); slotB = 1; }
            int npairs = (cl == ci) ? nA * (nA + 1) / 2 : nA * nB;
            for (int tt = lane; tt < npairs * 4; tt += 32) {
                int pr = tt >> 2, tile = tt & 3;
                int i, l;
                if (cl == ci) {
                    i = (int)((sqrtf(8.0f * (float)pr + 1.0f) - 1.0f) * 0.5f);
                    while (i * (i + 1) / 2 > pr) --i;
                    while ((i + 1) * (i + 2) / 2 <= pr) ++i;
                    l = pr - i * (i + 1) / 2;
                } else { i = pr / nB; l = pr % nB; }
                const int tr = (tile >> 1) * 5, tc = (tile & 1) * 5;
                const double* Fi = sF[w][i];
                const double* Wl = sW[w][slotB][l];
                const int cam_i = sCam[w][0][i], cam_l = sCam[w][slotB][l];
                double f[3][5], ww[3][5];
#pragma unroll
                for (int v = 0; v < 3; ++v)
#pragma unroll
                    for (int x = 0; x < 5; ++x) { f[v][x] = Fi[v * 10 + tr + x]; ww[v][x] = Wl[v * 10 + tc + x]; }
                const int blk = sink_block_id(sink, cam_i, cam_l);
#pragma unroll
                for (int x = 0; x < 5; ++x) {
#pragma unroll
                    for (int y = 0; y < 5; ++y) {
                        double v = f[0][x] * ww[0][y] + f[1][x] * ww[1][y] + f[2][x] * ww[2][y];
                        sink_add(sink, blk, cam_i, tr + x, cam_l, tc + y, -v);
                    }
                }
            }
            __syncwarp();
        }
    }
}


// ---------------------------------------------------------------------------------------------------------------------
// K2 (tiled): the Schur accumulation without per-point global atomics.
//
// A CTA owns a tile of TP consecutive points.  Tracks are created in capture order, so a tile touches few cameras: the
// CTA builds the sorted table of the (at most CMAX) distinct cameras of its tile and keeps the tile's whole contribution
//   S[cam_i, cam_l] -= sum_j F_ji^T E_cj^-1 F_jl        (lower block triangle, cam_i >= cam_l)
// in REGISTERS: "block warps" own 8 local camera pairs each, lane -> (pair, 5x5 sub-tile) with 25 FP64 accumulators, so
// every output entry has exactly one owner and no atomics are needed inside the tile.  Per batch of <= 16 points the
// warps first build E / E^-1 (shuffle reductions, the |det| > 1e-12 rule) and stage F_i = 2 Jp_i^T Jc_i and
// W_i = E^-1 F_i for the batch's observations in shared memory; then every owner lane scans the batch and adds
// F_i^T W_l wherever the point sees both cameras of its pair.  One extra warp owns the tile's rhs entries.  At the end of
// the tile each lane flushes its accumulators with red.global.add.f64: one atomic per touched entry per TILE instead of
// per point.  Points with more than kTileKMax observations, or with a camera outside the tile table, are flagged in
// `deferred` and handled by k_schur (per-point global atomics); the flags depend on the structure only, so the host
// learns at bind time (plan_only) whether that second launch is needed at all.
constexpr int kTileKMax = 16;    // longest track handled in the tile kernel
constexpr int kTileBatch = 16;   // points per staging batch
constexpr int kTileSO = 128;     // observations per staging batch
constexpr int kTileHash = 64;

__device__ __forceinline__ void hash_insert_pair(unsigned long long* keys, unsigned mask, int cam_i, int cam_l, int* overflow) {
    const unsigned long long key = ((unsigned long long)(unsigned)cam_i << 32) | (unsigned)cam_l;
    unsigned h = hash_pair(key, mask);
    for (unsigned probe = 0; probe <= mask; ++probe) {
        const unsigned long long prev = atomicCAS(&keys[h], kHashEmpty, key);
        if (prev == kHashEmpty || prev == key) return;
        h = (h + 1) & mask;
    }
    atomicExch(overflow, 1);
}

template <int CMAX>
struct SchurTileCfg {
    static constexpr int kBlocks = CMAX * (CMAX + 1) / 2;
    static constexpr int kBlockWarps = (kBlocks + 7) / 8;
    static constexpr int kWarps = kBlockWarps + 1;   // + the rhs warp
    static constexpr int kThreads = kWarps * 32;
};

template <int CMAX>
__global__ void __launch_bounds__(SchurTileCfg<CMAX>::kThreads) k_schur_tile(
    int64_t N, int64_t O, int tile_points, const int64_t* __restrict__ pt_begin, const int32_t* __restrict__ obs_cam, const double* __restrict__ J,
    double c, SchurSink sink, double* __restrict__ pinv, unsigned char* __restrict__ skipped, unsigned char* __restrict__ deferred, int plan_only,
    unsigned long long* __restrict__ plan_keys, unsigned plan_mask, int* __restrict__ plan_overflow) {
    using Cfg = SchurTileCfg<CMAX>;
    extern __shared__ double smem[];
    double* sF = smem;                         // [kTileSO][30]
    double* sW = smem + kTileSO * 30;          // [kTileSO][30]
    double* sT = sW + kTileSO * 30;            // [kTileBatch][3]
    int* sSlot = (int*)(sT + kTileBatch * 3);  // [kTileBatch][CMAX]
    int* sTab = sSlot + kTileBatch * CMAX;     // [CMAX] sorted camera ids of the tile
    int* sHash = sTab + CMAX;                  // [kTileHash]
    int* sMisc = sHash + kTileHash;            // [0] nLocal, [1] batch begin, [2] batch end
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int64_t p0 = (int64_t)blockIdx.x * tile_points;
    const int64_t p1 = min(N, p0 + (int64_t)tile_points);
    if (p0 >= N) return;

    // ---- camera table of the tile: hash-set insert of every observation's camera, then sort
    for (int i = tid; i < kTileHash; i += Cfg::kThreads) sHash[i] = -1;
    __syncthreads();
    {
        const int64_t ob = pt_begin[p0], oe = pt_begin[p1];
        for (int64_t o = ob + tid; o < oe; o += Cfg::kThreads) {
            int cam = obs_cam[o];
            unsigned h = ((unsigned)cam * 2654435761u) >> 26;   // 6 bits
            for (int probe = 0; probe < kTileHash; ++probe) {
                int prev = atomicCAS(&sHash[h], -1, cam);
                if (prev == -1 || prev == cam) break;
                h = (h + 1) & (kTileHash - 1);
            }   // a full table simply drops the camera: its points get deferred below
        }
    }
    __syncthreads();
    if (tid == 0) {
        int n = 0;
        for (int i = 0; i < kTileHash; ++i) {
            int cam = sHash[i];
            if (cam < 0) continue;
            // insertion into the sorted table, keeping the CMAX smallest ids
            int pos = n < CMAX ? n : CMAX;
            while (pos > 0 && sTab[pos - 1] > cam) --pos;
            if (pos >= CMAX) continue;
            int last = n < CMAX ? n : CMAX - 1;
            for (int q = last; q > pos; --q) sTab[q] = sTab[q - 1];
            sTab[pos] = cam;
            if (n < CMAX) ++n;
        }
        sMisc[0] = n;
    }
    __syncthreads();
    const int nLocal = sMisc[0];

    // ---- ownership: block warps -> 8 local pairs, lane -> (pair, 5x5 sub-tile)
    const bool is_block_warp = w < Cfg::kBlockWarps;
    int li = -1, ll = -1, tr = 0, tc = 0;
    if (is_block_warp) {
        int b = w * 8 + (lane >> 2);
        if (b < Cfg::kBlocks) {
            int r = (int)((sqrtf(8.0f * (float)b + 1.0f) - 1.0f) * 0.5f);
            while (r * (r + 1) / 2 > b) --r;
            while ((r + 1) * (r + 2) / 2 <= b) ++r;
            li = r; ll = b - r * (r + 1) / 2;
            if (li >= nLocal) { li = -1; ll = -1; }
        }
        tr = ((lane >> 1) & 1) * 5; tc = (lane & 1) * 5;
    }
    double acc[25];
#pragma unroll
    for (int i = 0; i < 25; ++i) acc[i] = 0.0;
    double racc[4] = {0.0, 0.0, 0.0, 0.0};   // rhs warp: entries e = lane + 32*q < CMAX*10

    int64_t bb = p0;
    while (bb < p1) {
        // batch = consecutive points with <= kTileSO observations in total and <= kTileBatch points
        if (tid == 0) {
            int64_t e = bb; int64_t base = pt_begin[bb];
            while (e < p1 && e - bb < kTileBatch && pt_begin[e + 1] - base <= kTileSO) ++e;
            if (e == bb) e = bb + 1;   // a single over-long track: deferred below
            sMisc[1] = (int)(e - bb);
        }
        __syncthreads();
        const int cnt = sMisc[1];
        const int64_t obase = pt_begin[bb];
        // ---- P1: per point E, E^-1, slots, staging
        for (int ptl = w; ptl < cnt; ptl += Cfg::kWarps) {
            const int64_t j = bb + ptl;
            const int64_t b = pt_begin[j], e = pt_begin[j + 1];
            const int k = (int)(e - b);
            if (lane < CMAX) sSlot[ptl * CMAX + lane] = -1;
            int loc = -1;
            bool bad = k > kTileKMax;
            if (!bad && lane < k) {
                int cam = obs_cam[b + lane];
                for (int q = 0; q < nLocal; ++q) if (sTab[q] == cam) loc = q;
                if (loc < 0) bad = true;
            }
            bad = __any_sync(0xffffffffu, bad);
            if (lane == 0) deferred[j] = bad ? 1 : 0;
            if (bad && plan_only && plan_keys != nullptr) {
                for (int pr = lane; pr < k * (k + 1) / 2; pr += 32) {
                    int i2 = (int)((sqrtf(8.0f * (float)pr + 1.0f) - 1.0f) * 0.5f);
                    while (i2 * (i2 + 1) / 2 > pr) --i2;
                    while ((i2 + 1) * (i2 + 2) / 2 <= pr) ++i2;
                    const int l2 = pr - i2 * (i2 + 1) / 2;
                    hash_insert_pair(plan_keys, plan_mask, obs_cam[b + i2], obs_cam[b + l2], plan_overflow);
                }
            }
            if (bad || plan_only) continue;
            double a9[9];
#pragma unroll
            for (int i = 0; i < 9; ++i) a9[i] = 0.0;
            double jp[6] = {0, 0, 0, 0, 0, 0};
            if (lane < k) {
                const int64_t o = b + lane;
                double rx = J[o], ry = J[O + o];
#pragma unroll
                for (int i = 0; i < 6; ++i) jp[i] = J[(int64_t)(2 + i) * O + o];
                a9[0] = jp[0] * jp[0] + jp[1] * jp[1];
                a9[1] = jp[0] * jp[2] + jp[1] * jp[3];
                a9[2] = jp[0] * jp[4] + jp[1] * jp[5];
                a9[3] = jp[2] * jp[2] + jp[3] * jp[3];
                a9[4] = jp[2] * jp[4] + jp[3] * jp[5];
                a9[5] = jp[4] * jp[4] + jp[5] * jp[5];
                a9[6] = jp[0] * rx + jp[1] * ry;
                a9[7] = jp[2] * rx + jp[3] * ry;
                a9[8] = jp[4] * rx + jp[5] * ry;
            }
#pragma unroll
            for (int i = 0; i < 9; ++i) {
                double v = a9[i];
#pragma unroll
                for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
                a9[i] = 2.0 * v;
            }
            double inv[6];
            const bool ok = point_block_inverse(a9, c, inv);
            if (lane == 0) {
                skipped[j] = ok ? 0 : 1;
#pragma unroll
                for (int i = 0; i < 6; ++i) pinv[(int64_t)i * N + j] = ok ? inv[i] : 0.0;
#pragma unroll
                for (int i = 0; i < 3; ++i) pinv[(int64_t)(6 + i) * N + j] = a9[6 + i];
            }
            if (!ok) continue;   // BA.cpp:1877-1881: no Schur contribution
            if (lane == 0) {
                sT[ptl * 3 + 0] = inv[0] * a9[6] + inv[1] * a9[7] + inv[2] * a9[8];
                sT[ptl * 3 + 1] = inv[1] * a9[6] + inv[3] * a9[7] + inv[4] * a9[8];
                sT[ptl * 3 + 2] = inv[2] * a9[6] + inv[4] * a9[7] + inv[5] * a9[8];
            }
            __syncwarp();
            const int sbase = (int)(b - obase);
            if (lane < k) {
                sSlot[ptl * CMAX + loc] = sbase + lane;
                const int64_t o = b + lane;
                double* Fi = sF + (sbase + lane) * 30;
                double* Wi = sW + (sbase + lane) * 30;
#pragma unroll
                for (int a = 0; a < 10; ++a) {
                    double j0 = J[(int64_t)(8 + 2 * a) * O + o], j1 = J[(int64_t)(9 + 2 * a) * O + o];
                    double f0 = 2.0 * (jp[0] * j0 + jp[1] * j1);
                    double f1 = 2.0 * (jp[2] * j0 + jp[3] * j1);
                    double f2 = 2.0 * (jp[4] * j0 + jp[5] * j1);
                    Fi[a] = f0; Fi[10 + a] = f1; Fi[20 + a] = f2;
                    Wi[a] = inv[0] * f0 + inv[1] * f1 + inv[2] * f2;
                    Wi[10 + a] = inv[1] * f0 + inv[3] * f1 + inv[4] * f2;
                    Wi[20 + a] = inv[2] * f0 + inv[4] * f1 + inv[5] * f2;
                }
            }
        }
        __syncthreads();
        // ---- P2: owners scan the batch
        if (!plan_only) {
            if (is_block_warp) {
                if (li >= 0) {
                    for (int ptl = 0; ptl < cnt; ++ptl) {
                        const int si = sSlot[ptl * CMAX + li], sl = sSlot[ptl * CMAX + ll];
                        if (si < 0 || sl < 0) continue;
                        const double* Fi = sF + si * 30 + tr;
                        const double* Wl = sW + sl * 30 + tc;
#pragma unroll
                        for (int v = 0; v < 3; ++v) {
                            double f[5], ww[5];
#pragma unroll
                            for (int x = 0; x < 5; ++x) { f[x] = Fi[v * 10 + x]; ww[x] = Wl[v * 10 + x]; }
#pragma unroll
                            for (int x = 0; x < 5; ++x)
#pragma unroll
                                for (int y = 0; y < 5; ++y) acc[x * 5 + y] += f[x] * ww[y];
                        }
                    }
                }
            } else {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int en = lane + 32 * q;
                    if (en >= CMAX * 10) continue;
                    const int lc = en / 10, a = en % 10;
                    if (lc >= nLocal) continue;
                    for (int ptl = 0; ptl < cnt; ++ptl) {
                        const int si = sSlot[ptl * CMAX + lc];
                        if (si < 0) continue;
                        const double* Fi = sF + si * 30;
                        racc[q] += Fi[a] * sT[ptl * 3 + 0] + Fi[10 + a] * sT[ptl * 3 + 1] + Fi[20 + a] * sT[ptl * 3 + 2];
                    }
                }
            }
        }
        __syncthreads();
        bb += cnt;
    }
    if (plan_only) {
        // structure of the block-sparse reduced camera system: every camera pair of the tile table (a superset of the pairs that
        // co-observe a point of the tile; the extra blocks simply stay zero)
        if (plan_keys != nullptr) {
            for (int b = tid; b < Cfg::kBlocks; b += Cfg::kThreads) {
                int r = (int)((sqrtf(8.0f * (float)b + 1.0f) - 1.0f) * 0.5f);
                while (r * (r + 1) / 2 > b) --r;
                while ((r + 1) * (r + 2) / 2 <= b) ++r;
                const int l2 = b - r * (r + 1) / 2;
                if (r < nLocal) hash_insert_pair(plan_keys, plan_mask, sTab[r], sTab[l2], plan_overflow);
            }
        }
        return;
    }
    // ---- flush: one red.global.add.f64 per touched entry per tile
    if (is_block_warp) {
        if (li >= 0) {
            const int cam_i = sTab[li], cam_l = sTab[ll];
            const int blk = sink_block_id(sink, cam_i, cam_l);
#pragma unroll
            for (int x = 0; x < 5; ++x) {
#pragma unroll
                for (int y = 0; y < 5; ++y) {
                    const double v = acc[x * 5 + y];
                    if (v != 0.0) sink_add(sink, blk, cam_i, tr + x, cam_l, tc + y, -v);
                }
            }
        }
    } else {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int en = lane + 32 * q;
            if (en >= CMAX * 10) continue;
            const int lc = en / 10, a = en % 10;
            if (lc >= nLocal) continue;
            if (racc[q] != 0.0) sink_add_rhs(sink, sTab[lc], a, racc[q]);
        }
    }
}

template <int CMAX>
static size_t schur_tile_smem() {
    return sizeof(double) * (kTileSO * 60 + kTileBatch * 3) + sizeof(int) * (kTileBatch * CMAX + CMAX + kTileHash + 4);
}

// ---------------------------------------------------------------------------------------------------------------------
// K2': LPP lanes per point.  t = sum_i F_i * df[cam_i] + g_p ;  dp = -Einv t ;  X_try = X + dp  (skipped points: dp = 0).
template <int LPP>
__global__ void __launch_bounds__(256) k_backsub(int64_t N, int64_t O, const int64_t* __restrict__ pt_begin, const int32_t* __restrict__ obs_cam,
                                                 const double* __restrict__ J, const double* __restrict__ df /*[10M] with gaps*/,
                                                 const double* __restrict__ pinv, const unsigned char* __restrict__ skipped,
                                                 const double* __restrict__ X, double* __restrict__ Xtry, double* __restrict__ dp_out) {
    int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t j = gid / LPP;
    int sub = (int)(gid % LPP);
    bool active = j < N;
    double t[3] = {0.0, 0.0, 0.0};
    bool skip = true;
    if (active) {
        skip = skipped[j] != 0;
        if (!skip) {
            int64_t b = pt_begin[j], e = pt_begin[j + 1];
            for (int64_t o = b + sub; o < e; o += LPP) {
                const double* d = df + (size_t)obs_cam[o] * 10;
                double s0 = 0.0, s1 = 0.0;
#pragma unroll
                for (int a = 0; a < 10; ++a) {
                    double da = d[a];
                    s0 += J[(int64_t)(8 + a * 2) * O + o] * da;
                    s1 += J[(int64_t)(9 + a * 2) * O + o] * da;
                }
#pragma unroll
                for (int v = 0; v < 3; ++v)
                    t[v] += 2.0 * (J[(int64_t)(2 + v * 2) * O + o] * s0 + J[(int64_t)(3 + v * 2) * O + o] * s1);
            }
        }
    }
#pragma unroll
    for (int v = 0; v < 3; ++v)
#pragma unroll
        for (int s = LPP / 2; s > 0; s >>= 1) t[v] += __shfl_xor_sync(0xffffffffu, t[v], s);
    if (active && sub == 0) {
        double d0 = 0.0, d1 = 0.0, d2 = 0.0;
        if (!skip) {
            double i0 = pinv[j], i1 = pinv[N + j], i2 = pinv[2 * N + j], i3 = pinv[3 * N + j], i4 = pinv[4 * N + j], i5 = pinv[5 * N + j];
            double u0 = t[0] + pinv[6 * N + j], u1 = t[1] + pinv[7 * N + j], u2 = t[2] + pinv[8 * N + j];
            d0 = -(i0 * u0 + i1 * u1 + i2 * u2);
            d1 = -(i1 * u0 + i3 * u1 + i4 * u2);
            d2 = -(i2 * u0 + i4 * u1 + i5 * u2);
        }
        Xtry[j] = X[j] + d0; Xtry[N + j] = X[N + j] + d1; Xtry[2 * N + j] = X[2 * N + j] + d2;
        if (dp_out != nullptr) { dp_out[3 * j] = d0; dp_out[3 * j + 1] = d1; dp_out[3 * j + 2] = d2; }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// K2 for scenes in which a point is seen by MANY cameras (demo-circle-grid: every point in every frame).  There the per-point
// pair enumeration is quadratic in the track length (50 cameras: 1275 pairs x 100 atomics per point) while the whole Schur
// complement is ONE dense contraction over the points,  D = Fall^T Wall,  Fall / Wall = the rows [3N x 10M] of F_j and E_cj^-1 F_j
// (zero where a camera does not see the point): "dense FP64 tensor work where the camera block is a real dense contraction".
// k_schur_rows builds the rows (one warp per point: E, damping, cofactor inverse with the |det| > 1e-12 rule, F_i, W_i; the rhs terms
// F_i^T E^-1 g_p go out as atomics, 10 per observation), launch_gemm_nt_dmma contracts them on DMMA with the K range split over the
// SMs, k_scatter_dense_schur subtracts the lower triangle into the gauge-reduced S.
constexpr int kRowsWarps = 2;
__global__ void __launch_bounds__(kRowsWarps * 32) k_schur_rows(int64_t N, int64_t O, const int64_t* __restrict__ pt_begin, const int32_t* __restrict__ obs_cam,
                                                    const double* __restrict__ J, double c, SchurSink sink, double* __restrict__ pinv, unsigned char* __restrict__ skipped,
                                                    int M, double* __restrict__ Fall, double* __restrict__ Wall) {
    const int lane = threadIdx.x & 31;
    const int64_t j = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (j >= N) return;
    const int64_t b = pt_begin[j], e = pt_begin[j + 1];
    double a9[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) a9[i] = 0.0;
    for (int64_t o = b + lane; o < e; o += 32) {
        const double rx = J[o], ry = J[O + o];
        double jp[6];
#pragma unroll
        for (int i = 0; i < 6; ++i) jp[i] = J[(int64_t)(2 + i) * O + o];
        a9[0] += jp[0] * jp[0] + jp[1] * jp[1];
        a9[1] += jp[0] * jp[2] + jp[1] * jp[3];
        a9[2] += jp[0] * jp[4] + jp[1] * jp[5];
        a9[3] += jp[2] * jp[2] + jp[3] * jp[3];
        a9[4] += jp[2] * jp[4] + jp[3] * jp[5];
        a9[5] += jp[4] * jp[4] + jp[5] * jp[5];
        a9[6] += jp[0] * rx + jp[1] * ry;
        a9[7] += jp[2] * rx + jp[3] * ry;
        a9[8] += jp[4] * rx + jp[5] * ry;
    }
#pragma unroll
    for (int i = 0; i < 9; ++i) {
        double v = a9[i];
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
        a9[i] = 2.0 * v;
    }
    double inv[6];
    const bool ok = point_block_inverse(a9, c, inv);
    if (lane == 0) {
        skipped[j] = ok ? 0 : 1;
#pragma unroll
        for (int i = 0; i < 6; ++i) pinv[(int64_t)i * N + j] = ok ? inv[i] : 0.0;
#pragma unroll
        for (int i = 0; i < 3; ++i) pinv[(int64_t)(6 + i) * N + j] = a9[6 + i];
    }
    if (!ok) return;   // rows stay zero (BA.cpp:1877-1881)
    const double t0 = inv[0] * a9[6] + inv[1] * a9[7] + inv[2] * a9[8];
    const double t1 = inv[1] * a9[6] + inv[3] * a9[7] + inv[4] * a9[8];
    const double t2 = inv[2] * a9[6] + inv[4] * a9[7] + inv[5] * a9[8];
    const int64_t ldr = (int64_t)M * kV;
    // 32 observations at a time: every lane forms the 3 x 10 blocks F_i, W_i of its observation, the warp stages them in shared memory
    // and writes the three rows out as runs of consecutive addresses (cameras of a track are mostly consecutive: 80-byte pieces written
    // lane by lane with a stride of 80 bytes cost 32 sectors per store instruction: 0.64 ms for the 240 MB at configs[1]).
    __shared__ double sF[kRowsWarps][3][320], sW[kRowsWarps][3][320];
    __shared__ int sCam[kRowsWarps][32];
    const int w = threadIdx.x >> 5;
    for (int64_t o0 = b; o0 < e; o0 += 32) {
        const int64_t o = o0 + lane;
        const bool in = o < e;
        const int cam = in ? obs_cam[o] : -1;
        sCam[w][lane] = cam;
        if (in) {
            double jp[6];
#pragma unroll
            for (int i = 0; i < 6; ++i) jp[i] = J[(int64_t)(2 + i) * O + o];
#pragma unroll
            for (int a = 0; a < kV; ++a) {
                const double j0 = J[(int64_t)(8 + 2 * a) * O + o], j1 = J[(int64_t)(9 + 2 * a) * O + o];
                const double f0 = 2.0 * (jp[0] * j0 + jp[1] * j1), f1 = 2.0 * (jp[2] * j0 + jp[3] * j1), f2 = 2.0 * (jp[4] * j0 + jp[5] * j1);
                sF[w][0][lane * kV + a] = f0; sF[w][1][lane * kV + a] = f1; sF[w][2][lane * kV + a] = f2;
                sW[w][0][lane * kV + a] = inv[0] * f0 + inv[1] * f1 + inv[2] * f2;
                sW[w][1][lane * kV + a] = inv[1] * f0 + inv[3] * f1 + inv[4] * f2;
                sW[w][2][lane * kV + a] = inv[2] * f0 + inv[4] * f1 + inv[5] * f2;
                sink_add_rhs(sink, cam, a, f0 * t0 + f1 * t1 + f2 * t2);
            }
        }
        __syncwarp();
#pragma unroll
        for (int t = 0; t < kV; ++t) {
            const int idx = lane + 32 * t;
            const int l2 = idx / kV, a = idx - kV * l2;
            const int cam2 = sCam[w][l2];
            if (cam2 >= 0) {
                const size_t off = (size_t)(3 * j) * ldr + (size_t)cam2 * kV + a;
#pragma unroll
                for (int v = 0; v < 3; ++v) { Fall[off + (size_t)v * ldr] = sF[w][v][idx]; Wall[off + (size_t)v * ldr] = sW[w][v][idx]; }
            }
        }
        __syncwarp();
    }
}
// S(red(r), red(c)) += D(r, c), D = -(Fall^T Wall) as the GEMM leaves it (C -= A B^T on a zeroed C), for the lower triangle of the full
// frame-variable space; the 7 gauge variables are skipped
__global__ void k_scatter_dense_schur(int n_full, const double* __restrict__ D, int unity, double* __restrict__ S, int64_t ld) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x, cidx = blockIdx.y;
    if (r >= n_full || r < cidx) return;
    const int rr = red_index(r / kV, r % kV, unity), rc = red_index(cidx / kV, cidx % kV, unity);
    if (rr < 0 || rc < 0) return;
    const double v = D[(size_t)cidx * n_full + r];
    if (v != 0.0) S[(size_t)rc * ld + rr] += v;
}

// K2' with one thread per OBSERVATION (every lane of a warp load carries data: the point-per-half-warp form above leaves 6 of 16 lanes
// idle at 10 observations per point and was bound by load latency at 58 % of the HBM peak).  A lane forms its observation's
// contribution F_i * df[cam_i]; the lanes of one point are contiguous (observations are point-major), so a segmented shuffle
// reduction gives the point's partial sum, which the segment's first lane adds to tacc[3][N].  A point spans at most two warps for tracks
// of up to 32 observations, and two partial sums add commutatively onto an exact zero: the result is bit-reproducible.
// k_backsub_finish then applies -Einv (t + g_p), writes the trial points and re-zeroes tacc for the next attempt.
__global__ void __launch_bounds__(256) k_backsub_obs(int64_t O, int64_t N, const int32_t* __restrict__ obs_pt, const int32_t* __restrict__ obs_cam,
                                                     const double* __restrict__ J, const double* __restrict__ df, const unsigned char* __restrict__ skipped,
                                                     double* __restrict__ tacc) {
    const int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const bool in = o < O;
    const int pt = in ? obs_pt[o] : -1;
    double c0 = 0.0, c1 = 0.0, c2 = 0.0;
    bool live = false;
    if (in) {
        const int cam = obs_cam[o];
        live = skipped[pt] == 0;
        if (live) {
            const double* d = df + (size_t)cam * 10;
            double s0 = 0.0, s1 = 0.0;
#pragma unroll
            for (int a = 0; a < 10; ++a) {
                const double da = d[a];
                s0 += J[(int64_t)(8 + a * 2) * O + o] * da;
                s1 += J[(int64_t)(9 + a * 2) * O + o] * da;
            }
            c0 = 2.0 * (J[(int64_t)2 * O + o] * s0 + J[(int64_t)3 * O + o] * s1);
            c1 = 2.0 * (J[(int64_t)4 * O + o] * s0 + J[(int64_t)5 * O + o] * s1);
            c2 = 2.0 * (J[(int64_t)6 * O + o] * s0 + J[(int64_t)7 * O + o] * s1);
        }
    }
    // segment heads: first lane of the warp or a new point id
    const int prev = __shfl_up_sync(0xffffffffu, pt, 1);
    const bool head = lane == 0 || pt != prev;
    const unsigned heads = __ballot_sync(0xffffffffu, head);
    const unsigned later = heads & ~((2u << lane) - 1u);            // heads strictly after this lane
    const int seg_end = later ? __ffs(later) - 1 : 32;              // first lane of the next segment
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const double v0 = __shfl_down_sync(0xffffffffu, c0, off), v1 = __shfl_down_sync(0xffffffffu, c1, off), v2 = __shfl_down_sync(0xffffffffu, c2, off);
        if (lane + off < seg_end) { c0 += v0; c1 += v1; c2 += v2; }
    }
    if (head && live) { atomicAdd(&tacc[pt], c0); atomicAdd(&tacc[N + pt], c1); atomicAdd(&tacc[2 * N + pt], c2); }
}
__global__ void __launch_bounds__(256) k_backsub_finish(int64_t N, double* __restrict__ tacc, const double* __restrict__ pinv, const unsigned char* __restrict__ skipped,
                                                        const double* __restrict__ X, double* __restrict__ Xtry, double* __restrict__ dp_out) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= N) return;
    double d0 = 0.0, d1 = 0.0, d2 = 0.0;
    if (skipped[j] == 0) {
        const double i0 = pinv[j], i1 = pinv[N + j], i2 = pinv[2 * N + j], i3 = pinv[3 * N + j], i4 = pinv[4 * N + j], i5 = pinv[5 * N + j];
        const double u0 = tacc[j] + pinv[6 * N + j], u1 = tacc[N + j] + pinv[7 * N + j], u2 = tacc[2 * N + j] + pinv[8 * N + j];
        d0 = -(i0 * u0 + i1 * u1 + i2 * u2);
        d1 = -(i1 * u0 + i3 * u1 + i4 * u2);
        d2 = -(i2 * u0 + i4 * u1 + i5 * u2);
        tacc[j] = 0.0; tacc[N + j] = 0.0; tacc[2 * N + j] = 0.0;
    }
    Xtry[j] = X[j] + d0; Xtry[N + j] = X[N + j] + d1; Xtry[2 * N + j] = X[2 * N + j] + d2;
    if (dp_out != nullptr) { dp_out[3 * j] = d0; dp_out[3 * j + 1] = d1; dp_out[3 * j + 2] = d2; }
}

// ---------------------------------------------------------------------------------------------------------------------
// K4: per-camera update (ApplyCorrections, BA.cpp:2021-2062): direct = SE3Inv(inverse); T_d += dT;
// R_d <- Rodrigues(dW)*R_d unless |dW| is ~0 (IsClose(0, ang), quirk Q6); inverse = SE3Inv(direct).
__global__ void k_cam_update(int M, const double* __restrict__ cams, const double* __restrict__ df, double* __restrict__ cams_try) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M) return;
    const double* c = cams + (size_t)i * 12;
    const double* d = df + (size_t)i * 10;
    double T[3] = {c[0], c[1], c[2]};
    double R[9];
#pragma unroll
    for (int j = 0; j < 9; ++j) R[j] = c[3 + j];
    // direct pose
    double Rd[9], Td[3];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) Rd[cc * 3 + r] = R[r * 3 + cc];
#pragma unroll
    for (int r = 0; r < 3; ++r) Td[r] = -(Rd[0 * 3 + r] * T[0] + Rd[1 * 3 + r] * T[1] + Rd[2 * 3 + r] * T[2]);
    Td[0] += d[4]; Td[1] += d[5]; Td[2] += d[6];
    double w0 = d[7], w1 = d[8], w2 = d[9];
    double ang = sqrt(w0 * w0 + w1 * w1 + w2 * w2);
    // IsClose(0, ang): |0 - ang| <= 1e-8 + 1e-5*|max(0, ang)|   (approx-alg.h:7-16)
    bool skip_rot = fabs(ang) <= (1e-8 + 1e-5 * fabs(fmax(0.0, ang)));
    double Rn[9];
    if (!skip_rot) {
        double n0 = w0 / ang, n1 = w1 / ang, n2 = w2 / ang;
        double s = sin(ang), co = cos(ang);
        double Sk[9] = {0.0, n2, -n1, -n2, 0.0, n0, n1, -n0, 0.0};  // column-major skew
        double A[9];                                                  // (1-c)*Sk
#pragma unroll
        for (int j = 0; j < 9; ++j) A[j] = (1.0 - co) * Sk[j];
        double rot[9];
#pragma unroll
        for (int cc = 0; cc < 3; ++cc)
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                double a2 = A[0 * 3 + r] * Sk[cc * 3 + 0] + A[1 * 3 + r] * Sk[cc * 3 + 1] + A[2 * 3 + r] * Sk[cc * 3 + 2];
                rot[cc * 3 + r] = ((r == cc ? 1.0 : 0.0) + s * Sk[cc * 3 + r]) + a2;
            }
#pragma unroll
        for (int cc = 0; cc < 3; ++cc)
#pragma unroll
            for (int r = 0; r < 3; ++r) Rn[cc * 3 + r] = rot[0 * 3 + r] * Rd[cc * 3 + 0] + rot[1 * 3 + r] * Rd[cc * 3 + 1] + rot[2 * 3 + r] * Rd[cc * 3 + 2];
    } else {
#pragma unroll
        for (int j = 0; j < 9; ++j) Rn[j] = Rd[j];
    }
    // back to inverse pose
    double* o = cams_try + (size_t)i * 12;
    double Ri[9];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) Ri[cc * 3 + r] = Rn[r * 3 + cc];
#pragma unroll
    for (int r = 0; r < 3; ++r) o[r] = -(Ri[0 * 3 + r] * Td[0] + Ri[1 * 3 + r] * Td[1] + Ri[2 * 3 + r] * Td[2]);
#pragma unroll
    for (int j = 0; j < 9; ++j) o[3 + j] = Ri[j];
}

// df (with gaps, [10M]) from the reduced solution (FillCorrectionsGapsFromNormalized, BA.cpp:1600-1679).
__global__ void k_expand_df(int M, const double* __restrict__ dfr, int unity, double* __restrict__ df) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= M * 10) return;
    int r = red_index(t / 10, t % 10, unity);
    df[t] = r >= 0 ? dfr[r] : 0.0;
}

// Points: X' = (R0*X + T0)*s (BA.cpp:179-186) or the inverse X = R0^T (X'*(1/s) - T0) (BA.cpp:187-191).
__global__ void k_normalize_points(int64_t N, double* __restrict__ X, const double* __restrict__ cam0 /*T[3],R[9]*/, double s, int revert) {
    int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= N) return;
    const double* T = cam0; const double* R = cam0 + 3;
    double x0 = X[j], x1 = X[N + j], x2 = X[2 * N + j];
    if (!revert) {
        double y0 = R[0] * x0 + R[3] * x1 + R[6] * x2 + T[0];
        double y1 = R[1] * x0 + R[4] * x1 + R[7] * x2 + T[1];
        double y2 = R[2] * x0 + R[5] * x1 + R[8] * x2 + T[2];
        X[j] = y0 * s; X[N + j] = y1 * s; X[2 * N + j] = y2 * s;
    } else {
        double is = 1.0 / s;
        double t0 = x0 * is - T[0], t1 = x1 * is - T[1], t2 = x2 * is - T[2];
        X[j] = R[0] * t0 + R[1] * t1 + R[2] * t2;
        X[N + j] = R[3] * t0 + R[4] * t1 + R[5] * t2;
        X[2 * N + j] = R[6] * t0 + R[7] * t1 + R[8] * t2;
    }
}

// AoS [3N] <-> planes [3][N]
__global__ void k_points_to_planes(int64_t N, const double* __restrict__ aos, double* __restrict__ planes) {
    int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= N) return;
    planes[j] = aos[3 * j]; planes[N + j] = aos[3 * j + 1]; planes[2 * N + j] = aos[3 * j + 2];
}
__global__ void k_planes_to_points(int64_t N, const double* __restrict__ planes, double* __restrict__ aos) {
    int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= N) return;
    aos[3 * j] = planes[j]; aos[3 * j + 1] = planes[N + j]; aos[3 * j + 2] = planes[2 * N + j];
}

// Parity hooks: E (from the stored point data is damped/inverted, so recompute the raw blocks) and F blocks from J.
__global__ void k_debug_point_blocks(int64_t N, int64_t O, const int64_t* __restrict__ pt_begin, const double* __restrict__ J,
                                     double* __restrict__ E /*[9N]*/, double* __restrict__ gp /*[3N]*/) {
    int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= N) return;
    double a[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int64_t o = pt_begin[j]; o < pt_begin[j + 1]; ++o) {
        double rx = J[o], ry = J[O + o], jp[6];
        for (int i = 0; i < 6; ++i) jp[i] = J[(int64_t)(2 + i) * O + o];
        a[0] += jp[0] * jp[0] + jp[1] * jp[1]; a[1] += jp[0] * jp[2] + jp[1] * jp[3]; a[2] += jp[0] * jp[4] + jp[1] * jp[5];
        a[3] += jp[2] * jp[2] + jp[3] * jp[3]; a[4] += jp[2] * jp[4] + jp[3] * jp[5]; a[5] += jp[4] * jp[4] + jp[5] * jp[5];
        a[6] += jp[0] * rx + jp[1] * ry; a[7] += jp[2] * rx + jp[3] * ry; a[8] += jp[4] * rx + jp[5] * ry;
    }
    double* e = E + 9 * j;
    e[0] = 2 * a[0]; e[1] = 2 * a[1]; e[2] = 2 * a[2]; e[3] = 2 * a[1]; e[4] = 2 * a[3]; e[5] = 2 * a[4]; e[6] = 2 * a[2]; e[7] = 2 * a[4]; e[8] = 2 * a[5];
    gp[3 * j] = 2 * a[6]; gp[3 * j + 1] = 2 * a[7]; gp[3 * j + 2] = 2 * a[8];
}
__global__ void k_debug_F_blocks(int64_t O, const double* __restrict__ J, double* __restrict__ F /*[30*O]*/) {
    int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= O) return;
    for (int v = 0; v < 3; ++v)
        for (int a = 0; a < 10; ++a)
            F[o * 30 + v * 10 + a] = 2.0 * (J[(int64_t)(2 + v * 2) * O + o] * J[(int64_t)(8 + a * 2) * O + o] + J[(int64_t)(3 + v * 2) * O + o] * J[(int64_t)(9 + a * 2) * O + o]);
}

// ---------------------------------------------------------------------------------------------------------------------
// Launch wrappers

static inline unsigned cdiv(int64_t a, int64_t b) { return (unsigned)((a + b - 1) / b); }

void launch_cam_prep(cudaStream_t st, int M, const double* cams, const double* K, int shared_K, double f0, double* camd) {
    k_cam_prep<<<cdiv(M, 128), 128, 0, st>>>(M, cams, K, shared_K, f0, camd);
}
void launch_jacobian(cudaStream_t st, int64_t O, const int32_t* obs_cam, const int32_t* obs_pt, const double* x, const double* y,
                     const double* X, int64_t N, const double* camd, double* J, double* Eacc) {
    if (O > 0) k_jacobian<<<cdiv(O, 256 * kObsRows), 256, 0, st>>>(O, obs_cam, obs_pt, x, y, X, N, camd, J, Eacc);
}
void launch_frame_blocks(cudaStream_t st, int M, const int64_t* cam_begin, const int32_t* c_pt, const double* c_x, const double* c_y,
                         const double* X, int64_t N, const double* camd, double* G, double* gf, int splits) {
    if (splits > 1) { cudaMemsetAsync(G, 0, sizeof(double) * 100 * (size_t)M, st); cudaMemsetAsync(gf, 0, sizeof(double) * 10 * (size_t)M, st); }
    k_frame_blocks<<<dim3(M, splits), 128, 0, st>>>(M, cam_begin, c_pt, c_x, c_y, X, N, camd, G, gf, splits);
}
int64_t residual_chunks(int64_t O) { return (O + kResChunk - 1) / kResChunk; }
int residual_chunk_slots() { return kCamTabSlots; }
void launch_chunk_tables(cudaStream_t st, int64_t O, const int32_t* obs_cam, const int32_t* obs_pt, int* chunk_cams, int* chunk_cnt, int* chunk_pts, unsigned char* obs_slot) {
    if (O > 0) k_chunk_tables<<<(unsigned)residual_chunks(O), 256, 0, st>>>(O, obs_cam, obs_pt, chunk_cams, chunk_cnt, reinterpret_cast<int2*>(chunk_pts), obs_slot);
}
void launch_schur_rows(cudaStream_t st, int64_t N, int64_t O, const int64_t* pt_begin, const int32_t* obs_cam, const double* J, double c, const SchurSink& sink,
                       double* pinv, unsigned char* skipped, int M, double* Fall, double* Wall) {
    if (N > 0) k_schur_rows<<<cdiv(N, kRowsWarps), kRowsWarps * 32, 0, st>>>(N, O, pt_begin, obs_cam, J, c, sink, pinv, skipped, M, Fall, Wall);
}
void launch_scatter_dense_schur(cudaStream_t st, int n_full, const double* D, int unity, double* S, int64_t ld) {
    if (n_full > 0) k_scatter_dense_schur<<<dim3(cdiv(n_full, 128), n_full), 128, 0, st>>>(n_full, D, unity, S, ld);
}
void launch_backsub_obs(cudaStream_t st, int64_t N, int64_t O, const int32_t* obs_pt, const int32_t* obs_cam, const double* J, const double* df, const double* pinv,
                        const unsigned char* skipped, const double* X, double* Xtry, double* dp_out, double* tacc) {
    if (O > 0) k_backsub_obs<<<cdiv(O, 256), 256, 0, st>>>(O, N, obs_pt, obs_cam, J, df, skipped, tacc);
    if (N > 0) k_backsub_finish<<<cdiv(N, 256), 256, 0, st>>>(N, tacc, pinv, skipped, X, Xtry, dp_out);
}
void launch_residual(cudaStream_t st, int64_t O, const int32_t* obs_cam, const int32_t* obs_pt, const double* x, const double* y,
                     const double* X, int64_t N, const double* camd, const int* chunk_cams, const int* chunk_cnt, const int* chunk_pts,
                     const unsigned char* obs_slot, double* partial, int nblocks, double* out) {
    k_residual<<<nblocks, 256, 0, st>>>(O, obs_cam, obs_pt, x, y, X, N, camd, chunk_cams, chunk_cnt, reinterpret_cast<const int2*>(chunk_pts), obs_slot, partial);
    k_sum_partials<<<1, 256, 0, st>>>(nblocks, partial, out);
}
void launch_fill_reduced(cudaStream_t st, int M, const double* G, const double* gf, double c, int unity, double* S, int64_t ld, double* rhs) {
    k_fill_reduced<<<M, 128, 0, st>>>(M, G, gf, c, unity, S, ld, rhs);
}
void launch_schur(cudaStream_t st, int64_t N, int64_t O, const int64_t* pt_begin, const int32_t* obs_cam, const double* J, double c, const SchurSink& sink,
                  double* pinv, unsigned char* skipped, const unsigned char* only_flagged) {
    if (N > 0) k_schur<<<cdiv(N, kSchurWarps), kSchurWarps * 32, 0, st>>>(N, O, pt_begin, obs_cam, J, c, sink, pinv, skipped, only_flagged, nullptr, nullptr);
}
void launch_schur_list(cudaStream_t st, int64_t N, int64_t O, const int64_t* pt_begin, const int32_t* obs_cam, const double* J, double c, const SchurSink& sink,
                       double* pinv, unsigned char* skipped, const int* list, const int* list_count, int list_cap) {
    if (N > 0 && list_cap > 0) k_schur<<<cdiv(list_cap, kSchurWarps), kSchurWarps * 32, 0, st>>>(N, O, pt_begin, obs_cam, J, c, sink, pinv, skipped, nullptr, list, list_count);
}
void launch_schur_tile(cudaStream_t st, int64_t N, int64_t O, int tile_points, const int64_t* pt_begin, const int32_t* obs_cam, const double* J, double c,
                       const SchurSink& sink, double* pinv, unsigned char* skipped, unsigned char* deferred, int plan_only,
                       unsigned long long* plan_keys, unsigned plan_mask, int* plan_overflow) {
    if (N <= 0) return;
    constexpr int CMAX = 12;
    static PerDeviceOnce once;
    const size_t smem = schur_tile_smem<CMAX>();
    if (once.first()) cudaFuncSetAttribute(k_schur_tile<CMAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k_schur_tile<CMAX><<<cdiv(N, tile_points), SchurTileCfg<CMAX>::kThreads, smem, st>>>(N, O, tile_points, pt_begin, obs_cam, J, c, sink, pinv, skipped,
                                                                                      deferred, plan_only, plan_keys, plan_mask, plan_overflow);
}
void launch_backsub(cudaStream_t st, int64_t N, int64_t O, const int64_t* pt_begin, const int32_t* obs_cam, const double* J, const double* df,
                    const double* pinv, const unsigned char* skipped, const double* X, double* Xtry, double* dp_out, int lanes_per_point) {
    if (N <= 0) return;
    if (lanes_per_point <= 4) k_backsub<4><<<cdiv(N * 4, 256), 256, 0, st>>>(N, O, pt_begin, obs_cam, J, df, pinv, skipped, X, Xtry, dp_out);
    else if (lanes_per_point <= 8) k_backsub<8><<<cdiv(N * 8, 256), 256, 0, st>>>(N, O, pt_begin, obs_cam, J, df, pinv, skipped, X, Xtry, dp_out);
    else if (lanes_per_point <= 16) k_backsub<16><<<cdiv(N * 16, 256), 256, 0, st>>>(N, O, pt_begin, obs_cam, J, df, pinv, skipped, X, Xtry, dp_out);
    else k_backsub<32><<<cdiv(N * 32, 256), 256, 0, st>>>(N, O, pt_begin, obs_cam, J, df, pinv, skipped, X, Xtry, dp_out);
}
void launch_cam_update(cudaStream_t st, int M, const double* cams, const double* df, double* cams_try) {
    k_cam_update<<<cdiv(M, 128), 128, 0, st>>>(M, cams, df, cams_try);
}
void launch_expand_df(cudaStream_t st, int M, const double* dfr, int unity, double* df) {
    k_expand_df<<<cdiv((int64_t)M * 10, 256), 256, 0, st>>>(M, dfr, unity, df);
}
void launch_normalize_points(cudaStream_t st, int64_t N, double* X, const double* cam0_dev, double s, int revert) {
    if (N > 0) k_normalize_points<<<cdiv(N, 256), 256, 0, st>>>(N, X, cam0_dev, s, revert);
}
void launch_points_to_planes(cudaStream_t st, int64_t N, const double* aos, double* planes) {
    if (N > 0) k_points_to_planes<<<cdiv(N, 256), 256, 0, st>>>(N, aos, planes);
}
void launch_planes_to_points(cudaStream_t st, int64_t N, const double* planes, double* aos) {
    if (N > 0) k_planes_to_points<<<cdiv(N, 256), 256, 0, st>>>(N, planes, aos);
}
void launch_debug_point_blocks(cudaStream_t st, int64_t N, int64_t O, const int64_t* pt_begin, const double* J, double* E, double* gp) {
    if (N > 0) k_debug_point_blocks<<<cdiv(N, 128), 128, 0, st>>>(N, O, pt_begin, J, E, gp);
}
void launch_debug_F_blocks(cudaStream_t st, int64_t O, const double* J, double* F) {
    if (O > 0) k_debug_F_blocks<<<cdiv(O, 128), 128, 0, st>>>(O, J, F);
}

}  // namespace srk
