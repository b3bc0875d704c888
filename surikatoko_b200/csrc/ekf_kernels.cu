// suriko-b200 — MonoSLAM EKF dense covariance chain on sm_100a behind include/srk/ekf_c_api.h.
//
// Replaces the Eigen algebra of ProcessFrame_StackedObservationsPerUpdateCore (EKF.cpp:977-1125) and PredictEstimVars
// (EKF.cpp:669-693), EKF.cpp = /root/reference/cpp_impl/suriko-engine/src/davison-mono-slam.cpp.  Device algorithm:
//   k_ekf_pht      PHt = P * H^T with the SPARSE H (13 camera columns + s point columns per observation row)        [n x 2m]
//   k_ekf_innov    S = H * PHt + meas_var * I, lower triangle                                                       [2m x 2m]
//   chol_kernels   S = L L^T (DMMA blocked Cholesky, inverted 64x64 diagonal blocks)
//   TRSM           Z = PHt * L^-T, block column by block column: 64-wide right solve + DMMA update of the columns to the right
//   k_ekf_state    x += Z * (L^-1 (z - h))
//   k_gemm_nt_dmma P -= Z * Z^T on the lower triangle (n x n x 2m DMMA), then mirrored
//   k_ekf_quat_*   quaternion renormalisation of x and of the 4 rows / columns of P (EKF.cpp:1652-1711)
//   k_ekf_nonneg_* rows / columns with a negative variance are zeroed (EKF.cpp:1739-1750)
//   k_ekf_predict* Pvv <- F Pvv F^T + G Q G^T, Pvm <- F Pvm, Pmv <- Pvm^T
// K = PHt S^-1 is never formed: K S K^T = Z Z^T and K (z - h) = Z L^-1 (z - h).
#include <cmath>
#include <cstring>
#include <string>
#include <cstdlib>
#include <vector>

#include "../../include/srk/ekf_c_api.h"
#include "kernels.h"

extern "C" void srk_internal_set_error(const char* s);   // engine.cu: the string behind srk_last_error()

namespace {

constexpr int kCam = 13;   // camera-state components (EKF.h kCamStateComps)

// PHt(r, j) = sum_c P(r, c) Hcam(j, c) + sum_c P(r, off_j + c) Hpt(j, c).  Threads along r (coalesced columns of P), 8 observation
// rows per CTA share the 13 camera columns through registers.
template <int S>
__global__ void __launch_bounds__(256) k_ekf_pht(int n, int m2, const double* __restrict__ P, const double* __restrict__ Hcam, const double* __restrict__ Hpt,
                                                 const int64_t* __restrict__ pt_off, double* __restrict__ PHt, int64_t ldz) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    const int j0 = blockIdx.y * 8;
    if (r >= n) return;
    double pc[kCam];
#pragma unroll
    for (int c = 0; c < kCam; ++c) pc[c] = P[(size_t)c * n + r];
    for (int j = j0; j < j0 + 8 && j < m2; ++j) {
        const double* hc = Hcam + (size_t)j * kCam;
        const double* hp = Hpt + (size_t)j * S;
        const int64_t off = pt_off[j >> 1];
        double acc = 0.0;
#pragma unroll
        for (int c = 0; c < kCam; ++c) acc += pc[c] * hc[c];
#pragma unroll
        for (int c = 0; c < S; ++c) acc += P[(size_t)(off + c) * n + r] * hp[c];
        PHt[(size_t)j * ldz + r] = acc;
    }
}
// S(i, j) = sum_c Hcam(i, c) PHt(c, j) + sum_c Hpt(i, c) PHt(off_i + c, j) + meas_var [i == j], lower triangle (i >= j), ld = lds
template <int S>
__global__ void __launch_bounds__(256) k_ekf_innov(int n, int m2, const double* __restrict__ PHt, const double* __restrict__ Hcam, const double* __restrict__ Hpt,
                                                   const int64_t* __restrict__ pt_off, double meas_var, double* __restrict__ Sm, int64_t lds, int64_t ldz) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y;
    if (i >= m2 || i < j) return;
    const double* col = PHt + (size_t)j * ldz;
    const double* hc = Hcam + (size_t)i * kCam;
    const double* hp = Hpt + (size_t)i * S;
    const int64_t off = pt_off[i >> 1];
    double acc = 0.0;
#pragma unroll
    for (int c = 0; c < kCam; ++c) acc += hc[c] * col[c];
#pragma unroll
    for (int c = 0; c < S; ++c) acc += hp[c] * col[off + c];
    if (i == j) acc += meas_var;
    Sm[(size_t)j * lds + i] = acc;
}
// w = z - h, written as ROW n of P H^T (element i at w[i * ldw]): the gain TRSM Z = [P H^T; w^T] L^-T then leaves L^-1 w in that row, and the
// forward substitution of the state update (a latency chain of 2m / 64 blocks, 0.35 ms at 2m = 4000) is not needed
__global__ void k_ekf_innovation_vec(int m2, const double* __restrict__ z, const double* __restrict__ h, double* __restrict__ w, int64_t ldw) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < m2) w[(size_t)i * ldw] = z[i] - h[i];
}
__global__ void k_ekf_take_row(int m2, const double* __restrict__ row, int64_t ld, double* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < m2) out[i] = row[(size_t)i * ld];
}
// x += Z w  (Z: n x m2 column-major); one warp per 32 rows, columns strided over the CTA's warps, fixed combination order
__global__ void __launch_bounds__(256) k_ekf_state(int n, int m2, const double* __restrict__ Z, int64_t ldz, const double* __restrict__ w, double* __restrict__ x) {
    __shared__ double part[8][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int r = blockIdx.x * 32 + lane;
    double acc = 0.0;
    if (r < n) for (int j = warp; j < m2; j += 8) acc += Z[(size_t)j * ldz + r] * w[j];
    part[warp][lane] = acc;
    __syncthreads();
    if (warp == 0 && r < n) {
        double s = 0.0;
        for (int k = 0; k < 8; ++k) s += part[k][lane];
        x[r] += s;
    }
}

// Quaternion renormalisation (EKF.cpp:1652-1711).  Step 1 (one CTA): decide (IsClose(1, |q|)), normalise x[3..7), build dq (4x4) into
// aux[0..16), flag into aux[16]; step 2: new column block  Cnew(r, j) = sum_k P(r, 3+k) dq(j, k)  for r outside 3..6 and the centre
// dq P44 dq^T, into tmp [n x 4]; step 3: write the four columns and rows back.
__global__ void k_ekf_quat_prepare(double* __restrict__ x, double* __restrict__ aux) {
    if (threadIdx.x != 0) return;
    const double q[4] = {x[3], x[4], x[5], x[6]};
    const double n2 = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
    const double len = sqrt(n2);
    const bool close = fabs(1.0 - len) <= (1.0e-8 + 1.0e-5 * fabs(fmax(1.0, len)));   // approx-alg.h:7-16
    aux[16] = close ? 0.0 : 1.0;
    if (close) return;
    for (int i = 0; i < 4; ++i) x[3 + i] = q[i] / len;
    const double mult = pow(n2, (double)-1.5f);
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            double v;
            if (i == j) { v = 0.0; for (int k = 0; k < 4; ++k) if (k != i) v += q[k] * q[k]; }
            else v = -q[i] * q[j];
            aux[i * 4 + j] = v * mult;
        }
}
__global__ void k_ekf_quat_columns(int n, const double* __restrict__ P, const double* __restrict__ aux, double* __restrict__ tmp) {
    if (aux[16] == 0.0) return;
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    double p[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) p[k] = P[(size_t)(3 + k) * n + r];
    if (r >= 3 && r < 7) {   // centre: dq * P44 * dq^T, row r-3
        double t[4];
        for (int c = 0; c < 4; ++c) { double s = 0.0; for (int k = 0; k < 4; ++k) s += aux[(r - 3) * 4 + k] * P[(size_t)(3 + c) * n + 3 + k]; t[c] = s; }
        for (int j = 0; j < 4; ++j) { double s = 0.0; for (int k = 0; k < 4; ++k) s += t[k] * aux[j * 4 + k]; tmp[(size_t)j * n + r] = s; }
    } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) tmp[(size_t)j * n + r] = (p[0] * aux[j * 4 + 0] + p[1] * aux[j * 4 + 1]) + (p[2] * aux[j * 4 + 2] + p[3] * aux[j * 4 + 3]);
    }
}
__global__ void k_ekf_quat_write(int n, double* __restrict__ P, const double* __restrict__ aux, const double* __restrict__ tmp) {
    if (aux[16] == 0.0) return;
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        double v = tmp[(size_t)j * n + r];
        if (r >= 3 && r < 7) v = 0.5 * (v + tmp[(size_t)(r - 3) * n + 3 + j]);   // centre block: (M + M^T)/2 as FixSymmetricMat does (EKF.cpp:4308)
        P[(size_t)(3 + j) * n + r] = v;           // column 3+j
        if (r < 3 || r >= 7) P[(size_t)r * n + 3 + j] = v;   // row 3+j (mirror); the centre block is written once, as columns
    }
}
// EnsureNonnegativeStateVariance (EKF.cpp:1739-1750)
__global__ void k_ekf_nonneg_flag(int n, const double* __restrict__ P, unsigned char* __restrict__ neg, int* __restrict__ any) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const bool bad = !(P[(size_t)i * n + i] >= 0.0);
    neg[i] = bad ? 1 : 0;
    if (bad) atomicOr(any, 1);
}
__global__ void k_ekf_nonneg_zero(int n, double* __restrict__ P, const unsigned char* __restrict__ neg, const int* __restrict__ any) {
    if (*any == 0) return;
    const int c = blockIdx.y;
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    if (neg[c] || neg[r]) P[(size_t)c * n + r] = 0.0;
}
// Predict: one CTA computes Pvv_new (13x13) into aux; a grid computes Pvm_new = F * Pvm column by column and mirrors it.
__global__ void k_ekf_predict_vv(int n, const double* __restrict__ P, const double* __restrict__ F, const double* __restrict__ Q, double* __restrict__ out) {
    __shared__ double FP[kCam][kCam];
    __shared__ double R2[kCam][kCam];
    const int t = threadIdx.x;
    if (t < kCam * kCam) {
        const int i = t % kCam, j = t / kCam;
        double s = 0.0;
        for (int k = 0; k < kCam; ++k) s += F[(size_t)k * kCam + i] * P[(size_t)j * n + k];   // (F Pvv)(i, j)
        FP[i][j] = s;
    }
    __syncthreads();
    if (t < kCam * kCam) {
        const int i = t % kCam, j = t / kCam;
        double s = 0.0;
        for (int k = 0; k < kCam; ++k) s += FP[i][k] * F[(size_t)k * kCam + j];               // (F Pvv F^T)(i, j) = sum_k FP(i,k) F(j,k)
        R2[i][j] = s + Q[(size_t)j * kCam + i];
    }
    __syncthreads();
    if (t < kCam * kCam) {   // exactly symmetric result, as FixSymmetricMat leaves it (EKF.cpp:692-693)
        const int i = t % kCam, j = t / kCam;
        out[(size_t)j * kCam + i] = 0.5 * (R2[i][j] + R2[j][i]);
    }
}
__global__ void k_ekf_predict_vm(int n, double* __restrict__ P, const double* __restrict__ F, const double* __restrict__ vv_new, const double* __restrict__ cam_new,
                                 double* __restrict__ x) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;   // column of P
    if (c >= n) return;
    if (c < kCam) {
        for (int i = 0; i < kCam; ++i) P[(size_t)c * n + i] = vv_new[(size_t)c * kCam + i];
        if (cam_new != nullptr) x[c] = cam_new[c];
        return;
    }
    double col[kCam], res[kCam];
#pragma unroll
    for (int k = 0; k < kCam; ++k) col[k] = P[(size_t)c * n + k];
#pragma unroll
    for (int i = 0; i < kCam; ++i) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < kCam; ++k) s += F[(size_t)k * kCam + i] * col[k];
        res[i] = s;
    }
#pragma unroll
    for (int i = 0; i < kCam; ++i) { P[(size_t)c * n + i] = res[i]; P[(size_t)i * n + c] = res[i]; }
}


// ---------------------------------------------------------------------------------------------------------------------
// 1-point RANSAC hypothesis scoring (OnePointRansac_GetConsensusMatches, EKF.cpp:1271-1391), all m hypotheses at once.
// Hypothesis i integrates ONLY matched point i into the state: S_i (2x2) = Hx Pxx Hx^T + mid + mid^T + Hy Pyy Hy^T + R (:1321-1326),
// x_i = x + (P[:, cam] Hx^T + P[:, pt_i] Hy^T) S_i^-1 (z_i - h_i) (:1331-1347), and its support is the number of matched points j
// whose projection under x_i lands within max_divergence pixels of their corner (:1349-1381).  The reference forms the whole
// n-vector x_i per hypothesis (m * n * 2 * (13 + s) multiply-adds and m^2 projections, one after the other); only the camera part
// and point j's own s components of x_i enter projection j, so the work is  m  small per-hypothesis solves (k_ransac_hyp)  and
// m^2 independent (s x (13 + s)) products + projections (k_ransac_support), which read every s x s block of P exactly once.
struct EkfCam { double fx, fy, cx, cy, dx, dy, k1, k2; int distort; };

__device__ __forceinline__ void ekf_rot_from_quat(const double* q, double (&R)[3][3]) {   // quat.cpp:75-91
    R[0][0] = q[0] * q[0] + q[1] * q[1] - q[2] * q[2] - q[3] * q[3];
    R[0][1] = 2 * (q[1] * q[2] - q[0] * q[3]);
    R[0][2] = 2 * (q[1] * q[3] + q[0] * q[2]);
    R[1][0] = 2 * (q[1] * q[2] + q[0] * q[3]);
    R[1][1] = q[0] * q[0] - q[1] * q[1] + q[2] * q[2] - q[3] * q[3];
    R[1][2] = 2 * (q[2] * q[3] - q[0] * q[1]);
    R[2][0] = 2 * (q[1] * q[3] - q[0] * q[2]);
    R[2][1] = 2 * (q[2] * q[3] + q[0] * q[1]);
    R[2][2] = q[0] * q[0] - q[1] * q[1] - q[2] * q[2] + q[3] * q[3];
}
// ProjectInternalSalientPoint (EKF.cpp:2947-2958) with the camera rotation already formed: A.22 / A.21 (scaled by the inverse
// distance), pinhole (:3021-3022), radial distortion (:2960-3005; the closed-form cube roots use the reference's float exponent).
template <int S>
__device__ __forceinline__ void ekf_project(const EkfCam& c, const double* pos, const double (&R)[3][3], const double* sp, double& h0, double& h1) {
    double v[3];
    if (S == 3) {
#pragma unroll
        for (int k = 0; k < 3; ++k) v[k] = sp[k] - pos[k];
    } else {
        const double cos_th = cos(sp[3]), sin_th = sin(sp[3]), cos_ph = cos(sp[4]), sin_ph = sin(sp[4]);
        const double m[3] = {cos_ph * sin_th, -sin_ph, cos_ph * cos_th};
#pragma unroll
        for (int k = 0; k < 3; ++k) v[k] = sp[5] * (sp[k] - pos[k]) + m[k];
    }
    double pc[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) pc[r] = R[0][r] * v[0] + R[1][r] * v[1] + R[2][r] * v[2];
    const double hu0 = c.cx - c.fx * pc[0] / pc[2], hu1 = c.cy - c.fy * pc[1] / pc[2];
    if (!c.distort) { h0 = hu0; h1 = hu1; return; }
    const double ru = sqrt((c.dx * (hu0 - c.cx)) * (c.dx * (hu0 - c.cx)) + (c.dy * (hu1 - c.cy)) * (c.dy * (hu1 - c.cy)));
    double rd;
    if (c.k2 != 0) {           // single real root of the increasing quintic rd + k1 rd^3 + k2 rd^5 = ru (Eigen PolynomialSolver in the reference)
        rd = ru;
        for (int it = 0; it < 60; ++it) {
            const double r2 = rd * rd;
            const double f = rd + c.k1 * r2 * rd + c.k2 * r2 * r2 * rd - ru;
            const double df = 1 + 3 * c.k1 * r2 + 5 * c.k2 * r2 * r2;
            const double step = f / df;
            rd -= step;
            if (fabs(step) <= 1e-17 * fabs(rd)) break;
        }
    } else if (c.k1 == 0) {
        rd = ru;
    } else {
        const double third = (double)(1.0f / 3);
        const double e = pow(9 * c.k1 * c.k1 * ru + sqrt(3 * c.k1 * c.k1 * c.k1 * (4 + 27 * c.k1 * ru * ru)), third);
        rd = (-2 * pow(3.0, third) * c.k1 + pow(2.0, third) * e * e) / (pow(6.0, 2.0 / 3) * c.k1 * e);
    }
    const double stretch = 1 + c.k1 * (rd * rd) + c.k2 * (rd * rd) * (rd * rd);
    h0 = c.cx + (hu0 - c.cx) / stretch;
    h1 = c.cy + (hu1 - c.cy) / stretch;
}

// hyp[i] = { u[13] = Hx^T w, v[S] = Hy^T w, dcam[13] = Pxx u + Pxy_i v },  w = S_i^-1 (z_i - h_i).  One thread per hypothesis.
constexpr int kHypStride = 13 + 6 + 13;
template <int S>
__global__ void __launch_bounds__(128) k_ransac_hyp(int n, int m, const double* __restrict__ P, const double* __restrict__ x, const double* __restrict__ Hcam,
                                                    const double* __restrict__ Hpt, const int64_t* __restrict__ off, const double* __restrict__ z, double meas_var,
                                                    EkfCam cam, double* __restrict__ hyp) {
    __shared__ double Pxx[kCam * kCam];
    for (int e = threadIdx.x; e < kCam * kCam; e += blockDim.x) Pxx[e] = P[(size_t)(e / kCam) * n + (e % kCam)];   // Pxx[c*13 + r] = P(r, c)
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const double* Hx = Hcam + (size_t)(2 * i) * kCam;
    const double* Hy = Hpt + (size_t)(2 * i) * S;
    const size_t oi = (size_t)off[i];
    double hx[2][kCam], hy[2][S];
#pragma unroll
    for (int a = 0; a < 2; ++a) {
#pragma unroll
        for (int c = 0; c < kCam; ++c) hx[a][c] = Hx[a * kCam + c];
#pragma unroll
        for (int c = 0; c < S; ++c) hy[a][c] = Hy[a * S + c];
    }
    // A = Pxx Hx^T (13 x 2),  B = Pxy Hy^T (13 x 2),  Cc = Pyy Hy^T (S x 2)
    double A[kCam][2], B[kCam][2], Cc[S][2];
#pragma unroll
    for (int r = 0; r < kCam; ++r)
#pragma unroll
        for (int a = 0; a < 2; ++a) {
            double t = 0.0;
#pragma unroll
            for (int c = 0; c < kCam; ++c) t += Pxx[c * kCam + r] * hx[a][c];
            A[r][a] = t;
            double t2 = 0.0;
#pragma unroll
            for (int c = 0; c < S; ++c) t2 += P[(oi + c) * (size_t)n + r] * hy[a][c];
            B[r][a] = t2;
        }
#pragma unroll
    for (int r = 0; r < S; ++r)
#pragma unroll
        for (int a = 0; a < 2; ++a) {
            double t = 0.0;
#pragma unroll
            for (int c = 0; c < S; ++c) t += P[(oi + c) * (size_t)n + oi + r] * hy[a][c];
            Cc[r][a] = t;
        }
    double Sm[2][2];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            double t1 = 0.0, mab = 0.0, mba = 0.0, t3 = 0.0;
#pragma unroll
            for (int r = 0; r < kCam; ++r) { t1 += hx[a][r] * A[r][b]; mab += hx[a][r] * B[r][b]; mba += hx[b][r] * B[r][a]; }
#pragma unroll
            for (int r = 0; r < S; ++r) t3 += hy[a][r] * Cc[r][b];
            Sm[a][b] = t1 + mab + mba + t3 + (a == b ? meas_var : 0.0);
        }
    const double idet = 1.0 / (Sm[0][0] * Sm[1][1] - Sm[0][1] * Sm[1][0]);
    double R[3][3];
    ekf_rot_from_quat(x + 3, R);
    double sp[S];
#pragma unroll
    for (int c = 0; c < S; ++c) sp[c] = x[oi + c];
    double h0, h1;
    ekf_project<S>(cam, x, R, sp, h0, h1);
    const double r0 = z[2 * i] - h0, r1 = z[2 * i + 1] - h1;
    const double w0 = (Sm[1][1] * r0 - Sm[0][1] * r1) * idet, w1 = (-Sm[1][0] * r0 + Sm[0][0] * r1) * idet;
    double* out = hyp + (size_t)i * kHypStride;
#pragma unroll
    for (int c = 0; c < kCam; ++c) out[c] = hx[0][c] * w0 + hx[1][c] * w1;                       // u
#pragma unroll
    for (int c = 0; c < S; ++c) out[kCam + c] = hy[0][c] * w0 + hy[1][c] * w1;                    // v
#pragma unroll
    for (int r = 0; r < kCam; ++r) out[kCam + 6 + r] = (A[r][0] + B[r][0]) * w0 + (A[r][1] + B[r][1]) * w1;   // camera part of K (z - h)
}

// grid (ceil(m / 256), m): CTA (jb, i) scores hypothesis i on the matched points j of its slice: point j's components of x_i are
// x_j + P[j.., cam] u_i + P[j.., pt_i] v_i (threads along j: every P column is read as one contiguous run), then the projection.
template <int S>
__global__ void __launch_bounds__(256) k_ransac_support(int n, int m, const double* __restrict__ P, const double* __restrict__ x, const int64_t* __restrict__ off,
                                                        const double* __restrict__ z, EkfCam cam, double max_div, const double* __restrict__ hyp,
                                                        int* __restrict__ support, unsigned* __restrict__ bits, int words) {
    __shared__ double sh[kHypStride];
    __shared__ double camn[kCam];
    __shared__ int cnt;
    const int i = blockIdx.y;
    if (threadIdx.x < kHypStride) sh[threadIdx.x] = hyp[(size_t)i * kHypStride + threadIdx.x];
    if (threadIdx.x == 0) cnt = 0;
    __syncthreads();
    if (threadIdx.x < kCam) camn[threadIdx.x] = x[threadIdx.x] + sh[kCam + 6 + threadIdx.x];
    __syncthreads();
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const size_t oi = (size_t)off[i];
    bool inl = false;
    if (j < m) {
        const size_t oj = (size_t)off[j];
        double sp[S];
#pragma unroll
        for (int r = 0; r < S; ++r) {
            double t = 0.0;
#pragma unroll
            for (int c = 0; c < kCam; ++c) t += P[(size_t)c * n + oj + r] * sh[c];
            double t2 = 0.0;
#pragma unroll
            for (int c = 0; c < S; ++c) t2 += P[(oi + c) * (size_t)n + oj + r] * sh[kCam + c];
            sp[r] = x[oj + r] + (t + t2);
        }
        double R[3][3];
        ekf_rot_from_quat(camn + 3, R);
        double h0, h1;
        ekf_project<S>(cam, camn, R, sp, h0, h1);
        const double d0 = z[2 * j] - h0, d1 = z[2 * j + 1] - h1;
        inl = sqrt(d0 * d0 + d1 * d1) < max_div;
    }
    const unsigned bal = __ballot_sync(0xffffffffu, inl);
    if ((threadIdx.x & 31) == 0) {
        const int wj = j >> 5;
        if (wj < words) bits[(size_t)i * words + wj] = bal;
        if (bal) atomicAdd(&cnt, __popc(bal));
    }
    __syncthreads();
    if (threadIdx.x == 0 && cnt > 0) atomicAdd(&support[i], cnt);
}


// Deriv_hd_by_cam_state_and_sal_pnt for every matched point (EKF.cpp:3067-3159; chain rule A.31-A.55 as in :2651-2865): one thread per
// point writes its two rows of Hcam [2m x 13] (velocity columns zero), Hpt [2m x S] and the projection hd -- what the reference builds
// point by point into a dense [2m x n] H (zero-filled, :3126-3127) before every update.
template <int S>
__global__ void __launch_bounds__(128) k_ekf_jacobians(int m, const double* __restrict__ x, const int64_t* __restrict__ off, EkfCam c, double* __restrict__ Hcam,
                                                       double* __restrict__ Hpt, double* __restrict__ hd_out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const size_t oi = (size_t)off[i];
    double sp[S];
#pragma unroll
    for (int k = 0; k < S; ++k) sp[k] = x[oi + k];
    double pos[3] = {x[0], x[1], x[2]};
    double Rw[3][3];
    ekf_rot_from_quat(x + 3, Rw);
    double h0, h1;
    ekf_project<S>(c, pos, Rw, sp, h0, h1);
    hd_out[2 * i] = h0; hd_out[2 * i + 1] = h1;
    double part2[3], rho = 1.0, cos_th = 0, sin_th = 0, cos_ph = 0, sin_ph = 0;
    if (S == 3) {
#pragma unroll
        for (int k = 0; k < 3; ++k) part2[k] = sp[k] - pos[k];
    } else {
        cos_th = cos(sp[3]); sin_th = sin(sp[3]); cos_ph = cos(sp[4]); sin_ph = sin(sp[4]);
        const double md[3] = {cos_ph * sin_th, -sin_ph, cos_ph * cos_th};
        rho = sp[5];
#pragma unroll
        for (int k = 0; k < 3; ++k) part2[k] = rho * (sp[k] - pos[k]) + md[k];
    }
    double hc[3];   // Rcw(r, k) = Rw[k][r]
#pragma unroll
    for (int r = 0; r < 3; ++r) hc[r] = Rw[0][r] * part2[0] + Rw[1][r] * part2[1] + Rw[2][r] * part2[2];
    double u00 = 1, u01 = 0, u10 = 0, u11 = 1;   // hd_by_hu
    if (c.distort) {
        const double ax = h0 - c.cx, ay = h1 - c.cy;
        const double rd = sqrt((c.dx * ax) * (c.dx * ax) + (c.dy * ay) * (c.dy * ay));
        const double stretch = 1 + c.k1 * (rd * rd) + c.k2 * (rd * rd) * (rd * rd);
        const double kk = c.k1 + 2 * c.k2 * (rd * rd);
        const double side = 2 * kk * ay * ax;
        const double r00 = stretch + 2 * kk * ((c.dx * ax) * (c.dx * ax)), r11 = stretch + 2 * kk * ((c.dy * ay) * (c.dy * ay));
        const double r10 = side * (c.dx * c.dx), r01 = side * (c.dy * c.dy);
        const double idet = 1.0 / (r00 * r11 - r01 * r10);
        u00 = r11 * idet; u01 = -r01 * idet; u10 = -r10 * idet; u11 = r00 * idet;
    }
    const double a00 = -c.fx / hc[2], a02 = c.fx * hc[0] / (hc[2] * hc[2]), a11 = -c.fy / hc[2], a12 = c.fy * hc[1] / (hc[2] * hc[2]);
    double D[2][3];
    D[0][0] = u00 * a00 + u01 * 0.0; D[0][1] = u00 * 0.0 + u01 * a11; D[0][2] = u00 * a02 + u01 * a12;
    D[1][0] = u10 * a00 + u11 * 0.0; D[1][1] = u10 * 0.0 + u11 * a11; D[1][2] = u10 * a02 + u11 * a12;
    double* Hx = Hcam + (size_t)(2 * i) * kCam;
    double* Hy = Hpt + (size_t)(2 * i) * S;
#pragma unroll
    for (int e = 0; e < 2 * kCam; ++e) Hx[e] = 0.0;
    const double scale = S == 3 ? 1.0 : rho;
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            double t = 0.0;
#pragma unroll
            for (int r = 0; r < 3; ++r) t += D[a][r] * (-scale * Rw[k][r]);
            Hx[a * kCam + k] = t;
        }
    const double q[4] = {x[3], -x[4], -x[5], -x[6]};
    const double dR[4][3][3] = {
        {{2 * q[0], -2 * q[3], 2 * q[2]}, {2 * q[3], 2 * q[0], -2 * q[1]}, {-2 * q[2], 2 * q[1], 2 * q[0]}},
        {{2 * q[1], 2 * q[2], 2 * q[3]}, {2 * q[2], -2 * q[1], -2 * q[0]}, {2 * q[3], 2 * q[0], -2 * q[1]}},
        {{-2 * q[2], 2 * q[1], 2 * q[0]}, {2 * q[1], 2 * q[2], 2 * q[3]}, {-2 * q[0], 2 * q[3], -2 * q[2]}},
        {{-2 * q[3], -2 * q[0], 2 * q[1]}, {2 * q[0], -2 * q[3], 2 * q[2]}, {2 * q[1], 2 * q[2], 2 * q[3]}}};
#pragma unroll
    for (int qi = 0; qi < 4; ++qi) {
        double col[3];
#pragma unroll
        for (int r = 0; r < 3; ++r) col[r] = dR[qi][r][0] * part2[0] + dR[qi][r][1] * part2[1] + dR[qi][r][2] * part2[2];
        const double sign = qi == 0 ? 1.0 : -1.0;
#pragma unroll
        for (int a = 0; a < 2; ++a) Hx[a * kCam + 3 + qi] = (D[a][0] * col[0] + D[a][1] * col[1] + D[a][2] * col[2]) * sign;
    }
    double dy[3][S];
    if (S == 3) {
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int k = 0; k < 3; ++k) dy[r][k] = Rw[k][r];
    } else {
        const double dth[3] = {cos_ph * cos_th, 0.0, -cos_ph * sin_th}, dph[3] = {-sin_ph * sin_th, -cos_ph, -sin_ph * cos_th};
        const double dp[3] = {sp[0] - pos[0], sp[1] - pos[1], sp[2] - pos[2]};
#pragma unroll
        for (int r = 0; r < 3; ++r) {
#pragma unroll
            for (int k = 0; k < 3; ++k) dy[r][k] = rho * Rw[k][r];
            dy[r][3 % S] = Rw[0][r] * dth[0] + Rw[1][r] * dth[1] + Rw[2][r] * dth[2];
            dy[r][4 % S] = Rw[0][r] * dph[0] + Rw[1][r] * dph[1] + Rw[2][r] * dph[2];
            dy[r][5 % S] = Rw[0][r] * dp[0] + Rw[1][r] * dp[1] + Rw[2][r] * dp[2];
        }
    }
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int k = 0; k < S; ++k) Hy[a * S + k] = D[a][0] * dy[0][k] + D[a][1] * dy[1][k] + D[a][2] * dy[2][k];
}

// ---------------------------------------------------------------------------------------------------------------------
struct ErrSink { ErrSink& operator=(const std::string& s) { srk_internal_set_error(s.c_str()); return *this; } ErrSink& operator=(const char* s) { srk_internal_set_error(s); return *this; } };
ErrSink g_ekf_error;
#define EKF_CUDA(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { g_ekf_error = std::string(#call) + ": " + cudaGetErrorString(e__); return SRK_E_CUDA; } } while (0)

struct DBuf {
    void* p = nullptr; size_t cap = 0;
    DBuf() = default;
    DBuf(const DBuf&) = delete;
    DBuf& operator=(const DBuf&) = delete;
    ~DBuf() { if (p != nullptr) cudaFree(p); }   // `delete handle` frees every buffer (r_hyp / r_support / r_bits included)
    cudaError_t ensure(size_t bytes) {
        if (p != nullptr && bytes <= cap) return cudaSuccess;
        if (p != nullptr) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = bytes < 256 ? 256 : bytes;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        // development aid: SRK_POISON=1 fills fresh buffers with NaN patterns, so that a read of memory nobody wrote shows up at once
        // (the fill runs on the legacy default stream, which the engine's non-blocking streams do not wait for: finish it before anybody writes the buffer)
        if (e == cudaSuccess && getenv("SRK_POISON") != nullptr) { cudaMemset(p, 0xFF, want); cudaDeviceSynchronize(); }
        return e;
    }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};
const char* kEkfFam[] = {"pht", "innov", "chol", "trsm", "syrk", "state", "predict", "ransac", "chol_trsm"};
enum { E_PHT = 0, E_INNOV, E_CHOL, E_TRSM, E_SYRK, E_STATE, E_PREDICT, E_RANSAC, E_CHOL_TRSM, E_COUNT };

struct Ekf {
    int device = 0;
    cudaStream_t own = nullptr, st = nullptr;
    cudaStream_t hi = nullptr;       // high-priority stream of the factorisation chain while the gain TRSM runs beside it (update_resident)
    std::vector<cudaEvent_t> evs;
    int64_t n = 0;
    DBuf P, x, PHt, S, ws, w, Hcam, Hpt, off, z, h, aux, tmp, neg, info, small;
    DBuf r_hyp, r_support, r_bits;   // 1-point RANSAC scoring: per-hypothesis vectors, support counts, inlier bit rows
    DBuf seq;                        // per-observation update: Jacobian of the point, K, K S
    DBuf Pgrow, xgrow, grow_in;      // covariance growth for new salient points: the grown copies (swapped in) and the small Jacobians
    int64_t launches = 0;
    bool timing = false;
    std::vector<cudaEvent_t> pend[E_COUNT];
    double total[E_COUNT] = {};
    int64_t count[E_COUNT] = {};
};
struct EScope {
    Ekf& e; int f; cudaEvent_t b = nullptr;
    EScope(Ekf& ek, int fam) : e(ek), f(fam) {
        if (!e.timing) return;
        cudaEvent_t a; cudaEventCreate(&a); cudaEventCreate(&b); cudaEventRecord(a, e.st);
        e.pend[f].push_back(a); e.pend[f].push_back(b);
    }
    ~EScope() { if (b != nullptr) cudaEventRecord(b, e.st); }
};
void ekf_resolve(Ekf& e) {
    cudaStreamSynchronize(e.st);
    for (int f = 0; f < E_COUNT; ++f) {
        for (size_t i = 0; i + 1 < e.pend[f].size(); i += 2) {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, e.pend[f][i], e.pend[f][i + 1]) == cudaSuccess) { e.total[f] += ms; e.count[f] += 1; }
            cudaEventDestroy(e.pend[f][i]); cudaEventDestroy(e.pend[f][i + 1]);
        }
        e.pend[f].clear();
    }
}

int update_resident(Ekf& e, int64_t m, const double* Hcam, const double* Hpt, const int64_t* pt_off, int s, const double* z, const double* hpred, double meas_var,
                    int32_t* info_out) {
    if (e.n <= 0) { g_ekf_error = "srk_ekf_update_resident before srk_ekf_set_state"; return SRK_E_NOT_BOUND; }
    if (m <= 0 || Hcam == nullptr || Hpt == nullptr || pt_off == nullptr || z == nullptr || hpred == nullptr || (s != 3 && s != 6)) {
        g_ekf_error = "bad update arguments (s must be 3 or 6)"; return SRK_E_INVALID_ARG;
    }
    for (int64_t i = 0; i < m; ++i) if (pt_off[i] < kCam || pt_off[i] + s > e.n) { g_ekf_error = "salient point offset out of range"; return SRK_E_INVALID_ARG; }
    EKF_CUDA(cudaSetDevice(e.device));
    const int n = (int)e.n, m2 = (int)(2 * m);
    const int64_t lds = ((int64_t)m2 + 7) & ~(int64_t)7;
    const int nz = n + 1;                                  // rows of the TRSM: P H^T and, as row n, the innovation z - h
    const int64_t ldz = ((int64_t)nz + 7) & ~(int64_t)7;  // 16-byte aligned columns for the cp.async operand loads
    cudaStream_t st = e.st;
    EKF_CUDA(e.PHt.ensure(sizeof(double) * (size_t)ldz * m2 + 64)); EKF_CUDA(e.S.ensure(sizeof(double) * (size_t)lds * m2));
    EKF_CUDA(e.ws.ensure(sizeof(double) * srk::dense_cholesky_dinv_doubles(m2))); EKF_CUDA(e.w.ensure(sizeof(double) * (size_t)lds));
    EKF_CUDA(e.Hcam.ensure(sizeof(double) * (size_t)m2 * kCam)); EKF_CUDA(e.Hpt.ensure(sizeof(double) * (size_t)m2 * s));
    EKF_CUDA(e.off.ensure(sizeof(int64_t) * (size_t)m)); EKF_CUDA(e.z.ensure(sizeof(double) * m2)); EKF_CUDA(e.h.ensure(sizeof(double) * m2));
    EKF_CUDA(e.aux.ensure(sizeof(double) * 32)); EKF_CUDA(e.tmp.ensure(sizeof(double) * 4 * (size_t)n)); EKF_CUDA(e.neg.ensure((size_t)n));
    EKF_CUDA(e.info.ensure(sizeof(int) * 4));
    EKF_CUDA(cudaMemcpyAsync(e.Hcam.p, Hcam, sizeof(double) * (size_t)m2 * kCam, cudaMemcpyHostToDevice, st));
    EKF_CUDA(cudaMemcpyAsync(e.Hpt.p, Hpt, sizeof(double) * (size_t)m2 * s, cudaMemcpyHostToDevice, st));
    EKF_CUDA(cudaMemcpyAsync(e.off.p, pt_off, sizeof(int64_t) * (size_t)m, cudaMemcpyHostToDevice, st));
    EKF_CUDA(cudaMemcpyAsync(e.z.p, z, sizeof(double) * m2, cudaMemcpyHostToDevice, st));
    EKF_CUDA(cudaMemcpyAsync(e.h.p, hpred, sizeof(double) * m2, cudaMemcpyHostToDevice, st));
    double* P = e.P.as<double>(); double* PHt = e.PHt.as<double>(); double* S = e.S.as<double>(); double* ws = e.ws.as<double>();
    {
        EScope sc(e, E_PHT);
        dim3 grid((n + 255) / 256, (m2 + 7) / 8);
        if (s == 3) k_ekf_pht<3><<<grid, 256, 0, st>>>(n, m2, P, e.Hcam.as<double>(), e.Hpt.as<double>(), e.off.as<int64_t>(), PHt, ldz);
        else k_ekf_pht<6><<<grid, 256, 0, st>>>(n, m2, P, e.Hcam.as<double>(), e.Hpt.as<double>(), e.off.as<int64_t>(), PHt, ldz);
        e.launches += 1;
    }
    {
        EScope sc(e, E_INNOV);
        EKF_CUDA(cudaMemsetAsync(S, 0, sizeof(double) * (size_t)lds * m2, st));
        dim3 grid((m2 + 255) / 256, m2);
        if (s == 3) k_ekf_innov<3><<<grid, 256, 0, st>>>(n, m2, PHt, e.Hcam.as<double>(), e.Hpt.as<double>(), e.off.as<int64_t>(), meas_var, S, lds, ldz);
        else k_ekf_innov<6><<<grid, 256, 0, st>>>(n, m2, PHt, e.Hcam.as<double>(), e.Hpt.as<double>(), e.off.as<int64_t>(), meas_var, S, lds, ldz);
        e.launches += 1;
    }
    k_ekf_innovation_vec<<<(m2 + 255) / 256, 256, 0, st>>>(m2, e.z.as<double>(), e.h.as<double>(), PHt + n, ldz);
    e.launches += 1;
    // Factorisation of S and the gain TRSM Z = PHt * L^-T side by side (chol_kernels.cu dense_cholesky_factor_trsm): the factorisation is a
    // latency chain on a handful of SMs, the TRSM needs of L only the panels that are already final.  Z overwrites PHt, which is scratch,
    // so the positive-definiteness check below may come after it.  SRK_EKF_OVERLAP=0: one after the other (development aid).
    bool overlapped = false;
    {
        static int overlap_env = -1;
        if (overlap_env < 0) { const char* oe = getenv("SRK_EKF_OVERLAP"); overlap_env = (oe != nullptr && oe[0] == '0') ? 0 : 1; }
        if (overlap_env && m2 >= 1024) {
            if (e.hi == nullptr) {
                int least = 0, greatest = 0;
                cudaDeviceGetStreamPriorityRange(&least, &greatest);
                if (cudaStreamCreateWithPriority(&e.hi, cudaStreamNonBlocking, greatest) != cudaSuccess) { cudaGetLastError(); e.hi = nullptr; }
            }
            const size_t need = (size_t)m2 / 256 + 4;
            while (e.hi != nullptr && e.evs.size() < need) {
                cudaEvent_t ev = nullptr;
                if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); break; }
                e.evs.push_back(ev);
            }
            if (e.hi != nullptr && e.evs.size() >= need) {
                EScope sc(e, E_CHOL_TRSM);
                const int64_t nl = srk::dense_cholesky_factor_trsm(e.hi, st, m2, S, lds, ws, e.info.as<int>(), nz, PHt, ldz, e.evs.data(), (int)e.evs.size());
                if (nl >= 0) { e.launches += nl; overlapped = true; }
            }
        }
    }
    if (!overlapped) {
        EScope sc(e, E_CHOL);
        e.launches += srk::dense_cholesky_factor(st, m2, S, lds, ws, e.info.as<int>());
    }
    {   // A failed factorisation must not reach the state: P and x are still untouched here (P*H^T and S are scratch).  The reference forms
        // a general inverse (EKF.cpp:1017-1019) and has no such failure; an innovation covariance that is not positive definite means the
        // filter has already diverged, and the caller is told instead of receiving NaNs.
        int hchol = 0;
        EKF_CUDA(cudaMemcpyAsync(&hchol, e.info.p, sizeof(int), cudaMemcpyDeviceToHost, st));
        EKF_CUDA(cudaStreamSynchronize(st));
        if (hchol != 0) {
            if (info_out != nullptr) *info_out = hchol;
            g_ekf_error = std::string("innovation covariance H P H^T + R is not numerically positive definite (pivot ") + std::to_string(hchol) + "); state left untouched";
            return SRK_E_NOT_POSDEF;
        }
    }
    if (!overlapped) {   // Z = PHt * L^-T; Z overwrites PHt.  Right-looking over 64-column blocks INSIDE a 256-column panel, then ONE update of everything
        // right of the panel with K = 256: a K = 64 update of the whole trailing matrix per block column re-reads and re-writes PHt
        // m2 / 64 times (12 GB at n = 6013, 2m = 4000 -- memory-bound at 14 TFLOP/s); with the panel the trailing traffic drops fourfold.
        EScope sc(e, E_TRSM);
        // inside a panel: ONE launch (48-row strips that walk the whole panel, chol_kernels.cu k_strip_trsm) instead of a right solve + a
        // K = 64 product per block column.  SRK_EKF_TRSM_PANEL: panel width (256 / 512), 0 = the launch-per-block-column form (development aid)
        static int panel_env = -1;
        if (panel_env < 0) { const char* pe = getenv("SRK_EKF_TRSM_PANEL"); panel_env = pe != nullptr ? atoi(pe) : 512; if (panel_env != 0 && panel_env != 256 && panel_env != 512) panel_env = 512; }
        const int kPanel = panel_env == 0 ? 512 : panel_env;
        for (int p0 = 0; p0 < m2; p0 += kPanel) {
            const int pend = m2 < p0 + kPanel ? m2 : p0 + kPanel;
            if (panel_env != 0) {
                srk::launch_strip_trsm(st, nz, pend - p0, PHt + (size_t)p0 * ldz, ldz, S + (size_t)p0 * lds + p0, lds, srk::dense_cholesky_dinv_block(ws, p0 / 64));
                e.launches += 1;
            }
            for (int k0 = p0; k0 < pend && panel_env == 0; k0 += 64) {
                srk::launch_block_right_solve(st, nz, PHt + (size_t)k0 * ldz, ldz, srk::dense_cholesky_dinv_block(ws, k0 / 64), m2 - k0);   // the last block may be ragged
                e.launches += 1;
                const int rest = pend - (k0 + 64);
                if (rest > 0) {   // the other blocks of the panel: PHt[:, k0+64:pend] -= Z_k (n x 64) * L[k0+64:pend, k0:k0+64]^T
                    srk::launch_gemm_nt_dmma(st, nz, rest, 64, PHt + (size_t)k0 * ldz, ldz, S + (size_t)k0 * lds + k0 + 64, lds, PHt + (size_t)(k0 + 64) * ldz, ldz, 0);
                    e.launches += 1;
                }
            }
            if (pend < m2) {      // PHt[:, pend:] -= Z_panel (n x 256) * L[pend:, p0:pend]^T
                srk::launch_gemm_nt_dmma(st, nz, m2 - pend, pend - p0, PHt + (size_t)p0 * ldz, ldz, S + (size_t)p0 * lds + pend, lds, PHt + (size_t)pend * ldz, ldz, 0);
                e.launches += 1;
            }
        }
    }
    {
        EScope sc(e, E_STATE);
        k_ekf_take_row<<<(m2 + 255) / 256, 256, 0, st>>>(m2, PHt + n, ldz, e.w.as<double>());      // L^-1 (z - h), carried through the TRSM as row n
        e.launches += 1;
        k_ekf_state<<<(n + 31) / 32, 256, 0, st>>>(n, m2, PHt, ldz, e.w.as<double>(), e.x.as<double>());
        e.launches += 1;
    }
    {
        EScope sc(e, E_SYRK);
        srk::launch_gemm_nt_dmma(st, n, n, m2, PHt, ldz, PHt, ldz, P, n, 1);
        srk::launch_mirror_lower(st, n, P, n);
        e.launches += 2;
    }
    {
        EScope sc(e, E_STATE);
        k_ekf_quat_prepare<<<1, 32, 0, st>>>(e.x.as<double>(), e.aux.as<double>());
        k_ekf_quat_columns<<<(n + 255) / 256, 256, 0, st>>>(n, P, e.aux.as<double>(), e.tmp.as<double>());
        k_ekf_quat_write<<<(n + 255) / 256, 256, 0, st>>>(n, P, e.aux.as<double>(), e.tmp.as<double>());
        EKF_CUDA(cudaMemsetAsync(e.info.as<int>() + 1, 0, sizeof(int), st));
        k_ekf_nonneg_flag<<<(n + 255) / 256, 256, 0, st>>>(n, P, e.neg.as<unsigned char>(), e.info.as<int>() + 1);
        k_ekf_nonneg_zero<<<dim3((n + 255) / 256, n), 256, 0, st>>>(n, P, e.neg.as<unsigned char>(), e.info.as<int>() + 1);
        e.launches += 5;
    }
    int hinfo = 0;
    EKF_CUDA(cudaMemcpyAsync(&hinfo, e.info.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    EKF_CUDA(cudaStreamSynchronize(st));
    EKF_CUDA(cudaGetLastError());
    if (info_out != nullptr) *info_out = hinfo;
    return SRK_OK;
}

int predict_resident(Ekf& e, const double* F13, const double* GQGt13, const double* cam_new) {
    if (e.n <= 0) { g_ekf_error = "srk_ekf_predict_resident before srk_ekf_set_state"; return SRK_E_NOT_BOUND; }
    if (F13 == nullptr || GQGt13 == nullptr) { g_ekf_error = "null F / GQGt"; return SRK_E_INVALID_ARG; }
    EKF_CUDA(cudaSetDevice(e.device));
    const int n = (int)e.n;
    EKF_CUDA(e.small.ensure(sizeof(double) * (3 * kCam * kCam + kCam)));
    double* dF = e.small.as<double>(); double* dQ = dF + kCam * kCam; double* dVV = dQ + kCam * kCam; double* dCam = dVV + kCam * kCam;
    EKF_CUDA(cudaMemcpyAsync(dF, F13, sizeof(double) * kCam * kCam, cudaMemcpyHostToDevice, e.st));
    EKF_CUDA(cudaMemcpyAsync(dQ, GQGt13, sizeof(double) * kCam * kCam, cudaMemcpyHostToDevice, e.st));
    if (cam_new != nullptr) EKF_CUDA(cudaMemcpyAsync(dCam, cam_new, sizeof(double) * kCam, cudaMemcpyHostToDevice, e.st));
    EScope sc(e, E_PREDICT);
    k_ekf_predict_vv<<<1, 192, 0, e.st>>>(n, e.P.as<double>(), dF, dQ, dVV);
    k_ekf_predict_vm<<<(n + 127) / 128, 128, 0, e.st>>>(n, e.P.as<double>(), dF, dVV, cam_new != nullptr ? dCam : nullptr, e.x.as<double>());
    e.launches += 2;
    return SRK_OK;
}


int measurement_jacobians(Ekf& e, int64_t m, const int64_t* pt_off, int s, const srk_ekf_camera* cp, double* Hcam, double* Hpt, double* h_pred) {
    if (e.n <= 0) { g_ekf_error = "srk_ekf_measurement_jacobians_resident before srk_ekf_set_state"; return SRK_E_NOT_BOUND; }
    if (m <= 0 || pt_off == nullptr || cp == nullptr || Hcam == nullptr || Hpt == nullptr || h_pred == nullptr || (s != 3 && s != 6)) {
        g_ekf_error = "bad Jacobian arguments (s must be 3 or 6)"; return SRK_E_INVALID_ARG;
    }
    for (int64_t i = 0; i < m; ++i) if (pt_off[i] < kCam || pt_off[i] + s > e.n) { g_ekf_error = "salient point offset out of range"; return SRK_E_INVALID_ARG; }
    EKF_CUDA(cudaSetDevice(e.device));
    const int mi = (int)m, m2 = (int)(2 * m);
    cudaStream_t st = e.st;
    EKF_CUDA(e.Hcam.ensure(sizeof(double) * (size_t)m2 * kCam)); EKF_CUDA(e.Hpt.ensure(sizeof(double) * (size_t)m2 * s));
    EKF_CUDA(e.off.ensure(sizeof(int64_t) * (size_t)m)); EKF_CUDA(e.h.ensure(sizeof(double) * m2));
    EKF_CUDA(cudaMemcpyAsync(e.off.p, pt_off, sizeof(int64_t) * (size_t)m, cudaMemcpyHostToDevice, st));
    EkfCam cam{cp->fx_pix, cp->fy_pix, cp->cx, cp->cy, cp->dx_mm, cp->dy_mm, cp->k1, cp->k2, cp->enable_distortion != 0 ? 1 : 0};
    {
        EScope sc(e, E_STATE);
        if (s == 3) k_ekf_jacobians<3><<<(mi + 127) / 128, 128, 0, st>>>(mi, e.x.as<double>(), e.off.as<int64_t>(), cam, e.Hcam.as<double>(), e.Hpt.as<double>(), e.h.as<double>());
        else k_ekf_jacobians<6><<<(mi + 127) / 128, 128, 0, st>>>(mi, e.x.as<double>(), e.off.as<int64_t>(), cam, e.Hcam.as<double>(), e.Hpt.as<double>(), e.h.as<double>());
        e.launches += 1;
    }
    EKF_CUDA(cudaMemcpyAsync(Hcam, e.Hcam.p, sizeof(double) * (size_t)m2 * kCam, cudaMemcpyDeviceToHost, st));
    EKF_CUDA(cudaMemcpyAsync(Hpt, e.Hpt.p, sizeof(double) * (size_t)m2 * s, cudaMemcpyDeviceToHost, st));
    EKF_CUDA(cudaMemcpyAsync(h_pred, e.h.p, sizeof(double) * m2, cudaMemcpyDeviceToHost, st));
    EKF_CUDA(cudaStreamSynchronize(st));
    EKF_CUDA(cudaGetLastError());
    return SRK_OK;
}

int ransac_consensus(Ekf& e, int64_t m, const double* Hcam, const double* Hpt, const int64_t* pt_off, int s, const double* z, double meas_var, const srk_ekf_camera* cp,
                     double max_div, int32_t* support_out, int32_t* best_out, unsigned char* best_inliers) {
    if (e.n <= 0) { g_ekf_error = "srk_ekf_ransac_consensus_resident before srk_ekf_set_state"; return SRK_E_NOT_BOUND; }
    if (m <= 0 || Hcam == nullptr || Hpt == nullptr || pt_off == nullptr || z == nullptr || cp == nullptr || (s != 3 && s != 6)) {
        g_ekf_error = "bad consensus arguments (s must be 3 or 6)"; return SRK_E_INVALID_ARG;
    }
    for (int64_t i = 0; i < m; ++i) if (pt_off[i] < kCam || pt_off[i] + s > e.n) { g_ekf_error = "salient point offset out of range"; return SRK_E_INVALID_ARG; }
    EKF_CUDA(cudaSetDevice(e.device));
    const int n = (int)e.n, mi = (int)m, m2 = (int)(2 * m);
    const int words = (mi + 31) / 32;
    cudaStream_t st = e.st;
    EKF_CUDA(e.Hcam.ensure(sizeof(double) * (size_t)m2 * kCam)); EKF_CUDA(e.Hpt.ensure(sizeof(double) * (size_t)m2 * s));
    EKF_CUDA(e.off.ensure(sizeof(int64_t) * (size_t)m)); EKF_CUDA(e.z.ensure(sizeof(double) * m2));
    EKF_CUDA(e.r_hyp.ensure(sizeof(double) * (size_t)m * kHypStride)); EKF_CUDA(e.r_support.ensure(sizeof(int) * (size_t)m));
    EKF_CUDA(e.r_bits.ensure(sizeof(unsigned) * (size_t)m * words));
    EKF_CUDA(cudaMemcpyAsync(e.Hcam.p, Hcam, sizeof(double) * (size_t)m2 * kCam, cudaMemcpyHostToDevice, st));
    EKF_CUDA(cudaMemcpyAsync(e.Hpt.p, Hpt, sizeof(double) * (size_t)m2 * s, cudaMemcpyHostToDevice, st));
    EKF_CUDA(cudaMemcpyAsync(e.off.p, pt_off, sizeof(int64_t) * (size_t)m, cudaMemcpyHostToDevice, st));
    EKF_CUDA(cudaMemcpyAsync(e.z.p, z, sizeof(double) * m2, cudaMemcpyHostToDevice, st));
    EKF_CUDA(cudaMemsetAsync(e.r_support.p, 0, sizeof(int) * (size_t)m, st));
    EkfCam cam{cp->fx_pix, cp->fy_pix, cp->cx, cp->cy, cp->dx_mm, cp->dy_mm, cp->k1, cp->k2, cp->enable_distortion != 0 ? 1 : 0};
    {
        EScope sc(e, E_RANSAC);
        const dim3 gs((mi + 255) / 256, mi);
        if (s == 3) {
            k_ransac_hyp<3><<<(mi + 127) / 128, 128, 0, st>>>(n, mi, e.P.as<double>(), e.x.as<double>(), e.Hcam.as<double>(), e.Hpt.as<double>(), e.off.as<int64_t>(),
                                                             e.z.as<double>(), meas_var, cam, e.r_hyp.as<double>());
            k_ransac_support<3><<<gs, 256, 0, st>>>(n, mi, e.P.as<double>(), e.x.as<double>(), e.off.as<int64_t>(), e.z.as<double>(), cam, max_div, e.r_hyp.as<double>(),
                                                    e.r_support.as<int>(), e.r_bits.as<unsigned>(), words);
        } else {
            k_ransac_hyp<6><<<(mi + 127) / 128, 128, 0, st>>>(n, mi, e.P.as<double>(), e.x.as<double>(), e.Hcam.as<double>(), e.Hpt.as<double>(), e.off.as<int64_t>(),
                                                             e.z.as<double>(), meas_var, cam, e.r_hyp.as<double>());
            k_ransac_support<6><<<gs, 256, 0, st>>>(n, mi, e.P.as<double>(), e.x.as<double>(), e.off.as<int64_t>(), e.z.as<double>(), cam, max_div, e.r_hyp.as<double>(),
                                                    e.r_support.as<int>(), e.r_bits.as<unsigned>(), words);
        }
        e.launches += 2;
    }
    std::vector<int32_t> sup((size_t)m);
    EKF_CUDA(cudaMemcpyAsync(sup.data(), e.r_support.p, sizeof(int) * (size_t)m, cudaMemcpyDeviceToHost, st));
    EKF_CUDA(cudaStreamSynchronize(st));
    int best = -1; int32_t best_cnt = 0;
    for (int64_t i = 0; i < m; ++i) if (sup[(size_t)i] > best_cnt) { best_cnt = sup[(size_t)i]; best = (int)i; }   // strictly more: the earliest maximum (EKF.cpp:1383)
    if (support_out != nullptr) std::memcpy(support_out, sup.data(), sizeof(int32_t) * (size_t)m);
    if (best_out != nullptr) *best_out = best;
    if (best_inliers != nullptr) {
        std::memset(best_inliers, 0, (size_t)m);
        if (best >= 0) {
            std::vector<unsigned> row((size_t)words);
            EKF_CUDA(cudaMemcpyAsync(row.data(), e.r_bits.as<unsigned>() + (size_t)best * words, sizeof(unsigned) * (size_t)words, cudaMemcpyDeviceToHost, st));
            EKF_CUDA(cudaStreamSynchronize(st));
            for (int64_t j = 0; j < m; ++j) best_inliers[j] = (row[(size_t)(j >> 5)] >> (j & 31)) & 1u;
        }
    }
    EKF_CUDA(cudaGetLastError());
    return SRK_OK;
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------------------------------
// Covariance growth for new salient points (AllocateAndInitStateForNewSalientPoint, EKF.cpp:2322-2396, with the P-dependent products of
// GetNewSphericalSalientPointCovar :2528-2546 and the XYZ conversion :2586-2592), for k new points in one pass.  The reference resizes the
// dense matrix once per point (conservativeResize: a temporary + a full copy each time); here the old block is copied once (a device
// copy at HBM speed) and the k*s new rows / columns are filled by one kernel:
//     P[new_i, old]   = Jy_i P[0:7, old]                      (bottom-left stripe, mirrored to the top right)
//     P[new_i, new_l] = Jy_i P[0:7, 0:7] Jy_l^T  (+ Qnew_i when i == l)
// which is what adding the points one after the other yields (P[0:7, new_l] = P[0:7, 0:7] Jy_l^T once point l exists).
__global__ void __launch_bounds__(256) k_ekf_grow(int n, int n2, int k, int s, const double* __restrict__ Pold, double* __restrict__ Pnew,
                                                  const double* __restrict__ Jy, const double* __restrict__ Q, int diag_only) {
    const int r = n + blockIdx.y;                 // new row
    const int kk = blockIdx.y / s, a = blockIdx.y - kk * s;
    const double* jy = Jy + ((size_t)kk * s + a) * 7;
    double j7[7];
#pragma unroll
    for (int q = 0; q < 7; ++q) j7[q] = jy[q];
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < n2; c += gridDim.x * blockDim.x) {
        double v = 0.0;
        if (c < n) {
            if (!diag_only) {
#pragma unroll
                for (int q = 0; q < 7; ++q) v += j7[q] * Pold[(size_t)c * n + q];
            }
            Pnew[(size_t)c * n2 + r] = v;
            Pnew[(size_t)r * n2 + c] = v;
        } else {
            const int ll = (c - n) / s, b = (c - n) - ll * s;
            if (c > r) continue;                  // the lower triangle of the new-new block; mirrored below
            if (!diag_only) {
                const double* jl = Jy + ((size_t)ll * s + b) * 7;
                for (int t = 0; t < 7; ++t) {
                    double w = 0.0;
#pragma unroll
                    for (int q = 0; q < 7; ++q) w += j7[q] * Pold[(size_t)t * n + q];
                    v += w * jl[t];
                }
            }
            // the reference assigns the auto-covariance block entry by entry as computed (:2386, :2394): Qnew is not symmetrised
            Pnew[(size_t)c * n2 + r] = ll == kk ? v + Q[((size_t)kk * s + a) * s + b] : v;     // P(r, c)
            Pnew[(size_t)r * n2 + c] = ll == kk ? v + Q[((size_t)kk * s + b) * s + a] : v;     // P(c, r)
        }
    }
}

// Projected 2-D covariance of every listed point at the resident state: covar2D = J P_in J^T with J = [d hd / d (camera position, quaternion)
// | d hd / d point] and P_in the matching blocks of P (GetSalientPointProjected2DPosWithUncertainty, EKF.cpp:3901-4025, the batched form the
// second stage of the 1-point RANSAC update needs, :1483).  One thread per point; cov [m][4] row-major, symmetrised as FixAlmostSymmetricMat does.
template <int S>
__global__ void __launch_bounds__(128) k_ekf_proj_cov(int m, int n, const double* __restrict__ P, const double* __restrict__ Hcam, const double* __restrict__ Hpt,
                                                      const int64_t* __restrict__ off, double* __restrict__ cov) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    constexpr int D = 7 + S;
    const size_t o = (size_t)off[i];
    double J[2][D];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
#pragma unroll
        for (int c = 0; c < 7; ++c) J[k][c] = Hcam[(size_t)(2 * i + k) * kCam + c];
#pragma unroll
        for (int c = 0; c < S; ++c) J[k][7 + c] = Hpt[(size_t)(2 * i + k) * S + c];
    }
    double c2[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
    for (int a = 0; a < D; ++a) {
        const size_t ra = a < 7 ? (size_t)a : o + (size_t)(a - 7);
        double t0 = 0.0, t1 = 0.0;      // (P_in J^T)[a][0..1]
        for (int b = 0; b < D; ++b) {
            const size_t cb = b < 7 ? (size_t)b : o + (size_t)(b - 7);
            const double p = P[cb * (size_t)n + ra];
            t0 += p * J[0][b]; t1 += p * J[1][b];
        }
        c2[0][0] += J[0][a] * t0; c2[0][1] += J[0][a] * t1; c2[1][0] += J[1][a] * t0; c2[1][1] += J[1][a] * t1;
    }
    const double offd = (c2[0][1] + c2[1][0]) / 2;
    cov[4 * (size_t)i] = c2[0][0]; cov[4 * (size_t)i + 1] = offd; cov[4 * (size_t)i + 2] = offd; cov[4 * (size_t)i + 3] = c2[1][1];
}

int projected_covariances(Ekf& e, int64_t m, const double* Hcam, const double* Hpt, const int64_t* pt_off, int s, double* cov) {
    if (e.n <= 0) { g_ekf_error = "srk_ekf_projected_covariances_resident before srk_ekf_set_state"; return SRK_E_NOT_BOUND; }
    if (m <= 0 || Hcam == nullptr || Hpt == nullptr || pt_off == nullptr || cov == nullptr || (s != 3 && s != 6)) { g_ekf_error = "bad arguments (s must be 3 or 6)"; return SRK_E_INVALID_ARG; }
    for (int64_t i = 0; i < m; ++i) if (pt_off[i] < kCam || pt_off[i] + s > e.n) { g_ekf_error = "salient point offset out of range"; return SRK_E_INVALID_ARG; }
    EKF_CUDA(cudaSetDevice(e.device));
    const int mi = (int)m, m2 = (int)(2 * m);
    cudaStream_t st = e.st;
    EKF_CUDA(e.Hcam.ensure(sizeof(double) * (size_t)m2 * kCam)); EKF_CUDA(e.Hpt.ensure(sizeof(double) * (size_t)m2 * s));
    EKF_CUDA(e.off.ensure(sizeof(int64_t) * (size_t)m)); EKF_CUDA(e.small.ensure(sizeof(double) * (3 * kCam * kCam + kCam)));
    EKF_CUDA(e.tmp.ensure(sizeof(double) * 4 * (size_t)(m > e.n ? m : e.n)));
    EKF_CUDA(cudaMemcpyAsync(e.Hcam.p, Hcam, sizeof(double) * (size_t)m2 * kCam, cudaMemcpyHostToDevice, st));
    EKF_CUDA(cudaMemcpyAsync(e.Hpt.p, Hpt, sizeof(double) * (size_t)m2 * s, cudaMemcpyHostToDevice, st));
    EKF_CUDA(cudaMemcpyAsync(e.off.p, pt_off, sizeof(int64_t) * (size_t)m, cudaMemcpyHostToDevice, st));
    {
        EScope sc(e, E_RANSAC);
        if (s == 3) k_ekf_proj_cov<3><<<(mi + 127) / 128, 128, 0, st>>>(mi, (int)e.n, e.P.as<double>(), e.Hcam.as<double>(), e.Hpt.as<double>(), e.off.as<int64_t>(), e.tmp.as<double>());
        else k_ekf_proj_cov<6><<<(mi + 127) / 128, 128, 0, st>>>(mi, (int)e.n, e.P.as<double>(), e.Hcam.as<double>(), e.Hpt.as<double>(), e.off.as<int64_t>(), e.tmp.as<double>());
        e.launches += 1;
    }
    EKF_CUDA(cudaMemcpyAsync(cov, e.tmp.p, sizeof(double) * 4 * (size_t)m, cudaMemcpyDeviceToHost, st));
    EKF_CUDA(cudaStreamSynchronize(st));
    EKF_CUDA(cudaGetLastError());
    return SRK_OK;
}

// ---------------------------------------------------------------------------------------------------------------------------------
// Per-observation update variants: ProcessFrame_OneObservationPerUpdate (EKF.cpp:1153-1269; 2x2 innovation per observed point) and
// ProcessFrame_OneComponentOfOneObservationPerUpdate (:1525-1650; scalar innovation per pixel component).  Every update re-derives the
// measurement Jacobian at the LATEST state, so the m (or 2m) updates form a dependent chain of rank-2 (rank-1) updates of the whole P:
// per update one launch for the Jacobian of the point, one for the innovation covariance + gain + state, one for P -= (K S) K^T (the HBM-bound
// one: 16 n^2 bytes), three for the quaternion normalisation.  FixSymmetricMat (:4308) is folded into the covariance kernel: the entry written
// is P(r, c) - (KS_r . K_c + KS_c . K_r) / 2.
// scratch (doubles): [0..25] Hx, [26..37] Hy, [38..39] hd  (written by k_ekf_jacobians with m = 1), then K [2n], KS [2n].
template <int S>
__global__ void __launch_bounds__(256) k_ekf_seq_gain(int n, const double* __restrict__ P, const double* __restrict__ jac, int64_t off, const double* __restrict__ z2,
                                                      double meas_var, int comp /* -1: both components, 0 / 1: one */, double* __restrict__ x,
                                                      double* __restrict__ K, double* __restrict__ KS) {
    __shared__ double sH[2][kCam + S];
    __shared__ double sS[2][2], sSi[2][2], sDelta[2];
    const int tid = threadIdx.x;
    const int k0 = comp < 0 ? 0 : comp, d = comp < 0 ? 2 : 1;
    if (tid < 2 * (kCam + S)) {
        const int k = tid / (kCam + S), c = tid - k * (kCam + S);
        sH[k][c] = c < kCam ? jac[k * kCam + c] : jac[26 + k * S + (c - kCam)];
    }
    __syncthreads();
    if (tid < d * d) {   // S = H P_in H^T + R over the camera (13) and the point (S) variables (:1208-1216, :1591-1597)
        const int a = k0 + tid / d, b = k0 + tid % d;
        double acc = 0.0;
        for (int i = 0; i < kCam + S; ++i) {
            const size_t ri = i < kCam ? (size_t)i : (size_t)off + (size_t)(i - kCam);
            double t = 0.0;
            for (int j = 0; j < kCam + S; ++j) {
                const size_t cj = j < kCam ? (size_t)j : (size_t)off + (size_t)(j - kCam);
                t += P[cj * (size_t)n + ri] * sH[b][j];
            }
            acc += sH[a][i] * t;
        }
        sS[tid / d][tid % d] = acc + (a == b ? meas_var : 0.0);
    }
    __syncthreads();
    if (tid == 0) {
        if (d == 2) { const double det = sS[0][0] * sS[1][1] - sS[0][1] * sS[1][0]; sSi[0][0] = sS[1][1] / det; sSi[0][1] = -sS[0][1] / det; sSi[1][0] = -sS[1][0] / det; sSi[1][1] = sS[0][0] / det; }
        else { sSi[0][0] = 1.0 / sS[0][0]; sSi[0][1] = sSi[1][0] = sSi[1][1] = 0.0; sS[0][1] = sS[1][0] = sS[1][1] = 0.0; }
        for (int a = 0; a < 2; ++a) sDelta[a] = a < d ? z2[k0 + a] - jac[38 + k0 + a] : 0.0;
    }
    __syncthreads();
    const int r = blockIdx.x * blockDim.x + tid;
    if (r >= n) return;
    double ph[2] = {0.0, 0.0};
    for (int a = 0; a < d; ++a) {
        double v = 0.0;
#pragma unroll
        for (int c = 0; c < kCam; ++c) v += P[(size_t)c * n + r] * sH[k0 + a][c];
#pragma unroll
        for (int c = 0; c < S; ++c) v += P[((size_t)off + c) * n + r] * sH[k0 + a][kCam + c];
        ph[a] = v;
    }
    double kk[2] = {0.0, 0.0}, ks[2] = {0.0, 0.0};
    for (int a = 0; a < d; ++a) kk[a] = ph[0] * sSi[0][a] + (d == 2 ? ph[1] * sSi[1][a] : 0.0);
    for (int a = 0; a < d; ++a) ks[a] = kk[0] * sS[0][a] + (d == 2 ? kk[1] * sS[1][a] : 0.0);
    x[r] += kk[0] * sDelta[0] + kk[1] * sDelta[1];
    K[2 * (size_t)r] = kk[0]; K[2 * (size_t)r + 1] = kk[1];
    KS[2 * (size_t)r] = ks[0]; KS[2 * (size_t)r + 1] = ks[1];
}
__global__ void __launch_bounds__(256) k_ekf_seq_cov(int n, double* __restrict__ P, const double* __restrict__ K, const double* __restrict__ KS) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x, c = blockIdx.y;
    if (r >= n) return;
    const double kr0 = K[2 * (size_t)r], kr1 = K[2 * (size_t)r + 1], sr0 = KS[2 * (size_t)r], sr1 = KS[2 * (size_t)r + 1];
    const double kc0 = K[2 * (size_t)c], kc1 = K[2 * (size_t)c + 1], sc0 = KS[2 * (size_t)c], sc1 = KS[2 * (size_t)c + 1];
    // P is symmetric on entry: (P - D) symmetrised = P - (D + D^T) / 2 with D(r, c) = KS_r . K_c
    P[(size_t)c * n + r] -= 0.5 * ((sr0 * kc0 + sr1 * kc1) + (sc0 * kr0 + sc1 * kr1));
}

int sequential_update(Ekf& e, int64_t m, const int64_t* pt_off, int s, const double* z, const srk_ekf_camera* cp, double meas_var, int per_component) {
    if (e.n <= 0) { g_ekf_error = "srk_ekf_sequential_update_resident before srk_ekf_set_state"; return SRK_E_NOT_BOUND; }
    if (m <= 0 || pt_off == nullptr || z == nullptr || cp == nullptr || (s != 3 && s != 6)) { g_ekf_error = "bad arguments (s must be 3 or 6)"; return SRK_E_INVALID_ARG; }
    for (int64_t i = 0; i < m; ++i) if (pt_off[i] < kCam || pt_off[i] + s > e.n) { g_ekf_error = "salient point offset out of range"; return SRK_E_INVALID_ARG; }
    EKF_CUDA(cudaSetDevice(e.device));
    const int n = (int)e.n;
    cudaStream_t st = e.st;
    EKF_CUDA(e.off.ensure(sizeof(int64_t) * (size_t)m)); EKF_CUDA(e.z.ensure(sizeof(double) * 2 * (size_t)m));
    EKF_CUDA(e.seq.ensure(sizeof(double) * (64 + 4 * (size_t)n)));
    EKF_CUDA(e.aux.ensure(sizeof(double) * 32)); EKF_CUDA(e.tmp.ensure(sizeof(double) * 4 * (size_t)n));
    EKF_CUDA(cudaMemcpyAsync(e.off.p, pt_off, sizeof(int64_t) * (size_t)m, cudaMemcpyHostToDevice, st));
    EKF_CUDA(cudaMemcpyAsync(e.z.p, z, sizeof(double) * 2 * (size_t)m, cudaMemcpyHostToDevice, st));
    EkfCam cam{cp->fx_pix, cp->fy_pix, cp->cx, cp->cy, cp->dx_mm, cp->dy_mm, cp->k1, cp->k2, cp->enable_distortion != 0 ? 1 : 0};
    double* jac = e.seq.as<double>(); double* K = jac + 64; double* KS = K + 2 * (size_t)n;
    EScope sc(e, E_STATE);
    for (int64_t i = 0; i < m; ++i) {
        const int passes = per_component ? 2 : 1;
        for (int pass = 0; pass < passes; ++pass) {
            const int comp = per_component ? pass : -1;
            if (s == 3) {
                k_ekf_jacobians<3><<<1, 128, 0, st>>>(1, e.x.as<double>(), e.off.as<int64_t>() + i, cam, jac, jac + 26, jac + 38);
                k_ekf_seq_gain<3><<<(n + 255) / 256, 256, 0, st>>>(n, e.P.as<double>(), jac, pt_off[i], e.z.as<double>() + 2 * i, meas_var, comp, e.x.as<double>(), K, KS);
            } else {
                k_ekf_jacobians<6><<<1, 128, 0, st>>>(1, e.x.as<double>(), e.off.as<int64_t>() + i, cam, jac, jac + 26, jac + 38);
                k_ekf_seq_gain<6><<<(n + 255) / 256, 256, 0, st>>>(n, e.P.as<double>(), jac, pt_off[i], e.z.as<double>() + 2 * i, meas_var, comp, e.x.as<double>(), K, KS);
            }
            k_ekf_seq_cov<<<dim3((n + 255) / 256, n), 256, 0, st>>>(n, e.P.as<double>(), K, KS);
            k_ekf_quat_prepare<<<1, 32, 0, st>>>(e.x.as<double>(), e.aux.as<double>());
            k_ekf_quat_columns<<<(n + 255) / 256, 256, 0, st>>>(n, e.P.as<double>(), e.aux.as<double>(), e.tmp.as<double>());
            k_ekf_quat_write<<<(n + 255) / 256, 256, 0, st>>>(n, e.P.as<double>(), e.aux.as<double>(), e.tmp.as<double>());
            e.launches += 6;
        }
    }
    EKF_CUDA(cudaStreamSynchronize(st));
    EKF_CUDA(cudaGetLastError());
    return SRK_OK;
}

int add_points_resident(Ekf& e, int64_t k, int s, const double* x_new, const double* Jy, const double* Qnew, int diag_only) {
    if (e.n <= 0) { g_ekf_error = "srk_ekf_add_points_resident before srk_ekf_set_state"; return SRK_E_NOT_BOUND; }
    if (k <= 0 || (s != 3 && s != 6) || x_new == nullptr || Jy == nullptr || Qnew == nullptr) { g_ekf_error = "bad add-points arguments (s must be 3 or 6)"; return SRK_E_INVALID_ARG; }
    const int64_t n = e.n, n2 = e.n + k * s;
    if (n2 > 60000) { g_ekf_error = "state too large"; return SRK_E_TOO_LARGE; }
    EKF_CUDA(cudaSetDevice(e.device));
    cudaStream_t st = e.st;
    EKF_CUDA(e.Pgrow.ensure(sizeof(double) * (size_t)n2 * n2 + 64)); EKF_CUDA(e.xgrow.ensure(sizeof(double) * (size_t)n2));
    EKF_CUDA(e.small.ensure(sizeof(double) * (3 * kCam * kCam + kCam)));
    EKF_CUDA(e.grow_in.ensure(sizeof(double) * (size_t)k * s * (7 + s)));
    double* dJy = e.grow_in.as<double>(); double* dQ = dJy + (size_t)k * s * 7;
    EKF_CUDA(cudaMemcpyAsync(dJy, Jy, sizeof(double) * (size_t)k * s * 7, cudaMemcpyHostToDevice, st));
    EKF_CUDA(cudaMemcpyAsync(dQ, Qnew, sizeof(double) * (size_t)k * s * s, cudaMemcpyHostToDevice, st));
    EKF_CUDA(cudaMemcpy2DAsync(e.Pgrow.p, sizeof(double) * (size_t)n2, e.P.p, sizeof(double) * (size_t)n, sizeof(double) * (size_t)n, (size_t)n, cudaMemcpyDeviceToDevice, st));
    EKF_CUDA(cudaMemcpyAsync(e.xgrow.p, e.x.p, sizeof(double) * (size_t)n, cudaMemcpyDeviceToDevice, st));
    EKF_CUDA(cudaMemcpyAsync(e.xgrow.as<double>() + n, x_new, sizeof(double) * (size_t)k * s, cudaMemcpyHostToDevice, st));
    {
        EScope sc(e, E_STATE);
        dim3 grid((unsigned)((n2 + 255) / 256 < 64 ? (n2 + 255) / 256 : 64), (unsigned)(k * s));
        k_ekf_grow<<<grid, 256, 0, st>>>((int)n, (int)n2, (int)k, s, e.P.as<double>(), e.Pgrow.as<double>(), dJy, dQ, diag_only);
        e.launches += 1;
    }
    EKF_CUDA(cudaStreamSynchronize(st));
    EKF_CUDA(cudaGetLastError());
    std::swap(e.P.p, e.Pgrow.p); std::swap(e.P.cap, e.Pgrow.cap);
    std::swap(e.x.p, e.xgrow.p); std::swap(e.x.cap, e.xgrow.cap);
    e.n = n2;
    return SRK_OK;
}

extern "C" {

int srk_ekf_create(void** h, int device) {
    if (h == nullptr) return SRK_E_INVALID_ARG;
    *h = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) { cudaGetLastError(); g_ekf_error = "no CUDA device (there is no CPU fallback)"; return SRK_E_NO_DEVICE; }
    if (device < 0 || device >= count) { g_ekf_error = "device id out of range"; return SRK_E_NO_DEVICE; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major != 10) { g_ekf_error = "this library holds sm_100a code only"; return SRK_E_NO_DEVICE; }
    cudaSetDevice(device);
    Ekf* e = new Ekf();
    e->device = device;
    if (cudaStreamCreateWithFlags(&e->own, cudaStreamNonBlocking) != cudaSuccess) { delete e; return SRK_E_CUDA; }
    e->st = e->own;
    *h = e;
    return SRK_OK;
}
void srk_ekf_destroy(void* h) {
    if (h == nullptr) return;
    Ekf* e = (Ekf*)h;
    cudaSetDevice(e->device);
    ekf_resolve(*e);
    if (e->hi != nullptr) cudaStreamDestroy(e->hi);
    for (cudaEvent_t ev : e->evs) cudaEventDestroy(ev);
    if (e->own != nullptr) cudaStreamDestroy(e->own);
    delete e;
}
int srk_ekf_set_stream(void* h, void* cuda_stream) {
    if (h == nullptr) return SRK_E_INVALID_ARG;
    Ekf* e = (Ekf*)h;
    cudaStreamSynchronize(e->st);
    e->st = cuda_stream != nullptr ? (cudaStream_t)cuda_stream : e->own;
    return SRK_OK;
}
int srk_ekf_set_state(void* h, int64_t n, const double* P, const double* x) {
    if (h == nullptr || P == nullptr || x == nullptr || n < kCam || n > 60000) { g_ekf_error = "bad state arguments"; return SRK_E_INVALID_ARG; }
    Ekf& e = *(Ekf*)h;
    EKF_CUDA(cudaSetDevice(e.device));
    EKF_CUDA(e.P.ensure(sizeof(double) * (size_t)n * n + 64)); EKF_CUDA(e.x.ensure(sizeof(double) * (size_t)n));
    EKF_CUDA(cudaMemcpyAsync(e.P.p, P, sizeof(double) * (size_t)n * n, cudaMemcpyHostToDevice, e.st));
    EKF_CUDA(cudaMemcpyAsync(e.x.p, x, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, e.st));
    EKF_CUDA(cudaStreamSynchronize(e.st));
    e.n = n;
    return SRK_OK;
}
int srk_ekf_get_state(void* h, double* P, double* x) {
    if (h == nullptr) return SRK_E_INVALID_ARG;
    Ekf& e = *(Ekf*)h;
    if (e.n <= 0) { g_ekf_error = "no resident state"; return SRK_E_NOT_BOUND; }
    EKF_CUDA(cudaSetDevice(e.device));
    if (P != nullptr) EKF_CUDA(cudaMemcpyAsync(P, e.P.p, sizeof(double) * (size_t)e.n * e.n, cudaMemcpyDeviceToHost, e.st));
    if (x != nullptr) EKF_CUDA(cudaMemcpyAsync(x, e.x.p, sizeof(double) * (size_t)e.n, cudaMemcpyDeviceToHost, e.st));
    EKF_CUDA(cudaStreamSynchronize(e.st));
    return SRK_OK;
}
int srk_ekf_predict_resident(void* h, const double* F13, const double* GQGt13, const double* cam_state_new) {
    if (h == nullptr) return SRK_E_INVALID_ARG;
    return predict_resident(*(Ekf*)h, F13, GQGt13, cam_state_new);
}
int srk_ekf_measurement_jacobians_resident(void* h, int64_t m, const int64_t* pt_off, int32_t s, const srk_ekf_camera* camera, double* Hcam, double* Hpt, double* h_pred) {
    if (h == nullptr) return SRK_E_INVALID_ARG;
    return measurement_jacobians(*(Ekf*)h, m, pt_off, s, camera, Hcam, Hpt, h_pred);
}
int srk_ekf_ransac_consensus_resident(void* h, int64_t m, const double* Hcam, const double* Hpt, const int64_t* pt_off, int32_t s, const double* z, double meas_var,
                                      const srk_ekf_camera* camera, double max_divergence_pix, int32_t* support, int32_t* best, unsigned char* best_inliers) {
    if (h == nullptr) return SRK_E_INVALID_ARG;
    return ransac_consensus(*(Ekf*)h, m, Hcam, Hpt, pt_off, s, z, meas_var, camera, max_divergence_pix, support, best, best_inliers);
}
int srk_ekf_add_points_resident(void* h, int64_t k, int32_t s, const double* x_new, const double* Jy, const double* Qnew, int32_t diag_only) {
    if (h == nullptr) return SRK_E_INVALID_ARG;
    return add_points_resident(*(Ekf*)h, k, s, x_new, Jy, Qnew, diag_only);
}
int srk_ekf_projected_covariances_resident(void* h, int64_t m, const double* Hcam, const double* Hpt, const int64_t* pt_off, int32_t s, double* cov) {
    if (h == nullptr) return SRK_E_INVALID_ARG;
    return projected_covariances(*(Ekf*)h, m, Hcam, Hpt, pt_off, s, cov);
}
int srk_ekf_sequential_update_resident(void* h, int64_t m, const int64_t* pt_off, int32_t s, const double* z, const srk_ekf_camera* camera, double meas_var,
                                       int32_t per_component) {
    if (h == nullptr) return SRK_E_INVALID_ARG;
    return sequential_update(*(Ekf*)h, m, pt_off, s, z, camera, meas_var, per_component);
}
int srk_ekf_state_size(void* h, int64_t* n) {
    if (h == nullptr || n == nullptr) return SRK_E_INVALID_ARG;
    *n = ((Ekf*)h)->n;
    return SRK_OK;
}
int srk_ekf_update_resident(void* h, int64_t m, const double* Hcam, const double* Hpt, const int64_t* pt_off, int32_t s, const double* z, const double* h_pred,
                            double meas_var, int32_t* info) {
    if (h == nullptr) return SRK_E_INVALID_ARG;
    return update_resident(*(Ekf*)h, m, Hcam, Hpt, pt_off, s, z, h_pred, meas_var, info);
}
int srk_ekf_update(void* h, int64_t n, int64_t m, double* P, double* x, const double* Hcam, const double* Hpt, const int64_t* pt_off, int32_t s, const double* z,
                   const double* h_pred, double meas_var) {
    int rc = srk_ekf_set_state(h, n, P, x);
    if (rc != SRK_OK) return rc;
    int32_t info = 0;
    rc = srk_ekf_update_resident(h, m, Hcam, Hpt, pt_off, s, z, h_pred, meas_var, &info);
    if (rc != SRK_OK) return rc;              // SRK_E_NOT_POSDEF included: the caller's P and x are not overwritten
    return srk_ekf_get_state(h, P, x);
}
int srk_ekf_predict(void* h, int64_t n, double* P, const double* F13, const double* GQGt13) {
    if (h == nullptr || P == nullptr) return SRK_E_INVALID_ARG;
    std::vector<double> x((size_t)n, 0.0);
    int rc = srk_ekf_set_state(h, n, P, x.data());
    if (rc != SRK_OK) return rc;
    rc = srk_ekf_predict_resident(h, F13, GQGt13, nullptr);
    if (rc != SRK_OK) return rc;
    return srk_ekf_get_state(h, P, nullptr);
}
int srk_ekf_set_timing(void* h, int enabled) {
    if (h == nullptr) return SRK_E_INVALID_ARG;
    Ekf& e = *(Ekf*)h;
    ekf_resolve(e);
    e.timing = enabled != 0;
    for (int f = 0; f < E_COUNT; ++f) { e.total[f] = 0; e.count[f] = 0; }
    return SRK_OK;
}
int srk_ekf_get_timing(void* h, const char* name, double* ms_total, int64_t* count) {
    if (h == nullptr || name == nullptr) return SRK_E_INVALID_ARG;
    Ekf& e = *(Ekf*)h;
    ekf_resolve(e);
    for (int f = 0; f < E_COUNT; ++f)
        if (std::strcmp(name, kEkfFam[f]) == 0) { if (ms_total) *ms_total = e.total[f]; if (count) *count = e.count[f]; return SRK_OK; }
    return SRK_E_INVALID_ARG;
}
int64_t srk_ekf_launches(void* h) { return h == nullptr ? 0 : ((Ekf*)h)->launches; }

}  // extern "C"
