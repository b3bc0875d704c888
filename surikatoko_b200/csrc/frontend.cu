// suriko-b200 — input front end of the two BA demos behind include/srk/frontend_c_api.h.
//
//   k_triangulate            Triangulate3DPointByLeastSquares (obs-geom.cpp:679-727), one thread per track: the 2k x 3 system
//                            is never formed; every corner contributes two rows that are folded into a 3x3 upper-triangular
//                            R and the rotated right-hand side by Givens rotations (a QR factorisation built row by row, the same
//                            least-squares minimiser as the reference's colPivHouseholderQr for a full-rank system), then one
//                            back substitution.  HBM traffic: 20 B per corner + the 96-byte projection matrix of its frame (L2).
//   srk_decompose_proj_mat   DecomposeProjMat (obs-geom.cpp:606-677), host: 3x3 inverses, a 3x3 Cholesky, a triangular inverse.
//   srk_read_matrix_from_file ReadMatrixFromFile (mat-serialization.cpp:12-87), host.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <mutex>
#include <sstream>
#include <string>
#include <vector>

#include <cuda_runtime.h>
#include "../../include/srk/frontend_c_api.h"

extern "C" void srk_internal_set_error(const char* s);   // engine.cu: the string behind srk_last_error()

namespace {

__device__ __forceinline__ void givens_fold(double (&R)[6], double (&qb)[3], double a0, double a1, double a2, double b) {
    // R = [r00 r01 r02; 0 r11 r12; 0 0 r22] packed as {r00, r01, r02, r11, r12, r22}; annihilate the new row (a0 a1 a2 | b) against it
    double row[3] = {a0, a1, a2};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const int d = c == 0 ? 0 : (c == 1 ? 3 : 5);       // index of R(c,c)
        const double x = R[d], y = row[c];
        if (y == 0.0) continue;
        const double h = sqrt(x * x + y * y);              // entries are O(1) in f0 units: no overflow guard (hypot) needed
        const double ih = 1.0 / h;
        const double cs = x * ih, sn = y * ih;
        R[d] = h;
#pragma unroll
        for (int k = c + 1; k < 3; ++k) {
            const int idx = d + (k - c);
            const double rv = R[idx], av = row[k];
            R[idx] = cs * rv + sn * av;
            row[k] = cs * av - sn * rv;
        }
        const double qv = qb[c];
        qb[c] = cs * qv + sn * b;
        b = cs * b - sn * qv;
        row[c] = 0.0;
    }
}

__global__ void __launch_bounds__(128) k_triangulate(int64_t n_tracks, const int64_t* __restrict__ track_begin, const int32_t* __restrict__ obs_frame,
                                                     const double* __restrict__ obs_xy, const double* __restrict__ proj, int n_frames, double f0,
                                                     double* __restrict__ out, int* __restrict__ err) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tracks) return;
    const int64_t b = track_begin[t], e = track_begin[t + 1];
    if (e - b < 2) { atomicOr(err, 1); return; }
    double R[6] = {0, 0, 0, 0, 0, 0}, qb[3] = {0, 0, 0};
    for (int64_t o = b; o < e; ++o) {
        const double x = obs_xy[2 * o], y = obs_xy[2 * o + 1];
        const int fr = obs_frame[o];
        if (fr < 0 || fr >= n_frames) { atomicOr(err, 2); return; }
        const double* P = proj + (size_t)fr * 12;                 // column-major 3x4: P(r,c) = P[c*3 + r]
        // obs-geom.cpp:694-709
        givens_fold(R, qb, x * P[2] - f0 * P[0], x * P[5] - f0 * P[3], x * P[8] - f0 * P[6], -(x * P[11] - f0 * P[9]));
        givens_fold(R, qb, y * P[2] - f0 * P[1], y * P[5] - f0 * P[4], y * P[8] - f0 * P[7], -(y * P[11] - f0 * P[10]));
    }
    const double z = qb[2] / R[5];
    const double yv = (qb[1] - R[4] * z) / R[3];
    const double xv = (qb[0] - R[1] * yv - R[2] * z) / R[0];
    out[3 * t] = xv; out[3 * t + 1] = yv; out[3 * t + 2] = z;
}

thread_local float g_last_kernel_ms = 0.f;   // CUDA-event time of the last k_triangulate launch on this thread (bench.py)

// grow-only device buffers + one stream, cached per process (the front end is called once per scene; the mutex serialises callers)
struct DevBuf {
    void* p = nullptr; size_t cap = 0;
    cudaError_t alloc(size_t bytes) {
        if (bytes < 16) bytes = 16;
        if (p != nullptr && bytes <= cap) return cudaSuccess;
        if (p != nullptr) { cudaFree(p); p = nullptr; cap = 0; }
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e == cudaSuccess) cap = bytes;
        return e;
    }
};
struct TriPool { int device = -1; DevBuf tb, fr, xy, pm, out, err; cudaStream_t st = nullptr; cudaEvent_t e0 = nullptr, e1 = nullptr; };
TriPool g_pool;
std::mutex g_pool_mutex;

#define FE_CUDA(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { srk_internal_set_error((std::string(#call) + ": " + cudaGetErrorString(e__)).c_str()); return SRK_E_CUDA; } } while (0)

// 3x3 helpers, column-major (r,c) -> [c*3 + r]
inline double det3(const double* m) {
    return m[0] * (m[4] * m[8] - m[7] * m[5]) - m[3] * (m[1] * m[8] - m[7] * m[2]) + m[6] * (m[1] * m[5] - m[4] * m[2]);
}
inline void inv3(const double* m, double* o) {
    const double d = det3(m), id = 1.0 / d;
    o[0] = (m[4] * m[8] - m[7] * m[5]) * id; o[3] = -(m[3] * m[8] - m[6] * m[5]) * id; o[6] = (m[3] * m[7] - m[6] * m[4]) * id;
    o[1] = -(m[1] * m[8] - m[7] * m[2]) * id; o[4] = (m[0] * m[8] - m[6] * m[2]) * id; o[7] = -(m[0] * m[7] - m[6] * m[1]) * id;
    o[2] = (m[1] * m[5] - m[4] * m[2]) * id; o[5] = -(m[0] * m[5] - m[3] * m[2]) * id; o[8] = (m[0] * m[4] - m[3] * m[1]) * id;
}
inline void mul3(const double* a, const double* b, double* c) {
    for (int cc = 0; cc < 3; ++cc) for (int r = 0; r < 3; ++r) c[cc * 3 + r] = a[0 * 3 + r] * b[cc * 3 + 0] + a[1 * 3 + r] * b[cc * 3 + 1] + a[2 * 3 + r] * b[cc * 3 + 2];
}
inline void tr3(const double* a, double* t) { for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) t[c * 3 + r] = a[r * 3 + c]; }

}  // namespace

extern "C" {

int srk_triangulate_tracks(int device, int64_t n_tracks, int64_t n_obs, int32_t n_frames, const int64_t* track_begin, const int32_t* obs_frame,
                           const double* obs_xy, const double* proj, double f0, double* points_out) {
    if (n_tracks < 0 || n_obs < 0 || n_frames <= 0 || track_begin == nullptr || proj == nullptr || (n_obs > 0 && (obs_frame == nullptr || obs_xy == nullptr)) ||
        (n_tracks > 0 && points_out == nullptr)) { srk_internal_set_error("null or negative-sized triangulation input"); return SRK_E_INVALID_ARG; }
    if (n_tracks == 0) return SRK_OK;
    if (track_begin[0] != 0 || track_begin[n_tracks] != n_obs) { srk_internal_set_error("track_begin must run from 0 to n_obs"); return SRK_E_INVALID_ARG; }
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0 || device < 0 || device >= count) { cudaGetLastError(); srk_internal_set_error("no CUDA device (there is no CPU fallback)"); return SRK_E_NO_DEVICE; }
    std::lock_guard<std::mutex> lock(g_pool_mutex);
    FE_CUDA(cudaSetDevice(device));
    TriPool& pl = g_pool;
    if (pl.device != device) {
        if (pl.device >= 0) {   // buffers of another device: drop them
            cudaSetDevice(pl.device);
            for (DevBuf* d : {&pl.tb, &pl.fr, &pl.xy, &pl.pm, &pl.out, &pl.err}) { if (d->p) cudaFree(d->p); d->p = nullptr; d->cap = 0; }
            if (pl.st) cudaStreamDestroy(pl.st);
            if (pl.e0) cudaEventDestroy(pl.e0);
            if (pl.e1) cudaEventDestroy(pl.e1);
            pl.st = nullptr; pl.e0 = pl.e1 = nullptr;
            cudaSetDevice(device);
        }
        pl.device = device;
    }
    if (pl.st == nullptr) { FE_CUDA(cudaStreamCreateWithFlags(&pl.st, cudaStreamNonBlocking)); FE_CUDA(cudaEventCreate(&pl.e0)); FE_CUDA(cudaEventCreate(&pl.e1)); }
    FE_CUDA(pl.tb.alloc(sizeof(int64_t) * (n_tracks + 1))); FE_CUDA(pl.fr.alloc(sizeof(int32_t) * n_obs)); FE_CUDA(pl.xy.alloc(sizeof(double) * 2 * n_obs));
    FE_CUDA(pl.pm.alloc(sizeof(double) * 12 * (size_t)n_frames)); FE_CUDA(pl.out.alloc(sizeof(double) * 3 * n_tracks)); FE_CUDA(pl.err.alloc(sizeof(int)));
    cudaStream_t st = pl.st;
    FE_CUDA(cudaMemcpyAsync(pl.tb.p, track_begin, sizeof(int64_t) * (n_tracks + 1), cudaMemcpyHostToDevice, st));
    FE_CUDA(cudaMemcpyAsync(pl.fr.p, obs_frame, sizeof(int32_t) * n_obs, cudaMemcpyHostToDevice, st));
    FE_CUDA(cudaMemcpyAsync(pl.xy.p, obs_xy, sizeof(double) * 2 * n_obs, cudaMemcpyHostToDevice, st));
    FE_CUDA(cudaMemcpyAsync(pl.pm.p, proj, sizeof(double) * 12 * (size_t)n_frames, cudaMemcpyHostToDevice, st));
    FE_CUDA(cudaMemsetAsync(pl.err.p, 0, sizeof(int), st));
    FE_CUDA(cudaEventRecord(pl.e0, st));
    k_triangulate<<<(unsigned)((n_tracks + 127) / 128), 128, 0, st>>>(n_tracks, (const int64_t*)pl.tb.p, (const int32_t*)pl.fr.p, (const double*)pl.xy.p,
                                                                   (const double*)pl.pm.p, n_frames, f0, (double*)pl.out.p, (int*)pl.err.p);
    FE_CUDA(cudaEventRecord(pl.e1, st));
    int h_err = 0;
    FE_CUDA(cudaMemcpyAsync(points_out, pl.out.p, sizeof(double) * 3 * n_tracks, cudaMemcpyDeviceToHost, st));
    FE_CUDA(cudaMemcpyAsync(&h_err, pl.err.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    FE_CUDA(cudaStreamSynchronize(st));
    FE_CUDA(cudaGetLastError());
    cudaEventElapsedTime(&g_last_kernel_ms, pl.e0, pl.e1);
    if (h_err & 2) { srk_internal_set_error("frame index out of range"); return SRK_E_INVALID_ARG; }
    if (h_err & 1) { srk_internal_set_error("Provide 2 or more projections of a 3D point (obs-geom.cpp:687)"); return SRK_E_INVALID_ARG; }
    return SRK_OK;
}

double srk_triangulate_last_kernel_ms(void) { return (double)g_last_kernel_ms; }

int srk_decompose_proj_mat(const double* P, double* scale_factor, double* K, double* direct_pose) {
    if (P == nullptr || scale_factor == nullptr || K == nullptr || direct_pose == nullptr) { srk_internal_set_error("null argument"); return SRK_E_INVALID_ARG; }
    double Q[9], q[3];
    for (int i = 0; i < 9; ++i) Q[i] = P[i];
    for (int i = 0; i < 3; ++i) q[i] = P[9 + i];
    int P_sign = 1;
    if (det3(Q) < 0) { P_sign = -1; for (double& v : Q) v = -v; for (double& v : q) v = -v; }   // R gets a positive determinant (:617-624)
    double Qi[9]; inv3(Q, Qi);
    double t[3];
    for (int r = 0; r < 3; ++r) t[r] = -(Qi[0 * 3 + r] * q[0] + Qi[1 * 3 + r] * q[1] + Qi[2 * 3 + r] * q[2]);   // t = -Q^-1 q (:627-628)
    double Qt[9], QQt[9], QQti[9];
    tr3(Q, Qt); mul3(Q, Qt, QQt); inv3(QQt, QQti);                                                             // (:631-634)
    // LLT of QQt^-1 (:637-643), lower L, column-major
    double L[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int j = 0; j < 3; ++j) {
        double d = QQti[j * 3 + j];
        for (int k = 0; k < j; ++k) d -= L[k * 3 + j] * L[k * 3 + j];
        if (!(d > 0.0)) return 1;
        L[j * 3 + j] = std::sqrt(d);
        for (int i = j + 1; i < 3; ++i) {
            double s = QQti[j * 3 + i];
            for (int k = 0; k < j; ++k) s -= L[k * 3 + i] * L[k * 3 + j];
            L[j * 3 + i] = s / L[j * 3 + j];
        }
    }
    double C[9]; tr3(L, C);                    // upper triangular (:646)
    double CQ[9], R[9];
    mul3(C, Q, CQ); tr3(CQ, R);                // R = (C Q)^T (:648)
    double Ci[9]; inv3(C, Ci);
    const double c_last = Ci[2 * 3 + 2];       // (:659-660)
    if (std::fabs(c_last) <= 1e-8) { srk_internal_set_error("det(P)<3 (obs-geom.cpp:661)"); return SRK_E_INVALID_ARG; }
    for (int i = 0; i < 9; ++i) K[i] = Ci[i] * (1 / c_last);
    *scale_factor = P_sign * c_last;
    for (int i = 0; i < 3; ++i) direct_pose[i] = t[i];
    for (int i = 0; i < 9; ++i) direct_pose[3 + i] = R[i];
    return SRK_OK;
}

int srk_read_matrix_from_file(const char* path, char delimiter, double* data, int64_t cap, int64_t* rows, int64_t* cols) {
    if (path == nullptr || rows == nullptr || cols == nullptr) { srk_internal_set_error("null argument"); return SRK_E_INVALID_ARG; }
    std::ifstream fs(path);
    if (!fs) { srk_internal_set_error((std::string("Can't open file ") + path).c_str()); return SRK_E_INVALID_ARG; }
    const char delims[2] = {delimiter, 0};
    std::string line;
    int64_t num_rows = 0, num_cols = -1, count = 0;
    while (std::getline(fs, line)) {
        int64_t cur = 0;
        std::vector<char> buf(line.begin(), line.end()); buf.push_back(0);
        char* save = nullptr;
        for (char* tok = strtok_r(buf.data(), delims, &save); tok != nullptr; tok = strtok_r(nullptr, delims, &save)) {
            std::istringstream iss{std::string(tok)};
            double num; iss >> num;
            if (iss.fail() || !iss.eof()) {   // the whole token must be a number (:55-64)
                srk_internal_set_error(("Can't parse number (" + std::string(tok) + ") on line " + std::to_string(num_rows)).c_str());
                return SRK_E_INVALID_ARG;
            }
            if (data != nullptr) { if (count >= cap) { srk_internal_set_error("matrix does not fit the buffer"); return SRK_E_TOO_LARGE; } data[count] = num; }
            ++count; ++cur;
        }
        if (num_cols == -1) num_cols = cur;
        else if (num_cols != cur) {
            srk_internal_set_error(("Data has inconsistent number of columns, row(0).columns=" + std::to_string(num_cols) + ", row(" + std::to_string(num_rows) +
                                    ").columns=" + std::to_string(cur)).c_str());
            return SRK_E_INVALID_ARG;
        }
        ++num_rows;
    }
    *rows = num_rows; *cols = num_cols == -1 ? 0 : num_cols;
    return SRK_OK;
}

}  // extern "C"
