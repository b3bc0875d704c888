// suriko-b200 — shared device helpers for the BA kernels (sm_100a).
// Per-observation formulas follow /root/reference/cpp_impl/suriko-engine/src/bundle-adj-kanatani.cpp ("BA.cpp").
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace srk {

// cudaFuncSetAttribute (the > 48 KB dynamic shared-memory opt-in) is a PER-DEVICE property: a host that opens handles on several GPUs
// in one process must set it once on each.  `static PerDeviceOnce once; if (once.first()) { ... }` replaces a process-wide flag.
struct PerDeviceOnce {
    unsigned long long seen[4] = {0, 0, 0, 0};   // 256 device ordinals
    bool first() {
        int d = 0;
        if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= 256) return true;
        const unsigned long long bit = 1ull << (d & 63);
        unsigned long long old = __atomic_fetch_or(&seen[d >> 6], bit, __ATOMIC_ACQ_REL);
        return (old & bit) == 0;
    }
    void forget() { int d = 0; if (cudaGetDevice(&d) == cudaSuccess && d >= 0 && d < 256) __atomic_fetch_and(&seen[d >> 6], ~(1ull << (d & 63)), __ATOMIC_ACQ_REL); }
};

constexpr int kV = 10;           // frame variables [fx fy u0 v0 | Tx Ty Tz | Wx Wy Wz]   (BA.h:102-118)
constexpr int kCamStride = 48;   // doubles per derived-camera record (384 B, 128-B aligned)

// Derived camera record, rebuilt whenever a pose changes (k_cam_prep).  All 3x3 are column-major (r,c)->[c*3+r].
//  [0..8] R  [9..11] T  [12..20] K  [21..29] KR=K*R  [30..32] Td=-R^T*T  [33..35] rot1  [36..38] rot2  [39..41] rot3
//  [42] 1/fx [43] u0/(f0*fx) [44] 1/fy [45] v0/(f0*fy) [46] 1/f0 [47] pad   -- the per-camera quotients of BA.cpp:1463-1474,
//  evaluated once per camera instead of once per observation (same operands, same IEEE division, same bits).
enum { CD_R = 0, CD_T = 9, CD_K = 12, CD_KR = 21, CD_TD = 30, CD_ROT1 = 33, CD_ROT2 = 36, CD_ROT3 = 39, CD_IFX = 42, CD_CU = 43, CD_IFY = 44, CD_CV = 45, CD_IF0 = 46 };

// Gauge-reduced index of frame variable (cam, a): frame 0 keeps [fx fy u0 v0], frame 1 drops T[unity], frames >= 2 keep
// all 10 (quirk Q13; BA.cpp:539-563, :1780-1823).  Returns -1 for the 7 removed variables.
__host__ __device__ __forceinline__ int red_index(int cam, int a, int unity) {
    if (cam == 0) return a < 4 ? a : -1;
    if (cam == 1) {
        if (a < 4) return 4 + a;
        int t = a - 4;
        if (t == unity) return -1;
        return 4 + a - (t > unity ? 1 : 0);
    }
    return cam * kV - 7 + a;
}

// Where the Schur accumulation kernels put their results.  Dense mode (blocks == nullptr): the gauge-reduced column-major
// matrix S (lower block triangle) and rhs of the dense Cholesky path.  Block-sparse mode: 10x10 blocks (row-major, rows = the
// variables of the LARGER camera index) addressed through an open-addressing hash of the camera pair, and a right-hand side
// in full frame-variable space [10M]; the 7 gauge variables are skipped (their rows / columns stay zero).
struct SchurSink {
    double* S; int64_t ld; double* rhs; int unity;
    const unsigned long long* hkeys; const int* hids; unsigned hmask; double* blocks;
};
constexpr unsigned long long kHashEmpty = 0xffffffffffffffffull;
__host__ __device__ __forceinline__ unsigned hash_pair(unsigned long long key, unsigned mask) {
    key ^= key >> 33; key *= 0xff51afd7ed558ccdull; key ^= key >> 33; key *= 0xc4ceb9fe1a85ec53ull; key ^= key >> 33;
    return (unsigned)key & mask;
}
// block id of the camera pair (cam_i >= cam_l), or -1 (dense mode / pair not in the structure)
__device__ __forceinline__ int sink_block_id(const SchurSink& s, int cam_i, int cam_l) {
    if (s.blocks == nullptr) return -1;
    const unsigned long long key = ((unsigned long long)(unsigned)cam_i << 32) | (unsigned)cam_l;
    unsigned h = hash_pair(key, s.hmask);
    for (unsigned probe = 0; probe <= s.hmask; ++probe) {
        const unsigned long long k = s.hkeys[h];
        if (k == key) return s.hids[h];
        if (k == kHashEmpty) return -1;
        h = (h + 1) & s.hmask;
    }
    return -1;
}
__device__ __forceinline__ void sink_add(const SchurSink& s, int blk, int cam_i, int a, int cam_l, int b, double v) {
    const int row = red_index(cam_i, a, s.unity), col = red_index(cam_l, b, s.unity);
    if (row < 0 || col < 0) return;
    if (s.blocks == nullptr) atomicAdd(&s.S[(size_t)col * s.ld + row], v);
    else if (blk >= 0) atomicAdd(&s.blocks[(size_t)blk * 100 + a * 10 + b], v);
}
__device__ __forceinline__ void sink_add_rhs(const SchurSink& s, int cam, int a, double v) {
    const int r = red_index(cam, a, s.unity);
    if (r < 0) return;
    if (s.blocks == nullptr) atomicAdd(&s.rhs[r], v);
    else atomicAdd(&s.rhs[(size_t)cam * 10 + a], v);
}

// pqr = K * (R*X + T)   (BA.cpp:469-470), natural left-to-right coefficient order.
__device__ __forceinline__ void project_pqr(const double* cd, double X0, double X1, double X2, double& p, double& q, double& r) {
    const double* R = cd + CD_R; const double* T = cd + CD_T; const double* K = cd + CD_K;
    double c0 = R[0] * X0 + R[3] * X1 + R[6] * X2 + T[0];
    double c1 = R[1] * X0 + R[4] * X1 + R[7] * X2 + T[1];
    double c2 = R[2] * X0 + R[5] * X1 + R[8] * X2 + T[2];
    p = K[0] * c0 + K[3] * c1 + K[6] * c2;
    q = K[1] * c0 + K[4] * c1 + K[7] * c2;
    r = K[2] * c0 + K[5] * c1 + K[8] * c2;
}

// One observation: residual rho = (p/r - x/f0, q/r - y/f0) (BA.cpp:475-479).  xs = x/f0, ys = y/f0 are formed once at bind time
// (k_prep_obs): the pixel and f0 never change during a solve, and the quotient is the same IEEE division.
__device__ __forceinline__ void obs_residual(const double* cd, double X0, double X1, double X2, double xs, double ys,
                                             double& rx, double& ry) {
    double p, q, r;
    project_pqr(cd, X0, X1, X2, p, q, r);
    rx = p / r - xs;
    ry = q / r - ys;
}

// One observation: residual + Jacobian rows J_a = (r*p_a - p*r_a, r*q_a - q*r_a) / r^2 for the 3 point variables
// (BA.cpp:1450-1455) and the 10 frame variables (BA.cpp:1457-1525, quirk Q3: f0 verbatim).  With these,
// formula 8 (BA.cpp:1528-1537) is 2*rho.J_a and formula 9 (BA.cpp:1540-1549) is 2*J_a.J_b.
// jp[v*2+comp], jc[a*2+comp].
__device__ __forceinline__ void obs_jacobian(const double* cd, double X0, double X1, double X2, double xs, double ys,
                                             double& rx, double& ry, double* __restrict__ jp, double* __restrict__ jc) {
    double p, q, r;
    project_pqr(cd, X0, X1, X2, p, q, r);
    rx = p / r - xs;
    ry = q / r - ys;
    const double ir2 = 1.0 / (r * r);
    const double* KR = cd + CD_KR;
#pragma unroll
    for (int v = 0; v < 3; ++v) {  // column v of P = K*[R|T]
        double pa = KR[v * 3 + 0], qa = KR[v * 3 + 1], ra = KR[v * 3 + 2];
        jp[v * 2 + 0] = (r * pa - p * ra) * ir2;
        jp[v * 2 + 1] = (r * qa - q * ra) * ir2;
    }
    // intrinsics: only one of (p_a, q_a) is non-zero and r_a = 0
    double pfx = cd[CD_IFX] * p - cd[CD_CU] * r;     // (1/fx) p - u0/(f0 fx) r
    double qfy = cd[CD_IFY] * q - cd[CD_CV] * r;     // (1/fy) q - v0/(f0 fy) r
    double pu0 = cd[CD_IF0] * r;                     // (1/f0) r
    jc[0] = (r * pfx) * ir2; jc[1] = 0.0;
    jc[2] = 0.0;             jc[3] = (r * qfy) * ir2;
    jc[4] = (r * pu0) * ir2; jc[5] = 0.0;
    jc[6] = 0.0;             jc[7] = (r * pu0) * ir2;
    const double* rot1 = cd + CD_ROT1; const double* rot2 = cd + CD_ROT2; const double* rot3 = cd + CD_ROT3;
#pragma unroll
    for (int t = 0; t < 3; ++t) {  // direct translation
        double pa = -rot1[t], qa = -rot2[t], ra = -rot3[t];
        jc[(4 + t) * 2 + 0] = (r * pa - p * ra) * ir2;
        jc[(4 + t) * 2 + 1] = (r * qa - q * ra) * ir2;
    }
    const double* Td = cd + CD_TD;
    double d0 = X0 - Td[0], d1 = X1 - Td[1], d2 = X2 - Td[2];
    double wp[3] = {rot1[1] * d2 - rot1[2] * d1, rot1[2] * d0 - rot1[0] * d2, rot1[0] * d1 - rot1[1] * d0};
    double wq[3] = {rot2[1] * d2 - rot2[2] * d1, rot2[2] * d0 - rot2[0] * d2, rot2[0] * d1 - rot2[1] * d0};
    double wr[3] = {rot3[1] * d2 - rot3[2] * d1, rot3[2] * d0 - rot3[0] * d2, rot3[0] * d1 - rot3[1] * d0};
#pragma unroll
    for (int t = 0; t < 3; ++t) {  // direct axis-angle
        jc[(7 + t) * 2 + 0] = (r * wp[t] - p * wr[t]) * ir2;
        jc[(7 + t) * 2 + 1] = (r * wq[t] - q * wr[t]) * ir2;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Per-CTA camera table.  Observations are point-major, so the 32 lanes of a warp touch ~10 different cameras; reading the
// 384-byte records straight from global memory costs one L1 wavefront per distinct camera per load (ncu: the LSU data pipe
// at 96 % with HBM at 8-33 %).  A CTA therefore collects the distinct cameras of its chunk of observations in a small
// shared-memory hash set, copies their records once (coalesced) into shared memory with an ODD stride, so that records of
// different cameras start in different banks and same-camera lanes broadcast, and every per-observation read becomes a
// conflict-free shared-memory load.  Cameras beyond kCamTabSlots per chunk (scattered visibility) are read from global memory.
constexpr int kCamTabSlots = 24;
constexpr int kCamTabHash = 64;
constexpr int kCamRecPad = kCamStride + 1;
struct CamTable {
    int key[kCamTabHash];
    int slot[kCamTabHash];
    int cam_of_slot[kCamTabSlots];
    int count;
    double rec[kCamTabSlots * kCamRecPad];
};
__device__ __forceinline__ void cam_table_reset(CamTable& t) {
    for (int i = threadIdx.x; i < kCamTabHash; i += blockDim.x) t.key[i] = -1;
    if (threadIdx.x == 0) t.count = 0;
}
// returns the hash position of `cam` (its slot is valid after the next barrier), or -1 when the hash set is full
__device__ __forceinline__ int cam_table_insert(CamTable& t, int cam) {
    unsigned h = ((unsigned)cam * 2654435761u) >> 26;   // 6 bits
    for (int probe = 0; probe < kCamTabHash; ++probe) {
        const int prev = atomicCAS(&t.key[h], -1, cam);
        if (prev == -1) {
            const int s = atomicAdd(&t.count, 1);
            if (s < kCamTabSlots) { t.slot[h] = s; t.cam_of_slot[s] = cam; } else t.slot[h] = -1;
            return (int)h;
        }
        if (prev == cam) return (int)h;
        h = (h + 1) & (kCamTabHash - 1);
    }
    return -1;
}
// after a barrier: hash position of a camera that was inserted before it (or -1: the set was full)
__device__ __forceinline__ int cam_table_find(const CamTable& t, int cam) {
    unsigned h = ((unsigned)cam * 2654435761u) >> 26;
    for (int probe = 0; probe < kCamTabHash; ++probe) {
        const int k = t.key[h];
        if (k == cam) return (int)h;
        if (k == -1) return -1;
        h = (h + 1) & (kCamTabHash - 1);
    }
    return -1;
}
// after a barrier: copies the records of the collected cameras (coalesced), caller issues the next barrier
__device__ __forceinline__ void cam_table_stage(CamTable& t, const double* __restrict__ camd) {
    const int n = min(t.count, kCamTabSlots) * kCamStride;
    for (int e = threadIdx.x; e < n; e += blockDim.x) {
        const int s = e / kCamStride, f = e - s * kCamStride;
        t.rec[s * kCamRecPad + f] = camd[(size_t)t.cam_of_slot[s] * kCamStride + f];
    }
}
__device__ __forceinline__ const double* cam_table_record(const CamTable& t, int hpos, int cam, const double* camd) {
    const int s = hpos >= 0 ? t.slot[hpos] : -1;
    return s >= 0 ? t.rec + s * kCamRecPad : camd + (size_t)cam * kCamStride;
}

// Damped 3x3 point block -> cofactor inverse with Eigen's computeInverseAndDetWithCheck rule |det| > 1e-12
// (BA.cpp:1873-1881, quirk Q5).  e = {E00,E01,E02,E11,E12,E22} (undamped, symmetric); inv gets the same packing.
__device__ __forceinline__ bool point_block_inverse(const double* e, double c, double* inv) {
    double m00 = e[0] * (1.0 + c), m11 = e[3] * (1.0 + c), m22 = e[5] * (1.0 + c);
    double m01 = e[1], m02 = e[2], m12 = e[4];
    double c00 = m11 * m22 - m12 * m12;
    double c10 = m12 * m02 - m01 * m22;   // cofactor(1,0) = m(2,1)*m(0,2) - m(2,2)*m(0,1)
    double c20 = m01 * m12 - m02 * m11;   // cofactor(2,0) = m(0,1)*m(1,2) - m(0,2)*m(1,1)
    double det = (c00 * m00 + c10 * m01) + c20 * m02;
    if (!(fabs(det) > 1e-12)) return false;
    double id = 1.0 / det;
    inv[0] = c00 * id; inv[1] = c10 * id; inv[2] = c20 * id;
    inv[3] = (m22 * m00 - m02 * m02) * id;
    inv[4] = (m02 * m01 - m00 * m12) * id;
    inv[5] = (m00 * m11 - m01 * m01) * id;
    return true;
}

}  // namespace srk
