// suriko-b200 — K3a: dense FP64 Cholesky solve of the reduced camera system S*df = rhs on sm_100a.
//
// Replaces Eigen's householderQr().solve at BA.cpp:1911.  S is symmetric positive definite (a Schur complement of the
// damped Gauss-Newton Hessian), so LL^T is the appropriate factorisation.  FP64 tensor work on Blackwell is DMMA:
// mma.sync.aligned.m8n8k4.f64 (tcgen05.mma has no f64 kind, TMEM accumulators are not available for doubles).
//
// Right-looking blocked algorithm on the column-major lower triangle (leading dimension ld, a multiple of 8):
//   panel width PB = 256, factored as four 64-column sub-panels:
//     k_potrf64_inv : 64x64 diagonal block -> L_kk in place and L_kk^-1 into dinv[kb] (kept for the substitutions)
//     k_panel_solve : rows below, X = A * L_kk^-T as a small GEMM against the inverted block (no serial TRSM)
//     k_syrk_dmma   : C -= X*X^T restricted to the rest of the current 256-panel (K = 64)
//   then ONE trailing update per panel with K = 256:
//     k_syrk_dmma   : 128x128 tiles, 3-stage cp.async pipeline over K chunks of 16, DMMA m8n8k4 with 32x64 register
//                     tiles per warp.  4x less trailing-matrix traffic than a 64-wide right-looking sweep.
//   k_trsv_coop     : forward / backward substitution as ONE cooperative kernel: per 64-block a GEMV with the stored
//                     L_kk^-1, a grid-wide panel GEMV for the rows below / above, one grid.sync per block.
#include <mutex>
#include <unordered_map>
#include <cooperative_groups.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include "kernels.h"

namespace cg = cooperative_groups;

namespace srk {

constexpr int NB = 64;     // diagonal block / sub-panel width
constexpr int PB = 256;    // panel width of the trailing update

// ---------------------------------------------------------------------------------------------------------------------
// 64x64 diagonal block.  Thread (r = tid & 63, grp = tid >> 6) keeps row r of the block, columns c == grp (mod 4), in 16
// registers, and the same slice of row r of W, which starts as the identity and receives the same row operations: after the
// 64 elimination steps T holds L up to the deferred column scalings and W the matching L^-1 (L(r,c) = T(r,c) rs_c,
// Linv(r,c) = W(r,c) rs_r, rs = 1/sqrt(pivot)).  A step publishes column j of T and row j of W through shared memory (double
// buffered, ONE barrier) and applies one DFMA per owned element.
// The step is bound by the ~200-instruction critical path of a single warp, not by FP64 throughput (measured with
// tools/potrf_micro.cu: 1870 -> 667 cycles per step), hence: the loop nest is (i unrolled) x (g rolled) so that every register
// index and almost every T-vs-W decision is static; the reciprocal of the pivot is an FP32 seed + Newton steps instead of the
// library division; rsqrt is taken once at the end.  info: 1-based index of the first non-positive pivot.
__device__ __forceinline__ double fast_rcp(double d) {   // 1/d to ~1 ulp
    double x = (double)__frcp_rn((float)d);
    x = x * (2.0 - d * x);
    x = x * (2.0 - d * x);
    x = x + x * (1.0 - d * x);
    return x;
}
template <bool CG>
__device__ void potrf64_inv_dev(int n, int k0, double* A, int64_t ld, double* dinv, int* info, unsigned char* F, int nblk) {
    constexpr int G = 4, E = NB / G;
    __shared__ double col[2][NB];
    __shared__ double wrow[2][NB];
    __shared__ double dv[NB];
    const int nb = min(NB, n - k0);
    const int tid = threadIdx.x;
    const int r = tid & 63, grp = tid >> 6;
    double t[E], w[E];
#pragma unroll
    for (int i = 0; i < E; ++i) {
        const int c = grp + G * i;
        t[i] = (r < nb && c < nb && r >= c) ? (CG ? __ldcg(A + (size_t)(k0 + c) * ld + k0 + r) : A[(size_t)(k0 + c) * ld + k0 + r]) : ((r == c) ? 1.0 : 0.0);
        w[i] = (r == c) ? 1.0 : 0.0;
    }
    if (tid == 0) F[(size_t)(k0 / NB) * nblk + k0 / NB] = 1;
#pragma unroll
    for (int i = 0; i < E; ++i) {
#pragma unroll 1
        for (int g = 0; g < G; ++g) {
            const int j = i * G + g;
            const int buf = g & 1;
            if (grp == g) col[buf][r] = t[i];                                     // T(r, j), unscaled
            if (r == j) {
#pragma unroll
                for (int i2 = 0; i2 <= i; ++i2) wrow[buf][grp + G * i2] = w[i2];  // W(j, c <= j), unscaled
            }
            __syncthreads();
            const double d = col[buf][j];
            if (tid == j) { dv[j] = d; if (!(d > 0.0) && j < nb) atomicCAS(info, 0, k0 + j + 1); }
            if (r > j) {
                const double a = col[buf][r] * fast_rcp(d);                       // L(r,j) L(c,j) = T(r,j) T(c,j) / d
#pragma unroll
                for (int i2 = 0; i2 < E; ++i2) {
                    const int c = grp + G * i2;
                    if (i2 < i) w[i2] -= a * wrow[buf][c];                        // c < j
                    else if (i2 > i) { if (c <= r) t[i2] -= a * col[buf][c]; }   // c > j
                    else { if (grp > g) { if (c <= r) t[i2] -= a * col[buf][c]; } else w[i2] -= a * wrow[buf][c]; }
                }
            }
        }
    }
    __syncthreads();
    const double rsr = rsqrt(dv[r]);
#pragma unroll
    for (int i = 0; i < E; ++i) {
        const int c = grp + G * i;
        if (r < nb && c < nb && r >= c) A[(size_t)(k0 + c) * ld + k0 + r] = t[i] * rsqrt(dv[c]);
        dinv[(size_t)c * NB + r] = (r >= c) ? w[i] * rsr : 0.0;                   // column-major 64x64: Linv(r, c)
    }
}
// ---------------------------------------------------------------------------------------------------------------------
// Blocked variant of the 64x64 diagonal-block kernel (used on the latency chain of the sparse-factor path): four 16-column
// panels.  The 16x16 diagonal block of a panel is factored by ONE warp entirely in registers (lane r holds row r, the
// pivot column travels by shuffles: no barrier inside the 16 steps, ~150 cycles per pivot instead of ~1000 with a CTA barrier
// per pivot); the rows below are solved by substitution, one thread per row; the trailing part of the tile is updated by all
// threads.  L^-1 is then assembled from the 16x16 diagonal inverses (substitution, one thread per column) and three levels of
// 16x16 block products  M_ij = -M_ii sum_k L_ik M_kj.  scratch: 2 * 64 * 65 doubles of shared memory.
__device__ long long g_potrf_prof[8];
constexpr int kPotrfLDT = 65;
constexpr size_t kPotrfScratchDoubles = 2 * NB * kPotrfLDT;
template <bool CG>
__device__ __noinline__ void potrf64_blk_dev(int n, int k0, double* A, int64_t ld, double* dinv, int* info, unsigned char* F, int nblk, double* scratch, long long* prof = nullptr,
                                             int info_base = 0, bool preloaded = false) {
    long long tp_ = clock64();
#define POTRF_T(slot) do { if (prof != nullptr && threadIdx.x == 0) { const long long now_ = clock64(); prof[slot] += now_ - tp_; tp_ = now_; } } while (0)
    constexpr int LDT = kPotrfLDT;
    double* T = scratch;                 // T[r*LDT + c]
    double* M = scratch + NB * LDT;      // M[r*LDT + c] = Linv(r, c)
    __shared__ double P[3][16][17];
    __shared__ double rsd[NB];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nb = min(NB, n - k0);
    __syncthreads();                     // the scratch region may still be read as operand tiles by the caller
    if (!preloaded) {                    // preloaded: the caller has left the tile in T and zeros in M (band_diag_update_into_scratch)
#pragma unroll
        for (int e = tid; e < NB * NB; e += 256) {
            const int c = e >> 6, r = e & 63;
            double v = (r == c) ? 1.0 : 0.0;
            if (r < nb && c < nb && r >= c) { const double* src = A + (size_t)(k0 + c) * ld + k0 + r; v = CG ? __ldcg(src) : *src; }
            T[r * LDT + c] = v;
            M[r * LDT + c] = 0.0;
        }
    }
    if (tid == 0) F[(size_t)(k0 / NB) * nblk + k0 / NB] = 1;
    __syncthreads();
    POTRF_T(0);
#pragma unroll 1                         // code size: the body is executed once per call, unrolling it only adds instruction-cache misses
    for (int p = 0; p < 4; ++p) {
        const int j0 = 16 * p;
        if (warp == 0) {
            // Division-free elimination on the chain: every stored entry carries a common scale s (1 <= s < 2^16), a step is
            //   a'(r,c) = 2^-e (d a(r,c) - a(r,j) a(c,j)),  e = exponent of the scaled pivot d,  s' = s * (d 2^-e),
            // i.e. shuffle + exponent trick + one multiply + one FMA per pivot; the sixteen 1/sqrt are taken afterwards in parallel:
            // L(r,j) = a(r,j) / sqrt(d_j s_{j-1}).
            const int r = lane & 15;     // lanes 16..31 mirror lanes 0..15
            double a[16];
#pragma unroll
            for (int c = 0; c < 16; ++c) a[c] = T[(j0 + r) * LDT + j0 + c];
            double s_prev = 1.0, my_sc = 1.0, my_sprev = 1.0;
            // the pivot column travels through a 16-double shared buffer (one store, broadcast vector loads) instead of 15 shuffles
            double* colb = &P[0][0][0];                            // P is idle during the factorisation; two buffers of 16
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                double* cb = colb + (j & 1) * 16;
                if (lane < 16) cb[r] = a[j];                       // scaled a(r, j); rows r < j hold stale upper entries nobody reads
                __syncwarp();
                const double d = cb[j];
                if (lane == 0 && !(d > 0.0) && j0 + j < nb) atomicCAS(info, 0, info_base + k0 + j0 + j + 1);
                const double p2 = __hiloint2double((2046 - ((__double2hiint(d) >> 20) & 0x7ff)) << 20, 0);   // 2^-e
                const double dm = d * p2;                          // in [1, 2)
                const double lr = a[j] * p2;
                if (lane == j) { my_sc = d * s_prev; my_sprev = s_prev; }
                s_prev *= dm;
#pragma unroll
                for (int c = j + 1; c < 16; ++c) a[c] = dm * a[c] - lr * cb[c];
            }
            const double rs = rsqrt(my_sc);
            if (lane < 16) rsd[j0 + lane] = rs * my_sprev;         // 1 / L(j,j)
#pragma unroll
            for (int c = 0; c < 16; ++c) {
                const double rc = __shfl_sync(0xffffffffu, rs, c);
                if (lane < 16) T[(j0 + r) * LDT + j0 + c] = (c <= r) ? a[c] * rc : 0.0;
            }
        }
        __syncthreads();
        POTRF_T(1);
        if (p < 3) {
            const int R = NB - (j0 + 16);          // rows below the diagonal block
            if (tid < R) {                         // X L16^T = A : x_c = (a_c - sum_{m<c} x_m L(c,m)) / L(c,c)
                const int i = j0 + 16 + tid;
                double x[16];
#pragma unroll
                for (int c = 0; c < 16; ++c) {
                    double sacc = T[i * LDT + j0 + c];
#pragma unroll
                    for (int m = 0; m < c; ++m) sacc -= x[m] * T[(j0 + c) * LDT + j0 + m];
                    x[c] = sacc * rsd[j0 + c];
                }
#pragma unroll
                for (int c = 0; c < 16; ++c) T[i * LDT + j0 + c] = x[c];
            }
            __syncthreads();
            POTRF_T(2);
            {   // trailing part of the tile: T(i, c) -= sum_m X(i, m) X(c, m), i >= c.  Thread (ti, tj) owns rows ti + 16 a, columns tj + 16 b.
                const int ti = tid >> 4, tj = tid & 15, nA = R >> 4;
                double acc9[3][3];
#pragma unroll
                for (int a2 = 0; a2 < 3; ++a2)
#pragma unroll
                    for (int b2 = 0; b2 < 3; ++b2) acc9[a2][b2] = 0.0;
                const double* xbase = T + (j0 + 16) * LDT + j0;
#pragma unroll
                for (int m0 = 0; m0 < 16; m0 += 4) {
                    double xi[3][4], xc[3][4];
#pragma unroll
                    for (int a2 = 0; a2 < 3; ++a2)
#pragma unroll
                        for (int m = 0; m < 4; ++m) {
                            xi[a2][m] = a2 < nA ? xbase[(ti + 16 * a2) * LDT + m0 + m] : 0.0;
                            xc[a2][m] = a2 < nA ? xbase[(tj + 16 * a2) * LDT + m0 + m] : 0.0;
                        }
#pragma unroll
                    for (int a2 = 0; a2 < 3; ++a2)
#pragma unroll
                        for (int b2 = 0; b2 <= a2; ++b2)
#pragma unroll
                            for (int m = 0; m < 4; ++m) acc9[a2][b2] += xi[a2][m] * xc[b2][m];
                }
#pragma unroll
                for (int a2 = 0; a2 < 3; ++a2)
#pragma unroll
                    for (int b2 = 0; b2 <= a2; ++b2) {
                        const int ii = ti + 16 * a2, cc = tj + 16 * b2;
                        if (a2 < nA && cc <= ii) T[(j0 + 16 + ii) * LDT + j0 + 16 + cc] -= acc9[a2][b2];
                    }
            }
            __syncthreads();
            POTRF_T(3);
        }
    }
    // ---- inverse: diagonal 16x16 blocks, one thread per column
    if (tid < NB) {
        const int b0 = tid & ~15, c = tid & 15;
        double x[16];
#pragma unroll
        for (int r = 0; r < 16; ++r) {
            double sacc = (r == c) ? 1.0 : 0.0;
#pragma unroll
            for (int m = 0; m < r; ++m) if (m >= c) sacc -= T[(b0 + r) * LDT + b0 + m] * x[m];
            x[r] = (r >= c) ? sacc * rsd[b0 + r] : 0.0;
        }
#pragma unroll
        for (int r = 0; r < 16; ++r) M[(b0 + r) * LDT + b0 + c] = x[r];
    }
    __syncthreads();
    POTRF_T(4);
    // ---- off-diagonal blocks, level d = i - j
    const int rr = tid >> 4, cc = tid & 15;
#pragma unroll 1
    for (int d = 1; d < 4; ++d) {
#pragma unroll 1
        for (int q = 0; q < 4 - d; ++q) {          // P_q = sum_{k=j}^{i-1} L_ik M_kj
            const int i = d + q, j = q;
            double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll 4
            for (int kk = 16 * j; kk < 16 * i; kk += 4) {
                s0 += T[(16 * i + rr) * LDT + kk] * M[kk * LDT + 16 * j + cc];
                s1 += T[(16 * i + rr) * LDT + kk + 1] * M[(kk + 1) * LDT + 16 * j + cc];
                s2 += T[(16 * i + rr) * LDT + kk + 2] * M[(kk + 2) * LDT + 16 * j + cc];
                s3 += T[(16 * i + rr) * LDT + kk + 3] * M[(kk + 3) * LDT + 16 * j + cc];
            }
            P[q][rr][cc] = (s0 + s1) + (s2 + s3);
        }
        __syncthreads();
#pragma unroll 1
        for (int q = 0; q < 4 - d; ++q) {          // M_ij = -M_ii P_q   (M_ii lower triangular)
            const int i = d + q, j = q;
            double s0 = 0.0, s1 = 0.0;      // M_ii is lower triangular: the terms m > rr are zeros of M
#pragma unroll
            for (int m = 0; m < 16; m += 2) { s0 += M[(16 * i + rr) * LDT + 16 * i + m] * P[q][m][cc]; s1 += M[(16 * i + rr) * LDT + 16 * i + m + 1] * P[q][m + 1][cc]; }
            M[(16 * i + rr) * LDT + 16 * j + cc] = -(s0 + s1);
        }
        __syncthreads();
    }
    POTRF_T(5);
#pragma unroll
    for (int e = tid; e < NB * NB; e += 256) {
        const int c = e >> 6, r = e & 63;
        if (r < nb && c < nb && r >= c) A[(size_t)(k0 + c) * ld + k0 + r] = T[r * LDT + c];
        dinv[(size_t)c * NB + r] = (r >= c) ? M[r * LDT + c] : 0.0;
    }
    __syncthreads();
    POTRF_T(6);
#undef POTRF_T
}

// The launch-per-operation path uses the blocked variant too (dynamic shared memory = its scratch); the row-operation variant
// above stays as the reference implementation behind SRK_POTRF=rowops.
__global__ void __launch_bounds__(256) k_potrf64_inv(int n, int k0, double* __restrict__ A, int64_t ld, double* __restrict__ dinv, int* __restrict__ info,
                                                     unsigned char* __restrict__ F, int nblk, int rowops) {
    extern __shared__ __align__(16) double potrf_scratch[];
    if (rowops) potrf64_inv_dev<false>(n, k0, A, ld, dinv, info, F, nblk);
    else potrf64_blk_dev<false>(n, k0, A, ld, dinv, info, F, nblk, potrf_scratch);
}

// ---------------------------------------------------------------------------------------------------------------------
// Rows below the diagonal block: X = A * L_kk^-T, i.e. X(r,c) = sum_{m<=c} A(r,m) * Linv(c,m).  128 rows per CTA.
constexpr int PS_ROWS = 128;
__global__ void __launch_bounds__(256) k_panel_solve(int n, int k0, double* __restrict__ A, int64_t ld, const double* __restrict__ dinv,
                                                     unsigned char* __restrict__ F, int nblk) {
    extern __shared__ double sm[];
    double* sLi = sm;                    // sLi[m*NB + c] = Linv(c, m): for a fixed m the 32 columns of a thread are contiguous
    double* sA = sm + NB * NB;           // sA[m*PS_ROWS + r]
    const int tid = threadIdx.x;
    const int r0 = k0 + NB + blockIdx.x * PS_ROWS;
    for (int e = tid; e < NB * NB; e += 256) { int m = e >> 6, cc = e & 63; sLi[m * NB + cc] = dinv[(size_t)m * NB + cc]; }
    // 64x64 block structure of L: F[kb][rb] != 0 iff the block (rows rb, columns kb) holds a non-zero.  An all-zero input tile
    // stays zero (X = 0 * Linv^T), so the CTA records that and leaves; the trailing update and the substitutions skip it.
    int nz_lo = 0, nz_hi = 0;
    for (int e = tid; e < NB * PS_ROWS; e += 256) {
        int m = e / PS_ROWS, r = e % PS_ROWS;
        double v = (r0 + r < n) ? A[(size_t)(k0 + m) * ld + r0 + r] : 0.0;
        sA[e] = v;
        if (v != 0.0) { if (r < NB) nz_lo = 1; else nz_hi = 1; }
    }
    nz_lo = __syncthreads_or(nz_lo);
    nz_hi = __syncthreads_or(nz_hi);
    if (tid == 0) {
        const int rb = r0 / NB;
        F[(size_t)(k0 / NB) * nblk + rb] = (unsigned char)nz_lo;
        if (rb + 1 < nblk) F[(size_t)(k0 / NB) * nblk + rb + 1] = (unsigned char)nz_hi;
    }
    if (!(nz_lo | nz_hi)) return;
    const int r = tid & (PS_ROWS - 1), cbase = (tid >> 7) * 32;
    double acc[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) acc[i] = 0.0;
    for (int m = 0; m < cbase + 32; ++m) {
        const double a = sA[m * PS_ROWS + r];
        const double* li = sLi + m * NB + cbase;
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[i] += a * li[i];   // Linv(c, m) is zero for m > c
    }
    if (r0 + r < n) {
#pragma unroll
        for (int i = 0; i < 32; ++i) A[(size_t)(k0 + cbase + i) * ld + r0 + r] = acc[i];
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// C(origin.., origin..col_end) -= P * P^T on the lower triangle, P = A[:, kcol0 .. kcol0+K).
// Persistent CTAs over a work list that every CTA derives from F (the 64x64 block structure of the panel): only pairs of
// NON-ZERO row tiles are visited, so a banded / block-sparse reduced camera system costs what its fill costs while a dense
// one runs every tile.  Two instantiations share the code: 128x128 tiles (8 warps of 32x64, best DMMA efficiency) when there
// are at least kHeavyPairs tile pairs, 64x64 tiles (4 warps of 32x32, four times the CTAs per unit of work) otherwise, so
// that a short work list is not bound by the DMMA throughput of a handful of SMs.  Both are launched; the one whose regime
// does not apply returns at once.
constexpr int KC = 16;           // K chunk per pipeline stage
constexpr int STAGES = 3;
constexpr int kHeavyPairs = 148;
constexpr int kMaxRowBlocks = 512;   // n <= 32768

__device__ __forceinline__ void dmma_m8n8k4(double& d0, double& d1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, int src_bytes) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gsrc), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

template <int TILE>
struct SyrkCfg {
    static constexpr int kWarpsR = TILE / 32;                    // warps along rows (32 rows each)
    static constexpr int kNJ = TILE == 128 ? 8 : 4;              // 8-column fragments per warp
    static constexpr int kWarpsC = TILE / (kNJ * 8);
    static constexpr int kThreads = kWarpsR * kWarpsC * 32;
    static constexpr int kSLD = TILE + 4;                        // padded [k][row] stride: conflict-free fragment loads, 16-B aligned rows
    static constexpr size_t kSmem = sizeof(double) * (2 * STAGES * KC * kSLD) + sizeof(short) * kMaxRowBlocks;
};

template <int TILE>
__global__ void __launch_bounds__(SyrkCfg<TILE>::kThreads) k_syrk_dmma(int n, double* __restrict__ A, int64_t ld, int kcol0, int K, int origin, int col_end,
                                                                      const unsigned char* __restrict__ F, int nblk) {
    using Cfg = SyrkCfg<TILE>;
    constexpr int SLD = Cfg::kSLD, NT = Cfg::kThreads, NJ = Cfg::kNJ;
    extern __shared__ double sm[];
    double* sA = sm;                              // [STAGES][KC][SLD]
    double* sB = sm + STAGES * KC * SLD;          // [STAGES][KC][SLD]
    short* nzt = (short*)(sm + 2 * STAGES * KC * SLD);
    __shared__ int s_m, s_m128;
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;

    // ---- work list: non-zero row tiles of this panel (tile index relative to origin), identical in every CTA.  All threads test
    // tiles in parallel (the F bytes of a tile are independent loads), warp ballots + a scan over the warp counts compact them.
    {
        __shared__ int s_wcnt[8];
        const int kb0 = kcol0 / NB, kb1 = (kcol0 + K) / NB;
        const int ntile = (n - origin + TILE - 1) / TILE;
        const int ntile128 = (n - origin + 127) / 128;
        int m_run = 0, m128_run = 0;
        for (int base = 0; base < max(ntile, ntile128); base += NT) {
            const int t = base + tid;
            int nz = 0, nz128 = 0;
            if (t < ntile) {
                const int rb0 = (origin + t * TILE) / NB;
                for (int rb = rb0; rb < rb0 + TILE / NB && rb < nblk; ++rb)
                    for (int kb = kb0; kb < kb1; ++kb) nz |= F[(size_t)kb * nblk + rb];
            }
            if (TILE == 128) nz128 = nz;
            else if (t < ntile128) {
                const int rb0 = (origin + t * 128) / NB;
                for (int rb = rb0; rb < rb0 + 2 && rb < nblk; ++rb)
                    for (int kb = kb0; kb < kb1; ++kb) nz128 |= F[(size_t)kb * nblk + rb];
            }
            const unsigned bal = __ballot_sync(0xffffffffu, nz != 0);
            if (lane == 0) s_wcnt[warp] = __popc(bal);
            const int c128 = __syncthreads_count(nz128 != 0);   // also publishes s_wcnt
            int off = m_run, tot = 0;
            for (int w2 = 0; w2 < NT / 32; ++w2) { if (w2 < warp) off += s_wcnt[w2]; tot += s_wcnt[w2]; }
            if (nz) nzt[off + __popc(bal & ((1u << lane) - 1))] = (short)t;
            m_run += tot; m128_run += c128;
            __syncthreads();
        }
        if (tid == 0) { s_m = m_run; s_m128 = m128_run; }
    }
    __syncthreads();
    const int m = s_m;
    const bool heavy = s_m128 * (s_m128 + 1) / 2 >= kHeavyPairs;
    if (heavy != (TILE == 128)) return;
    const int npairs = m * (m + 1) / 2;
    const int nk = K / KC;
    const int wr = (warp % Cfg::kWarpsR) * 32;
    const int wc = (warp / Cfg::kWarpsR) * (NJ * 8);
    const int g = lane >> 2, tg = lane & 3;

    for (int p = blockIdx.x; p < npairs; p += gridDim.x) {
        int a = (int)((sqrtf(8.0f * (float)p + 1.0f) - 1.0f) * 0.5f);
        while (a * (a + 1) / 2 > p) --a;
        while ((a + 1) * (a + 2) / 2 <= p) ++a;
        const int bb = p - a * (a + 1) / 2;
        const int i0 = origin + (int)nzt[a] * TILE, j0 = origin + (int)nzt[bb] * TILE;
        if (j0 >= col_end) continue;

        auto load_chunk = [&](int stage, int kc) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int v = tid + NT * q;            // KC * TILE / 2 16-byte vectors per operand tile
                const int k = v / (TILE / 2), rp = (v % (TILE / 2)) * 2;
                const size_t colo = (size_t)(kcol0 + kc * KC + k) * ld;
                {
                    const int row = i0 + rp;
                    int bytes = (n - row) * 8; bytes = bytes > 16 ? 16 : (bytes < 0 ? 0 : bytes);
                    cp_async16(sA + (stage * KC + k) * SLD + rp, A + colo + (bytes > 0 ? row : 0), bytes);
                }
                {
                    const int row = j0 + rp;
                    int bytes = (n - row) * 8; bytes = bytes > 16 ? 16 : (bytes < 0 ? 0 : bytes);
                    cp_async16(sB + (stage * KC + k) * SLD + rp, A + colo + (bytes > 0 ? row : 0), bytes);
                }
            }
        };
        load_chunk(0, 0); cp_async_commit();
        if (nk > 1) load_chunk(1, 1);
        cp_async_commit();

        double acc[4][NJ][2];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < NJ; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }

        for (int kc = 0; kc < nk; ++kc) {
            cp_async_wait<1>();
            __syncthreads();
            if (kc + 2 < nk) load_chunk((kc + 2) % STAGES, kc + 2);
            cp_async_commit();
            const double* cA = sA + (kc % STAGES) * KC * SLD;
            const double* cB = sB + (kc % STAGES) * KC * SLD;
#pragma unroll
            for (int ks = 0; ks < KC; ks += 4) {
                double af[4], bf[NJ];
                const double* pa = cA + (ks + tg) * SLD + wr + g;
                const double* pb = cB + (ks + tg) * SLD + wc + g;
#pragma unroll
                for (int i = 0; i < 4; ++i) af[i] = pa[i * 8];
#pragma unroll
                for (int j = 0; j < NJ; ++j) bf[j] = pb[j * 8];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < NJ; ++j) dmma_m8n8k4(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
            }
        }
        cp_async_wait<0>();
        __syncthreads();   // the stage buffers are free for the next tile
        // epilogue: C(row, col) -= acc on the lower triangle inside [origin, col_end).  All loads of the read-modify-write are
        // issued before the first store: behind a store to A the compiler must assume aliasing and would serialise them.
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                const int row = i0 + wr + i * 8 + g;
                const int col = j0 + wc + j * 8 + tg * 2;
                const bool ok0 = row < n && col < col_end && row >= col, ok1 = row < n && col + 1 < col_end && row >= col + 1;
                acc[i][j][0] = (ok0 ? A[(size_t)col * ld + row] : 0.0) - acc[i][j][0];
                acc[i][j][1] = (ok1 ? A[(size_t)(col + 1) * ld + row] : 0.0) - acc[i][j][1];
            }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                const int row = i0 + wr + i * 8 + g;
                const int col = j0 + wc + j * 8 + tg * 2;
                if (row < n) {
                    if (col < col_end && row >= col) A[(size_t)col * ld + row] = acc[i][j][0];
                    if (col + 1 < col_end && row >= col + 1) A[(size_t)(col + 1) * ld + row] = acc[i][j][1];
                }
            }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Right-side solve of one 64-column block against the inverted diagonal block, on the tensor pipe:
//   X(rows, 0..63) <- X * Linv^T,  X(r, c) = sum_q X(r, q) Linv(c, q)      (in place; a CTA owns 128 full rows, so it reads its
// operands completely before it writes).  Used by the Cholesky panel (rows below the diagonal block; F receives the 64x64
// block structure, all-zero tiles are recorded and left alone) and by the EKF TRSM (F == nullptr).
constexpr int RS_ROWS = 128;
constexpr int RS_SLDA = RS_ROWS + 4;
constexpr int RS_SLDB = NB + 4;
constexpr size_t kRightSolveSmem = sizeof(double) * (NB * RS_SLDA + NB * RS_SLDB);
// ncols (<= 64): columns of the block that exist (a ragged last block of a caller whose matrix ends there: neither read nor written beyond)
__global__ void __launch_bounds__(256) k_right_solve_dmma(int rows, int row0, double* __restrict__ A, int64_t lda, const double* __restrict__ dinv,
                                                          unsigned char* __restrict__ Frow, int nblk, int ncols) {
    extern __shared__ double sm[];
    double* sA = sm;                       // [q][row]
    double* sB = sm + NB * RS_SLDA;        // [q][c] = Linv(c, q)
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int r0 = row0 + blockIdx.x * RS_ROWS;
#pragma unroll
    for (int it = 0; it < 16; ++it) {      // 64 x 128 doubles = 4096 16-byte vectors
        const int v = tid + 256 * it;
        const int q = v >> 6, rp = (v & 63) * 2;
        const int row = r0 + rp;
        int bytes = (rows - row) * 8; bytes = bytes > 16 ? 16 : (bytes < 0 ? 0 : bytes);
        if (q >= ncols) bytes = 0;
        cp_async16(sA + q * RS_SLDA + rp, A + (bytes > 0 ? (size_t)q * lda + row : 0), bytes);
    }
#pragma unroll
    for (int it = 0; it < 8; ++it) {       // 64 x 64 doubles = 2048 vectors; dinv[q*64 + c] = Linv(c, q)
        const int v = tid + 256 * it;
        const int q = v >> 5, cp = (v & 31) * 2;
        cp_async16(sB + q * RS_SLDB + cp, dinv + (size_t)q * NB + cp, 16);
    }
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    if (Frow != nullptr) {                 // block structure of L: which 64-row halves of this tile hold a non-zero
        int nz_lo = 0, nz_hi = 0;
#pragma unroll
        for (int it = 0; it < 16; ++it) {
            const int v = tid + 256 * it;
            const int q = v >> 6, rp = (v & 63) * 2;
            const double a0 = sA[q * RS_SLDA + rp], a1 = sA[q * RS_SLDA + rp + 1];
            if (a0 != 0.0 || a1 != 0.0) { if (rp < NB) nz_lo = 1; else nz_hi = 1; }
        }
        nz_lo = __syncthreads_or(nz_lo);
        nz_hi = __syncthreads_or(nz_hi);
        if (tid == 0) {
            const int rb = r0 / NB;
            Frow[rb] = (unsigned char)nz_lo;
            if (rb + 1 < nblk) Frow[rb + 1] = (unsigned char)nz_hi;
        }
        if (!(nz_lo | nz_hi)) return;      // X = 0 * Linv^T stays zero
    }
    const int wr = (warp & 3) * 32, wc = (warp >> 2) * 32;
    const int g = lane >> 2, tg = lane & 3;
    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }
#pragma unroll 4
    for (int ks = 0; ks < NB; ks += 4) {
        double af[4], bf[4];
        const double* pa = sA + (ks + tg) * RS_SLDA + wr + g;
        const double* pb = sB + (ks + tg) * RS_SLDB + wc + g;
#pragma unroll
        for (int i = 0; i < 4; ++i) af[i] = pa[i * 8];
#pragma unroll
        for (int j = 0; j < 4; ++j) bf[j] = pb[j * 8];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) dmma_m8n8k4(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int row = r0 + wr + i * 8 + g;
            const int c = wc + j * 8 + tg * 2;
            if (row < rows) { if (c < ncols) A[(size_t)c * lda + row] = acc[i][j][0]; if (c + 1 < ncols) A[(size_t)(c + 1) * lda + row] = acc[i][j][1]; }
        }
}

// ---------------------------------------------------------------------------------------------------------------------
// Sparse-factor path of the factorisation.  The reduced camera system of a scene with localized visibility is block-banded
// (plus the wrap-around rows of a closed camera ring): per 64-column step there are a handful of non-zero tiles, the ~1.4 GFLOP of
// the whole factorisation are nothing, and the run time of the launch-per-operation path above is the latency of ~630 dependent
// kernel launches.  Here ONE thread-block cluster of 8 CTAs walks the block columns inside a single kernel:
//     sync A | every CTA lists the non-zero row tiles of column k (F) | right solves X = A(r,k) Linv_k^T, one tile per CTA |
//     sync B | tile-pair updates C(ra,rb) -= X_a X_b^T spread over CTAs 1..7, fill marked in F |
//            | CTA 0: update of the next diagonal tile, then potrf of column k+1 (look-ahead: overlaps the other updates)
// barrier.cluster (release/acquire, ~0.2 us) orders the global-memory tile traffic between the CTAs; tile operands are read with
// cp.async.cg / ld.global.cg (L2, never a stale L1 line).  F starts as the non-zero tile pattern of the input (k_tile_pattern),
// fill is tracked symbolically as updates are applied, so the structure handed to the substitutions is exact.
constexpr int kBandCluster = 8;
constexpr int kCholMaxParts = CholPartition::kMaxParts;
constexpr int kBandTS = NB + 4;        // [k][row] tile stride in shared memory: 68 = 4 (mod 16), conflict-free fragment loads
constexpr int kBandMaxFill = 4;        // sparse path when the input has <= kBandMaxFill * nblk non-zero off-diagonal tiles
struct BandSmem {
    double A[NB * kBandTS];
    double B[NB * kBandTS];
    int rows[kMaxRowBlocks];
    int m;
};
constexpr size_t kBandSmemRequest = sizeof(BandSmem) > 120 * 1024 ? sizeof(BandSmem) : 120 * 1024;
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
__device__ __forceinline__ unsigned cluster_cta_rank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r)); return r; }

// non-zero pattern of the lower 64x64 tiles of A (column-major): F0[c*nblk + r] = 1 iff tile (r, c), r > c, holds a non-zero; count += 1
__global__ void __launch_bounds__(256) k_tile_pattern(int n, const double* __restrict__ A, int64_t ld, int nblk, unsigned char* __restrict__ F, int* __restrict__ count) {
    const int c = blockIdx.x, r = blockIdx.y;
    if (r <= c) return;
    const int tid = threadIdx.x;
    int nz = 0;
#pragma unroll 4
    for (int it = 0; it < 8; ++it) {        // 64 columns x 32 pairs of rows
        const int v = tid + 256 * it;
        const int q = v >> 5, rp = (v & 31) * 2;
        const int row = r * NB + rp, col = c * NB + q;
        if (col < n && row + 1 < n) { const double2 x = *reinterpret_cast<const double2*>(A + (size_t)col * ld + row); nz |= (x.x != 0.0) | (x.y != 0.0); }
        else if (col < n && row < n) nz |= A[(size_t)col * ld + row] != 0.0;
    }
    nz = __syncthreads_or(nz);
    if (tid == 0 && nz) { F[(size_t)c * nblk + r] = 1; atomicAdd(count, 1); }
}

// s[q*TS + rr] = g(row0 + rr, col0 + q), zero outside the matrix
__device__ __forceinline__ void band_load_tile(double* s, const double* g, int64_t ld, int row0, int col0, int n) {
    const int tid = threadIdx.x;
#pragma unroll
    for (int it = 0; it < 8; ++it) {
        const int v = tid + 256 * it;
        const int q = v >> 5, rp = (v & 31) * 2;
        const int row = row0 + rp;
        int bytes = (n - row) * 8; bytes = bytes > 16 ? 16 : (bytes < 0 ? 0 : bytes);
        if (col0 + q >= n) bytes = 0;
        cp_async16(s + q * kBandTS + rp, g + (bytes > 0 ? (size_t)(col0 + q) * ld + row : 0), bytes);
    }
}
// acc = A_tile(64 x 64) * B_tile(64 x 64)^T from the [k][row] shared-memory tiles; warp (wr, wc) owns 32 x 16
__device__ __forceinline__ void band_tile_nt(const double* sA, const double* sB, double (&acc)[4][2][2]) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wr = (warp & 1) * 32, wc = (warp >> 1) * 16;
    const int g = lane >> 2, tg = lane & 3;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }
#pragma unroll 4
    for (int ks = 0; ks < NB; ks += 4) {
        double af[4], bf[2];
        const double* pa = sA + (ks + tg) * kBandTS + wr + g;
        const double* pb = sB + (ks + tg) * kBandTS + wc + g;
#pragma unroll
        for (int i = 0; i < 4; ++i) af[i] = pa[i * 8];
#pragma unroll
        for (int j = 0; j < 2; ++j) bf[j] = pb[j * 8];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 2; ++j) dmma_m8n8k4(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
    }
}

// One 64x64x64 tile operation of the cluster kernel (one copy of the code: it runs a few times per step, every copy would be
// another set of instruction-cache misses).  mode 0: A(ra,k) <- A(ra,k) Linv_k^T in place;  mode 1: A(ra,rb) -= X(ra,k) X(rb,k)^T.
// atomic (mode 1 only): the destination tile lies in the separator block shared by several clusters (nested-dissection order,
// see dense_cholesky_set_partition): the update goes out as red.global.add.f64 instead of a read-modify-write.
__device__ __noinline__ void band_tile_op(BandSmem& sm, int n, double* A, int64_t ld, const double* dinv, unsigned char* F, int nblk, int mode, int ra, int rb, int k,
                                          bool atomic) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int wr = (warp & 1) * 32, wc = (warp >> 1) * 16, g = lane >> 2, tg = lane & 3;
    __syncthreads();
    band_load_tile(sm.A, A, ld, ra * NB, k * NB, n);
    if (mode == 0) {   // sB[q*TS + c] = Linv(c, q) = dinv_k[q*64 + c]
        const double* dk = dinv + (size_t)k * NB * NB;
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            const int v = tid + 256 * it;
            const int q = v >> 5, cp = (v & 31) * 2;
            cp_async16(sm.B + q * kBandTS + cp, dk + (size_t)q * NB + cp, 16);
        }
    } else {
        band_load_tile(sm.B, A, ld, rb * NB, k * NB, n);
    }
    cp_async_commit(); cp_async_wait<0>();
    __syncthreads();
    double acc[4][2][2];
    band_tile_nt(sm.A, sm.B, acc);
    const int cbase = (mode == 0 ? k : rb) * NB;
    double cur[4][2][2];
    if (mode == 1 && !atomic) {   // all 16 loads of the read-modify-write first: behind a store the compiler has to assume aliasing and serialises them
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 2; ++j)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int row = ra * NB + wr + 8 * i + g, col = cbase + wc + 8 * j + 2 * tg + e;
                    cur[i][j][e] = (row < n && col < n && row >= col) ? __ldcg(A + (size_t)col * ld + row) : 0.0;
                }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int row = ra * NB + wr + 8 * i + g, col = cbase + wc + 8 * j + 2 * tg + e;
                if (row >= n || col >= n) continue;
                double* p = A + (size_t)col * ld + row;
                if (mode == 0) *p = acc[i][j][e];
                else if (row >= col) { if (atomic) atomicAdd(p, -acc[i][j][e]); else *p = cur[i][j][e] - acc[i][j][e]; }
            }
    if (mode == 1 && tid == 0 && ra != rb) F[(size_t)rb * nblk + ra] = 1;   // fill (or already non-zero)
}

// The update of the NEXT diagonal tile, A(k1,k1) - X(k1,k) X(k1,k)^T, left directly in the scratch layout of potrf64_blk_dev (T = the tile, lower
// triangle, identity outside the matrix; M = 0) instead of going back to global memory and being read again by the factorisation that follows at
// once on the same CTA: one operand load instead of two, no store + reload round trip through L2 on the chain.
__device__ __noinline__ void band_diag_update_into_scratch(BandSmem& sm, int n, const double* A, int64_t ld, int k1, int k, double* scratch) {
    constexpr int LDT = kPotrfLDT;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int wr = (warp & 1) * 32, wc = (warp >> 1) * 16, g = lane >> 2, tg = lane & 3;
    __syncthreads();
    band_load_tile(sm.A, A, ld, k1 * NB, k * NB, n);
    cp_async_commit();
    const int nb = min(NB, n - k1 * NB);
    double cur[4][2][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int r = wr + 8 * i + g, c = wc + 8 * j + 2 * tg + e;
                cur[i][j][e] = (r < nb && c < nb && r >= c) ? __ldcg(A + (size_t)(k1 * NB + c) * ld + k1 * NB + r) : 0.0;
            }
    cp_async_wait<0>();
    __syncthreads();
    double acc[4][2][2];
    band_tile_nt(sm.A, sm.A, acc);
    __syncthreads();                     // every warp is done with the operand tile: the scratch layout overwrites it
    double* T = scratch;
    double* M = scratch + NB * LDT;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int r = wr + 8 * i + g, c = wc + 8 * j + 2 * tg + e;
                double v = (r == c) ? 1.0 : 0.0;
                if (r < nb && c < nb && r >= c) v = cur[i][j][e] - acc[i][j][e];
                T[r * LDT + c] = v;
                M[r * LDT + c] = 0.0;
            }
}

// development aid: clock64() per phase, CTA 0 (slots 0..5: wait A, list, solve, wait B, diag update, potrf) and CTA 1 (8..11: wait A, solve, wait B, updates)
__device__ long long g_band_prof[16];
#define BAND_T(slot) do { if (tid == 0 && rank <= 1 && blockIdx.x < kBandCluster) { const long long now_ = clock64(); g_band_prof[rank * 8 + (slot)] += now_ - t_prev; t_prev = now_; } } while (0)
// Column ranges of the clusters (nested-dissection order of the reduced camera system, solve_order.h): in phase 0 cluster c
// factors the block columns [k0[c], k1[c]) of its own part -- parts do not touch each other's tiles; what they contribute to the
// separator block (rows and columns >= ksep) is added atomically -- and in phase 1 ONE cluster factors [ksep, nblk).
// An unpartitioned factorisation is phase 1 with ksep = 0.
using BandParts = CholPartition;
__global__ void __cluster_dims__(kBandCluster, 1, 1) __launch_bounds__(256, 1)
k_band_chol(int n, double* A, int64_t ld, double* dinv, int* info, unsigned char* F, int nblk, const BandParts bp, int phase, int info_base) {
    extern __shared__ __align__(16) unsigned char band_raw[];
    BandSmem& sm = *reinterpret_cast<BandSmem*>(band_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int rank = (int)cluster_cta_rank();
    const int cl = blockIdx.x / kBandCluster;
    const int kbeg = phase == 0 ? bp.k0[cl] : bp.ksep;
    const int kend = phase == 0 ? bp.k1[cl] : nblk;
    const int shared_from = phase == 0 ? bp.ksep : 0x7fffffff;      // tiles (ra, rb) with rb >= shared_from are shared between clusters
    if (kbeg >= kend) return;                                        // uniform over the cluster

    auto solve_tile = [&](int r, int k) { band_tile_op(sm, n, A, ld, dinv, F, nblk, 0, r, r, k, false); };
    auto update_pair = [&](int ra, int rb, int k) { band_tile_op(sm, n, A, ld, dinv, F, nblk, 1, ra, rb, k, rb >= shared_from); };

    static_assert(2 * NB * kBandTS >= (int)kPotrfScratchDoubles, "potrf scratch must fit the two operand tiles");
    if (rank == 0) potrf64_blk_dev<true>(n, kbeg * NB, A, ld, dinv + (size_t)kbeg * NB * NB, info, F, nblk, sm.A, nullptr, info_base);
    long long t_prev = clock64();
    for (int k = kbeg; k < kend; ++k) {
        cluster_sync_all();                                   // A: potrf(k) and the updates of step k-1 are visible
        BAND_T(0);
        // non-zero row tiles of column k, ascending
        if (warp == 0) {
            int m = 0;
            for (int base = k + 1; base < nblk; base += 32) {
                const int r = base + lane;
                const bool nz = r < nblk && __ldcg(F + (size_t)k * nblk + r) != 0;
                const unsigned bal = __ballot_sync(0xffffffffu, nz);
                if (nz) sm.rows[m + __popc(bal & ((1u << lane) - 1))] = r;
                m += __popc(bal);
            }
            if (lane == 0) sm.m = m;
        }
        __syncthreads();
        BAND_T(1);
        const int m = sm.m;
        for (int i = rank; i < m; i += kBandCluster) solve_tile(sm.rows[i], k);
        BAND_T(2);
        cluster_sync_all();                                   // B: every X(r,k) is visible
        BAND_T(3);
        const bool next_in_list = m > 0 && sm.rows[0] == k + 1;
        const int npairs = m * (m + 1) / 2;
        if (rank == 0) {
            // the next diagonal tile belongs to this cluster alone when it lies inside its own range: its update can stay on the CTA
            const bool fuse = next_in_list && k + 1 < kend && k + 1 < shared_from;
            if (fuse) band_diag_update_into_scratch(sm, n, A, ld, k + 1, k, sm.A);
            else if (next_in_list) update_pair(k + 1, k + 1, k);
            __syncthreads();
            BAND_T(4);
            if (k + 1 < kend) potrf64_blk_dev<true>(n, (k + 1) * NB, A, ld, dinv + (size_t)(k + 1) * NB * NB, info, F, nblk, sm.A, g_potrf_prof, info_base, fuse);
            BAND_T(5);
        } else {
            for (int p = (next_in_list ? 1 : 0) + (rank - 1); p < npairs; p += kBandCluster - 1) {
                int a = (int)((sqrtf(8.0f * (float)p + 1.0f) - 1.0f) * 0.5f);
                while (a * (a + 1) / 2 > p) --a;
                while ((a + 1) * (a + 2) / 2 <= p) ++a;
                const int b = p - a * (a + 1) / 2;
                update_pair(sm.rows[a], sm.rows[b], k);
            }
            BAND_T(4);
        }
    }
}
void dense_cholesky_band_profile_report() {
    long long h[16];
    if (cudaMemcpyFromSymbol(h, g_band_prof, sizeof(h)) != cudaSuccess) return;
    const char* n0[6] = {"wait A", "list", "solve", "wait B", "diag update", "potrf"};
    printf("  band kernel, CTA 0 (cycles):");
    for (int i = 0; i < 6; ++i) printf(" %s=%lld", n0[i], h[i]);
    printf("\n  band kernel, CTA 1 (cycles): wait A=%lld list=%lld solve=%lld wait B=%lld updates=%lld\n", h[8], h[9], h[10], h[11], h[12]);
    printf("  trsv cluster, CTA 0 (cycles): diag=%lld cluster wait=%lld tile ops=%lld\n", h[13], h[14], h[15]);
    long long z[16] = {0};
    cudaMemcpyToSymbol(g_band_prof, z, sizeof(z));
    long long hp[8];
    if (cudaMemcpyFromSymbol(hp, g_potrf_prof, sizeof(hp)) == cudaSuccess) {
        printf("  potrf64_blk (cycles): load=%lld potrf16=%lld trsm=%lld syrk=%lld inv_diag=%lld inv_offdiag=%lld store=%lld\n", hp[0], hp[1], hp[2], hp[3], hp[4], hp[5], hp[6]);
        cudaMemcpyToSymbol(g_potrf_prof, z, sizeof(hp));
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Triangular solves with the stored inverses of the diagonal blocks, ONE cooperative launch and no grid-wide barrier.
//   forward  (L y = b):   blocks in ascending order:  y_k = Linv_k b_k ;   b_j -= L(j,k) y_k   for j > k
//   backward (L^T x = y): blocks in descending order: x_k = Linv_k^T b_k ; b_j -= L(k,j)^T x_k for j < k
// Row block j of the right-hand side is owned by CTA (order(j) mod G): every update of b_j happens inside one CTA in
// program order.  The owner of block k turns b_k into the solved y_k, publishes it (ybuf + release flag = epoch) and the
// other CTAs pick it up with an acquire spin only when they own a non-zero tile in that block column (F, the 64x64 block
// structure of L).  All CTAs are co-resident (cooperative launch), so the spins cannot deadlock.
__device__ __forceinline__ int ld_acquire(const int* p) { int v; asm volatile("ld.acquire.gpu.global.s32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void st_release(int* p, int v) { asm volatile("st.release.gpu.global.s32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory"); }

// Sparse-factor path of the substitutions: when L has few non-zero 64x64 tiles (block-banded reduced camera systems) the whole solve
// is a latency chain, and handing solved blocks from CTA to CTA through global memory (k_trsv_flags) costs ~6 us per block.
// Here ONE CTA walks the chain with the right-hand side resident in shared memory and no inter-CTA traffic at all.
//   * k_build_trsv_lists (once per factorisation) turns F into two flat op lists, forward and backward: per block the diagonal
//     op (k,k) followed by one op per non-zero tile (k,j); op = (k << 16) | j.
//   * k_trsv_single walks a list.  Every op is a 64x64 GEMV slice: lane = (row & 7, q) owns the partial sum over columns
//     c == q (mod 4) of one row, 16 operands per thread, reduced over q with two shuffles (no shared-memory reduction).  The
//     operands of op t+2 are loaded into registers before op t is computed, so the L2/HBM latency of the tiles is off the chain;
//     tile ops of one block column write different row blocks and need no barrier between them: 2 barriers per BLOCK.
// nz_tiles (device) decides which of the two kernels does the work; both are launched, the other returns at once.
constexpr int kTrsvSparseMaxN = 24576;      // b in shared memory: 192 KB
constexpr int kTrsvSparseMaxBlk = kTrsvSparseMaxN / NB;
constexpr int kTrsvSparseFill = 6;          // sparse path while nz_tiles <= kTrsvSparseFill * nblk
__host__ __device__ __forceinline__ bool trsv_use_sparse(int n, int nblk, int nz_tiles) { return n <= kTrsvSparseMaxN && nz_tiles <= kTrsvSparseFill * nblk; }

// out[0] = number of non-zero 64x64 tiles of L (diagonal included), out[1] = ops per list; listF / listB: (kTrsvSparseFill+1)*nblk ints each.
__global__ void __launch_bounds__(1024) k_build_trsv_lists(int n, int nblk, const unsigned char* __restrict__ F, int* __restrict__ out, int* __restrict__ listF,
                                                           int* __restrict__ listB) {
    __shared__ int cntF[kTrsvSparseMaxBlk], cntB[kTrsvSparseMaxBlk], offF[kTrsvSparseMaxBlk], offB[kTrsvSparseMaxBlk];
    __shared__ int s_total;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;
    if (n > kTrsvSparseMaxN) {   // dense path only: just the tile count
        int c = 0;
        for (int e = tid; e < nblk * nblk; e += blockDim.x) c += F[e] != 0;
        for (int s2 = 16; s2 > 0; s2 >>= 1) c += __shfl_xor_sync(0xffffffffu, c, s2);
        if (tid == 0) s_total = 0;
        __syncthreads();
        if (lane == 0) atomicAdd(&s_total, c);
        __syncthreads();
        if (tid == 0) { out[0] = s_total; out[1] = 0; }
        return;
    }
    for (int k = tid; k < nblk; k += blockDim.x) { cntF[k] = 0; cntB[k] = 0; }
    __syncthreads();
    // pass 1: off-diagonal non-zero tiles per block column (forward) / per block row (backward)
    for (int k = warp; k < nblk; k += nwarp) {
        int c = 0;
        for (int j = k + 1 + lane; j < nblk; j += 32)
            if (F[(size_t)k * nblk + j] != 0) { ++c; atomicAdd(&cntB[j], 1); }
        for (int s2 = 16; s2 > 0; s2 >>= 1) c += __shfl_xor_sync(0xffffffffu, c, s2);
        if (lane == 0) cntF[k] = c;
    }
    __syncthreads();
    if (tid == 0) {
        int run = 0;
        for (int k = 0; k < nblk; ++k) { offF[k] = run; run += 1 + cntF[k]; }
        const int total = run;
        run = 0;
        for (int k = nblk - 1; k >= 0; --k) { offB[k] = run; run += 1 + cntB[k]; }
        s_total = total;
        out[0] = total; out[1] = total;
    }
    __syncthreads();
    if (!trsv_use_sparse(n, nblk, s_total)) return;
    for (int k = tid; k < nblk; k += blockDim.x) {
        listF[offF[k]] = (k << 16) | k;
        listB[offB[k]] = (k << 16) | k;
        cntB[k] = 0;   // reused as the fill cursor of the backward list
    }
    __syncthreads();
    for (int k = warp; k < nblk; k += nwarp) {
        int run = 0;
        for (int j0 = k + 1; j0 < nblk; j0 += 32) {
            const int j = j0 + lane;
            const bool nz = j < nblk && F[(size_t)k * nblk + j] != 0;
            const unsigned bal = __ballot_sync(0xffffffffu, nz);
            if (nz) {
                listF[offF[k] + 1 + run + __popc(bal & ((1u << lane) - 1))] = (k << 16) | j;      // solved block k updates row block j > k
                listB[offB[j] + 1 + atomicAdd(&cntB[j], 1)] = (j << 16) | k;                       // solved block j updates row block k < j
            }
            run += __popc(bal);
        }
    }
}

__device__ __forceinline__ void trsv_load_op(int op, int n, const double* __restrict__ L, int64_t ld, const double* __restrict__ dinv, int backward, int r, int q,
                                             double (&v)[16]) {
    const int k = op >> 16, j = op & 0xffff;
    if (k == j) {
        const double* Di = dinv + (size_t)k * NB * NB;      // Di[c*NB + r] = Linv(r, c), zeros above the diagonal
        if (!backward) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = Di[(size_t)(q + 4 * i) * NB + r];      // y_r = sum_c Linv(r,c) b_c
        } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = Di[(size_t)r * NB + q + 4 * i];        // x_r = sum_c Linv(c,r) b_c
        }
    } else {
        const int k0 = k * NB, j0 = j * NB;
        if (!backward) {                                    // b_j[r] -= sum_c L(j0+r, k0+c) y_c   (block k < j is full)
            const bool ok = j0 + r < n;
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = ok ? L[(size_t)(k0 + q + 4 * i) * ld + j0 + r] : 0.0;
        } else {                                            // b_j[r] -= sum_c L(k0+c, j0+r) x_c   (block k > j may be the ragged last one)
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = (k0 + q + 4 * i < n) ? L[(size_t)(j0 + r) * ld + k0 + q + 4 * i] : 0.0;
        }
    }
}

__global__ void __launch_bounds__(256, 1) k_trsv_single(int n, const double* __restrict__ L, int64_t ld, const double* __restrict__ dinv, double* __restrict__ b,
                                                        const int* __restrict__ stats, const int* __restrict__ list, int backward) {
    extern __shared__ double sb[];          // [nblk*NB] the whole right-hand side, padded to full blocks; then the op list (ints)
    __shared__ double yk[NB];
    const int nblk = (n + NB - 1) / NB;
    if (!trsv_use_sparse(n, nblk, stats[0])) return;
    const int nops = stats[1];
    int* sop = (int*)(sb + (size_t)nblk * NB);
    const int tid = threadIdx.x;
    const int q = tid & 3, r = tid >> 2;    // 4 consecutive lanes share a row
    for (int i = tid; i < nblk * NB; i += 256) sb[i] = i < n ? b[i] : 0.0;
    for (int i = tid; i < nops; i += 256) sop[i] = list[i];
    __syncthreads();
    double vA[16], vB[16], vC[16];
    if (nops > 0) trsv_load_op(sop[0], n, L, ld, dinv, backward, r, q, vA);
    if (nops > 1) trsv_load_op(sop[1], n, L, ld, dinv, backward, r, q, vB);
    // one op: the operands of op t are in `cur` (loaded two ops ago); the operands of op t+2 go to `pre`.  The three buffers
    // rotate by name (the loop is unrolled by 3), never by copying: a register copy would wait for the load it copies.
    auto step = [&](int t, double (&cur)[16], double (&pre)[16]) {
        if (t >= nops) return;
        if (t + 2 < nops) trsv_load_op(sop[t + 2], n, L, ld, dinv, backward, r, q, pre);
        const int op = sop[t];
        const int k = op >> 16, j = op & 0xffff;
        const bool diag = k == j;
        if (diag) __syncthreads();          // the pending updates of row block k are complete
        const double* x = diag ? sb + k * NB : yk;
        double s = 0.0;
#pragma unroll
        for (int i = 0; i < 16; ++i) s += cur[i] * x[q + 4 * i];
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        if (diag) {
            if (q == 0) { yk[r] = s; if (k * NB + r < n) b[k * NB + r] = s; }
            __syncthreads();
        } else if (q == 0) {
            sb[j * NB + r] -= s;
        }
    };
    for (int t = 0; t < nops; t += 3) {
        step(t, vA, vC);
        step(t + 1, vB, vA);
        step(t + 2, vC, vB);
    }
}

// Cluster version of the sparse-factor substitutions: the op list is split over the 8 CTAs of one thread-block cluster by the
// owner of the row block an op writes (block j belongs to CTA j mod 8, which keeps b_j in shared memory).  The owner of block k
// turns b_k into the solved y_k (GEMV with the stored inverse) and publishes it in global memory; ONE barrier.cluster per block
// column makes it visible; every CTA then applies its own tiles of that column.  A single CTA can keep ~64 KB of tile loads in
// flight (~40 GB/s: the 30 MB of tiles took 0.8 ms); eight CTAs with a 4-deep cp.async ring each keep ~0.8 MB in flight.
constexpr int kTrsvClRing = 4;
constexpr int kTrsvClLDT = NB + 2;                      // 66: 16-byte aligned rows; forward ops read [c][r] (conflict-free), backward ops read
                                                        // [r][c] with 16-byte loads (row stride 33 x 16 B: conflict-free per quarter warp)
constexpr int kTrsvClMaxOwn = (kTrsvSparseMaxBlk + kBandCluster - 1) / kBandCluster;
struct TrsvClSmem {
    double ring[kTrsvClRing][NB * kTrsvClLDT];
    double bown[kTrsvClMaxOwn][NB];
    double yk[2][NB];                                   // y_k of the current / next block column, pushed by its owner (DSMEM)
    int ops[(kTrsvSparseFill + 1) * kTrsvSparseMaxBlk];
    int nmy;
};
// ring slot <- the 64x64 operand of `op` as it lies in memory: element (major m, minor i) lands at slot[m*66 + i].
// forward: diag Linv(i, m), tile L(j0+i, k0+m);  backward: diag Linv(i, m) read transposed, tile L(k0+i, j0+m).
__device__ __forceinline__ void trsv_cl_issue(double* slot, int op, int n, const double* __restrict__ L, int64_t ld, const double* __restrict__ dinv, int backward) {
    const int k = op >> 16, j = op & 0xffff;
    const int tid = threadIdx.x;
    const double* base; int64_t smaj; int row0;
    if (k == j) { base = dinv + (size_t)k * NB * NB; smaj = NB; row0 = 0; }
    else if (!backward) { base = L + (size_t)(k * NB) * ld + j * NB; smaj = ld; row0 = j * NB; }
    else { base = L + (size_t)(j * NB) * ld + k * NB; smaj = ld; row0 = k * NB; }
    const int lim = (k == j) ? NB : n - row0;          // valid minor indices: [0, lim)
    const int mi = (tid & 31) * 2;
    int bytes = (lim - mi) * 8; bytes = bytes > 16 ? 16 : (bytes < 0 ? 0 : bytes);
    const double* src = base + (bytes > 0 ? mi : 0);
#pragma unroll
    for (int it = 0; it < 8; ++it) {
        const int mj = (tid >> 5) + 8 * it;
        cp_async16(slot + mj * kTrsvClLDT + mi, bytes > 0 ? src + (size_t)mj * smaj : base, bytes);
    }
}
// Partitioned factor (BandParts): the substitutions run as four launches -- forward over the parts (one cluster each; what a part
// subtracts from the separator rows is collected from zero and added atomically at the end), forward + backward over the separator
// (one cluster), backward over the parts (the separator's solved blocks are read as given multipliers first).
//   phase 0: blocks [k0[c], k1[c]) of part c;   phase 1: blocks [ksep, nblk) (ksep = 0: the whole unpartitioned system).
__global__ void __cluster_dims__(kBandCluster, 1, 1) __launch_bounds__(256, 1)
k_trsv_cluster(int n, const double* __restrict__ L, int64_t ld, const double* __restrict__ dinv, double* b, const int* __restrict__ stats,
               const int* __restrict__ list, int backward, const BandParts bp, int phase) {
    extern __shared__ __align__(16) unsigned char trsv_raw[];
    TrsvClSmem& sm = *reinterpret_cast<TrsvClSmem*>(trsv_raw);
    const int nblk = (n + NB - 1) / NB;
    if (!trsv_use_sparse(n, nblk, stats[0])) return;      // uniform over the cluster
    const int nops = stats[1];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int cl = blockIdx.x / kBandCluster;
    const int kbeg = phase == 0 ? bp.k0[cl] : bp.ksep;
    const int kend = phase == 0 ? bp.k1[cl] : nblk;
    const int ksep = phase == 0 ? bp.ksep : nblk;         // phase 0: blocks >= ksep belong to the separator
    if (kbeg >= kend) return;
    // my ops, in list order: an op (k, j) writes row block j.
    //   forward:  solved block k in my range; j in my range or (phase 0) in the separator
    //   backward: j in my range; k in my range or (phase 0) in the separator
    if (warp == 0) {
        int m = 0;
        for (int base = 0; base < nops; base += 32) {
            const int t = base + lane;
            const int op = t < nops ? list[t] : 0;
            const int k = op >> 16, j = op & 0xffff;
            const bool kin = k >= kbeg && k < kend, jin = j >= kbeg && j < kend;
            const bool sel = backward ? (jin && (kin || k >= ksep)) : (kin && (jin || j >= ksep));
            const bool mine = t < nops && sel && (j % kBandCluster) == rank;
            const unsigned bal = __ballot_sync(0xffffffffu, mine);
            if (mine) sm.ops[m + __popc(bal & ((1u << lane) - 1))] = op;
            m += __popc(bal);
        }
        if (lane == 0) sm.nmy = m;
    }
    for (int jb = rank; jb < nblk; jb += kBandCluster) {
        const bool own = jb >= kbeg && jb < kend;
        if (!own && jb < ksep) continue;
        for (int i = tid; i < NB; i += 256) sm.bown[jb / kBandCluster][i] = (own && jb * NB + i < n) ? b[jb * NB + i] : 0.0;   // separator rows: partial sums from zero
    }
    __syncthreads();
    const int nmy = sm.nmy;
    for (int i = 0; i < kTrsvClRing - 1; ++i) { if (i < nmy) trsv_cl_issue(sm.ring[i], sm.ops[i], n, L, ld, dinv, backward); cp_async_commit(); }
    int t = 0;
    // one op: the operand tile of op t is in ring[t % R]; x = the 64 multipliers; thread tid < 64 returns its dot product
    auto run_op = [&](const double* x) -> double {
        { const int nx = t + kTrsvClRing - 1; if (nx < nmy) trsv_cl_issue(sm.ring[nx % kTrsvClRing], sm.ops[nx], n, L, ld, dinv, backward); cp_async_commit(); }
        cp_async_wait<kTrsvClRing - 1>();
        __syncthreads();
        double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
        if (tid < NB) {
            const double* T = sm.ring[t % kTrsvClRing];
            if (!backward) {
#pragma unroll
                for (int c = 0; c < NB; c += 4) {
                    s0 += T[c * kTrsvClLDT + tid] * x[c]; s1 += T[(c + 1) * kTrsvClLDT + tid] * x[c + 1];
                    s2 += T[(c + 2) * kTrsvClLDT + tid] * x[c + 2]; s3 += T[(c + 3) * kTrsvClLDT + tid] * x[c + 3];
                }
            } else {
                const double2* T2 = reinterpret_cast<const double2*>(T + tid * kTrsvClLDT);
                const double2* x2 = reinterpret_cast<const double2*>(x);
#pragma unroll
                for (int c = 0; c < NB / 2; c += 2) {
                    const double2 a0 = T2[c], a1 = T2[c + 1], x0 = x2[c], x1 = x2[c + 1];
                    s0 += a0.x * x0.x; s1 += a0.y * x0.y; s2 += a1.x * x1.x; s3 += a1.y * x1.y;
                }
            }
        }
        return (s0 + s1) + (s2 + s3);
    };
    cluster.sync();                                           // every CTA of the cluster is running before the first remote store
    int sidx = 0;
    auto apply_column = [&](int k) {                          // my tiles of block column k, multipliers in yk[sidx & 1]
        const double* yk = sm.yk[sidx & 1];
        while (t < nmy && (sm.ops[t] >> 16) == k) {
            const int j = sm.ops[t] & 0xffff;
            const double d = run_op(yk);
            if (tid < NB) sm.bown[j / kBandCluster][tid] -= d;
            ++t;
            __syncthreads();                                  // the ring slot and b_j are settled before the next op touches them
        }
    };
    if (backward && phase == 0) {
        // the separator is solved: its blocks are plain multipliers, every CTA reads them from global memory itself
        for (int k = nblk - 1; k >= ksep; --k, ++sidx) {
            if (!(t < nmy && (sm.ops[t] >> 16) == k)) continue;   // no tile of mine in this column (CTA-uniform)
            if (tid < NB) sm.yk[sidx & 1][tid] = (k * NB + tid < n) ? __ldcg(b + k * NB + tid) : 0.0;
            __syncthreads();
            apply_column(k);
        }
        sidx = (sidx + 1) & ~1;
        cluster.sync();                                       // nobody still reads its local yk when the first remote push arrives
    }
    for (int s2 = 0; s2 < kend - kbeg; ++s2, ++sidx) {
        const int k = backward ? kend - 1 - s2 : kbeg + s2;
        if (t < nmy && sm.ops[t] == ((k << 16) | k)) {      // I own block k: every update of b_k has been applied (list order)
            double* bk = sm.bown[k / kBandCluster];
            const double y = run_op(bk);
            if (tid < NB) {
                if (k * NB + tid < n) b[k * NB + tid] = y;
#pragma unroll
                for (int dst = 0; dst < kBandCluster; ++dst) *cluster.map_shared_rank(&sm.yk[sidx & 1][tid], dst) = y;   // DSMEM push to the whole cluster
            }
            ++t;
            __syncthreads();
        }
        cluster_sync_all();                                   // y_k has arrived everywhere
        apply_column(k);
    }
    if (!backward && phase == 0) {                            // what this part subtracts from the separator's right-hand side
        for (int jb = rank; jb < nblk; jb += kBandCluster) {
            if (jb < ksep) continue;
            if (tid < NB && jb * NB + tid < n) { const double v = sm.bown[jb / kBandCluster][tid]; if (v != 0.0) atomicAdd(b + jb * NB + tid, v); }
        }
    }
}

constexpr int kTrsvOwn = 2;          // owned row blocks per CTA whose operands are kept in shared memory
constexpr int kTrsvFCacheMax = 65536;  // bytes of F cached in shared memory (n <= 16384)

__global__ void __launch_bounds__(256) k_trsv_flags(int n, const double* __restrict__ L, int64_t ld, const double* __restrict__ dinv, double* b,
                                                    double* ybuf, int* flags, const unsigned char* __restrict__ F, int epoch, int backward, int cacheF,
                                                    const int* __restrict__ nz_tiles) {
    if (trsv_use_sparse(n, (n + NB - 1) / NB, nz_tiles[0])) return;   // k_trsv_single does it
    extern __shared__ double smt[];
    double* sDi = smt;                          // [kTrsvOwn][64*64]  Linv of my blocks, column-major
    double* sLt = smt + kTrsvOwn * NB * NB;     // [kTrsvOwn][64*64]  the sub-diagonal tile next to each of my blocks, column-major
    unsigned char* sF = (unsigned char*)(smt + 2 * kTrsvOwn * NB * NB);
    __shared__ double bown[kTrsvOwn][NB];
    __shared__ double yk[NB];
    __shared__ double part[4][NB];
    const int tid = threadIdx.x;
    const int nblk = (n + NB - 1) / NB;
    const int G = gridDim.x, me = blockIdx.x;
    const int r = tid & 63, q = tid >> 6;
    const unsigned char* Fp = F;
    if (cacheF) {
        for (int e = tid; e < nblk * nblk; e += 256) sF[e] = F[e];
        Fp = sF;
    }
    // ---- prefetch: everything on the critical path of my blocks lives in shared memory before the chain reaches me
    for (int o = 0; o < kTrsvOwn; ++o) {
        const int so = me + o * G;
        if (so >= nblk) break;
        const int kb = backward ? nblk - 1 - so : so;
        const double* Di = dinv + (size_t)kb * NB * NB;
        for (int e = tid; e < NB * NB; e += 256) sDi[o * NB * NB + e] = Di[e];
        if (tid < NB) bown[o][tid] = (kb * NB + tid < n) ? b[kb * NB + tid] : 0.0;
        if (so >= 1) {
            const int lo = backward ? kb : kb - 1;          // the tile L(rows lo+1, cols lo)
            for (int e = tid; e < NB * NB; e += 256) {
                const int c = e >> 6, rr = e & 63;
                const int row = (lo + 1) * NB + rr, colg = lo * NB + c;
                sLt[o * NB * NB + e] = (row < n) ? L[(size_t)colg * ld + row] : 0.0;
            }
        }
    }
    __syncthreads();

    for (int s = 0; s < nblk; ++s) {
        const int kb = backward ? nblk - 1 - s : s;
        const int k0 = kb * NB;
        const int nbk = min(NB, n - k0);
        const bool owner = (s % G) == me;
        int s2_first = s + ((me - s) % G + G) % G;
        if (s2_first == s) s2_first += G;
        bool any_tile = false;
        for (int s2 = s2_first; s2 < nblk; s2 += G) {
            const int jb = backward ? nblk - 1 - s2 : s2;
            any_tile |= (backward ? Fp[(size_t)jb * nblk + kb] : Fp[(size_t)kb * nblk + jb]) != 0;
        }
        if (owner) {
            const int o = (s - me) / G;
            const double* Di = o < kTrsvOwn ? sDi + o * NB * NB : dinv + (size_t)kb * NB * NB;   // Di[c*NB + r] = Linv(r, c)
            if (o >= kTrsvOwn) {
                if (tid < NB) yk[tid] = tid < nbk ? __ldcg(b + k0 + tid) : 0.0;
                __syncthreads();
            }
            const double* bk = o < kTrsvOwn ? bown[o] : yk;
            double sacc = 0.0;
            if (!backward) { for (int c = q; c <= r; c += 4) sacc += Di[(size_t)c * NB + r] * bk[c]; }                  // y_r = sum_c Linv(r,c) b_c
            else { for (int c = r + ((q - r) & 3); c < NB; c += 4) sacc += Di[(size_t)r * NB + c] * bk[c]; }           // x_r = sum_c Linv(c,r) b_c
            part[q][r] = sacc;
            __syncthreads();
            if (tid < NB) {
                const double v = (part[0][tid] + part[1][tid]) + (part[2][tid] + part[3][tid]);
                yk[tid] = v;
                if (tid < nbk) { __stcg(ybuf + k0 + tid, v); __stcg(b + k0 + tid, v); }
            }
            __syncthreads();                       // orders the 64 stores before the release below (cumulativity through the barrier)
            if (tid == 0) st_release(flags + s, epoch);
        } else {
            if (!any_tile) continue;   // uniform across the CTA
            if (tid == 0) { while (ld_acquire(flags + s) != epoch) {} }
            __syncthreads();
            if (tid < NB) yk[tid] = tid < nbk ? __ldcg(ybuf + k0 + tid) : 0.0;
            __syncthreads();
        }
        for (int s2 = s2_first; s2 < nblk; s2 += G) {
            const int jb = backward ? nblk - 1 - s2 : s2;
            if ((backward ? Fp[(size_t)jb * nblk + kb] : Fp[(size_t)kb * nblk + jb]) == 0) continue;
            const int o2 = (s2 - me) / G;
            const int j0 = jb * NB;
            const int nbj = min(NB, n - j0);
            double sacc = 0.0;
            if (s2 == s + 1 && o2 < kTrsvOwn) {    // the sub-diagonal tile, prefetched: sLt[c*64 + rr] = L(lo+1 rows, lo cols)
                const double* T = sLt + o2 * NB * NB;
                if (!backward) { for (int c = q; c < NB; c += 4) sacc += T[c * NB + r] * yk[c]; }        // b_j[r] -= sum_c L(j0+r, k0+c) y_c
                else { for (int c = q; c < NB; c += 4) sacc += T[r * NB + c] * yk[c]; }                  // b_j[r] -= sum_c L(k0+c, j0+r) x_c
            } else if (r < nbj) {
                if (!backward) { for (int c = q; c < nbk; c += 4) sacc += L[(size_t)(k0 + c) * ld + j0 + r] * yk[c]; }
                else { for (int c = q; c < nbk; c += 4) sacc += L[(size_t)(j0 + r) * ld + k0 + c] * yk[c]; }
            }
            part[q][r] = sacc;
            __syncthreads();
            if (tid < nbj) {
                const double dlt = (part[0][tid] + part[1][tid]) + (part[2][tid] + part[3][tid]);
                if (o2 < kTrsvOwn) bown[o2][tid] -= dlt;
                else __stcg(b + j0 + tid, __ldcg(b + j0 + tid) - dlt);
            }
            __syncthreads();
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Mirror the lower triangle into the upper one, 32x32 tiles through shared memory (both sides coalesced).
__global__ void __launch_bounds__(256) k_mirror_lower(int n, double* __restrict__ A, int64_t ld) {
    __shared__ double t[32][33];
    const int ti = blockIdx.y, tj = blockIdx.x;   // tile row / tile col; only ti >= tj does work
    if (ti < tj) return;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
    for (int cc = ty; cc < 32; cc += 8) {
        int r = ti * 32 + tx, c = tj * 32 + cc;
        t[cc][tx] = (r < n && c < n) ? A[(size_t)c * ld + r] : 0.0;
    }
    __syncthreads();
    for (int cc = ty; cc < 32; cc += 8) {
        // destination element (row = tj*32 + tx, col = ti*32 + cc) = source (row = ti*32 + cc, col = tj*32 + tx)
        int r = tj * 32 + tx, c = ti * 32 + cc;
        if (r < n && c < n && c > r) A[(size_t)c * ld + r] = t[tx][cc];
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// r = b - A*x for symmetric A with BOTH triangles stored (row i is read as the contiguous column i), accumulated in
// double-double (two-prod / two-sum), one warp per row.  This is the extended-precision residual of the refinement.
__device__ __forceinline__ void dd_add(double& hi, double& lo, double bh, double bl) {
    double s = __dadd_rn(hi, bh);
    double bb = __dadd_rn(s, -hi);
    double e = __dadd_rn(__dadd_rn(hi, -__dadd_rn(s, -bb)), __dadd_rn(bh, -bb));
    e = __dadd_rn(e, __dadd_rn(lo, bl));
    hi = __dadd_rn(s, e);
    lo = __dadd_rn(e, -__dadd_rn(hi, -s));
}
__global__ void __launch_bounds__(256) k_residual_dd(int n, const double* __restrict__ A, int64_t ld, const double* __restrict__ x,
                                                     const double* __restrict__ b, double* __restrict__ r) {
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= n) return;
    const double* col = A + (size_t)i * ld;
    double hi = 0.0, lo = 0.0;
    for (int j = lane; j < n; j += 32) {
        double a = col[j], xv = x[j];
        double p = __dmul_rn(a, xv);
        double e = __fma_rn(a, xv, -p);
        dd_add(hi, lo, p, e);
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
        double oh = __shfl_xor_sync(0xffffffffu, hi, s), ol = __shfl_xor_sync(0xffffffffu, lo, s);
        dd_add(hi, lo, oh, ol);
    }
    if (lane == 0) {
        double rh = b[i], rl = 0.0;
        dd_add(rh, rl, -hi, -lo);
        r[i] = rh;
    }
}
__global__ void k_axpy1(int n, const double* __restrict__ d, double* __restrict__ x) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[i] += d[i];
}

// ---------------------------------------------------------------------------------------------------------------------
// Strip triangular solve  X <- X * L^-T  for a panel of W <= 64*NBLK columns: X is (rows x W), L the W x W lower-triangular factor of
// the panel's diagonal block, dinv the stored inverses of its 64x64 diagonal blocks.  The launch-per-operation form of this (a 64-wide
// right solve, then a K = 64 product for the rest of the panel, per block column) is a chain of 2*W/64 launches of ~20 us each that
// do ~9 us of arithmetic.  Here ONE CTA owns a strip of 48 rows and walks the whole panel: the strip lives in the accumulator
// registers (warp w holds columns 8w..8w+7 of every 64-column block, 6 row fragments: 96 doubles per thread), the solved block
// column goes through shared memory as the A operand of the products that follow, and the tiles of L and the diagonal inverses stream
// through a three-slot cp.async ring in the order they are needed:  D_0, L_10 .. L_(n-1)0, D_1, L_21 ..   48 rows x 126 CTAs cover the
// 6013 rows of the EKF gain with one wave on 148 SMs.
constexpr int kStripRF = 6;
constexpr int kStripRows = 8 * kStripRF;
constexpr int kStripXS = kStripRows + 4;       // [k][row] stride, 52 = 4 (mod 16): conflict-free fragment loads
struct StripSmem {
    double ring[3][NB * kBandTS];              // tiles of L / diagonal inverses, [k][column]
    double xin[NB * kStripXS];                 // the current block column before its diagonal solve, [k][row]
    double xout[NB * kStripXS];                // ... and solved
};
template <int NBLK>
__global__ void __launch_bounds__(256, 1) k_strip_trsm(int rows, int W, double* __restrict__ X, int64_t ldx, const double* __restrict__ L, int64_t ldl,
                                                       const double* __restrict__ dinv) {
    extern __shared__ __align__(16) unsigned char strip_raw[];
    StripSmem& sm = *reinterpret_cast<StripSmem*>(strip_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, tg = lane & 3;
    const int r0 = blockIdx.x * kStripRows;
    const int nblk = (W + NB - 1) / NB;
    // tile stream: (j, jp) with jp == j the diagonal inverse of block j, jp > j the tile L(jp, j)
    int ij = 0, ijp = 0;
    auto issue_next = [&](int slot) {
        if (ij < nblk) {
            double* dst = sm.ring[slot];
            const double* src; int64_t stride; int lim;
            if (ijp == ij) { src = dinv + (size_t)ij * NB * NB; stride = NB; lim = NB; }
            else { src = L + (size_t)(ij * NB) * ldl + ijp * NB; stride = ldl; lim = W - ijp * NB; }      // rows of L past W: zero
            const int cp = (tid & 31) * 2;
            int bytes = (lim - cp) * 8; bytes = bytes > 16 ? 16 : (bytes < 0 ? 0 : bytes);
#pragma unroll
            for (int it = 0; it < 8; ++it) {
                const int k = (tid >> 5) + 8 * it;
                cp_async16(dst + k * kBandTS + cp, bytes > 0 ? src + (size_t)k * stride + cp : src, bytes);
            }
            if (++ijp >= nblk) { ++ij; ijp = ij; }
        }
        cp_async_commit();
    };
    double acc[kStripRF][NBLK][2];
#pragma unroll
    for (int jb = 0; jb < NBLK; ++jb)
#pragma unroll
        for (int i = 0; i < kStripRF; ++i)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int row = r0 + 8 * i + g, col = jb * NB + 8 * warp + 2 * tg + e;
                acc[i][jb][e] = (row < rows && col < W) ? X[(size_t)col * ldx + row] : 0.0;
            }
    issue_next(0);
    issue_next(1);
    int t = 0;
#pragma unroll
    for (int j = 0; j < NBLK; ++j) {
        if (j >= nblk) break;
        // block column j leaves the registers: operand of its own diagonal solve
#pragma unroll
        for (int i = 0; i < kStripRF; ++i)
#pragma unroll
            for (int e = 0; e < 2; ++e) { sm.xin[(8 * warp + 2 * tg + e) * kStripXS + 8 * i + g] = acc[i][j][e]; acc[i][j][e] = 0.0; }
        cp_async_wait<1>();
        __syncthreads();                               // tile t (= D_j) has landed, xin is complete, every warp is done with tile t-1
        issue_next((t + 2) % 3);
        {
            const double* D = sm.ring[t % 3];
#pragma unroll 2
            for (int ks = 0; ks < NB; ks += 4) {       // X_j(r, c) = sum_k xin(r, k) Linv(c, k)
                const double bf = D[(ks + tg) * kBandTS + 8 * warp + g];
#pragma unroll
                for (int i = 0; i < kStripRF; ++i) dmma_m8n8k4(acc[i][j][0], acc[i][j][1], sm.xin[(ks + tg) * kStripXS + 8 * i + g], bf);
            }
        }
#pragma unroll
        for (int i = 0; i < kStripRF; ++i)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int row = r0 + 8 * i + g, col = j * NB + 8 * warp + 2 * tg + e;
                sm.xout[(8 * warp + 2 * tg + e) * kStripXS + 8 * i + g] = -acc[i][j][e];      // negated: the products below subtract
                if (row < rows && col < W) X[(size_t)col * ldx + row] = acc[i][j][e];
            }
        ++t;
        __syncthreads();
#pragma unroll
        for (int jp = j + 1; jp < NBLK; ++jp) {
            if (jp >= nblk) break;
            cp_async_wait<1>();
            __syncthreads();
            issue_next((t + 2) % 3);
            const double* T = sm.ring[t % 3];
#pragma unroll 2
            for (int ks = 0; ks < NB; ks += 4) {       // X_jp(r, c) -= sum_k X_j(r, k) L(jp*64 + c, j*64 + k)
                const double bf = T[(ks + tg) * kBandTS + 8 * warp + g];
#pragma unroll
                for (int i = 0; i < kStripRF; ++i) dmma_m8n8k4(acc[i][jp][0], acc[i][jp][1], sm.xout[(ks + tg) * kStripXS + 8 * i + g], bf);
            }
            ++t;
        }
    }
    cp_async_wait<0>();
}
void launch_strip_trsm(cudaStream_t st, int rows, int W, double* X, int64_t ldx, const double* L, int64_t ldl, const double* dinv) {
    if (rows <= 0 || W <= 0) return;
    static PerDeviceOnce once;
    if (once.first()) {
        cudaFuncSetAttribute(k_strip_trsm<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(StripSmem));
        cudaFuncSetAttribute(k_strip_trsm<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(StripSmem));
    }
    const int grid = (rows + kStripRows - 1) / kStripRows;
    if (W <= 4 * NB) k_strip_trsm<4><<<grid, 256, sizeof(StripSmem), st>>>(rows, W, X, ldx, L, ldl, dinv);
    else k_strip_trsm<8><<<grid, 256, sizeof(StripSmem), st>>>(rows, W, X, ldx, L, ldl, dinv);
}

// ---------------------------------------------------------------------------------------------------------------------
constexpr size_t kPanelSolveSmem = sizeof(double) * (NB * NB + NB * PS_ROWS);

static int g_coop_blocks = 0;
static int g_potrf_rowops = 0;
static int g_sms = 148;
static void launch_syrk(cudaStream_t st, int n, double* A, int64_t ld, int kcol0, int K, int origin, int col_end, const unsigned char* F, int nblk);
// substitution epoch of a factor: the flags of k_trsv_flags live in the factor's own workspace, so the counter is kept per workspace
// (handles that alternate, or run on different host threads, do not see each other's epochs)
static std::mutex g_epoch_mu;
static std::unordered_map<const void*, int> g_epochs;
static void epoch_reset(const void* ws) { std::lock_guard<std::mutex> g(g_epoch_mu); if (g_epochs.size() > 4096) g_epochs.clear(); g_epochs[ws] = 0; }
static int epoch_next(const void* ws) { std::lock_guard<std::mutex> g(g_epoch_mu); return ++g_epochs[ws]; }
static void set_attrs_once() {
    static PerDeviceOnce once;
    if (!once.first()) return;
    cudaFuncSetAttribute(k_panel_solve, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPanelSolveSmem);
    cudaFuncSetAttribute(k_potrf64_inv, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(double) * kPotrfScratchDoubles));
    { const char* e = getenv("SRK_POTRF"); g_potrf_rowops = (e != nullptr && e[0] == 'r') ? 1 : 0; }
    cudaFuncSetAttribute(k_right_solve_dmma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kRightSolveSmem);
    cudaFuncSetAttribute(k_syrk_dmma<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SyrkCfg<128>::kSmem);
    cudaFuncSetAttribute(k_syrk_dmma<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SyrkCfg<64>::kSmem);
    g_sms = 148; { int d = 0; cudaGetDevice(&d); cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, d); }
    int dev = 0, sms = 148, per_sm = 1;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaFuncSetAttribute(k_trsv_flags, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(double) * 2 * kTrsvOwn * NB * NB + kTrsvFCacheMax));
    g_coop_blocks = sms * (per_sm < 1 ? 1 : 1);   // one CTA per SM: fewer spinners, same bandwidth
}

// Workspace layout (doubles): [dinv: nblk*64*64][ybuf: nblk*64][flags: nblk ints, padded][F: nblk*nblk bytes, padded][op lists: 2 x 7*nblk ints][tile pattern of one panel's diagonal block: 64 bytes]
constexpr size_t kFPanelBytes = 64;     // (PB / NB)^2 = 16 used
static inline int chol_nblk(int n) { return (n + NB - 1) / NB; }
size_t dense_cholesky_dinv_doubles(int n) {
    const size_t nblk = (size_t)chol_nblk(n);
    return nblk * NB * NB + nblk * NB + (nblk + 1) / 2 + 8 + (nblk * nblk + 7) / 8 + 8 + (size_t)(kTrsvSparseFill + 1) * nblk + 8 + kFPanelBytes / 8;   // the +8 after the flags also holds the tile / op counters
}
static inline double* ws_ybuf(double* ws, int n) { return ws + (size_t)chol_nblk(n) * NB * NB; }
static inline int* ws_flags(double* ws, int n) { return (int*)(ws_ybuf(ws, n) + (size_t)chol_nblk(n) * NB); }
static inline int* ws_nzt(double* ws, int n) { return ws_flags(ws, n) + chol_nblk(n) + 1; }
static inline unsigned char* ws_F(double* ws, int n) { return (unsigned char*)(ws_ybuf(ws, n) + (size_t)chol_nblk(n) * NB + (chol_nblk(n) + 1) / 2 + 8); }
static inline unsigned char* ws_fpanel(double* ws, int n) {
    const size_t nblk = (size_t)chol_nblk(n);
    return (unsigned char*)(ws_ybuf(ws, n) + nblk * NB + (nblk + 1) / 2 + 8 + (nblk * nblk + 7) / 8 + 8 + ((kTrsvSparseFill + 1) * nblk * 2 + 1) / 2 + 8);
}
static inline int* ws_list(double* ws, int n, int backward) {
    const size_t nblk = (size_t)chol_nblk(n);
    int* base = (int*)(ws_ybuf(ws, n) + nblk * NB + (nblk + 1) / 2 + 8 + (nblk * nblk + 7) / 8 + 8);
    return base + (backward ? (size_t)(kTrsvSparseFill + 1) * nblk : 0);
}

// Development aid: SRK_CHOL_PROFILE=1 times every kernel type of the factorisation with events (serialised, no graph).
static bool g_prof = false;
static double g_prof_ms[4] = {0, 0, 0, 0};
static int g_prof_n[4] = {0, 0, 0, 0};
struct ProfScope {
    int k; cudaStream_t st; cudaEvent_t a, b;
    ProfScope(int kind, cudaStream_t s) : k(kind), st(s) { if (g_prof) { cudaEventCreate(&a); cudaEventCreate(&b); cudaEventRecord(a, st); } }
    ~ProfScope() { if (g_prof) { cudaEventRecord(b, st); cudaEventSynchronize(b); float ms = 0; cudaEventElapsedTime(&ms, a, b); g_prof_ms[k] += ms; g_prof_n[k]++; cudaEventDestroy(a); cudaEventDestroy(b); } }
};
void dense_cholesky_profile_report() {
    const char* names[4] = {"potrf64_inv", "panel_solve", "syrk_inpanel", "syrk_trailing"};
    for (int i = 0; i < 4; ++i) if (g_prof_n[i]) printf("  %-14s n=%4d total %8.3f ms avg %7.1f us\n", names[i], g_prof_n[i], g_prof_ms[i], 1e3 * g_prof_ms[i] / g_prof_n[i]);
    for (int i = 0; i < 4; ++i) { g_prof_ms[i] = 0; g_prof_n[i] = 0; }
}

static bool ensure_band_attr();
// Panel-wise dense factorisation (default): per 256-column panel THREE dependent steps instead of 4 x (potrf, right solve, in-panel
// update) + trailing update --
//   A  the 256 x 256 diagonal block is factored by one cluster (k_band_chol on the block as a 4-tile dense system: potrf look-ahead,
//      tile solves and updates spread over 8 CTAs, no launch in between);
//   B  the rows below are solved against it in one launch (k_strip_trsm, 48-row strips that walk the whole panel);
//   C  trailing update with K = 256 (k_syrk_dmma).
// Every tile of a dense factor is non-zero: F is set to ones (the substitutions and k_syrk_dmma read it).
static int64_t enqueue_factor_panels(cudaStream_t st, int n, double* A, int64_t ld, double* ws, int* info_dev) {
    int64_t launches = 0;
    const int nblk = chol_nblk(n);
    unsigned char* F = ws_F(ws, n);
    unsigned char* Fp = ws_fpanel(ws, n);
    cudaMemsetAsync(info_dev, 0, sizeof(int), st);
    cudaMemsetAsync(ws_flags(ws, n), 0, sizeof(int) * nblk, st);
    cudaMemsetAsync(F, 1, (size_t)nblk * nblk, st);
    cudaMemsetAsync(Fp, 1, kFPanelBytes, st);
    CholPartition whole;
    for (int p0 = 0; p0 < n; p0 += PB) {
        const int pend = min(n, p0 + PB), W = pend - p0;
        double* di = ws + (size_t)(p0 / NB) * NB * NB;
        double* App = A + (size_t)p0 * ld + p0;
        k_band_chol<<<kBandCluster, 256, kBandSmemRequest, st>>>(W, App, ld, di, info_dev, Fp, chol_nblk(W), whole, 1, p0); ++launches;
        if (pend < n) {
            launch_strip_trsm(st, n - pend, W, A + (size_t)p0 * ld + pend, ld, App, ld, di); ++launches;
            launch_syrk(st, n, A, ld, p0, W, pend, n, F, nblk); launches += 2;
        }
    }
    k_build_trsv_lists<<<1, 1024, 0, st>>>(n, nblk, F, ws_nzt(ws, n), ws_list(ws, n, 0), ws_list(ws, n, 1)); ++launches;
    return launches;
}

// Factorisation of the m x m matrix S on the stream `hi` with the triangular solve Z <- Z * L^-T (Z: rows x m) following it panel by panel on
// the stream `lo` (EKF: S = innovation covariance, Z = P H^T).  The factorisation is a latency chain that leaves most SMs idle (one cluster
// for the diagonal block, <= 78 CTAs for the strip solve), the solve is throughput work on every SM and needs of L only the panels that are
// already final: run one after the other they cost 3.4 + 4.3 ms at m = 4000, rows = 6013.  `hi` should have a higher priority than `lo`; the
// products of the solve are launched one CTA per tile so that the chain's kernels get SMs at tile granularity.  ev: >= m / 256 + 2 events.
// On return `lo` has waited for everything enqueued on `hi`.
int64_t dense_cholesky_factor_trsm(cudaStream_t hi, cudaStream_t lo, int m, double* S, int64_t lds, double* ws, int* info_dev, int rows, double* Z, int64_t ldz,
                                   cudaEvent_t* ev, int nev) {
    set_attrs_once();
    epoch_reset(ws);
    const int npanel = (m + PB - 1) / PB;
    if (nev < npanel + 2 || !ensure_band_attr() || ((uintptr_t)S & 15) != 0 || (lds & 1) != 0) return -1;
    int64_t launches = 0;
    const int nblk = chol_nblk(m);
    unsigned char* F = ws_F(ws, m);
    unsigned char* Fp = ws_fpanel(ws, m);
    cudaEventRecord(ev[npanel], lo);                       // S and Z are produced on `lo`
    cudaStreamWaitEvent(hi, ev[npanel], 0);
    cudaMemsetAsync(info_dev, 0, sizeof(int), hi);
    cudaMemsetAsync(ws_flags(ws, m), 0, sizeof(int) * nblk, hi);
    cudaMemsetAsync(F, 1, (size_t)nblk * nblk, hi);
    cudaMemsetAsync(Fp, 1, kFPanelBytes, hi);
    CholPartition whole;
    constexpr int kZPanel = 2 * PB;                        // the solve walks 512-column panels = two panels of the factorisation
    for (int p0 = 0, p = 0; p0 < m; p0 += PB, ++p) {
        const int pend = min(m, p0 + PB), W = pend - p0;
        double* di = ws + (size_t)(p0 / NB) * NB * NB;
        double* App = S + (size_t)p0 * lds + p0;
        k_band_chol<<<kBandCluster, 256, kBandSmemRequest, hi>>>(W, App, lds, di, info_dev, Fp, chol_nblk(W), whole, 1, p0); ++launches;
        if (pend < m) { launch_strip_trsm(hi, m - pend, W, S + (size_t)p0 * lds + pend, lds, App, lds, di); ++launches; }
        cudaEventRecord(ev[p], hi);                        // columns [p0, pend) of L are final
        if (pend < m) { launch_syrk(hi, m, S, lds, p0, W, pend, m, F, nblk); launches += 2; }
        if ((p & 1) == 1 || pend == m) {
            const int q0 = (p0 / kZPanel) * kZPanel;
            cudaStreamWaitEvent(lo, ev[p], 0);
            launch_strip_trsm(lo, rows, pend - q0, Z + (size_t)q0 * ldz, ldz, S + (size_t)q0 * lds + q0, lds, ws + (size_t)(q0 / NB) * NB * NB); ++launches;
            if (pend < m) {
                launch_gemm_nt_dmma(lo, rows, m - pend, pend - q0, Z + (size_t)q0 * ldz, ldz, S + (size_t)q0 * lds + pend, lds, Z + (size_t)pend * ldz, ldz, 0, 0, 1);
                ++launches;
            }
        }
    }
    k_build_trsv_lists<<<1, 1024, 0, hi>>>(m, nblk, F, ws_nzt(ws, m), ws_list(ws, m, 0), ws_list(ws, m, 1)); ++launches;
    cudaEventRecord(ev[npanel + 1], hi);
    cudaStreamWaitEvent(lo, ev[npanel + 1], 0);
    return launches;
}

static int64_t enqueue_factor(cudaStream_t st, int n, double* A, int64_t ld, double* ws, int* info_dev) {
    static int panels = -1;      // SRK_CHOL_DENSE=steps: the launch-per-64-columns form (development aid; also the profiled one)
    if (panels < 0) { const char* e = getenv("SRK_CHOL_DENSE"); panels = (e != nullptr && e[0] == 's') ? 0 : 1; }
    if (panels && !g_prof && n >= 2 * NB && ((uintptr_t)A & 15) == 0 && (ld & 1) == 0 && ensure_band_attr()) return enqueue_factor_panels(st, n, A, ld, ws, info_dev);
    int64_t launches = 0;
    const int nblk = chol_nblk(n);
    unsigned char* F = ws_F(ws, n);
    cudaMemsetAsync(info_dev, 0, sizeof(int), st);
    cudaMemsetAsync(ws_flags(ws, n), 0, sizeof(int) * nblk, st);
    cudaMemsetAsync(F, 0, (size_t)nblk * nblk, st);
    for (int p0 = 0; p0 < n; p0 += PB) {
        const int pend = min(n, p0 + PB);
        for (int k0 = p0; k0 < pend; k0 += NB) {
            double* di = ws + (size_t)(k0 / NB) * NB * NB;
            { ProfScope ps(0, st); k_potrf64_inv<<<1, 256, sizeof(double) * kPotrfScratchDoubles, st>>>(n, k0, A, ld, di, info_dev, F, nblk, g_potrf_rowops); } ++launches;
            const int below = n - (k0 + NB);
            if (below <= 0) continue;
            { ProfScope ps(1, st); k_right_solve_dmma<<<(below + RS_ROWS - 1) / RS_ROWS, 256, kRightSolveSmem, st>>>(n, k0 + NB, A + (size_t)k0 * ld, ld, di, F + (size_t)(k0 / NB) * nblk, nblk, NB); } ++launches;
            const int origin = k0 + NB;
            if (origin < pend) {   // rest of the current panel, K = 64
                { ProfScope ps(2, st); launch_syrk(st, n, A, ld, k0, NB, origin, pend, F, nblk); } launches += 2;
            }
        }
        if (pend < n) {            // trailing matrix, K = panel width (pend - p0 == PB whenever pend < n)
            { ProfScope ps(3, st); launch_syrk(st, n, A, ld, p0, pend - p0, pend, n, F, nblk); } launches += 2;
        }
    }
    k_build_trsv_lists<<<1, 1024, 0, st>>>(n, nblk, F, ws_nzt(ws, n), ws_list(ws, n, 0), ws_list(ws, n, 1)); ++launches;
    return launches;
}

// The launch sequence of a factorisation is static for given (n, buffers): it is captured once into a CUDA graph and
// replayed, which removes ~800 host launch calls per solve from the critical path.
struct FactorGraph { int n = 0; double* A = nullptr; int64_t ld = 0; double* ws = nullptr; int* info = nullptr; cudaGraphExec_t exec = nullptr; int64_t launches = 0; };
static thread_local FactorGraph g_fg;   // per host thread: a handle is driven by one thread at a time (ba_c_api.h threading contract)

// Sparse-factor path: tile pattern of the input, then (when it is sparse enough) the single-cluster kernel.  Returns the number of
// launches, or 0 when the matrix is not sparse and the launch-per-operation path has to run.
static bool ensure_band_attr() {
    static PerDeviceOnce once;
    if (once.first()) {
        if (cudaFuncSetAttribute(k_band_chol, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBandSmemRequest) != cudaSuccess) { cudaGetLastError(); once.forget(); return false; }
    }
    return true;
}
// the second level of a two-level partition seen as a one-level one: parts = the second-level separators, separator = the top one
static CholPartition mid_level(const CholPartition& p) {
    CholPartition m;
    m.nparts = p.nmids; m.ksep = p.msep;
    for (int i = 0; i < p.nmids; ++i) { m.k0[i] = p.m0[i]; m.k1[i] = p.m1[i]; }
    return m;
}
static int64_t try_band_factor(cudaStream_t st, int n, double* A, int64_t ld, double* ws, int* info_dev, const CholPartition* part, const unsigned char* pattern_dev,
                               int pattern_count) {
    const int nblk = chol_nblk(n);
    if (nblk < 2 || nblk > kMaxRowBlocks) return 0;
    static int force = -1;    // SRK_CHOL_PATH=dense disables the sparse path, =band forces it (development aid)
    if (force < 0) { const char* e = getenv("SRK_CHOL_PATH"); force = e == nullptr ? 0 : (e[0] == 'd' ? 1 : (e[0] == 'b' ? 2 : 0)); }
    if (force == 1) return 0;
    if (!ensure_band_attr()) return 0;
    // The kernel is a latency chain: two CTAs of different clusters on one SM would take turns on every step.  Asking for more than half
    // of the SM's shared memory keeps it at one CTA per SM (SRK_BAND_SMEM_PAD=0: the bare request, development aid).
    static int pad = -1;
    if (pad < 0) { const char* e = getenv("SRK_BAND_SMEM_PAD"); pad = (e != nullptr && e[0] == '0') ? 0 : 1; }
    const size_t band_smem = pad ? kBandSmemRequest : sizeof(BandSmem);
    unsigned char* F = ws_F(ws, n);
    int* cnt = ws_nzt(ws, n);
    cudaMemsetAsync(info_dev, 0, sizeof(int), st);
    cudaMemsetAsync(ws_flags(ws, n), 0, sizeof(int) * nblk, st);
    cudaMemsetAsync(F, 0, (size_t)nblk * nblk, st);
    cudaMemsetAsync(cnt, 0, sizeof(int) * 2, st);
    int h = -1;
    int64_t nl = 2;
    if (pattern_dev != nullptr) {   // the caller knows the tile structure (solve_order.h): no pass over the matrix, no host round trip
        if (cudaMemcpyAsync(F, pattern_dev, (size_t)nblk * nblk, cudaMemcpyDeviceToDevice, st) != cudaSuccess) { cudaGetLastError(); return 0; }
        h = pattern_count; nl = 1;
    } else {
        k_tile_pattern<<<dim3(nblk, nblk), 256, 0, st>>>(n, A, ld, nblk, F, cnt);
        if (cudaMemcpyAsync(&h, cnt, sizeof(int), cudaMemcpyDeviceToHost, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess) { cudaGetLastError(); return 0; }
    }
    const int max_fill = (part != nullptr && part->nparts > 0) ? 2 * kBandMaxFill : kBandMaxFill;   // a partitioned pattern carries the separator rows as well
    if (h < 0 || (force != 2 && h > max_fill * nblk)) return 0;
    if (part != nullptr && part->nparts > 0) {
        k_band_chol<<<kBandCluster * part->nparts, 256, band_smem, st>>>(n, A, ld, ws, info_dev, F, nblk, *part, 0, 0);
        CholPartition top = *part;
        if (part->nmids > 0) {      // second-level separators: the same kernel with their ranges as the parts and the top separator as the shared block
            const CholPartition mids = mid_level(*part);
            k_band_chol<<<kBandCluster * mids.nparts, 256, band_smem, st>>>(n, A, ld, ws, info_dev, F, nblk, mids, 0, 0);
            top.ksep = part->msep;
            nl += 1;
        }
        k_band_chol<<<kBandCluster, 256, band_smem, st>>>(n, A, ld, ws, info_dev, F, nblk, top, 1, 0);
        nl += 2;
    } else {
        CholPartition whole;
        k_band_chol<<<kBandCluster, 256, band_smem, st>>>(n, A, ld, ws, info_dev, F, nblk, whole, 1, 0);
        nl += 1;
    }
    if (cudaGetLastError() != cudaSuccess) return 0;
    k_build_trsv_lists<<<1, 1024, 0, st>>>(n, nblk, F, ws_nzt(ws, n), ws_list(ws, n, 0), ws_list(ws, n, 1));
    return nl;
}

// true when a factorisation with this caller-provided pattern is certain to take the sparse-factor (cluster) path
bool dense_cholesky_pattern_ok(int n, int pattern_count, bool partitioned) {
    const int nblk = chol_nblk(n);
    const char* e = getenv("SRK_CHOL_PATH");
    if (e != nullptr && e[0] == 'd') return false;
    const char* pe = getenv("SRK_CHOL_PROFILE");
    if (pe != nullptr && pe[0] == '1') return false;
    if (nblk < 2 || nblk > kMaxRowBlocks) return false;
    return pattern_count <= (partitioned ? 2 * kBandMaxFill : kBandMaxFill) * nblk;
}
int64_t dense_cholesky_factor(cudaStream_t st, int n, double* A, int64_t ld, double* ws, int* info_dev, const CholPartition* part, const unsigned char* pattern_dev,
                              int pattern_count) {
    set_attrs_once();
    epoch_reset(ws);
    static int prof_env = -1;
    if (prof_env < 0) { const char* e = getenv("SRK_CHOL_PROFILE"); prof_env = (e != nullptr && e[0] == '1') ? 1 : 0; }
    g_prof = prof_env == 1;
    if (g_prof) return enqueue_factor(st, n, A, ld, ws, info_dev);
    {
        cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
        cudaStreamIsCapturing(st, &cs);
        if (cs == cudaStreamCaptureStatusNone) {   // the pattern read-back synchronises: not inside somebody's capture
            const int64_t nl = try_band_factor(st, n, A, ld, ws, info_dev, part, pattern_dev, pattern_count);
            if (nl > 0) return nl;
        }
    }
    if (g_fg.exec != nullptr && g_fg.n == n && g_fg.A == A && g_fg.ld == ld && g_fg.ws == ws && g_fg.info == info_dev) {
        if (cudaGraphLaunch(g_fg.exec, st) == cudaSuccess) return g_fg.launches;
        cudaGetLastError();
    }
    if (g_fg.exec != nullptr) { cudaGraphExecDestroy(g_fg.exec); g_fg.exec = nullptr; }
    cudaGraph_t graph = nullptr;
    if (chol_nblk(n) >= 8 && cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
        const int64_t launches = enqueue_factor(st, n, A, ld, ws, info_dev);
        if (cudaStreamEndCapture(st, &graph) == cudaSuccess && graph != nullptr) {
            cudaGraphExec_t exec = nullptr;
            if (cudaGraphInstantiate(&exec, graph, 0) == cudaSuccess) {
                cudaGraphDestroy(graph);
                g_fg.n = n; g_fg.A = A; g_fg.ld = ld; g_fg.ws = ws; g_fg.info = info_dev; g_fg.exec = exec; g_fg.launches = launches;
                if (cudaGraphLaunch(exec, st) == cudaSuccess) return launches;
            } else {
                cudaGraphDestroy(graph);
            }
        }
        cudaGetLastError();
    }
    return enqueue_factor(st, n, A, ld, ws, info_dev);
}

static void launch_syrk(cudaStream_t st, int n, double* A, int64_t ld, int kcol0, int K, int origin, int col_end, const unsigned char* F, int nblk) {
    k_syrk_dmma<128><<<g_sms, SyrkCfg<128>::kThreads, SyrkCfg<128>::kSmem, st>>>(n, A, ld, kcol0, K, origin, col_end, F, nblk);
    k_syrk_dmma<64><<<g_sms * 3, SyrkCfg<64>::kThreads, SyrkCfg<64>::kSmem, st>>>(n, A, ld, kcol0, K, origin, col_end, F, nblk);
}

// true when a factor with nz_tiles non-zero 64x64 tiles (diagonal included) is certain to take the sparse-factor substitution kernels
bool dense_cholesky_trsv_is_sparse(int n, int64_t nz_tiles) { return trsv_use_sparse(n, chol_nblk(n), (int)(nz_tiles > 0x7fffffff ? 0x7fffffff : nz_tiles)); }
static int64_t trsv(cudaStream_t st, int n, const double* L, int64_t ld, double* ws, double* b, int backward, const CholPartition* part, bool sparse_certain) {
    set_attrs_once();
    const int nblk = chol_nblk(n);
    int blocks = g_coop_blocks < nblk ? g_coop_blocks : nblk;
    if (blocks < 1) blocks = 1;
    const double* dinv = ws;
    double* ybuf = ws_ybuf(ws, n);
    int* flags = ws_flags(ws, n);
    const unsigned char* F = ws_F(ws, n);
    int epoch = epoch_next(ws);
    const int* nzt = ws_nzt(ws, n);
    if (n <= kTrsvSparseMaxN) {
        static PerDeviceOnce once;
        if (once.first()) cudaFuncSetAttribute(k_trsv_single, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(double) * kTrsvSparseMaxN + sizeof(int) * (kTrsvSparseFill + 1) * kTrsvSparseMaxBlk));
        static int use_cluster = -1;
        if (use_cluster < 0) {
            const char* e = getenv("SRK_TRSV_CLUSTER");
            use_cluster = (e != nullptr && e[0] == '0') ? 0 : 1;
        }
        static PerDeviceOnce once_cl;
        if (use_cluster && once_cl.first() && cudaFuncSetAttribute(k_trsv_cluster, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TrsvClSmem)) != cudaSuccess) { cudaGetLastError(); use_cluster = 0; }
        if (use_cluster) {
            const int* lst = ws_list(ws, n, backward);
            if (part != nullptr && part->nparts > 0) {
                // forward: leaves, second-level separators, top;  backward: the reverse (solved blocks of the later levels are plain multipliers)
                CholPartition top = *part, mids;
                const bool two = part->nmids > 0;
                if (two) { mids = mid_level(*part); top.ksep = part->msep; }
                if (!backward) {
                    k_trsv_cluster<<<kBandCluster * part->nparts, 256, sizeof(TrsvClSmem), st>>>(n, L, ld, dinv, b, nzt, lst, backward, *part, 0);
                    if (two) k_trsv_cluster<<<kBandCluster * mids.nparts, 256, sizeof(TrsvClSmem), st>>>(n, L, ld, dinv, b, nzt, lst, backward, mids, 0);
                }
                k_trsv_cluster<<<kBandCluster, 256, sizeof(TrsvClSmem), st>>>(n, L, ld, dinv, b, nzt, lst, backward, top, 1);
                if (backward) {
                    if (two) k_trsv_cluster<<<kBandCluster * mids.nparts, 256, sizeof(TrsvClSmem), st>>>(n, L, ld, dinv, b, nzt, lst, backward, mids, 0);
                    k_trsv_cluster<<<kBandCluster * part->nparts, 256, sizeof(TrsvClSmem), st>>>(n, L, ld, dinv, b, nzt, lst, backward, *part, 0);
                }
            } else {
                CholPartition whole;
                k_trsv_cluster<<<kBandCluster, 256, sizeof(TrsvClSmem), st>>>(n, L, ld, dinv, b, nzt, lst, backward, whole, 1);
            }
        }
        else k_trsv_single<<<1, 256, sizeof(double) * (size_t)nblk * NB + sizeof(int) * (size_t)(kTrsvSparseFill + 1) * nblk, st>>>(n, L, ld, dinv, b, nzt, ws_list(ws, n, backward),
                                                                                                                                  backward);
        if (g_prof) {
            int h[2] = {0, 0};
            cudaStreamSynchronize(st); cudaMemcpy(h, nzt, sizeof(h), cudaMemcpyDeviceToHost);
            printf("  trsv: nblk=%d non-zero tiles=%d ops=%d sparse path=%d\n", nblk, h[0], h[1], (int)trsv_use_sparse(n, nblk, h[0]));
        }
    }
    if (sparse_certain && n <= kTrsvSparseMaxN && !g_prof) return 1 + ((part != nullptr && part->nparts > 0) ? (part->nmids > 0 ? 2 : 1) : 0);   // the caller knows the tile count: no need to launch the dense kernel just to see it return
    int cacheF = nblk * nblk <= kTrsvFCacheMax ? 1 : 0;
    size_t smem = sizeof(double) * 2 * kTrsvOwn * NB * NB + (cacheF ? (size_t)((nblk * nblk + 15) & ~15) : 0);
    void* args[] = {(void*)&n, (void*)&L, (void*)&ld, (void*)&dinv, (void*)&b, (void*)&ybuf, (void*)&flags, (void*)&F, (void*)&epoch, (void*)&backward,
                    (void*)&cacheF, (void*)&nzt};
    cudaLaunchCooperativeKernel((void*)k_trsv_flags, dim3(blocks), dim3(256), args, smem, st);
    return 2 + ((part != nullptr && part->nparts > 0 && n <= kTrsvSparseMaxN) ? (part->nmids > 0 ? 2 : 1) : 0);
}
// Structure of the last factor held in `ws`: non-zero 64x64 tiles of L and the flops the factorisation actually executed
// (potrf 64^3/3 per diagonal block, 64^3 per off-diagonal tile for the right solve, 2*64^3 per pair of non-zero tiles of a
// block column for the symmetric update) -- what a roofline may count when zero tiles are skipped.  Synchronises `st`.
int dense_cholesky_stats(cudaStream_t st, int n, const double* ws, int64_t* nblk_out, int64_t* nz_tiles, double* factor_flops) {
    const int nblk = chol_nblk(n);
    const unsigned char* F = ws_F(const_cast<double*>(ws), n);
    unsigned char* h = (unsigned char*)malloc((size_t)nblk * nblk);
    if (h == nullptr) return -1;
    if (cudaMemcpyAsync(h, F, (size_t)nblk * nblk, cudaMemcpyDeviceToHost, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess) { free(h); return -1; }
    const double b3 = 64.0 * 64.0 * 64.0;
    double fl = 0.0; int64_t nz = 0;
    for (int k = 0; k < nblk; ++k) {
        int64_t ck = 0;
        for (int j = k + 1; j < nblk; ++j) ck += h[(size_t)k * nblk + j] != 0;
        nz += 1 + ck;
        fl += b3 / 3.0 + (double)ck * b3 + (double)(ck * (ck + 1) / 2) * 2.0 * b3;
    }
    free(h);
    if (nblk_out) *nblk_out = nblk;
    if (nz_tiles) *nz_tiles = nz;
    if (factor_flops) *factor_flops = fl;
    return 0;
}
int64_t dense_cholesky_forward(cudaStream_t st, int n, const double* L, int64_t ld, double* ws, double* b, const CholPartition* part, bool sparse_certain) {
    return trsv(st, n, L, ld, ws, b, 0, part, sparse_certain);
}
int64_t dense_cholesky_backward(cudaStream_t st, int n, const double* L, int64_t ld, double* ws, double* b, const CholPartition* part, bool sparse_certain) {
    return trsv(st, n, L, ld, ws, b, 1, part, sparse_certain);
}


// ---------------------------------------------------------------------------------------------------------------------
// General C -= A * B^T on DMMA (the TRSM and covariance updates of the EKF chain).  Same pipeline as k_syrk_dmma: operands are
// read in [k][row] order straight from the column-major matrices, 128x128 tiles, persistent CTAs over the tile list.
// ksplit > 1: blockIdx.y owns the K slice [y * Kper, (y + 1) * Kper) and ADDS its product atomically (a short-and-wide product -- few
// output tiles, long contraction -- would otherwise keep only a handful of SMs busy)
constexpr int kGemmKC64 = 12;     // K chunk of the 64-row variant
constexpr int kGemmKC = 16;      // K chunk per stage of the general product (32 halves the barriers but measured no faster: SYRK 5.04 -> 4.99 ms, TRSM 4.28 -> 4.35)
// TM = rows of a CTA tile: 128 (8 warps, one CTA per SM) or 64 (4 warps of the same 32 x 64 warp tile, <= 168 registers: THREE CTAs per SM whose
// barriers and epilogues are out of step with each other -- the shape cuBLAS's own DGEMM on this GPU uses, cutlass d884gemm 64x128_16x3).
template <int TM>
__global__ void __launch_bounds__(TM * 2, TM == 64 ? 3 : 1) k_gemm_nt_dmma(int m, int n, int Ktot, const double* __restrict__ A, int64_t lda, const double* __restrict__ B, int64_t ldb,
                                                                           double* __restrict__ C, int64_t ldc, int lower_only, int ksplit) {
    constexpr int TN = 128, SLDA = TM + 4, SLDB = TN + 4, NJ = 8, KC = TM == 64 ? kGemmKC64 : kGemmKC, NT = TM * 2;   // 64-row tiles: 3 x 19 KB of shared memory per CTA, three CTAs per SM
    if (ksplit > 1) {
        const int Kper = (((Ktot + ksplit - 1) / ksplit) + 47) / 48 * 48;      // a multiple of both K chunk sizes
        const int kb = (int)blockIdx.y * Kper;
        if (kb >= Ktot) return;
        A += (size_t)kb * lda; B += (size_t)kb * ldb;
        Ktot = Ktot - kb < Kper ? Ktot - kb : Kper;
    }
    const int K = Ktot;
    extern __shared__ double sm[];
    double* sA = sm;
    double* sB = sm + STAGES * KC * SLDA;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tm = (m + TM - 1) / TM, tn = (n + TN - 1) / TN;
    const int nk = (K + KC - 1) / KC;
    const int wr = (warp % (TM / 32)) * 32, wc = (warp / (TM / 32)) * 64;
    const int g = lane >> 2, tg = lane & 3;
    // Work list.  lower_only: ONLY the tiles that touch the lower triangle, numbered down the block columns -- striding over the full
    // tm x tn grid and skipping the upper ones left the CTAs with 6 to 10 tiles each (ncu: the DMMA pipe 84 % busy while an SM was
    // active, 71 % of the elapsed time).  Every tile of the list costs the same, so a plain round-robin is balanced to one tile.
    // block column c of the tile grid reaches the lower triangle from tile row (c * TN) / TM on
    auto first_row = [&](int c) { const int f = (c * TN) / TM; return f < tm ? f : tm; };
    int ntiles = tm * tn;
    if (lower_only) { ntiles = 0; for (int c = 0; c < tn; ++c) ntiles += tm - first_row(c); }
    auto tile_of = [&](int t, int& ti, int& tj) {
        if (!lower_only) { ti = t % tm; tj = t / tm; return; }
        int c = 0, rem = t;
        while (c < tn - 1 && rem >= tm - first_row(c)) { rem -= tm - first_row(c); ++c; }
        ti = first_row(c) + rem; tj = c;
    };
    auto load_chunk = [&](int i0, int j0, int stage, int kc) {
#pragma unroll
        for (int q = 0; q < KC * (TM / 2) / NT; ++q) {          // A: KC x TM doubles
            const int v = tid + NT * q;
            const int k = v / (TM / 2), rp = (v % (TM / 2)) * 2;
            const int kk = kc * KC + k;
            const int row = i0 + rp;
            int bytes = (kk < K) ? (m - row) * 8 : 0; bytes = bytes > 16 ? 16 : (bytes < 0 ? 0 : bytes);
            cp_async16(sA + (stage * KC + k) * SLDA + rp, A + (bytes > 0 ? (size_t)kk * lda + row : 0), bytes);
        }
#pragma unroll
        for (int q = 0; q < KC * (TN / 2) / NT; ++q) {          // B: KC x TN doubles
            const int v = tid + NT * q;
            const int k = v / (TN / 2), rp = (v % (TN / 2)) * 2;
            const int kk = kc * KC + k;
            const int row = j0 + rp;
            int bytes = (kk < K) ? (n - row) * 8 : 0; bytes = bytes > 16 ? 16 : (bytes < 0 ? 0 : bytes);
            cp_async16(sB + (stage * KC + k) * SLDB + rp, B + (bytes > 0 ? (size_t)kk * ldb + row : 0), bytes);
        }
    };
    // the first two chunks of a tile are requested BEFORE the epilogue of the tile before it: its read-modify-write of C then overlaps them
    bool primed = false;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
        int ti, tj;
        tile_of(t, ti, tj);
        const int i0 = ti * TM, j0 = tj * TN;
        if (!primed) {
            load_chunk(i0, j0, 0, 0); cp_async_commit();
            if (nk > 1) load_chunk(i0, j0, 1, 1);
            cp_async_commit();
        }
        double acc[4][NJ][2];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < NJ; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }
        for (int kc = 0; kc < nk; ++kc) {
            cp_async_wait<1>();
            __syncthreads();
            if (kc + 2 < nk) load_chunk(i0, j0, (kc + 2) % STAGES, kc + 2);
            cp_async_commit();
            const double* cA = sA + (kc % STAGES) * KC * SLDA;
            const double* cB = sB + (kc % STAGES) * KC * SLDB;
#pragma unroll
            for (int ks = 0; ks < KC; ks += 4) {
                double af[4], bf[NJ];
                const double* pa = cA + (ks + tg) * SLDA + wr + g;
                const double* pb = cB + (ks + tg) * SLDB + wc + g;
#pragma unroll
                for (int i = 0; i < 4; ++i) af[i] = pa[i * 8];
#pragma unroll
                for (int j = 0; j < NJ; ++j) bf[j] = pb[j * 8];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < NJ; ++j) dmma_m8n8k4(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
            }
        }
        cp_async_wait<0>();
        __syncthreads();
        primed = false;
        if (t + (int)gridDim.x < ntiles) {      // every stage is free now: request the next tile's first two chunks, then do the epilogue
            int ni, nj;
            tile_of(t + (int)gridDim.x, ni, nj);
            load_chunk(ni * TM, nj * TN, 0, 0); cp_async_commit();
            if (nk > 1) load_chunk(ni * TM, nj * TN, 1, 1);
            cp_async_commit();
            primed = true;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                const int row = i0 + wr + i * 8 + g;
                const int col = j0 + wc + j * 8 + tg * 2;
                const bool ok0 = row < m && col < n && (!lower_only || row >= col), ok1 = row < m && col + 1 < n && (!lower_only || row >= col + 1);
                if (ksplit > 1) {
                    if (ok0) atomicAdd(&C[(size_t)col * ldc + row], -acc[i][j][0]);
                    if (ok1) atomicAdd(&C[(size_t)(col + 1) * ldc + row], -acc[i][j][1]);
                    continue;
                }
                acc[i][j][0] = (ok0 ? C[(size_t)col * ldc + row] : 0.0) - acc[i][j][0];     // every load of the read-modify-write before the first store
                acc[i][j][1] = (ok1 ? C[(size_t)(col + 1) * ldc + row] : 0.0) - acc[i][j][1];
            }
        if (ksplit > 1) continue;
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                const int row = i0 + wr + i * 8 + g;
                const int col = j0 + wc + j * 8 + tg * 2;
                if (row < m) {
                    if (col < n && (!lower_only || row >= col)) C[(size_t)col * ldc + row] = acc[i][j][0];
                    if (col + 1 < n && (!lower_only || row >= col + 1)) C[(size_t)(col + 1) * ldc + row] = acc[i][j][1];
                }
            }
    }
}
// the operands need 16-byte aligned columns: lda, ldb even and base pointers 16-byte aligned (the callers guarantee it)
void launch_gemm_nt_dmma(cudaStream_t st, int m, int n, int K, const double* A, int64_t lda, const double* B, int64_t ldb, double* C, int64_t ldc, int lower_only,
                         int allow_split_k, int cta_per_tile) {
    set_attrs_once();
    static PerDeviceOnce once;
    const size_t smem128 = sizeof(double) * (STAGES * kGemmKC * ((128 + 4) + (128 + 4)));
    const size_t smem64 = sizeof(double) * (STAGES * kGemmKC64 * ((64 + 4) + (128 + 4)));
    if (once.first()) {
        cudaFuncSetAttribute(k_gemm_nt_dmma<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem128);
        cudaFuncSetAttribute(k_gemm_nt_dmma<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem64);
    }
    if (m <= 0 || n <= 0 || K <= 0) return;
    static int tile_env = -1;        // SRK_GEMM_TILE=64: 64 x 128 CTA tiles, three CTAs per SM (measured: SYRK 5.20 ms against 5.09 ms with 128 x 128)
    if (tile_env < 0) { const char* e = getenv("SRK_GEMM_TILE"); tile_env = (e != nullptr && atoi(e) == 64) ? 64 : 128; }
    const int TM = tile_env;
    const int tm = (m + TM - 1) / TM, tn = (n + 127) / 128;
    int tiles = tm * tn;                                                  // tiles that do work
    if (lower_only) { tiles = 0; for (int c = 0; c < tn; ++c) { int f = (c * 128) / TM; if (f > tm) f = tm; tiles += tm - f; } }
    int ksplit = 1;
    if (allow_split_k && tiles < g_sms / 2 && K >= 4096) { ksplit = (2 * g_sms + tiles - 1) / tiles; const int maxs = K / 1024; if (ksplit > maxs) ksplit = maxs; if (ksplit < 1) ksplit = 1; }
    // cta_per_tile: one CTA per tile instead of persistent CTAs -- an SM is handed back after every tile, so the kernels of a
    // higher-priority stream (the factorisation chain that runs beside the EKF gain) get in at tile granularity
    const int resident = TM == 64 ? 3 * g_sms : g_sms;
    const int gx = (cta_per_tile || tiles < resident) ? tiles : resident;
    if (TM == 64) k_gemm_nt_dmma<64><<<dim3(gx, ksplit), 128, smem64, st>>>(m, n, K, A, lda, B, ldb, C, ldc, lower_only, ksplit);
    else k_gemm_nt_dmma<128><<<dim3(gx, ksplit), 256, smem128, st>>>(m, n, K, A, lda, B, ldb, C, ldc, lower_only, ksplit);
}

// X <- X * Linv^T on an (rows x 64) block column: X(r,c) = sum_{q<=c} X(r,q) * Linv(c,q).
__global__ void __launch_bounds__(256) k_block_right_solve(int rows, double* __restrict__ A, int64_t lda, const double* __restrict__ dinv) {
    extern __shared__ double sm[];
    double* sLi = sm;                    // sLi[q*NB + c] = Linv(c, q)
    double* sA = sm + NB * NB;           // sA[q*PS_ROWS + r]
    const int tid = threadIdx.x;
    const int r0 = blockIdx.x * PS_ROWS;
    for (int e = tid; e < NB * NB; e += 256) sLi[e] = dinv[e];   // dinv[q*NB + c] = Linv(c, q) (column q)
    for (int e = tid; e < NB * PS_ROWS; e += 256) {
        const int q = e / PS_ROWS, r = e % PS_ROWS;
        sA[e] = (r0 + r < rows) ? A[(size_t)q * lda + r0 + r] : 0.0;
    }
    __syncthreads();
    const int r = tid & (PS_ROWS - 1), cbase = (tid >> 7) * 32;
    double acc[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) acc[i] = 0.0;
    for (int q = 0; q < cbase + 32; ++q) {
        const double a = sA[q * PS_ROWS + r];
        const double* li = sLi + q * NB + cbase;
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[i] += a * li[i];
    }
    if (r0 + r < rows) {
#pragma unroll
        for (int i = 0; i < 32; ++i) A[(size_t)(cbase + i) * lda + r0 + r] = acc[i];
    }
}
void launch_block_right_solve(cudaStream_t st, int rows, double* A, int64_t lda, const double* dinv_block, int ncols) {
    set_attrs_once();
    if (rows > 0 && ncols > 0) k_right_solve_dmma<<<(rows + RS_ROWS - 1) / RS_ROWS, 256, kRightSolveSmem, st>>>(rows, 0, A, lda, dinv_block, nullptr, 0, ncols < NB ? ncols : NB);
}

void launch_mirror_lower(cudaStream_t st, int n, double* A, int64_t ld) {
    int T = (n + 31) / 32;
    k_mirror_lower<<<dim3(T, T), 256, 0, st>>>(n, A, ld);
}
void launch_residual_dd(cudaStream_t st, int n, const double* A, int64_t ld, const double* x, const double* b, double* r) {
    k_residual_dd<<<(n + 7) / 8, 256, 0, st>>>(n, A, ld, x, b, r);
}
void launch_axpy1(cudaStream_t st, int n, const double* d, double* x) { k_axpy1<<<(n + 255) / 256, 256, 0, st>>>(n, d, x); }

}  // namespace srk
