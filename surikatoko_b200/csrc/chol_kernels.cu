// suriko-b200 — K3a: dense FP64 Cholesky solve of the reduced camera system S*df = rhs on sm_100a.
//
// Replaces Eigen's householderQr().solve at BA.cpp:1911.  S is symmetric positive definite (a Schur complement of the
// damped Gauss-Newton Hessian), so LL^T is the appropriate factorisation.  FP64 tensor work on Blackwell is DMMA:
// mma.sync.aligned.m8n8k4.f64 (tcgen05.mma has no f64 kind, TMEM accumulators are not available for doubles).
//
// Right-looking blocked algorithm, panel width NB = 64, column-major lower triangle, leading dimension ld (multiple of 2):
//   k_potrf64 : diagonal 64x64 tile in shared memory (one CTA) + forward substitution of the matching rhs slice
//   k_trsm64  : panel rows below the tile,  P <- P * L_kk^-T  in shared memory, fused rhs update  b_rows -= P*y_k
//   k_syrk    : trailing update  C -= P*P^T  on 128x128 lower tiles, K = 64, DMMA with register tiles of 32x64 per warp
//   k_back64  : backward substitution L^T x = y, one launch per block column (right-looking on row panels)
#include <math.h>
#include "kernels.h"

namespace srk {

constexpr int NB = 64;

// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_potrf64(int n, int k0, double* __restrict__ A, int64_t ld, double* __restrict__ b, int* __restrict__ info) {
    __shared__ double T[NB][NB + 1];  // T[c][r]
    __shared__ double y[NB];
    const int nb = min(NB, n - k0);
    const int tid = threadIdx.x;
    for (int e = tid; e < NB * NB; e += 256) {
        int c = e / NB, r = e % NB;
        T[c][r] = (r < nb && c < nb && r >= c) ? A[(size_t)(k0 + c) * ld + k0 + r] : 0.0;
    }
    if (tid < NB) y[tid] = (b != nullptr && tid < nb) ? b[k0 + tid] : 0.0;
    __syncthreads();
    for (int j = 0; j < nb; ++j) {
        double d = T[j][j];
        if (tid == 0 && !(d > 0.0) && atomicCAS(info, 0, k0 + j + 1) == 0) {}
        double sd = sqrt(d);
        __syncthreads();
        // scale column j
        if (tid >= j && tid < nb) T[j][tid] = (tid == j) ? sd : T[j][tid] / sd;
        __syncthreads();
        // rank-1 update of the trailing lower part: T[c][r] -= T[j][r]*T[j][c], j < c <= r
        int m = nb - j - 1;
        for (int e = tid; e < m * m; e += 256) {
            int c = j + 1 + e / m, r = j + 1 + e % m;
            if (r >= c) T[c][r] -= T[j][r] * T[j][c];
        }
        // no barrier needed here: the next iteration's first barrier orders these writes before they are read
        __syncthreads();
    }
    // forward substitution  L y = b_k  (column oriented)
    for (int j = 0; j < nb; ++j) {
        if (tid == 0) y[j] = y[j] / T[j][j];
        __syncthreads();
        if (tid > j && tid < nb) y[tid] -= T[j][tid] * y[j];
        __syncthreads();
    }
    for (int e = tid; e < NB * NB; e += 256) {
        int c = e / NB, r = e % NB;
        if (r < nb && c < nb && r >= c) A[(size_t)(k0 + c) * ld + k0 + r] = T[c][r];
    }
    if (b != nullptr && tid < nb) b[k0 + tid] = y[tid];
}

// ---------------------------------------------------------------------------------------------------------------------
// Rows [k0+NB, n) of the panel: 128 rows per CTA, one thread per row.  smem: L_kk (64x64) and the row tile (64 x 128).
constexpr int TR = 128;
__global__ void __launch_bounds__(TR) k_trsm64(int n, int k0, double* __restrict__ A, int64_t ld, double* __restrict__ b) {
    extern __shared__ double sm[];
    double* L = sm;                 // L[c*NB + r]  (column-major tile)
    double* Xt = sm + NB * NB;      // Xt[c*TR + row]
    double* y = Xt + NB * TR;       // y[NB]
    const int tid = threadIdx.x;
    const int r0 = k0 + NB + blockIdx.x * TR;
    const int rows = min(TR, n - r0);
    for (int e = tid; e < NB * NB; e += TR) {
        int c = e / NB, r = e % NB;
        L[e] = (r >= c) ? A[(size_t)(k0 + c) * ld + k0 + r] : 0.0;
    }
    if (tid < NB) y[tid] = b != nullptr ? b[k0 + tid] : 0.0;
    for (int c = 0; c < NB; ++c) Xt[c * TR + tid] = tid < rows ? A[(size_t)(k0 + c) * ld + r0 + tid] : 0.0;
    __syncthreads();
    double acc = 0.0;
    for (int c = 0; c < NB; ++c) {
        double s = Xt[c * TR + tid];
#pragma unroll 8
        for (int m = 0; m < c; ++m) s -= Xt[m * TR + tid] * L[m * NB + c];   // L(c,m)
        s = s / L[c * NB + c];
        Xt[c * TR + tid] = s;
        acc += s * y[c];
    }
    if (tid < rows) {
        for (int c = 0; c < NB; ++c) A[(size_t)(k0 + c) * ld + r0 + tid] = Xt[c * TR + tid];
        if (b != nullptr) b[r0 + tid] -= acc;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Trailing update C -= P*P^T on lower 128x128 tiles; P = A[k0+NB.., k0..k0+63] (already solved by k_trsm64).
constexpr int TM = 128;
constexpr int SLD = TM + 4;  // padded row stride of the [k][row] shared tiles: conflict-free DMMA fragment loads

__device__ __forceinline__ void dmma_m8n8k4(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(256, 1) k_syrk(int n, int k0, double* __restrict__ A, int64_t ld) {
    extern __shared__ double sm[];
    double* sA = sm;               // [NB][SLD]  rows of the tile-row block
    double* sB = sm + NB * SLD;    // [NB][SLD]  rows of the tile-column block
    // decode lower-triangular tile index
    const int t0 = k0 + NB;
    int bt = blockIdx.x;
    int ti = (int)((sqrtf(8.0f * (float)bt + 1.0f) - 1.0f) * 0.5f);
    while (ti * (ti + 1) / 2 > bt) --ti;
    while ((ti + 1) * (ti + 2) / 2 <= bt) ++ti;
    const int tj = bt - ti * (ti + 1) / 2;
    const int i0 = t0 + ti * TM, j0 = t0 + tj * TM;
    const int tid = threadIdx.x;
    // load the two 128x64 panel tiles (coalesced along rows)
    for (int e = tid; e < NB * TM; e += 256) {
        int k = e / TM, r = e % TM;
        int gi = i0 + r, gj = j0 + r;
        sA[k * SLD + r] = gi < n ? A[(size_t)(k0 + k) * ld + gi] : 0.0;
        sB[k * SLD + r] = gj < n ? A[(size_t)(k0 + k) * ld + gj] : 0.0;
    }
    __syncthreads();
    const int warp = tid >> 5, lane = tid & 31;
    const int wr = (warp & 3) * 32;   // warp row offset in the tile (4 warps along rows)
    const int wc = (warp >> 2) * 64;  // warp col offset (2 warps along cols)
    const int g = lane >> 2, tg = lane & 3;
    double acc[4][8][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }
#pragma unroll 2
    for (int ks = 0; ks < NB; ks += 4) {
        double a[4], bfr[8];
        const double* pa = sA + (ks + tg) * SLD + wr + g;
        const double* pb = sB + (ks + tg) * SLD + wc + g;
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = pa[i * 8];
#pragma unroll
        for (int j = 0; j < 8; ++j) bfr[j] = pb[j * 8];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) dmma_m8n8k4(acc[i][j][0], acc[i][j][1], a[i], bfr[j]);
    }
    // epilogue: C(row, col) -= acc, lower part only
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            int row = i0 + wr + i * 8 + g;
            int col = j0 + wc + j * 8 + tg * 2;
            if (row < n) {
                if (col < n && row >= col) A[(size_t)col * ld + row] -= acc[i][j][0];
                if (col + 1 < n && row >= col + 1) A[(size_t)(col + 1) * ld + row] -= acc[i][j][1];
            }
        }
}

// ---------------------------------------------------------------------------------------------------------------------
// Backward substitution, block column kb (descending).  x lives in b.  grid.x = kb+1 CTAs:
//   CTA j < kb : b_j -= L(kb-block rows, j-block cols)^T * x_kb         (x_kb final from the previous launch)
//   ... except that x_kb itself must first be produced: launch order is  solve_diag(kb) ; update(j<kb) — two tiny
//   kernels would double the launch count, so CTA j == kb-1 of launch kb also solves the diagonal block kb-1 after its
//   own update, and launch kb == last block starts with a diagonal-only launch.
__global__ void __launch_bounds__(128) k_back64(int n, int kb, int first, const double* __restrict__ A, int64_t ld, double* __restrict__ b) {
    __shared__ double xk[NB];
    __shared__ double yj[NB];
    __shared__ double T[NB][NB + 1];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int j = blockIdx.x;  // target block column
    const int k0 = kb * NB;
    if (first) {
        // diagonal solve of the last block only
        const int nb = min(NB, n - k0);
        for (int e = tid; e < NB * NB; e += 128) { int c = e / NB, r = e % NB; T[c][r] = (r < nb && c < nb && r >= c) ? A[(size_t)(k0 + c) * ld + k0 + r] : 0.0; }
        if (tid < NB) yj[tid] = tid < nb ? b[k0 + tid] : 0.0;
        __syncthreads();
        for (int c = nb - 1; c >= 0; --c) {   // L^T x = y : x_c = (y_c - sum_{r>c} L(r,c) x_r) / L(c,c)
            if (tid == 0) yj[c] = yj[c] / T[c][c];
            __syncthreads();
            if (tid < c) yj[tid] -= T[tid][c] * yj[c];
            __syncthreads();
        }
        if (tid < nb) b[k0 + tid] = yj[tid];
        return;
    }
    const int nbk = min(NB, n - k0);
    const int j0 = j * NB;
    if (tid < NB) { xk[tid] = tid < nbk ? b[k0 + tid] : 0.0; yj[tid] = b[j0 + tid]; }
    __syncthreads();
    // y_j[c] -= sum_r L(k0+r, j0+c) * x_k[r] : each warp handles 16 columns, lanes over rows
    for (int cc = 0; cc < 16; ++cc) {
        int c = warp * 16 + cc;
        const double* col = A + (size_t)(j0 + c) * ld + k0;
        double s = 0.0;
        for (int r = lane; r < nbk; r += 32) s += col[r] * xk[r];
#pragma unroll
        for (int sft = 16; sft > 0; sft >>= 1) s += __shfl_xor_sync(0xffffffffu, s, sft);
        if (lane == 0) yj[c] -= s;
    }
    __syncthreads();
    if (j == kb - 1) {
        for (int e = tid; e < NB * NB; e += 128) { int c = e / NB, r = e % NB; T[c][r] = (r >= c) ? A[(size_t)(j0 + c) * ld + j0 + r] : 0.0; }
        __syncthreads();
        for (int c = NB - 1; c >= 0; --c) {
            if (tid == 0) yj[c] = yj[c] / T[c][c];
            __syncthreads();
            if (tid < c) yj[tid] -= T[tid][c] * yj[c];
            __syncthreads();
        }
    }
    if (tid < NB) b[j0 + tid] = yj[tid];
}

// ---------------------------------------------------------------------------------------------------------------------
// Forward substitution L y = b with an existing factor (used by the refinement steps), block column kb ascending:
//   first launch: diagonal solve of block 0.  Launch kb: CTA j (j = kb+1 .. nblk-1) applies  b_j -= L(j,kb) * y_kb ;
//   the CTA of block kb+1 then owns a complete right-hand side and solves its diagonal block.
__global__ void __launch_bounds__(128) k_fwd64(int n, int kb, int first, const double* __restrict__ A, int64_t ld, double* __restrict__ b) {
    __shared__ double yk[NB];
    __shared__ double yj[NB];
    __shared__ double part[2][NB];
    __shared__ double T[NB][NB + 1];
    const int tid = threadIdx.x;
    auto diag_solve = [&](int j0) {
        const int nb = min(NB, n - j0);
        for (int e = tid; e < NB * NB; e += 128) { int c = e / NB, r = e % NB; T[c][r] = (r < nb && c < nb && r >= c) ? A[(size_t)(j0 + c) * ld + j0 + r] : 0.0; }
        __syncthreads();
        for (int c = 0; c < nb; ++c) {
            if (tid == 0) yj[c] = yj[c] / T[c][c];
            __syncthreads();
            if (tid > c && tid < nb) yj[tid] -= T[c][tid] * yj[c];
            __syncthreads();
        }
        if (tid < nb) b[j0 + tid] = yj[tid];
    };
    if (first) {
        if (tid < NB) yj[tid] = tid < n ? b[tid] : 0.0;
        __syncthreads();
        diag_solve(0);
        return;
    }
    const int k0 = kb * NB;
    const int j = kb + 1 + blockIdx.x;
    const int j0 = j * NB;
    const int nbj = min(NB, n - j0);
    if (tid < NB) { yk[tid] = b[k0 + tid]; yj[tid] = tid < nbj ? b[j0 + tid] : 0.0; }
    __syncthreads();
    {
        const int r = tid & (NB - 1), half = tid >> 6;
        double s = 0.0;
        if (r < nbj)
            for (int c = half; c < NB; c += 2) s += A[(size_t)(k0 + c) * ld + j0 + r] * yk[c];
        part[half][r] = s;
    }
    __syncthreads();
    if (tid < NB) yj[tid] -= part[0][tid] + part[1][tid];
    __syncthreads();
    if (j == kb + 1) diag_solve(j0);
    else if (tid < nbj) b[j0 + tid] = yj[tid];
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

static void set_attrs_once() {
    static bool attr_set = false;
    if (attr_set) return;
    cudaFuncSetAttribute(k_trsm64, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(double) * (NB * NB + NB * TR + NB)));
    cudaFuncSetAttribute(k_syrk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(double) * (2 * NB * SLD)));
    attr_set = true;
}

int64_t dense_cholesky_factor(cudaStream_t st, int n, double* A, int64_t ld, double* b, int* info_dev) {
    set_attrs_once();
    const size_t trsm_smem = sizeof(double) * (NB * NB + NB * TR + NB);
    const size_t syrk_smem = sizeof(double) * (2 * NB * SLD);
    int64_t launches = 0;
    cudaMemsetAsync(info_dev, 0, sizeof(int), st);
    const int nblk = (n + NB - 1) / NB;
    for (int kb = 0; kb < nblk; ++kb) {
        const int k0 = kb * NB;
        k_potrf64<<<1, 256, 0, st>>>(n, k0, A, ld, b, info_dev); ++launches;
        const int below = n - (k0 + NB);
        if (below > 0) {
            k_trsm64<<<(below + TR - 1) / TR, TR, trsm_smem, st>>>(n, k0, A, ld, b); ++launches;
            const int T = (below + TM - 1) / TM;
            k_syrk<<<T * (T + 1) / 2, 256, syrk_smem, st>>>(n, k0, A, ld); ++launches;
        }
    }
    return launches;
}
int64_t dense_cholesky_forward(cudaStream_t st, int n, const double* L, int64_t ld, double* b) {
    const int nblk = (n + NB - 1) / NB;
    int64_t launches = 0;
    k_fwd64<<<1, 128, 0, st>>>(n, 0, 1, L, ld, b); ++launches;
    for (int kb = 0; kb + 1 < nblk; ++kb) { k_fwd64<<<nblk - kb - 1, 128, 0, st>>>(n, kb, 0, L, ld, b); ++launches; }
    return launches;
}
int64_t dense_cholesky_backward(cudaStream_t st, int n, const double* L, int64_t ld, double* b) {
    const int nblk = (n + NB - 1) / NB;
    int64_t launches = 0;
    k_back64<<<1, 128, 0, st>>>(n, nblk - 1, 1, L, ld, b); ++launches;
    for (int kb = nblk - 1; kb >= 1; --kb) { k_back64<<<kb, 128, 0, st>>>(n, kb, 0, L, ld, b); ++launches; }
    return launches;
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
