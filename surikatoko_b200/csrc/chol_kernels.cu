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
    if (tid < NB) y[tid] = tid < nb ? b[k0 + tid] : 0.0;
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
    if (tid < nb) b[k0 + tid] = y[tid];
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
    if (tid < NB) y[tid] = b[k0 + tid];
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
        b[r0 + tid] -= acc;
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
__global__ void __launch_bounds__(128) k_back64(int n, int kb, int first, double* __restrict__ A, int64_t ld, double* __restrict__ b) {
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

__global__ void k_symv_lower(int n, const double* __restrict__ A, int64_t ld, const double* __restrict__ x, double* __restrict__ y) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double s = 0.0;
    for (int j = 0; j <= i; ++j) s += A[(size_t)j * ld + i] * x[j];
    for (int j = i + 1; j < n; ++j) s += A[(size_t)i * ld + j] * x[j];
    y[i] = s;
}
__global__ void k_mirror_lower(int n, double* __restrict__ A, int64_t ld) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (int64_t)n * n) return;
    int c = (int)(e / n), r = (int)(e % n);
    if (r > c) A[(size_t)r * ld + c] = A[(size_t)c * ld + r];
}

size_t dense_cholesky_work_doubles(int) { return 0; }

int64_t dense_cholesky_solve(cudaStream_t st, int n, double* A, int64_t ld, double* b, int* info_dev, double*) {
    static bool attr_set = false;
    const size_t trsm_smem = sizeof(double) * (NB * NB + NB * TR + NB);
    const size_t syrk_smem = sizeof(double) * (2 * NB * SLD);
    if (!attr_set) {
        cudaFuncSetAttribute(k_trsm64, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)trsm_smem);
        cudaFuncSetAttribute(k_syrk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)syrk_smem);
        attr_set = true;
    }
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
    // backward: L^T x = y
    k_back64<<<1, 128, 0, st>>>(n, nblk - 1, 1, A, ld, b); ++launches;
    for (int kb = nblk - 1; kb >= 1; --kb) { k_back64<<<kb, 128, 0, st>>>(n, kb, 0, A, ld, b); ++launches; }
    return launches;
}

void launch_symv_lower(cudaStream_t st, int n, const double* A, int64_t ld, const double* x, double* y) {
    k_symv_lower<<<(n + 127) / 128, 128, 0, st>>>(n, A, ld, x, y);
}
void launch_mirror_lower(cudaStream_t st, int n, double* A, int64_t ld) {
    int64_t tot = (int64_t)n * n;
    k_mirror_lower<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(n, A, ld);
}

}  // namespace srk
