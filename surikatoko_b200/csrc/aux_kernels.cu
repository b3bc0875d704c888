// suriko-b200 — small glue kernels of the LM control path.
#include "kernels.h"

namespace srk {

// One attempt's scalars for the host: slots[0..world) = per-rank error partial (only this rank's slot is non-zero, so a
// sum all-reduce moves the partials exactly and every rank adds them in rank order), slots[world] = non-finite flag,
// slots[world+1] = number of skipped points (quirk Q5).
__global__ void k_pack_attempt(const double* __restrict__ err_sum, const int* __restrict__ finite_flag,
                               const unsigned long long* __restrict__ skipped_cnt, int rank, int world, double* __restrict__ slots) {
    int t = threadIdx.x;
    if (t < world) slots[t] = (t == rank) ? err_sum[0] : 0.0;
    if (t == 0) {
        slots[world] = finite_flag != nullptr ? (double)finite_flag[0] : 0.0;
        slots[world + 1] = skipped_cnt != nullptr ? (double)skipped_cnt[0] : 0.0;
        slots[world + 2] = finite_flag != nullptr ? (double)(finite_flag[1] != 0) : 0.0;   // flags[2]: Cholesky info (first non-positive pivot)
    }
}

// X_try = X + corrections (AoS [3N], BA.cpp:2003-2017) for the parity hook srk_ba_debug_apply.
__global__ void k_add_points_aos(int64_t N, const double* __restrict__ X, const double* __restrict__ corr, double* __restrict__ Xtry) {
    int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= N) return;
    Xtry[j] = X[j] + corr[3 * j];
    Xtry[N + j] = X[N + j] + corr[3 * j + 1];
    Xtry[2 * N + j] = X[2 * N + j] + corr[3 * j + 2];
}

// ---- multi-GPU: all-reduce only the non-zero 64x64 tiles of the reduced camera system -------------------------------------------
// Every rank's S covers the camera pairs of its own points; the union over ranks is still block-banded for a scene with localized
// visibility (~4 % of the tiles at configs[2]: 16 MB instead of 0.8 GB on the wire).  mask (doubles, so that the f64 sum all-reduce
// callback can carry it) -> union over ranks -> tile list -> gather / all-reduce / scatter.
constexpr int kTile = 64;
__global__ void __launch_bounds__(256) k_tile_mask(int n, const double* __restrict__ S, int64_t ld, int nblk, double* __restrict__ mask) {
    const int c = blockIdx.x, r = blockIdx.y;
    if (r < c) return;
    const int tid = threadIdx.x;
    int nz = (r == c) ? 1 : 0;          // diagonal tiles always travel
    if (!nz) {
#pragma unroll 4
        for (int it = 0; it < 16; ++it) {
            const int e = tid + 256 * it;
            const int q = e >> 6, rr = e & 63;
            const int row = r * kTile + rr, col = c * kTile + q;
            if (row < n && col < n) nz |= S[(size_t)col * ld + row] != 0.0;
        }
    }
    nz = __syncthreads_or(nz);
    if (tid == 0) mask[(size_t)c * nblk + r] = nz ? 1.0 : 0.0;
}
// list[i] = c * nblk + r of the i-th tile of the union pattern (column-major tile order), count[0] = number of tiles; one CTA
__global__ void __launch_bounds__(1024) k_tile_list(int nblk, const double* __restrict__ mask, int* __restrict__ list, int* __restrict__ count) {
    __shared__ int warp_sums[32];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int total = nblk * nblk;
    for (int base = 0; base < total; base += 1024) {
        const int e = base + threadIdx.x;
        const int c = e / nblk, r = e - c * nblk;
        const int occ = (e < total && r >= c && mask[e] > 0.0) ? 1 : 0;
        const unsigned bal = __ballot_sync(0xffffffffu, occ);
        if (lane == 0) warp_sums[w] = __popc(bal);
        __syncthreads();
        int off = carry;
        for (int i = 0; i < w; ++i) off += warp_sums[i];
        if (occ) list[off + __popc(bal & ((1u << lane) - 1))] = e;
        __syncthreads();
        if (threadIdx.x == 0) { int t = 0; for (int i = 0; i < 32; ++i) t += warp_sums[i]; carry += t; }
        __syncthreads();
    }
    if (threadIdx.x == 0) count[0] = carry;
}
// dir 0: packed[i] <- tile list[i] of S (zeros outside the matrix), packed[count*4096 ..] <- rhs;  dir 1: the reverse
__global__ void __launch_bounds__(256) k_tile_pack(int n, double* __restrict__ S, int64_t ld, int nblk, const int* __restrict__ list, int count,
                                                   double* __restrict__ rhs, int64_t nrhs, double* __restrict__ packed, int dir) {
    const int i = blockIdx.x, tid = threadIdx.x;
    if (i >= count) {   // the last CTAs carry the right-hand side
        double* pr = packed + (size_t)count * kTile * kTile;
        for (int64_t e = (int64_t)(i - count) * 256 + tid; e < nrhs; e += (int64_t)(gridDim.x - count) * 256) { if (dir == 0) pr[e] = rhs[e]; else rhs[e] = pr[e]; }
        return;
    }
    const int t = list[i], c = t / nblk, r = t - c * nblk;
    double* pt = packed + (size_t)i * kTile * kTile;
#pragma unroll 4
    for (int it = 0; it < 16; ++it) {
        const int e = tid + 256 * it;
        const int q = e >> 6, rr = e & 63;
        const int row = r * kTile + rr, col = c * kTile + q;
        const bool in = row < n && col < n;
        if (dir == 0) pt[e] = in ? S[(size_t)col * ld + row] : 0.0;
        else if (in) S[(size_t)col * ld + row] = pt[e];
    }
}
void launch_tile_mask(cudaStream_t st, int n, const double* S, int64_t ld, double* mask) {
    const int nblk = (n + kTile - 1) / kTile;
    k_tile_mask<<<dim3(nblk, nblk), 256, 0, st>>>(n, S, ld, nblk, mask);
}
void launch_tile_list(cudaStream_t st, int n, const double* mask, int* list, int* count) {
    k_tile_list<<<1, 1024, 0, st>>>((n + kTile - 1) / kTile, mask, list, count);
}
void launch_tile_pack(cudaStream_t st, int n, double* S, int64_t ld, const int* list, int count, double* rhs, int64_t nrhs, double* packed, int dir) {
    k_tile_pack<<<count + 8, 256, 0, st>>>(n, S, ld, (n + kTile - 1) / kTile, list, count, rhs, nrhs, packed, dir);
}

void launch_pack_attempt(cudaStream_t st, const double* err_sum, const int* finite_flag, const unsigned long long* skipped_cnt, int rank, int world,
                         double* slots) {
    k_pack_attempt<<<1, world < 32 ? 32 : ((world + 31) / 32) * 32, 0, st>>>(err_sum, finite_flag, skipped_cnt, rank, world, slots);
}
void launch_add_points_aos(cudaStream_t st, int64_t N, const double* X, const double* corr_aos, double* Xtry) {
    if (N > 0) k_add_points_aos<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(N, X, corr_aos, Xtry);
}

}  // namespace srk
