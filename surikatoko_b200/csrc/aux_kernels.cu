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

void launch_pack_attempt(cudaStream_t st, const double* err_sum, const int* finite_flag, const unsigned long long* skipped_cnt, int rank, int world,
                         double* slots) {
    k_pack_attempt<<<1, world < 32 ? 32 : ((world + 31) / 32) * 32, 0, st>>>(err_sum, finite_flag, skipped_cnt, rank, world, slots);
}
void launch_add_points_aos(cudaStream_t st, int64_t N, const double* X, const double* corr_aos, double* Xtry) {
    if (N > 0) k_add_points_aos<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(N, X, corr_aos, Xtry);
}

}  // namespace srk
