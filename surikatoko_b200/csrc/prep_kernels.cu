// suriko-b200 — bind-time index construction on the device (runs once per srk_ba_bind, not per LM iteration).
//   k_prep_obs       validate the (pnt_ind, frame_ind) ordering, split pixel pairs into planes of x/f0, y/f0, build the point CSR
//                    (pt_begin) by boundary detection, histogram observations per camera
//   k_scan_counts    exclusive scan of the per-camera histogram (M is small: one CTA)
//   k_scatter_by_cam camera-major copy of (point index, x, y) used by the frame pass (k_frame_blocks)
//   k_finite_flag / k_count_skipped  tiny reductions the LM control needs (allFinite, BA.cpp:1912, :1953)
#include "kernels.h"
#include "prep.h"

namespace srk {

// The structure half (indices only) and the value half (pixels) are separate kernels: the indices arrive first over PCIe and the
// whole structure pass of a bind (CSR, camera histogram / scatter positions, chunk tables, Schur plan, solve order) runs on a side
// stream while the 16 bytes per observation of pixel data are still in flight.
__global__ void k_prep_index(int64_t O, int64_t N, int M, const int32_t* __restrict__ obs_cam, const int32_t* __restrict__ obs_pt, int64_t* __restrict__ pt_begin,
                             unsigned long long* __restrict__ cam_count, int* __restrict__ err_flag) {
    int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int c = -1;
    if (o < O) {
        int p = obs_pt[o];
        c = obs_cam[o];
        if (p < 0 || p >= N || c < 0 || c >= M) { atomicOr(err_flag, 1); c = -1; }
        else {
            int pp = o > 0 ? obs_pt[o - 1] : -1, pc = o > 0 ? obs_cam[o - 1] : -1;
            if (p < pp || (p == pp && c <= pc)) atomicOr(err_flag, 2);
            if (p != pp) for (int q = (pp < 0 ? 0 : pp + 1); q <= p; ++q) pt_begin[q] = o;
            if (o == O - 1) for (int64_t q = p + 1; q <= N; ++q) pt_begin[q] = O;
        }
    }
    const unsigned peers = __match_any_sync(0xffffffffu, c);
    if (c >= 0 && (threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&cam_count[c], (unsigned long long)__popc(peers));
}
// camera-major position of every observation (obs_pos) and the camera-major point ids
__global__ void k_scatter_index(int64_t O, int64_t N, int M, const int32_t* __restrict__ obs_cam, const int32_t* __restrict__ obs_pt, unsigned long long* __restrict__ cursor,
                                int32_t* __restrict__ c_pt, unsigned* __restrict__ obs_pos) {
    int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int c = o < O ? obs_cam[o] : -1;
    const int p = o < O ? obs_pt[o] : -1;
    if (c < 0 || c >= M || p < 0 || p >= N) c = -1;                  // invalid input is reported by k_prep_index; do not index with it
    const unsigned peers = __match_any_sync(0xffffffffu, c);
    const int lane = threadIdx.x & 31, leader = __ffs(peers) - 1;
    unsigned long long base = 0;
    if (c >= 0 && lane == leader) base = atomicAdd(&cursor[c], (unsigned long long)__popc(peers));   // one atomic per distinct camera of the warp
    base = __shfl_sync(0xffffffffu, base, leader);
    if (c < 0) { if (o < O) obs_pos[o] = 0xffffffffu; return; }
    const unsigned long long pos = base + __popc(peers & ((1u << lane) - 1));
    c_pt[pos] = p; obs_pos[o] = (unsigned)pos;
}
// x/f0, y/f0 of BA.cpp:475-479, formed once, in point-major planes and in the camera-major copy
__global__ void k_prep_xy(int64_t O, const double* __restrict__ obs_xy, double f0, const unsigned* __restrict__ obs_pos, double* __restrict__ x, double* __restrict__ y,
                          double* __restrict__ c_x, double* __restrict__ c_y) {
    int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= O) return;
    const double2 xy = reinterpret_cast<const double2*>(obs_xy)[o];
    const double xs = xy.x / f0, ys = xy.y / f0;
    x[o] = xs; y[o] = ys;
    const unsigned pos = obs_pos[o];
    if (pos != 0xffffffffu) { c_x[pos] = xs; c_y[pos] = ys; }
}
__global__ void k_prep_obs(int64_t O, int64_t N, int M, const int32_t* __restrict__ obs_cam, const int32_t* __restrict__ obs_pt,
                           const double* __restrict__ obs_xy, double f0, double* __restrict__ x, double* __restrict__ y, int64_t* __restrict__ pt_begin,
                           unsigned long long* __restrict__ cam_count, int* __restrict__ err_flag) {
    int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int c = -1;
    if (o < O) {
        int p = obs_pt[o];
        c = obs_cam[o];
        if (p < 0 || p >= N || c < 0 || c >= M) { atomicOr(err_flag, 1); c = -1; }
        else {
            int pp = o > 0 ? obs_pt[o - 1] : -1, pc = o > 0 ? obs_cam[o - 1] : -1;
            if (p < pp || (p == pp && c <= pc)) atomicOr(err_flag, 2);
            const double2 xy = reinterpret_cast<const double2*>(obs_xy)[o];
            x[o] = xy.x / f0; y[o] = xy.y / f0;   // x/f0, y/f0 of BA.cpp:475-479, formed once
            if (p != pp) for (int q = (pp < 0 ? 0 : pp + 1); q <= p; ++q) pt_begin[q] = o;
            if (o == O - 1) for (int64_t q = p + 1; q <= N; ++q) pt_begin[q] = O;
        }
    }
    // histogram with one atomic per distinct camera of the warp (point-major order: ~10 cameras per 32 observations)
    const unsigned peers = __match_any_sync(0xffffffffu, c);
    if (c >= 0 && (threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&cam_count[c], (unsigned long long)__popc(peers));
}

__global__ void k_scan_counts(int M, const unsigned long long* __restrict__ cnt, int64_t* __restrict__ cam_begin, unsigned long long* __restrict__ cursor) {
    // single CTA, sequential chunks of 1024 with a running carry
    __shared__ unsigned long long sm[1024];
    __shared__ unsigned long long carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < M; base += 1024) {
        int i = base + threadIdx.x;
        unsigned long long v = i < M ? cnt[i] : 0ULL;
        sm[threadIdx.x] = v;
        __syncthreads();
        for (int off = 1; off < 1024; off <<= 1) {
            unsigned long long t = threadIdx.x >= off ? sm[threadIdx.x - off] : 0ULL;
            __syncthreads();
            sm[threadIdx.x] += t;
            __syncthreads();
        }
        unsigned long long excl = carry + sm[threadIdx.x] - v;
        if (i < M) { cam_begin[i] = (int64_t)excl; cursor[i] = excl; }
        __syncthreads();
        if (threadIdx.x == 1023) carry += sm[1023];
        __syncthreads();
    }
    if (threadIdx.x == 0) cam_begin[M] = (int64_t)carry;
}

__global__ void k_scatter_by_cam(int64_t O, const int32_t* __restrict__ obs_cam, const int32_t* __restrict__ obs_pt, const double* __restrict__ x,
                                 const double* __restrict__ y, unsigned long long* __restrict__ cursor, int32_t* __restrict__ c_pt,
                                 double* __restrict__ c_x, double* __restrict__ c_y) {
    int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int c = o < O ? obs_cam[o] : -1;
    const unsigned peers = __match_any_sync(0xffffffffu, c);
    const int lane = threadIdx.x & 31, leader = __ffs(peers) - 1;
    unsigned long long base = 0;
    if (c >= 0 && lane == leader) base = atomicAdd(&cursor[c], (unsigned long long)__popc(peers));   // one atomic per distinct camera of the warp
    base = __shfl_sync(0xffffffffu, base, leader);
    if (c < 0) return;
    const unsigned long long pos = base + __popc(peers & ((1u << lane) - 1));
    c_pt[pos] = obs_pt[o]; c_x[pos] = x[o]; c_y[pos] = y[o];
}

__global__ void k_finite_flag(int64_t n, const double* __restrict__ v, int* __restrict__ flag) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && !isfinite(v[i])) atomicOr(flag, 1);
}
__global__ void k_count_skipped(int64_t N, const unsigned char* __restrict__ skipped, unsigned long long* __restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned v = (i < N && skipped[i]) ? 1u : 0u;
    unsigned b = __ballot_sync(0xffffffffu, v);
    if ((threadIdx.x & 31) == 0 && b) atomicAdd(out, (unsigned long long)__popc(b));
}

static inline unsigned cdiv(int64_t a, int64_t b) { return (unsigned)((a + b - 1) / b); }

void launch_prep_obs(cudaStream_t st, int64_t O, int64_t N, int M, const int32_t* obs_cam, const int32_t* obs_pt, const double* obs_xy, double f0,
                     double* x, double* y, int64_t* pt_begin, unsigned long long* cam_count, int* err_flag) {
    if (O > 0) k_prep_obs<<<cdiv(O, 256), 256, 0, st>>>(O, N, M, obs_cam, obs_pt, obs_xy, f0, x, y, pt_begin, cam_count, err_flag);
}
void launch_prep_index(cudaStream_t st, int64_t O, int64_t N, int M, const int32_t* obs_cam, const int32_t* obs_pt, int64_t* pt_begin, unsigned long long* cam_count,
                       int* err_flag) {
    if (O > 0) k_prep_index<<<cdiv(O, 256), 256, 0, st>>>(O, N, M, obs_cam, obs_pt, pt_begin, cam_count, err_flag);
}
void launch_scatter_index(cudaStream_t st, int64_t O, int64_t N, int M, const int32_t* obs_cam, const int32_t* obs_pt, unsigned long long* cursor, int32_t* c_pt,
                          unsigned* obs_pos) {
    if (O > 0) k_scatter_index<<<cdiv(O, 256), 256, 0, st>>>(O, N, M, obs_cam, obs_pt, cursor, c_pt, obs_pos);
}
void launch_prep_xy(cudaStream_t st, int64_t O, const double* obs_xy, double f0, const unsigned* obs_pos, double* x, double* y, double* c_x, double* c_y) {
    if (O > 0) k_prep_xy<<<cdiv(O, 256), 256, 0, st>>>(O, obs_xy, f0, obs_pos, x, y, c_x, c_y);
}
void launch_scan_counts(cudaStream_t st, int M, const unsigned long long* cnt, int64_t* cam_begin, unsigned long long* cursor) {
    k_scan_counts<<<1, 1024, 0, st>>>(M, cnt, cam_begin, cursor);
}
void launch_scatter_by_cam(cudaStream_t st, int64_t O, const int32_t* obs_cam, const int32_t* obs_pt, const double* x, const double* y,
                           unsigned long long* cursor, int32_t* c_pt, double* c_x, double* c_y) {
    if (O > 0) k_scatter_by_cam<<<cdiv(O, 256), 256, 0, st>>>(O, obs_cam, obs_pt, x, y, cursor, c_pt, c_x, c_y);
}
void launch_finite_flag(cudaStream_t st, int64_t n, const double* v, int* flag) {
    if (n > 0) k_finite_flag<<<cdiv(n, 256), 256, 0, st>>>(n, v, flag);
}
void launch_count_skipped(cudaStream_t st, int64_t N, const unsigned char* skipped, unsigned long long* out) {
    if (N > 0) k_count_skipped<<<cdiv(N, 256), 256, 0, st>>>(N, skipped, out);
}

}  // namespace srk
