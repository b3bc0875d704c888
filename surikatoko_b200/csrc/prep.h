// suriko-b200 — launch wrappers of the bind-time index kernels (prep_kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace srk {
void launch_prep_obs(cudaStream_t st, int64_t O, int64_t N, int M, const int32_t* obs_cam, const int32_t* obs_pt, const double* obs_xy, double f0,
                     double* x, double* y, int64_t* pt_begin, unsigned long long* cam_count, int* err_flag);
// the same work split into a structure half (indices) and a value half (pixels), see prep_kernels.cu
void launch_prep_index(cudaStream_t st, int64_t O, int64_t N, int M, const int32_t* obs_cam, const int32_t* obs_pt, int64_t* pt_begin, unsigned long long* cam_count,
                       int* err_flag);
void launch_scatter_index(cudaStream_t st, int64_t O, int64_t N, int M, const int32_t* obs_cam, const int32_t* obs_pt, unsigned long long* cursor, int32_t* c_pt,
                          unsigned* obs_pos);
void launch_prep_xy(cudaStream_t st, int64_t O, const double* obs_xy, double f0, const unsigned* obs_pos, double* x, double* y, double* c_x, double* c_y);
void launch_scan_counts(cudaStream_t st, int M, const unsigned long long* cnt, int64_t* cam_begin, unsigned long long* cursor);
void launch_scatter_by_cam(cudaStream_t st, int64_t O, const int32_t* obs_cam, const int32_t* obs_pt, const double* x, const double* y,
                           unsigned long long* cursor, int32_t* c_pt, double* c_x, double* c_y);
void launch_finite_flag(cudaStream_t st, int64_t n, const double* v, int* flag);
void launch_count_skipped(cudaStream_t st, int64_t N, const unsigned char* skipped, unsigned long long* out);
}  // namespace srk
