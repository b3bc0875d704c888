// suriko-b200 — launch wrappers of the hand-written kernels (ba_kernels.cu, chol_kernels.cu, pcg_kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace srk {

void launch_cam_prep(cudaStream_t st, int M, const double* cams, const double* K, int shared_K, double f0, double* camd);
void launch_jacobian(cudaStream_t st, int64_t O, const int32_t* obs_cam, const int32_t* obs_pt, const double* x, const double* y,
                     const double* X, int64_t N, const double* camd, double* J);
void launch_frame_blocks(cudaStream_t st, int M, const int64_t* cam_begin, const int32_t* c_pt, const double* c_x, const double* c_y,
                         const double* X, int64_t N, const double* camd, double* G, double* gf, int splits);
void launch_residual(cudaStream_t st, int64_t O, const int32_t* obs_cam, const int32_t* obs_pt, const double* x, const double* y,
                     const double* X, int64_t N, const double* camd, double* partial, int nblocks, double* out);
void launch_fill_reduced(cudaStream_t st, int M, const double* G, const double* gf, double c, int unity, double* S, int64_t ld, double* rhs);
void launch_schur(cudaStream_t st, int64_t N, int64_t O, const int64_t* pt_begin, const int32_t* obs_cam, const double* J, double c, int unity,
                  double* S, int64_t ld, double* rhs, double* pinv, unsigned char* skipped);
void launch_backsub(cudaStream_t st, int64_t N, int64_t O, const int64_t* pt_begin, const int32_t* obs_cam, const double* J, const double* df,
                    const double* pinv, const unsigned char* skipped, const double* X, double* Xtry, double* dp_out, int lanes_per_point);
void launch_cam_update(cudaStream_t st, int M, const double* cams, const double* df, double* cams_try);
void launch_expand_df(cudaStream_t st, int M, const double* dfr, int unity, double* df);
void launch_normalize_points(cudaStream_t st, int64_t N, double* X, const double* cam0_dev, double s, int revert);
void launch_points_to_planes(cudaStream_t st, int64_t N, const double* aos, double* planes);
void launch_planes_to_points(cudaStream_t st, int64_t N, const double* planes, double* aos);
void launch_debug_point_blocks(cudaStream_t st, int64_t N, int64_t O, const int64_t* pt_begin, const double* J, double* E, double* gp);
void launch_debug_F_blocks(cudaStream_t st, int64_t O, const double* J, double* F);

// Dense reduced-camera solve (chol_kernels.cu).  A is column-major n x n with leading dimension ld, lower triangle
// referenced; b[n] is overwritten by the solution.  Returns the number of kernels launched; *info (device int) is set
// to the 1-based index of the first non-positive pivot (0 = success).
int64_t dense_cholesky_solve(cudaStream_t st, int n, double* A, int64_t ld, double* b, int* info_dev, double* work);
size_t dense_cholesky_work_doubles(int n);
// y = A*x using only the lower triangle of symmetric A (for parity hooks / refinement).
void launch_symv_lower(cudaStream_t st, int n, const double* A, int64_t ld, const double* x, double* y);
// mirror the lower triangle into the upper one (parity hook output)
void launch_mirror_lower(cudaStream_t st, int n, double* A, int64_t ld);

}  // namespace srk
