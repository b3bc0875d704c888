/* suriko-b200 -- C ABI of the input front end of the two BA demos: what turns projection
 * matrices + corner tracks into the scene that srk_ba_solve refines.
 *
 *   srk_triangulate_tracks      Triangulate3DPointByLeastSquares   obs-geom.cpp:679-727   (batched over tracks, on the device)
 *   srk_decompose_proj_mat      DecomposeProjMat                   obs-geom.cpp:606-677   (host; 36 matrices in the dinosaur demo)
 *   srk_read_matrix_from_file   ReadMatrixFromFile                 mat-serialization.cpp:12-87   (host)
 *
 * Plain pointers and sizes, no exceptions: 0 = ok, negative = error (srk_last_error() has the text).  Matrices are
 * column-major like the Eigen objects of the reference. */
#ifndef SRK_FRONTEND_C_API_H
#define SRK_FRONTEND_C_API_H
#include <stdint.h>
#include "ba_c_api.h"
#ifdef __cplusplus
extern "C" {
#endif

/* One 3D point per track, the least-squares solution of  x P(2,:) X - f0 P(0,:) X = 0,  y P(2,:) X - f0 P(1,:) X = 0  over the
 * track's corners (obs-geom.cpp:694-709; the reference solves with colPivHouseholderQr, here a streaming Givens QR per track:
 * same minimiser).  proj: [12 * n_frames], 3x4 per frame, column-major, already f0-scaled like proj_mat_per_frame_f0scaled of the
 * demo; track_begin: [n_tracks + 1] offsets into obs_frame / obs_xy; obs_xy: [2 * n_obs] pixels.  Every track needs >= 2 corners
 * (the reference CHECKs, :687).  device: CUDA device ordinal.  All pointers are host memory. */
SRK_API int srk_triangulate_tracks(int device, int64_t n_tracks, int64_t n_obs, int32_t n_frames, const int64_t* track_begin,
                                   const int32_t* obs_frame, const double* obs_xy, const double* proj, double f0, double* points_out);

/* CUDA-event duration (ms) of the triangulation kernel of the last srk_triangulate_tracks call on this thread (for bench.py). */
SRK_API double srk_triangulate_last_kernel_ms(void);

/* P[3x4] -> scale_factor, K[9] (upper triangular, K(2,2) = 1), direct camera pose {T[3], R[9] col-major} with
 * P = scale_factor * K * R^T * [I | -T]   (obs-geom.cpp:606-677).  Returns 1 when Q Q^T is not positive definite (the reference
 * returns false), 0 on success. */
SRK_API int srk_decompose_proj_mat(const double* P34, double* scale_factor, double* K, double* direct_pose);

/* Row-major text matrix with a single delimiter character (mat-serialization.cpp:12-87).  data may be null to query the shape;
 * cap = capacity of data in doubles.  Errors: file cannot be opened, a token is not a number, ragged rows, capacity too small. */
SRK_API int srk_read_matrix_from_file(const char* path, char delimiter, double* data, int64_t cap, int64_t* rows, int64_t* cols);

#ifdef __cplusplus
}
#endif
#endif /* SRK_FRONTEND_C_API_H */
