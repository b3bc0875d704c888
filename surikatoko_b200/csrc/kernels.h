// suriko-b200 — launch wrappers of the hand-written kernels (ba_kernels.cu, chol_kernels.cu, pcg_kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "common.cuh"

namespace srk {

void launch_cam_prep(cudaStream_t st, int M, const double* cams, const double* K, int shared_K, double f0, double* camd);
void launch_jacobian(cudaStream_t st, int64_t O, const int32_t* obs_cam, const int32_t* obs_pt, const double* x, const double* y,
                     const double* X, int64_t N, const double* camd, double* J, double* Eacc = nullptr /* [9N], zero on entry: per-point sums of Jp^T Jp (6) and Jp^T rho (3) */);
void launch_frame_blocks(cudaStream_t st, int M, const int64_t* cam_begin, const int32_t* c_pt, const double* c_x, const double* c_y,
                         const double* X, int64_t N, const double* camd, double* G, double* gf, int splits);
// bind-time structure pass of K1' (per-chunk camera lists + per-observation slots), see k_chunk_tables
int64_t residual_chunks(int64_t O);
int residual_chunk_slots();
void launch_chunk_tables(cudaStream_t st, int64_t O, const int32_t* obs_cam, const int32_t* obs_pt, int* chunk_cams, int* chunk_cnt, int* chunk_pts /*int2 per chunk*/, unsigned char* obs_slot);
void launch_residual(cudaStream_t st, int64_t O, const int32_t* obs_cam, const int32_t* obs_pt, const double* x, const double* y,
                     const double* X, int64_t N, const double* camd, const int* chunk_cams, const int* chunk_cnt, const int* chunk_pts,
                     const unsigned char* obs_slot, double* partial, int nblocks, double* out);
void launch_fill_reduced(cudaStream_t st, int M, const double* G, const double* gf, double c, int unity, double* S, int64_t ld, double* rhs);
struct SchurSink;   // common.cuh
void launch_schur(cudaStream_t st, int64_t N, int64_t O, const int64_t* pt_begin, const int32_t* obs_cam, const double* J, double c, const SchurSink& sink,
                  double* pinv, unsigned char* skipped, const unsigned char* only_flagged);
// the same kernel over a device-side list of points (count read on the device; grid sized for list_cap)
void launch_schur_list(cudaStream_t st, int64_t N, int64_t O, const int64_t* pt_begin, const int32_t* obs_cam, const double* J, double c, const SchurSink& sink,
                       double* pinv, unsigned char* skipped, const int* list, const int* list_count, int list_cap);
// K2, third form (schur_v3.cu): per-point factor kernel + bind-time tile tables + single-operand DMMA tile kernel
void launch_point_factor(cudaStream_t st, int64_t N, int64_t O, const int64_t* pt_begin, const double* J, double c, double* pinv, unsigned char* skipped,
                         double* gi, double* uvec, int* exc_list, int* exc_count, int exc_cap);
// the same from the per-point sums K1 left in Eacc (launch_jacobian): one thread per point, no pass over the observations
void launch_point_finish(cudaStream_t st, int64_t N, const double* Eacc, double c, double* pinv, unsigned char* skipped, double* gi, double* uvec, int* exc_list,
                         int* exc_count, int exc_cap);
void launch_schur_tables(cudaStream_t st, int64_t N, int tile_points, const int64_t* pt_begin, const int32_t* obs_cam, int* tile_tab, int* tile_n,
                         unsigned char* obs_slot, unsigned short* pt_mask, unsigned char* deferred);
void launch_schur_v3(cudaStream_t st, int64_t N, int64_t O, int tile_points, const int64_t* pt_begin, const int32_t* obs_pt, const double* J,
                     const SchurSink& sink, const double* gi, const double* uvec, const unsigned char* skipped, const int* tile_tab, const int* tile_n,
                     const unsigned char* obs_slot, const unsigned short* pt_mask);
// Tiled Schur accumulation (register-owned camera-pair blocks per tile of points); flags the points it leaves to launch_schur.
// plan_only: structure pass -- deferred flags and (plan_keys != nullptr) the camera-pair hash of the block-sparse system.
// DMMA formulation of the same accumulation (schur_mma.cu); same deferral rule as the plan pass of launch_schur_tile.
void launch_schur_mma(cudaStream_t st, int64_t N, int64_t O, int tile_points, const int64_t* pt_begin, const int32_t* obs_cam, const double* J, double c,
                      const SchurSink& sink, double* pinv, unsigned char* skipped, unsigned char* deferred);
// dense-rows form of K2 (points seen by many cameras): rows of F and E^-1 F [3N x 10M] -> launch_gemm_nt_dmma -> S -= lower(D)
void launch_schur_rows(cudaStream_t st, int64_t N, int64_t O, const int64_t* pt_begin, const int32_t* obs_cam, const double* J, double c, const SchurSink& sink,
                       double* pinv, unsigned char* skipped, int M, double* Fall, double* Wall);
void launch_scatter_dense_schur(cudaStream_t st, int n_full, const double* D, int unity, double* S, int64_t ld);
void launch_schur_plan(cudaStream_t st, int64_t N, int tile_points, const int64_t* pt_begin, const int32_t* obs_cam, unsigned char* deferred,
                       unsigned long long* n_deferred);
void launch_schur_tile(cudaStream_t st, int64_t N, int64_t O, int tile_points, const int64_t* pt_begin, const int32_t* obs_cam, const double* J, double c,
                       const SchurSink& sink, double* pinv, unsigned char* skipped, unsigned char* deferred, int plan_only,
                       unsigned long long* plan_keys, unsigned plan_mask, int* plan_overflow);
void launch_backsub(cudaStream_t st, int64_t N, int64_t O, const int64_t* pt_begin, const int32_t* obs_cam, const double* J, const double* df,
                    const double* pinv, const unsigned char* skipped, const double* X, double* Xtry, double* dp_out, int lanes_per_point);
// K2' with one thread per observation + per-point finish; tacc: [3N] doubles, zero on entry, zero again on exit
void launch_backsub_obs(cudaStream_t st, int64_t N, int64_t O, const int32_t* obs_pt, const int32_t* obs_cam, const double* J, const double* df, const double* pinv,
                        const unsigned char* skipped, const double* X, double* Xtry, double* dp_out, double* tacc);
void launch_cam_update(cudaStream_t st, int M, const double* cams, const double* df, double* cams_try);
void launch_expand_df(cudaStream_t st, int M, const double* dfr, int unity, double* df);
void launch_normalize_points(cudaStream_t st, int64_t N, double* X, const double* cam0_dev, double s, int revert);
void launch_points_to_planes(cudaStream_t st, int64_t N, const double* aos, double* planes);
void launch_planes_to_points(cudaStream_t st, int64_t N, const double* planes, double* aos);
void launch_debug_point_blocks(cudaStream_t st, int64_t N, int64_t O, const int64_t* pt_begin, const double* J, double* E, double* gp);
void launch_debug_F_blocks(cudaStream_t st, int64_t O, const double* J, double* F);

// Dense reduced-camera solve (chol_kernels.cu).  A is column-major n x n with leading dimension ld (multiple of 8), lower
// triangle referenced and overwritten by L; ws (dense_cholesky_dinv_doubles(n) doubles) receives the inverses of the 64x64
// diagonal blocks of L, the 64x64 block structure of L (zero tiles are skipped everywhere) and the substitution flags.  *info_dev is set to the 1-based index of the
// first non-positive pivot (0 = success).  All return launch counts.
size_t dense_cholesky_dinv_doubles(int n);
void dense_cholesky_profile_report();
void dense_cholesky_band_profile_report();
// column-major 64x64 inverse of the kb-th diagonal block of L inside the workspace
inline const double* dense_cholesky_dinv_block(const double* ws, int kb) { return ws + (size_t)kb * 64 * 64; }
// C (m x n, ldc) -= A (m x K, lda) * B (n x K, ldb)^T, all column-major, DMMA 128x128 tiles; lower_only: only tiles / entries with row >= col.
// allow_split_k: few output tiles and a long contraction -> the K range is split over blockIdx.y and the slices are added atomically
void launch_gemm_nt_dmma(cudaStream_t st, int m, int n, int K, const double* A, int64_t lda, const double* B, int64_t ldb, double* C, int64_t ldc, int lower_only,
                         int allow_split_k = 0, int cta_per_tile = 0);
// Factor S (m x m, lower, as dense_cholesky_factor) on stream `hi` while Z (rows x m) <- Z * L^-T follows panel by panel on stream `lo`;
// `lo` must be the stream that produced S and Z, and has waited for all of `hi` on return.  ev: at least m / 256 + 3 events owned by the
// caller.  Returns the number of launches, -1 when the arguments do not fit (the caller then runs the two steps one after the other).
int64_t dense_cholesky_factor_trsm(cudaStream_t hi, cudaStream_t lo, int m, double* S, int64_t lds, double* ws, int* info_dev, int rows, double* Z, int64_t ldz,
                                   cudaEvent_t* ev, int nev);
// X (rows x W, W <= 512) <- X * L^-T with L the W x W lower-triangular factor at L (ldl) and dinv the stored inverses of its 64x64 diagonal
// blocks (dense_cholesky_dinv_block of the first one): the whole panel in one launch, 48-row strips.  X, L 16-byte aligned, ldx, ldl even.
void launch_strip_trsm(cudaStream_t st, int rows, int W, double* X, int64_t ldx, const double* L, int64_t ldl, const double* dinv);
// X (rows x 64 at A, lda) <- X * Linv^T with Linv a column-major 64x64 lower-triangular inverse (right-side triangular solve of one block column)
// ncols: how many of the 64 columns exist (the caller's matrix may end inside the last block)
void launch_block_right_solve(cudaStream_t st, int rows, double* A, int64_t lda, const double* dinv_block, int ncols = 64);
// Optional scheduling hint for block-sparse systems ordered by nested dissection (solve_order.h): the block columns (64 wide) of
// part p are [k0[p], k1[p]); no non-zero tile couples two different parts; [ksep, nblk) is the separator block, ordered last.
// The parts are then factored / substituted by one thread-block cluster each, concurrently, the separator afterwards.
// Two-level dissection (nmids > 0): the parts are the LEAVES, [ksep, msep) holds the second-level separators -- block columns
// [m0[i], m1[i]) of separator i, which couples two leaves to each other; different second-level separators do not touch -- and
// [msep, nblk) is the top separator.  Order of work: leaves (concurrently), second-level separators (concurrently), top.
struct CholPartition {
    static constexpr int kMaxParts = 32;
    int nparts = 0; int ksep = 0; int k0[kMaxParts]; int k1[kMaxParts];
    int nmids = 0; int msep = 0; int m0[kMaxParts / 2]; int m1[kMaxParts / 2];
};
// pattern_dev (optional, device, nblk x nblk bytes, [c*nblk + r] = 1 for a strictly lower tile that may hold a non-zero) + its count:
// the caller's knowledge of the tile structure replaces the pattern pass over A and its host round trip.  A must then be zero in
// every tile outside the pattern that the factorisation fills (SolveOrder::l_all_tiles).
int64_t dense_cholesky_factor(cudaStream_t st, int n, double* A, int64_t ld, double* ws, int* info_dev, const CholPartition* part = nullptr,
                              const unsigned char* pattern_dev = nullptr, int pattern_count = 0);
bool dense_cholesky_pattern_ok(int n, int pattern_count, bool partitioned);
int dense_cholesky_stats(cudaStream_t st, int n, const double* ws, int64_t* nblk, int64_t* nz_tiles, double* factor_flops);
// sparse_certain: the caller knows (symbolic factorisation) that L is sparse enough for the sparse-factor kernels: the dense kernel is not launched
bool dense_cholesky_trsv_is_sparse(int n, int64_t nz_tiles);
int64_t dense_cholesky_forward(cudaStream_t st, int n, const double* L, int64_t ld, double* ws, double* b, const CholPartition* part = nullptr, bool sparse_certain = false);
int64_t dense_cholesky_backward(cudaStream_t st, int n, const double* L, int64_t ld, double* ws, double* b, const CholPartition* part = nullptr, bool sparse_certain = false);
// mirror the lower triangle into the upper one
void launch_mirror_lower(cudaStream_t st, int n, double* A, int64_t ld);
// multi-GPU exchange of the non-zero 64x64 tiles of S only (aux_kernels.cu)
void launch_tile_mask(cudaStream_t st, int n, const double* S, int64_t ld, double* mask);
void launch_tile_list(cudaStream_t st, int n, const double* mask, int* list, int* count);
void launch_tile_pack(cudaStream_t st, int n, double* S, int64_t ld, const int* list, int count, double* rhs, int64_t nrhs, double* packed, int dir);
// r = b - A*x, symmetric A with both triangles stored, double-double accumulation
void launch_residual_dd(cudaStream_t st, int n, const double* A, int64_t ld, const double* x, const double* b, double* r);
void launch_axpy1(cudaStream_t st, int n, const double* d, double* x);

// aux_kernels.cu
void launch_pack_attempt(cudaStream_t st, const double* err_sum, const int* finite_flag, const unsigned long long* skipped_cnt, int rank, int world,
                         double* slots);
void launch_add_points_aos(cudaStream_t st, int64_t N, const double* X, const double* corr_aos, double* Xtry);


}  // namespace srk
