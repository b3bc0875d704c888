/*
 * suriko-b200 — C ABI of the B200-native bundle-adjustment engine.
 *
 * Drop-in boundary for suriko-engine's Kanatani BA path.  "BA.h" / "BA.cpp" below are
 *   /root/reference/cpp_impl/suriko-engine/include/suriko/bundle-adj-kanatani.h
 *   /root/reference/cpp_impl/suriko-engine/src/bundle-adj-kanatani.cpp
 * Every entry point is extern "C", takes plain pointers and sizes (host memory unless the name says "dev"),
 * never throws and never aborts: the return value is 0 on success or a negative SRK_E_* code.
 *
 * Threading contract (same as the reference, BA.h:136-163): one handle = one CUDA device + stream; calls on a
 * handle are serialised by the caller; a handle may be reused across problems (work buffers are cached).
 *
 * There is no CPU fallback: srk_ba_create fails with SRK_E_NO_DEVICE when no CUDA device is usable.
 */
#ifndef SRK_BA_C_API_H
#define SRK_BA_C_API_H

#include <stdint.h>

#if defined(__GNUC__)
#define SRK_API __attribute__((visibility("default")))
#else
#define SRK_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define SRK_OK 0
#define SRK_E_INVALID_ARG (-1)   /* null pointer, f0 ~ 0 (BA.cpp:420), bad unity_comp_ind (BA.cpp:628), unsorted observations */
#define SRK_E_NO_DEVICE (-2)     /* no CUDA device / wrong architecture */
#define SRK_E_CUDA (-3)          /* a CUDA call failed; srk_last_error() has the text */
#define SRK_E_NOT_BOUND (-4)     /* run/fetch/debug call before srk_ba_bind */
#define SRK_E_TOO_LARGE (-5)     /* reduced camera system does not fit the selected solver */
#define SRK_E_NOT_POSDEF (-6)    /* EKF: the innovation covariance is not numerically positive definite; state left untouched */

/* Stop reasons: BA.cpp:751, :866-868, :882; 5 = NormalizeSceneInplace failed (BA.cpp:681-682, empty string in the
 * reference); 6 = max_outer_iters reached (an extension: the reference has no iteration cap, quirk Q9). */
#define SRK_STOP_NONE 0
#define SRK_STOP_ABS_ERR_THRESHOLD 1      /* "abs err threshold"            -> returns true  */
#define SRK_STOP_SMALL_ERR_CHANGE 2       /* "small relative err change"    -> returns true  */
#define SRK_STOP_HESSIAN_OVERFLOW 3       /* "hessian overflow"             -> returns false */
#define SRK_STOP_ERR_CONVERGED 4          /* "err converged to limit value" -> returns false */
#define SRK_STOP_NORMALIZATION_FAILED 5
#define SRK_STOP_MAX_ITERS 6

/* Reduced-camera solver selection (north_star kernel 3). */
#define SRK_SOLVER_AUTO 0          /* dense Cholesky when the reduced system is small/dense enough, else PCG */
#define SRK_SOLVER_DENSE_CHOLESKY 1
#define SRK_SOLVER_BLOCK_PCG 2

/*
 * Flat problem.  Replaces the (FragmentMap&, vector<SE3Transform>&, CornerTrackRepository&, K) argument pack of
 * BundleAdjustmentKanatani::ComputeInplace (BA.h:179-184).  The adapter in include/suriko_compat/ builds it from the
 * reference's containers through GetCorner/EachCorner (quirks Q10, Q11 of SURVEY.md section 8).
 */
typedef struct srk_ba_problem {
    int64_t n_cams, n_points, n_obs;
    const int32_t* obs_cam;    /* [n_obs]   frame_ind                                                              */
    const int32_t* obs_point;  /* [n_obs]   pnt_ind = running index over tracks that own a SalientPointId (BA.cpp:1161-1171);
                                            observations sorted by (pnt_ind, frame_ind), at most one per pair      */
    const double* obs_xy;      /* [2*n_obs] pixels as stored in CornerData.pixel_coord (obs-geom.h:245-249)        */
    double* points;            /* [3*n_points] in/out, indexed by pnt_ind                                          */
    double* cams;              /* [12*n_cams] in/out: T[3] then R column-major[9] — byte-compatible with
                                            suriko::SE3Transform (obs-geom.h:177-190); inverse (camera-from-world) poses */
    const double* K;           /* [9] (shared) or [9*n_cams], column-major — byte-compatible with Eigen::Matrix<Scalar,3,3>;
                                            never written (quirk Q2: BA.cpp:2026-2035 drops the intrinsic corrections) */
    int32_t shared_K;          /* 1: K is one matrix for all frames (BA.h shared_intrinsic_cam_mat), 0: one per frame */
    double f0;                 /* numerical-stability scale (BA.h:128-130)                                         */
} srk_ba_problem;

typedef struct srk_ba_options {
    int32_t has_err_change;            /* BundleAdjustmentKanataniTermCriteria::AllowedReprojErrRelativeChange (BA.h:84-85) */
    double err_change;
    int32_t has_max_hessian_factor;    /* ...::MaxHessianFactor (BA.h:89-90)                                         */
    double max_hessian_factor;
    int32_t unity_comp_ind;            /* BA.h:134 unity_t1_comp_ind_, default 1                                      */
    double unity_comp_value;           /* BA.h:133 unity_t1_comp_value_, default 1.0                                  */
    int32_t max_outer_iters;           /* 0 = unlimited (reference behaviour)                                         */
    int32_t solver;                    /* SRK_SOLVER_*                                                                */
    int32_t pcg_max_iters;             /* 0 = default                                                                 */
    double pcg_rel_tol;                /* 0 = default (1e-13)                                                         */
    int32_t refine_steps;              /* dense solve: iterative-refinement steps with a double-double residual (default 1) */
} srk_ba_options;

typedef struct srk_ba_report {
    int32_t converged;                 /* return value of ComputeInplace                                             */
    int32_t stop_reason;               /* SRK_STOP_*; srk_stop_reason_string() gives OptimizationStatusString()      */
    int32_t outer_iters;               /* outer LM iterations started                                                */
    int32_t attempts;                  /* solve attempts (EstimateCorrections + ApplyCorrections + ReprojError)      */
    double err_initial, err_final;     /* sum of squared residuals in (pix/f0)^2 (BA.cpp:479-482)                     */
    double hessian_factor_final;
    int64_t seen_points;               /* number of observations (seen_points_count, BA.cpp:483)                      */
    double* err_trace;                 /* optional [err_trace_cap]: accepted error after each successful outer iteration */
    int32_t err_trace_cap, err_trace_len;
    double* attempt_trace;             /* optional [4*attempt_trace_cap]: hessian_factor, err_new, accepted, skipped_points */
    int32_t attempt_trace_cap, attempt_trace_len;
    int64_t gpu_launches;              /* kernels launched by this call                                              */
    int32_t solver_used;               /* SRK_SOLVER_DENSE_CHOLESKY or SRK_SOLVER_BLOCK_PCG                           */
    int32_t pcg_iters_last;
    double pcg_rel_res_last;           /* ||rhs - S x|| / ||rhs|| the last PCG solve stopped at (0 on the dense path)          */
    int32_t factor_failures;           /* attempts whose Cholesky factorisation met a non-positive pivot: each counts as a failed
                                          attempt and is retried with hessian_factor * 10 (the reference's QR, BA.cpp:1911, returns
                                          finite corrections there, which then fail the decrease test, :841); attempt_trace holds
                                          err_new = +inf for them                                                          */
    int32_t reserved_;
} srk_ba_report;

SRK_API void srk_ba_default_options(srk_ba_options* opt);
SRK_API const char* srk_stop_reason_string(int32_t stop_reason);
SRK_API const char* srk_last_error(void);
SRK_API int srk_abi_version(void);

/* Handle life cycle.  device_ids may be null (device 0); n_devices must be 1 (one process per GPU; multi-GPU runs use
 * one handle per rank plus srk_ba_set_allreduce). */
SRK_API int srk_ba_create(void** h, const int* device_ids, int n_devices);
SRK_API void srk_ba_destroy(void* h);
/* Run every kernel of this handle on the caller's CUDA stream (a cudaStream_t, e.g. torch.cuda.current_stream().cuda_stream)
 * instead of the handle's own stream; null restores the own stream.  Needed when the all-reduce callback is stream-ordered. */
SRK_API int srk_ba_set_stream(void* h, void* cuda_stream);

/* One-call path == BundleAdjustmentKanatani::ComputeInplace (BA.cpp:617-718): upload, normalise, LM loop, revert,
 * download.  points/cams are refined in place. */
SRK_API int srk_ba_solve(void* h, srk_ba_problem* problem, const srk_ba_options* opt, srk_ba_report* rep);

/* == static BundleAdjustmentKanatani::ReprojError (BA.cpp:589-600): no normalisation, state untouched. */
SRK_API int srk_ba_reproj_error(void* h, const srk_ba_problem* problem, double* err, int64_t* seen_points);

/* Split path (what srk_ba_solve is made of); keeps the scene resident in HBM between calls. */
SRK_API int srk_ba_bind(void* h, const srk_ba_problem* problem, const srk_ba_options* opt);  /* H2D + NormalizeSceneInplace; returns 1 if normalisation failed */
SRK_API int srk_ba_run(void* h, const srk_ba_options* opt, srk_ba_report* rep);              /* ComputeOnNormalizedWorld on the resident state */
SRK_API int srk_ba_reset(void* h);                                                           /* resident state <- state as bound (device copy) */
SRK_API int srk_ba_fetch(void* h, srk_ba_problem* problem);                                  /* RevertNormalization + D2H into points/cams */

/* Multi-GPU plumbing: the host (torch.distributed / NCCL, or gloo in CPU tests) provides an in-place sum all-reduce over
 * `count` doubles at device pointer `dev`, ordered on CUDA stream `stream`.  Each rank binds its own shard of points with
 * all cameras replicated; the engine reduces G, g_f, S, rhs and the error partials. */
typedef int (*srk_allreduce_fn)(void* user, double* dev, int64_t count, void* stream);
SRK_API int srk_ba_set_allreduce(void* h, srk_allreduce_fn fn, void* user, int rank, int world_size);
/* The same exchange without a host callback (no Python in the loop): the library binds NCCL itself (dlopen of libnccl.so.2, so
 * single-GPU users need no NCCL) and all-reduces on the handle's stream with ncclAllReduce(ncclDouble, ncclSum) -- SURVEY.md
 * section 7 step 6 / 8e.  Create the 128-byte id on rank 0 (srk_nccl_unique_id), hand it to every rank over any channel (MPI,
 * a file, torch.distributed), then every rank calls srk_ba_nccl_init (collective: it returns when all ranks have joined).
 * srk_ba_set_nccl_comm adopts a communicator the host already owns (an ncclComm_t; it is not destroyed with the handle). */
#define SRK_NCCL_UNIQUE_ID_BYTES 128
SRK_API int srk_nccl_unique_id(unsigned char* id128);
SRK_API int srk_ba_nccl_init(void* h, const unsigned char* id128, int rank, int world_size);
SRK_API int srk_ba_set_nccl_comm(void* h, void* nccl_comm, int rank, int world_size);

/* Parity hooks: one derivative pass (+ one two-phase solve at damping c when c >= 0) on the resident normalised state.
 * Any output may be null.  Layouts match oracle/srk_oracle_capi.cpp::srk_oracle_derivs_and_solve:
 *   gradE[3N+10M]; E[9N] per point row-major 3x3; G[100M] per frame row-major 10x10; Fblk[30*n_obs] per observation
 *   row-major 3x10; S[n_f*n_f] column-major (lower triangle valid, upper mirrored); rhs[n_f]; skipped[N];
 *   corrections[3N+10M] (gaps re-inserted, BA.cpp:1600-1679). */
SRK_API int srk_ba_debug_derivs_and_solve(void* h, double c, double* gradE, double* E, double* G, double* Fblk, double* S, double* rhs,
                                  unsigned char* skipped, double* corrections);
/* Same with the reduced-camera solver chosen explicitly (SRK_SOLVER_DENSE_CHOLESKY / SRK_SOLVER_BLOCK_PCG); with the PCG solver
 * S / rhs are the block-sparse system scattered into the same dense layout, *pcg_iters the iterations it took. */
SRK_API int srk_ba_debug_derivs_and_solve_ex(void* h, double c, int32_t solver, double* gradE, double* E, double* G, double* Fblk, double* S, double* rhs,
                                             unsigned char* skipped, double* corrections, int32_t* pcg_iters);
/* Resident (normalised) state as the engine holds it. */
SRK_API int srk_ba_debug_get_state(void* h, double* points, double* cams);
/* Host-only inspection hook of the elimination order (csrc/solve_order.cu), no device needed: n_groups groups of unknowns (one per
 * camera: group_size[g] unknowns, consecutive), adj[g*n_groups + h] != 0 iff two groups are coupled.  Returns 0 when the natural order
 * is kept, else the number of parts; pos[sum(group_size)] = ordered index of every unknown, *ordered_n = size with padding,
 * part_k0 / part_k1 [32] = 64-column block ranges of the parts, *ksep = first block column of the separator. */
SRK_API int srk_ba_debug_build_order(int32_t n_groups, const int32_t* group_size, const unsigned char* adj, int32_t* pos, int64_t* ordered_n, int32_t* part_k0,
                                     int32_t* part_k1, int32_t* ksep);
/* Same inputs: the second level.  Returns the number of second-level separators; mid_k0 / mid_k1 [16] = their block ranges, all inside
 * [ksep, *msep); *msep = first block column of the top separator.  A second-level separator couples exactly two leaves. */
SRK_API int srk_ba_debug_order_levels(int32_t n_groups, const int32_t* group_size, const unsigned char* adj, int32_t* mid_k0, int32_t* mid_k1, int32_t* msep);
/* ApplyCorrections (BA.cpp:1997-2063) of `corrections` to the resident state, then ReprojError. */
SRK_API int srk_ba_debug_apply(void* h, const double* corrections, double* err_new);

/* Kernel timing for bench.py: CUDA-event time (ms) of the last launch of each kernel family on this handle's stream,
 * enabled with srk_ba_set_timing(h, 1).  names: "jacobian", "frame_blocks", "schur", "solve", "backsub", "update", "residual". */
SRK_API int srk_ba_set_timing(void* h, int enabled);
SRK_API int srk_ba_get_timing(void* h, const char* name, double* ms_last, double* ms_total, int64_t* launches);
/* Structure of the last dense Cholesky factor of the reduced camera system: number of 64x64 block rows, non-zero tiles of L,
 * and the floating-point operations the factorisation executed (zero tiles are skipped, so a block-banded system costs what
 * its fill costs).  bench.py uses it for the FP64 roofline. */
SRK_API int srk_ba_solve_stats(void* h, int64_t* n_f, int64_t* block_rows, int64_t* nonzero_tiles, double* factor_flops);
/* Order in which the last dense solve factored the reduced camera system (the reference factors in capture order, BA.cpp:1911; a
 * symmetric permutation does not change the solution).  parts = 0: capture order, one chain of block_rows dependent block columns.
 * parts > 0: nested-dissection order -- `parts` independent groups of cameras factored concurrently (the longest is max_part_blocks
 * block columns) followed by a separator of separator_blocks block columns; ordered_n = n_f + padding to 64-column part boundaries.
 * SRK_SOLVE_ORDER=0 in the environment of srk_ba_create forces capture order. */
SRK_API int srk_ba_solve_order(void* h, int64_t* ordered_n, int64_t* parts, int64_t* max_part_blocks, int64_t* separator_blocks);
/* Second level of the dissection: when mid_separators > 0 the parts above are LEAVES, pairs of which are coupled through a second-level
 * separator (the longest is max_mid_blocks block columns); the dependent chain of the factorisation is then
 * max_part_blocks + max_mid_blocks + separator_blocks block columns.  SRK_SOLVE_LEVELS=1 keeps the one-level order. */
SRK_API int srk_ba_solve_levels(void* h, int64_t* mid_separators, int64_t* max_mid_blocks);
/* Block-sparse (PCG) path: stored 10x10 blocks of the reduced camera system (lower block triangle incl. the diagonal) and the PCG
 * iterations executed since srk_ba_set_timing(h, 1) -- the work bench.py's roofline of that solve is computed from. */
SRK_API int srk_ba_pcg_stats(void* h, int64_t* nnz_blocks, int64_t* iters_since_timing);

#ifdef __cplusplus
}
#endif
#endif /* SRK_BA_C_API_H */
