// suriko-b200 — K3b: block-Jacobi preconditioned conjugate gradients on the block-sparse reduced camera system
// (pcg_kernels.cu).  Used when the reduced system is too large for a dense factorisation (north_star kernel 3).
//
// The system lives in full frame-variable space [10M] with 10x10 blocks per co-visible camera pair (lower block triangle,
// rows = variables of the larger camera index); the 7 gauge variables are decoupled identity rows with a zero right-hand
// side, so the solution has exact zeros there (BA.cpp:1600-1679) and equals the gauge-reduced solve everywhere else.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/srk/ba_c_api.h"
#include "common.cuh"

namespace srk {

struct PcgWorkspace {
    // structure (built once per bind)
    bool structure_valid = false;
    unsigned hmask = 0;
    unsigned long long* hkeys = nullptr; size_t hkeys_cap = 0;
    int* hids = nullptr; size_t hids_cap = 0;
    int nnzb = 0;
    int* blk_cams = nullptr; size_t blk_cams_cap = 0;       // [2*nnzb] (cam_i, cam_l)
    int64_t* row_ptr = nullptr; size_t row_ptr_cap = 0;     // [M+1]
    int* row_ent = nullptr; size_t row_ent_cap = 0;         // [2*nent] (block id | transposed << 30, column camera)
    int* diag_id = nullptr; size_t diag_id_cap = 0;         // [M]
    unsigned long long* cnt = nullptr; size_t cnt_cap = 0;  // [M+2] scratch counters / cursors
    int* misc = nullptr;                                    // [4] overflow flag, block counter
    // values
    double* blocks = nullptr; size_t blocks_cap = 0;        // [100*nnzb]
    double* rhs = nullptr; size_t vec_cap = 0;              // [10M] and the PCG vectors r, z, p, y
    double *r = nullptr, *z = nullptr, *p = nullptr, *y = nullptr;
    double* diag = nullptr; size_t diag_cap = 0;            // [100M] diagonal blocks, then their inverses
    double* scal = nullptr;                                 // [16] device scalars, two parities of {rz, bb, rr, pAp}
    double* partials = nullptr; size_t partials_cap = 0;    // per-CTA partial sums of the dot products (fixed-order second stage)
    double* h_scal = nullptr;                               // pinned mirror
};

// Block structure from the tile plan of the Schur kernel (camera pairs per tile + pairs of the deferred points).
int pcg_build_structure(PcgWorkspace& ws, cudaStream_t st, int64_t N, int64_t O, int M, int tile_points, const int64_t* pt_begin, const int32_t* obs_cam,
                        unsigned char* deferred, int64_t* launches);
// Zero the values, put the damped G blocks / identity gauge rows on the diagonal (rank 0 only) and rhs = -g_f; returns the sink
// the Schur kernels accumulate into.
int pcg_begin(PcgWorkspace& ws, cudaStream_t st, int M, const double* G, const double* gf, double c, int unity, int rank, SchurSink* sink, int64_t* launches);
// Solves S x = rhs; x: [10M] (zeros at the gauge variables).  Multi-rank: ws holds this rank's partial S; the right-hand side,
// the diagonal blocks and every mat-vec result are all-reduced through `ar`.
int pcg_solve(PcgWorkspace& ws, cudaStream_t st, int M, double* x, int max_iters, double rel_tol, int world, srk_allreduce_fn ar, void* ar_user,
              int64_t* launches, int32_t* iters_out, double* rel_res_out);
// Dense copy of the assembled system in gauge-reduced space (parity hook): S [nf x nf] column-major with leading dimension ld, rhs [nf].
int pcg_debug_to_dense(PcgWorkspace& ws, cudaStream_t st, int M, int unity, double* S, int64_t ld, double* rhs, int64_t* launches);
void pcg_release(PcgWorkspace& ws);

}  // namespace srk
