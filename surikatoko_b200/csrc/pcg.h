// suriko-b200 — K3b: block-Jacobi preconditioned conjugate gradients on the reduced camera system (pcg_kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/srk/ba_c_api.h"

namespace srk {

struct PcgWorkspace {
    void* buf[16] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    size_t cap[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    double* h_scal = nullptr;  // pinned host scalars
};

// Solves (G_c - sum_j F_j^T E_cj^-1 F_j) x = sum_j F_j^T E_cj^-1 g_pj - g_f  without forming the matrix: the operator is
// applied per point from the stored Jacobian rows.  Also fills pinv / skipped exactly as k_schur does.  x: [ld >= n_f].
int pcg_schur_solve(PcgWorkspace& ws, cudaStream_t st, int64_t N, int64_t O, int M, int unity, double c, const int64_t* pt_begin,
                    const int32_t* obs_cam, const double* J, const double* G, const double* gf, double* pinv, unsigned char* skipped, double* x,
                    int max_iters, double rel_tol, int rank, int world, srk_allreduce_fn ar, void* ar_user, int64_t* launches, int32_t* iters_out,
                    int timing);
void pcg_release(PcgWorkspace& ws);

}  // namespace srk
