// suriko-b200 — K3b placeholder: the matrix-free block-Jacobi PCG is not built yet; the entry point fails loudly.
#include "pcg.h"

namespace srk {
int pcg_schur_solve(PcgWorkspace&, cudaStream_t, int64_t, int64_t, int, int, double, const int64_t*, const int32_t*, const double*, const double*,
                    const double*, double*, unsigned char*, double*, int, double, int, int, srk_allreduce_fn, void*, int64_t*, int32_t*, int) {
    return SRK_E_TOO_LARGE;
}
void pcg_release(PcgWorkspace&) {}
}  // namespace srk
