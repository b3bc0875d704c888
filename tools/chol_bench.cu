// Micro-benchmark of the dense Cholesky pieces (development aid, not part of the library).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I surikatoko_b200/csrc tools/chol_bench.cu surikatoko_b200/_lib/chol_kernels.o -o tools/chol_bench
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include "kernels.h"
#include "solve_order.h"
using namespace srk;
int main(int argc, char** argv) {
    int n = argc > 1 ? atoi(argv[1]) : 9993;
    int band = argc > 2 ? atoi(argv[2]) : 100;   // half bandwidth; 0 = dense
    int nd = argc > 3 ? atoi(argv[3]) : 0;       // 1: nested-dissection order + partitioned factorisation (solve_order.h)
    int64_t ld = (n + 7) & ~7;
    std::vector<double> h((size_t)ld * n, 0.0);
    srand(1);
    for (int c = 0; c < n; ++c) {
        for (int r = c; r < n; ++r) {
            bool in = band == 0 || (r - c) <= band || (r - c) >= n - band;
            if (!in) continue;
            double v = (r == c) ? (band == 0 ? n : 4.0 * band) + 1.0 : (rand() / (double)RAND_MAX - 0.5);
            h[(size_t)c * ld + r] = v;
        }
    }
    double *A, *L, *ws, *b; int* info;
    cudaMalloc(&A, sizeof(double) * ld * n); cudaMalloc(&L, sizeof(double) * ld * n);
    cudaMalloc(&ws, sizeof(double) * dense_cholesky_dinv_doubles(n)); cudaMalloc(&b, sizeof(double) * ld); cudaMalloc(&info, 4);
    cudaMemcpy(A, h.data(), sizeof(double) * ld * n, cudaMemcpyHostToDevice);
    std::vector<double> hb(ld, 1.0);
    cudaStream_t st; cudaStreamCreate(&st);
    cudaEvent_t e0, e1, e2, e3; cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2); cudaEventCreate(&e3);
    for (int it = 0; it < 4; ++it) {
        cudaMemcpyAsync(L, A, sizeof(double) * ld * n, cudaMemcpyDeviceToDevice, st);
        cudaMemcpyAsync(b, hb.data(), sizeof(double) * ld, cudaMemcpyHostToDevice, st);
        cudaEventRecord(e0, st);
        int64_t nl = dense_cholesky_factor(st, n, L, ld, ws, info);
        cudaEventRecord(e1, st);
        dense_cholesky_forward(st, n, L, ld, ws, b);
        cudaEventRecord(e2, st);
        dense_cholesky_backward(st, n, L, ld, ws, b);
        cudaEventRecord(e3, st);
        cudaStreamSynchronize(st);
        float f, a, c; cudaEventElapsedTime(&f, e0, e1); cudaEventElapsedTime(&a, e1, e2); cudaEventElapsedTime(&c, e2, e3);
        int hinfo; cudaMemcpy(&hinfo, info, 4, cudaMemcpyDeviceToHost);
        printf("n=%d band=%d iter %d: factor %.3f ms (%lld launches, %.2f TFLOP/s dense-equivalent) fwd %.3f ms bwd %.3f ms info=%d err=%s\n", n, band, it, f,
               (long long)nl, (double)n * n * n / 3.0 / (f * 1e-3) / 1e12, a, c, hinfo, cudaGetErrorString(cudaGetLastError()));
        if (it == 3) { dense_cholesky_profile_report(); dense_cholesky_band_profile_report(); }
        else if (it == 2) { cudaDeviceSynchronize(); long long z[16] = {0}; (void)z; }
    }
    if (nd) {
        // groups of 10 unknowns (the first two 4 and 9, as in the gauge-reduced camera system), coupling read off the matrix
        std::vector<int> gsize; { int left = n; gsize.push_back(4); gsize.push_back(9); left -= 13; while (left > 0) { gsize.push_back(left < 10 ? left : 10); left -= 10; } }
        const int G = (int)gsize.size();
        std::vector<int> gof(n); { int i = 0; for (int g = 0; g < G; ++g) for (int a = 0; a < gsize[g]; ++a) gof[i++] = g; }
        std::vector<unsigned char> adj((size_t)G * G, 0);
        for (int c = 0; c < n; ++c) for (int r = c; r < n; ++r) if (h[(size_t)c * ld + r] != 0.0) { adj[(size_t)gof[r] * G + gof[c]] = 1; adj[(size_t)gof[c] * G + gof[r]] = 1; }
        SolveOrder o = build_solve_order(G, gsize.data(), adj.data());
        printf("order: active=%d n=%d np=%d levels=%d sep_levels=%d parts=%d max_part_blocks=%d ksep=%d sep_blocks=%d\n", (int)o.active, o.n, o.np, o.levels, o.sep_levels,
               o.part.nparts, o.max_part_blocks, o.part.ksep, o.sep_blocks);
        if (o.active) {
            const int np = o.np; const int64_t ldp = (np + 7) & ~7;
            double *Lp, *wsp, *bp, *xb; int* srcd;
            cudaMalloc(&Lp, sizeof(double) * ldp * np); cudaMalloc(&wsp, sizeof(double) * dense_cholesky_dinv_doubles(np)); cudaMalloc(&bp, sizeof(double) * ldp);
            cudaMalloc(&xb, sizeof(double) * ld); cudaMalloc(&srcd, sizeof(int) * np);
            cudaMemcpy(srcd, o.src.data(), sizeof(int) * np, cudaMemcpyHostToDevice);
            cudaEvent_t ep; cudaEventCreate(&ep);
            for (int it = 0; it < 4; ++it) {
                cudaMemcpyAsync(b, hb.data(), sizeof(double) * ld, cudaMemcpyHostToDevice, st);
                cudaEventRecord(ep, st);
                launch_permute_sym(st, n, A, ld, 0, np, srcd, Lp, ldp);
                launch_gather_vec(st, np, srcd, b, bp);
                cudaEventRecord(e0, st);
                int64_t nl = dense_cholesky_factor(st, np, Lp, ldp, wsp, info, &o.part);
                cudaEventRecord(e1, st);
                dense_cholesky_forward(st, np, Lp, ldp, wsp, bp, &o.part);
                cudaEventRecord(e2, st);
                dense_cholesky_backward(st, np, Lp, ldp, wsp, bp, &o.part);
                cudaEventRecord(e3, st);
                launch_scatter_vec(st, np, srcd, bp, b);
                cudaStreamSynchronize(st);
                float pm, f, a, c; cudaEventElapsedTime(&pm, ep, e0); cudaEventElapsedTime(&f, e0, e1); cudaEventElapsedTime(&a, e1, e2); cudaEventElapsedTime(&c, e2, e3);
                int hinfo; cudaMemcpy(&hinfo, info, 4, cudaMemcpyDeviceToHost);
                printf("ND n=%d np=%d iter %d: permute %.3f ms factor %.3f ms (%lld launches) fwd %.3f ms bwd %.3f ms info=%d err=%s\n", n, np, it, pm, f, (long long)nl, a, c, hinfo,
                       cudaGetErrorString(cudaGetLastError()));
            }
        }
    }
    // residual check: A x = 1
    std::vector<double> x(ld); cudaMemcpy(x.data(), b, sizeof(double) * ld, cudaMemcpyDeviceToHost);
    double maxr = 0;
    for (int r = 0; r < n; r += 97) {
        double sacc = 0;
        for (int c = 0; c < n; ++c) { double v = r >= c ? h[(size_t)c * ld + r] : h[(size_t)r * ld + c]; sacc += v * x[c]; }
        maxr = fmax(maxr, fabs(sacc - 1.0));
    }
    printf("max |A x - 1| over sampled rows = %.3e\n", maxr);
    return 0;
}
