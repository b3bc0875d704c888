// Development aid: what a consumer-style DMMA loop can reach on the FP64 tensor pipe.  One CTA per SM, W warps, every warp F independent
// accumulator fragments, one DMMA per fragment per k-step; operands either stay in registers (mode 0) or are reloaded from shared memory
// every k-step (mode 1: all loads at the top of the k-step, mode 2: rolling -- each operand right after its last use).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probe/dmma_micro tools/probe/dmma_micro.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
template <int F, int NOPS, int MODE>
__global__ void __launch_bounds__(512, 1) k(int iters, double* out, long long* cyc) {
    __shared__ double sm[44 * 124];
    for (int i = threadIdx.x; i < 44 * 124; i += blockDim.x) sm[i] = 1e-3 * (i % 17);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const double* p = sm + (lane & 3) * 124 + (lane >> 2);
    double acc[F][2], x[NOPS];
#pragma unroll
    for (int f = 0; f < F; ++f) { acc[f][0] = 0; acc[f][1] = 0; }
#pragma unroll
    for (int o = 0; o < NOPS; ++o) x[o] = p[8 * o];
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
        const double* q = p + ((it & 7) * 4) * 124;
        if (MODE == 1) {
#pragma unroll
            for (int o = 0; o < NOPS; ++o) x[o] = q[8 * o];
        }
#pragma unroll
        for (int f = 0; f < F; ++f) {
            dmma(acc[f][0], acc[f][1], x[f % NOPS], x[(f * 7 + 3) % NOPS]);
            if (MODE == 2 && f >= F - NOPS) x[f - (F - NOPS)] = q[8 * (f - (F - NOPS))];
        }
    }
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int f = 0; f < F; ++f) s += acc[f][0] + acc[f][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int F, int NOPS, int MODE>
void run(const char* name, int warps) {
    double* out; long long* cyc;
    cudaMalloc(&out, sizeof(double) * 148 * 512); cudaMalloc(&cyc, sizeof(long long) * 148);
    const int iters = 20000;
    k<F, NOPS, MODE><<<148, 32 * warps>>>(iters, out, cyc);
    k<F, NOPS, MODE><<<148, 32 * warps>>>(iters, out, cyc);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 148; ++i) avg += (double)h[i]; avg /= 148;
    const double dm_per_smsp = (double)iters * F * warps / 4.0;
    printf("%-44s warps %2d F %2d: %.1f cycles per DMMA per sub-partition (16 = pipe peak) -> %.0f %% of peak  [%s]\n", name, warps, F, avg / dm_per_smsp,
           100.0 * 16.0 * dm_per_smsp / avg, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out); cudaFree(cyc);
}
int main() {
    run<30, 15, 0>("registers only", 4);
    run<30, 15, 1>("operands reloaded at the top of a k-step", 4);
    run<30, 15, 2>("rolling reloads", 4);
    run<15, 8, 0>("registers only", 8);
    run<15, 8, 1>("operands reloaded at the top of a k-step", 8);
    run<15, 8, 2>("rolling reloads", 8);
    run<32, 12, 0>("registers only", 8);
    run<32, 12, 1>("operands reloaded at the top (SYRK shape)", 8);
    run<8, 6, 1>("2x4 piece: 8 fragments, 6 operands, reload at top", 16);
    run<8, 6, 2>("2x4 piece: 8 fragments, 6 operands, rolling", 16);
    run<8, 6, 1>("2x4 piece: 8 fragments, 6 operands, reload at top", 12);
    run<10, 7, 1>("10 fragments, 7 operands, reload at top", 12);
    run<15, 8, 0>("registers only", 16);
    run<15, 8, 1>("operands reloaded at the top of a k-step", 16);
    return 0;
}
