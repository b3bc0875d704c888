// FP64 latency / throughput microbenchmark (development aid).
#include <cstdio>
__global__ void k_lat(double* out, long long* cyc, double x0, int mode) {
    double x = x0 + threadIdx.x * 1e-9, y = 1.000001;
    long long c0 = clock64();
    if (mode == 0) { for (int i = 0; i < 1024; ++i) x = fma(x, y, 1e-9); }
    else if (mode == 1) { for (int i = 0; i < 256; ++i) x = rsqrt(x) + 1.5; }
    else if (mode == 2) { for (int i = 0; i < 256; ++i) x = 1.0 / x + 1.5; }
    else if (mode == 3) { for (int i = 0; i < 256; ++i) x = sqrt(x) + 1.5; }
    else if (mode == 4) { float f = (float)x; for (int i = 0; i < 1024; ++i) f = fmaf(f, 1.000001f, 1e-9f); x = f; }
    long long c1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[mode] = c1 - c0;
}
__global__ void k_thr(double* out, long long* cyc, double x0) {
    double a[8];
    for (int i = 0; i < 8; ++i) a[i] = x0 + i + threadIdx.x * 1e-9;
    long long c0 = clock64();
    for (int it = 0; it < 512; ++it)
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = fma(a[i], 1.000001, 1e-9);
    long long c1 = clock64();
    double s = 0; for (int i = 0; i < 8; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = c1 - c0;
}
int main() {
    double* out; long long* cyc; cudaMalloc(&out, 8 * 1024 * 1024); cudaMalloc(&cyc, 64);
    long long h[8];
    const char* names[5] = {"dependent DFMA x1024", "dependent rsqrt(+add) x256", "dependent 1/x(+add) x256", "dependent sqrt(+add) x256", "dependent FFMA x1024"};
    int reps[5] = {1024, 256, 256, 256, 1024};
    for (int m = 0; m < 5; ++m) {
        k_lat<<<1, 32>>>(out, cyc, 1.5, m); cudaDeviceSynchronize();
        cudaMemcpy(h, cyc, 64, cudaMemcpyDeviceToHost);
        printf("%-28s: %.1f cycles per op (1 warp)\n", names[m], (double)h[m] / reps[m]);
    }
    for (int warps = 1; warps <= 32; warps *= 2) {
        k_thr<<<148, warps * 32>>>(out, cyc, 1.5); cudaDeviceSynchronize();
        cudaMemcpy(h, cyc, 64, cudaMemcpyDeviceToHost);
        double per_sm = (double)warps * 32 * 8 * 512 / (double)h[0];
        printf("independent DFMA, %2d warps/SM: %.2f FMA lanes per clock per SM -> %.2f TFLOP/s at 1.9 GHz x 148 SMs\n", warps, per_sm, per_sm * 2 * 1.9e9 * 148 / 1e12);
    }
    return 0;
}
