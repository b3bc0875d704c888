// Variants of the 64x64 potrf step loop, to find what bounds it (development aid).
#include <cstdio>
constexpr int NB = 64;
template <int VAR>
__global__ void __launch_bounds__(256) k(double* A, long long* cyc) {
    extern __shared__ double smp[];
    double (*T)[NB + 1] = reinterpret_cast<double (*)[NB + 1]>(smp);
    double (*W)[NB + 1] = reinterpret_cast<double (*)[NB + 1]>(smp + NB * (NB + 1));
    const int tid = threadIdx.x, r = tid & 63, grp = tid >> 6;
    for (int e = tid; e < NB * NB; e += 256) { int c = e >> 6, rr = e & 63; T[c][rr] = (rr >= c) ? A[c * NB + rr] : 0.0; W[c][rr] = (rr == c) ? 1.0 : 0.0; }
    __syncthreads();
    long long c0 = clock64();
#pragma unroll 1
    for (int j = 0; j < NB; ++j) {
        const double d = T[j][j];
        double rs;
        if (VAR == 2) rs = d * 0.001; else rs = rsqrt(d);
        const double lr = T[j][r] * rs;
        double lc[16], cur[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const int c = grp + 4 * i;
            const bool inT = c > j;
            if (VAR == 1) { lc[i] = T[j][c]; cur[i] = T[c][r]; }
            else { lc[i] = inT ? T[j][c] : W[c][j]; cur[i] = inT ? T[c][r] : W[c][r]; }
        }
        if (VAR != 3) __syncthreads();
        if (VAR == 4) {   // branch-free stores: always write back own elements
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const int c = grp + 4 * i;
                const bool inT = c > j;
                const bool act = (r > j) && (!inT || c <= r);
                const double v = act ? cur[i] - lr * (lc[i] * rs) : ((r == j && !inT) ? cur[i] * rs : cur[i]);
                if (inT) T[c][r] = v; else W[c][r] = v;
            }
            if (grp == (j & 3) && r >= j) T[j][r] = (r == j) ? d * rs : lr;
        } else if (r > j) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const int c = grp + 4 * i;
                const double v = cur[i] - lr * (lc[i] * rs);
                if (VAR == 1) { if (c > j && c <= r) T[c][r] = v; }
                else { if (c > j) { if (c <= r) T[c][r] = v; } else W[c][r] = v; }
            }
            if (grp == (j & 3)) T[j][r] = lr;
        } else if (r == j) {
            if (VAR != 1) {
#pragma unroll
                for (int i = 0; i < 16; ++i) { const int c = grp + 4 * i; if (c <= j) W[c][j] = cur[i] * rs; }
            }
            if (grp == (j & 3)) T[j][j] = d * rs;
        }
        if (VAR != 3) __syncthreads();
    }
    long long c1 = clock64();
    for (int e = tid; e < NB * NB; e += 256) { int c = e >> 6, rr = e & 63; A[c * NB + rr] = T[c][rr] + W[c][rr]; }
    if (tid == 0) cyc[VAR] = c1 - c0;
}
__global__ void __launch_bounds__(256) k5(double* A, long long* cyc) {
    extern __shared__ double smp[];
    double (*T)[NB + 1] = reinterpret_cast<double (*)[NB + 1]>(smp);
    double (*W)[NB + 1] = reinterpret_cast<double (*)[NB + 1]>(smp + NB * (NB + 1));
    double* rsv = smp + 2 * NB * (NB + 1);
    const int tid = threadIdx.x, r = tid & 63, grp = tid >> 6;
    for (int e = tid; e < NB * NB; e += 256) { int c = e >> 6, rr = e & 63; T[c][rr] = (rr >= c) ? A[c * NB + rr] : 0.0; W[c][rr] = (rr == c) ? 1.0 : 0.0; }
    __syncthreads();
    long long c0 = clock64();
#pragma unroll 1
    for (int j = 0; j < NB; ++j) {
        if (r > j) {
            const double d = T[j][j];
            const double a = T[j][r] * (1.0 / d);
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const int c = grp + 4 * i;
                if (c > j) { if (c <= r) T[c][r] -= a * T[j][c]; }
                else W[c][r] -= a * W[c][j];
            }
        } else if (tid == j) rsv[j] = rsqrt(T[j][j]);
        __syncthreads();
    }
    long long c1 = clock64();
    for (int e = tid; e < NB * NB; e += 256) { int c = e >> 6, rr = e & 63; A[c * NB + rr] = T[c][rr] * rsv[c] + W[c][rr] * rsv[rr]; }
    if (tid == 0) cyc[5] = c1 - c0;
}
__global__ void __launch_bounds__(256) k6(double* A, long long* cyc) {
    __shared__ double col[2][NB];
    __shared__ double wrow[2][NB];
    __shared__ double rsv[NB];
    const int tid = threadIdx.x, r = tid & 63, grp = tid >> 6;
    double t[16], w[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { const int c = grp + 4 * i; t[i] = (r >= c) ? A[c * NB + r] : 0.0; w[i] = (r == c) ? 1.0 : 0.0; }
    long long c0 = clock64();
#pragma unroll 1
    for (int j = 0; j < NB; ++j) {
        const int buf = j & 1;
#pragma unroll
        for (int i = 0; i < 16; ++i) if (grp + 4 * i == j) col[buf][r] = t[i];
        if (r == j) {
#pragma unroll
            for (int i = 0; i < 16; ++i) wrow[buf][grp + 4 * i] = w[i];
        }
        __syncthreads();
        const double d = col[buf][j];
        if (tid == j) rsv[j] = rsqrt(d);
        if (r > j) {
            const double a = col[buf][r] * (1.0 / d);
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const int c = grp + 4 * i;
                if (c > j) { if (c <= r) t[i] -= a * col[buf][c]; }
                else w[i] -= a * wrow[buf][c];
            }
        }
    }
    __syncthreads();
    long long c1 = clock64();
#pragma unroll
    for (int i = 0; i < 16; ++i) { const int c = grp + 4 * i; A[c * NB + r] = t[i] * rsv[c] + w[i] * rsv[r]; }
    if (tid == 0) cyc[6] = c1 - c0;
}
__global__ void __launch_bounds__(256) k7(double* A, long long* cyc) {
    __shared__ double col[2][NB];
    __shared__ double wrow[2][NB];
    __shared__ double rsv[NB];
    const int tid = threadIdx.x, r = tid & 63, grp = tid >> 6;
    double t[16], w[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { const int c = grp + 4 * i; t[i] = (r >= c) ? A[c * NB + r] : 0.0; w[i] = (r == c) ? 1.0 : 0.0; }
    long long c0 = clock64();
#pragma unroll 1
    for (int j = 0; j < NB; ++j) {
        const int buf = j & 1;
        double mine = 0.0;
#pragma unroll
        for (int i = 0; i < 16; ++i) mine = (grp + 4 * i == j) ? t[i] : mine;
        if ((j & 3) == grp) col[buf][r] = mine;
        if (r == j) {
#pragma unroll
            for (int i = 0; i < 16; ++i) wrow[buf][grp + 4 * i] = w[i];
        }
        __syncthreads();
        const double d = col[buf][j];
        if (tid == j) rsv[j] = rsqrt(d);
        const double a = (r > j) ? col[buf][r] * (1.0 / d) : 0.0;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const int c = grp + 4 * i;
            const double cv = col[buf][c], wv = wrow[buf][c];
            const bool inT = c > j;
            const double val = inT ? cv : wv;
            const double delta = a * val;
            t[i] -= (inT && c <= r) ? delta : 0.0;
            w[i] -= inT ? 0.0 : delta;
        }
    }
    __syncthreads();
    long long c1 = clock64();
#pragma unroll
    for (int i = 0; i < 16; ++i) { const int c = grp + 4 * i; A[c * NB + r] = t[i] * rsv[c] + w[i] * rsv[r]; }
    if (tid == 0) cyc[7] = c1 - c0;
}
template <int G>   // G column groups, blockDim = 64*G, 64/G columns per thread
__global__ void __launch_bounds__(64 * G) k8(double* A, long long* cyc) {
    constexpr int E = NB / G;
    __shared__ double col[2][NB];
    __shared__ double wrow[2][NB];
    __shared__ double rsv[NB];
    const int tid = threadIdx.x, r = tid & 63, grp = tid >> 6;
    double t[E], w[E];
#pragma unroll
    for (int i = 0; i < E; ++i) { const int c = grp + G * i; t[i] = (r >= c) ? A[c * NB + r] : 0.0; w[i] = (r == c) ? 1.0 : 0.0; }
    long long c0 = clock64();
#pragma unroll 1
    for (int j = 0; j < NB; ++j) {
        const int buf = j & 1;
#pragma unroll
        for (int i = 0; i < E; ++i) if (grp + G * i == j) col[buf][r] = t[i];
        if (r == j) {
#pragma unroll
            for (int i = 0; i < E; ++i) wrow[buf][grp + G * i] = w[i];
        }
        __syncthreads();
        const double d = col[buf][j];
        if (tid == j) rsv[j] = rsqrt(d);
        if (r > j) {
            const double a = col[buf][r] * (1.0 / d);
#pragma unroll
            for (int i = 0; i < E; ++i) {
                const int c = grp + G * i;
                if (c > j) { if (c <= r) t[i] -= a * col[buf][c]; }
                else w[i] -= a * wrow[buf][c];
            }
        }
    }
    __syncthreads();
    long long c1 = clock64();
#pragma unroll
    for (int i = 0; i < E; ++i) { const int c = grp + G * i; A[c * NB + r] = t[i] * rsv[c] + w[i] * rsv[r]; }
    if (tid == 0) cyc[G == 8 ? 4 : 3] = c1 - c0;
}
__device__ __forceinline__ double fast_rcp(double d) {   // 1/d to ~1 ulp: FP32 seed + 3 Newton steps in FP64 (shorter dependent chain than the library division)
    double x = (double)__frcp_rn((float)d);
    x = x * (2.0 - d * x);
    x = x * (2.0 - d * x);
    x = x + x * (1.0 - d * x);
    return x;
}
template <int G>
__global__ void __launch_bounds__(64 * G) k9(double* A, long long* cyc) {
    constexpr int E = NB / G;
    __shared__ double col[2][NB];
    __shared__ double wrow[2][NB];
    __shared__ double dv[NB];
    const int tid = threadIdx.x, r = tid & 63, grp = tid >> 6;
    double t[E], w[E];
#pragma unroll
    for (int i = 0; i < E; ++i) { const int c = grp + G * i; t[i] = (r >= c) ? A[c * NB + r] : 0.0; w[i] = (r == c) ? 1.0 : 0.0; }
    long long c0 = clock64();
#pragma unroll
    for (int i = 0; i < E; ++i) {
#pragma unroll 1
        for (int g = 0; g < G; ++g) {
            const int j = i * G + g;
            const int buf = g & 1;
            if (grp == g) col[buf][r] = t[i];
            if (r == j) {
#pragma unroll
                for (int i2 = 0; i2 <= i; ++i2) wrow[buf][grp + G * i2] = w[i2];
            }
            __syncthreads();
            const double d = col[buf][j];
            if (tid == j) dv[j] = d;
            if (r > j) {
                const double a = col[buf][r] * fast_rcp(d);
#pragma unroll
                for (int i2 = 0; i2 < E; ++i2) {
                    const int c = grp + G * i2;
                    if (i2 < i) w[i2] -= a * wrow[buf][c];                       // c < j
                    else if (i2 > i) { if (c <= r) t[i2] -= a * col[buf][c]; }  // c > j
                    else { if (grp > g) { if (c <= r) t[i2] -= a * col[buf][c]; } else w[i2] -= a * wrow[buf][c]; }
                }
            }
        }
    }
    __syncthreads();
    long long c1 = clock64();
#pragma unroll
    for (int i = 0; i < E; ++i) { const int c = grp + G * i; A[c * NB + r] = t[i] * rsqrt(dv[c]) + w[i] * rsqrt(dv[r]); }
    if (tid == 0) cyc[G == 8 ? 4 : 3] = c1 - c0;
}
int main() {
    double* A; long long* cyc; cudaMalloc(&A, NB * NB * 8); cudaMalloc(&cyc, 64);
    double h[NB * NB];
    const size_t sm = sizeof(double) * 2 * NB * (NB + 1);
    cudaFuncSetAttribute(k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm); cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    cudaFuncSetAttribute(k<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm); cudaFuncSetAttribute(k<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    cudaFuncSetAttribute(k<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    for (int rep = 0; rep < 2; ++rep) {
        for (int v = 0; v < 5; ++v) {
            for (int c = 0; c < NB; ++c) for (int r = 0; r < NB; ++r) h[c * NB + r] = (r == c) ? 70.0 : 0.5 / (1 + abs(r - c));
            cudaMemcpy(A, h, sizeof(h), cudaMemcpyHostToDevice);
            if (v == 0) k<0><<<1, 256, sm>>>(A, cyc); if (v == 1) k<1><<<1, 256, sm>>>(A, cyc); if (v == 2) k<2><<<1, 256, sm>>>(A, cyc);
            if (v == 3) k<3><<<1, 256, sm>>>(A, cyc); if (v == 4) k<4><<<1, 256, sm>>>(A, cyc);
            cudaDeviceSynchronize();
        }
    }
    cudaFuncSetAttribute(k5, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm + 512);
    for (int rep = 0; rep < 2; ++rep) {
        for (int c = 0; c < NB; ++c) for (int r = 0; r < NB; ++r) h[c * NB + r] = (r == c) ? 70.0 : 0.5 / (1 + abs(r - c));
        cudaMemcpy(A, h, sizeof(h), cudaMemcpyHostToDevice);
        k5<<<1, 256, sm + 512>>>(A, cyc); cudaDeviceSynchronize();
    }
    for (int rep = 0; rep < 2; ++rep) {
        for (int c = 0; c < NB; ++c) for (int r = 0; r < NB; ++r) h[c * NB + r] = (r == c) ? 70.0 : 0.5 / (1 + abs(r - c));
        cudaMemcpy(A, h, sizeof(h), cudaMemcpyHostToDevice);
        k6<<<1, 256>>>(A, cyc); cudaDeviceSynchronize();
    }
    { long long hc2[8]; cudaMemcpy(hc2, cyc, 64, cudaMemcpyDeviceToHost); printf("register resident, rolled, 1 barrier: loop %lld cycles (%.0f per step)\n", hc2[6], hc2[6] / 64.0);
      double o[NB*NB]; cudaMemcpy(o, A, sizeof(o), cudaMemcpyDeviceToHost); printf("  L(0,0)+Linv(0,0)=%.6f (expect %.6f)\n", o[0], sqrt(70.0) + 1/sqrt(70.0)); }
    for (int rep = 0; rep < 2; ++rep) {
        for (int c = 0; c < NB; ++c) for (int r = 0; r < NB; ++r) h[c * NB + r] = (r == c) ? 70.0 : 0.5 / (1 + abs(r - c));
        cudaMemcpy(A, h, sizeof(h), cudaMemcpyHostToDevice);
        k7<<<1, 256>>>(A, cyc); cudaDeviceSynchronize();
    }
    { long long hc2[8]; cudaMemcpy(hc2, cyc, 64, cudaMemcpyDeviceToHost); printf("register resident, branch-free: loop %lld cycles (%.0f per step)\n", hc2[7], hc2[7] / 64.0);
      double o[NB*NB]; cudaMemcpy(o, A, sizeof(o), cudaMemcpyDeviceToHost); printf("  L(0,0)+Linv(0,0)=%.6f (expect %.6f)\n", o[0], sqrt(70.0) + 1/sqrt(70.0)); }
    for (int rep = 0; rep < 2; ++rep) {
        for (int g = 0; g < 2; ++g) {
            for (int c = 0; c < NB; ++c) for (int r = 0; r < NB; ++r) h[c * NB + r] = (r == c) ? 70.0 : 0.5 / (1 + abs(r - c));
            cudaMemcpy(A, h, sizeof(h), cudaMemcpyHostToDevice);
            if (g == 0) k8<8><<<1, 512>>>(A, cyc); else k8<16><<<1, 1024>>>(A, cyc);
            cudaDeviceSynchronize();
            long long hc2[8]; cudaMemcpy(hc2, cyc, 64, cudaMemcpyDeviceToHost);
            double o[NB*NB]; cudaMemcpy(o, A, sizeof(o), cudaMemcpyDeviceToHost);
            if (rep == 1) printf("register resident, %d threads: loop %lld cycles (%.0f per step) check %.6f\n", g == 0 ? 512 : 1024, hc2[g == 0 ? 4 : 3], hc2[g == 0 ? 4 : 3] / 64.0, o[0]);
        }
    }
    for (int rep = 0; rep < 2; ++rep) {
        for (int g = 0; g < 2; ++g) {
            for (int c = 0; c < NB; ++c) for (int r = 0; r < NB; ++r) h[c * NB + r] = (r == c) ? 70.0 : 0.5 / (1 + abs(r - c));
            cudaMemcpy(A, h, sizeof(h), cudaMemcpyHostToDevice);
            if (g == 0) k9<8><<<1, 512>>>(A, cyc); else k9<4><<<1, 256>>>(A, cyc);
            cudaDeviceSynchronize();
            long long hc2[8]; cudaMemcpy(hc2, cyc, 64, cudaMemcpyDeviceToHost);
            double o[NB*NB]; cudaMemcpy(o, A, sizeof(o), cudaMemcpyDeviceToHost);
            if (rep == 1) printf("static-split loops, %d threads: loop %lld cycles (%.0f per step) check %.9f %.9f\n", g == 0 ? 512 : 256, hc2[g == 0 ? 4 : 3], hc2[g == 0 ? 4 : 3] / 64.0, o[0], o[5 * NB + 9]);
        }
    }
    long long hc[8]; cudaMemcpy(hc, cyc, 64, cudaMemcpyDeviceToHost);
    printf("deferred scaling, 1 barrier: loop %lld cycles (%.0f per step)\n", hc[5], hc[5] / 64.0);
    const char* nm[5] = {"baseline", "T only (no W)", "no rsqrt", "no barriers", "branch-free stores"};
    for (int v = 0; v < 5; ++v) printf("%-20s loop %lld cycles (%.0f per step) err=%s\n", nm[v], hc[v], hc[v] / 64.0, cudaGetErrorString(cudaGetLastError()));
    return 0;
}
