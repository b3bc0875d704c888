// suriko-b200 — K3b: block-sparse reduced camera system + block-Jacobi PCG (see pcg.h).
#include <cooperative_groups.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include "kernels.h"
#include "pcg.h"
#include "prep.h"

namespace srk {

namespace {

template <class T>
cudaError_t ensure(T*& p, size_t& cap, size_t count) {
    if (p != nullptr && count <= cap) return cudaSuccess;
    if (p != nullptr) cudaFree(p);
    p = nullptr; cap = 0;
    size_t want = count < 64 ? 64 : count;
    cudaError_t e = cudaMalloc((void**)&p, want * sizeof(T));
    if (e == cudaSuccess) cap = want;
    return e;
}
inline unsigned cdiv(int64_t a, int64_t b) { return (unsigned)((a + b - 1) / b); }

// ---- structure ----------------------------------------------------------------------------------------------------------
__global__ void k_insert_diagonals(int M, unsigned long long* keys, unsigned mask, int* overflow) {
    int cam = blockIdx.x * blockDim.x + threadIdx.x;
    if (cam >= M) return;
    const unsigned long long key = ((unsigned long long)(unsigned)cam << 32) | (unsigned)cam;
    unsigned h = hash_pair(key, mask);
    for (unsigned probe = 0; probe <= mask; ++probe) {
        const unsigned long long prev = atomicCAS(&keys[h], kHashEmpty, key);
        if (prev == kHashEmpty || prev == key) return;
        h = (h + 1) & mask;
    }
    atomicExch(overflow, 1);
}
__global__ void k_count_blocks(unsigned cap, const unsigned long long* __restrict__ keys, int* __restrict__ counter) {
    unsigned s = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned v = (s < cap && keys[s] != kHashEmpty) ? 1u : 0u;
    unsigned b = __ballot_sync(0xffffffffu, v);
    if ((threadIdx.x & 31) == 0 && b) atomicAdd(counter, __popc(b));
}
// ids in slot order (deterministic given the table): exclusive scan over the occupancy, one CTA walking the table
__global__ void k_assign_ids(unsigned cap, const unsigned long long* __restrict__ keys, int* __restrict__ ids, int* __restrict__ blk_cams,
                             unsigned long long* __restrict__ row_cnt) {
    __shared__ int warp_sums[32];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (unsigned base = 0; base < cap; base += 1024) {
        const unsigned s = base + threadIdx.x;
        const unsigned long long key = s < cap ? keys[s] : kHashEmpty;
        const int occ = key != kHashEmpty ? 1 : 0;
        const unsigned bal = __ballot_sync(0xffffffffu, occ);
        const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
        if (lane == 0) warp_sums[w] = __popc(bal);
        __syncthreads();
        int off = carry;
        for (int i = 0; i < w; ++i) off += warp_sums[i];
        if (occ) {
            const int id = off + __popc(bal & ((1u << lane) - 1));
            ids[s] = id;
            const int ci = (int)(key >> 32), cl = (int)(key & 0xffffffffu);
            blk_cams[2 * id] = ci; blk_cams[2 * id + 1] = cl;
            atomicAdd(&row_cnt[ci], 1ULL);
            if (ci != cl) atomicAdd(&row_cnt[cl], 1ULL);
        }
        __syncthreads();
        if (threadIdx.x == 0) { int t = 0; for (int i = 0; i < 32; ++i) t += warp_sums[i]; carry += t; }
        __syncthreads();
    }
}
__global__ void k_fill_rows(int nnzb, const int* __restrict__ blk_cams, unsigned long long* __restrict__ cursor, int* __restrict__ row_ent, int* __restrict__ diag_id) {
    int id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= nnzb) return;
    const int ci = blk_cams[2 * id], cl = blk_cams[2 * id + 1];
    unsigned long long pos = atomicAdd(&cursor[ci], 1ULL);
    row_ent[2 * pos] = id; row_ent[2 * pos + 1] = cl;
    if (ci != cl) {
        pos = atomicAdd(&cursor[cl], 1ULL);
        row_ent[2 * pos] = id | (1 << 30); row_ent[2 * pos + 1] = ci;
    } else {
        diag_id[ci] = id;
    }
}
// fixed summation order of the mat-vec: every row's entries sorted by column camera
__global__ void k_sort_rows(int M, const int64_t* __restrict__ row_ptr, int* __restrict__ row_ent) {
    int cam = blockIdx.x * blockDim.x + threadIdx.x;
    if (cam >= M) return;
    const int64_t b = row_ptr[cam], e = row_ptr[cam + 1];
    for (int64_t i = b + 1; i < e; ++i) {
        const int id = row_ent[2 * i], col = row_ent[2 * i + 1];
        int64_t j = i - 1;
        while (j >= b && row_ent[2 * j + 1] > col) { row_ent[2 * j + 2] = row_ent[2 * j]; row_ent[2 * j + 3] = row_ent[2 * j + 1]; --j; }
        row_ent[2 * j + 2] = id; row_ent[2 * j + 3] = col;
    }
}

// ---- values -------------------------------------------------------------------------------------------------------------
// diagonal blocks <- damped G (fill_matG, BA.cpp:1780-1823) with the gauge variables turned into identity rows; rhs <- -g_f
__global__ void k_bsr_fill_diag(int M, const double* __restrict__ G, const double* __restrict__ gf, double c, int unity, const int* __restrict__ diag_id,
                                double* __restrict__ blocks, double* __restrict__ rhs) {
    const int cam = blockIdx.x, t = threadIdx.x;
    if (t < 100) {
        const int a = t / 10, b = t % 10;
        const bool ka = red_index(cam, a, unity) >= 0, kb = red_index(cam, b, unity) >= 0;
        double v;
        if (ka && kb) { v = G[(size_t)cam * 100 + t]; if (a == b) v *= 1.0 + c; }
        else v = (a == b) ? 1.0 : 0.0;
        blocks[(size_t)diag_id[cam] * 100 + t] = v;
    } else if (t < 110) {
        const int a = t - 100;
        rhs[(size_t)cam * 10 + a] = red_index(cam, a, unity) >= 0 ? -gf[(size_t)cam * 10 + a] : 0.0;
    }
}
__global__ void k_gather_diag(int M, const int* __restrict__ diag_id, const double* __restrict__ blocks, double* __restrict__ diag) {
    const int cam = blockIdx.x, t = threadIdx.x;
    if (t < 100) diag[(size_t)cam * 100 + t] = blocks[(size_t)diag_id[cam] * 100 + t];
}
// block-Jacobi preconditioner: in-place inverse of the (symmetric positive definite) 10x10 diagonal blocks, Gauss-Jordan
__global__ void k_invert_diag(int M, double* __restrict__ diag) {
    const int cam = blockIdx.x * blockDim.x + threadIdx.x;
    if (cam >= M) return;
    double a[10][10], inv[10][10];
    double* d = diag + (size_t)cam * 100;
    for (int i = 0; i < 10; ++i) for (int j = 0; j < 10; ++j) { a[i][j] = 0.5 * (d[i * 10 + j] + d[j * 10 + i]); inv[i][j] = i == j ? 1.0 : 0.0; }
    for (int k = 0; k < 10; ++k) {
        const double piv = 1.0 / a[k][k];
        for (int j = 0; j < 10; ++j) { a[k][j] *= piv; inv[k][j] *= piv; }
        for (int i = 0; i < 10; ++i) {
            if (i == k) continue;
            const double f = a[i][k];
            for (int j = 0; j < 10; ++j) { a[i][j] -= f * a[k][j]; inv[i][j] -= f * inv[k][j]; }
        }
    }
    for (int i = 0; i < 10; ++i) for (int j = 0; j < 10; ++j) d[i * 10 + j] = inv[i][j];
}

// ---- PCG iteration on the whole GPU, deterministic --------------------------------------------------------------------------
// One iteration = three grid-wide kernels (the vector half used to be ONE CTA streaming 12 MB per iteration through one SM:
// 0.16 ms per iteration at configs[4]):
//   k_bsr_spmv_dot   y = S p  and per-CTA partials of p.y                       (k_dot_partials after the all-reduce of y when multi-GPU)
//   k_pcg_update     alpha = rz / sum(partials); x += alpha p; r -= alpha y; z = Minv r; per-CTA partials of r.z and r.r
//   k_pcg_direction  beta = sum(partials r.z) / rz; p = z + beta p; block 0 publishes the scalars of the next parity
// Every dot product is a two-stage sum with a fixed shape: shuffle tree inside a warp, warps in order, CTAs in order (every CTA
// re-reduces the partial array itself, so no extra launch and identical bits everywhere).
// scal[par*8 + {0: rz, 1: bb, 2: rr, 3: pAp}]: an iteration reads parity par and writes parity par^1.
constexpr int kPcgThreads = 256;   // full warps (the reductions shuffle with a full mask)
constexpr int kPcgPerCta = 250;    // entries per CTA of the vector kernels: 25 cameras of 10 variables, threads 250..255 idle

__device__ __forceinline__ double cta_sum_fixed(double v, double* red) {   // fixed shape: lanes (tree), warps in order
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    double t = 0.0;
    for (int i = 0; i < (int)((blockDim.x + 31) >> 5); ++i) t += red[i];
    return t;
}
// sum of partial[0..n) in a fixed order, identical in every CTA: thread t adds the entries t, t+T, ... then the CTA tree
__device__ __forceinline__ double reduce_partials(const double* __restrict__ partial, int n, int stride, int off, double* red) {
    double v = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) v += partial[(size_t)i * stride + off];
    return cta_sum_fixed(v, red);
}

// y = S v (10 threads per block row, 3 rows per warp, entries in column order) + partial of v.y per CTA
__global__ void __launch_bounds__(128) k_bsr_spmv_dot(int M, const int64_t* __restrict__ row_ptr, const int* __restrict__ row_ent, const double* __restrict__ blocks,
                                                      const double* __restrict__ v, double* __restrict__ y, double* __restrict__ partial) {
    __shared__ double red[4];
    const int lane = threadIdx.x & 31, warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int sub = lane / 10, a = lane % 10;
    const int cam = warp * 3 + sub;
    double dot = 0.0;
    if (sub < 3 && cam < M) {
        double acc = 0.0;
        for (int64_t e = row_ptr[cam]; e < row_ptr[cam + 1]; ++e) {
            const int ent = row_ent[2 * e], col = row_ent[2 * e + 1];
            const double* B = blocks + (size_t)(ent & 0x3fffffff) * 100;
            const double* vc = v + (size_t)col * 10;
            if (ent & (1 << 30)) {
#pragma unroll
                for (int b = 0; b < 10; ++b) acc += B[b * 10 + a] * vc[b];
            } else {
#pragma unroll
                for (int b = 0; b < 10; ++b) acc += B[a * 10 + b] * vc[b];
            }
        }
        y[(size_t)cam * 10 + a] = acc;
        dot = v[(size_t)cam * 10 + a] * acc;
    }
    if (partial != nullptr) {
        const double t = cta_sum_fixed(dot, red);
        if (threadIdx.x == 0) partial[blockIdx.x] = t;
    }
}
__global__ void __launch_bounds__(256) k_dot_partials(int n, const double* __restrict__ a, const double* __restrict__ b, double* __restrict__ partial) {
    __shared__ double red[8];
    double v = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) v += a[i] * b[i];
    const double t = cta_sum_fixed(v, red);
    if (threadIdx.x == 0) partial[blockIdx.x] = t;
}
// x = 0, r = b, z = Minv b, p = z; partials of r.z and b.b  (10 lanes per camera)
__global__ void __launch_bounds__(kPcgThreads) k_pcg_init(int M, const double* __restrict__ b, const double* __restrict__ Minv, double* __restrict__ x, double* __restrict__ r,
                                                          double* __restrict__ z, double* __restrict__ p, double* __restrict__ partial) {
    __shared__ double red[8];
    double rz = 0.0, bb = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < M * 10; i += gridDim.x * blockDim.x) {
        const int cam = i / 10, a = i - cam * 10;
        const double* Mi = Minv + (size_t)cam * 100 + a * 10;
        const double* bc = b + (size_t)cam * 10;
        double zi = 0.0;
#pragma unroll
        for (int k = 0; k < 10; ++k) zi += Mi[k] * bc[k];
        const double bi = bc[a];
        x[i] = 0.0; r[i] = bi; z[i] = zi; p[i] = zi;
        rz += bi * zi; bb += bi * bi;
    }
    rz = cta_sum_fixed(rz, red); bb = cta_sum_fixed(bb, red);
    if (threadIdx.x == 0) { partial[2 * blockIdx.x] = rz; partial[2 * blockIdx.x + 1] = bb; }
}
__global__ void __launch_bounds__(256) k_pcg_init_scal(int nparts, const double* __restrict__ partial, double* __restrict__ scal) {
    __shared__ double red[8];
    const double rz = reduce_partials(partial, nparts, 2, 0, red), bb = reduce_partials(partial, nparts, 2, 1, red);
    if (threadIdx.x == 0) { scal[0] = rz; scal[1] = bb; scal[2] = bb; scal[3] = 0.0; }
}
// alpha from the partials of p.y; x += alpha p; r -= alpha y; z = Minv r (the 10 lanes of a camera exchange r through shared memory);
// partials of r.z and r.r.  The grid covers M cameras: CTA c owns cameras [24 c, 24 c + 24).
__global__ void __launch_bounds__(kPcgThreads) k_pcg_update(int M, int n_dot_parts, const double* __restrict__ dot_partial, const double* __restrict__ Minv,
                                                            const double* __restrict__ y, const double* __restrict__ p, double* __restrict__ x, double* __restrict__ r,
                                                            double* __restrict__ z, const double* __restrict__ scal_in, double* __restrict__ partial) {
    __shared__ double red[8];
    __shared__ double rs[kPcgThreads];
    const double pAp = reduce_partials(dot_partial, n_dot_parts, 1, 0, red);
    const double rz = scal_in[0];
    const double alpha = (pAp != 0.0) ? rz / pAp : 0.0;
    const int i = blockIdx.x * kPcgPerCta + threadIdx.x;
    const int lc = threadIdx.x / 10, a = threadIdx.x - lc * 10;
    double ri = 0.0;
    const bool in = threadIdx.x < kPcgPerCta && i < M * 10;
    if (in) {
        x[i] += alpha * p[i];
        ri = r[i] - alpha * y[i];
        r[i] = ri;
    }
    rs[threadIdx.x] = ri;
    __syncthreads();
    double rzn = 0.0, rr = 0.0;
    if (in) {
        const int cam = i / 10;
        const double* Mi = Minv + (size_t)cam * 100 + a * 10;
        double zi = 0.0;
#pragma unroll
        for (int k = 0; k < 10; ++k) zi += Mi[k] * rs[lc * 10 + k];
        z[i] = zi;
        rzn = ri * zi; rr = ri * ri;
    }
    rzn = cta_sum_fixed(rzn, red); rr = cta_sum_fixed(rr, red);
    if (threadIdx.x == 0) { partial[3 * blockIdx.x] = rzn; partial[3 * blockIdx.x + 1] = rr; partial[3 * blockIdx.x + 2] = pAp; }
}
// beta from the partials of r.z; p = z + beta p; block 0 publishes the scalars of the next parity
__global__ void __launch_bounds__(kPcgThreads) k_pcg_direction(int M, int nparts, const double* __restrict__ partial, const double* __restrict__ z, double* __restrict__ p,
                                                               const double* __restrict__ scal_in, double* __restrict__ scal_out) {
    __shared__ double red[8];
    const double rzn = reduce_partials(partial, nparts, 3, 0, red);
    const double rz = scal_in[0];
    const double beta = (rz != 0.0) ? rzn / rz : 0.0;
    const int i = blockIdx.x * kPcgPerCta + threadIdx.x;
    if (threadIdx.x < kPcgPerCta && i < M * 10) p[i] = z[i] + beta * p[i];
    if (blockIdx.x == 0) {
        const double rr = reduce_partials(partial, nparts, 3, 1, red);
        if (threadIdx.x == 0) { scal_out[0] = rzn; scal_out[1] = scal_in[1]; scal_out[2] = rr; scal_out[3] = partial[2]; }
    }
}

// ---- the whole PCG loop as ONE persistent cooperative kernel (single GPU) ---------------------------------------------------------
// The launch-per-phase form above spends most of an iteration between kernels: three dependent grid-wide launches of ~10 us of work each
// (configs[4]: 57 us per iteration for a system that is L2 resident).  Here one co-resident grid walks the iterations itself with three
// grid barriers per iteration; every CTA re-reduces the per-CTA partials in the same fixed order, so all CTAs see bit-identical scalars,
// take the convergence decision in lock-step without a flag, and the iteration stays deterministic (for a given grid).  A warp owns
// three cameras at a time (10 lanes each) in every phase; z = Minv r exchanges r through shuffles.
// scal[0..3] = {rz, bb, rr, -} from k_pcg_init / k_pcg_init_scal; out[0] = iterations, out[1] = rr, out[2] = bb.
constexpr int kPcgPersistThreads = 256;
__global__ void __launch_bounds__(kPcgPersistThreads) k_pcg_persistent(int M, const int64_t* __restrict__ row_ptr, const int* __restrict__ row_ent,
                                                                       const double* __restrict__ blocks, const double* __restrict__ Minv, double* __restrict__ x,
                                                                       double* __restrict__ r, double* __restrict__ z, double* __restrict__ p, double* __restrict__ y,
                                                                       double* __restrict__ part_dot, double* __restrict__ part_vec, const double* __restrict__ scal,
                                                                       int max_iters, double tol2, double* __restrict__ out) {
    namespace cgx = cooperative_groups;
    cgx::grid_group grid = cgx::this_grid();
    __shared__ double red[8];
    const int lane = threadIdx.x & 31;
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, GW = (gridDim.x * blockDim.x) >> 5;
    const int sub = lane / 10, a = lane - 10 * sub;
    const int ntrip = (M + 2) / 3;
    const int G = (int)gridDim.x;
    double rz = scal[0];
    const double bb = scal[1];
    double rr = scal[2];
    int it = 0;
    if (!(bb > 0.0)) max_iters = 0;
    for (; it < max_iters; ++it) {
        // ---- y = S p, partial of p.y
        double dot = 0.0;
        for (int t = gw; t < ntrip; t += GW) {
            const int cam = 3 * t + sub;
            if (sub < 3 && cam < M) {
                double acc = 0.0;
                for (int64_t e = row_ptr[cam]; e < row_ptr[cam + 1]; ++e) {
                    const int ent = row_ent[2 * e], col = row_ent[2 * e + 1];
                    const double* B = blocks + (size_t)(ent & 0x3fffffff) * 100;
                    const double* vc = p + (size_t)col * 10;
                    if (ent & (1 << 30)) {
#pragma unroll
                        for (int b = 0; b < 10; ++b) acc += B[b * 10 + a] * vc[b];
                    } else {
#pragma unroll
                        for (int b = 0; b < 10; ++b) acc += B[a * 10 + b] * vc[b];
                    }
                }
                y[(size_t)cam * 10 + a] = acc;
                dot += p[(size_t)cam * 10 + a] * acc;
            }
        }
        dot = cta_sum_fixed(dot, red);
        if (threadIdx.x == 0) part_dot[blockIdx.x] = dot;
        grid.sync();
        // ---- alpha; x += alpha p; r -= alpha y; z = Minv r; partials of r.z and r.r
        const double pAp = reduce_partials(part_dot, G, 1, 0, red);
        const double alpha = (pAp != 0.0) ? rz / pAp : 0.0;
        double rzn = 0.0, rrn = 0.0;
        for (int t = gw; t < ntrip; t += GW) {
            const int cam = 3 * t + sub;
            const bool in = sub < 3 && cam < M;
            double ri = 0.0;
            if (in) {
                const size_t i = (size_t)cam * 10 + a;
                x[i] += alpha * p[i];
                ri = r[i] - alpha * y[i];
                r[i] = ri;
            }
            double zi = 0.0;
#pragma unroll
            for (int k = 0; k < 10; ++k) {
                const double rk = __shfl_sync(0xffffffffu, ri, (sub < 3 ? sub : 0) * 10 + k);
                if (in) zi += Minv[(size_t)cam * 100 + a * 10 + k] * rk;
            }
            if (in) { z[(size_t)cam * 10 + a] = zi; rzn += ri * zi; rrn += ri * ri; }
        }
        rzn = cta_sum_fixed(rzn, red); rrn = cta_sum_fixed(rrn, red);
        if (threadIdx.x == 0) { part_vec[2 * blockIdx.x] = rzn; part_vec[2 * blockIdx.x + 1] = rrn; }
        grid.sync();
        // ---- beta; p = z + beta p; convergence (every CTA holds the same scalars)
        const double rz_new = reduce_partials(part_vec, G, 2, 0, red);
        rr = reduce_partials(part_vec, G, 2, 1, red);
        const double beta = (rz != 0.0) ? rz_new / rz : 0.0;
        for (int t = gw; t < ntrip; t += GW) {
            const int cam = 3 * t + sub;
            if (sub < 3 && cam < M) { const size_t i = (size_t)cam * 10 + a; p[i] = z[i] + beta * p[i]; }
        }
        rz = rz_new;
        grid.sync();
        if (!(rr == rr) || rr <= tol2 * bb) { ++it; break; }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) { out[0] = (double)it; out[1] = rr; out[2] = bb; }
}

// parity hook: scatter the block-sparse system into the dense gauge-reduced layout
__global__ void k_bsr_to_dense(int nnzb, const int* __restrict__ blk_cams, const double* __restrict__ blocks, int unity, double* __restrict__ S, int64_t ld) {
    const int id = blockIdx.x, t = threadIdx.x;
    if (id >= nnzb || t >= 100) return;
    const int ci = blk_cams[2 * id], cl = blk_cams[2 * id + 1];
    const int a = t / 10, b = t % 10;
    const int row = red_index(ci, a, unity), col = red_index(cl, b, unity);
    if (row < 0 || col < 0) return;
    const double v = blocks[(size_t)id * 100 + t];
    S[(size_t)col * ld + row] = v;
    S[(size_t)row * ld + col] = v;
}
__global__ void k_full_to_reduced(int M, int unity, const double* __restrict__ full, double* __restrict__ red) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= M * 10) return;
    const int r = red_index(t / 10, t % 10, unity);
    if (r >= 0) red[r] = full[t];
}

}  // namespace

#define PCG_CUDA(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return SRK_E_CUDA; } while (0)

int pcg_build_structure(PcgWorkspace& ws, cudaStream_t st, int64_t N, int64_t O, int M, int tile_points, const int64_t* pt_begin, const int32_t* obs_cam,
                        unsigned char* deferred, int64_t* launches) {
    // capacity: a camera couples with at most a few dozen others in a localized scene; 64 slots per camera, power of two
    size_t cap = 1024;
    while (cap < (size_t)M * 64) cap <<= 1;
    if (cap > ((size_t)1 << 30)) return SRK_E_TOO_LARGE;
    PCG_CUDA(ensure(ws.hkeys, ws.hkeys_cap, cap)); PCG_CUDA(ensure(ws.hids, ws.hids_cap, cap));
    PCG_CUDA(ensure(ws.cnt, ws.cnt_cap, (size_t)M + 2)); PCG_CUDA(ensure(ws.row_ptr, ws.row_ptr_cap, (size_t)M + 1));
    PCG_CUDA(ensure(ws.diag_id, ws.diag_id_cap, (size_t)M));
    if (ws.misc == nullptr) PCG_CUDA(cudaMalloc((void**)&ws.misc, sizeof(int) * 4));
    if (ws.scal == nullptr) PCG_CUDA(cudaMalloc((void**)&ws.scal, sizeof(double) * 16));
    if (ws.h_scal == nullptr) PCG_CUDA(cudaMallocHost((void**)&ws.h_scal, sizeof(double) * 8));
    ws.hmask = (unsigned)(cap - 1);
    PCG_CUDA(cudaMemsetAsync(ws.hkeys, 0xff, sizeof(unsigned long long) * cap, st));
    PCG_CUDA(cudaMemsetAsync(ws.misc, 0, sizeof(int) * 4, st));
    PCG_CUDA(cudaMemsetAsync(ws.cnt, 0, sizeof(unsigned long long) * ((size_t)M + 2), st));
    k_insert_diagonals<<<cdiv(M, 256), 256, 0, st>>>(M, ws.hkeys, ws.hmask, ws.misc);
    SchurSink none{};
    launch_schur_tile(st, N, O, tile_points, pt_begin, obs_cam, nullptr, 0.0, none, nullptr, nullptr, deferred, 1, ws.hkeys, ws.hmask, ws.misc);
    k_count_blocks<<<cdiv((int64_t)cap, 256), 256, 0, st>>>((unsigned)cap, ws.hkeys, ws.misc + 1);
    *launches += 3;
    int h[2] = {0, 0};
    PCG_CUDA(cudaMemcpyAsync(h, ws.misc, sizeof(int) * 2, cudaMemcpyDeviceToHost, st));
    PCG_CUDA(cudaStreamSynchronize(st));
    if (h[0] != 0 || (size_t)h[1] * 10 > cap * 7) return SRK_E_TOO_LARGE;   // table overflow / load factor above 0.7
    ws.nnzb = h[1];
    PCG_CUDA(ensure(ws.blk_cams, ws.blk_cams_cap, (size_t)2 * ws.nnzb));
    PCG_CUDA(ensure(ws.row_ent, ws.row_ent_cap, (size_t)4 * ws.nnzb));
    PCG_CUDA(ensure(ws.blocks, ws.blocks_cap, (size_t)100 * ws.nnzb));
    k_assign_ids<<<1, 1024, 0, st>>>((unsigned)cap, ws.hkeys, ws.hids, ws.blk_cams, ws.cnt);
    launch_scan_counts(st, M, ws.cnt, ws.row_ptr, ws.cnt);   // cnt becomes the per-row cursor
    k_fill_rows<<<cdiv(ws.nnzb, 256), 256, 0, st>>>(ws.nnzb, ws.blk_cams, ws.cnt, ws.row_ent, ws.diag_id);
    k_sort_rows<<<cdiv(M, 128), 128, 0, st>>>(M, ws.row_ptr, ws.row_ent);
    *launches += 4;
    size_t nv = (size_t)M * 10;
    if (ws.rhs == nullptr || ws.vec_cap < nv) {
        if (ws.rhs != nullptr) cudaFree(ws.rhs);
        ws.rhs = nullptr; ws.vec_cap = 0;
        PCG_CUDA(cudaMalloc((void**)&ws.rhs, sizeof(double) * nv * 5));
        ws.vec_cap = nv;
    }
    ws.r = ws.rhs + nv; ws.z = ws.r + nv; ws.p = ws.z + nv; ws.y = ws.p + nv;
    PCG_CUDA(ensure(ws.diag, ws.diag_cap, (size_t)100 * M));
    PCG_CUDA(cudaStreamSynchronize(st));
    PCG_CUDA(cudaGetLastError());
    ws.structure_valid = true;
    return SRK_OK;
}

int pcg_begin(PcgWorkspace& ws, cudaStream_t st, int M, const double* G, const double* gf, double c, int unity, int rank, SchurSink* sink, int64_t* launches) {
    if (!ws.structure_valid) return SRK_E_NOT_BOUND;
    PCG_CUDA(cudaMemsetAsync(ws.blocks, 0, sizeof(double) * 100 * (size_t)ws.nnzb, st));
    PCG_CUDA(cudaMemsetAsync(ws.rhs, 0, sizeof(double) * 10 * (size_t)M, st));
    if (rank == 0) { k_bsr_fill_diag<<<M, 128, 0, st>>>(M, G, gf, c, unity, ws.diag_id, ws.blocks, ws.rhs); *launches += 1; }
    sink->S = nullptr; sink->ld = 0; sink->rhs = ws.rhs; sink->unity = unity;
    sink->hkeys = ws.hkeys; sink->hids = ws.hids; sink->hmask = ws.hmask; sink->blocks = ws.blocks;
    return SRK_OK;
}

int pcg_solve(PcgWorkspace& ws, cudaStream_t st, int M, double* x, int max_iters, double rel_tol, int world, srk_allreduce_fn ar, void* ar_user,
              int64_t* launches, int32_t* iters_out, double* rel_res_out) {
    if (!ws.structure_valid) return SRK_E_NOT_BOUND;
    const int n = M * 10;
    if (max_iters <= 0) max_iters = 4 * n < 20000 ? 4 * n : 20000;
    if (rel_tol <= 0.0) rel_tol = 1e-13;
    const bool multi = ar != nullptr && world > 1;
    if (multi && ar(ar_user, ws.rhs, n, (void*)st) != 0) return SRK_E_CUDA;
    k_gather_diag<<<M, 128, 0, st>>>(M, ws.diag_id, ws.blocks, ws.diag);
    if (multi && ar(ar_user, ws.diag, (int64_t)100 * M, (void*)st) != 0) return SRK_E_CUDA;
    k_invert_diag<<<cdiv(M, 64), 64, 0, st>>>(M, ws.diag);
    const int g_spmv = (int)cdiv((int64_t)((M + 2) / 3) * 32, 128);       // CTAs of the mat-vec = partials of p.y (single GPU)
    const int g_vec = (int)cdiv((int64_t)n, kPcgPerCta);                  // CTAs of the vector kernels (25 cameras each)
    const int g_dot = 296;                                                 // CTAs of the separate dot product (multi-GPU)
    const int g_init = g_vec < 1184 ? g_vec : 1184;
    {
        const size_t need = (size_t)(g_spmv > g_dot ? g_spmv : g_dot) + 3 * (size_t)g_vec + 2 * (size_t)g_init + 16 + 3 * 1024 + 8;   // + the persistent kernel's per-CTA partials
        PCG_CUDA(ensure(ws.partials, ws.partials_cap, need));
    }
    double* part_dot = ws.partials;
    double* part_vec = ws.partials + (g_spmv > g_dot ? g_spmv : g_dot);
    double* part_init = part_vec + 3 * (size_t)g_vec;
    k_pcg_init<<<g_init, kPcgThreads, 0, st>>>(M, ws.rhs, ws.diag, x, ws.r, ws.z, ws.p, part_init);
    k_pcg_init_scal<<<1, 256, 0, st>>>(g_init, part_init, ws.scal);
    *launches += 4;
    const int check_every = 16;
    int it = 0;
    double rel = 1.0;
    const double tol2 = rel_tol * rel_tol;
    static int persist = -1;   // SRK_PCG_PERSISTENT=0: the launch-per-phase loop on one GPU too (cross-check)
    if (persist < 0) { const char* e = getenv("SRK_PCG_PERSISTENT"); persist = (e != nullptr && e[0] == '0') ? 0 : 1; }
    if (!multi && persist) {
        int dev = 0, sms = 148, occ = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_pcg_persistent, kPcgPersistThreads, 0) == cudaSuccess && occ >= 1) {
            const int per_sm = occ < 4 ? occ : 4;
            int G = sms * per_sm;
            if (G > 1024) G = 1024;
            double* pd = ws.partials + (g_spmv > g_dot ? g_spmv : g_dot) + 3 * (size_t)g_vec + 2 * (size_t)g_init + 16;
            double* pv = pd + G;
            double* outp = ws.scal + 8;      // parity-1 slot of the scalar block: free on this path
            int Mi = M, mi = max_iters; double t2 = tol2;
            const int64_t* rp = ws.row_ptr; const int* re = ws.row_ent; const double* bl = ws.blocks; const double* mv = ws.diag; const double* sc = ws.scal;
            double *xx = x, *rr_ = ws.r, *zz = ws.z, *pp = ws.p, *yy = ws.y;
            void* args[] = {&Mi, &rp, &re, &bl, &mv, &xx, &rr_, &zz, &pp, &yy, &pd, &pv, &sc, &mi, &t2, &outp};
            if (cudaLaunchCooperativeKernel((void*)k_pcg_persistent, dim3(G), dim3(kPcgPersistThreads), args, 0, st) == cudaSuccess) {
                *launches += 1;
                PCG_CUDA(cudaMemcpyAsync(ws.h_scal, outp, sizeof(double) * 3, cudaMemcpyDeviceToHost, st));
                PCG_CUDA(cudaStreamSynchronize(st));
                it = (int)ws.h_scal[0];
                const double rrv = ws.h_scal[1], bbv = ws.h_scal[2];
                rel = bbv > 0.0 ? sqrt(rrv / bbv) : 0.0;
                *iters_out = it;
                if (rel_res_out != nullptr) *rel_res_out = rel;
                PCG_CUDA(cudaGetLastError());
                return SRK_OK;
            }
            cudaGetLastError();   // cooperative launch refused: the loop below
        }
    }
    while (it < max_iters) {
        const int batch = (max_iters - it) < check_every ? (max_iters - it) : check_every;
        for (int k = 0; k < batch; ++k) {
            const int par = (it + k) & 1;
            double* s_in = ws.scal + 8 * par; double* s_out = ws.scal + 8 * (par ^ 1);
            int n_dot = g_spmv;
            if (!multi) {
                k_bsr_spmv_dot<<<g_spmv, 128, 0, st>>>(M, ws.row_ptr, ws.row_ent, ws.blocks, ws.p, ws.y, part_dot);
            } else {
                k_bsr_spmv_dot<<<g_spmv, 128, 0, st>>>(M, ws.row_ptr, ws.row_ent, ws.blocks, ws.p, ws.y, nullptr);
                if (ar(ar_user, ws.y, n, (void*)st) != 0) return SRK_E_CUDA;
                k_dot_partials<<<g_dot, 256, 0, st>>>(n, ws.p, ws.y, part_dot);
                n_dot = g_dot; *launches += 1;
            }
            k_pcg_update<<<g_vec, kPcgThreads, 0, st>>>(M, n_dot, part_dot, ws.diag, ws.y, ws.p, x, ws.r, ws.z, s_in, part_vec);
            k_pcg_direction<<<g_vec, kPcgThreads, 0, st>>>(M, g_vec, part_vec, ws.z, ws.p, s_in, s_out);
        }
        *launches += 3 * batch;
        it += batch;
        PCG_CUDA(cudaMemcpyAsync(ws.h_scal, ws.scal + 8 * (it & 1), sizeof(double) * 4, cudaMemcpyDeviceToHost, st));
        PCG_CUDA(cudaStreamSynchronize(st));
        const double bb = ws.h_scal[1], rr = ws.h_scal[2];
        if (!(bb > 0.0)) { rel = 0.0; break; }
        rel = sqrt(rr / bb);
        if (!(rr == rr) || rr <= tol2 * bb) break;   // NaN (breakdown) or converged
    }
    *iters_out = it;
    if (rel_res_out != nullptr) *rel_res_out = rel;
    PCG_CUDA(cudaGetLastError());
    return SRK_OK;
}

int pcg_debug_to_dense(PcgWorkspace& ws, cudaStream_t st, int M, int unity, double* S, int64_t ld, double* rhs, int64_t* launches) {
    if (!ws.structure_valid) return SRK_E_NOT_BOUND;
    const int nf = M * 10 - 7;
    PCG_CUDA(cudaMemsetAsync(S, 0, sizeof(double) * (size_t)ld * nf, st));
    k_bsr_to_dense<<<ws.nnzb, 128, 0, st>>>(ws.nnzb, ws.blk_cams, ws.blocks, unity, S, ld);
    k_full_to_reduced<<<cdiv((int64_t)M * 10, 256), 256, 0, st>>>(M, unity, ws.rhs, rhs);
    *launches += 2;
    return SRK_OK;
}

void pcg_release(PcgWorkspace& ws) {
    void* ptrs[] = {ws.hkeys, ws.hids, ws.blk_cams, ws.row_ptr, ws.row_ent, ws.diag_id, ws.cnt, ws.misc, ws.blocks, ws.rhs, ws.diag, ws.scal, ws.partials};
    for (void* p : ptrs) if (p != nullptr) cudaFree(p);
    if (ws.h_scal != nullptr) cudaFreeHost(ws.h_scal);
    ws = PcgWorkspace();
}

}  // namespace srk
