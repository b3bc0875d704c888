// suriko-b200 — K2 on the FP64 tensor pipe: per-point 3x3 blocks + Schur-complement accumulation as a tiled DMMA contraction.
//
// Replaces the Schur loop of EstimateCorrectionsDecomposedInTwoPhases (BA.cpp:1859-1900):
//     S[cam_i, cam_l] -= sum_j F_ji^T E_cj^-1 F_jl ,   rhs[cam_i] += sum_j F_ji^T E_cj^-1 g_pj
// A CTA owns a tile of `tile_points` consecutive points.  Tracks are created in capture order, so a tile touches few cameras;
// the CTA builds the sorted table of the (at most kMmaCams) distinct cameras of its tile.  With row index m = slot*10 + a
// (slot = position of the camera in the table, a = frame variable) the tile's whole contribution is ONE dense contraction
//     D[m][n] = sum_{(j,v)} Fall[(j,v)][m] * Wall[(j,v)][n],     Fall_j = F_j scattered to the slots point j sees (zero elsewhere),
//                                                                Wall_j = E_cj^-1 Fall_j,     k = (j,v) runs over 3 rows per point,
// i.e. a [120 x 3P] x [3P x 120] product whose lower block triangle is what the reference subtracts.  Zero-filling the
// unseen slots (a point sees ~10 of the ~12 cameras of its tile) costs ~35 % extra multiply-adds and buys a regular
// mma.sync.m8n8k4.f64 pipeline with every operand read from shared memory exactly once per 8x8 output tile.
//
// Warp-specialised, 512 threads, one CTA per SM, batches of 16 points, two operand buffers:
//   producer warps 8..15 : half-warp per point.  Lanes over the point's observations reduce E = 2 sum Jp^T Jp and g_p with shuffles,
//                          apply the damping and the cofactor inverse with the |det| > 1e-12 rule (quirk Q5), store E^-1 / g_p / flags
//                          for K2', then lanes over the table slots build F_i = 2 Jp_i^T Jc_i and W_i = E^-1 F_i (or zeros for a
//                          slot the point does not see) and store them into the k-major operand buffers of batch b+1
//   consumer warps 0..7  : DMMA over batch b.  The 120x120 lower triangle is cut into 15 super-blocks of 24x24 (3x3 fragments);
//                          warp w accumulates super-blocks w and w+8 in registers across ALL batches of the tile; warp 7 owns one
//                          super-block and the tile's rhs entries (a GEMV against t_j = E_cj^-1 g_pj).
//   one __syncthreads per batch; at the end of the tile every accumulator is flushed with one red.global.add.f64 per touched
//   entry (per TILE, not per point).
// Points with more than 16 observations or with a camera outside the tile table are flagged in `deferred` and handled by
// k_schur (per-point atomics), exactly like the bind-time plan (k_schur_tile, plan_only) predicts.
#include "common.cuh"
#include "kernels.h"

namespace srk {

constexpr int kMmaCams = 12;              // cameras per tile table (must match the plan kernel's CMAX)
constexpr int kMmaRows = kMmaCams * 10;   // 120
constexpr int kMmaSLD = 124;              // row stride of the operand buffers: 124 = 12 (mod 16) -> conflict-free fragment loads
constexpr int kMmaBP = 16;                // points per batch
constexpr int kMmaK = 3 * kMmaBP;         // 48 contraction rows per batch
constexpr int kMmaThreads = 512;
constexpr int kMmaHash = 64;

struct MmaSmem {
    double F[2][kMmaK * kMmaSLD];
    double W[2][kMmaK * kMmaSLD];
    double T[2][kMmaK];
    int tab[kMmaCams];
    int hash[kMmaHash];
    int slot_obs[kMmaBP][16];
    int blk[kMmaCams * kMmaCams];
    int n_local;
};

__device__ __forceinline__ void cta_barrier() { asm volatile("bar.sync 0;\n" ::: "memory"); }
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(kMmaThreads, 1) k_schur_mma(int64_t N, int64_t O, int tile_points, const int64_t* __restrict__ pt_begin,
                                                              const int32_t* __restrict__ obs_cam, const double* __restrict__ J, double c, SchurSink sink,
                                                              double* __restrict__ pinv, unsigned char* __restrict__ skipped,
                                                              unsigned char* __restrict__ deferred) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    MmaSmem& sm = *reinterpret_cast<MmaSmem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int64_t p0 = (int64_t)blockIdx.x * tile_points;
    const int64_t p1 = min(N, p0 + (int64_t)tile_points);
    if (p0 >= N) return;

    // ---- camera table of the tile (identical to the plan kernel: hash-set insert of every observation's camera, then sort)
    for (int i = tid; i < kMmaHash; i += kMmaThreads) sm.hash[i] = -1;
    __syncthreads();
    {
        const int64_t ob = pt_begin[p0], oe = pt_begin[p1];
        for (int64_t o = ob + tid; o < oe; o += kMmaThreads) {
            const int cam = obs_cam[o];
            unsigned h = ((unsigned)cam * 2654435761u) >> 26;
            for (int probe = 0; probe < kMmaHash; ++probe) {
                const int prev = atomicCAS(&sm.hash[h], -1, cam);
                if (prev == -1 || prev == cam) break;
                h = (h + 1) & (kMmaHash - 1);
            }
        }
    }
    __syncthreads();
    if (tid == 0) {
        int n = 0;
        for (int i = 0; i < kMmaHash; ++i) {
            const int cam = sm.hash[i];
            if (cam < 0) continue;
            int pos = n < kMmaCams ? n : kMmaCams;
            while (pos > 0 && sm.tab[pos - 1] > cam) --pos;
            if (pos >= kMmaCams) continue;
            const int last = n < kMmaCams ? n : kMmaCams - 1;
            for (int q = last; q > pos; --q) sm.tab[q] = sm.tab[q - 1];
            sm.tab[pos] = cam;
            if (n < kMmaCams) ++n;
        }
        sm.n_local = n;
    }
    __syncthreads();
    const int nLocal = sm.n_local;
    if (sink.blocks != nullptr) {   // block ids of the tile's camera pairs (block-sparse sink), looked up once per tile
        for (int e = tid; e < kMmaCams * kMmaCams; e += kMmaThreads) {
            const int si = e / kMmaCams, sl = e % kMmaCams;
            sm.blk[e] = (si < nLocal && sl <= si) ? sink_block_id(sink, sm.tab[si], sm.tab[sl]) : -1;
        }
    }

    const int nbatch = (int)((p1 - p0 + kMmaBP - 1) / kMmaBP);
    const bool producer = w >= 8;

    // ---- producer: stage one batch (16 points, half-warp per point) into buffer `buf`
    auto stage = [&](int b, int buf) {
        const int h = lane >> 4, hl = lane & 15;
        const int pl = 2 * (w - 8) + h;                    // point of the batch
        const int64_t j = p0 + (int64_t)b * kMmaBP + pl;
        const bool valid = j < p1;
        int64_t kb = 0; int k = 0;
        if (valid) { kb = pt_begin[j]; k = (int)(pt_begin[j + 1] - kb); }
        bool bad = k > 16;
        int loc = -1;
        if (valid && !bad && hl < k) {
            const int cam = obs_cam[kb + hl];
            for (int q = 0; q < nLocal; ++q) if (sm.tab[q] == cam) loc = q;
            if (loc < 0) bad = true;
        }
        const unsigned bal = __ballot_sync(0xffffffffu, bad);
        bad = ((bal >> (16 * h)) & 0xffffu) != 0;
        if (valid && hl == 0) deferred[j] = bad ? 1 : 0;
        const bool use = valid && !bad;
        double a9[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) a9[i] = 0.0;
        if (use && hl < k) {
            const int64_t o = kb + hl;
            const double rx = J[o], ry = J[O + o];
            double jp[6];
#pragma unroll
            for (int i = 0; i < 6; ++i) jp[i] = J[(int64_t)(2 + i) * O + o];
            a9[0] = jp[0] * jp[0] + jp[1] * jp[1];
            a9[1] = jp[0] * jp[2] + jp[1] * jp[3];
            a9[2] = jp[0] * jp[4] + jp[1] * jp[5];
            a9[3] = jp[2] * jp[2] + jp[3] * jp[3];
            a9[4] = jp[2] * jp[4] + jp[3] * jp[5];
            a9[5] = jp[4] * jp[4] + jp[5] * jp[5];
            a9[6] = jp[0] * rx + jp[1] * ry;
            a9[7] = jp[2] * rx + jp[3] * ry;
            a9[8] = jp[4] * rx + jp[5] * ry;
        }
#pragma unroll
        for (int i = 0; i < 9; ++i) {
            double v = a9[i];
#pragma unroll
            for (int s = 8; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);   // stays inside the half-warp
            a9[i] = 2.0 * v;
        }
        double inv[6];
        const bool ok = point_block_inverse(a9, c, inv);
        if (use && hl == 0) {
            skipped[j] = ok ? 0 : 1;
#pragma unroll
            for (int i = 0; i < 6; ++i) pinv[(int64_t)i * N + j] = ok ? inv[i] : 0.0;
#pragma unroll
            for (int i = 0; i < 3; ++i) pinv[(int64_t)(6 + i) * N + j] = a9[6 + i];
        }
        const bool contrib = use && ok;   // BA.cpp:1877-1881: a non-invertible point block contributes nothing
        if (hl < 3) {
            double t = 0.0;
            if (contrib) {
                if (hl == 0) t = inv[0] * a9[6] + inv[1] * a9[7] + inv[2] * a9[8];
                else if (hl == 1) t = inv[1] * a9[6] + inv[3] * a9[7] + inv[4] * a9[8];
                else t = inv[2] * a9[6] + inv[4] * a9[7] + inv[5] * a9[8];
            }
            sm.T[buf][3 * pl + hl] = t;
        }
        // slot -> observation of this point
        sm.slot_obs[pl][hl] = -1;
        __syncwarp();
        if (contrib && hl < k) sm.slot_obs[pl][loc] = hl;
        __syncwarp();
        if (hl < kMmaCams) {
            const int oi = sm.slot_obs[pl][hl];
            double* Fr = sm.F[buf] + (3 * pl) * kMmaSLD + 10 * hl;
            double* Wr = sm.W[buf] + (3 * pl) * kMmaSLD + 10 * hl;
            if (oi >= 0) {
                const int64_t o = kb + oi;
                double jp[6];
#pragma unroll
                for (int i = 0; i < 6; ++i) jp[i] = J[(int64_t)(2 + i) * O + o];
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    double jc[10];
#pragma unroll
                    for (int i = 0; i < 10; ++i) jc[i] = J[(int64_t)(8 + 10 * half + i) * O + o];
#pragma unroll
                    for (int a2 = 0; a2 < 5; a2 += 2) {
                        // variables a = 5*half + a2 (and a + 1 when it stays inside this half)
                        const int a = 5 * half + a2;
                        const int na = (a2 + 1 < 5) ? 2 : 1;
                        double f[3][2], ww[3][2];
#pragma unroll
                        for (int u = 0; u < 2; ++u) {
                            const double j0 = u < na ? jc[(a2 + u) * 2] : 0.0, j1 = u < na ? jc[(a2 + u) * 2 + 1] : 0.0;
                            f[0][u] = 2.0 * (jp[0] * j0 + jp[1] * j1);
                            f[1][u] = 2.0 * (jp[2] * j0 + jp[3] * j1);
                            f[2][u] = 2.0 * (jp[4] * j0 + jp[5] * j1);
                            ww[0][u] = inv[0] * f[0][u] + inv[1] * f[1][u] + inv[2] * f[2][u];
                            ww[1][u] = inv[1] * f[0][u] + inv[3] * f[1][u] + inv[4] * f[2][u];
                            ww[2][u] = inv[2] * f[0][u] + inv[4] * f[1][u] + inv[5] * f[2][u];
                        }
#pragma unroll
                        for (int v = 0; v < 3; ++v)
#pragma unroll
                            for (int u = 0; u < 2; ++u)
                                if (u < na) { Fr[v * kMmaSLD + a + u] = f[v][u]; Wr[v * kMmaSLD + a + u] = ww[v][u]; }
                    }
                }
            } else {
#pragma unroll
                for (int v = 0; v < 3; ++v)
#pragma unroll
                    for (int a = 0; a < 10; a += 2) {
                        *reinterpret_cast<double2*>(Fr + v * kMmaSLD + a) = make_double2(0.0, 0.0);
                        *reinterpret_cast<double2*>(Wr + v * kMmaSLD + a) = make_double2(0.0, 0.0);
                    }
            }
        }
    };

    // Producers and consumers run separate loops with the same number of CTA-wide barriers (bar.sync 0 counts arrivals, it
    // does not care from which instruction they come), so the consumers' accumulators are not live in the producer code.
    if (producer) {
        stage(0, 0);
        cta_barrier();
        for (int b = 0; b < nbatch; ++b) {
            if (b + 1 < nbatch) stage(b + 1, (b & 1) ^ 1);
            cta_barrier();
        }
        return;
    }

    // ---- consumers: super-block sb = I(I+1)/2 + Jc covers rows [24 I, 24 I + 24) x columns [24 Jc, 24 Jc + 24)
    const int g = lane >> 2, tg = lane & 3;
    int sbI[2] = {0, 0}, sbJ[2] = {0, 0};
    int nsb = 0;
    for (int q = 0; q < 2; ++q) {
        const int sb = w + 8 * q;
        if (sb >= 15) break;
        int I = 0;
        while ((I + 1) * (I + 2) / 2 <= sb) ++I;
        sbI[q] = I; sbJ[q] = sb - I * (I + 1) / 2;
        nsb = q + 1;
    }
    double acc[2][3][3][2];
#pragma unroll
    for (int q = 0; q < 2; ++q)
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int jj = 0; jj < 3; ++jj) { acc[q][i][jj][0] = 0.0; acc[q][i][jj][1] = 0.0; }
    double racc[4] = {0.0, 0.0, 0.0, 0.0};   // warp 7: rhs entries m = lane + 32*q < 120

    cta_barrier();
    for (int b = 0; b < nbatch; ++b) {
        const int buf = b & 1;
        const double* F = sm.F[buf];
        const double* W = sm.W[buf];
        const double* pa0 = F + tg * kMmaSLD + 24 * sbI[0] + g;
        const double* pb0 = W + tg * kMmaSLD + 24 * sbJ[0] + g;
        const double* pa1 = F + tg * kMmaSLD + 24 * sbI[1] + g;
        const double* pb1 = W + tg * kMmaSLD + 24 * sbJ[1] + g;
#pragma unroll 4
        for (int ks = 0; ks < kMmaK / 4; ++ks) {
            const int off = ks * 4 * kMmaSLD;
            double af[3], bf[3];
#pragma unroll
            for (int i = 0; i < 3; ++i) { af[i] = pa0[off + 8 * i]; bf[i] = pb0[off + 8 * i]; }
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int jj = 0; jj < 3; ++jj) dmma884(acc[0][i][jj][0], acc[0][i][jj][1], af[i], bf[jj]);
            if (nsb > 1) {
#pragma unroll
                for (int i = 0; i < 3; ++i) { af[i] = pa1[off + 8 * i]; bf[i] = pb1[off + 8 * i]; }
#pragma unroll
                for (int i = 0; i < 3; ++i)
#pragma unroll
                    for (int jj = 0; jj < 3; ++jj) dmma884(acc[1][i][jj][0], acc[1][i][jj][1], af[i], bf[jj]);
            }
        }
        if (w == 7) {   // rhs += Fall^T t
            const double* T = sm.T[buf];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int m = lane + 32 * q;
                if (m < kMmaRows) {
                    double s = racc[q];
#pragma unroll 8
                    for (int kk = 0; kk < kMmaK; ++kk) s += F[kk * kMmaSLD + m] * T[kk];
                    racc[q] = s;
                }
            }
        }
        cta_barrier();
    }

    // ---- flush: one red.global.add.f64 per touched entry per tile
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        if (q >= nsb) break;
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int jj = 0; jj < 3; ++jj)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const double v = acc[q][i][jj][e];
                    if (v == 0.0) continue;
                    const int row = 24 * sbI[q] + 8 * i + g, col = 24 * sbJ[q] + 8 * jj + 2 * tg + e;
                    const int si = row / 10, a = row - 10 * si, sl = col / 10, bq = col - 10 * sl;
                    if (si >= nLocal || sl > si) continue;
                    const int cam_i = sm.tab[si], cam_l = sm.tab[sl];
                    const int blk = sink.blocks != nullptr ? sm.blk[si * kMmaCams + sl] : -1;
                    if (si != sl) {
                        sink_add(sink, blk, cam_i, a, cam_l, bq, -v);
                    } else if (a >= bq) {          // diagonal block: the lower triangle is computed once and mirrored inside the block
                        sink_add(sink, blk, cam_i, a, cam_l, bq, -v);
                        if (a != bq) sink_add(sink, blk, cam_i, bq, cam_l, a, -v);
                    }
                }
    }
    if (w == 7) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int m = lane + 32 * q;
            if (m >= kMmaRows) continue;
            const int s = m / 10, a = m - 10 * s;
            if (s < nLocal && racc[q] != 0.0) sink_add_rhs(sink, sm.tab[s], a, racc[q]);
        }
    }
}

void launch_schur_mma(cudaStream_t st, int64_t N, int64_t O, int tile_points, const int64_t* pt_begin, const int32_t* obs_cam, const double* J, double c,
                      const SchurSink& sink, double* pinv, unsigned char* skipped, unsigned char* deferred) {
    if (N <= 0) return;
    static bool attr = false;
    if (!attr) { cudaFuncSetAttribute(k_schur_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(MmaSmem)); attr = true; }
    const unsigned grid = (unsigned)((N + tile_points - 1) / tile_points);
    k_schur_mma<<<grid, kMmaThreads, sizeof(MmaSmem), st>>>(N, O, tile_points, pt_begin, obs_cam, J, c, sink, pinv, skipped, deferred);
}

}  // namespace srk
