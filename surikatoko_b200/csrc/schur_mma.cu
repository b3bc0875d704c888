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
    int gidx[kMmaRows];    // local row (slot*10 + a) -> gauge-reduced global index, -1 for a removed variable or an empty slot
    int n_local;
};

__device__ __forceinline__ void cta_barrier() { asm volatile("bar.sync 0;\n" ::: "memory"); }
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// One super-block (NI x NJ fragments of 8x8; DIAG: only the fragments on and below its diagonal) over one batch: straight-line DMMA
// code for every shape, chosen per tile -- a predicate per fragment inside the loop was measured slower than computing the zeros.
template <int NI, int NJ, bool DIAG>
__device__ __forceinline__ void sb_batch(double (&acc)[3][3][2], const double* __restrict__ pa, const double* __restrict__ pb) {
#pragma unroll 4
    for (int ks = 0; ks < kMmaK / 4; ++ks) {
        const int off = ks * 4 * kMmaSLD;
        double af[NI], bf[NJ];
#pragma unroll
        for (int i = 0; i < NI; ++i) af[i] = pa[off + 8 * i];
#pragma unroll
        for (int j = 0; j < NJ; ++j) bf[j] = pb[off + 8 * j];
#pragma unroll
        for (int i = 0; i < NI; ++i)
#pragma unroll
            for (int jj = 0; jj < NJ; ++jj)
                if (!DIAG || jj <= i) dmma884(acc[i][jj][0], acc[i][jj][1], af[i], bf[jj]);
    }
}
// shape code: off-diagonal (ni - 1) * 3 + (nj - 1) in 0..8, diagonal 9 + (ni - 1) in 9..11, -1 = nothing to do
__device__ __forceinline__ void sb_dispatch(int code, double (&acc)[3][3][2], const double* pa, const double* pb) {
    switch (code) {
        case 0: sb_batch<1, 1, false>(acc, pa, pb); break;
        case 1: sb_batch<1, 2, false>(acc, pa, pb); break;
        case 2: sb_batch<1, 3, false>(acc, pa, pb); break;
        case 3: sb_batch<2, 1, false>(acc, pa, pb); break;
        case 4: sb_batch<2, 2, false>(acc, pa, pb); break;
        case 5: sb_batch<2, 3, false>(acc, pa, pb); break;
        case 6: sb_batch<3, 1, false>(acc, pa, pb); break;
        case 7: sb_batch<3, 2, false>(acc, pa, pb); break;
        case 8: sb_batch<3, 3, false>(acc, pa, pb); break;
        case 9: sb_batch<1, 1, true>(acc, pa, pb); break;
        case 10: sb_batch<2, 2, true>(acc, pa, pb); break;
        case 11: sb_batch<3, 3, true>(acc, pa, pb); break;
        default: break;
    }
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
    if (tid < kMmaHash) {   // rank of every distinct camera = number of smaller ones; the table keeps the kMmaCams smallest ids, sorted
        const int cam = sm.hash[tid];
        int rank = 0, total = 0;
        for (int i = 0; i < kMmaHash; ++i) {
            const int other = sm.hash[i];
            total += other >= 0;
            rank += (other >= 0 && other < cam);
        }
        if (cam >= 0 && rank < kMmaCams) sm.tab[rank] = cam;
        if (tid == 0) sm.n_local = total < kMmaCams ? total : kMmaCams;
    }
    __syncthreads();
    const int nLocal = sm.n_local;
    if (sink.blocks != nullptr) {   // block ids of the tile's camera pairs (block-sparse sink), looked up once per tile
        for (int e = tid; e < kMmaCams * kMmaCams; e += kMmaThreads) {
            const int si = e / kMmaCams, sl = e % kMmaCams;
            sm.blk[e] = (si < nLocal && sl <= si) ? sink_block_id(sink, sm.tab[si], sm.tab[sl]) : -1;
        }
    }

    for (int m = tid; m < kMmaRows; m += kMmaThreads) {
        const int slot = m / 10, a = m - 10 * slot;
        sm.gidx[m] = slot < nLocal ? red_index(sm.tab[slot], a, sink.unity) : -1;
    }

    const int nbatch = (int)((p1 - p0 + kMmaBP - 1) / kMmaBP);
    const bool producer = w >= 8;

    // ---- producer: stage one batch (16 points, half-warp per point) into buffer `buf`.  (kb, k) = first observation and track
    // length of this half-warp's point, loaded one batch ahead; every global load of the batch is issued before the first use.
    const int h = lane >> 4, hl = lane & 15;
    auto load_track = [&](int b, int64_t& kb, int& k) {
        const int64_t j = p0 + (int64_t)b * kMmaBP + 2 * (w - 8) + h;
        kb = 0; k = 0;
        if (b < nbatch && j < p1) { kb = pt_begin[j]; k = (int)(pt_begin[j + 1] - kb); }
    };
    auto load_cam = [&](int64_t kb, int k) { return (k <= 16 && hl < k) ? obs_cam[kb + hl] : -1; };
    auto stage = [&](int b, int buf, int64_t kb, int k, int cam) {
        const int pl = 2 * (w - 8) + h;                    // point of the batch
        const int64_t j = p0 + (int64_t)b * kMmaBP + pl;
        const bool valid = j < p1;
        bool bad = k > 16;
        const bool active = valid && !bad && hl < k;       // this lane carries one observation of the point
        double rx = 0.0, ry = 0.0, jp[6], jc[20];
        if (active) {
            const int64_t o = kb + hl;
            rx = J[o]; ry = J[O + o];
#pragma unroll
            for (int i = 0; i < 6; ++i) jp[i] = J[(int64_t)(2 + i) * O + o];
#pragma unroll
            for (int i = 0; i < 20; ++i) jc[i] = J[(int64_t)(8 + i) * O + o];
        } else {
#pragma unroll
            for (int i = 0; i < 6; ++i) jp[i] = 0.0;
#pragma unroll
            for (int i = 0; i < 20; ++i) jc[i] = 0.0;
        }
        int loc = -1;
        if (active) {
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
        if (use && active) {
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
            for (int s2 = 8; s2 > 0; s2 >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s2);   // stays inside the half-warp
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
        // which table slots does this point fill?
        sm.slot_obs[pl][hl] = 0;
        __syncwarp();
        if (contrib && active) sm.slot_obs[pl][loc] = 1;
        __syncwarp();
        if (contrib && active) {
            double* Fr = sm.F[buf] + (3 * pl) * kMmaSLD + 10 * loc;
            double* Wr = sm.W[buf] + (3 * pl) * kMmaSLD + 10 * loc;
            // F_i = (2 Jp_i)^T Jc_i and W_i = E^-1 F_i = (E^-1 (2 Jp_i)^T) Jc_i: both are [3 x 2][2 x 10] products once the two 3x2 factors
            // are formed (18 multiply-adds per observation instead of 90 for E^-1 applied to every column of F).  The producers share
            // the FP64 pipe with the consumers' DMMAs and wait behind them: every instruction less here shortens the batch.
            double p2[6], qv[6];
#pragma unroll
            for (int i = 0; i < 6; ++i) p2[i] = 2.0 * jp[i];
            // qv[v*2 + comp] = sum_u inv(v,u) * p2[u*2 + comp]
#pragma unroll
            for (int cpt = 0; cpt < 2; ++cpt) {
                qv[0 + cpt] = inv[0] * p2[0 + cpt] + inv[1] * p2[2 + cpt] + inv[2] * p2[4 + cpt];
                qv[2 + cpt] = inv[1] * p2[0 + cpt] + inv[3] * p2[2 + cpt] + inv[4] * p2[4 + cpt];
                qv[4 + cpt] = inv[2] * p2[0 + cpt] + inv[4] * p2[2 + cpt] + inv[5] * p2[4 + cpt];
            }
#pragma unroll
            for (int a = 0; a < 10; a += 2) {
                double f[3][2], ww[3][2];
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const double j0 = jc[(a + u) * 2], j1 = jc[(a + u) * 2 + 1];
#pragma unroll
                    for (int v = 0; v < 3; ++v) {
                        f[v][u] = p2[2 * v] * j0 + p2[2 * v + 1] * j1;
                        ww[v][u] = qv[2 * v] * j0 + qv[2 * v + 1] * j1;
                    }
                }
#pragma unroll
                for (int v = 0; v < 3; ++v) {
                    *reinterpret_cast<double2*>(Fr + v * kMmaSLD + a) = make_double2(f[v][0], f[v][1]);
                    *reinterpret_cast<double2*>(Wr + v * kMmaSLD + a) = make_double2(ww[v][0], ww[v][1]);
                }
            }
        }
        if (hl < kMmaCams && sm.slot_obs[pl][hl] == 0) {   // slots the point does not see (all of them for an absent / skipped / deferred point)
            double* Fr = sm.F[buf] + (3 * pl) * kMmaSLD + 10 * hl;
            double* Wr = sm.W[buf] + (3 * pl) * kMmaSLD + 10 * hl;
#pragma unroll
            for (int v = 0; v < 3; ++v)
#pragma unroll
                for (int a = 0; a < 10; a += 2) {
                    *reinterpret_cast<double2*>(Fr + v * kMmaSLD + a) = make_double2(0.0, 0.0);
                    *reinterpret_cast<double2*>(Wr + v * kMmaSLD + a) = make_double2(0.0, 0.0);
                }
        }
    };

    // Producers and consumers run separate loops with the same number of CTA-wide barriers (bar.sync 0 counts arrivals, it
    // does not care from which instruction they come), so the consumers' accumulators are not live in the producer code.
    if (producer) {
        // software pipeline over batches: track extents two batches ahead, camera ids one batch ahead, Jacobian rows in the batch itself
        int64_t kb0, kb1; int k0, k1, cam0, cam1;
        load_track(0, kb0, k0);
        load_track(1, kb1, k1);
        cam0 = load_cam(kb0, k0);
        cam1 = load_cam(kb1, k1);
        stage(0, 0, kb0, k0, cam0);
        cta_barrier();
        for (int b = 0; b < nbatch; ++b) {   // consumers work on batch b while batch b+1 is staged
            kb0 = kb1; k0 = k1; cam0 = cam1;
            load_track(b + 2, kb1, k1);
            if (b + 1 < nbatch) stage(b + 1, (b & 1) ^ 1, kb0, k0, cam0);
            cam1 = load_cam(kb1, k1);
            cta_barrier();
        }
        return;
    }

    // ---- consumers: super-block sb = I(I+1)/2 + Jc covers rows [24 I, 24 I + 24) x columns [24 Jc, 24 Jc + 24)
    const int g = lane >> 2, tg = lane & 3;
    // Assignment of the 15 super-blocks (24x24) to the 8 consumer warps.  A warp issues its DMMAs at a fixed rate, so the warp with
    // the most fragments sets the pace of a batch: a diagonal super-block (d) only computes the 6 fragments on and below its diagonal,
    // and fragments in rows / columns beyond the tile's last camera slot (10 * nLocal) are not computed at all -- with 11 cameras in
    // the table the 15th fragment row is empty and the super-blocks of row I = 4 shrink to 2 fragment rows.  Fragments per warp with
    // 11 (12) cameras: w0 {1,10} 15 (18), w4 {3,11} 15 (18), w1 {4,12} 15 (18), w5 {6,13} 15 (18), w2 {7,0d} 15 (15), w6 {8,2d} 15 (15),
    // w3 {5d,9d} 12 (12), w7 {14d} 3 (6) + the rhs GEMV.  It was 18 for every warp, zeros included.
    int sbI[2] = {0, 0}, sbJ[2] = {0, 0}, code[2] = {-1, -1};
    int nsb = 0;
    {
        constexpr int kSbOf[8][2] = {{1, 10}, {4, 12}, {7, 0}, {5, 9}, {3, 11}, {6, 13}, {8, 2}, {14, -1}};
        const int lim = 10 * nLocal;
        for (int q = 0; q < 2; ++q) {
            const int sb = kSbOf[w][q];
            if (sb < 0) break;
            int I = 0;
            while ((I + 1) * (I + 2) / 2 <= sb) ++I;
            sbI[q] = I; sbJ[q] = sb - I * (I + 1) / 2;
            nsb = q + 1;
            int ni = (lim - 24 * sbI[q] + 7) / 8; ni = ni < 0 ? 0 : (ni > 3 ? 3 : ni);
            int nj = (lim - 24 * sbJ[q] + 7) / 8; nj = nj < 0 ? 0 : (nj > 3 ? 3 : nj);
            if (ni > 0 && nj > 0) code[q] = sbI[q] == sbJ[q] ? 9 + (ni - 1) : (ni - 1) * 3 + (nj - 1);
        }
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
        sb_dispatch(code[0], acc[0], F + tg * kMmaSLD + 24 * sbI[0] + g, W + tg * kMmaSLD + 24 * sbJ[0] + g);
        if (nsb > 1) sb_dispatch(code[1], acc[1], F + tg * kMmaSLD + 24 * sbI[1] + g, W + tg * kMmaSLD + 24 * sbJ[1] + g);
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

    // ---- flush: one red.global.add.f64 per touched entry per tile.  Index work comes from the per-tile tables (gidx, blk); the
    // loops stay unrolled (register accumulators) but the body is a handful of instructions.
    const bool dense = sink.blocks == nullptr;
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        if (q >= nsb) break;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const int row = 24 * sbI[q] + 8 * i + g;
            const int ri = sm.gidx[row], si = row / 10;
#pragma unroll
            for (int jj = 0; jj < 3; ++jj)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const double v = acc[q][i][jj][e];
                    const int col = 24 * sbJ[q] + 8 * jj + 2 * tg + e;
                    const int ci = sm.gidx[col], sl = col / 10;
                    // lower block triangle; inside a diagonal block the lower triangle (row >= col) is computed once
                    if (v == 0.0 || ri < 0 || ci < 0 || sl > si || (sl == si && col > row)) continue;
                    if (dense) {
                        atomicAdd(&sink.S[(size_t)ci * sink.ld + ri], -v);
                    } else {
                        const int blk = sm.blk[si * kMmaCams + sl];
                        if (blk < 0) continue;
                        const int a = row - 10 * si, bq = col - 10 * sl;
                        atomicAdd(&sink.blocks[(size_t)blk * 100 + a * 10 + bq], -v);
                        if (sl == si && a != bq) atomicAdd(&sink.blocks[(size_t)blk * 100 + bq * 10 + a], -v);   // stored diagonal blocks are full
                    }
                }
        }
    }
    if (w == 7) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int m = lane + 32 * q;
            if (m >= kMmaRows || racc[q] == 0.0) continue;
            const int r = sm.gidx[m];
            if (r < 0) continue;
            const int slot = m / 10;
            if (dense) atomicAdd(&sink.rhs[r], racc[q]);
            else atomicAdd(&sink.rhs[(size_t)sm.tab[slot] * 10 + (m - 10 * slot)], racc[q]);
        }
    }
}

// Bind-time plan of the dense path: which points the tile kernel will leave to the per-point kernel (structure only: more than
// 16 observations, or a camera outside the tile's table) and how many there are.  Same table construction as k_schur_mma; one
// thread per point.
__global__ void __launch_bounds__(256) k_schur_plan(int64_t N, int tile_points, const int64_t* __restrict__ pt_begin, const int32_t* __restrict__ obs_cam,
                                                    unsigned char* __restrict__ deferred, unsigned long long* __restrict__ n_deferred) {
    __shared__ int hash[kMmaHash];
    __shared__ int tab[kMmaCams];
    __shared__ int n_local;
    const int tid = threadIdx.x;
    const int64_t p0 = (int64_t)blockIdx.x * tile_points;
    const int64_t p1 = min(N, p0 + (int64_t)tile_points);
    if (p0 >= N) return;
    for (int i = tid; i < kMmaHash; i += blockDim.x) hash[i] = -1;
    __syncthreads();
    const int64_t ob = pt_begin[p0], oe = pt_begin[p1];
    for (int64_t o = ob + tid; o < oe; o += blockDim.x) {
        const int cam = obs_cam[o];
        unsigned h = ((unsigned)cam * 2654435761u) >> 26;
        for (int probe = 0; probe < kMmaHash; ++probe) {
            const int prev = atomicCAS(&hash[h], -1, cam);
            if (prev == -1 || prev == cam) break;
            h = (h + 1) & (kMmaHash - 1);
        }
    }
    __syncthreads();
    if (tid < kMmaHash) {
        const int cam = hash[tid];
        int rank = 0, total = 0;
        for (int i = 0; i < kMmaHash; ++i) { const int other = hash[i]; total += other >= 0; rank += (other >= 0 && other < cam); }
        if (cam >= 0 && rank < kMmaCams) tab[rank] = cam;
        if (tid == 0) n_local = total < kMmaCams ? total : kMmaCams;
    }
    __syncthreads();
    const int nLocal = n_local;
    int mine = 0;
    for (int64_t j = p0 + tid; j < p1; j += blockDim.x) {
        const int64_t kb = pt_begin[j];
        const int k = (int)(pt_begin[j + 1] - kb);
        bool bad = k > 16;
        for (int i = 0; i < k && !bad; ++i) {
            const int cam = obs_cam[kb + i];
            bool found = false;
            for (int q = 0; q < nLocal; ++q) found |= tab[q] == cam;
            bad = !found;
        }
        deferred[j] = bad ? 1 : 0;
        mine += bad;
    }
    mine = __syncthreads_count(mine) > 0 ? mine : 0;
    for (int s2 = 16; s2 > 0; s2 >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, s2);
    if ((tid & 31) == 0 && mine > 0) atomicAdd(n_deferred, (unsigned long long)mine);
}

void launch_schur_plan(cudaStream_t st, int64_t N, int tile_points, const int64_t* pt_begin, const int32_t* obs_cam, unsigned char* deferred,
                       unsigned long long* n_deferred) {
    if (N <= 0) return;
    k_schur_plan<<<(unsigned)((N + tile_points - 1) / tile_points), 256, 0, st>>>(N, tile_points, pt_begin, obs_cam, deferred, n_deferred);
}

void launch_schur_mma(cudaStream_t st, int64_t N, int64_t O, int tile_points, const int64_t* pt_begin, const int32_t* obs_cam, const double* J, double c,
                      const SchurSink& sink, double* pinv, unsigned char* skipped, unsigned char* deferred) {
    if (N <= 0) return;
    static PerDeviceOnce once;
    if (once.first()) cudaFuncSetAttribute(k_schur_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(MmaSmem));
    const unsigned grid = (unsigned)((N + tile_points - 1) / tile_points);
    k_schur_mma<<<grid, kMmaThreads, sizeof(MmaSmem), st>>>(N, O, tile_points, pt_begin, obs_cam, J, c, sink, pinv, skipped, deferred);
}

}  // namespace srk
