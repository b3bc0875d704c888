// suriko-b200 — K2, third form: the Schur complement of a tile of points as ONE symmetric rank-k update on the FP64 tensor pipe.
//
// Replaces the Schur loop of EstimateCorrectionsDecomposedInTwoPhases (BA.cpp:1859-1900):
//     S[cam_i, cam_l] -= sum_j F_ji^T E_cj^-1 F_jl ,   rhs[cam_i] += sum_j F_ji^T E_cj^-1 g_pj
// What ncu said about the second form (schur_mma.cu, profiles/r01_ncu_schur_hot_lines.txt): the FP64 pipe is shared by the consumers'
// DMMAs and the producers' vector arithmetic; a producer DFMA queues behind 16-cycle DMMAs, so the ~256 FP64 instructions a producer
// lane spent per batch (E reduction by shuffles, cofactor inverse, F and W = E^-1 F) made the PRODUCERS the bottleneck through pipe
// arbitration: the consumers sat 30 % of their time at the batch barrier with the DMMA sub-pipe 50 % busy.  This form takes the work
// off that pipe instead of scheduling around it:
//   * k_point_factor (a small streaming kernel, once per attempt): per point E = 2 sum Jp^T Jp, g_p, the damped cofactor inverse with the
//     |det| > 1e-12 rule (quirk Q5, BA.cpp:1873-1881) -> pinv / skipped for K2' exactly as before, plus the Cholesky factor of the damped
//     block E_c = L L^T as Gi = L^-1 (6 doubles) and u = Gi g_p.  With V_j = Gi_j F_j:   F^T E_c^-1 F = V^T V,   F^T E_c^-1 g_p = V^T u,
//     so the tile contraction needs ONE operand (half the shared memory, half the stores) and the producers no reductions at all.
//   * bind-time tables (k_schur_tables; the observation order never changes): the sorted camera table of every tile, the table slot of
//     every observation (1 byte), the slot mask of every point -- no per-launch hash / sort / search.
//   * k_schur_v3: 8 producer warps, one LANE PER OBSERVATION (every lane of every load carries data): Q = Gi (2 Jp^T) (3x2), V = Q Jc
//     (3x10: 72 FP64 instructions per lane instead of ~256), written to a 4-deep ring of [48 x 124] operand stages; unseen slots are
//     zero-filled from the point masks.  8 consumer warps run D += V^T V over the lower triangle of 8x8 fragments with mma.sync.m8n8k4.f64,
//     super-blocks dealt so that the four SM sub-partitions carry 27 / 27 / 27 / 24 + rhs fragments (30 / 30 / 30 / 15 before).
//     Producers and consumers meet through named barriers per ring stage (bar.arrive / bar.sync: full[s], empty[s]) -- the producers run
//     up to three batches ahead, so a slow batch (a DRAM round trip) no longer stalls the tensor pipe.
// A point whose damped block passes the |det| rule but is not numerically positive definite (never for a sum of squares with damping
// unless rounding dominates) has no real factor: it contributes through the exception list to the per-point kernel (two operands).
#include <stdlib.h>
#include "common.cuh"
#include "kernels.h"

namespace srk {

constexpr int kV3Cams = 12;              // cameras per tile table (== the plan kernel's CMAX, schur_mma.cu kMmaCams)
constexpr int kV3Rows = kV3Cams * 10;    // 120
constexpr int kV3SLD = 124;              // row stride of an operand stage: 124 = 12 (mod 16) -> conflict-free fragment loads
constexpr int kV3BP = 16;                // points per batch
constexpr int kV3K = 3 * kV3BP;          // 48 contraction rows per batch
constexpr int kV3Threads = 512;
constexpr int kV3Stages = 4;
constexpr int kV3Hash = 64;

struct V3Smem {
    double V[kV3Stages][kV3K * kV3SLD];
    double U[kV3Stages][kV3K];
    double G[kV3Stages][kV3BP * 6];      // Gi of the batch's points
    int gidx[kV3Rows];
    int blk[kV3Cams * kV3Cams];
    int tab[kV3Cams];
};

__device__ __forceinline__ void bar_sync_named(int id, int count) { asm volatile("bar.sync %0, %1;\n" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void bar_arrive_named(int id, int count) { asm volatile("bar.arrive %0, %1;\n" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void dmma884_v3(double& d0, double& d1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// a9 = {E00, E01, E02, E11, E12, E22, g_p0, g_p1, g_p2} of point j -> pinv / skipped (K2'), Gi, u (tile kernel), exception list
__device__ __forceinline__ void point_factor_store(int64_t j, int64_t N, const double* a9, double c, double* __restrict__ pinv, unsigned char* __restrict__ skipped,
                                                   double* __restrict__ gi, double* __restrict__ uvec, int* __restrict__ exc_list, int* __restrict__ exc_count,
                                                   int exc_cap) {
    double inv[6];
    const bool ok = point_block_inverse(a9, c, inv);
    skipped[j] = ok ? 0 : 1;
#pragma unroll
    for (int i = 0; i < 6; ++i) pinv[(int64_t)i * N + j] = ok ? inv[i] : 0.0;
#pragma unroll
    for (int i = 0; i < 3; ++i) pinv[(int64_t)(6 + i) * N + j] = a9[6 + i];
    // Cholesky of the damped block (same damped entries as point_block_inverse), Gi = L^-1
    double g[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0}, u[3] = {0.0, 0.0, 0.0};   // g = {Gi00, Gi10, Gi11, Gi20, Gi21, Gi22}
    if (ok) {
        const double m00 = a9[0] * (1.0 + c), m11 = a9[3] * (1.0 + c), m22 = a9[5] * (1.0 + c), m01 = a9[1], m02 = a9[2], m12 = a9[4];
        bool pd = m00 > 0.0;
        const double l00 = sqrt(m00);
        const double l10 = m01 / l00, l20 = m02 / l00;
        const double d11 = m11 - l10 * l10;
        pd = pd && d11 > 0.0;
        const double l11 = sqrt(d11);
        const double l21 = (m12 - l20 * l10) / l11;
        const double d22 = m22 - l20 * l20 - l21 * l21;
        pd = pd && d22 > 0.0;
        const double l22 = sqrt(d22);
        const double i00 = 1.0 / l00, i11 = 1.0 / l11, i22 = 1.0 / l22;
        const double g10 = -l10 * i00 * i11, g21 = -l21 * i11 * i22;
        const double g20 = -(l20 * i00 + l21 * g10) * i22;
        pd = pd && isfinite(i00) && isfinite(i11) && isfinite(i22) && isfinite(g10) && isfinite(g20) && isfinite(g21);
        if (pd) {
            g[0] = i00; g[1] = g10; g[2] = i11; g[3] = g20; g[4] = g21; g[5] = i22;
            u[0] = g[0] * a9[6];
            u[1] = g[1] * a9[6] + g[2] * a9[7];
            u[2] = g[3] * a9[6] + g[4] * a9[7] + g[5] * a9[8];
        } else {   // invertible by the reference's rule but without a real factor: the two-operand per-point kernel takes it
            const int idx = atomicAdd(exc_count, 1);
            if (idx < exc_cap) exc_list[idx] = (int)j;
        }
    }
#pragma unroll
    for (int i = 0; i < 6; ++i) gi[(int64_t)i * N + j] = g[i];
#pragma unroll
    for (int i = 0; i < 3; ++i) uvec[(int64_t)i * N + j] = u[i];
}

// ---------------------------------------------------------------------------------------------------------------------------------
// Per-point blocks of one attempt.  Half-warp per point (lanes over its observations, strided for long tracks).
__global__ void __launch_bounds__(256) k_point_factor(int64_t N, int64_t O, const int64_t* __restrict__ pt_begin, const double* __restrict__ J, double c,
                                                      double* __restrict__ pinv, unsigned char* __restrict__ skipped, double* __restrict__ gi,
                                                      double* __restrict__ uvec, int* __restrict__ exc_list, int* __restrict__ exc_count, int exc_cap) {
    const int hl = threadIdx.x & 15;
    const int64_t j = ((int64_t)blockIdx.x * 256 + threadIdx.x) >> 4;
    const bool valid = j < N;
    int64_t kb = 0; int k = 0;
    if (valid) { kb = pt_begin[j]; k = (int)(pt_begin[j + 1] - kb); }
    double a9[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) a9[i] = 0.0;
    for (int i = hl; i < k; i += 16) {
        const int64_t o = kb + i;
        const double rx = J[o], ry = J[O + o];
        double jp[6];
#pragma unroll
        for (int q = 0; q < 6; ++q) jp[q] = J[(int64_t)(2 + q) * O + o];
        a9[0] += jp[0] * jp[0] + jp[1] * jp[1];
        a9[1] += jp[0] * jp[2] + jp[1] * jp[3];
        a9[2] += jp[0] * jp[4] + jp[1] * jp[5];
        a9[3] += jp[2] * jp[2] + jp[3] * jp[3];
        a9[4] += jp[2] * jp[4] + jp[3] * jp[5];
        a9[5] += jp[4] * jp[4] + jp[5] * jp[5];
        a9[6] += jp[0] * rx + jp[1] * ry;
        a9[7] += jp[2] * rx + jp[3] * ry;
        a9[8] += jp[4] * rx + jp[5] * ry;
    }
#pragma unroll
    for (int i = 0; i < 9; ++i) {
        double v = a9[i];
#pragma unroll
        for (int s2 = 8; s2 > 0; s2 >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s2);   // stays inside the half-warp
        a9[i] = 2.0 * v;
    }
    if (!valid || hl != 0) return;
    point_factor_store(j, N, a9, c, pinv, skipped, gi, uvec, exc_list, exc_count, exc_cap);
}

// The same from the per-point sums K1 accumulated (k_jacobian, Eacc [9N]: sum Jp^T Jp (6), sum Jp^T rho (3)): one thread per point.
__global__ void __launch_bounds__(256) k_point_finish(int64_t N, const double* __restrict__ Eacc, double c, double* __restrict__ pinv,
                                                      unsigned char* __restrict__ skipped, double* __restrict__ gi, double* __restrict__ uvec,
                                                      int* __restrict__ exc_list, int* __restrict__ exc_count, int exc_cap) {
    const int64_t j = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (j >= N) return;
    double a9[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) a9[i] = 2.0 * Eacc[(int64_t)i * N + j];
    point_factor_store(j, N, a9, c, pinv, skipped, gi, uvec, exc_list, exc_count, exc_cap);
}

// ---------------------------------------------------------------------------------------------------------------------------------
// Bind-time tables of the tile kernel.  Same table construction and the same deferral rule as k_schur_plan (schur_mma.cu): the
// kV3Cams smallest distinct camera ids of the tile, sorted; a point is deferred when it has more than 16 observations or sees a camera
// outside the table.  Outputs: tile_tab [tiles x 12], tile_n [tiles], obs_slot [O] (point-in-batch << 4 | table slot; slot 0xF for observations of deferred points),
// pt_mask [N] (bit s = the point sees table slot s; 0 for deferred points), deferred [N] (rewritten with the same values).
__global__ void __launch_bounds__(256) k_schur_tables(int64_t N, int tile_points, const int64_t* __restrict__ pt_begin, const int32_t* __restrict__ obs_cam,
                                                      int* __restrict__ tile_tab, int* __restrict__ tile_n, unsigned char* __restrict__ obs_slot,
                                                      unsigned short* __restrict__ pt_mask, unsigned char* __restrict__ deferred) {
    __shared__ int hash[kV3Hash];
    __shared__ int tab[kV3Cams];
    __shared__ int n_local;
    const int tid = threadIdx.x;
    const int64_t p0 = (int64_t)blockIdx.x * tile_points;
    const int64_t p1 = min(N, p0 + (int64_t)tile_points);
    if (p0 >= N) return;
    for (int i = tid; i < kV3Hash; i += blockDim.x) hash[i] = -1;
    if (tid < kV3Cams) tab[tid] = -1;
    __syncthreads();
    const int64_t ob = pt_begin[p0], oe = pt_begin[p1];
    for (int64_t o = ob + tid; o < oe; o += blockDim.x) {
        const int cam = obs_cam[o];
        unsigned h = ((unsigned)cam * 2654435761u) >> 26;
        for (int probe = 0; probe < kV3Hash; ++probe) {
            const int prev = atomicCAS(&hash[h], -1, cam);
            if (prev == -1 || prev == cam) break;
            h = (h + 1) & (kV3Hash - 1);
        }
    }
    __syncthreads();
    if (tid < kV3Hash) {
        const int cam = hash[tid];
        int rank = 0, total = 0;
        for (int i = 0; i < kV3Hash; ++i) { const int other = hash[i]; total += other >= 0; rank += (other >= 0 && other < cam); }
        if (cam >= 0 && rank < kV3Cams) tab[rank] = cam;
        if (tid == 0) n_local = total < kV3Cams ? total : kV3Cams;
    }
    __syncthreads();
    const int nLocal = n_local;
    if (tid < kV3Cams) tile_tab[(size_t)blockIdx.x * kV3Cams + tid] = tab[tid];
    if (tid == 0) tile_n[blockIdx.x] = nLocal;
    for (int64_t j = p0 + tid; j < p1; j += blockDim.x) {
        const int64_t kb = pt_begin[j];
        const int k = (int)(pt_begin[j + 1] - kb);
        bool bad = k > 16;
        unsigned mask = 0;
        for (int i = 0; i < k && !bad; ++i) {
            const int cam = obs_cam[kb + i];
            int loc = -1;
            for (int q = 0; q < nLocal; ++q) if (tab[q] == cam) loc = q;
            if (loc < 0) bad = true; else mask |= 1u << loc;
        }
        deferred[j] = bad ? 1 : 0;
        pt_mask[j] = bad ? (unsigned short)0 : (unsigned short)mask;
        const int pl = (int)((j - p0) % kV3BP);       // the point's position inside its batch of the tile (tiles start on batch boundaries)
        for (int i = 0; i < k; ++i) {
            int loc = 0xF;
            if (!bad) { const int cam = obs_cam[kb + i]; for (int q = 0; q < nLocal; ++q) if (tab[q] == cam) loc = q; }
            obs_slot[kb + i] = (unsigned char)((pl << 4) | loc);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------------------
#include "schur_v3_consume.inc"
#include "schur_v3_consume4.inc"

// CW = consumer warps: 8 (two per SM sub-partition, 15 accumulator fragments each; the rhs GEMV on consumer warp 7) or 4 (ONE per sub-partition
// with 30 fragments: 30 independent DMMAs per k-step and operand load).  Eight producer warps either way.
// RP: the rhs GEMV (rhs += V^T u, vector FP64) runs with the producers instead of on the last consumer warp.  tools/probe/dmma_micro.cu: ONE
// warp per sub-partition issues a DMMA every ~25 cycles (62 % of the pipe), TWO saturate it -- so every cycle a consumer warp spends on anything
// but DMMAs (the GEMV took ~20 % of warp 7's time) idles its sub-partition's pipe, and all consumers wait for the slowest at the batch barrier.
template <int CW, bool RP>
__global__ void __launch_bounds__(32 * (CW + 8), 1) k_schur_v3(int64_t N, int64_t O, int tile_points, const int64_t* __restrict__ pt_begin,
                                                            const int32_t* __restrict__ obs_pt, const double* __restrict__ J, SchurSink sink,
                                                            const double* __restrict__ gi, const double* __restrict__ uvec,
                                                            const unsigned char* __restrict__ skipped, const int* __restrict__ tile_tab,
                                                            const int* __restrict__ tile_n, const unsigned char* __restrict__ obs_slot,
                                                            const unsigned short* __restrict__ pt_mask, int dbg_mode) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    V3Smem& sm = *reinterpret_cast<V3Smem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int64_t p0 = (int64_t)blockIdx.x * tile_points;
    const int64_t p1 = min(N, p0 + (int64_t)tile_points);
    if (p0 >= N) return;
    const int nbatch = (int)((p1 - p0 + kV3BP - 1) / kV3BP);
    const int nLocal = tile_n[blockIdx.x];
    constexpr int NT = 32 * (CW + 8);
    if (RP) {           // the flush tables up front: the producers flush the rhs themselves
        if (tid < kV3Cams) sm.tab[tid] = tile_tab[(size_t)blockIdx.x * kV3Cams + tid];
        __syncthreads();
        for (int m = tid; m < kV3Rows; m += NT) {
            const int slot = m / 10, a = m - 10 * slot;
            sm.gidx[m] = slot < nLocal ? red_index(sm.tab[slot], a, sink.unity) : -1;
        }
        if (sink.blocks != nullptr) {
            for (int e = tid; e < kV3Cams * kV3Cams; e += NT) {
                const int si = e / kV3Cams, sl = e % kV3Cams;
                sm.blk[e] = (si < nLocal && sl <= si) ? sink_block_id(sink, sm.tab[si], sm.tab[sl]) : -1;
            }
        }
        __syncthreads();
    }

    // dbg_mode (timing experiments only, results are garbage): 1 = consumers alone (no producers, no barriers), 2 = producers alone
    if (dbg_mode == 1 && w >= CW) return;
    if (dbg_mode == 2 && w < CW) return;
    if (w >= CW) {
        // ================= producers: one lane per observation of the batch
        // Every global address of a batch is known before the batch starts: the observation extents come one batch ahead, the per-point
        // Gi / u / slot masks are addressed by the batch's first point id (staged to shared memory by helper lanes) and the position of an
        // observation's point inside the batch travels in the slot byte -- ONE round trip per batch, no dependent loads.
        const int ptid = tid - 32 * CW;
        double rhs_acc = 0.0;                         // RP: threads 0..239 own (column m, half of the batch's rows) of rhs += V^T u
        auto rhs_of_stage = [&](int st2) {
            if (ptid >= 2 * kV3Rows) return;
            const int m = ptid % kV3Rows, k0 = (ptid / kV3Rows) * (kV3K / 2);
            const double* Vq = sm.V[st2];
            const double* U = sm.U[st2];
            double sacc = rhs_acc;
#pragma unroll 8
            for (int kk = k0; kk < k0 + kV3K / 2; ++kk) sacc += Vq[kk * kV3SLD + m] * U[kk];
            rhs_acc = sacc;
        };
        int64_t ob = pt_begin[p0], oe = pt_begin[min(p1, p0 + (int64_t)kV3BP)];
        for (int b = 0; b < nbatch; ++b) {
            const int s = b % kV3Stages;
            const int64_t pb0 = p0 + (int64_t)b * kV3BP;
            const int64_t pb1 = min(p1, pb0 + kV3BP);
            int64_t ob_n = 0, oe_n = 0;
            if (b + 1 < nbatch) { ob_n = oe; oe_n = pt_begin[min(p1, pb0 + 2 * (int64_t)kV3BP)]; }
            if (b >= kV3Stages && dbg_mode == 0) bar_sync_named(1 + kV3Stages + s, NT);     // empty[s]: the consumers are done with batch b - stages
            double* Vs = sm.V[s];
            // ---- this lane's observation (first pass): loads in flight before anything is waited for
            int64_t o = ob + ptid;
            bool have = o < oe;
            int sb = 0xF;
            double jp[6], jc[20];
            if (have) {
                sb = obs_slot[o];
#pragma unroll
                for (int i = 0; i < 6; ++i) jp[i] = J[(int64_t)(2 + i) * O + o];
#pragma unroll
                for (int i = 0; i < 20; ++i) jc[i] = J[(int64_t)(8 + i) * O + o];
            }
            // ---- rhs += V^T u of the batch BEFORE this one, while this batch's loads are in flight (the stage is only rewritten by the
            // producers themselves, a ring turn later)
            if (RP && b > 0) rhs_of_stage((b - 1) % kV3Stages);
            // ---- helper lanes: per-point data of the batch
            if (ptid < kV3BP * 6) {
                const int pl = ptid / 6, i = ptid - 6 * pl;
                const int64_t j = pb0 + pl;
                sm.G[s][ptid] = j < pb1 ? gi[(int64_t)i * N + j] : 0.0;
            } else if (ptid < kV3BP * 6 + kV3K) {     // u = Gi g_p of the batch's points (zero for absent / deferred / skipped ones)
                const int e = ptid - kV3BP * 6;
                const int pl = e / 3, v = e - 3 * pl;
                const int64_t j = pb0 + pl;
                double val = 0.0;
                if (j < pb1 && pt_mask[j] != 0) val = uvec[(int64_t)v * N + j];
                sm.U[s][e] = val;
            }
            if (ptid < kV3BP * kV3Cams) {           // slots the point does not see (all of them for an absent / deferred point)
                const int pl = ptid / kV3Cams, sl = ptid - pl * kV3Cams;
                const int64_t j = pb0 + pl;
                const unsigned mask = j < pb1 ? (unsigned)pt_mask[j] : 0u;
                if (((mask >> sl) & 1u) == 0u) {
                    double* r = Vs + (3 * pl) * kV3SLD + 10 * sl;
#pragma unroll
                    for (int v = 0; v < 3; ++v)
#pragma unroll
                        for (int a = 0; a < 10; a += 2) *reinterpret_cast<double2*>(r + v * kV3SLD + a) = make_double2(0.0, 0.0);
                }
            }
            bar_sync_named(10, 256);                 // producers only: G[s] is complete
            for (;;) {
                if (have && (sb & 0xF) != 0xF) {
                    const int pl = sb >> 4, sl = sb & 0xF;
                    const double* g = sm.G[s] + 6 * pl;   // {g00, g10, g11, g20, g21, g22}; all zero for a skipped point (BA.cpp:1877-1881: no contribution)
                    // Q = Gi (2 Jp^T): q[v][comp], jp[u*2 + comp]
                    double q[3][2];
#pragma unroll
                    for (int cpt = 0; cpt < 2; ++cpt) {
                        const double a0 = 2.0 * jp[0 + cpt], a1 = 2.0 * jp[2 + cpt], a2 = 2.0 * jp[4 + cpt];
                        q[0][cpt] = g[0] * a0;
                        q[1][cpt] = g[1] * a0 + g[2] * a1;
                        q[2][cpt] = g[3] * a0 + g[4] * a1 + g[5] * a2;
                    }
                    const bool live = g[0] != 0.0;
                    double* r = Vs + (3 * pl) * kV3SLD + 10 * sl;
#pragma unroll
                    for (int a = 0; a < 10; a += 2) {
#pragma unroll
                        for (int v = 0; v < 3; ++v) {
                            double x0 = q[v][0] * jc[a * 2] + q[v][1] * jc[a * 2 + 1];
                            double x1 = q[v][0] * jc[a * 2 + 2] + q[v][1] * jc[a * 2 + 3];
                            if (!live) { x0 = 0.0; x1 = 0.0; }
                            *reinterpret_cast<double2*>(r + v * kV3SLD + a) = make_double2(x0, x1);
                        }
                    }
                }
                o += 256;                             // batches with more than 256 observations (deferred long tracks inside the range): further passes
                if (!__any_sync(0xffffffffu, o < oe)) break;
                have = o < oe;
                if (have) {
                    sb = obs_slot[o];
#pragma unroll
                    for (int i = 0; i < 6; ++i) jp[i] = J[(int64_t)(2 + i) * O + o];
#pragma unroll
                    for (int i = 0; i < 20; ++i) jc[i] = J[(int64_t)(8 + i) * O + o];
                }
            }
            if (RP) bar_sync_named(10, 256);                                        // every producer's rows of V[s] are written (the rhs below reads them)
            if (dbg_mode == 0) bar_arrive_named(1 + s, NT);                // full[s]
            ob = ob_n; oe = oe_n;
        }
        if (RP) rhs_of_stage((nbatch - 1) % kV3Stages);
        if (RP && ptid < 2 * kV3Rows && rhs_acc != 0.0) {
            const int m = ptid % kV3Rows;
            const int r = sm.gidx[m];
            if (r >= 0) {
                const int slot = m / 10;
                if (sink.blocks == nullptr) atomicAdd(&sink.rhs[r], rhs_acc);
                else atomicAdd(&sink.rhs[(size_t)sm.tab[slot] * 10 + (m - 10 * slot)], rhs_acc);
            }
        }
        return;
    }

    // ================= consumers
    // Fragment lists per (warp, R) and the straight-line DMMA code that walks them: schur_v3_consume.inc (tools/gen_schur_v3_consume.py).
    // R = fragment rows of the tile's lower triangle; every warp walks both of its pieces in ONE k-loop (15 DMMAs back to back per k-step
    // at R = 15, operand fragments loaded once per k-step); 30 fragments per SM sub-partition at R = 15, 28 / 28 / 28 / 21 + rhs at R = 14.
    const int g8 = lane >> 2, tg = lane & 3;
    int R = (10 * nLocal + 7) / 8;
    R = R < 1 ? 1 : (R > 15 ? 15 : R);
    if constexpr (CW == 4) {
        double acc[kV4NFMax][2];
#pragma unroll
        for (int f = 0; f < kV4NFMax; ++f) { acc[f][0] = 0.0; acc[f][1] = 0.0; }
        for (int b = 0; b < nbatch; ++b) {
            const int s = b % kV3Stages;
            if (dbg_mode == 0) bar_sync_named(1 + s, NT);                          // full[s]
            v4_consume_generated(w, R, acc, sm.V[s] + tg * kV3SLD + g8);
            if (b + kV3Stages < nbatch && dbg_mode == 0) bar_arrive_named(1 + kV3Stages + s, NT);   // empty[s]
        }
        const bool dense = sink.blocks == nullptr;
        const int nfr = kV4NFrag[R - kV3GenRMin][w];
#pragma unroll
        for (int f = 0; f < kV4NFMax; ++f) {
            if (f >= nfr) break;
            const int rc = kV4FragRC[R - kV3GenRMin][w][f];
            const int row = 8 * (rc >> 4) + g8;
            if (row >= kV3Rows) continue;
            const int ri = sm.gidx[row], si = row / 10;
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const double v = acc[f][e];
                const int col = 8 * (rc & 15) + 2 * tg + e;
                if (col >= kV3Rows) continue;
                const int ci = sm.gidx[col], sl = col / 10;
                if (v == 0.0 || ri < 0 || ci < 0 || sl > si || (sl == si && col > row)) continue;
                if (dense) {
                    atomicAdd(&sink.S[(size_t)ci * sink.ld + ri], -v);
                } else {
                    const int blk = sm.blk[si * kV3Cams + sl];
                    if (blk < 0) continue;
                    const int a = row - 10 * si, bq = col - 10 * sl;
                    atomicAdd(&sink.blocks[(size_t)blk * 100 + a * 10 + bq], -v);
                    if (sl == si && a != bq) atomicAdd(&sink.blocks[(size_t)blk * 100 + bq * 10 + a], -v);
                }
            }
        }
        return;
    } else {
    double acc[15][2];
#pragma unroll
    for (int f = 0; f < 15; ++f) { acc[f][0] = 0.0; acc[f][1] = 0.0; }
    double racc[4] = {0.0, 0.0, 0.0, 0.0};   // warp 7: rhs entries m = lane + 32*q < 120

    for (int b = 0; b < nbatch; ++b) {
        const int s = b % kV3Stages;
        if (dbg_mode == 0) bar_sync_named(1 + s, NT);                      // full[s]
        const double* V = sm.V[s];
        v3_consume_generated(w, R, acc, V + tg * kV3SLD + g8);
        if (!RP && w == 7) {   // rhs += V^T u
            const double* U = sm.U[s];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int m = lane + 32 * q;
                if (m < kV3Rows) {
                    double sacc = racc[q];
#pragma unroll 8
                    for (int kk = 0; kk < kV3K; ++kk) sacc += V[kk * kV3SLD + m] * U[kk];
                    racc[q] = sacc;
                }
            }
        }
        if (b + kV3Stages < nbatch && dbg_mode == 0) bar_arrive_named(1 + kV3Stages + s, NT);   // empty[s]
    }

    // ---- flush: one red.global.add.f64 per touched entry per tile
    if (!RP) {
        if (tid < kV3Cams) sm.tab[tid] = tile_tab[(size_t)blockIdx.x * kV3Cams + tid];
        bar_sync_named(9, 256);
        for (int m = tid; m < kV3Rows; m += 256) {
            const int slot = m / 10, a = m - 10 * slot;
            sm.gidx[m] = slot < nLocal ? red_index(sm.tab[slot], a, sink.unity) : -1;
        }
        if (sink.blocks != nullptr) {
            for (int e = tid; e < kV3Cams * kV3Cams; e += 256) {
                const int si = e / kV3Cams, sl = e % kV3Cams;
                sm.blk[e] = (si < nLocal && sl <= si) ? sink_block_id(sink, sm.tab[si], sm.tab[sl]) : -1;
            }
        }
        bar_sync_named(9, 256);
    }
    const bool dense = sink.blocks == nullptr;
    const int nfr = kV3NFrag[R - kV3GenRMin][w];
#pragma unroll
    for (int f = 0; f < 15; ++f) {
        if (f >= nfr) break;
        const int rc = kV3FragRC[R - kV3GenRMin][w][f];
        const int row = 8 * (rc >> 4) + g8;
        if (row >= kV3Rows) continue;
        const int ri = sm.gidx[row], si = row / 10;
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const double v = acc[f][e];
            const int col = 8 * (rc & 15) + 2 * tg + e;
            if (col >= kV3Rows) continue;
            const int ci = sm.gidx[col], sl = col / 10;
            // lower block triangle; inside a diagonal block the lower triangle (row >= col) is computed once
            if (v == 0.0 || ri < 0 || ci < 0 || sl > si || (sl == si && col > row)) continue;
            if (dense) {
                atomicAdd(&sink.S[(size_t)ci * sink.ld + ri], -v);
            } else {
                const int blk = sm.blk[si * kV3Cams + sl];
                if (blk < 0) continue;
                const int a = row - 10 * si, bq = col - 10 * sl;
                atomicAdd(&sink.blocks[(size_t)blk * 100 + a * 10 + bq], -v);
                if (sl == si && a != bq) atomicAdd(&sink.blocks[(size_t)blk * 100 + bq * 10 + a], -v);   // stored diagonal blocks are full
            }
        }
    }
    if (!RP && w == 7) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int m = lane + 32 * q;
            if (m >= kV3Rows || racc[q] == 0.0) continue;
            const int r = sm.gidx[m];
            if (r < 0) continue;
            const int slot = m / 10;
            if (dense) atomicAdd(&sink.rhs[r], racc[q]);
            else atomicAdd(&sink.rhs[(size_t)sm.tab[slot] * 10 + (m - 10 * slot)], racc[q]);
        }
    }

    }
}

void launch_point_factor(cudaStream_t st, int64_t N, int64_t O, const int64_t* pt_begin, const double* J, double c, double* pinv, unsigned char* skipped,
                         double* gi, double* uvec, int* exc_list, int* exc_count, int exc_cap) {
    if (N <= 0) return;
    k_point_factor<<<(unsigned)((N * 16 + 255) / 256), 256, 0, st>>>(N, O, pt_begin, J, c, pinv, skipped, gi, uvec, exc_list, exc_count, exc_cap);
}

void launch_point_finish(cudaStream_t st, int64_t N, const double* Eacc, double c, double* pinv, unsigned char* skipped, double* gi, double* uvec, int* exc_list,
                         int* exc_count, int exc_cap) {
    if (N <= 0) return;
    k_point_finish<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(N, Eacc, c, pinv, skipped, gi, uvec, exc_list, exc_count, exc_cap);
}

void launch_schur_tables(cudaStream_t st, int64_t N, int tile_points, const int64_t* pt_begin, const int32_t* obs_cam, int* tile_tab, int* tile_n,
                         unsigned char* obs_slot, unsigned short* pt_mask, unsigned char* deferred) {
    if (N <= 0) return;
    k_schur_tables<<<(unsigned)((N + tile_points - 1) / tile_points), 256, 0, st>>>(N, tile_points, pt_begin, obs_cam, tile_tab, tile_n, obs_slot, pt_mask, deferred);
}

void launch_schur_v3(cudaStream_t st, int64_t N, int64_t O, int tile_points, const int64_t* pt_begin, const int32_t* obs_pt, const double* J,
                     const SchurSink& sink, const double* gi, const double* uvec, const unsigned char* skipped, const int* tile_tab, const int* tile_n,
                     const unsigned char* obs_slot, const unsigned short* pt_mask) {
    if (N <= 0) return;
    static PerDeviceOnce once;
    if (once.first()) {
        cudaFuncSetAttribute(k_schur_v3<8, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(V3Smem));
        cudaFuncSetAttribute(k_schur_v3<8, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(V3Smem));
        cudaFuncSetAttribute(k_schur_v3<4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(V3Smem));
    }
    static int cons = -1;        // consumer warps: 4 = one per SM sub-partition with 30 fragments each (default: K2 1.77 -> 1.70 ms), SRK_V3_CONS=8 = two with 15 each
    if (cons < 0) { const char* e = getenv("SRK_V3_CONS"); cons = e != nullptr ? atoi(e) : 4; if (cons != 8 && cons != 80) cons = 4; }   // 8: two per sub-partition, rhs with the producers; 80: ... rhs on consumer warp 7 (round 2's first form)
    const unsigned grid = (unsigned)((N + tile_points - 1) / tile_points);
    static int dbg_mode = -1;
    if (dbg_mode < 0) { const char* e = getenv("SRK_V3_DEBUG"); dbg_mode = e != nullptr ? atoi(e) : 0; }
    if (cons == 4) k_schur_v3<4, true><<<grid, 32 * 12, sizeof(V3Smem), st>>>(N, O, tile_points, pt_begin, obs_pt, J, sink, gi, uvec, skipped, tile_tab, tile_n, obs_slot, pt_mask, dbg_mode);
    else if (cons == 8) k_schur_v3<8, true><<<grid, kV3Threads, sizeof(V3Smem), st>>>(N, O, tile_points, pt_begin, obs_pt, J, sink, gi, uvec, skipped, tile_tab, tile_n, obs_slot, pt_mask, dbg_mode);
    else k_schur_v3<8, false><<<grid, kV3Threads, sizeof(V3Smem), st>>>(N, O, tile_points, pt_begin, obs_pt, J, sink, gi, uvec, skipped, tile_tab, tile_n, obs_slot, pt_mask, dbg_mode);
}

}  // namespace srk
