// suriko-b200 — nested-dissection order of the reduced camera system for the sparse-factor Cholesky (chol_kernels.cu).
//
// The reduced camera system S couples two cameras only when they observe a common point (BA.cpp:1859-1900 accumulates
// F_j^T E_j^-1 F_j over the cameras of point j).  In capture order S is block-banded and the factorisation is ONE chain of
// n/64 dependent block columns.  Ordered as [part 0 | part 1 | ... | separator], with no coupling between different parts, the
// parts are independent chains that run concurrently (one thread-block cluster each) and only the separator block is left
// for a second, short chain.  The order is found on the host from the camera co-visibility graph (structure only, constant
// over the LM iterations): BFS level structure from a pseudo-peripheral camera, a few whole levels as separators, connected
// components of the rest as parts.  Every part and the separator start on a 64-column boundary (padding unknowns with a unit
// diagonal), so that no 64x64 tile is shared by two parts.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <vector>
#include "kernels.h"

namespace srk {

struct SolveOrder {
    bool active = false;        // false: keep the natural order (dense / small / nothing to gain)
    int n = 0;                  // unknowns of the natural system
    int np = 0;                 // unknowns of the ordered system (n + padding)
    std::vector<int> pos;       // [n]  natural index -> ordered index
    std::vector<int> src;       // [np] ordered index -> natural index, -1 = padding
    CholPartition part{};       // block-column ranges of the parts and of the separator
    int levels = 0, sep_levels = 0, sep_blocks = 0, max_part_blocks = 0, max_mid_blocks = 0;   // diagnostics (blocks of the top separator, the longest leaf, the longest second-level separator)
    // 64x64 tile structure (from the same graph, so a superset of whatever the damping / skip rule leaves non-zero), which lets
    // every pass over the system touch its ~4 % non-zero tiles instead of n^2 entries.  Tiles are packed (row << 16) | col, row >= col.
    std::vector<int> s_tiles;            // natural-order system: lower tiles that can hold a non-zero
    std::vector<int> res_ptr, res_ent;   // per natural row block R: entries (other << 1) | t;  t = 0: tile (R, other);  t = 1: tile (other, R) transposed
    std::vector<int> l_in_tiles;         // ordered system: lower tiles of P S P^T that can hold a non-zero (every diagonal tile included)
    std::vector<int> l_all_tiles;        // ... plus the fill of the factorisation (symbolic, tile level)
    std::vector<unsigned char> l_pattern;   // [nblk x nblk] l_pattern[c*nblk + r] = 1 for the strictly lower tiles of l_in_tiles
    int l_pattern_count = 0;
};

// groups: consecutive runs of unknowns that stay together (one camera each); gsize[g] unknowns; adj[g*G + h] != 0 iff the groups are
// coupled (symmetric, diagonal ignored).  Returns an inactive order when a partition would not shorten the chain.
SolveOrder build_solve_order(int G, const int* gsize, const unsigned char* adj);

// L (np x np, ldp, lower triangle + diagonal) <- P S P^T; S is n x n (ld) with BOTH triangles stored when `mirrored`, else the lower one.
// Padding unknowns get a unit diagonal.  src_dev: [np] ordered -> natural (-1 = padding).
void launch_permute_sym(cudaStream_t st, int n, const double* S, int64_t ld, int mirrored, int np, const int* src_dev, double* L, int64_t ldp);
// Tile-list forms (SolveOrder::s_tiles / l_in_tiles / l_all_tiles / res_*), one CTA per tile or row block:
void launch_zero_tiles(cudaStream_t st, int n, double* A, int64_t ld, const int* tiles, int count);
// listed tiles of L <- P S P^T read from the LOWER triangle of S only
void launch_permute_tiles(cudaStream_t st, int n, const double* S, int64_t ld, int np, const int* src_dev, double* L, int64_t ldp, const int* tiles, int count);
// r = b - S x, S symmetric with only the lower tiles stored, double-double accumulation in a fixed order
void launch_residual_dd_tiles(cudaStream_t st, int n, const double* S, int64_t ld, const double* x, const double* b, double* r, const int* res_ptr, const int* res_ent);
// out[i] = src[i] >= 0 ? in[src[i]] : 0     (natural -> ordered vector)
void launch_gather_vec(cudaStream_t st, int np, const int* src_dev, const double* in, double* out);
// out[src[i]] = in[i] for src[i] >= 0       (ordered -> natural vector)
void launch_scatter_vec(cudaStream_t st, int np, const int* src_dev, const double* in, double* out);
// adj[ci*M + cl] = adj[cl*M + ci] = 1 for every pair of cameras that observe a common point
void launch_cam_adjacency(cudaStream_t st, int64_t N, const int64_t* pt_begin, const int32_t* obs_cam, int M, unsigned char* adj);
// the f64 sum all-reduce of the multi-GPU path carries the union over ranks: bytes -> doubles -> (all-reduce) -> bytes
void launch_bytes_to_doubles(cudaStream_t st, int64_t n, const unsigned char* in, double* out);
void launch_doubles_to_bytes(cudaStream_t st, int64_t n, const double* in, unsigned char* out);

}  // namespace srk
