// suriko-b200 — nested-dissection order of the reduced camera system (host) and the kernels that move a system into it.
// See solve_order.h.  The reference solves the system in capture order with a dense Householder QR (BA.cpp:1911); any
// symmetric permutation gives the same solution up to rounding, and the engine refines the solution against the natural-order
// system afterwards (k_residual_dd), so the order is purely a scheduling decision.
#include "solve_order.h"
#include <algorithm>
#include <numeric>
#include <queue>
#include <cstring>
#include <cstdlib>

namespace srk {

namespace {

constexpr int kTile = 64;
// clusters of 8 CTAs that are co-resident on 148 SMs at one CTA per SM: cudaOccupancyMaxActiveClusters says 15 on B200 (a cluster stays
// inside a GPC, and the GPCs have 16 to 20 SMs); a 16th part would run as a second wave.  SRK_SOLVE_MAX_PARTS overrides (development aid)
static int max_concurrent_parts() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("SRK_SOLVE_MAX_PARTS"); v = e != nullptr ? atoi(e) : 15; if (v < 2) v = 2; if (v > CholPartition::kMaxParts) v = CholPartition::kMaxParts; }
    return v;
}
#define kMaxConcurrentParts max_concurrent_parts()

struct Graph {
    int G;
    std::vector<std::vector<int>> nb;
};

// BFS levels over ALL components (later components continue the level numbering); returns the number of levels
int bfs_levels(const Graph& g, int root, std::vector<int>& level) {
    level.assign(g.G, -1);
    int maxl = -1;
    auto run = [&](int r, int base) {
        std::queue<int> q;
        level[r] = base; q.push(r);
        while (!q.empty()) {
            const int u = q.front(); q.pop();
            maxl = std::max(maxl, level[u]);
            for (int v : g.nb[u]) if (level[v] < 0) { level[v] = level[u] + 1; q.push(v); }
        }
    };
    run(root, 0);
    for (int u = 0; u < g.G; ++u) if (level[u] < 0) run(u, maxl + 1);
    return maxl + 1;
}

// chain length of the separator block when it is factored as a dense block of s tiles by one cluster: one step per block column,
// plus the tile-pair updates that do not hide behind the look-ahead potrf (4 rounds of 7 pairs do)
double dense_chain_cost(int s) {
    double c = 0.0;
    for (int j = 0; j < s; ++j) {
        const int m = s - j - 1;
        const int rounds = (m * (m + 1) / 2 + 6) / 7;
        c += 1.0 + 0.16 * std::max(0, rounds - 4);
    }
    return c;
}

struct Candidate {
    double cost = 1e300;
    std::vector<char> is_sep_level;
    std::vector<int> comp;                  // component id per group (-1 = separator)
    std::vector<int> bin_of_comp;
    int nbins = 0, sep_tiles = 0, max_bin_tiles = 0, nsep_levels = 0;
};

// connected components of the groups outside the separator levels
struct Components {
    std::vector<int> comp;                  // component id per group (-1 = separator)
    std::vector<int> comp_vars;
    int sep_vars = 0, nsep_levels = 0;
};
void find_components(const Graph& g, const int* gsize, const std::vector<int>& level, const std::vector<char>& sep_level, Components& c, std::vector<int>& queue) {
    c.comp.assign(g.G, -1);
    c.comp_vars.clear();
    c.sep_vars = 0; c.nsep_levels = 0;
    queue.resize(g.G);
    int ncomp = 0;
    for (int u = 0; u < g.G; ++u) {
        if (sep_level[level[u]]) { c.sep_vars += gsize[u]; continue; }
        if (c.comp[u] >= 0) continue;
        int head = 0, tail = 0, vars = 0;
        c.comp[u] = ncomp; queue[tail++] = u;
        while (head < tail) {
            const int a = queue[head++];
            vars += gsize[a];
            for (int b : g.nb[a]) if (c.comp[b] < 0 && !sep_level[level[b]]) { c.comp[b] = ncomp; queue[tail++] = b; }
        }
        c.comp_vars.push_back(vars);
        ++ncomp;
    }
    for (char sl : sep_level) c.nsep_levels += sl ? 1 : 0;
}
// longest-processing-time packing of the components into at most max_bins bins, and the chain-length cost of the result
// (comp / is_sep_level are filled in by the caller for a candidate that wins: they are copies)
Candidate pack_components(const Components& cs, int max_bins) {
    Candidate c;
    const int ncomp = (int)cs.comp_vars.size();
    c.nsep_levels = cs.nsep_levels;
    if (ncomp < 2) return c;
    const int nbins = std::min(ncomp, max_bins);
    std::vector<int> order(ncomp);
    std::iota(order.begin(), order.end(), 0);
    std::sort(order.begin(), order.end(), [&](int a, int b) { return cs.comp_vars[a] != cs.comp_vars[b] ? cs.comp_vars[a] > cs.comp_vars[b] : a < b; });
    std::vector<int> bin_vars(nbins, 0);
    c.bin_of_comp.assign(ncomp, 0);
    for (int ci : order) {
        int best = 0;
        for (int b = 1; b < nbins; ++b) if (bin_vars[b] < bin_vars[best]) best = b;
        c.bin_of_comp[ci] = best;
        bin_vars[best] += cs.comp_vars[ci];
    }
    c.nbins = nbins;
    int maxb = 0;
    for (int b = 0; b < nbins; ++b) maxb = std::max(maxb, (bin_vars[b] + kTile - 1) / kTile);
    c.max_bin_tiles = maxb;
    c.sep_tiles = (cs.sep_vars + kTile - 1) / kTile;
    c.cost = (double)maxb + dense_chain_cost(c.sep_tiles);
    return c;
}

}  // namespace

SolveOrder build_solve_order(int G, const int* gsize, const unsigned char* adj) {
    SolveOrder o;
    int n = 0;
    for (int g = 0; g < G; ++g) n += gsize[g];
    o.n = n;
    const int nblk0 = (n + kTile - 1) / kTile;
    if (G < 8 || nblk0 < 24) return o;
    Graph g; g.G = G; g.nb.resize(G);
    for (int a = 0; a < G; ++a) {          // rows are mostly zero: test eight cells at a time
        const unsigned char* row = adj + (size_t)a * G;
        int b = 0;
        for (; b + 8 <= G; b += 8) {
            uint64_t w; memcpy(&w, row + b, 8);
            if (w == 0) continue;
            for (int t = b; t < b + 8; ++t) if (row[t] != 0 && t != a) { g.nb[a].push_back(t); }
        }
        for (; b < G; ++b) if (row[b] != 0 && b != a) g.nb[a].push_back(b);
    }
    for (int a = 0; a < G; ++a)            // symmetrise (the device kernel marks both orders; a caller's matrix may hold one triangle)
        for (int b : g.nb[a]) if (adj[(size_t)b * G + a] == 0) g.nb[b].push_back(a);
    for (int a = 0; a < G; ++a) { std::sort(g.nb[a].begin(), g.nb[a].end()); g.nb[a].erase(std::unique(g.nb[a].begin(), g.nb[a].end()), g.nb[a].end()); }
    // pseudo-peripheral root: the farthest, lowest-degree group of a few successive BFS runs
    std::vector<int> level;
    int root = 0, nlevels = 0;
    for (int it = 0; it < 3; ++it) {
        nlevels = bfs_levels(g, root, level);
        int far = root;
        for (int u = 0; u < G; ++u) {
            if (level[u] > level[far] || (level[u] == level[far] && g.nb[u].size() < g.nb[far].size())) far = u;
        }
        if (it < 2) root = far;
    }
    nlevels = bfs_levels(g, root, level);
    o.levels = nlevels;
    if (nlevels < 5) return o;
    std::vector<int64_t> cum(nlevels + 1, 0);
    {
        std::vector<int> w(nlevels, 0);
        for (int u = 0; u < G; ++u) w[level[u]] += gsize[u];
        for (int l = 0; l < nlevels; ++l) cum[l + 1] = cum[l] + w[l];
    }
    // candidates: q whole levels as separators, spaced evenly in unknowns (scheme 0) or with half-width end gaps (scheme 1: on a ring
    // the end gaps are one component each, the inner gaps split into two)
    // One-level candidates fill up to kMaxConcurrentParts bins; two-level candidates half as many, because every bin becomes two leaves.
    static int two_level = -1;      // SRK_SOLVE_LEVELS=1 keeps the one-level order (development aid)
    if (two_level < 0) { const char* e = getenv("SRK_SOLVE_LEVELS"); two_level = (e != nullptr && e[0] == '1') ? 0 : 1; }
    std::vector<int> by_level(G);
    std::iota(by_level.begin(), by_level.end(), 0);
    std::stable_sort(by_level.begin(), by_level.end(), [&](int a, int b) { return level[a] < level[b]; });
    struct Piece { std::vector<int> groups; int vars = 0; };
    struct Plan { std::vector<Piece> leaves, mids; int chain = 0; bool cut = false; };
    // second level: a bin that holds ONE component is cut at the BFS level that halves its unknowns.  Edges join neighbouring levels
    // only, so the level separates what lies below it from what lies above: two leaves and a second-level separator that only those
    // two leaves (and the top separator) touch.  The chain of a bin of b tiles shrinks from b to ~b/2 + (tiles of one level).
    auto plan_of = [&](const Candidate& c, bool allow_cut) {
        Plan pl;
        const int ncomp = (int)c.bin_of_comp.size();
        std::vector<std::vector<int>> bin_groups(c.nbins);
        std::vector<int> bin_ncomp(c.nbins, 0), bin_vars(c.nbins, 0);
        for (int b = 0; b < c.nbins; ++b)
            for (int ci = 0; ci < ncomp; ++ci) {
                if (c.bin_of_comp[ci] != b) continue;
                ++bin_ncomp[b];
                for (int u : by_level) if (c.comp[u] == ci) { bin_groups[b].push_back(u); bin_vars[b] += gsize[u]; }
            }
        auto tiles = [](int v) { return (v + kTile - 1) / kTile; };
        std::vector<int> cut_level(c.nbins, -1), cut_chain(c.nbins, 0);
        int nleaves = c.nbins, chain_before = 0, chain_after = 0;
        std::vector<int> cand(c.nbins);
        std::iota(cand.begin(), cand.end(), 0);
        std::stable_sort(cand.begin(), cand.end(), [&](int a, int b) { return bin_vars[a] > bin_vars[b]; });
        for (int b : cand) {
            cut_chain[b] = tiles(bin_vars[b]);
            if (!allow_cut || nleaves + 1 > kMaxConcurrentParts || bin_ncomp[b] != 1 || bin_groups[b].empty()) continue;
            const int lo = level[bin_groups[b].front()], hi = level[bin_groups[b].back()];
            if (hi - lo < 2) continue;
            std::vector<int> w(hi - lo + 1, 0);
            for (int u : bin_groups[b]) w[level[u] - lo] += gsize[u];
            int best_l = -1, best_worst = 0x7fffffff, below = w[0];
            for (int l = lo + 1; l < hi; ++l) {
                const int above = bin_vars[b] - below - w[l - lo];
                const int worst = std::max(below, above);
                if (below > 0 && above > 0 && worst < best_worst) { best_worst = worst; best_l = l; }
                below += w[l - lo];
            }
            if (best_l < 0) continue;
            const int cut_tiles = tiles(best_worst) + tiles(w[best_l - lo]);
            if (cut_tiles + 2 > tiles(bin_vars[b])) continue;             // nothing to gain
            cut_level[b] = best_l;
            cut_chain[b] = cut_tiles;
            ++nleaves;
        }
        // the chain is the longest bin: cuts pay only when they shorten THAT by two steps or more, and only bins longer than the new
        // longest one need cutting (every cut costs padding)
        for (int b = 0; b < c.nbins; ++b) { chain_before = std::max(chain_before, tiles(bin_vars[b])); chain_after = std::max(chain_after, cut_chain[b]); }
        for (int b = 0; b < c.nbins; ++b)
            if (chain_after + 2 > chain_before || tiles(bin_vars[b]) <= chain_after) cut_level[b] = -1;
        pl.chain = chain_after + 2 > chain_before ? chain_before : chain_after;
        for (int b = 0; b < c.nbins; ++b) {
            if (cut_level[b] < 0) { Piece p; p.groups = bin_groups[b]; p.vars = bin_vars[b]; pl.leaves.push_back(std::move(p)); continue; }
            Piece lo_p, hi_p, mid_p;
            for (int u : bin_groups[b]) {
                Piece& dst = level[u] < cut_level[b] ? lo_p : (level[u] == cut_level[b] ? mid_p : hi_p);
                dst.groups.push_back(u); dst.vars += gsize[u];
            }
            pl.leaves.push_back(std::move(lo_p)); pl.leaves.push_back(std::move(hi_p)); pl.mids.push_back(std::move(mid_p));
            pl.cut = true;
        }
        return pl;
    };
    Candidate best, best2;
    Components comps;
    std::vector<int> queue;
    const int qmax = std::min(24, (nlevels - 1) / 2);
    for (int q = 1; q <= qmax; ++q) {
        for (int scheme = 0; scheme < 2; ++scheme) {
            std::vector<char> sep(nlevels, 0);
            int placed = 0;
            for (int i = 1; i <= q; ++i) {
                const double frac = scheme == 0 ? (double)i / (q + 1) : ((double)i - 0.5) / q;
                const int64_t target = (int64_t)(frac * (double)n);
                int l = (int)(std::upper_bound(cum.begin(), cum.end(), target) - cum.begin()) - 1;
                l = std::max(1, std::min(nlevels - 2, l));
                while (l < nlevels - 2 && sep[l]) ++l;
                if (!sep[l]) { sep[l] = 1; ++placed; }
            }
            if (placed == 0) continue;
            find_components(g, gsize, level, sep, comps, queue);
            if (comps.comp_vars.size() < 2) continue;
            Candidate c = pack_components(comps, kMaxConcurrentParts);
            if (c.cost < best.cost) { best = std::move(c); best.comp = comps.comp; best.is_sep_level = sep; }
            if (two_level) {
                Candidate c2 = pack_components(comps, kMaxConcurrentParts / 2);
                if (c2.nbins >= 2) {        // estimate: every bin halves, plus a second-level separator of about two tiles
                    c2.cost = (double)((c2.max_bin_tiles + 1) / 2 + 2) + dense_chain_cost(c2.sep_tiles);
                    if (c2.cost < best2.cost) { best2 = std::move(c2); best2.comp = comps.comp; best2.is_sep_level = sep; }
                }
            }
        }
    }
    Plan plan;
    bool have_plan = false;
    if (two_level && best2.nbins >= 2) {        // the estimate has to survive the real cut
        Plan p2 = plan_of(best2, true);
        const double cost2 = (double)p2.chain + dense_chain_cost(best2.sep_tiles);
        if (p2.cut && cost2 + 1.0 < best.cost) { best = std::move(best2); best.cost = cost2; plan = std::move(p2); have_plan = true; }
    }
    if (!(best.cost < 0.75 * (double)nblk0) || best.nbins < 2) return o;
    if (!have_plan) plan = plan_of(best, false);
    const std::vector<Piece>& leaves = plan.leaves;
    const std::vector<Piece>& mids = plan.mids;

    // positions: leaves in order (inside a leaf components in id order, inside a component groups by (level, index)), then the
    // second-level separators, the top separator last; every piece starts on a 64-column boundary
    std::vector<int> gstart(G + 1, 0);
    for (int u = 0; u < G; ++u) gstart[u + 1] = gstart[u] + gsize[u];
    o.pos.assign(n, -1);
    int cursor = 0;
    auto place = [&](int u) { for (int a = 0; a < gsize[u]; ++a) o.pos[gstart[u] + a] = cursor++; };
    o.part.nparts = 0;
    for (const Piece& p : leaves) {
        const int k0 = cursor / kTile;
        for (int u : p.groups) place(u);
        cursor = (cursor + kTile - 1) / kTile * kTile;
        const int k1 = cursor / kTile;
        if (k1 > k0) { o.part.k0[o.part.nparts] = k0; o.part.k1[o.part.nparts] = k1; ++o.part.nparts; o.max_part_blocks = std::max(o.max_part_blocks, k1 - k0); }
    }
    o.part.ksep = cursor / kTile;
    o.part.nmids = 0;
    for (const Piece& p : mids) {
        const int k0 = cursor / kTile;
        for (int u : p.groups) place(u);
        cursor = (cursor + kTile - 1) / kTile * kTile;
        const int k1 = cursor / kTile;
        if (k1 > k0) { o.part.m0[o.part.nmids] = k0; o.part.m1[o.part.nmids] = k1; ++o.part.nmids; o.max_mid_blocks = std::max(o.max_mid_blocks, k1 - k0); }
    }
    o.part.msep = cursor / kTile;
    for (int u : by_level) if (best.comp[u] < 0) place(u);
    o.np = cursor;
    o.src.assign(o.np, -1);
    for (int i = 0; i < n; ++i) o.src[o.pos[i]] = i;
    o.sep_levels = best.nsep_levels;
    o.sep_blocks = (o.np + kTile - 1) / kTile - o.part.msep;
    o.active = o.part.nparts >= 2;
    if (!o.active) return o;

    // ---- tile structure of the natural and of the ordered system
    const int nb0 = nblk0, nb1 = (o.np + kTile - 1) / kTile;
    std::vector<unsigned char> sm((size_t)nb0 * nb0, 0), lm((size_t)nb1 * nb1, 0);   // [c*nb + r], r >= c
    auto mark_pair = [&](int a, int b) {       // every tile the block of groups (a, b) touches, mirrored into the lower triangle
        const int a0 = gstart[a], a1 = gstart[a + 1] - 1, b0 = gstart[b], b1 = gstart[b + 1] - 1;
        for (int R = a0 / kTile; R <= a1 / kTile; ++R)
            for (int C = b0 / kTile; C <= b1 / kTile; ++C) { const int r = std::max(R, C), c = std::min(R, C); sm[(size_t)c * nb0 + r] = 1; }
        // ordered positions of a group are consecutive
        const int pa0 = o.pos[a0], pa1 = o.pos[a1], pb0 = o.pos[b0], pb1 = o.pos[b1];
        for (int R = pa0 / kTile; R <= pa1 / kTile; ++R)
            for (int C = pb0 / kTile; C <= pb1 / kTile; ++C) { const int r = std::max(R, C), c = std::min(R, C); lm[(size_t)c * nb1 + r] = 1; }
    };
    for (int a = 0; a < G; ++a) { mark_pair(a, a); for (int b : g.nb[a]) if (b < a) mark_pair(a, b); }
    for (int k = 0; k < nb1; ++k) lm[(size_t)k * nb1 + k] = 1;                          // padding unknowns sit on the diagonal
    for (int c = 0; c < nb0; ++c) for (int r = c; r < nb0; ++r) if (sm[(size_t)c * nb0 + r]) o.s_tiles.push_back((r << 16) | c);
    o.res_ptr.assign(nb0 + 1, 0);
    for (int R = 0; R < nb0; ++R) {
        for (int C = 0; C <= R; ++C) if (sm[(size_t)C * nb0 + R]) o.res_ent.push_back((C << 1) | 0);
        for (int R2 = R + 1; R2 < nb0; ++R2) if (sm[(size_t)R * nb0 + R2]) o.res_ent.push_back((R2 << 1) | 1);
        o.res_ptr[R + 1] = (int)o.res_ent.size();
    }
    o.l_pattern.assign((size_t)nb1 * nb1, 0);
    for (int c = 0; c < nb1; ++c)
        for (int r = c; r < nb1; ++r)
            if (lm[(size_t)c * nb1 + r]) { o.l_in_tiles.push_back((r << 16) | c); if (r > c) { o.l_pattern[(size_t)c * nb1 + r] = 1; ++o.l_pattern_count; } }
    // symbolic factorisation, tile level: column k's rows fill each other's tiles
    std::vector<int> rows;
    for (int k = 0; k < nb1; ++k) {
        rows.clear();
        for (int r = k + 1; r < nb1; ++r) if (lm[(size_t)k * nb1 + r]) rows.push_back(r);
        for (size_t i = 0; i < rows.size(); ++i) for (size_t j = 0; j <= i; ++j) lm[(size_t)rows[j] * nb1 + rows[i]] = 1;
    }
    for (int c = 0; c < nb1; ++c) for (int r = c; r < nb1; ++r) if (lm[(size_t)c * nb1 + r]) o.l_all_tiles.push_back((r << 16) | c);
    return o;
}

// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_permute_sym(int n, const double* __restrict__ S, int64_t ld, int mirrored, int np, const int* __restrict__ src,
                                                     double* __restrict__ L, int64_t ldp) {
    const int jd = blockIdx.x;
    const int js = src[jd];
    double* out = L + (size_t)jd * ldp;
    if (js < 0) {
        for (int id = jd + threadIdx.x; id < np; id += 256) out[id] = id == jd ? 1.0 : 0.0;
        return;
    }
    const double* col = S + (size_t)js * ld;
    for (int id = jd + threadIdx.x; id < np; id += 256) {
        const int is = src[id];
        double v = 0.0;
        if (is >= 0) v = (mirrored || is >= js) ? col[is] : S[(size_t)is * ld + js];
        out[id] = v;
    }
}
__global__ void __launch_bounds__(256) k_zero_tiles(int n, double* __restrict__ A, int64_t ld, const int* __restrict__ tiles) {
    const int t = tiles[blockIdx.x];
    const int r0 = (t >> 16) * kTile, c0 = (t & 0xffff) * kTile;
    for (int e = threadIdx.x; e < kTile * kTile; e += 256) {
        const int r = r0 + (e & 63), c = c0 + (e >> 6);
        if (r < n && c < n) A[(size_t)c * ld + r] = 0.0;
    }
}
__global__ void __launch_bounds__(256) k_permute_tiles(int n, const double* __restrict__ S, int64_t ld, int np, const int* __restrict__ src, double* __restrict__ L,
                                                       int64_t ldp, const int* __restrict__ tiles) {
    __shared__ int srow[kTile], scol[kTile];
    const int t = tiles[blockIdx.x];
    const int r0 = (t >> 16) * kTile, c0 = (t & 0xffff) * kTile;
    if (threadIdx.x < kTile) srow[threadIdx.x] = r0 + threadIdx.x < np ? src[r0 + threadIdx.x] : -2;
    else if (threadIdx.x < 2 * kTile) scol[threadIdx.x - kTile] = c0 + threadIdx.x - kTile < np ? src[c0 + threadIdx.x - kTile] : -2;
    __syncthreads();
    for (int e = threadIdx.x; e < kTile * kTile; e += 256) {
        const int rr = e & 63, cc = e >> 6;
        const int id = r0 + rr, jd = c0 + cc;
        const int is = srow[rr], js = scol[cc];
        if (is == -2 || js == -2 || id < jd) continue;
        double v;
        if (is < 0 || js < 0) v = id == jd ? 1.0 : 0.0;
        else v = is >= js ? S[(size_t)js * ld + is] : S[(size_t)is * ld + js];
        L[(size_t)jd * ldp + id] = v;
    }
}
__device__ __forceinline__ void dd_acc(double& hi, double& lo, double bh, double bl) {   // (hi, lo) += (bh, bl), two-sum
    double s = __dadd_rn(hi, bh);
    double bb = __dadd_rn(s, -hi);
    double e = __dadd_rn(__dadd_rn(hi, -__dadd_rn(s, -bb)), __dadd_rn(bh, -bb));
    e = __dadd_rn(e, __dadd_rn(lo, bl));
    hi = __dadd_rn(s, e);
    lo = __dadd_rn(e, -__dadd_rn(hi, -s));
}
// one CTA per row block R; thread (row = tid & 63, q = tid >> 6) sums the columns c == q (mod 4) of its row over the block's tiles
__global__ void __launch_bounds__(256) k_residual_dd_tiles(int n, const double* __restrict__ S, int64_t ld, const double* __restrict__ x, const double* __restrict__ b,
                                                           double* __restrict__ r, const int* __restrict__ res_ptr, const int* __restrict__ res_ent) {
    __shared__ double xs[kTile];
    __shared__ double ph[4][kTile], pl[4][kTile];
    const int R = blockIdx.x, row = threadIdx.x & 63, q = threadIdx.x >> 6;
    const int i = R * kTile + row;
    double hi = 0.0, lo = 0.0;
    for (int t = res_ptr[R]; t < res_ptr[R + 1]; ++t) {
        const int ent = res_ent[t];
        const int other = ent >> 1, tr = ent & 1;
        __syncthreads();
        if (threadIdx.x < kTile) xs[threadIdx.x] = other * kTile + threadIdx.x < n ? x[other * kTile + threadIdx.x] : 0.0;
        __syncthreads();
        if (i >= n) continue;
#pragma unroll 4
        for (int cc = q; cc < kTile; cc += 4) {
            const int j = other * kTile + cc;
            if (j >= n) break;
            // tile (R, other): S(i, j) lower as stored;  transposed tile (other, R): S(i, j) = S(j, i), j > i
            double a;
            if (!tr) { if (other == R && j > i) a = S[(size_t)i * ld + j]; else a = S[(size_t)j * ld + i]; }
            else a = S[(size_t)i * ld + j];
            const double xv = xs[cc];
            const double p = __dmul_rn(a, xv);
            const double e = __fma_rn(a, xv, -p);
            dd_acc(hi, lo, p, e);
        }
    }
    ph[q][row] = hi; pl[q][row] = lo;
    __syncthreads();
    if (q == 0 && i < n) {
        for (int k = 1; k < 4; ++k) dd_acc(hi, lo, ph[k][row], pl[k][row]);
        double rh = b[i], rl = 0.0;
        dd_acc(rh, rl, -hi, -lo);
        r[i] = rh;
    }
}
__global__ void k_gather_vec(int np, const int* __restrict__ src, const double* __restrict__ in, double* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < np) { const int s = src[i]; out[i] = s >= 0 ? in[s] : 0.0; }
}
__global__ void k_scatter_vec(int np, const int* __restrict__ src, const double* __restrict__ in, double* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < np) { const int s = src[i]; if (s >= 0) out[s] = in[i]; }
}
// one thread per point: every pair of its cameras is marked (both orders); the test before the store keeps the write traffic
// to the first few points of a pair
__global__ void k_cam_adjacency(int64_t N, const int64_t* __restrict__ pt_begin, const int32_t* __restrict__ obs_cam, int M, unsigned char* adj) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= N) return;
    const int64_t b = pt_begin[j], e = pt_begin[j + 1];
    for (int64_t a = b; a < e; ++a) {
        const int ca = obs_cam[a];
        for (int64_t c = b; c < a; ++c) {
            const int cc = obs_cam[c];
            unsigned char* p = adj + (size_t)ca * M + cc;
            if (*p == 0) { *p = 1; adj[(size_t)cc * M + ca] = 1; }
        }
    }
}

void launch_permute_sym(cudaStream_t st, int n, const double* S, int64_t ld, int mirrored, int np, const int* src_dev, double* L, int64_t ldp) {
    if (np > 0) k_permute_sym<<<np, 256, 0, st>>>(n, S, ld, mirrored, np, src_dev, L, ldp);
}
void launch_zero_tiles(cudaStream_t st, int n, double* A, int64_t ld, const int* tiles, int count) {
    if (count > 0) k_zero_tiles<<<count, 256, 0, st>>>(n, A, ld, tiles);
}
void launch_permute_tiles(cudaStream_t st, int n, const double* S, int64_t ld, int np, const int* src_dev, double* L, int64_t ldp, const int* tiles, int count) {
    if (count > 0) k_permute_tiles<<<count, 256, 0, st>>>(n, S, ld, np, src_dev, L, ldp, tiles);
}
void launch_residual_dd_tiles(cudaStream_t st, int n, const double* S, int64_t ld, const double* x, const double* b, double* r, const int* res_ptr, const int* res_ent) {
    const int nb = (n + kTile - 1) / kTile;
    if (nb > 0) k_residual_dd_tiles<<<nb, 256, 0, st>>>(n, S, ld, x, b, r, res_ptr, res_ent);
}
void launch_gather_vec(cudaStream_t st, int np, const int* src_dev, const double* in, double* out) {
    if (np > 0) k_gather_vec<<<(np + 255) / 256, 256, 0, st>>>(np, src_dev, in, out);
}
void launch_scatter_vec(cudaStream_t st, int np, const int* src_dev, const double* in, double* out) {
    if (np > 0) k_scatter_vec<<<(np + 255) / 256, 256, 0, st>>>(np, src_dev, in, out);
}
__global__ void k_bytes_to_doubles(int64_t n, const unsigned char* __restrict__ in, double* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[i] != 0 ? 1.0 : 0.0;
}
__global__ void k_doubles_to_bytes(int64_t n, const double* __restrict__ in, unsigned char* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[i] != 0.0 ? 1 : 0;
}
void launch_cam_adjacency(cudaStream_t st, int64_t N, const int64_t* pt_begin, const int32_t* obs_cam, int M, unsigned char* adj) {
    if (N > 0) k_cam_adjacency<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(N, pt_begin, obs_cam, M, adj);
}
void launch_bytes_to_doubles(cudaStream_t st, int64_t n, const unsigned char* in, double* out) {
    if (n > 0) k_bytes_to_doubles<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, in, out);
}
void launch_doubles_to_bytes(cudaStream_t st, int64_t n, const double* in, unsigned char* out) {
    if (n > 0) k_doubles_to_bytes<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, in, out);
}

}  // namespace srk
