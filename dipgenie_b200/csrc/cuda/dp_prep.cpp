// dp_prep.cpp — see dp_prep.h.
#include "dp_prep.h"

#include <algorithm>

namespace dg {

bool build_dip_plan(const DipGraphView& g, DipPlan& p) {
    p = DipPlan();
    if (g.n_levels < 1 || g.R < 0 || !g.level_off || !g.adj_off) { p.error = "bad arguments"; return false; }
    const int L = g.n_levels;
    p.L = L; p.R = g.R;
    p.level_off.assign(g.level_off, g.level_off + L + 1);
    const int32_t V = p.level_off[L];
    p.V = V;
    if (p.level_off[0] != 0) { p.error = "level_off[0] must be 0"; return false; }
    // approximator.cpp:375-377 assumes a single source on level 0 (dp_cur has R+1 cells, :535)
    if (p.level_off[1] - p.level_off[0] != 1) { p.error = "level 0 must hold exactly one vertex"; return false; }
    for (int l = 0; l < L; ++l) {
        int32_t k = p.level_off[l + 1] - p.level_off[l];
        if (k <= 0) { p.error = "empty level"; return false; }
        if (k >= 65535) { p.error = "level wider than 65534 vertices"; return false; }
        p.kmax = std::max(p.kmax, k);
    }
    if (g.adj_off[0] != 0) { p.error = "adj_off[0] must be 0"; return false; }

    // ---- in-edge CSR (gather form of approximator.cpp:640-649) ----
    p.in_off.assign((size_t)V + 1, 0);
    const int64_t E = g.adj_off[V];
    for (int l = 0; l < L; ++l) {
        const int32_t lo = p.level_off[l], hi = p.level_off[l + 1];
        const int32_t nlo = (l + 1 < L) ? p.level_off[l + 1] : V, nhi = (l + 1 < L) ? p.level_off[l + 2] : V;
        uint64_t El = 0;
        for (int32_t u = lo; u < hi; ++u) {
            if (g.adj_off[u + 1] < g.adj_off[u]) { p.error = "adj_off not monotone"; return false; }
            El += (uint64_t)(g.adj_off[u + 1] - g.adj_off[u]);
            for (int64_t e = g.adj_off[u]; e < g.adj_off[u + 1]; ++e) {
                const int32_t v = g.adj_dst[e];
                if (v < nlo || v >= nhi) { p.error = "edge does not span exactly one level"; return false; }
                ++p.in_off[(size_t)v + 1];
            }
        }
        if (l + 1 < L) p.cell_updates += (uint64_t)(g.R + 1) * El * El;
    }
    for (int32_t v = 0; v < V; ++v) {
        p.max_indeg = std::max(p.max_indeg, p.in_off[(size_t)v + 1]);
        p.in_off[(size_t)v + 1] += p.in_off[v];
    }
    p.n_in = E;
    p.in_edge.assign((size_t)E, 0);
    {
        std::vector<int32_t> fill(p.in_off.begin(), p.in_off.end() - 1);
        for (int l = 0; l + 1 < L; ++l) {
            const int32_t lo = p.level_off[l], hi = p.level_off[l + 1];
            for (int32_t u = lo; u < hi; ++u)
                for (int64_t e = g.adj_off[u]; e < g.adj_off[u + 1]; ++e)
                    p.in_edge[(size_t)fill[g.adj_dst[e]]++] = (uint32_t)(u - lo) | ((uint32_t)g.adj_w[e] << IN_W_SHIFT);
        }
    }

    // ---- per-transition colour masks (approximator.cpp:431-453 + :269-311 as popcounts) ----
    p.lvlW.assign(L, 0); p.msrc_off.assign(L, 0); p.mdst_off.assign(L, 0);
    std::vector<int32_t> local(g.n_colours > 0 ? g.n_colours : 1, -1);
    std::vector<int32_t> uni;
    for (int l = 0; l + 1 < L; ++l) {
        const int32_t lo = p.level_off[l], mid = p.level_off[l + 1], hi = p.level_off[l + 2];
        if (g.col_off[hi] == g.col_off[lo]) continue;   // no colours on either level
        uni.clear();
        for (int64_t c = g.col_off[lo]; c < g.col_off[hi]; ++c) {
            const int32_t col = g.col_val[c];
            if (col < 0 || col >= g.n_colours) { p.error = "colour id out of range"; return false; }
            if (local[col] < 0) { local[col] = 0; uni.push_back(col); }
        }
        std::sort(uni.begin(), uni.end());
        for (size_t x = 0; x < uni.size(); ++x) local[uni[x]] = (int32_t)x;
        const int W = (int)((uni.size() + 63) / 64);
        p.lvlW[l] = W; p.Wmax = std::max(p.Wmax, W);
        p.msrc_off[l] = (int64_t)p.masks.size();
        p.masks.resize(p.masks.size() + (size_t)(mid - lo) * 2 * W, 0);
        p.mdst_off[l] = (int64_t)p.masks.size();
        p.masks.resize(p.masks.size() + (size_t)(hi - mid) * 2 * W, 0);
        for (int32_t v = lo; v < hi; ++v) {
            uint64_t* m = (v < mid) ? &p.masks[(size_t)p.msrc_off[l] + (size_t)(v - lo) * 2 * W]
                                    : &p.masks[(size_t)p.mdst_off[l] + (size_t)(v - mid) * 2 * W];
            for (int64_t c = g.col_off[v]; c < g.col_off[v + 1]; ++c) {
                const int32_t col = g.col_val[c];
                const int b = local[col];
                const int half = (g.colour_is_hom[col] == 1) ? 0 : W;   // hom words first, then het
                m[half + (b >> 6)] |= 1ull << (b & 63);
            }
        }
        for (int32_t col : uni) local[col] = -1;
    }

    // ---- predecessor-code offsets and accounting ----
    p.pred_off.assign((size_t)L + 1, 0);
    for (int l = 0; l < L; ++l) {
        const uint64_t k = (uint64_t)(p.level_off[l + 1] - p.level_off[l]);
        p.pred_off[(size_t)l + 1] = p.pred_off[l] + (int64_t)((uint64_t)(g.R + 1) * k * k);
        if (l >= 1) p.cells += (uint64_t)(g.R + 1) * k * k;
        if (l + 1 < L) {
            const uint64_t k2 = (uint64_t)(p.level_off[l + 2] - p.level_off[l + 1]);
            p.algo_bytes += (uint64_t)(g.R + 1) * (4 * k * k + 5 * k2 * k2);
        }
    }
    return true;
}

void plan_participants(DipPlan& p, int grid, int cells_per_cta) {
    const int L = p.L;
    p.P.assign(L, 1); p.bar_target.assign(L, 0); p.bar_edge.assign(L, 0);
    if (grid < 1) grid = 1;
    if (cells_per_cta < 1) cells_per_cta = 1;
    for (int l = 0; l + 1 < L; ++l) {
        const uint64_t k2 = (uint64_t)(p.level_off[l + 2] - p.level_off[l + 1]);
        const uint64_t cells = (uint64_t)(p.R + 1) * k2 * k2;
        uint64_t want = (cells + (uint64_t)cells_per_cta - 1) / (uint64_t)cells_per_cta;
        p.P[l] = (int32_t)std::min<uint64_t>(std::max<uint64_t>(want, 1), (uint64_t)grid);
    }
    // A grid barrier follows transition l unless both it and the next one run on CTA 0 alone.
    uint32_t acc = 0;
    for (int l = 0; l + 1 < L; ++l) {
        const bool last = (l + 2 >= L);
        const int pn = last ? 1 : p.P[l + 1];
        const bool edge = (p.P[l] > 1) || (pn > 1);
        p.bar_edge[l] = edge ? 1 : 0;
        if (edge) acc += (uint32_t)p.P[l];
        p.bar_target[l] = acc;
    }
}

}  // namespace dg
