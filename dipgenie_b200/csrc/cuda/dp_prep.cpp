// dp_prep.cpp — see dp_prep.h.
#include "dp_prep.h"

#include <algorithm>
#include <cstring>

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

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

void plan_sweep(DipPlan& p, const SweepShape& sh) {
    const int L = p.L;
    const int grid = sh.grid < 1 ? 1 : sh.grid;
    const uint64_t cpc = sh.cells_per_cta < 1 ? 1 : (uint64_t)sh.cells_per_cta;
    p.P.assign(L, 1); p.bar_target.assign(L, 0); p.bar_edge.assign(L, 0);
    p.mode.assign(L, MODE_GLOBAL); p.flags.assign(L, 0); p.rec_off.assign(L, -1);
    p.records.clear(); p.n_fast = p.n_staged = p.n_global = 0;
    const int T = L - 1;                       // number of transitions
    std::vector<size_t> rec_bytes(L, 0);
    auto cells_of = [&](int l) { const uint64_t k = (uint64_t)(p.level_off[l + 1] - p.level_off[l]); return (uint64_t)(p.R + 1) * k * k; };
    for (int l = 0; l < T; ++l) {
        const int32_t k = p.level_off[l + 1] - p.level_off[l], k2 = p.level_off[l + 2] - p.level_off[l + 1];
        const int64_t n_in = (int64_t)p.in_off[p.level_off[l + 2]] - (int64_t)p.in_off[p.level_off[l + 1]];
        const int W = p.lvlW[l];
        size_t b = sizeof(RecHeader) + align_up((size_t)(k2 + 1) * 2, 4);
        b += (size_t)n_in * 4; b = align_up(b, 8);
        b += (size_t)(k + k2) * 2 * W * 8; b = align_up(b, 16);
        rec_bytes[l] = b;
        const bool staged_ok = b <= (size_t)sh.stage_bytes && n_in < 65536;
        const bool fits = cells_of(l) <= (uint64_t)sh.tile_cells && cells_of(l + 1) <= (uint64_t)sh.tile_cells;
        p.mode[l] = !staged_ok ? MODE_GLOBAL : (fits ? MODE_FAST : MODE_STAGED);
        if (p.mode[l] == MODE_FAST) { p.P[l] = 1; ++p.n_fast; }
        else {
            const uint64_t want = (cells_of(l + 1) + cpc - 1) / cpc;
            p.P[l] = (int32_t)std::min<uint64_t>(std::max<uint64_t>(want, 1), (uint64_t)grid);
            if (p.mode[l] == MODE_STAGED) ++p.n_staged; else ++p.n_global;
        }
    }
    // A grid barrier follows transition l unless both it and the next one run on CTA 0 alone.
    uint32_t acc = 0;
    for (int l = 0; l < T; ++l) {
        const int pn = (l + 1 < T) ? p.P[l + 1] : 1;
        const bool edge = (p.P[l] > 1) || (pn > 1);
        p.bar_edge[l] = edge ? 1 : 0;
        if (edge) acc += (uint32_t)p.P[l];
        p.bar_target[l] = acc;
    }
    for (int l = 0; l < T; ++l) {
        uint16_t f = 0;
        if (l > 0 && p.bar_edge[l - 1]) f |= REC_WAIT;
        if (p.bar_edge[l]) f |= REC_ARRIVE;
        // layer l lives in shared memory iff it is produced and consumed by FAST transitions (level 0: by the kernel prologue)
        const bool fast = p.mode[l] == MODE_FAST;
        if (fast && (l == 0 || p.mode[l - 1] == MODE_FAST)) f |= REC_SRC_SMEM;
        if (fast && l + 1 < T && p.mode[l + 1] == MODE_FAST) f |= REC_DST_SMEM;
        p.flags[l] = f;
    }
    // pack records
    size_t total = 0;
    for (int l = 0; l < T; ++l) if (p.mode[l] != MODE_GLOBAL) { p.rec_off[l] = (int64_t)total; total += rec_bytes[l]; }
    p.records.assign(total, 0);
    for (int l = 0; l < T; ++l) {
        if (p.mode[l] == MODE_GLOBAL) continue;
        uint8_t* r = p.records.data() + p.rec_off[l];
        const int32_t mid = p.level_off[l + 1];
        const int32_t k = mid - p.level_off[l], k2 = p.level_off[l + 2] - mid;
        const int32_t e0 = p.in_off[mid], e1 = p.in_off[p.level_off[l + 2]];
        const int W = p.lvlW[l];
        RecHeader h;
        h.k = (uint16_t)k; h.k2 = (uint16_t)k2; h.W = (uint16_t)W; h.flags = p.flags[l];
        h.n_in = (uint32_t)(e1 - e0); h.bytes = (uint32_t)rec_bytes[l];
        h.P = (uint32_t)p.P[l]; h.wait_target = (l > 0) ? p.bar_target[l - 1] : 0;
        h.pred_off2 = p.pred_off[l + 1];
        memcpy(r, &h, sizeof h);
        size_t o = sizeof(RecHeader);
        uint16_t* off2 = reinterpret_cast<uint16_t*>(r + o);
        for (int32_t x = 0; x <= k2; ++x) off2[x] = (uint16_t)(p.in_off[mid + x] - e0);
        o += align_up((size_t)(k2 + 1) * 2, 4);
        if (e1 > e0) memcpy(r + o, &p.in_edge[e0], (size_t)(e1 - e0) * 4);
        o += (size_t)(e1 - e0) * 4; o = align_up(o, 8);
        if (W > 0) {
            memcpy(r + o, &p.masks[(size_t)p.msrc_off[l]], (size_t)k * 2 * W * 8);
            o += (size_t)k * 2 * W * 8;
            memcpy(r + o, &p.masks[(size_t)p.mdst_off[l]], (size_t)k2 * 2 * W * 8);
        }
    }
}

}  // namespace dg
