// dp_prep.cpp — see dp_prep.h.  Levels are independent for everything but a few prefix sums, so the
// planner runs level-parallel on the host cores (OpenMP; compiles and runs serially without it).
#include "dp_prep.h"

#include <algorithm>
#include <atomic>
#include <cstring>

#if defined(_OPENMP)
#include <omp.h>
#endif

namespace dg {

namespace {
int n_threads() {
#if defined(_OPENMP)
    return std::max(1, std::min(omp_get_max_threads(), 32));
#else
    return 1;
#endif
}
size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
}  // namespace

bool build_dip_plan(const DipGraphView& g, DipPlan& p) {
    p = DipPlan();
    if (g.n_levels < 1 || g.R < 0 || !g.level_off || !g.adj_off || !g.col_off) { p.error = "bad arguments"; return false; }
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
    if (g.adj_off[V] < 0 || (g.adj_off[V] > 0 && (!g.adj_dst || !g.adj_w))) { p.error = "adj_dst / adj_w missing"; return false; }
    if (g.col_off[0] != 0 || g.col_off[V] < 0 || (g.col_off[V] > 0 && (!g.col_val || !g.colour_is_hom))) { p.error = "col_val / colour_is_hom missing"; return false; }
    const int NT = n_threads();
    (void)NT;
    std::atomic<int> bad(0);      // 1 adj_off, 2 edge span, 3 colour id

    // ---- in-edge CSR (gather form of approximator.cpp:640-649) ----
    // The in-edges of level l+1 are the out-edges of level l and nothing else, so the CSR of level l+1 starts at
    // adj_off[level_off[l]] and every level is counted, prefixed and filled on its own — one pass over a level's edges to count,
    // one (cache-hot) to scatter, no global prefix sum.  Sources are visited in ascending position: the lists come out in the
    // tie-break order (smaller i first).
    const int64_t E = g.adj_off[V];
    p.in_off.assign((size_t)V + 1, 0);
    p.n_in = E;
    p.in_edge.assign((size_t)E, 0);
    p.in_dst.assign((size_t)E, 0);
    uint64_t cell_updates = 0;
    int32_t max_indeg = 0;
    std::atomic<int> mixed(0);
    for (int32_t u = 0; u < V && !bad; ++u) if (g.adj_off[u + 1] < g.adj_off[u]) bad = 1;
    if (bad == 1) { p.error = "adj_off not monotone"; return false; }
    if (E > 0x7FFFFFFF) { p.error = "more than 2^31 edges"; return false; }
#pragma omp parallel num_threads(NT) reduction(+ : cell_updates) reduction(max : max_indeg)
    {
        std::vector<int32_t> cur;
#pragma omp for schedule(static)
        for (int l = 0; l < L; ++l) {
            const int32_t lo = p.level_off[l], mid = p.level_off[l + 1];
            const int64_t e0 = g.adj_off[lo], e1 = g.adj_off[mid];
            if (l + 1 == L) { if (e1 != e0) bad = 2; continue; }          // (an edge out of the last level spans no level)
            const int32_t hi = p.level_off[l + 2], k2 = hi - mid;
            cur.assign((size_t)k2 + 1, 0);
            bool ok = true;
            for (int64_t e = e0; e < e1; ++e) {
                const int32_t v = g.adj_dst[e];
                if (v < mid || v >= hi) { bad = 2; ok = false; break; }
                if (g.adj_w[e] > 1) { bad = 4; ok = false; break; }        // weights are 0 / 1 (ExpandedGraph: lane edge / recombination edge)
                ++cur[(size_t)(v - mid) + 1];
            }
            if (!ok) continue;
            int32_t run = (int32_t)e0;
            for (int32_t j = 0; j < k2; ++j) {
                const int32_t c = cur[(size_t)j + 1];
                max_indeg = std::max(max_indeg, c);
                p.in_off[(size_t)mid + j] = run;
                cur[(size_t)j] = run;                 // cursor of destination j
                run += c;
            }
            for (int32_t u = lo; u < mid; ++u)
                for (int64_t e = g.adj_off[u]; e < g.adj_off[u + 1]; ++e) {
                    const int32_t j = g.adj_dst[e] - mid;
                    const size_t slot = (size_t)cur[(size_t)j]++;
                    const uint32_t x = (uint32_t)(u - lo) | ((uint32_t)g.adj_w[e] << IN_W_SHIFT);
                    // parallel edges u -> v sit next to each other in v's list (sources ascend): differing weights are refused below
                    if (slot > (size_t)p.in_off[(size_t)mid + j] && ((p.in_edge[slot - 1] ^ x) & IN_POS_MASK) == 0 && p.in_edge[slot - 1] != x) mixed = 1;
                    p.in_edge[slot] = x;
                    p.in_dst[slot] = (uint16_t)j;
                }
            const uint64_t El = (uint64_t)(e1 - e0);
            cell_updates += (uint64_t)(g.R + 1) * El * El;
        }
    }
    p.in_off[(size_t)V] = (int32_t)E;
    if (bad == 2) { p.error = "edge does not span exactly one level"; return false; }
    if (bad == 4) { p.error = "edge weight above 1"; return false; }
    p.cell_updates = cell_updates;
    p.max_indeg = max_indeg;

    // Parallel edges u -> v of DIFFERENT weight: their candidates reach one destination cell from two layers of the same
    // source pair; on a tie the reference keeps whichever its racing relax loop (approximator.cpp:627-701, `omp for
    // collapse(3) schedule(guided)`) writes first — the result depends on thread timing, there is nothing to be
    // bit-exact with.  (The reference's own front end never builds such edges; same-weight duplicates are fine: identical
    // candidates.)  Found while the lists were filled.
    if (mixed) { p.error = "parallel edges of differing weight between one vertex pair (the reference's tie-break is thread-order dependent for them)"; return false; }

    // ---- per-transition colour masks (approximator.cpp:431-453 + :269-311 as popcounts) ----
    // pass 1: size of the local colour universe of every transition; prefix; pass 2: fill
    p.lvlW.assign(L, 0); p.msrc_off.assign(L, 0); p.mdst_off.assign(L, 0);
    const size_t ncol = (size_t)(g.n_colours > 0 ? g.n_colours : 1);
    int Wmax = 0;
    int64_t vbound = 0;
#pragma omp parallel num_threads(NT) reduction(max : Wmax) reduction(+ : vbound)
    {
        std::vector<int32_t> seen(ncol, -1);
#pragma omp for schedule(static)
        for (int l = 0; l < L - 1; ++l) {
            const int32_t lo = p.level_off[l], hi = p.level_off[l + 2];
            if (g.col_off[hi] == g.col_off[lo]) continue;   // no colours on either level
            int n = 0;
            for (int64_t c = g.col_off[lo]; c < g.col_off[hi]; ++c) {
                const int32_t col = g.col_val[c];
                if (col < 0 || col >= g.n_colours) { bad = 3; break; }
                if (seen[col] != l) { seen[col] = l; ++n; }
            }
            const int W = (n + 63) / 64;
            p.lvlW[l] = W;
            Wmax = std::max(Wmax, W);
            vbound += n;          // delta = |hom ∩| + |het △| <= colours present on the two levels
        }
    }
    if (bad == 3) { p.error = "colour id out of range"; return false; }
    p.Wmax = Wmax;
    p.value_bound = vbound;
    size_t words = 0;
    for (int l = 0; l + 1 < L; ++l) {
        const size_t W = (size_t)p.lvlW[l];
        if (!W) continue;
        const int32_t lo = p.level_off[l], mid = p.level_off[l + 1], hi = p.level_off[l + 2];
        p.msrc_off[l] = (int64_t)words; words += (size_t)(mid - lo) * 2 * W;
        p.mdst_off[l] = (int64_t)words; words += (size_t)(hi - mid) * 2 * W;
    }
    p.masks.assign(words, 0);
#pragma omp parallel num_threads(NT)
    {
        std::vector<int32_t> local(ncol, -1);
        std::vector<int32_t> uni;
#pragma omp for schedule(static)
        for (int l = 0; l < L - 1; ++l) {
            const int W = p.lvlW[l];
            if (!W) continue;
            const int32_t lo = p.level_off[l], mid = p.level_off[l + 1], hi = p.level_off[l + 2];
            uni.clear();
            for (int64_t c = g.col_off[lo]; c < g.col_off[hi]; ++c) {
                const int32_t col = g.col_val[c];
                if (local[col] < 0) { local[col] = 0; uni.push_back(col); }
            }
            // (bit positions in first-seen order: the masks only ever meet in popcounts of AND / XOR within this transition, any
            // bijection of its colours onto bits gives the same pair scores)
            for (size_t x = 0; x < uni.size(); ++x) local[uni[x]] = (int32_t)x;
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

namespace {

// Lane-form blocks of level l+1: starts of the <= 32-wide in-edge blocks followed by n_in.  A block never cuts
// the group of a destination with <= 32 in-edges; a destination with more (allowed only with `allow_long`, at
// most LANE_MAX_LONG per level) is cut into slice blocks of its own.  Returns the number of blocks, 0 when the
// level cannot use the lane form.  `out` / `long_j` may be null (count only); *rounds = ceil(log2(longest group
// of <= 32)), *n_long = destinations with more than 32 in-edges.
int lane_blocks(const DipPlan& p, int l, bool allow_long, uint16_t* out, uint16_t* long_j, int* rounds, int* n_long) {
    const int32_t mid = p.level_off[l + 1], k2 = p.level_off[l + 2] - mid;
    const int32_t e0 = p.in_off[mid];
    const int64_t n_in = (int64_t)p.in_off[mid + k2] - e0;
    if (n_long) *n_long = 0;
    if (n_in <= 0 || n_in >= 65536) return 0;
    int32_t longest = 0, start = 0;
    int nb = 0, nl = 0;
    if (out) out[0] = 0;
    auto cut = [&](int32_t at) {                 // a block boundary at in-edge `at` (ignored if the open block is empty)
        if (at > start) { ++nb; if (out) out[nb] = (uint16_t)at; start = at; }
    };
    for (int32_t x = 0; x < k2; ++x) {
        const int32_t gs = p.in_off[mid + x] - e0, ge = p.in_off[mid + x + 1] - e0, deg = ge - gs;
        if (deg < 1) return 0;
        if (deg > 32) {
            if (!allow_long || nl >= LANE_MAX_LONG) return 0;
            if (long_j) long_j[nl] = (uint16_t)x;
            ++nl;
            cut(gs);
            for (int32_t at = gs + 32; at < ge; at += 32) cut(at);
            cut(ge);
            continue;
        }
        longest = std::max(longest, deg);
        if (ge - start > 32) cut(gs);
    }
    if (start < n_in) { ++nb; if (out) out[nb] = (uint16_t)n_in; }
    if (rounds) { int r = 0; while ((1 << r) < longest) ++r; *rounds = r; }
    if (n_long) *n_long = nl;
    return nb;
}

}  // namespace

// Checkpoints of the traceback: every cell of a checkpoint level walks back to the next checkpoint (dip_anc_kernel), so a
// checkpoint costs (R+1) k^2 walkers times its distance to the next one — the narrow levels are the cheap ones.  From `cur` the
// next checkpoint is the narrowest level T/2 .. 3T/2 below; when that one is still wide (above 1.6 x the 10th percentile of
// the widths: a bubble of the graph that spans the whole window), the search goes on down to 16 T below for the first level
// under that bound — the walkers of the narrow level `cur` pay the longer way, a wide level's (quadratically more) never start.
std::vector<int32_t> choose_checkpoints(const std::vector<int32_t>& level_off, int T) {
    const int L = (int)level_off.size() - 1;
    std::vector<int32_t> cp;
    if (T < 1) T = 1;
    auto width = [&](int l) { return level_off[l + 1] - level_off[l]; };
    int32_t bound = 0;
    {
        std::vector<int32_t> w((size_t)L);
        for (int l = 0; l < L; ++l) w[(size_t)l] = width(l);
        std::nth_element(w.begin(), w.begin() + L / 10, w.end());
        bound = w[(size_t)(L / 10)] + (w[(size_t)(L / 10)] * 3 + 4) / 5;
    }
    int cur = L - 1;
    cp.push_back(cur);
    while (cur > 0) {
        const int hi = cur - std::max(1, T / 2), lo = cur - (T + T / 2);
        if (lo <= 0) { cp.push_back(0); break; }
        int best = hi;
        for (int l = hi; l >= lo; --l)
            if (width(l) < width(best)) best = l;
        if (width(best) > bound) {
            const int far = std::max(1, cur - 16 * T);
            for (int l = lo - 1; l >= far; --l) {
                if (width(l) < width(best)) best = l;
                if (width(best) <= bound) break;
            }
        }
        cp.push_back(best);
        cur = best;
    }
    return cp;
}

void plan_tasks(DipPlan& p, const SweepShape& sh) {
    const int L = p.L, T = L - 1;                  // T = number of transitions
    const int G = sh.grid < 1 ? 1 : sh.grid;
    const int NR = sh.replicas < 1 ? 1 : sh.replicas, RK = sh.rank;      // row-sharded over NR ranks (dp_prep.h)
    const int GG = G * NR;                                                // global CTAs
    const uint32_t CT = sh.threads < 1 ? 1u : (uint32_t)sh.threads;
    const size_t slot = (size_t)sh.slot_bytes;
    const int lrc = (sh.lane_rc == LANE_RC_BIG) ? LANE_RC_BIG : LANE_RC_SMALL;
    p.grid = G;
    p.P.assign(L, 1); p.bar_target.assign(L, 0); p.bar_edge.assign(L, 0); p.narrow.assign(L, 0);
    p.rec_off.assign(L, -1); p.delta_off.assign(L, -1); p.delta_list.clear(); p.delta_elems = 0;
    p.records.clear(); p.tasks.clear(); p.task_begin.assign((size_t)G + 1, 0);
    p.n_narrow = p.n_wide = p.n_tasks_global = p.n_tasks_masks = 0;
    if (T <= 0) return;
    const int NT = n_threads();
    (void)NT;
    auto width = [&](int l) { return p.level_off[l + 1] - p.level_off[l]; };
    // tile cells a layer set needs: R+1 layers rounded up to whole lane-form chunks, so that the loads of the last
    // chunk's layers above R (never stored) stay inside the tile
    const uint64_t layers_padded = (uint64_t)((p.R + lrc) / lrc) * lrc;
    auto cells_of = [&](int l) { const uint64_t k = (uint64_t)width(l); return layers_padded * k * k; };
    auto nin_of = [&](int l) { return (int64_t)p.in_off[p.level_off[l + 2]] - (int64_t)p.in_off[p.level_off[l + 1]]; };

    // ---- 1. lane-form blocks, record sizes, pair-score matrix layout, placement of the layers ----
    std::vector<int32_t> nblk(T, 0);
    std::vector<uint8_t> rec_staged(T, 0), rec_inplace(T, 0), seg_rounds(T, 0), nlong(T, 0);
    std::vector<uint32_t> rec_bytes(T, 0);
#pragma omp parallel for schedule(static) num_threads(NT)
    for (int l = 0; l < T; ++l) {
        int rounds = 0, nl = 0;
        nblk[l] = lane_blocks(p, l, sh.allow_long, nullptr, nullptr, &rounds, &nl);
        if (nblk[l] > 0 && nl > 0 && nl * (p.R + 1) > LANE_SCRATCH_ENTRIES) { nblk[l] = 0; nl = 0; }     // not even one row fits the scratch
        seg_rounds[l] = (uint8_t)rounds;
        nlong[l] = (uint8_t)(nblk[l] > 0 ? nl : 0);
        const int64_t n_in = nin_of(l);
        const size_t rb = rec_bytes_for(width(l + 1), n_in, nblk[l], nlong[l]);
        rec_staged[l] = (sizeof(TaskHdr) + rb <= slot && n_in < 65536) ? 1 : 0;
        // a record too big for a slot still exists when the level can run in lane form: the lanes then read it in place
        // (wide panels: thousands of in-edges per level)
        rec_inplace[l] = (!rec_staged[l] && sh.allow_long && nblk[l] > 0 && n_in < 65536 && rb < ((size_t)1 << 31)) ? 1 : 0;
        rec_bytes[l] = (rec_staged[l] || rec_inplace[l]) ? (uint32_t)rb : 0u;
    }
    size_t rec_total = 0;
    for (int l = 0; l < T; ++l) {
        const int64_t n_in = nin_of(l);
        if (rec_staged[l] || rec_inplace[l]) { p.rec_off[l] = (int64_t)rec_total; rec_total += rec_bytes[l]; }
        // (u16 pair scores: a transition with more than 65535 colours takes the on-the-fly masks, whose delta is an int)
        if (p.lvlW[l] > 0 && (int64_t)p.lvlW[l] * 64 <= 65535 && n_in <= sh.delta_max_in && (p.delta_elems + n_in * n_in) * 2 <= sh.delta_budget) {
            p.delta_off[l] = p.delta_elems;
            p.delta_elems += (int64_t)align_up((size_t)(n_in * n_in), 8);
            p.delta_list.push_back(l);
        }
        const bool masks_mode = p.lvlW[l] > 0 && p.delta_off[l] < 0;     // too wide for a matrix: never narrow
        p.narrow[l] = (rec_staged[l] && !masks_mode && cells_of(l) <= (uint64_t)sh.tile_cells &&
                       cells_of(l + 1) <= (uint64_t)sh.tile_cells) ? 1 : 0;
        if (p.narrow[l]) ++p.n_narrow; else ++p.n_wide;
    }
    p.records.assign(rec_total, 0);
#pragma omp parallel for schedule(static) num_threads(NT)
    for (int l = 0; l < T; ++l) {
        if (!rec_staged[l] && !rec_inplace[l]) continue;
        uint8_t* r = p.records.data() + p.rec_off[l];
        const int32_t mid = p.level_off[l + 1], k2 = width(l + 1);
        const int32_t e0 = p.in_off[mid], e1 = p.in_off[p.level_off[l + 2]];
        uint16_t* off2 = reinterpret_cast<uint16_t*>(r);
        for (int32_t x = 0; x <= k2; ++x) off2[x] = (uint16_t)(p.in_off[mid + x] - e0);
        if (e1 > e0) {
            memcpy(r + rec_edge_offset(k2), &p.in_edge[e0], (size_t)(e1 - e0) * 4);
            memcpy(r + rec_dst_offset(k2, e1 - e0), &p.in_dst[e0], (size_t)(e1 - e0) * 2);
        }
        if (nblk[l] > 0)
            lane_blocks(p, l, sh.allow_long, reinterpret_cast<uint16_t*>(r + rec_bstart_offset(k2, e1 - e0)),
                        reinterpret_cast<uint16_t*>(r + rec_long_offset(k2, e1 - e0, nblk[l])), nullptr, nullptr);
    }

    // ---- 2. participants: narrow -> CTA 0 alone; wide -> P = min(G, k2) CTAs ----
    for (int l = 0; l < T; ++l) p.P[l] = p.narrow[l] ? 1 : std::max(1, std::min(GG, (int)width(l + 1)));

    // ---- 3. barrier schedule: a grid barrier follows transition l unless both it and the next run on CTA 0 alone ----
    uint32_t acc = 0, acc_local = 0;
    std::vector<uint32_t> loc_target((size_t)T, 0), loc_n((size_t)T, 0);     // this rank's share of the arrivals (row-sharded sweep)
    for (int l = 0; l < T; ++l) {
        const int pn = (l + 1 < T) ? p.P[l + 1] : 1;
        bool edge = (p.P[l] > 1) || (pn > 1);
        if (NR > 1) edge = !p.narrow[l] || (l + 1 < T && !p.narrow[l + 1]);   // (a one-CTA wide transition still has to reach the peers)
        p.bar_edge[l] = edge ? 1 : 0;
        if (edge) acc += (uint32_t)(p.narrow[l] ? NR : p.P[l]);      // a narrow transition runs (and arrives) once per rank
        if (edge && NR > 1) {
            // arriving CTAs of this rank: its CTA 0 for a narrow transition, else the global CTAs c < P with c % NR == RK
            loc_n[l] = p.narrow[l] ? 1u : (uint32_t)((p.P[l] - RK + NR - 1) / NR);
            if (!p.narrow[l] && RK >= p.P[l]) loc_n[l] = 0;
            acc_local += loc_n[l];
            loc_target[l] = acc_local;
        }
        p.bar_target[l] = acc;
    }

    // ---- 4. task streams: every host thread compiles a contiguous range of levels into per-CTA pieces, which
    // are then concatenated in thread order (= level order) ----
    std::vector<std::vector<std::vector<TaskHdr>>> piece((size_t)NT, std::vector<std::vector<TaskHdr>>((size_t)G));
    std::vector<int64_t> n_glob((size_t)NT, 0), n_mask((size_t)NT, 0);
#pragma omp parallel num_threads(NT)
    {
#if defined(_OPENMP)
        const int tno = omp_get_thread_num(), tcount = omp_get_num_threads();
#else
        const int tno = 0, tcount = 1;
#endif
        const int l_begin = (int)((int64_t)T * tno / tcount), l_end = (int)((int64_t)T * (tno + 1) / tcount);
        std::vector<std::vector<TaskHdr>>& stream = piece[(size_t)tno];
        std::vector<int32_t> cut;
        for (int l = l_begin; l < l_end; ++l) {
            const int32_t mid = p.level_off[l + 1], k = width(l), k2 = width(l + 1);
            const int64_t n_in = nin_of(l);
            const int32_t e0 = p.in_off[mid];
            const int P = p.P[l];
            // row partition, balanced by candidates (indeg(i') * n_in) plus a per-pair overhead
            cut.assign((size_t)P + 1, 0);
            cut[(size_t)P] = k2;
            if (P > 1) {
                uint64_t total = 0;
                for (int32_t x = 0; x < k2; ++x) total += (uint64_t)(p.in_off[mid + x + 1] - p.in_off[mid + x]) * (uint64_t)n_in + 4u * (uint64_t)k2;
                uint64_t accw = 0;
                int32_t x = 0;
                for (int q = 1; q < P; ++q) {
                    const uint64_t want = total * (uint64_t)q / (uint64_t)P;
                    const int32_t lo_x = cut[(size_t)q - 1] + 1, hi_x = k2 - (P - q);   // >= 1 row per chunk, now and later
                    while (x < hi_x && (x < lo_x || accw < want)) {
                        accw += (uint64_t)(p.in_off[mid + x + 1] - p.in_off[mid + x]) * (uint64_t)n_in + 4u * (uint64_t)k2;
                        ++x;
                    }
                    cut[(size_t)q] = x;
                }
            }
            const size_t rb = rec_staged[l] ? rec_bytes[l] : 0;      // bytes the producer stages
            uint32_t base_flags = 0;
            // layer l lives in shared memory iff it is produced and consumed by narrow transitions (level 0: by the kernel prologue)
            if (p.narrow[l] && (l == 0 || p.narrow[l - 1])) base_flags |= TK_SRC_SMEM;
            if (p.narrow[l] && l + 1 < T && p.narrow[l + 1]) base_flags |= TK_DST_SMEM;
            if (p.lvlW[l] > 0) base_flags |= (p.delta_off[l] >= 0) ? TK_DELTA : TK_DELTA_MASKS;
            if (!rec_staged[l]) base_flags |= TK_REC_GLOBAL;
            const bool same_place = ((base_flags & TK_SRC_SMEM) != 0) == ((base_flags & TK_DST_SMEM) != 0);   // hand-overs use the pair form
            const uint64_t kk2 = (uint64_t)k2 * (uint64_t)k2;
            const int32_t max_rows = (int32_t)std::max<uint64_t>(1, std::min<uint64_t>(65535, ((1ull << 32) - 1) / std::max<uint64_t>(kk2, 1)));
            for (int c = 0; c < P; ++c) {
                // global CTA c of a wide transition belongs to rank c % NR; narrow ones (P == 1) run on every rank's CTA 0
                if (NR > 1 && !p.narrow[l] && c % NR != RK) continue;
                const int lc = p.narrow[l] ? 0 : c / NR;
                const int32_t ra = cut[(size_t)c], rbnd = cut[(size_t)c + 1];
                int32_t x = ra;
                bool first = true;
                while (x < rbnd) {
                    TaskHdr h;
                    memset(&h, 0, sizeof h);
                    h.level = l; h.k = (uint16_t)k; h.k2 = (uint16_t)k2; h.i0 = (uint16_t)x;
                    h.n_in = (uint32_t)((rec_staged[l] || rec_inplace[l]) ? n_in : 0);
                    h.flags = base_flags;
                    h.pred_off2 = p.pred_off[l + 1];
                    int32_t y = std::min(rbnd, x + max_rows);
                    if (nlong[l] > 0)          // rows x long destinations x layers of a task share the CTA's scratch words
                        y = std::min(y, x + std::max(1, LANE_SCRATCH_ENTRIES / ((int)nlong[l] * (p.R + 1))));
                    if (rec_staged[l]) {
                        h.rec_off16 = (uint32_t)(p.rec_off[l] / 16);
                        h.rec_bytes = (uint32_t)rb;
                        if (base_flags & TK_DELTA) {
                            // stage as many whole rows of the matrix as the slot still holds
                            const size_t avail = slot - sizeof(TaskHdr) - rb;
                            const int64_t s_el = (int64_t)(p.in_off[mid + x] - e0) * n_in;
                            const int64_t s_al = s_el & ~(int64_t)7;                  // 16-byte aligned (matrix base is)
                            int32_t yy = x;
                            int64_t bytes = 0;
                            while (yy < y) {
                                const int64_t e_el = (int64_t)(p.in_off[mid + yy + 1] - e0) * n_in;
                                const int64_t b = (int64_t)align_up((size_t)((e_el - s_al) * 2), 16);
                                if ((size_t)b > avail) break;
                                bytes = b; ++yy;
                            }
                            if (yy > x) {
                                y = yy;
                                if (bytes > 0) {
                                    h.flags |= TK_DELTA_STAGED;
                                    h.delta_off16 = (uint32_t)((p.delta_off[l] + s_al) / 8);
                                    h.delta_bytes = (uint32_t)bytes;
                                    h.delta_skew = (uint32_t)(s_el - s_al);
                                }
                            } else {
                                y = x + 1;          // one row larger than a slot: read its matrix rows in place
                            }
                        }
                    } else {
                        if (rec_inplace[l]) h.rec_off16 = (uint32_t)(p.rec_off[l] / 16);      // read in place; rec_bytes stays 0
                        ++n_glob[(size_t)tno];
                    }
                    if (base_flags & TK_DELTA_MASKS) ++n_mask[(size_t)tno];
                    h.i1 = (uint16_t)y;
                    const uint32_t npairs = (uint32_t)(y - x) * (uint32_t)k2;
                    h.m_k2 = make_magic((uint32_t)k2);
                    h.m_pairs = make_magic(npairs);
                    const int rc = choose_rc(npairs, p.R, CT);
                    const uint32_t nchunk = (uint32_t)(p.R + rc) / (uint32_t)rc;
                    h.rc = (uint16_t)rc;
                    h.groups = (uint16_t)(npairs >= CT ? 1u : std::max(1u, std::min(nchunk, CT / npairs)));
                    h.n_active = std::min<uint32_t>(CT, (uint32_t)h.groups * npairs);
                    // lane form: record staged, level eligible, pair scores absent or in a matrix (staged rows or in place)
                    // (a matrix whose rows do not fit the slot is read in place from HBM/L2: wide levels, long rows)
                    const bool scores_ok = !(base_flags & TK_DELTA_MASKS);
                    const uint32_t lchunks = (uint32_t)(p.R + lrc) / (uint32_t)lrc;
                    const uint64_t witems = (uint64_t)(y - x) * (uint64_t)std::max(nblk[l], 1) * lchunks;   // upper bound (rp >= 1)
                    if ((rec_staged[l] || rec_inplace[l]) && nblk[l] > 0 && scores_ok && same_place && CT % 32 == 0 &&
                        witems * (uint64_t)std::max<int64_t>(nblk[l], y - x) < (1ull << 32)) {
                        h.flags |= TK_LANES;
                        h.nblk = (uint16_t)nblk[l];
                        h.rp = (uint16_t)((nblk[l] == 1) ? std::max<int64_t>(1, std::min<int64_t>(32 / n_in, y - x)) : 1);
                        h.nrg = (uint32_t)((y - x + h.rp - 1) / h.rp);
                        h.m_nblk = make_magic((uint32_t)nblk[l]);
                        h.m_nrg = make_magic(h.nrg);
                        h.m_nin = make_magic((uint32_t)n_in);
                        // (levels with long destinations run the unpacked lane task, which is compiled for LANE_RC_SMALL)
                        const int trc = nlong[l] > 0 ? LANE_RC_SMALL : lrc;
                        h.rc = (uint16_t)trc;
                        h.n_witems = h.nrg * (uint32_t)nblk[l] * ((uint32_t)(p.R + trc) / (uint32_t)trc);
                        h.rounds = seg_rounds[l];
                        h.bstart_off = (uint32_t)rec_bstart_offset(k2, n_in);
                        if (nlong[l] > 0) {
                            h.flags |= TK_LONG;
                            h.n_long = nlong[l];
                            h.long_off = (uint32_t)rec_long_offset(k2, n_in, nblk[l]);
                        }
                        h.n_active = std::min<uint32_t>(CT, 32u * h.n_witems);
                    }
                    if (first && l > 0 && p.bar_edge[l - 1]) { h.flags |= TK_WAIT; h.wait_target = p.bar_target[l - 1]; }
                    if (y >= rbnd) {
                        h.flags |= TK_BAR;
                        if (p.bar_edge[l]) { h.flags |= TK_ARRIVE; h.arrive_local_target = loc_target[l]; h.arrive_n = loc_n[l]; }
                        if (NR > 1 && !p.narrow[l]) { h.flags |= TK_PUSH; h.push_i0 = (uint16_t)ra; h.push_i1 = (uint16_t)rbnd; }
                    }
                    stream[(size_t)lc].push_back(h);
                    first = false;
                    x = y;
                }
            }
        }
    }
    for (int t = 0; t < NT; ++t) { p.n_tasks_global += n_glob[(size_t)t]; p.n_tasks_masks += n_mask[(size_t)t]; }
    // offsets: CTA-major, then host-thread (level) order
    std::vector<size_t> dst_off((size_t)NT * (size_t)G, 0);
    size_t total = 0;
    for (int c = 0; c < G; ++c) {
        p.task_begin[(size_t)c] = (int64_t)total;
        for (int t = 0; t < NT; ++t) { dst_off[(size_t)t * (size_t)G + (size_t)c] = total; total += piece[(size_t)t][(size_t)c].size(); }
    }
    p.task_begin[(size_t)G] = (int64_t)total;
    p.tasks.resize(total);
#pragma omp parallel for schedule(static) num_threads(NT)
    for (int t = 0; t < NT; ++t)
        for (int c = 0; c < G; ++c) {
            const std::vector<TaskHdr>& v = piece[(size_t)t][(size_t)c];
            if (!v.empty()) memcpy(p.tasks.data() + dst_off[(size_t)t * (size_t)G + (size_t)c], v.data(), v.size() * sizeof(TaskHdr));
        }
}

}  // namespace dg
