// dp_plan4.cpp — see dp_plan4.h.
#include "dp_plan4.h"

#include <algorithm>
#include <cstring>

namespace dg {

namespace {

// Barrier schedule of a directory (monotone arrival counter, dp_sweep4.cuh): a transition between HBM-resident levels is
// shared by all CTAs (wait for everything before it, then every CTA arrives); a hand-over from shared memory to HBM runs
// on CTA 0, which arrives when the level is whole; a hand-over back waits for everything before it.
void schedule(Plan4Dir& d, int grid) {
    uint32_t cum = 0;
    d.wide_list.clear();
    for (size_t x = 0; x < d.dir.size(); ++x) {
        ProgDir& e = d.dir[x];
        const bool ss = e.flags & PF_SRC_SMEM, ds = e.flags & PF_DST_SMEM;
        e.flags &= ~(uint32_t)(PF_ALL_CTAS | PF_WAIT | PF_ARRIVE);
        e.wait_target = 0;
        if (!ss && !ds) { e.flags |= PF_ALL_CTAS | PF_WAIT | PF_ARRIVE; e.wait_target = cum; cum += (uint32_t)grid; d.wide_list.push_back((int32_t)x); }
        else if (ss && !ds) { e.flags |= PF_ARRIVE; cum += 1; }
        else if (!ss && ds) { e.flags |= PF_WAIT; e.wait_target = cum; }
    }
    d.final_target = cum;
    d.n = (int32_t)d.dir.size();
}

}  // namespace

bool plan4_build(const DipPlan& p, const DipGraphView& g, const Sweep4Shape& shape, int rc, Plan4& q, std::string& why) {
    q.shape = shape;
    const int L = p.L;
    q.L = L; q.R = p.R; q.rc = rc;
    q.nchunk = (p.R + rc) / rc;
    q.RL = q.nchunk * rc;
    if (L < 2) { why = "single level"; return false; }
    if (p.value_bound >= KEY_VALUE_LIMIT) { why = "DP values may exceed the packed key"; return false; }
    if ((int64_t)shape.kn * shape.kn > (int64_t)shape.cells() || shape.kn > 32 || shape.kn < 2) { why = "bad shape"; return false; }
    if (p.kmax >= 32768) { why = "level wider than 32767 vertices"; return false; }
    const int32_t V = p.V;
    const int kn = shape.kn;
    q.cls_list.assign((size_t)V, 0);
    q.vinfo.assign((size_t)V, 0);
    q.vslot.assign((size_t)V, 0);
    q.lvl_dom.assign((size_t)L, 0);
    q.lvl_n1.assign((size_t)L, 0); q.lvl_np.assign((size_t)L, 0); q.lvl_m.assign((size_t)L, 0); q.lvl_z.assign((size_t)L, 0); q.lvl_dm.assign((size_t)L, 0);
    q.mpre_off.assign((size_t)L, 0);
    q.tflags.assign((size_t)L - 1, 0);

    // ---- slots and placement (sequential over the levels, O(V)) ----
    std::vector<uint8_t> passive((size_t)V, 0);
    {
        std::vector<uint16_t> sfree, hfree;          // free slots of the two tiles (stacks)
        for (int s = kn - 1; s >= 1; --s) sfree.push_back((uint16_t)s);
        uint32_t hnext = 0;
        std::vector<uint8_t> inherited;
        q.vslot[0] = 0;
        int64_t hmax = 1;
        for (int l = 0; l + 1 < L; ++l) {
            const int32_t lo = p.level_off[l], mid = p.level_off[l + 1], hi = p.level_off[l + 2];
            const int32_t k = mid - lo, k2 = hi - mid;
            inherited.assign((size_t)k, 0);
            int np = 0;
            for (int32_t v = mid; v < hi; ++v) {
                if (p.in_off[(size_t)v + 1] - p.in_off[v] != 1) continue;
                const uint32_t e = p.in_edge[(size_t)p.in_off[v]];
                const int32_t u = lo + (int32_t)(e & 0xFFFFu);
                if ((e >> 16) != 0 || inherited[(size_t)(u - lo)]) continue;
                if (g.col_off[v + 1] != g.col_off[v] || g.col_off[u + 1] != g.col_off[u]) continue;
                passive[v] = 1; inherited[(size_t)(u - lo)] = 1; ++np;
            }
            const uint8_t dom = q.lvl_dom[l];
            bool relocate;
            uint8_t ndom;
            if (dom == 0) { relocate = k + (k2 - np) > kn; ndom = relocate ? 1 : 0; }
            else { relocate = k2 <= kn * 3 / 4; ndom = relocate ? 0 : 1; }
            q.lvl_dom[l + 1] = ndom;
            uint32_t f = 0;
            if (dom == 0) f |= PF_SRC_SMEM;
            if (ndom == 0) f |= PF_DST_SMEM;
            if (relocate) {
                f |= PF_RELOCATE;
                ++q.n_relocate;
                for (int32_t v = mid; v < hi; ++v) { passive[v] = 0; q.vslot[v] = (uint16_t)(v - mid); }
                if (ndom == 0) {        // everything of the HBM tile is released; the shared-memory tile starts compact
                    sfree.clear();
                    for (int s = kn - 1; s >= k2; --s) sfree.push_back((uint16_t)s);
                    hfree.clear(); hnext = 0;
                } else {
                    hfree.clear(); hnext = (uint32_t)k2;
                    sfree.clear();
                    for (int s = kn - 1; s >= 0; --s) sfree.push_back((uint16_t)s);
                }
            } else {
                for (int32_t v = mid; v < hi; ++v) {
                    if (passive[v]) {
                        const uint32_t e = p.in_edge[(size_t)p.in_off[v]];
                        q.vslot[v] = q.vslot[lo + (int32_t)(e & 0xFFFFu)];
                    } else if (ndom == 0) {
                        if (sfree.empty()) { why = "internal: shared-memory slots exhausted"; return false; }
                        q.vslot[v] = sfree.back(); sfree.pop_back();
                    } else {
                        if (!hfree.empty()) { q.vslot[v] = hfree.back(); hfree.pop_back(); }
                        else { if (hnext >= 65535u) { why = "more than 65534 HBM slots"; return false; } q.vslot[v] = (uint16_t)hnext++; }
                    }
                }
                for (int32_t u = lo; u < mid; ++u)
                    if (!inherited[(size_t)(u - lo)]) (ndom == 0 ? sfree : hfree).push_back(q.vslot[u]);
            }
            if (ndom == 1) hmax = std::max<int64_t>(hmax, (int64_t)hnext);
            q.tflags[l] = f;
        }
        q.hstride = (int32_t)hmax;
    }

    // ---- class tables (level-parallel) ----
    int64_t mp = 0;
    for (int l = 0; l < L; ++l) {
        const int32_t lo = p.level_off[l], hi = p.level_off[l + 1];
        uint32_t n1 = 0, np = 0, m = 0, z = 0;
        for (int32_t v = lo; v < hi; ++v) {
            const int32_t d = p.in_off[(size_t)v + 1] - p.in_off[v];
            if (d == 1) { ++n1; np += passive[v]; } else if (d >= 2) ++m; else ++z;
        }
        q.lvl_n1[l] = n1; q.lvl_np[l] = np; q.lvl_m[l] = m; q.lvl_z[l] = z;
        q.mpre_off[l] = mp;
        mp += (int64_t)m + 1;
    }
    q.mpre.assign((size_t)mp, 0);
    std::vector<uint32_t> lvl_dmax((size_t)L, 0);      // largest in-degree of the level
    uint64_t max_cand = 0;
#pragma omp parallel for schedule(static) reduction(max : max_cand)
    for (int l = 0; l < L; ++l) {
        const int32_t lo = p.level_off[l], hi = p.level_off[l + 1];
        const uint32_t n1 = q.lvl_n1[l], np = q.lvl_np[l], m = q.lvl_m[l], na = n1 - np;
        uint32_t a = 0, ap = 0, b = 0, c = 0, dm = 0, dmax = 0;
        uint32_t* P = q.mpre.data() + q.mpre_off[l];
        uint16_t* list = q.cls_list.data() + lo;
        for (int32_t v = lo; v < hi; ++v) {
            const uint32_t d = (uint32_t)(p.in_off[(size_t)v + 1] - p.in_off[v]);
            const uint16_t pos = (uint16_t)(v - lo);
            if (d == 1 && !passive[v]) list[a++] = pos;
            else if (d == 1) list[na + ap++] = pos;
            else if (d >= 2) list[n1 + b++] = pos;
            else list[n1 + m + c++] = pos;
        }
        // (every class stays in POSITION order.  Slot order was worth something while the HBM tile was layer-major — the cells of
        // a block then sat next to each other; with the cell-major tile a cell is its own 96-byte record, and the four sorts
        // per level were 15 % of the planner's time.  The tie-break never depended on it: it orders the candidates of one cell.)
        for (uint32_t x = 0; x < n1; ++x) q.vinfo[(size_t)lo + list[x]] = x;
        for (uint32_t x = 0; x < m; ++x) {
            const int32_t v = lo + list[n1 + x];
            const uint32_t d = (uint32_t)(p.in_off[(size_t)v + 1] - p.in_off[v]);
            q.vinfo[v] = x | (1u << 30); P[x] = dm; dm += d; dmax = std::max(dmax, d);
        }
        for (uint32_t x = 0; x < (uint32_t)(hi - lo) - n1 - m; ++x) q.vinfo[(size_t)lo + list[n1 + m + x]] = x | (2u << 30);
        P[m] = dm;
        q.lvl_dm[l] = dm;
        lvl_dmax[l] = dmax;
        max_cand = std::max<uint64_t>(max_cand, (uint64_t)dmax * dmax);
    }
    q.max_cand = (uint32_t)std::min<uint64_t>(max_cand, 0xFFFFFFFFu);
    if (max_cand > PROG_MAX_CAND) { why = "a cell has more candidates than the packed ordinal holds"; return false; }

    // ---- headers, offsets ----
    q.hdr.assign((size_t)L - 1, ProgHdr());
    q.prog_off.assign((size_t)L, 0);
    q.pred_off.assign((size_t)L + 1, 0);
    std::vector<uint64_t> bytes((size_t)L - 1, 0);
    uint64_t written = 0, total = 0;
#pragma omp parallel for schedule(static) reduction(+ : written, total)
    for (int l = 0; l < L - 1; ++l) {
        const int32_t k = p.level_off[l + 1] - p.level_off[l], k2 = p.level_off[l + 2] - p.level_off[l + 1];
        LevelClass c;
        c.k2 = (uint32_t)k2; c.n1 = q.lvl_n1[l + 1]; c.np = q.lvl_np[l + 1]; c.m = q.lvl_m[l + 1]; c.z = q.lvl_z[l + 1]; c.dm = q.lvl_dm[l + 1];
        c.list = q.cls_list.data() + p.level_off[l + 1];
        c.mpre = q.mpre.data() + q.mpre_off[l + 1];
        const bool relocate = q.tflags[l] & PF_RELOCATE;
        const ProgCounts n = prog_counts(c, relocate);
        const bool both_smem = (q.tflags[l] & PF_SRC_SMEM) && (q.tflags[l] & PF_DST_SMEM);
        // (a cell of more than PROG_KEY_CAND candidates needs the generic path's warp form: striped ordinals)
        const bool compact = both_smem && (int64_t)p.lvlW[l] * 64 <= (int64_t)PROG_COMPACT_DELTA &&
                             (uint64_t)lvl_dmax[(size_t)l + 1] * lvl_dmax[(size_t)l + 1] <= PROG_KEY_CAND;
        if (compact) q.tflags[l] |= PF_COMPACT;
        ProgHdr& h = q.hdr[l];
        h.k = (uint16_t)k; h.k2 = (uint16_t)k2;
        h.n_copy = (uint32_t)n.n_copy; h.n_multi = (uint32_t)n.n_multi; h.n_cand = (uint32_t)n.n_cand;
        h.n_big = (uint32_t)n.n_big; h.n_dead = (uint32_t)n.n_dead; h.max_n = n.max_n;
        h.n_passive = relocate ? 0u : c.np;
        h.n_mm = c.m * c.m;
        h.n_giant = 0;
        if ((uint64_t)lvl_dmax[(size_t)l + 1] * lvl_dmax[(size_t)l + 1] > PROG_KEY_CAND)
            for (uint32_t x = 0; x < c.m; ++x)
                for (uint32_t y = 0; y < c.m; ++y)
                    h.n_giant += (uint64_t)(c.mpre[x + 1] - c.mpre[x]) * (c.mpre[y + 1] - c.mpre[y]) > PROG_KEY_CAND;
        const ProgLayout lay = prog_layout(compact, n.n_copy, n.n_multi, n.n_cand, n.n_big, n.n_dead);
        h.off_cell = (uint32_t)lay.cell; h.off_cand = (uint32_t)lay.cand; h.off_big = (uint32_t)lay.big; h.off_dead = (uint32_t)lay.dead;
        bytes[l] = lay.end > 0xFFFFFFFFull ? ~0ull : (uint64_t)lay.end;
        written += n.n_copy + n.n_multi + n.n_dead;
        total += (uint64_t)k2 * (uint64_t)k2;
    }
    q.cells_written = written; q.cells_total = total;
    for (int l = 0; l + 1 < L; ++l) q.max_giant = std::max(q.max_giant, q.hdr[l].n_giant);
    if (q.max_giant > 1024) { why = "more than 1024 cells above 1024 candidates in one level"; return false; }     // dp_sweep4.cuh: GIANT_LIST_MAX
    for (int l = 0; l + 1 < L; ++l) {
        if (bytes[l] == ~0ull) { why = "a transition's program exceeds 4 GB"; return false; }
        q.prog_off[(size_t)l + 1] = q.prog_off[l] + (bytes[l] + PROG_ALIGN - 1) / PROG_ALIGN * PROG_ALIGN;
        if (q.prog_off[l] / PROG_ALIGN > 0xFFFFFFFFull) { why = "program larger than 256 GB"; return false; }
    }
    q.prog_bytes = q.prog_off[(size_t)L - 1];
    for (int l = 0; l < L; ++l) {
        const int64_t nm = l >= 1 ? (int64_t)q.hdr[(size_t)l - 1].n_multi : 0;
        q.pred_off[(size_t)l + 1] = q.pred_off[l] + (int64_t)q.RL * nm;
        if (l >= 1) q.hdr[(size_t)l - 1].pred_off = (uint64_t)q.pred_off[l];
    }
    q.pred_elems = q.pred_off[L];

    // ---- directories ----
    q.full.dir.reserve((size_t)L - 1);          // (no regrowth: the storage is a monotonic page-locked block)
    q.timed.dir.reserve((size_t)L - 1);
    for (int l = 0; l + 1 < L; ++l) {
        ProgDir d;
        memset(&d, 0, sizeof d);
        d.off64 = (uint32_t)(q.prog_off[l] / PROG_ALIGN);
        d.flags = q.tflags[l];
        d.level = l;
        d.prog16 = (uint32_t)(bytes[l] / 16);
        if (bytes[l] + sizeof(ProgDir) <= (uint64_t)shape.slot_bytes) { d.flags |= PF_STAGED; d.stage_bytes = (uint32_t)bytes[l]; }
        else d.stage_bytes = (uint32_t)sizeof(ProgHdr);
        if ((d.flags & PF_SRC_SMEM) && (d.flags & PF_DST_SMEM)) ++q.n_smem_trans;
        q.full.dir.push_back(d);
        const ProgHdr& h = q.hdr[l];
        if ((d.flags & PF_RELOCATE) || h.n_copy + h.n_multi + h.n_dead > 0) q.timed.dir.push_back(d);
        else ++q.n_skipped;
    }
    schedule(q.full, shape.grid);
    schedule(q.timed, shape.grid);
    // cell records padded to whole 32-byte sectors — unless the tile then outgrows the L2 (126 MB on B200: 40 layers x 824^2
    // slots = 130 MB at 192 bytes per cell, 114 MB at 168), where an even word count (8-byte stores) has to do
    q.gcs = (q.RL + 2 + 7) / 8 * 8;
    if ((int64_t)q.gcs * q.hstride * q.hstride * 4 > ((int64_t)96 << 20)) q.gcs = (q.RL + 2 + 1) / 2 * 2;
    q.gtile_cells = (int64_t)q.gcs * q.hstride * q.hstride;
    const int32_t last = p.level_off[L - 1];
    const uint32_t ls = q.vslot[last];
    q.sink_cell = ls * (uint32_t)(q.lvl_dom[L - 1] == 0 ? kn : q.hstride) + ls;
    q.last_smem = q.lvl_dom[L - 1] == 0;
    return true;
}

ProgLevelIn plan4_level_in(const DipPlan& p, const Plan4& q, int l) {
    ProgLevelIn in;
    const int32_t lo = p.level_off[l], mid = p.level_off[l + 1];
    in.k = (uint32_t)(mid - lo); in.k2 = (uint32_t)(p.level_off[l + 2] - mid);
    in.in_off = p.in_off.data() + mid;
    in.in_edge = p.in_edge.data();
    in.cls.k2 = in.k2; in.cls.n1 = q.lvl_n1[l + 1]; in.cls.np = q.lvl_np[l + 1]; in.cls.m = q.lvl_m[l + 1]; in.cls.z = q.lvl_z[l + 1];
    in.cls.dm = q.lvl_dm[l + 1];
    in.cls.list = q.cls_list.data() + mid;
    in.cls.mpre = q.mpre.data() + q.mpre_off[l + 1];
    in.slot_src = q.vslot.data() + lo; in.slot_dst = q.vslot.data() + mid;
    in.stride_src = q.lvl_dom[l] == 0 ? (uint32_t)q.shape.kn : (uint32_t)q.hstride;
    in.stride_dst = q.lvl_dom[l + 1] == 0 ? (uint32_t)q.shape.kn : (uint32_t)q.hstride;
    in.relocate = (q.tflags[l] & PF_RELOCATE) != 0;
    in.W = p.lvlW[l];
    in.msrc = in.W ? p.masks.data() + p.msrc_off[l] : nullptr;
    in.mdst = in.W ? p.masks.data() + p.mdst_off[l] : nullptr;
    return in;
}

void prog_fill_level_host(const DipPlan& p, const Plan4& q, int l, uint8_t* out) {
    const ProgHdr& h = q.hdr[l];
    const bool compact = q.tflags[l] & PF_COMPACT;
    const ProgLayout lay = prog_layout(compact, h.n_copy, h.n_multi, h.n_cand, h.n_big, h.n_dead);
    memset(out, 0, lay.end);
    memcpy(out, &h, sizeof h);
    const ProgLevelIn in = plan4_level_in(p, q, l);
    for (uint64_t t = 0; t < h.n_copy; ++t) {
        const CopyDesc d = make_copy(in, t);
        if (compact) reinterpret_cast<uint32_t*>(out + lay.copy)[t] = pack_copy_c(d);
        else { uint32_t* w = reinterpret_cast<uint32_t*>(out + lay.copy) + 4 * t; w[0] = d.src | (d.w << 30); w[1] = d.dst; w[2] = d.delta; w[3] = 0; }
    }
    uint32_t nb = 0;
    for (uint64_t t = 0; t < h.n_multi; ++t) {
        const MultiCell c = multi_cell(in, t);
        const uint32_t dst = dst_cell(in, c.i2, c.j2);
        if (compact) { uint32_t* w = reinterpret_cast<uint32_t*>(out + lay.cell) + 2 * t; w[0] = dst | (c.n << 16); w[1] = (uint32_t)c.cand_off; }
        else { uint32_t* w = reinterpret_cast<uint32_t*>(out + lay.cell) + 4 * t; w[0] = dst; w[1] = c.n; w[2] = (uint32_t)c.cand_off; w[3] = 0; }
        for (uint32_t o = 0; o < c.n; ++o) {
            const CandDesc d = make_cand(in, c, o);
            if (compact) reinterpret_cast<uint32_t*>(out + lay.cand)[c.cand_off + o] = pack_cand_c(d);
            else { uint32_t* w = reinterpret_cast<uint32_t*>(out + lay.cand) + 2 * (c.cand_off + o); w[0] = d.src | (d.w << 30); w[1] = d.delta; }
        }
        if (c.n >= PROG_BIG_MIN) reinterpret_cast<uint32_t*>(out + lay.big)[nb++] = (uint32_t)t;
    }
    for (uint64_t x = 0; x < h.n_dead; ++x) reinterpret_cast<uint32_t*>(out + lay.dead)[x] = dead_cell(in, x);
}

}  // namespace dg
