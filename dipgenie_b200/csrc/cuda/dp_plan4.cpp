// dp_plan4.cpp — see dp_plan4.h.
#include "dp_plan4.h"

#include <algorithm>
#include <cstring>

namespace dg {

bool plan4_build(const DipPlan& p, const Sweep4Shape& shape, int rc, Plan4& q, std::string& why) {
    q = Plan4();
    q.shape = shape;
    const int L = p.L;
    q.L = L; q.R = p.R; q.rc = rc;
    q.nchunk = (p.R + rc) / rc;
    q.RL = q.nchunk * rc;
    if (L < 2) { why = "single level"; return false; }
    if (p.value_bound >= KEY_VALUE_LIMIT) { why = "DP values may exceed the packed key"; return false; }
    if ((int64_t)shape.kn * shape.kn > ((int64_t)1 << shape.slog)) { why = "bad shape"; return false; }
    if (p.kmax >= 32768) { why = "level wider than 32767 vertices"; return false; }
    const int32_t V = p.V;
    q.cls_list.assign((size_t)V, 0);
    q.vinfo.assign((size_t)V, 0);
    q.lvl_n1.assign((size_t)L, 0); q.lvl_m.assign((size_t)L, 0); q.lvl_z.assign((size_t)L, 0); q.lvl_dm.assign((size_t)L, 0);
    q.mpre_off.assign((size_t)L, 0);
    // class tables
    int64_t mp = 0;
    for (int l = 0; l < L; ++l) {
        const int32_t lo = p.level_off[l], hi = p.level_off[l + 1];
        uint32_t n1 = 0, m = 0, z = 0;
        for (int32_t v = lo; v < hi; ++v) {
            const int32_t d = p.in_off[(size_t)v + 1] - p.in_off[v];
            if (d == 1) ++n1; else if (d >= 2) ++m; else ++z;
        }
        q.lvl_n1[l] = n1; q.lvl_m[l] = m; q.lvl_z[l] = z;
        q.mpre_off[l] = mp;
        mp += (int64_t)m + 1;
    }
    q.mpre.assign((size_t)mp, 0);
    uint64_t max_cand = 0;
#pragma omp parallel for schedule(static) reduction(max : max_cand)
    for (int l = 0; l < L; ++l) {
        const int32_t lo = p.level_off[l], hi = p.level_off[l + 1];
        const uint32_t n1 = q.lvl_n1[l], m = q.lvl_m[l];
        uint32_t a = 0, b = 0, c = 0, dm = 0, dmax = 0;
        uint32_t* P = q.mpre.data() + q.mpre_off[l];
        uint16_t* list = q.cls_list.data() + lo;
        for (int32_t v = lo; v < hi; ++v) {
            const uint32_t d = (uint32_t)(p.in_off[(size_t)v + 1] - p.in_off[v]);
            const uint16_t pos = (uint16_t)(v - lo);
            if (d == 1) { q.vinfo[v] = a; list[a++] = pos; }
            else if (d >= 2) { q.vinfo[v] = b | (1u << 30); P[b] = dm; dm += d; dmax = std::max(dmax, d); list[n1 + b++] = pos; }
            else { q.vinfo[v] = c | (2u << 30); list[n1 + m + c++] = pos; }
        }
        P[m] = dm;
        q.lvl_dm[l] = dm;
        max_cand = std::max<uint64_t>(max_cand, (uint64_t)dmax * dmax);
    }
    q.max_cand = (uint32_t)std::min<uint64_t>(max_cand, 0xFFFFFFFFu);
    if (max_cand > PROG_MAX_CAND) { why = "a cell has more candidates than the packed ordinal holds"; return false; }

    // headers, offsets, directory
    q.hdr.assign((size_t)L - 1, ProgHdr());
    q.dir.assign((size_t)L - 1, ProgDir());
    q.prog_off.assign((size_t)L, 0);
    q.pred_off.assign((size_t)L + 1, 0);
    std::vector<uint64_t> bytes((size_t)L - 1, 0);
#pragma omp parallel for schedule(static)
    for (int l = 0; l < L - 1; ++l) {
        const int32_t k = p.level_off[l + 1] - p.level_off[l], k2 = p.level_off[l + 2] - p.level_off[l + 1];
        LevelClass c;
        c.k2 = (uint32_t)k2; c.n1 = q.lvl_n1[l + 1]; c.m = q.lvl_m[l + 1]; c.z = q.lvl_z[l + 1]; c.dm = q.lvl_dm[l + 1];
        c.list = q.cls_list.data() + p.level_off[l + 1];
        c.mpre = q.mpre.data() + q.mpre_off[l + 1];
        const ProgCounts n = prog_counts(c);
        const bool compact = k <= PROG_COMPACT_K && k2 <= PROG_COMPACT_K && k <= shape.kn && k2 <= shape.kn &&
                             (int64_t)p.lvlW[l] * 64 <= (int64_t)PROG_COMPACT_DELTA;
        ProgHdr& h = q.hdr[l];
        h.k = (uint16_t)k; h.k2 = (uint16_t)k2;
        h.n_copy = (uint32_t)n.n_copy; h.n_multi = (uint32_t)n.n_multi; h.n_cand = (uint32_t)n.n_cand;
        h.n_big = (uint32_t)n.n_big; h.n_dead = (uint32_t)n.n_dead; h.max_n = n.max_n;
        uint32_t f = compact ? PF_COMPACT : 0u;
        if (k <= shape.kn) f |= PF_SRC_SMEM;
        if (k2 <= shape.kn) f |= PF_DST_SMEM;
        q.dir[l].flags = f;
        const ProgLayout lay = prog_layout(compact, n.n_copy, n.n_multi, n.n_cand, n.n_big, n.n_dead);
        h.off_cell = (uint32_t)lay.cell; h.off_cand = (uint32_t)lay.cand; h.off_big = (uint32_t)lay.big; h.off_dead = (uint32_t)lay.dead;
        if (lay.end > 0xFFFFFFFFull) bytes[l] = ~0ull; else
        bytes[l] = lay.end;
    }
    uint32_t cum = 0;
    int64_t kg = 0;
    for (int l = 0; l + 1 < L; ++l) {
        if (bytes[l] == ~0ull) { why = "a transition's program exceeds 4 GB"; return false; }
        q.prog_off[(size_t)l + 1] = q.prog_off[l] + bytes[l];
        ProgDir& d = q.dir[l];
        const bool ss = d.flags & PF_SRC_SMEM, ds = d.flags & PF_DST_SMEM;
        if (!ss) kg = std::max<int64_t>(kg, q.hdr[l].k);
        if (!ds) kg = std::max<int64_t>(kg, q.hdr[l].k2);
        if (!ss && !ds) { d.flags |= PF_ALL_CTAS | PF_WAIT | PF_ARRIVE; d.wait_target = cum; cum += (uint32_t)shape.grid; q.wide_list.push_back(l); }
        else if (ss && !ds) { d.flags |= PF_ARRIVE; cum += 1; }
        else if (!ss && ds) { d.flags |= PF_WAIT; d.wait_target = cum; }
        else ++q.n_smem_trans;
        if (q.prog_off[l] / 16 > 0xFFFFFFFFull) { why = "program larger than 64 GB"; return false; }
        d.off16 = (uint32_t)(q.prog_off[l] / 16);
        if (bytes[l] + sizeof(ProgDir) <= (uint64_t)shape.slot_bytes) { d.flags |= PF_STAGED; d.stage_bytes = (uint32_t)bytes[l]; }
        else d.stage_bytes = (uint32_t)sizeof(ProgHdr);
    }
    q.final_target = cum;
    q.prog_bytes = q.prog_off[(size_t)L - 1];
    for (int l = 0; l < L; ++l) {
        const int64_t nm = l >= 1 ? (int64_t)q.hdr[(size_t)l - 1].n_multi : 0;
        q.pred_off[(size_t)l + 1] = q.pred_off[l] + (int64_t)q.RL * nm;
        if (l >= 1) q.hdr[(size_t)l - 1].pred_off = (uint64_t)q.pred_off[l];
    }
    q.pred_elems = q.pred_off[L];
    q.gpad = 2 * kg * kg;
    q.gtile_cells = q.gpad + (int64_t)q.RL * kg * kg;
    return true;
}

ProgLevelIn plan4_level_in(const DipPlan& p, const Plan4& q, int l) {
    ProgLevelIn in;
    const int32_t mid = p.level_off[l + 1];
    in.k = (uint32_t)(mid - p.level_off[l]); in.k2 = (uint32_t)(p.level_off[l + 2] - mid);
    in.in_off = p.in_off.data() + mid;
    in.in_edge = p.in_edge.data();
    in.cls.k2 = in.k2; in.cls.n1 = q.lvl_n1[l + 1]; in.cls.m = q.lvl_m[l + 1]; in.cls.z = q.lvl_z[l + 1]; in.cls.dm = q.lvl_dm[l + 1];
    in.cls.list = q.cls_list.data() + mid;
    in.cls.mpre = q.mpre.data() + q.mpre_off[l + 1];
    in.W = p.lvlW[l];
    in.msrc = in.W ? p.masks.data() + p.msrc_off[l] : nullptr;
    in.mdst = in.W ? p.masks.data() + p.mdst_off[l] : nullptr;
    return in;
}

void prog_fill_level_host(const DipPlan& p, const Plan4& q, int l, uint8_t* out) {
    const ProgHdr& h = q.hdr[l];
    const bool compact = q.dir[l].flags & PF_COMPACT;
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
        const uint32_t dst = c.i2 * in.k2 + c.j2;
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
