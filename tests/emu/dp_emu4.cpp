// tests/emu/dp_emu4.cpp — TEST-ONLY CPU emulation of the level-program sweep (engine v4).
// Runs the exact host planning (dp_prep.cpp + dp_plan4.cpp), builds every transition's program with the very
// descriptor functions the device builder runs (dp_prog.h), then interprets the programs the way
// dip_sweep4_kernel does — same in-place tiles (one shared-memory tile of layer stride 1 << slog, one HBM tile of layer
// stride hstride^2, two dead padding layers below layer 0), same slots, same packed keys, same predecessor-code
// slots, the timed directory's skipping of idle transitions — with the thread grid replaced by serial loops.  It also
// checks what the design relies on: no cell that a transition reads is written by it, every written cell exactly once.
// `-m "not gpu"` tests check value, s_het, edge lists and the per-level checksums of every live cell against the
// oracle without a GPU.  Never part of the product library.
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../dipgenie_b200/csrc/cuda/dp_cell.h"
#include "../../dipgenie_b200/csrc/cuda/dp_plan4.h"
#include "../../dipgenie_b200/csrc/cuda/dp_prep.h"
#include "../../dipgenie_b200/csrc/cuda/dp_prog.h"

using namespace dg;

namespace {

struct Tiles {
    int slog; int64_t gpad, hkk;
    std::vector<int32_t> s, g;
    int32_t& at(bool smem, int r, uint32_t idx) {
        if (smem) return s[(size_t)(((int64_t)(r + 2) << slog) + idx)];
        return g[(size_t)(gpad + (int64_t)r * hkk + idx)];
    }
};

bool plan(const DipGraphView& gv, const int32_t* shape, DipPlan& p, Plan4& q) {
    if (!build_dip_plan(gv, p)) { if (getenv("EMU4_VERBOSE")) fprintf(stderr, "build_dip_plan: %s\n", p.error.c_str()); return false; }
    Sweep4Shape sh;
    int rc = 10;
    if (shape) {
        if (shape[0] > 0) sh.slog = shape[0];
        if (shape[1] > 0) sh.kn = shape[1];
        if (shape[2] > 0) sh.slot_bytes = shape[2];
        if (shape[3] > 0) sh.grid = shape[3];
        if (shape[4] > 0) rc = shape[4];
    }
    std::string why;
    const bool ok = plan4_build(p, gv, sh, rc, q, why);
    if (!ok && getenv("EMU4_VERBOSE")) fprintf(stderr, "plan4_build: %s\n", why.c_str());
    return ok;
}

}  // namespace

// shape: [slog, kn, slot_bytes, grid, rc] (0 = default).  counts: [transitions in shared memory, transitions over all
// CTAs, compact programs, staged programs, big cells, program bytes, code elements, max candidates, relocations,
// skipped transitions, cells written, cells of all levels]
extern "C" int emu4_dp_diploid(int32_t n_levels, const int32_t* level_off, const int64_t* adj_off,
                               const int32_t* adj_dst, const uint8_t* adj_w, const int64_t* col_off,
                               const int32_t* col_val, const uint8_t* colour_is_hom, int32_t n_colours, int32_t R,
                               int32_t* sink_value, int32_t* sink_s_het, int32_t* p1_edges, int32_t* n_p1,
                               int32_t* p2_edges, int32_t* n_p2, uint64_t* level_checksum, uint64_t* level_live,
                               const int32_t* shape, int64_t* counts) {
    DipGraphView gv;
    gv.n_levels = n_levels; gv.level_off = level_off; gv.adj_off = adj_off; gv.adj_dst = adj_dst; gv.adj_w = adj_w;
    gv.col_off = col_off; gv.col_val = col_val; gv.colour_is_hom = colour_is_hom; gv.n_colours = n_colours; gv.R = R;
    DipPlan p;
    Plan4 q;
    {
        DipPlan probe;
        if (!build_dip_plan(gv, probe)) return -1;
    }
    if (!plan(gv, shape, p, q)) return -2;
    const Sweep4Shape& sh = q.shape;
    const int L = p.L, RL = q.RL;

    Tiles T;
    T.slog = sh.slog; T.hkk = (int64_t)q.hstride * q.hstride; T.gpad = 2 * T.hkk;      // (the emulator keeps its own layer-major model of the HBM tile)
    T.s.assign((size_t)(RL + 2) << sh.slog, 0x5A5A5A5A);                 // (garbage: nothing may rely on zeros)
    std::fill(T.s.begin(), T.s.begin() + ((size_t)2 << sh.slog), V4_DEAD);
    T.g.assign((size_t)std::max<int64_t>(T.gpad + (int64_t)RL * T.hkk, 1), 0x5A5A5A5A);
    std::fill(T.g.begin(), T.g.begin() + (size_t)T.gpad, V4_DEAD);
    // level 0: one vertex in slot 0 of the shared-memory tile, every layer starts at 0 (approximator.cpp:535)
    for (int r = 0; r < RL; ++r) T.at(true, r, 0) = r <= R ? 0 : V4_DEAD;

    std::vector<uint16_t> pred((size_t)std::max<int64_t>(q.pred_elems, 1), 0xABCD);
    std::vector<uint64_t> sum((size_t)L, FOLD_BASIS), live((size_t)L, 0);
    std::vector<uint8_t> prog;
    int64_t n_compact = 0, n_staged = 0, n_big = 0, n_all = 0;
    // the two directories: same transitions in level order, the timed one without the idle ones
    std::vector<int> timed_at((size_t)L, -1);
    for (size_t x = 0; x < q.timed.dir.size(); ++x) timed_at[(size_t)q.timed.dir[x].level] = (int)x;
    if ((int)q.full.dir.size() != L - 1) return -9;
    for (const Plan4Dir* D : {&q.full, &q.timed}) {                     // barrier schedules
        uint32_t cum = 0;
        size_t nw = 0;
        for (size_t x = 0; x < D->dir.size(); ++x) {
            const ProgDir& d = D->dir[x];
            const bool ss = d.flags & PF_SRC_SMEM, ds = d.flags & PF_DST_SMEM;
            if ((d.flags & PF_WAIT) && d.wait_target != cum) return -12;
            if (d.flags & PF_ARRIVE) cum += (d.flags & PF_ALL_CTAS) ? (uint32_t)sh.grid : 1u;
            if (((d.flags & PF_ALL_CTAS) != 0) != (!ss && !ds)) return -13;
            if (!ss && ds && !(d.flags & PF_WAIT)) return -13;
            if (ss && !ds && !(d.flags & PF_ARRIVE)) return -13;
            if (d.flags & PF_ALL_CTAS) { if (nw >= D->wide_list.size() || D->wide_list[nw] != (int32_t)x) return -13; ++nw; }
        }
        if (nw != D->wide_list.size() || cum != D->final_target) return -13;
    }
    std::vector<uint8_t> written, readset;
    for (int l = 0; l + 1 < L; ++l) {
        const ProgDir& d = q.full.dir[l];
        if (d.level != l) return -9;
        const size_t bytes = (size_t)d.prog16 * 16, room = (size_t)(q.prog_off[(size_t)l + 1] - q.prog_off[l]);
        if (room < bytes || room - bytes >= PROG_ALIGN) return -11;      // programs start PROG_ALIGN-aligned
        prog.assign(bytes, 0);
        prog_fill_level_host(p, q, l, prog.data());
        ProgHdr h;
        memcpy(&h, prog.data(), sizeof h);
        const bool compact = d.flags & PF_COMPACT, ss = d.flags & PF_SRC_SMEM, ds = d.flags & PF_DST_SMEM, reloc = d.flags & PF_RELOCATE;
        if ((d.flags & PF_STAGED) ? (d.stage_bytes != bytes || bytes + sizeof(ProgDir) > (size_t)sh.slot_bytes) : d.stage_bytes != sizeof(ProgHdr)) return -10;
        if ((uint64_t)d.off64 * PROG_ALIGN != q.prog_off[l]) return -11;
        if (reloc != (ss != ds)) return -16;
        if (ss != (q.lvl_dom[l] == 0) || ds != (q.lvl_dom[l + 1] == 0)) return -16;
        const ProgLayout lay = prog_layout(compact, h.n_copy, h.n_multi, h.n_cand, h.n_big, h.n_dead);
        if (lay.end != bytes || h.off_cell != lay.cell || h.off_cand != lay.cand || h.off_big != lay.big || h.off_dead != lay.dead) return -14;
        const int k = h.k, k2 = h.k2;
        const int64_t kk2 = (int64_t)k2 * k2;
        const ProgLevelIn in = plan4_level_in(p, q, l);
        const int64_t cap_s = ss ? (int64_t)sh.kn * sh.kn : T.hkk, cap_d = ds ? (int64_t)sh.kn * sh.kn : T.hkk;
        const bool idle = timed_at[l] < 0;
        if (idle && (reloc || h.n_copy + h.n_multi + h.n_dead != 0)) return -17;
        if (!idle) {
            const ProgDir& dt = q.timed.dir[(size_t)timed_at[l]];
            if (dt.off64 != d.off64 || dt.stage_bytes != d.stage_bytes || ((dt.flags ^ d.flags) & ~(uint32_t)(PF_WAIT | PF_ARRIVE)) != 0) return -17;
            n_compact += compact; n_staged += (d.flags & PF_STAGED) != 0; n_big += h.n_big; n_all += (d.flags & PF_ALL_CTAS) != 0;
        }
        written.assign((size_t)cap_d, 0);
        readset.assign((size_t)cap_s, 0);
        auto fold = [&](int r, uint32_t i2, uint32_t j2, int32_t val, uint32_t i, uint32_t j) {
            if (r > R || val < 0) return;
            ++live[(size_t)l + 1];
            sum[(size_t)l + 1] += cell_fold((uint64_t)r * kk2 + (uint64_t)i2 * k2 + j2, val >> V4_SHIFT, (int)i, (int)j);
        };
        const uint32_t* wcopy = reinterpret_cast<const uint32_t*>(prog.data() + lay.copy);
        const uint32_t* wcell = reinterpret_cast<const uint32_t*>(prog.data() + lay.cell);
        const uint32_t* wcand = reinterpret_cast<const uint32_t*>(prog.data() + lay.cand);
        const uint32_t* wbig = reinterpret_cast<const uint32_t*>(prog.data() + lay.big);
        const uint32_t* wdead = reinterpret_cast<const uint32_t*>(prog.data() + lay.dead);
        for (uint32_t t = 0; t < h.n_copy; ++t) {
            CopyDesc c;
            if (compact) c = unpack_copy_c(wcopy[t]);
            else { c.src = wcopy[4 * t] & 0x3FFFFFFFu; c.w = wcopy[4 * t] >> 30; c.dst = wcopy[4 * t + 1]; c.delta = wcopy[4 * t + 2]; }
            if (c.dst >= cap_d || c.src >= cap_s || written[c.dst]) return -20;
            written[c.dst] = 1; readset[c.src] = 1;
            uint32_t i2, j2, i, wi, j, wj;
            copy_pair(in, t, i2, j2);
            in_edge_at(in, i2, 0, i, wi); in_edge_at(in, j2, 0, j, wj);
            std::vector<int32_t> vals((size_t)RL);
            for (int r = 0; r < RL; ++r) vals[r] = T.at(ss, r - (int)c.w, c.src) + (int32_t)(c.delta << V4_SHIFT);
            for (int r = 0; r < RL; ++r) { T.at(ds, r, c.dst) = vals[r]; fold(r, i2, j2, vals[r], i, j); }
        }
        uint32_t nb_seen = 0;
        std::vector<int32_t> keys((size_t)RL);
        for (uint32_t t = 0; t < h.n_multi; ++t) {
            CellDesc c;
            if (compact) { c.dst = wcell[2 * t] & 1023u; c.n = wcell[2 * t] >> 16; c.cand_off = wcell[2 * t + 1]; }
            else { c.dst = wcell[4 * t]; c.n = wcell[4 * t + 1]; c.cand_off = wcell[4 * t + 2]; }
            if (c.dst >= cap_d || written[c.dst] || c.n < 2 || c.n > PROG_MAX_CAND || (uint64_t)c.cand_off + c.n > h.n_cand) return -21;
            written[c.dst] = 1;
            if (c.n >= PROG_BIG_MIN) { if (nb_seen >= h.n_big || wbig[nb_seen] != t) return -22; ++nb_seen; }
            else if (c.n > h.max_n) return -23;
            const MultiCell mc = multi_cell(in, t);
            const bool striped = c.n > PROG_KEY_CAND;       // the kernel's warp form: round in the key, lane from a ballot, code = ordinal
            std::vector<uint32_t> ords((size_t)RL, 0);
            for (int r = 0; r < RL; ++r) {
                int32_t key = V4_DEAD;
                for (uint32_t o = 0; o < c.n; ++o) {
                    CandDesc e;
                    if (compact) e = unpack_cand_c(wcand[c.cand_off + o]);
                    else { e.src = wcand[2 * (c.cand_off + o)] & 0x3FFFFFFFu; e.w = wcand[2 * (c.cand_off + o)] >> 30; e.delta = wcand[2 * (c.cand_off + o) + 1]; }
                    if (e.src >= cap_s) return -24;
                    readset[e.src] = 1;
                    if (striped) {      // first strict maximum of the value, ordinal kept beside it
                        const int32_t cv = T.at(ss, r - (int)e.w, e.src) + (int32_t)(e.delta << V4_SHIFT);
                        if (o == 0 || cv > key) { key = cv; ords[r] = o; }
                        continue;
                    }
                    const int32_t cand = T.at(ss, r - (int)e.w, e.src) + (int32_t)((e.delta << V4_SHIFT) + (V4_ORD_MASK - o));
                    key = std::max(key, cand);
                }
                keys[r] = key;
            }
            for (int r = 0; r < RL; ++r) {
                const int32_t key = keys[r], val = (int32_t)((uint32_t)key & ~V4_ORD_MASK);
                T.at(ds, r, c.dst) = val;
                pred[(size_t)(h.pred_off + (uint64_t)r * h.n_multi + t)] = striped ? (uint16_t)ords[r] : (uint16_t)key;
                if (val >= 0 && r <= R) {
                    const uint32_t o = striped ? ords[r] : V4_ORD_MASK - ((uint32_t)key & V4_ORD_MASK);
                    const uint32_t e1 = o / mc.d2, e2 = o - e1 * mc.d2;
                    uint32_t i, wi, j, wj;
                    in_edge_at(in, mc.i2, e1, i, wi); in_edge_at(in, mc.j2, e2, j, wj);
                    fold(r, mc.i2, mc.j2, val, i, j);
                }
            }
        }
        if (nb_seen != h.n_big) return -25;
        for (uint32_t x = 0; x < h.n_dead; ++x) {
            const uint32_t dst = wdead[x];
            if (dst >= cap_d || written[dst]) return -26;
            written[dst] = 1;
            for (int r = 0; r < RL; ++r) T.at(ds, r, dst) = V4_DEAD;
        }
        // passive pairs: the same memory in both levels, untouched
        if ((uint64_t)h.n_copy + h.n_multi + h.n_dead + (uint64_t)h.n_passive * h.n_passive != (uint64_t)kk2) return -27;
        for (uint64_t x = 0; x < (uint64_t)h.n_passive * h.n_passive; ++x) {
            uint32_t i2, j2, i, wi, j, wj;
            passive_pair(in, x, i2, j2);
            in_edge_at(in, i2, 0, i, wi); in_edge_at(in, j2, 0, j, wj);
            const uint32_t dc = dst_cell(in, i2, j2);
            if (dc != src_cell(in, i, j) || wi + wj != 0 || written[dc]) return -28;
            written[dc] = 1;
            for (int r = 0; r < RL; ++r) fold(r, i2, j2, T.at(ds, r, dc), i, j);
        }
        if (ss == ds)                                                  // in place: nothing that was read may have been written
            for (size_t x = 0; x < readset.size(); ++x) if (readset[x] && written[x] && !(x < (size_t)cap_d && false)) {
                // (a passive cell is "written" in the bookkeeping above but not in memory: reads of it are fine)
                bool passive_cell = false;
                for (uint64_t y = 0; y < (uint64_t)h.n_passive * h.n_passive && !passive_cell; ++y) {
                    uint32_t i2, j2; passive_pair(in, y, i2, j2);
                    passive_cell = dst_cell(in, i2, j2) == x;
                }
                if (!passive_cell) return -29;
            }
        (void)k;
    }
    // sink cell (r = R, 0, 0) of the last level and the walk back through the codes
    const bool last_smem = q.last_smem;
    if (last_smem != (q.lvl_dom[L - 1] == 0) || q.full.n != L - 1 || q.timed.n != (int32_t)q.timed.dir.size()) return -9;
    const int32_t raw = T.at(last_smem, R, q.sink_cell);
    *sink_value = raw < 0 ? NEG_INF : (raw >> V4_SHIFT);
    *sink_s_het = 0; *n_p1 = 0; *n_p2 = 0;
    if (level_checksum) for (int l = 0; l < L; ++l) { level_checksum[l] = sum[l]; level_live[l] = live[l]; }
    if (counts) {
        counts[0] = q.n_smem_trans; counts[1] = n_all; counts[2] = n_compact; counts[3] = n_staged; counts[4] = n_big;
        counts[5] = (int64_t)q.prog_bytes; counts[6] = q.pred_elems; counts[7] = q.max_cand; counts[8] = q.n_relocate; counts[9] = q.n_skipped;
        counts[10] = (int64_t)q.cells_written; counts[11] = (int64_t)q.cells_total;
    }
    if (raw < 0) return 0;
    TraceView v;
    v.L = L; v.R = R; v.level_off = p.level_off.data(); v.in_off = p.in_off.data(); v.in_edge = p.in_edge.data();
    v.lvlW = p.lvlW.data(); v.msrc_off = p.msrc_off.data(); v.mdst_off = p.mdst_off.data(); v.masks = p.masks.data();
    v.pred_off = q.pred_off.data();
    v.vinfo = q.vinfo.data(); v.lvl_n1 = q.lvl_n1.data(); v.lvl_m = q.lvl_m.data(); v.RL = RL;
    TraceState s = {R, 0, 0};
    const int cap = R + 2;
    std::vector<int32_t> a((size_t)2 * cap), b((size_t)2 * cap);
    int32_t n1 = 0, n2 = 0, sh_het = 0;
    const int rc2 = trace_segment<uint16_t>(v, pred.data(), L - 1, 0, s, a.data(), &n1, b.data(), &n2, cap, &sh_het);
    if (rc2) return -30 + rc2;
    *sink_s_het = sh_het; *n_p1 = n1; *n_p2 = n2;
    for (int x = 0; x < n1; ++x) { p1_edges[2 * x] = a[2 * (n1 - 1 - x)]; p1_edges[2 * x + 1] = a[2 * (n1 - 1 - x) + 1]; }
    for (int x = 0; x < n2; ++x) { p2_edges[2 * x] = b[2 * (n2 - 1 - x)]; p2_edges[2 * x + 1] = b[2 * (n2 - 1 - x) + 1]; }
    return 0;
}

// The whole program as the HOST builder writes it (prog_fill_level_host), for the byte-for-byte comparison with what
// the device builder (prog_fill_kernel) wrote.  shape as above.  Returns the size; fills `out` when it is large enough.
extern "C" int64_t emu4_build_program(int32_t n_levels, const int32_t* level_off, const int64_t* adj_off,
                                      const int32_t* adj_dst, const uint8_t* adj_w, const int64_t* col_off,
                                      const int32_t* col_val, const uint8_t* colour_is_hom, int32_t n_colours, int32_t R,
                                      const int32_t* shape, uint8_t* out, int64_t cap) {
    DipGraphView gv;
    gv.n_levels = n_levels; gv.level_off = level_off; gv.adj_off = adj_off; gv.adj_dst = adj_dst; gv.adj_w = adj_w;
    gv.col_off = col_off; gv.col_val = col_val; gv.colour_is_hom = colour_is_hom; gv.n_colours = n_colours; gv.R = R;
    DipPlan p;
    Plan4 q;
    {
        DipPlan probe;
        if (!build_dip_plan(gv, probe)) return -1;
    }
    if (!plan(gv, shape, p, q)) return -2;
    if (out && cap >= (int64_t)q.prog_bytes) {
#pragma omp parallel for schedule(dynamic, 64)
        for (int l = 0; l < p.L - 1; ++l) prog_fill_level_host(p, q, l, out + q.prog_off[l]);
    }
    return (int64_t)q.prog_bytes;
}

// Per-transition plan statistics (tools/plan_stats.py): out[l * 8 ..] = flags, k, k2, n_copy, n_multi, n_cand, n_big, n_dead.
extern "C" int64_t emu4_plan_stats(int32_t n_levels, const int32_t* level_off, const int64_t* adj_off,
                                   const int32_t* adj_dst, const uint8_t* adj_w, const int64_t* col_off,
                                   const int32_t* col_val, const uint8_t* colour_is_hom, int32_t n_colours, int32_t R,
                                   const int32_t* shape, int64_t* out) {
    DipGraphView gv;
    gv.n_levels = n_levels; gv.level_off = level_off; gv.adj_off = adj_off; gv.adj_dst = adj_dst; gv.adj_w = adj_w;
    gv.col_off = col_off; gv.col_val = col_val; gv.colour_is_hom = colour_is_hom; gv.n_colours = n_colours; gv.R = R;
    DipPlan p;
    Plan4 q;
    const bool ok = plan(gv, shape, p, q);
    if (!ok && (int)q.hdr.size() != p.L - 1) return -2;      // (a plan rejected for its size still has its headers: report them)
    for (int l = 0; l < p.L - 1; ++l) {
        const ProgHdr& h = q.hdr[l];
        int64_t* o = out + (size_t)l * 8;
        o[0] = ok ? q.full.dir[l].flags : q.tflags[l]; o[1] = h.k; o[2] = h.k2; o[3] = h.n_copy; o[4] = h.n_multi; o[5] = h.n_cand; o[6] = h.n_big; o[7] = h.n_dead;
    }
    return ok ? (int64_t)q.prog_bytes : -3;
}
