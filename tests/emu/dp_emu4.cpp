// tests/emu/dp_emu4.cpp — TEST-ONLY CPU emulation of the level-program sweep (engine v4).
// Runs the exact host planning (dp_prep.cpp + dp_plan4.cpp), builds every transition's program with the very
// descriptor functions the device builder runs (dp_prog.h), then interprets the programs the way
// dip_sweep4_kernel does — same tile layouts (shared-memory tiles of stride 1 << slog with two dead padding layers,
// HBM tiles of stride k^2 behind gpad dead cells), same packed keys, same predecessor-code slots — with the thread
// grid replaced by serial loops.  `-m "not gpu"` tests check value, s_het, edge lists and the per-level checksums of
// every live cell against the oracle without a GPU.  Never part of the product library.
#include <algorithm>
#include <cstdint>
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
    int slog; int64_t gpad; int RL;
    std::vector<int32_t> s[2], g[2];
    // address of (layer r, cell idx) of level parity `par`, placement smem / HBM with layer stride kk
    int32_t& at(bool smem, int par, int64_t kk, int r, uint32_t idx) {
        if (smem) return s[par][(size_t)(((int64_t)(r + 2) << slog) + idx)];
        return g[par][(size_t)(gpad + (int64_t)r * kk + idx)];
    }
};

}  // namespace

// shape: [slog, kn, slot_bytes, grid, rc] (0 = default).  counts: [transitions in shared memory, transitions over all
// CTAs, compact programs, staged programs, big cells, program bytes, code elements, max candidates]
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
    if (!build_dip_plan(gv, p)) return -1;
    Sweep4Shape sh;
    int rc = 10;
    if (shape) {
        if (shape[0] > 0) sh.slog = shape[0];
        if (shape[1] > 0) sh.kn = shape[1];
        if (shape[2] > 0) sh.slot_bytes = shape[2];
        if (shape[3] > 0) sh.grid = shape[3];
        if (shape[4] > 0) rc = shape[4];
    }
    Plan4 q;
    std::string why;
    if (!plan4_build(p, sh, rc, q, why)) return -2;
    const int L = p.L, RL = q.RL;

    Tiles T;
    T.slog = sh.slog; T.gpad = q.gpad; T.RL = RL;
    for (int x = 0; x < 2; ++x) {
        T.s[x].assign((size_t)(RL + 2) << sh.slog, 0x5A5A5A5A);                 // (garbage: nothing may rely on zeros)
        std::fill(T.s[x].begin(), T.s[x].begin() + ((size_t)2 << sh.slog), V4_DEAD);
        T.g[x].assign((size_t)std::max<int64_t>(q.gtile_cells, 1), 0x5A5A5A5A);
        std::fill(T.g[x].begin(), T.g[x].begin() + (size_t)q.gpad, V4_DEAD);
    }
    // level 0: one vertex, every layer starts at 0 (approximator.cpp:535); it always lives in shared memory
    for (int r = 0; r < RL; ++r) T.at(true, 0, 1, r, 0) = r <= R ? 0 : V4_DEAD;

    std::vector<uint16_t> pred((size_t)std::max<int64_t>(q.pred_elems, 1), 0xABCD);
    std::vector<uint64_t> sum((size_t)L, FOLD_BASIS), live((size_t)L, 0);
    std::vector<uint8_t> prog;
    int64_t n_compact = 0, n_staged = 0, n_big = 0, n_all = 0;
    uint32_t cum = 0;
    for (int l = 0; l + 1 < L; ++l) {
        const ProgDir& d = q.dir[l];
        const size_t bytes = (size_t)(q.prog_off[(size_t)l + 1] - q.prog_off[l]);
        prog.assign(bytes, 0);
        prog_fill_level_host(p, q, l, prog.data());
        ProgHdr h;
        memcpy(&h, prog.data(), sizeof h);
        const bool compact = d.flags & PF_COMPACT, ss = d.flags & PF_SRC_SMEM, ds = d.flags & PF_DST_SMEM;
        n_compact += compact; n_staged += (d.flags & PF_STAGED) != 0; n_big += h.n_big; n_all += (d.flags & PF_ALL_CTAS) != 0;
        if ((d.flags & PF_STAGED) ? (d.stage_bytes != bytes || bytes + 16 > (size_t)sh.slot_bytes) : d.stage_bytes != sizeof(ProgHdr)) return -10;
        if ((uint64_t)d.off16 * 16 != q.prog_off[l]) return -11;
        // barrier schedule: a waiting level sees every arrival issued so far, no more
        if ((d.flags & PF_WAIT) && d.wait_target != cum) return -12;
        if (d.flags & PF_ARRIVE) cum += (d.flags & PF_ALL_CTAS) ? (uint32_t)sh.grid : 1u;
        if (((d.flags & PF_ALL_CTAS) != 0) != (!ss && !ds)) return -13;
        const ProgLayout lay = prog_layout(compact, h.n_copy, h.n_multi, h.n_cand, h.n_big, h.n_dead);
        if (lay.end != bytes) return -14;
        const int k = h.k, k2 = h.k2;
        const int64_t kk = (int64_t)k * k, kk2 = (int64_t)k2 * k2;
        if (ss && kk > ((int64_t)1 << sh.slog)) return -15;
        if (ds && kk2 > ((int64_t)1 << sh.slog)) return -15;
        const int sp = l & 1, dp = (l + 1) & 1;
        std::vector<uint8_t> written((size_t)kk2, 0);
        auto fold = [&](int r, uint32_t dst, int32_t val, uint32_t src) {
            if (r > R || val < 0) return;
            ++live[(size_t)l + 1];
            sum[(size_t)l + 1] += cell_fold((uint64_t)r * kk2 + dst, val >> V4_SHIFT, (int)(src / k), (int)(src % k));
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
            if (c.dst >= kk2 || c.src >= kk || written[c.dst]) return -20;
            written[c.dst] = 1;
            for (int r = 0; r < RL; ++r) {
                const int32_t v = T.at(ss, sp, kk, r - (int)c.w, c.src) + (int32_t)(c.delta << V4_SHIFT);
                T.at(ds, dp, kk2, r, c.dst) = v;
                fold(r, c.dst, v, c.src);
            }
        }
        uint32_t nb_seen = 0;
        for (uint32_t t = 0; t < h.n_multi; ++t) {
            CellDesc c;
            if (compact) { c.dst = wcell[2 * t] & 1023u; c.n = wcell[2 * t] >> 16; c.cand_off = wcell[2 * t + 1]; }
            else { c.dst = wcell[4 * t]; c.n = wcell[4 * t + 1]; c.cand_off = wcell[4 * t + 2]; }
            if (c.dst >= kk2 || written[c.dst] || c.n < 2 || c.n > PROG_MAX_CAND || (uint64_t)c.cand_off + c.n > h.n_cand) return -21;
            written[c.dst] = 1;
            if (c.n >= PROG_BIG_MIN) { if (nb_seen >= h.n_big || wbig[nb_seen] != t) return -22; ++nb_seen; }
            else if (c.n > h.max_n) return -23;
            for (int r = 0; r < RL; ++r) {
                int32_t key = V4_DEAD;
                for (uint32_t o = 0; o < c.n; ++o) {
                    CandDesc e;
                    if (compact) e = unpack_cand_c(wcand[c.cand_off + o]);
                    else { e.src = wcand[2 * (c.cand_off + o)] & 0x3FFFFFFFu; e.w = wcand[2 * (c.cand_off + o)] >> 30; e.delta = wcand[2 * (c.cand_off + o) + 1]; }
                    if (e.src >= kk) return -24;
                    const int32_t cand = T.at(ss, sp, kk, r - (int)e.w, e.src) + (int32_t)((e.delta << V4_SHIFT) + (V4_ORD_MASK - o));
                    key = std::max(key, cand);
                }
                const int32_t val = (int32_t)((uint32_t)key & ~V4_ORD_MASK);
                T.at(ds, dp, kk2, r, c.dst) = val;
                pred[(size_t)(h.pred_off + (uint64_t)r * h.n_multi + t)] = (uint16_t)key;
                if (val >= 0 && r <= R) {
                    const uint32_t o = V4_ORD_MASK - ((uint32_t)key & V4_ORD_MASK);
                    const uint32_t src = compact ? unpack_cand_c(wcand[c.cand_off + o]).src : (wcand[2 * (c.cand_off + o)] & 0x3FFFFFFFu);
                    fold(r, c.dst, val, src);
                }
            }
        }
        if (nb_seen != h.n_big) return -25;
        for (uint32_t x = 0; x < h.n_dead; ++x) {
            const uint32_t dst = wdead[x];
            if (dst >= kk2 || written[dst]) return -26;
            written[dst] = 1;
            for (int r = 0; r < RL; ++r) T.at(ds, dp, kk2, r, dst) = V4_DEAD;
        }
        for (uint8_t b : written) if (!b) return -27;
    }
    // sink cell (r = R, 0, 0) of the last level and the walk back through the codes
    const int32_t ks = p.level_off[L] - p.level_off[L - 1];
    const bool last_smem = ks <= sh.kn;
    const int32_t raw = T.at(last_smem, (L - 1) & 1, (int64_t)ks * ks, R, 0);
    *sink_value = raw < 0 ? NEG_INF : (raw >> V4_SHIFT);
    *sink_s_het = 0; *n_p1 = 0; *n_p2 = 0;
    if (level_checksum) for (int l = 0; l < L; ++l) { level_checksum[l] = sum[l]; level_live[l] = live[l]; }
    if (counts) {
        counts[0] = q.n_smem_trans; counts[1] = n_all; counts[2] = n_compact; counts[3] = n_staged; counts[4] = n_big;
        counts[5] = (int64_t)q.prog_bytes; counts[6] = q.pred_elems; counts[7] = q.max_cand;
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
    if (!build_dip_plan(gv, p)) return -1;
    Sweep4Shape sh;
    int rc = 10;
    if (shape) {
        if (shape[0] > 0) sh.slog = shape[0];
        if (shape[1] > 0) sh.kn = shape[1];
        if (shape[2] > 0) sh.slot_bytes = shape[2];
        if (shape[3] > 0) sh.grid = shape[3];
        if (shape[4] > 0) rc = shape[4];
    }
    Plan4 q;
    std::string why;
    if (!plan4_build(p, sh, rc, q, why)) return -2;
    if (out && cap >= (int64_t)q.prog_bytes) {
#pragma omp parallel for schedule(dynamic, 64)
        for (int l = 0; l < p.L - 1; ++l) prog_fill_level_host(p, q, l, out + q.prog_off[l]);
    }
    return (int64_t)q.prog_bytes;
}
