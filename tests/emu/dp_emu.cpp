// tests/emu/dp_emu.cpp — TEST-ONLY CPU emulation of the CUDA diploid sweep.
// Runs the exact host planning (dp_prep.cpp) and the exact per-cell / traceback code (dp_cell.h)
// the kernels run, with the thread grid replaced by a serial loop, so that `-m "not gpu"` tests can
// check the gather formulation, the mask construction, the predecessor codes and the traceback
// against the oracle without a GPU.  Never part of the product library.
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../dipgenie_b200/csrc/cuda/dp_cell.h"
#include "../../dipgenie_b200/csrc/cuda/dp_prep.h"

using namespace dg;

template <class PredT>
static int run(const DipPlan& p, int32_t* sink_value, int32_t* sink_s_het, int32_t* p1, int32_t* n1,
               int32_t* p2, int32_t* n2, uint64_t* level_checksum, uint64_t* level_live) {
    const int R = p.R, L = p.L;
    std::vector<PredT> pred((size_t)p.pred_off[L]);
    std::vector<int32_t> cur((size_t)(R + 1), 0), next;
    constexpr int SH = (sizeof(PredT) == 2) ? 8 : 16;
    for (int l = 0; l + 1 < L; ++l) {
        Transition t;
        t.k = p.level_off[l + 1] - p.level_off[l];
        t.k2 = p.level_off[l + 2] - p.level_off[l + 1];
        t.W = p.lvlW[l];
        t.in_off = p.in_off.data() + p.level_off[l + 1];
        t.in_edge = p.in_edge.data();
        t.msrc = p.masks.data() + p.msrc_off[l];
        t.mdst = p.masks.data() + p.mdst_off[l];
        const size_t ncell = (size_t)(R + 1) * t.k2 * t.k2;
        next.assign(ncell, NEG_INF);
        uint64_t h = FOLD_BASIS, live = 0;
        const int32_t* src = cur.data();
        for (size_t c = 0; c < ncell; ++c) {
            const int j2 = (int)(c % t.k2), i2 = (int)((c / t.k2) % t.k2), r2 = (int)(c / ((size_t)t.k2 * t.k2));
            uint32_t code;
            const uint64_t key = relax_cell(t, [src](int64_t idx) { return src[idx]; }, r2, i2, j2, code);
            next[c] = key_value(key);
            pred[(size_t)p.pred_off[l + 1] + c] = key ? (PredT)(((code >> 16) << SH) | (code & 0xFFFFu)) : (PredT)~(PredT)0;
            if (key) {
                ++live;
                h += cell_fold(c, next[c], 0xFFFF - (int)((key >> 16) & 0xFFFF), 0xFFFF - (int)(key & 0xFFFF));
            }
        }
        if (level_checksum) { level_checksum[l + 1] = h; level_live[l + 1] = live; }
        cur.swap(next);
    }
    const size_t ks = (size_t)(p.level_off[L] - p.level_off[L - 1]);
    *sink_value = cur[(size_t)R * ks * ks];   // cell (r=R,0,0) of the last level (:730, :775)
    TraceView v;
    v.L = L; v.R = R; v.level_off = p.level_off.data(); v.in_off = p.in_off.data(); v.in_edge = p.in_edge.data();
    v.lvlW = p.lvlW.data(); v.msrc_off = p.msrc_off.data(); v.mdst_off = p.mdst_off.data();
    v.masks = p.masks.data(); v.pred_off = p.pred_off.data();
    std::vector<int32_t> a(2 * (R + 2)), b(2 * (R + 2));
    int rc = traceback<PredT>(v, pred.data(), *sink_value, a.data(), n1, b.data(), n2, R + 2, sink_s_het);
    if (rc == -1) { *n1 = 0; *n2 = 0; return 0; }
    if (rc) return rc;
    for (int x = 0; x < *n1; ++x) { p1[2 * x] = a[2 * (*n1 - 1 - x)]; p1[2 * x + 1] = a[2 * (*n1 - 1 - x) + 1]; }
    for (int x = 0; x < *n2; ++x) { p2[2 * x] = b[2 * (*n2 - 1 - x)]; p2[2 * x + 1] = b[2 * (*n2 - 1 - x) + 1]; }
    return 0;
}

extern "C" int emu_dp_diploid(int32_t n_levels, const int32_t* level_off, const int64_t* adj_off,
                              const int32_t* adj_dst, const uint8_t* adj_w, const int64_t* col_off,
                              const int32_t* col_val, const uint8_t* colour_is_hom, int32_t n_colours, int32_t R,
                              int32_t* sink_value, int32_t* sink_s_het, int32_t* p1_edges, int32_t* n_p1,
                              int32_t* p2_edges, int32_t* n_p2, uint64_t* level_checksum, uint64_t* level_live,
                              int32_t force_pred32) {
    DipGraphView g;
    g.n_levels = n_levels; g.level_off = level_off; g.adj_off = adj_off; g.adj_dst = adj_dst; g.adj_w = adj_w;
    g.col_off = col_off; g.col_val = col_val; g.colour_is_hom = colour_is_hom; g.n_colours = n_colours; g.R = R;
    DipPlan p;
    if (!build_dip_plan(g, p)) return -1;
    plan_participants(p, 148, 2048);
    if (p.max_indeg <= 255 && !force_pred32)
        return run<uint16_t>(p, sink_value, sink_s_het, p1_edges, n_p1, p2_edges, n_p2, level_checksum, level_live);
    return run<uint32_t>(p, sink_value, sink_s_het, p1_edges, n_p1, p2_edges, n_p2, level_checksum, level_live);
}
