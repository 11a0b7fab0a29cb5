// tests/emu/dp_emu.cpp — TEST-ONLY CPU emulation of the CUDA diploid sweep.
// Runs the exact host planning (dp_prep.cpp: in-edge CSR, colour masks, modes, packed records, barrier
// schedule) and the exact per-cell / record-view / traceback code (dp_cell.h) the kernels run, with the
// thread grid replaced by serial loops and the shared-memory tiles / record stages by host buffers, so
// that `-m "not gpu"` tests can check the gather formulation, the record packing, the tile placement
// flags, the monotone barrier targets, the predecessor codes and the traceback against the oracle
// without a GPU.  Never part of the product library.
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../dipgenie_b200/csrc/cuda/dp_cell.h"
#include "../../dipgenie_b200/csrc/cuda/dp_prep.h"

using namespace dg;

template <int RC, bool HAS_MASK, class PredT, class OffT, class Load, class Store>
static void items(const TransitionT<OffT>& t, int R, Load load, Store store, PredT* pl, uint64_t& h, uint64_t& live) {
    constexpr int SH = (sizeof(PredT) == 2) ? 8 : 16;
    const uint32_t k2 = (uint32_t)t.k2, npairs = k2 * k2, nchunk = (uint32_t)(R + RC) / RC;
    for (uint64_t x = 0; x < (uint64_t)npairs * nchunk; ++x) {
        const uint32_t chunk = (uint32_t)(x / npairs), pair = (uint32_t)(x - (uint64_t)chunk * npairs);
        const uint32_t i2 = pair / k2, j2 = pair - i2 * k2;
        uint64_t best[RC]; uint32_t code[RC];
        relax_pair<RC, HAS_MASK>(t, load, R, (int)chunk * RC, (int)i2, (int)j2, best, code);
        for (int rr = 0; rr < RC; ++rr) {
            const int r2 = (int)chunk * RC + rr;
            if (r2 > R) continue;
            const uint64_t c = (uint64_t)r2 * npairs + pair, key = best[rr];
            store(c, key_value(key));
            pl[c] = key ? (PredT)(((code[rr] >> 16) << SH) | (code[rr] & 0xFFFFu)) : (PredT) ~(PredT)0;
            if (key) {
                ++live;
                h += cell_fold(c, key_value(key), 0xFFFF - (int)((key >> 16) & 0xFFFF), 0xFFFF - (int)(key & 0xFFFF));
            }
        }
    }
}

// same dispatch as sweep_dispatch() in dp_diploid.cu
template <class PredT, class OffT, class Load, class Store>
static void cells(const TransitionT<OffT>& t, int R, uint64_t nthreads, Load load, Store store, PredT* pl, uint64_t& h, uint64_t& live) {
    const int rc = choose_rc((uint64_t)t.k2 * t.k2, R, nthreads);
    if (t.W > 0) {
        switch (rc) {
            case 8: items<8, true>(t, R, load, store, pl, h, live); break;
            case 4: items<4, true>(t, R, load, store, pl, h, live); break;
            case 2: items<2, true>(t, R, load, store, pl, h, live); break;
            default: items<1, true>(t, R, load, store, pl, h, live); break;
        }
    } else {
        switch (rc) {
            case 8: items<8, false>(t, R, load, store, pl, h, live); break;
            case 4: items<4, false>(t, R, load, store, pl, h, live); break;
            case 2: items<2, false>(t, R, load, store, pl, h, live); break;
            default: items<1, false>(t, R, load, store, pl, h, live); break;
        }
    }
}

template <class PredT>
static int run(const DipPlan& p, const SweepShape& sh, int32_t* sink_value, int32_t* sink_s_het, int32_t* p1,
               int32_t* n1, int32_t* p2, int32_t* n2, uint64_t* level_checksum, uint64_t* level_live) {
    const int R = p.R, L = p.L;
    std::vector<PredT> pred((size_t)p.pred_off[L]);
    const size_t widest = (size_t)(R + 1) * p.kmax * p.kmax;
    // poison values make a wrong tile-placement flag visible
    std::vector<int32_t> g0(widest, 0x5A5A5A5A), g1(widest, 0x5A5A5A5A), s0(sh.tile_cells, 0x3C3C3C3C), s1(sh.tile_cells, 0x3C3C3C3C);
    for (int r = 0; r <= R; ++r) { g0[r] = 0; s0[r] = 0; }
    std::vector<uint8_t> stage((size_t)sh.stage_bytes + 16);
    uint32_t counter = 0;
    for (int l = 0; l + 1 < L; ++l) {
        uint64_t h = FOLD_BASIS, live = 0;
        PredT* pl = pred.data() + p.pred_off[l + 1];
        if (p.mode[l] != MODE_GLOBAL) {
            RecHeader hd;
            memcpy(&hd, p.records.data() + p.rec_off[l], sizeof hd);
            if (hd.bytes > (uint32_t)sh.stage_bytes || hd.bytes % 16 != 0 || p.rec_off[l] % 16 != 0) return -10;
            memcpy(stage.data(), p.records.data() + p.rec_off[l], hd.bytes);
            TransitionT<uint16_t> t;
            record_view(stage.data(), hd, t);
            if ((hd.flags & REC_WAIT) && counter < hd.wait_target) return -11;       // would dead-lock on the GPU
            const bool ssm = hd.flags & REC_SRC_SMEM, dsm = hd.flags & REC_DST_SMEM;
            if ((ssm || dsm) && (p.mode[l] != MODE_FAST || hd.P != 1)) return -12;
            const int32_t* src = ssm ? ((l & 1) ? s1.data() : s0.data()) : ((l & 1) ? g1.data() : g0.data());
            int32_t* dst = dsm ? ((l & 1) ? s0.data() : s1.data()) : ((l & 1) ? g0.data() : g1.data());
            if ((ssm && (size_t)(R + 1) * t.k * t.k > (size_t)sh.tile_cells) ||
                (dsm && (size_t)(R + 1) * t.k2 * t.k2 > (size_t)sh.tile_cells)) return -13;
            if (hd.pred_off2 != p.pred_off[l + 1]) return -14;
            cells<PredT>(t, R, (uint64_t)hd.P * sh.cells_per_cta / 4, [src](int64_t i) { return src[i]; }, [dst](size_t c, int32_t v) { dst[c] = v; }, pl, h, live);
            if (hd.flags & REC_ARRIVE) counter += hd.P;
        } else {
            Transition t;
            t.k = p.level_off[l + 1] - p.level_off[l];
            t.k2 = p.level_off[l + 2] - p.level_off[l + 1];
            t.W = p.lvlW[l];
            t.in_off = p.in_off.data() + p.level_off[l + 1];
            t.in_edge = p.in_edge.data();
            t.msrc = p.masks.data() + p.msrc_off[l];
            t.mdst = p.masks.data() + p.mdst_off[l];
            if ((p.flags[l] & REC_WAIT) && counter < p.bar_target[l - 1]) return -11;
            const int32_t* src = (l & 1) ? g1.data() : g0.data();
            int32_t* dst = (l & 1) ? g0.data() : g1.data();
            cells<PredT>(t, R, (uint64_t)p.P[l] * sh.cells_per_cta / 4, [src](int64_t i) { return src[i]; }, [dst](size_t c, int32_t v) { dst[c] = v; }, pl, h, live);
            if (p.flags[l] & REC_ARRIVE) counter += (uint32_t)p.P[l];
        }
        if (counter != p.bar_target[l]) return -15;
        if (level_checksum) { level_checksum[l + 1] = h; level_live[l + 1] = live; }
    }
    const size_t ks = (size_t)(p.level_off[L] - p.level_off[L - 1]);
    const std::vector<int32_t>& last = ((L - 1) & 1) ? g1 : g0;     // the sink layer must be in global memory
    *sink_value = last[(size_t)R * ks * ks];   // cell (r=R,0,0) of the last level (:730, :775)
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

// shape: [grid, cells_per_cta, tile_cells, stage_bytes] (0 = kernel default)
extern "C" int emu_dp_diploid(int32_t n_levels, const int32_t* level_off, const int64_t* adj_off,
                              const int32_t* adj_dst, const uint8_t* adj_w, const int64_t* col_off,
                              const int32_t* col_val, const uint8_t* colour_is_hom, int32_t n_colours, int32_t R,
                              int32_t* sink_value, int32_t* sink_s_het, int32_t* p1_edges, int32_t* n_p1,
                              int32_t* p2_edges, int32_t* n_p2, uint64_t* level_checksum, uint64_t* level_live,
                              int32_t force_pred32, const int32_t* shape, int64_t* mode_counts) {
    DipGraphView g;
    g.n_levels = n_levels; g.level_off = level_off; g.adj_off = adj_off; g.adj_dst = adj_dst; g.adj_w = adj_w;
    g.col_off = col_off; g.col_val = col_val; g.colour_is_hom = colour_is_hom; g.n_colours = n_colours; g.R = R;
    DipPlan p;
    if (!build_dip_plan(g, p)) return -1;
    SweepShape sh;
    sh.grid = 148; sh.cells_per_cta = 2048; sh.tile_cells = 16384; sh.stage_bytes = 16384;
    if (shape) {
        if (shape[0] > 0) sh.grid = shape[0];
        if (shape[1] > 0) sh.cells_per_cta = shape[1];
        if (shape[2] > 0) sh.tile_cells = shape[2];
        if (shape[3] > 0) sh.stage_bytes = shape[3];
    }
    plan_sweep(p, sh);
    if (mode_counts) { mode_counts[0] = p.n_fast; mode_counts[1] = p.n_staged; mode_counts[2] = p.n_global; }
    if (p.max_indeg <= 255 && !force_pred32)
        return run<uint16_t>(p, sh, sink_value, sink_s_het, p1_edges, n_p1, p2_edges, n_p2, level_checksum, level_live);
    return run<uint32_t>(p, sh, sink_value, sink_s_het, p1_edges, n_p1, p2_edges, n_p2, level_checksum, level_live);
}
