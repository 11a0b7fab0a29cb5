// tests/emu/dp_emu.cpp — TEST-ONLY CPU emulation of the CUDA diploid sweep.
// Runs the exact host planning (dp_prep.cpp: in-edge CSR, colour masks, task streams, packed records,
// pair-score matrix layout, barrier schedule) and the exact per-cell / traceback code (dp_cell.h) the
// kernels run, with the thread grid replaced by serial loops, the shared-memory tiles and task slots by
// host buffers and the bulk copies by memcpy, so that `-m "not gpu"` tests can check the gather
// formulation, the task/record packing, the slot layout (alignment, skew, sizes), the tile placement
// flags, the monotone barrier targets, the predecessor codes and the checkpointed traceback against the
// oracle without a GPU.  Never part of the product library.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>

#include "../../dipgenie_b200/csrc/cuda/dp_cell.h"
#include "../../dipgenie_b200/csrc/cuda/dp_prep.h"

using namespace dg;

namespace {

struct Emu {
    const DipPlan& p;
    const SweepShape& sh;
    std::vector<uint16_t> delta;
    std::vector<int32_t> g0, g1, s0, s1;
    std::vector<uint64_t> sum, live;
    bool bad_rc = false;
    int shift = 0;                     // layers hold value << shift (packed keys)
    std::vector<uint64_t> scratch;     // TK_LONG combine words
    int64_t n_lane_tasks = 0;
    std::vector<uint8_t> written;      // cells of the current destination level stored so far (lane form: exactly once)
    Emu(const DipPlan& pp, const SweepShape& s) : p(pp), sh(s) {}
};

// same thread -> item mapping as sweep_items() in dp_diploid.cu, threads run one after the other
template <class PredT, class OffT, int RC, bool MASKS>
void items(Emu& e, const TransitionT<OffT>& t, const TaskHdr& h, const int32_t* src, int32_t* dst, PredT* pl) {
    constexpr int SH = (sizeof(PredT) == 2) ? 8 : 16;
    const int R = e.p.R;
    const uint32_t CT = (uint32_t)e.sh.threads;
    const uint32_t k2 = h.k2, npairs = (uint32_t)(h.i1 - h.i0) * k2, nchunk = (uint32_t)(R + RC) / RC, groups = h.groups;
    auto load = [src](int64_t idx) { return src[idx]; };
    for (uint32_t tid = 0; tid < CT; ++tid) {
        uint32_t g = 0, pp = tid;
        if (groups > 1) {
            g = h.m_pairs ? div_magic(tid, h.m_pairs) : tid;
            pp = tid - g * npairs;
            if (g >= groups) continue;
        }
        for (uint32_t pair = pp; pair < npairs; pair += CT) {
            const uint32_t ir = h.m_k2 ? div_magic(pair, h.m_k2) : pair;
            const uint32_t j2 = pair - ir * k2, i2 = (uint32_t)h.i0 + ir;
            for (uint32_t chunk = g; chunk < nchunk; chunk += groups) {
                int32_t best[RC]; uint32_t code[RC];
                relax_pair<RC, MASKS, int64_t>(t, load, R, (int)chunk * RC, (int)i2, (int)j2, best, code);
                for (int rr = 0; rr < RC; ++rr) {
                    const int r2 = (int)chunk * RC + rr;
                    if (r2 > R) continue;
                    const uint64_t c = ((uint64_t)r2 * k2 + i2) * k2 + j2;
                    const bool lv = best[rr] >= 0;
                    e.written[c] = 1;
                    dst[c] = lv ? best[rr] : NEG_INF;
                    pl[c] = lv ? (PredT)(((code[rr] >> 16) << SH) | (code[rr] & 0xFFFFu)) : (PredT) ~(PredT)0;
                    if (lv) {
                        const int pi = (int)(t.in_edge[(int32_t)t.in_off[i2] + (int32_t)(code[rr] >> 16)] & 0xFFFFu);
                        const int pj = (int)(t.in_edge[(int32_t)t.in_off[j2] + (int32_t)(code[rr] & 0xFFFFu)] & 0xFFFFu);
                        ++e.live[h.level + 1];
                        e.sum[h.level + 1] += cell_fold(c, best[rr] >> e.shift, pi, pj);
                    }
                }
            }
        }
    }
}

template <class PredT, class OffT, bool MASKS>
void by_rc(Emu& e, const TransitionT<OffT>& t, const TaskHdr& h, const int32_t* src, int32_t* dst, PredT* pl) {
    if (h.rc != DIP_RC) { e.bad_rc = true; return; }
    items<PredT, OffT, DIP_RC, MASKS>(e, t, h, src, dst, pl);
}

template <class PredT, class OffT>
void by_dm(Emu& e, const TransitionT<OffT>& t, const TaskHdr& h, const int32_t* src, int32_t* dst, PredT* pl) {
    if (h.flags & TK_DELTA_MASKS) by_rc<PredT, OffT, true>(e, t, h, src, dst, pl);
    else by_rc<PredT, OffT, false>(e, t, h, src, dst, pl);
}

// the lane form (dp_diploid.cu: lane_task / lane_task_packed / long_finalize), warp by warp, the 32 lanes emulated
// in lock-step arrays
template <class PredT, int RC>
int lane_items(Emu& e, const uint8_t* slot, const TaskHdr& h, const int32_t* src, int32_t* dst, PredT* pl) {
    constexpr int SH = (sizeof(PredT) == 2) ? 8 : 16;
    const int R = e.p.R;
    const uint32_t NCW = (uint32_t)e.sh.threads / 32;
    const uint8_t* rec = slot + sizeof(TaskHdr);
    const uint16_t* in_off = reinterpret_cast<const uint16_t*>(rec);
    const uint32_t* in_edge = reinterpret_cast<const uint32_t*>(rec + rec_edge_offset(h.k2));
    const uint16_t* in_dst = reinterpret_cast<const uint16_t*>(rec + rec_dst_offset(h.k2, h.n_in));
    if (h.bstart_off != rec_bstart_offset(h.k2, h.n_in) || h.bstart_off + 2u * (h.nblk + 1u) > h.rec_bytes) return -50;
    const uint16_t* bstart = reinterpret_cast<const uint16_t*>(rec + h.bstart_off);
    const bool staged = (h.flags & TK_DELTA_STAGED) != 0;
    const bool packed = e.shift != 0;
    const bool is_long = (h.flags & TK_LONG) != 0;
    const uint16_t* long_j = nullptr;
    if (packed && e.shift != KEY_SHIFT) return -63;
    if (is_long) {
        if (!packed || h.n_long < 1 || h.n_long > (uint32_t)LANE_MAX_LONG || h.long_off != rec_long_offset(h.k2, h.n_in, h.nblk) ||
            h.long_off + 2u * h.n_long > h.rec_bytes) return -64;
        if ((uint64_t)(h.i1 - h.i0) * h.n_long * (uint64_t)(R + 1) > (uint64_t)LANE_SCRATCH_ENTRIES) return -65;
        long_j = reinterpret_cast<const uint16_t*>(rec + h.long_off);
        e.scratch.assign((size_t)(h.i1 - h.i0) * h.n_long * (size_t)(R + 1), 0);       // zeroed by all threads + barrier
    } else if (h.n_long) return -64;
    if (h.flags & TK_DELTA_MASKS) return -51;
    const bool has_delta = staged || (h.flags & TK_DELTA);
    const uint16_t* delta = staged ? reinterpret_cast<const uint16_t*>(rec + h.rec_bytes) + h.delta_skew - (size_t)in_off[h.i0] * h.n_in
                                   : ((h.flags & TK_DELTA) ? e.delta.data() + e.p.delta_off[h.level] : nullptr);     // in place
    const uint32_t k = h.k, k2 = h.k2, n_in = h.n_in, kk = k * k, kk2 = k2 * k2;
    const uint32_t nchunk = (uint32_t)(R + RC) / RC;
    if (h.n_witems != h.nrg * h.nblk * nchunk || h.nblk < 1 || h.rp < 1 || (h.nblk > 1 && h.rp != 1)) return -52;
    if ((uint64_t)h.nrg * h.rp < (uint64_t)(h.i1 - h.i0) || bstart[0] != 0 || bstart[h.nblk] != n_in) return -53;
    auto store_cell = [&](uint64_t c, bool lv, int32_t val, uint32_t cd, uint32_t a0, uint32_t s0) -> int {
        if (e.written[c]) return -59;
        e.written[c] = 1;
        dst[c] = lv ? val : NEG_INF;
        pl[c] = lv ? (PredT)(((cd >> 16) << SH) | (cd & 0xFFFFu)) : (PredT) ~(PredT)0;
        if (lv) {
            const int pi = (int)(in_edge[a0 + (cd >> 16)] & 0xFFFFu);
            const int pj = (int)(in_edge[s0 + (cd & 0xFFFFu)] & 0xFFFFu);
            ++e.live[h.level + 1];
            e.sum[h.level + 1] += cell_fold(c, val >> e.shift, pi, pj);
        }
        return 0;
    };
    for (uint32_t warp = 0; warp < NCW; ++warp)
        for (uint32_t wi = warp; wi < h.n_witems; wi += NCW) {
            const uint32_t q = h.m_nblk ? div_magic(wi, h.m_nblk) : wi, b = wi - q * h.nblk;
            const uint32_t chunk = h.m_nrg ? div_magic(q, h.m_nrg) : q, rg = q - chunk * h.nrg;
            if (b >= h.nblk || rg >= h.nrg || chunk >= nchunk) return -54;
            const uint32_t bs = bstart[b], be = bstart[b + 1];
            if (be - bs > 32 || be <= bs) return -55;
            const int r0 = (int)chunk * RC;
            int32_t best[32][RC]; uint32_t code[32][RC];
            bool valid[32]; uint32_t pos[32], seg[32], row[32], j2[32], a0[32], a1[32], s0[32], el[32], e2c[32], jv[32]; int wv[32];
            bool any_long = false, any_slice = false;
            for (uint32_t lane = 0; lane < 32; ++lane) {
                uint32_t rs = 0; el[lane] = lane;
                if (h.nblk == 1) { rs = h.m_nin ? div_magic(lane, h.m_nin) : lane; el[lane] = lane - rs * n_in; }
                row[lane] = h.i0 + rg * h.rp + rs;
                const uint32_t e2 = bs + el[lane];
                valid[lane] = rs < h.rp && row[lane] < h.i1 && e2 < be;
                const uint32_t rowc = valid[lane] ? row[lane] : h.i0;
                e2c[lane] = valid[lane] ? e2 : bs;
                a0[lane] = in_off[rowc];
                a1[lane] = valid[lane] ? in_off[rowc + 1] : a0[lane];
                const uint32_t y = in_edge[e2c[lane]];
                j2[lane] = in_dst[e2c[lane]];
                s0[lane] = in_off[j2[lane]];
                pos[lane] = e2c[lane] - s0[lane]; seg[lane] = in_off[j2[lane] + 1] - s0[lane];
                if (valid[lane] && pos[lane] >= seg[lane]) return -56;
                if (valid[lane] && seg[lane] <= 32 && (s0[lane] < bs || s0[lane] + seg[lane] > be)) return -56;   // a block cut a short group
                if (valid[lane] && seg[lane] > 32 && (!is_long || s0[lane] > bs || s0[lane] + seg[lane] < be)) return -56;  // a slice holds one long group only
                jv[lane] = y & 0xFFFFu; wv[lane] = (int)(y >> 16);
                if (valid[lane] && (seg[lane] > 32 || a1[lane] - a0[lane] > 32)) any_long = true;
                if (valid[lane] && seg[lane] > 32) any_slice = true;
            }
            const bool unp = is_long;                        // TK_LONG tasks run the unpacked lane task
            if (any_long && !is_long) return -66;
            const bool use_packed = packed && !unp;
            for (uint32_t lane = 0; lane < 32; ++lane) {
                for (int rr = 0; rr < RC; ++rr) { best[lane][rr] = -1; code[lane][rr] = 0xFFFFFFFFu; }
                for (uint32_t e1 = a0[lane]; e1 < a1[lane]; ++e1) {
                    const uint32_t x = in_edge[e1];
                    const uint32_t base = (x & 0xFFFFu) * k + jv[lane];
                    const int w = (int)(x >> 16) + wv[lane];
                    const int d = has_delta ? (int)delta[(size_t)e1 * n_in + e2c[lane]] : 0;
                    const uint32_t cd = ((e1 - a0[lane]) << 16) | pos[lane];
                    if (use_packed && ((e1 - a0[lane]) > 31 || pos[lane] > 31)) return -62;
                    const int32_t dp = (int32_t)(((uint32_t)d << KEY_SHIFT) | ((31u - (e1 - a0[lane])) << KEY_ORD_BITS) | (31u - pos[lane]));
                    for (int rr = 0; rr < RC; ++rr) {
                        int r = r0 + rr - w;
                        const bool ok = r >= 0 && r0 + rr <= R;
                        r = r < 0 ? 0 : (r > R ? R : r);
                        const int32_t sv = src[(size_t)r * kk + base];
                        if (use_packed) {                       // one word per layer: key = value << 10 | ordinals
                            const int32_t key = ok ? sv + dp : -1;
                            if (key > best[lane][rr]) best[lane][rr] = key;
                        } else {
                            const int32_t c = sv + (d << e.shift);
                            if (ok && c > best[lane][rr]) { best[lane][rr] = c; code[lane][rr] = cd; }
                        }
                    }
                }
            }
            const uint32_t nrd = (unp && any_slice) ? 5u : h.rounds;
            for (uint32_t rd = 0, off = 1; rd < nrd; ++rd, off <<= 1) {
                int32_t nb[32][RC]; uint32_t nc[32][RC];
                for (uint32_t lane = 0; lane < 32; ++lane)
                    for (int rr = 0; rr < RC; ++rr) {
                        const uint32_t o = lane + off < 32 ? lane + off : lane;      // __shfl_down_sync semantics
                        const int32_t ob = best[o][rr]; const uint32_t oc = code[o][rr];
                        nb[lane][rr] = best[lane][rr]; nc[lane][rr] = code[lane][rr];
                        const bool partner = (unp && any_slice) ? (valid[lane] && el[lane] + off < be - bs) : (pos[lane] + off < seg[lane]);
                        const bool take = use_packed ? (ob > best[lane][rr]) : (ob > best[lane][rr] || (ob == best[lane][rr] && oc < code[lane][rr]));
                        if (partner && take) {
                            if (lane + off >= 32) return -57;
                            nb[lane][rr] = ob; nc[lane][rr] = oc;
                        }
                    }
                memcpy(best, nb, sizeof best); memcpy(code, nc, sizeof code);
            }
            for (uint32_t lane = 0; lane < 32; ++lane) {
                if (!valid[lane]) continue;
                if (unp && any_slice) {
                    if (el[lane] != 0) continue;                 // the slice's leader: into the scratch words
                    uint32_t g = 0;
                    while (g + 1 < h.n_long && long_j[g] != j2[lane]) ++g;
                    if (long_j[g] != j2[lane]) return -67;
                    for (int rr = 0; rr < RC; ++rr) {
                        const int r2 = r0 + rr;
                        if (r2 > R || best[lane][rr] < 0) continue;
                        uint64_t& w = e.scratch[((size_t)(row[lane] - h.i0) * h.n_long + g) * (size_t)(R + 1) + r2];
                        const uint64_t K = ((uint64_t)((uint32_t)best[lane][rr] + 1u) << 32) | (uint64_t)(0xFFFFFFFFu - code[lane][rr]);
                        if (K > w) w = K;
                    }
                    continue;
                }
                if (pos[lane] != 0) continue;
                if (!unp && (1u << h.rounds) < seg[lane]) return -58;
                for (int rr = 0; rr < RC; ++rr) {
                    const int r2 = r0 + rr;
                    if (r2 > R) continue;
                    const uint64_t c = (uint64_t)r2 * kk2 + (uint64_t)row[lane] * k2 + j2[lane];
                    const bool lv = best[lane][rr] >= 0;
                    int32_t val = best[lane][rr];
                    uint32_t cd = code[lane][rr];
                    if (use_packed && lv) {
                        const uint32_t inv = ~(uint32_t)val & 1023u;
                        cd = ((inv >> KEY_ORD_BITS) << 16) | (inv & 31u);
                        val = (int32_t)((uint32_t)val & ~1023u);
                    }
                    if (int rcs = store_cell(c, lv, val, cd, a0[lane], s0[lane])) return rcs;
                }
            }
        }
    if (is_long) {                                               // long_finalize, after a block barrier
        const uint32_t RP1 = (uint32_t)R + 1u, G = h.n_long;
        for (uint32_t idx = 0; idx < (h.i1 - h.i0) * G * RP1; ++idx) {
            const uint64_t K = e.scratch[idx];
            const uint32_t q = idx / RP1, r2 = idx - q * RP1, rowrel = q / G, g = q - rowrel * G;
            const uint32_t row = h.i0 + rowrel, j2 = long_j[g];
            const bool lv = K != 0;
            const int32_t val = lv ? (int32_t)((uint32_t)(K >> 32) - 1u) : NEG_INF;
            const uint32_t cd = lv ? 0xFFFFFFFFu - (uint32_t)K : 0xFFFFFFFFu;
            const uint64_t c = (uint64_t)r2 * kk2 + (uint64_t)row * k2 + j2;
            if (int rcs = store_cell(c, lv, val, cd, in_off[row], in_off[j2])) return rcs;
        }
    }
    return 0;
}

// One task, as the producer warp (three bulk copies into the slot) and the compute warps execute it.  `c` = the local
// CTA running it; `h` receives the header as the compute warps see it.
template <class PredT>
int exec_task(Emu& e, const DipPlan& p, const SweepShape& sh, const TaskHdr& gh, int l, int c, std::vector<uint8_t>& slot,
              std::vector<uint8_t>& row_done, PredT* pl, TaskHdr& h) {
    const int R = p.R;
        // --- what the producer warp does: three bulk copies into the slot ---
        if ((size_t)sizeof(TaskHdr) + gh.rec_bytes + gh.delta_bytes > (size_t)sh.slot_bytes) return -10;
        if (gh.rec_bytes % 16 || gh.delta_bytes % 16) return -10;
        std::fill(slot.begin(), slot.end(), (uint8_t)0xEE);
        memcpy(slot.data(), &gh, sizeof(TaskHdr));
        if (gh.rec_bytes) {
            if ((size_t)gh.rec_off16 * 16 + gh.rec_bytes > p.records.size()) return -10;
            memcpy(slot.data() + sizeof(TaskHdr), p.records.data() + (size_t)gh.rec_off16 * 16, gh.rec_bytes);
        }
        if (gh.delta_bytes) {
            if ((size_t)gh.delta_off16 * 8 + gh.delta_bytes / 2 > e.delta.size() + 8) return -10;
            const size_t avail = (e.delta.size() - (size_t)gh.delta_off16 * 8) * 2;
            memcpy(slot.data() + sizeof(TaskHdr) + gh.rec_bytes, e.delta.data() + (size_t)gh.delta_off16 * 8,
                   std::min<size_t>(gh.delta_bytes, avail));
        }
        // --- what the compute warps do ---
        memcpy(&h, slot.data(), sizeof h);
        const bool ssm = h.flags & TK_SRC_SMEM, dsm = h.flags & TK_DST_SMEM;
        if ((ssm || dsm) && (!p.narrow[l] || c != 0)) return -12;
        const size_t lp = (size_t)((R + sh.lane_rc) / sh.lane_rc) * sh.lane_rc;       // whole lane-form chunks
        if ((ssm && lp * h.k * h.k > (size_t)sh.tile_cells) || (dsm && lp * h.k2 * h.k2 > (size_t)sh.tile_cells)) return -13;
        if (h.pred_off2 != p.pred_off[l + 1] || h.i0 >= h.i1 || h.i1 > h.k2) return -14;
        if ((uint64_t)(h.i1 - h.i0) * h.k2 * h.k2 >= (1ull << 32)) return -14;    // div_magic exactness
        const int32_t* src = ssm ? ((l & 1) ? e.s1.data() : e.s0.data()) : ((l & 1) ? e.g1.data() : e.g0.data());
        int32_t* dst = dsm ? ((l & 1) ? e.s0.data() : e.s1.data()) : ((l & 1) ? e.g0.data() : e.g1.data());
        for (int x = h.i0; x < h.i1; ++x) { if (row_done[x]) return -16; row_done[x] = 1; }
        if (h.flags & TK_LANES) {
            if (ssm != dsm) return -60;
            if (h.rc != ((h.flags & TK_LONG) ? LANE_RC_SMALL : e.sh.lane_rc)) return -61;
            const uint8_t* view = slot.data();
            TaskHdr hv = h;
            std::vector<uint8_t> inplace;
            if (h.flags & TK_REC_GLOBAL) {
                // record too big for a slot: the lanes read it in place (a.records + rec_off16 * 16); nothing of it is staged
                if (ssm || h.rec_bytes != 0 || (h.flags & TK_DELTA_STAGED) || p.rec_off[l] < 0 || (int64_t)h.rec_off16 * 16 != p.rec_off[l]) return -66;
                const size_t len = rec_bytes_for(h.k2, h.n_in, h.nblk, (int)h.n_long);
                if ((size_t)p.rec_off[l] + len > p.records.size()) return -66;
                if (sizeof(TaskHdr) + len <= (size_t)sh.slot_bytes) return -66;         // would have been staged
                inplace.resize(sizeof(TaskHdr) + len);
                memcpy(inplace.data() + sizeof(TaskHdr), p.records.data() + p.rec_off[l], len);
                hv.rec_bytes = (uint32_t)len;
                view = inplace.data();
            }
            const int lrc = hv.rc == LANE_RC_BIG ? lane_items<PredT, LANE_RC_BIG>(e, view, hv, src, dst, pl)
                                                 : lane_items<PredT, LANE_RC_SMALL>(e, view, hv, src, dst, pl);
            if (lrc) return lrc;
            ++e.n_lane_tasks;
        } else if (!(h.flags & TK_REC_GLOBAL)) {
            TransitionT<uint16_t> t;
            t.k = h.k; t.k2 = h.k2;
            t.in_off = reinterpret_cast<const uint16_t*>(slot.data() + sizeof(TaskHdr));
            t.in_edge = reinterpret_cast<const uint32_t*>(slot.data() + sizeof(TaskHdr) + rec_edge_offset(h.k2));
            t.delta = nullptr; t.dstride = h.n_in; t.e1_base = 0; t.e2_base = 0; t.dshift = e.shift; t.W = 0; t.msrc = t.mdst = nullptr;
            if (h.flags & TK_DELTA_STAGED) {
                t.delta = reinterpret_cast<const uint16_t*>(slot.data() + sizeof(TaskHdr) + h.rec_bytes) + h.delta_skew;
                t.e1_base = (int32_t)t.in_off[h.i0];
            } else if (h.flags & TK_DELTA) {
                t.delta = e.delta.data() + p.delta_off[l];
            } else if (h.flags & TK_DELTA_MASKS) {
                t.W = p.lvlW[l]; t.msrc = p.masks.data() + p.msrc_off[l]; t.mdst = p.masks.data() + p.mdst_off[l];
            }
            by_dm<PredT, uint16_t>(e, t, h, src, dst, pl);
        } else {
            const int32_t mid = p.level_off[l + 1], ebase = p.in_off[mid];
            TransitionT<int32_t> t;
            t.k = h.k; t.k2 = h.k2;
            t.in_off = p.in_off.data() + mid; t.in_edge = p.in_edge.data();
            t.delta = nullptr; t.dstride = p.in_off[p.level_off[l + 2]] - ebase; t.e1_base = ebase; t.e2_base = ebase; t.dshift = e.shift;
            t.W = 0; t.msrc = t.mdst = nullptr;
            if (h.flags & TK_DELTA) t.delta = e.delta.data() + p.delta_off[l];
            else if (h.flags & TK_DELTA_MASKS) {
                t.W = p.lvlW[l]; t.msrc = p.masks.data() + p.msrc_off[l]; t.mdst = p.masks.data() + p.mdst_off[l];
            }
            by_dm<PredT, int32_t>(e, t, h, src, dst, pl);
        }    return 0;
}

// K7: checkpointed traceback (dip_anc_kernel / dip_hop_kernel / dip_seg_kernel / dip_merge_kernel)
template <class PredT>
int traceback(const DipPlan& p, const Emu& e, const std::vector<PredT>& pred, int trace_T, int32_t* sink_value, int32_t* sink_s_het,
              int32_t* p1, int32_t* n1, int32_t* p2, int32_t* n2) {
    const int R = p.R, L = p.L;
    const size_t ks = (size_t)(p.level_off[L] - p.level_off[L - 1]);
    const std::vector<int32_t>& last = ((L - 1) & 1) ? e.g1 : e.g0;     // the sink layer must be in global memory
    *sink_value = last[(size_t)R * ks * ks];   // cell (r=R,0,0) of the last level (:730, :775)
    *sink_value = *sink_value < 0 ? NEG_INF : (*sink_value >> e.shift);
    TraceView v;
    v.L = L; v.R = R; v.level_off = p.level_off.data(); v.in_off = p.in_off.data(); v.in_edge = p.in_edge.data();
    v.lvlW = p.lvlW.data(); v.msrc_off = p.msrc_off.data(); v.mdst_off = p.mdst_off.data();
    v.masks = p.masks.data(); v.pred_off = p.pred_off.data();
    const std::vector<int32_t> cp = choose_checkpoints(p.level_off, trace_T);
    if (cp.front() != L - 1 || cp.back() != 0) return -43;
    for (size_t x = 1; x < cp.size(); ++x) if (cp[x] >= cp[x - 1]) return -43;
    const int M = (int)cp.size() - 1;
    const int cap = R + 2;
    *n1 = 0; *n2 = 0; *sink_s_het = 0;
    if (*sink_value == NEG_INF) return 0;
    auto state_of = [&](int64_t cell, int l, TraceState& s) {
        const int32_t k = p.level_off[l + 1] - p.level_off[l];
        s.r = (int32_t)(cell / ((int64_t)k * k));
        const int32_t rem = (int32_t)(cell - (int64_t)s.r * k * k);
        s.i2 = rem / k; s.j2 = rem - s.i2 * k;
    };
    std::vector<int32_t> a, b;
    int64_t cur_cell = (int64_t)R * ks * ks;
    int sh_total = 0;
    for (int m = 0; m < M; ++m) {
        // hop: the ancestor of cur_cell at cp[m+1] (what dip_anc_kernel stores for every cell of cp[m])
        TraceState s;
        state_of(cur_cell, cp[m], s);
        TraceState w = s;
        for (int l = cp[m] - 1; l >= cp[m + 1]; --l) {
            int wu, wv, i2, j2;
            if (!trace_step<PredT>(v, pred.data(), l, w, wu, wv, i2, j2)) return -40;
        }
        // segment walk
        std::vector<int32_t> sa(2 * cap), sb(2 * cap);
        int32_t sn1 = 0, sn2 = 0, ssh = 0;
        const int rc = trace_segment<PredT>(v, pred.data(), cp[m], cp[m + 1], s, sa.data(), &sn1, sb.data(), &sn2, cap, &ssh);
        if (rc) return rc == -2 ? -2 : -41;
        if (s.r != w.r || s.i2 != w.i2 || s.j2 != w.j2) return -42;
        a.insert(a.end(), sa.begin(), sa.begin() + 2 * sn1);
        b.insert(b.end(), sb.begin(), sb.begin() + 2 * sn2);
        sh_total += ssh;
        const int32_t k = p.level_off[cp[m + 1] + 1] - p.level_off[cp[m + 1]];
        cur_cell = ((int64_t)w.r * k + w.i2) * k + w.j2;
    }
    if ((int)a.size() > 2 * cap || (int)b.size() > 2 * cap) return -2;
    *n1 = (int32_t)a.size() / 2; *n2 = (int32_t)b.size() / 2; *sink_s_het = sh_total;
    for (int x = 0; x < *n1; ++x) { p1[2 * x] = a[2 * (*n1 - 1 - x)]; p1[2 * x + 1] = a[2 * (*n1 - 1 - x) + 1]; }
    for (int x = 0; x < *n2; ++x) { p2[2 * x] = b[2 * (*n2 - 1 - x)]; p2[2 * x + 1] = b[2 * (*n2 - 1 - x) + 1]; }
    return 0;
}

template <class PredT>
int run(const DipPlan& p, const SweepShape& sh, int trace_T, bool no_pack, int64_t* n_lane_tasks, int32_t* sink_value, int32_t* sink_s_het, int32_t* p1,
        int32_t* n1, int32_t* p2, int32_t* n2, uint64_t* level_checksum, uint64_t* level_live) {
    const int R = p.R, L = p.L;
    Emu e(p, sh);
    e.shift = (p.value_bound < KEY_VALUE_LIMIT && !no_pack) ? KEY_SHIFT : 0;    // same rule as dip_plan_host
    std::vector<PredT> pred((size_t)p.pred_off[L]);
    const size_t widest = (size_t)(R + sh.lane_rc) * p.kmax * p.kmax;
    // poison values make a wrong tile-placement flag visible
    e.g0.assign(widest, 0x5A5A5A5A); e.g1.assign(widest, 0x5A5A5A5A);
    e.s0.assign(sh.tile_cells, 0x3C3C3C3C); e.s1.assign(sh.tile_cells, 0x3C3C3C3C);
    for (int r = 0; r <= R; ++r) { e.g0[r] = 0; e.s0[r] = 0; }
    e.sum.assign(L, FOLD_BASIS); e.live.assign(L, 0);

    // K4: pair-score matrices (dip_delta_kernel)
    e.delta.assign((size_t)p.delta_elems, 0xDEAD);
    for (int l : p.delta_list) {
        const int32_t mid = p.level_off[l + 1], e0 = p.in_off[mid];
        const uint32_t n_in = (uint32_t)(p.in_off[p.level_off[l + 2]] - e0);
        if (p.delta_off[l] % 8 != 0) return -20;
        uint16_t* D = e.delta.data() + p.delta_off[l];
        for (uint32_t x = 0; x < n_in * n_in; ++x) {
            const uint32_t e1 = x / n_in, e2 = x - e1 * n_in;
            D[x] = (uint16_t)mask_delta(p.lvlW[l], p.masks.data() + p.msrc_off[l], p.masks.data() + p.mdst_off[l],
                                        (int)(p.in_edge[e0 + e1] & 0xFFFFu), (int)(p.in_edge[e0 + e2] & 0xFFFFu),
                                        (int)p.in_dst[e0 + e1], (int)p.in_dst[e0 + e2]);
        }
    }

    // K5: task streams, executed level by level, CTA by CTA (any order the barriers allow is equivalent)
    const int G = p.grid;
    std::vector<int64_t> cur(p.task_begin.begin(), p.task_begin.end() - 1);
    std::vector<uint8_t> slot((size_t)sh.slot_bytes + 16);
    uint32_t counter = 0;
    for (int l = 0; l + 1 < L; ++l) {
        PredT* pl = pred.data() + p.pred_off[l + 1];
        std::vector<uint8_t> row_done((size_t)(p.level_off[l + 2] - p.level_off[l + 1]), 0);
        e.written.assign((size_t)(R + 1) * row_done.size() * row_done.size(), 0);
        uint32_t arrivals = 0;
        int participants = 0;
        for (int c = 0; c < G; ++c) {
            bool first = true, closed = false;
            while (cur[c] < p.task_begin[(size_t)c + 1] && p.tasks[(size_t)cur[c]].level == l) {
                const TaskHdr& gh = p.tasks[(size_t)cur[c]++];
                if (closed) return -30;                         // a task after the CTA's TK_BAR task of this level
                TaskHdr h;
                if (int rc = exec_task<PredT>(e, p, sh, gh, l, c, slot, row_done, pl, h)) return rc;
                const bool needs_wait = l > 0 && p.bar_edge[l - 1];
                if (first) {
                    if (needs_wait != ((h.flags & TK_WAIT) != 0)) return -11;
                    if (needs_wait && (h.wait_target != p.bar_target[l - 1] || counter < h.wait_target)) return -11;   // would race / dead-lock
                } else if (h.flags & TK_WAIT) return -11;
                first = false;
                if (c >= p.P[l]) return -12;
                if (h.flags & TK_BAR) closed = true;
                if (h.flags & TK_ARRIVE) { if (!(h.flags & TK_BAR)) return -17; ++arrivals; }
            }
            if (!first) { ++participants; if (!closed) return -18; }
        }
        for (uint8_t d : row_done) if (!d) return -19;
        for (uint8_t d : e.written) if (!d) return -25;             // every cell of the level was stored              // every destination row belongs to exactly one task
        if (participants != p.P[l]) return -21;
        if (p.bar_edge[l] ? (arrivals != (uint32_t)p.P[l]) : (arrivals != 0)) return -22;
        counter += arrivals;
        if (counter != p.bar_target[l]) return -15;
        if (level_checksum) { level_checksum[l + 1] = e.sum[l + 1]; level_live[l + 1] = e.live[l + 1]; }
    }
    for (int c = 0; c < G; ++c) if (cur[c] != p.task_begin[(size_t)c + 1]) return -23;
    if (e.bad_rc) return -24;
    if (n_lane_tasks) *n_lane_tasks = e.n_lane_tasks;

    return traceback<PredT>(p, e, pred, trace_T, sink_value, sink_s_het, p1, n1, p2, n2);
}

// The row-sharded sweep (dg_dip_create_sharded): N ranks, each with its own plan (same graph, its own rank), layers,
// predecessor codes and counter; levels are executed rank by rank, then the TK_PUSH row copies and the broadcast
// arrivals are applied.  Checks: every destination row of a wide transition belongs to exactly one task over all ranks,
// every rank's counter meets the (identical) barrier targets, all ranks end with identical layers and codes.
template <class PredT>
int run_sharded(const DipGraphView& g, const SweepShape& sh0, int N, int trace_T, bool no_pack, int32_t* sink_value, int32_t* sink_s_het,
                int32_t* p1, int32_t* n1, int32_t* p2, int32_t* n2, int64_t* counts) {
    std::vector<std::unique_ptr<DipPlan>> plans;
    std::vector<SweepShape> shapes((size_t)N, sh0);
    std::vector<std::unique_ptr<Emu>> emus;
    std::vector<std::vector<PredT>> preds((size_t)N);
    for (int r = 0; r < N; ++r) {
        plans.emplace_back(new DipPlan());
        if (!build_dip_plan(g, *plans[r])) return -1;
        shapes[r].replicas = N; shapes[r].rank = r;
        plan_tasks(*plans[r], shapes[r]);
    }
    const DipPlan& p0 = *plans[0];
    const int R = p0.R, L = p0.L, G = p0.grid;
    const size_t widest = (size_t)(R + sh0.lane_rc) * p0.kmax * p0.kmax;
    for (int r = 0; r < N; ++r) {
        const DipPlan& p = *plans[r];
        if (p.bar_target != p0.bar_target || p.bar_edge != p0.bar_edge || p.P != p0.P || p.narrow != p0.narrow) return -70;
        emus.emplace_back(new Emu(p, shapes[r]));
        Emu& e = *emus[r];
        e.shift = (p.value_bound < KEY_VALUE_LIMIT && !no_pack) ? KEY_SHIFT : 0;
        preds[r].assign((size_t)p.pred_off[L], (PredT)0x7B7B);
        e.g0.assign(widest, 0x5A5A5A5A); e.g1.assign(widest, 0x5A5A5A5A);
        e.s0.assign(sh0.tile_cells, 0x3C3C3C3C); e.s1.assign(sh0.tile_cells, 0x3C3C3C3C);
        for (int x = 0; x <= R; ++x) { e.g0[x] = 0; e.s0[x] = 0; }
        e.sum.assign(L, FOLD_BASIS); e.live.assign(L, 0);
        e.delta.assign((size_t)p.delta_elems, 0xDEAD);
        for (int l : p.delta_list) {
            const int32_t mid = p.level_off[l + 1], e0 = p.in_off[mid];
            const uint32_t n_in = (uint32_t)(p.in_off[p.level_off[l + 2]] - e0);
            uint16_t* D = e.delta.data() + p.delta_off[l];
            for (uint32_t x = 0; x < n_in * n_in; ++x) {
                const uint32_t e1 = x / n_in, e2 = x - e1 * n_in;
                D[x] = (uint16_t)mask_delta(p.lvlW[l], p.masks.data() + p.msrc_off[l], p.masks.data() + p.mdst_off[l],
                                            (int)(p.in_edge[e0 + e1] & 0xFFFFu), (int)(p.in_edge[e0 + e2] & 0xFFFFu),
                                            (int)p.in_dst[e0 + e1], (int)p.in_dst[e0 + e2]);
            }
        }
    }
    std::vector<std::vector<int64_t>> cur((size_t)N);
    for (int r = 0; r < N; ++r) cur[r].assign(plans[r]->task_begin.begin(), plans[r]->task_begin.end() - 1);
    std::vector<uint8_t> slot((size_t)sh0.slot_bytes + 16);
    std::vector<uint32_t> counter((size_t)N, 0), local_count((size_t)N, 0);
    int64_t n_push = 0, n_wide_tasks = 0;
    for (int l = 0; l + 1 < L; ++l) {
        const size_t k2 = (size_t)(p0.level_off[l + 2] - p0.level_off[l + 1]);
        const bool narrow = p0.narrow[l] != 0;
        std::vector<uint8_t> rows_shared(k2, 0);
        uint32_t arrivals = 0;
        int participants = 0;
        struct Push { int rank; int i0, i1; };
        std::vector<Push> pushes;
        for (int r = 0; r < N; ++r) {
            const DipPlan& p = *plans[r];
            Emu& e = *emus[r];
            PredT* pl = preds[r].data() + p.pred_off[l + 1];
            std::vector<uint8_t> rows_own(k2, 0);
            std::vector<uint8_t>& row_done = narrow ? rows_own : rows_shared;
            e.written.assign((size_t)(R + 1) * k2 * k2, 0);
            for (int c = 0; c < G; ++c) {
                bool first = true, closed = false;
                while (cur[r][c] < p.task_begin[(size_t)c + 1] && p.tasks[(size_t)cur[r][c]].level == l) {
                    const TaskHdr& gh = p.tasks[(size_t)cur[r][c]++];
                    if (closed) return -30;
                    TaskHdr h;
                    if (int rc = exec_task<PredT>(e, p, shapes[r], gh, l, c, slot, row_done, pl, h)) return rc;
                    const bool needs_wait = l > 0 && p.bar_edge[l - 1];
                    if (first) {
                        if (needs_wait != ((h.flags & TK_WAIT) != 0)) return -11;
                        if (needs_wait && (h.wait_target != p.bar_target[l - 1] || counter[r] < h.wait_target)) return -11;
                    } else if (h.flags & TK_WAIT) return -11;
                    first = false;
                    if (narrow ? c != 0 : (c * N + r >= p.P[l])) return -12;          // local CTA c of rank r = global CTA c * N + r
                    if (!narrow) ++n_wide_tasks;
                    if (h.flags & TK_BAR) closed = true;
                    if (h.flags & TK_ARRIVE) {
                        if (!(h.flags & TK_BAR)) return -17;
                        // counted on the rank's local word; the CTA that completes the rank's share forwards it
                        if (++local_count[r] > h.arrive_local_target) return -77;
                        if (local_count[r] == h.arrive_local_target) arrivals += h.arrive_n;
                    }
                    if (((h.flags & TK_PUSH) != 0) != (!narrow && (h.flags & TK_BAR))) return -71;
                    if (h.flags & TK_PUSH) {
                        if (h.flags & TK_DST_SMEM) return -72;
                        if (h.push_i0 > h.i0 || h.push_i1 != h.i1) return -72;       // the CTA's whole range ends with this task
                        pushes.push_back({r, (int)h.push_i0, (int)h.push_i1});
                        ++n_push;
                    }
                }
                if (!first) { ++participants; if (!closed) return -18; }
            }
            if (narrow) {
                for (uint8_t d : rows_own) if (!d) return -19;
                for (uint8_t d : e.written) if (!d) return -25;
            }
        }
        if (!narrow) for (uint8_t d : rows_shared) if (!d) return -19;
        if (participants != (narrow ? N : p0.P[l])) return -21;
        if (p0.bar_edge[l] ? (arrivals != (uint32_t)(narrow ? N : p0.P[l])) : (arrivals != 0)) return -22;
        // the pushes: rows of layer l+1 and of its codes, from the owner's global tile to every peer's
        std::vector<uint8_t> pushed(k2, 0);
        for (const Push& q : pushes) {
            const std::vector<int32_t>& src = (l & 1) ? emus[q.rank]->g0 : emus[q.rank]->g1;
            const PredT* ps = preds[q.rank].data() + p0.pred_off[l + 1];
            for (int x = q.i0; x < q.i1; ++x) { if (pushed[x]) return -73; pushed[x] = 1; }
            for (int t = 0; t < N; ++t) {
                if (t == q.rank) continue;
                std::vector<int32_t>& dst = (l & 1) ? emus[t]->g0 : emus[t]->g1;
                PredT* pd = preds[t].data() + p0.pred_off[l + 1];
                for (int rr = 0; rr <= R; ++rr)
                    for (size_t x = (size_t)q.i0 * k2; x < (size_t)q.i1 * k2; ++x) {
                        dst[(size_t)rr * k2 * k2 + x] = src[(size_t)rr * k2 * k2 + x];
                        pd[(size_t)rr * k2 * k2 + x] = ps[(size_t)rr * k2 * k2 + x];
                    }
            }
        }
        if (!narrow) for (uint8_t d : pushed) if (!d) return -74;
        for (int r = 0; r < N; ++r) {                       // arrivals are broadcast
            counter[r] += arrivals;
            if (counter[r] != p0.bar_target[l]) return -15;
        }
    }
    for (int r = 0; r < N; ++r) {
        for (int c = 0; c < G; ++c) if (cur[r][c] != plans[r]->task_begin[(size_t)c + 1]) return -23;
        if (emus[r]->bad_rc) return -24;
        if (preds[r] != preds[0]) return -75;               // every rank holds the complete codes
        const std::vector<int32_t>& last = ((L - 1) & 1) ? emus[r]->g1 : emus[r]->g0;
        const std::vector<int32_t>& last0 = ((L - 1) & 1) ? emus[0]->g1 : emus[0]->g0;
        const size_t ks = (size_t)(p0.level_off[L] - p0.level_off[L - 1]);
        for (size_t x = 0; x < (size_t)(R + 1) * ks * ks; ++x) if (last[x] != last0[x]) return -76;
    }
    if (counts) { counts[0] = p0.n_narrow; counts[1] = p0.n_wide; counts[2] = n_wide_tasks; counts[3] = n_push; }
    const int last_rank = N - 1;                            // any rank can trace
    return traceback<PredT>(*plans[last_rank], *emus[last_rank], preds[last_rank], trace_T, sink_value, sink_s_het, p1, n1, p2, n2);
}

}  // namespace

// shape: [grid, threads, tile_cells, slot_bytes, delta_max_in, trace_T, no_pack, no_long] (0 = kernel default)
// counts: [narrow transitions, wide transitions, tasks, tasks with in-place records, tasks with on-the-fly masks, matrices,
//          tasks run in lane form, TK_LONG tasks, lane-form tasks with in-place records]
extern "C" int emu_dp_diploid(int32_t n_levels, const int32_t* level_off, const int64_t* adj_off,
                              const int32_t* adj_dst, const uint8_t* adj_w, const int64_t* col_off,
                              const int32_t* col_val, const uint8_t* colour_is_hom, int32_t n_colours, int32_t R,
                              int32_t* sink_value, int32_t* sink_s_het, int32_t* p1_edges, int32_t* n_p1,
                              int32_t* p2_edges, int32_t* n_p2, uint64_t* level_checksum, uint64_t* level_live,
                              int32_t force_pred32, const int32_t* shape, int64_t* counts) {
    DipGraphView g;
    g.n_levels = n_levels; g.level_off = level_off; g.adj_off = adj_off; g.adj_dst = adj_dst; g.adj_w = adj_w;
    g.col_off = col_off; g.col_val = col_val; g.colour_is_hom = colour_is_hom; g.n_colours = n_colours; g.R = R;
    DipPlan p;
    if (!build_dip_plan(g, p)) return -1;
    SweepShape sh;
    sh.grid = 32; sh.threads = 480; sh.tile_cells = 16384; sh.slot_bytes = 4096;
    int trace_T = 128;
    bool no_pack = false;
    if (shape) {
        if (shape[0] > 0) sh.grid = shape[0];
        if (shape[1] > 0) sh.threads = shape[1];
        if (shape[2] > 0) sh.tile_cells = shape[2];
        if (shape[3] > 0) sh.slot_bytes = shape[3];
        if (shape[4] > 0) sh.delta_max_in = shape[4];
        if (shape[4] < 0) sh.delta_max_in = 0;
        if (shape[5] > 0) trace_T = shape[5];
        no_pack = shape[6] != 0;
    }
    sh.lane_rc = (p.value_bound < KEY_VALUE_LIMIT && !no_pack && R + 1 >= LANE_RC_BIG) ? LANE_RC_BIG : LANE_RC_SMALL;   // as dip_plan_host
    sh.allow_long = (p.value_bound < KEY_VALUE_LIMIT && !no_pack) && !(shape && shape[7] != 0);
    plan_tasks(p, sh);
    if (counts) {
        counts[0] = p.n_narrow; counts[1] = p.n_wide; counts[2] = (int64_t)p.tasks.size();
        counts[3] = p.n_tasks_global; counts[4] = p.n_tasks_masks; counts[5] = (int64_t)p.delta_list.size();
        counts[7] = 0; counts[8] = 0;
        for (const TaskHdr& t : p.tasks) {
            if (t.flags & TK_LONG) ++counts[7];
            if ((t.flags & TK_LANES) && (t.flags & TK_REC_GLOBAL)) ++counts[8];      // lane form over an in-place record
        }
    }
    if (p.max_indeg <= 255 && !force_pred32)
        return run<uint16_t>(p, sh, trace_T, no_pack, counts ? counts + 6 : nullptr, sink_value, sink_s_het, p1_edges, n_p1, p2_edges, n_p2, level_checksum, level_live);
    return run<uint32_t>(p, sh, trace_T, no_pack, counts ? counts + 6 : nullptr, sink_value, sink_s_het, p1_edges, n_p1, p2_edges, n_p2, level_checksum, level_live);
}

// Row-sharded sweep over `n_ranks` emulated GPUs (dg_dip_create_sharded).  counts: [narrow, wide, wide tasks over all ranks, pushes]
extern "C" int emu_dp_diploid_sharded(int32_t n_levels, const int32_t* level_off, const int64_t* adj_off,
                                      const int32_t* adj_dst, const uint8_t* adj_w, const int64_t* col_off,
                                      const int32_t* col_val, const uint8_t* colour_is_hom, int32_t n_colours, int32_t R,
                                      int32_t n_ranks, int32_t grid, int32_t tile_cells,
                                      int32_t* sink_value, int32_t* sink_s_het, int32_t* p1_edges, int32_t* n_p1,
                                      int32_t* p2_edges, int32_t* n_p2, int64_t* counts) {
    DipGraphView g;
    g.n_levels = n_levels; g.level_off = level_off; g.adj_off = adj_off; g.adj_dst = adj_dst; g.adj_w = adj_w;
    g.col_off = col_off; g.col_val = col_val; g.colour_is_hom = colour_is_hom; g.n_colours = n_colours; g.R = R;
    DipPlan probe;
    if (!build_dip_plan(g, probe)) return -1;
    SweepShape sh;
    sh.grid = grid > 0 ? grid : 4; sh.threads = 480; sh.tile_cells = tile_cells > 0 ? tile_cells : 16384; sh.slot_bytes = 4096;
    const bool packed = probe.value_bound < KEY_VALUE_LIMIT;
    sh.lane_rc = (packed && R + 1 >= LANE_RC_BIG) ? LANE_RC_BIG : LANE_RC_SMALL;
    sh.allow_long = packed;
    if (probe.max_indeg <= 255)
        return run_sharded<uint16_t>(g, sh, n_ranks, 128, false, sink_value, sink_s_het, p1_edges, n_p1, p2_edges, n_p2, counts);
    return run_sharded<uint32_t>(g, sh, n_ranks, 128, false, sink_value, sink_s_het, p1_edges, n_p1, p2_edges, n_p2, counts);
}
