// dp_cell.h — the per-cell arithmetic and the task format of the diploid DP, shared verbatim by the
// CUDA kernels (dp_diploid.cu) and by the CPU kernel-logic emulator used in the `-m "not gpu"` tests
// (tests/emu/dp_emu.cpp).  Nothing here is a fallback: the product path only ever runs it on the device.
//
// Reference semantics (src/approximator.cpp:627-701): destination cell (r2,i',j') of level l+1 receives
//     max over edges (i -w1-> i'), (j -w2-> j'), r = r2-w1-w2 >= 0, src(r,i,j) live
//         of  src(r,i,j).value + delta(i,j,i',j')
// ties broken towards smaller i, then smaller j (:657-659).  In-edge lists are kept in ascending source
// position, so the *first strict maximum* in (e1, e2) iteration order is exactly that winner.  Live
// values are >= 0 (level 0 starts at 0, :535; deltas are set sizes), dead cells hold NEG_INF, so a
// running maximum initialised to -1 never accepts a dead source (NEG_INF + delta << -1).
// The winner's in-edge ordinals (e1 within in(i'), e2 within in(j')) are the predecessor code.
#pragma once
#include <cstddef>
#include <cstdint>

#if defined(__CUDACC__)
#define DG_HD __host__ __device__ __forceinline__
#else
#define DG_HD inline
#endif

namespace dg {

constexpr int32_t NEG_INF = INT32_MIN / 4;   // approximator.cpp:413

DG_HD int popc64(uint64_t x) {
#if defined(__CUDA_ARCH__)
    return __popcll(x);
#else
    return __builtin_popcountll(x);
#endif
}

// ---- tasks ---------------------------------------------------------------------------------------
// The host planner (dp_prep.cpp: plan_tasks) compiles the sweep into one task stream per CTA.  A task
// is "rows [i0,i1) of the destination level of transition `level`"; the 128-byte header below travels
// with the task's record into a shared-memory slot (three bulk copies on one mbarrier: header, record,
// delta slice).  Narrow transitions are one task on CTA 0 with both DP layers in shared memory; wide
// ones are split by destination row over P CTAs with the layers in HBM/L2 and closed by a
// monotone-counter grid barrier.
enum : uint32_t {
    TK_WAIT = 1,           // grid-level wait (counter >= wait_target) before the task
    TK_ARRIVE = 2,         // grid-level arrive after the task (implies TK_BAR)
    TK_SRC_SMEM = 4,       // source layer lives in CTA 0's shared-memory tile
    TK_DST_SMEM = 8,       // destination layer goes to CTA 0's shared-memory tile
    TK_BAR = 16,           // block barrier after the task (last task of this CTA for the level)
    TK_DELTA = 32,         // the transition has a pair-score matrix (colours on level l or l+1)
    TK_DELTA_STAGED = 64,  // ... and this task's rows of it are staged in the slot
    TK_REC_GLOBAL = 128,   // in-edge lists are read in place from the global CSR (record too big for a slot)
    TK_DELTA_MASKS = 256,  // no matrix was materialised (too wide): popcount the colour masks on the fly
    TK_LANES = 512,        // lane form (below); otherwise the pair form
    TK_LONG = 1024,        // lane form with destinations of more than 32 in-edges (slice blocks + scratch combine)
    TK_PUSH = 4096,        // row-sharded sweep over several GPUs: after the CTA's rows of the level are whole, copy rows
                           // [push_i0, push_i1) of the destination layer and of its predecessor codes to every peer GPU
};

// Two evaluation forms of a task, same results:
//  * lane form (TK_LANES, the fast path): a warp takes a group of `rp` destination rows x one block of <= 32
//    in-edges e2 of the destination level; lane = (row slot, e2).  Every lane walks the in-edges e1 of its row
//    (same trip count for the lanes of a row: no divergence inside a row) and keeps the first strict maximum
//    for its e2; the lanes of one destination column j' (adjacent: in-edges are grouped by destination, and a
//    block never cuts a group) are then combined by a segmented warp-shuffle maximum (value desc, code asc),
//    and the first lane of each group stores the cell.  Needs: staged record, pair scores absent or staged,
//    every destination of the level with >= 1 in-edges.  A destination with more than 32 in-edges (a
//    recombination vertex of a panel with more than 32 walks) is cut into slice blocks of 32 lanes; each slice is
//    reduced over the whole warp and the slices of one cell meet in a 64-bit shared-memory scratch word
//    (atomic maximum of value+1 << 32 | ~code), which a second pass after a block barrier turns into the cell.
//  * pair form: one thread per (destination pair, chunk of layers) looping over in(i') x in(j') — any shape,
//    any placement of records and scores (the general fallback).
struct TaskHdr {             // 128 bytes
    // fetch part (read by the producer lane before the copies are issued)
    uint32_t rec_off16;      // record offset in `records`, units of 16 bytes
    uint32_t rec_bytes;      // multiple of 16; 0: nothing staged besides the header
    uint32_t delta_off16;    // 16-byte-aligned start of the staged delta slice in the delta buffer, units of 16 bytes
    uint32_t delta_bytes;    // multiple of 16; 0: no staged slice
    // exec part
    int32_t level;           // transition level -> level+1
    uint32_t flags;
    uint32_t wait_target;
    uint32_t delta_skew;     // u16 elements between the staged slice start and the first delta row of row i0
    uint16_t k, k2;          // widths of the two levels
    uint16_t i0, i1;         // destination rows of this task
    uint32_t n_in;           // in-edges of level l+1 (valid unless TK_REC_GLOBAL)
    uint32_t n_active;       // threads of the CTA that own work (warps beyond skip the task)
    int64_t pred_off2;       // offset (cells) of level l+1's predecessor codes
    uint32_t m_k2;           // pair form: magic of k2
    uint32_t m_pairs;        //            magic of npairs = (i1-i0) * k2
    uint16_t rc;             // layers per work item (pair form: DIP_RC; lane form: LANE_RC_SMALL / LANE_RC_BIG)
    uint16_t groups;         // pair form: thread groups sharing the pairs (>= 1); chunks are dealt round-robin
    uint16_t nblk;           // lane form: blocks of the in-edge range
    uint16_t rp;             //            destination rows per warp item (> 1 only when nblk == 1)
    uint32_t nrg;            //            row groups = ceil((i1-i0) / rp)
    uint32_t m_nblk, m_nrg;  //            magics of nblk, nrg
    uint32_t m_nin;          //            magic of n_in (lane -> row slot)
    uint32_t n_witems;       //            warp items = nrg * nblk * chunks
    uint32_t rounds;         //            shuffle rounds of the segmented maximum = ceil(log2(longest group))
    uint32_t bstart_off;     //            byte offset of bstart[] inside the record
    uint32_t n_long;         //            destinations with more than 32 in-edges (TK_LONG), <= LANE_MAX_LONG
    uint32_t long_off;       //            byte offset of long_j[] (their positions, u16) inside the record
    uint16_t push_i0, push_i1;   // TK_PUSH: the CTA's whole row range of the level (byte 108)
    // row-sharded sweep, TK_ARRIVE: the CTAs of one rank first count themselves on a local word; the one that brings it
    // to arrive_local_target (cumulative over the levels) forwards arrive_n arrivals to every rank's counter
    uint32_t arrive_local_target;   // byte 112
    uint32_t arrive_n;              // byte 116
    uint32_t pad[2];
};
static_assert(sizeof(TaskHdr) == 128, "TaskHdr layout");

// x / d for 0 <= x < 2^16 * ... : with m = floor(2^32/d)+1 the product (x*m)>>32 is exact whenever
// x * d < 2^32 (here x < 2^20 threads/pairs and d < 2^12 in every use; the planner checks).
DG_HD uint32_t div_magic(uint32_t x, uint32_t m) {
#if defined(__CUDA_ARCH__)
    return __umulhi(x, m);
#else
    return (uint32_t)(((uint64_t)x * m) >> 32);
#endif
}
inline uint32_t make_magic(uint32_t d) { return d <= 1 ? 0u : (uint32_t)((1ull << 32) / d) + 1u; }

// One transition (level l -> l+1) as the kernels see it.  OffT = uint16_t when in_off points into a
// staged record (offsets relative to the record's own in_edge array), int32_t when in_off / in_edge
// are the global in-edge CSR (absolute offsets).
template <class OffT>
struct TransitionT {
    int32_t k;                 // |level l|
    int32_t k2;                // |level l+1|
    const OffT* in_off;        // indexed by destination position, k2+1 entries
    const uint32_t* in_edge;   // entry = source position | weight << 16
    // pair-score matrix D[e1][e2] over the in-edges of level l+1 (approximator.cpp:604-624 precomputed):
    // delta of candidate (e1,e2) = delta[(e1 - e1_base) * dstride + (e2 - e2_base)]
    const uint16_t* delta;
    int32_t dstride, e1_base, e2_base;
    int32_t dshift;            // layers hold value << dshift (0 or KEY_SHIFT): deltas are shifted alike
    // on-the-fly form (TK_DELTA_MASKS): colour bit-masks of the two levels
    int32_t W;                 // 64-bit mask words per set
    const uint64_t* msrc;      // [k ][2W]  hom words then het words
    const uint64_t* mdst;      // [k2][2W]
};

// record = in_off2 u16[k2+1] | pad16 | in_edge u32[n_in] | pad16 | in_dst u16[n_in] | pad16 | bstart u16[nblk+1] | pad16
//          | long_j u16[n_long] | pad16
// (16-byte aligned, self-contained; in_dst = destination position of every in-edge, bstart = first in-edge of
// every lane-form block, bstart[nblk] = n_in, long_j = destinations with more than 32 in-edges)
DG_HD size_t rec_edge_offset(int k2) { return (((size_t)k2 + 1) * 2 + 15) & ~(size_t)15; }
DG_HD size_t rec_dst_offset(int k2, int64_t n_in) { return rec_edge_offset(k2) + (((size_t)n_in * 4 + 15) & ~(size_t)15); }
DG_HD size_t rec_bstart_offset(int k2, int64_t n_in) { return rec_dst_offset(k2, n_in) + (((size_t)n_in * 2 + 15) & ~(size_t)15); }
DG_HD size_t rec_long_offset(int k2, int64_t n_in, int nblk) { return rec_bstart_offset(k2, n_in) + ((((size_t)nblk + 1) * 2 + 15) & ~(size_t)15); }
DG_HD size_t rec_bytes_for(int k2, int64_t n_in, int nblk, int n_long) { return rec_long_offset(k2, n_in, nblk) + (((size_t)n_long * 2 + 15) & ~(size_t)15); }
constexpr int LANE_MAX_LONG = 16;            // long destinations per level the lane form accepts
constexpr int LANE_SCRATCH_ENTRIES = 1024;   // 64-bit scratch words per CTA: rows x long destinations x layers of a task

// delta = |(Hom u1 ∪ Hom v1) ∩ (Hom u2 ∪ Hom v2)| + |(Het u1 ∪ Het v1) △ (Het u2 ∪ Het v2)|
// (approximator.cpp:614-619).  `het_only` returns just the second term (dp_entry::s_het, :662).
DG_HD int mask_delta(int W, const uint64_t* msrc, const uint64_t* mdst, int i, int j, int i2, int j2, bool het_only = false) {
    if (W == 0) return 0;
    const uint64_t* si = msrc + (int64_t)i * 2 * W;
    const uint64_t* sj = msrc + (int64_t)j * 2 * W;
    const uint64_t* di = mdst + (int64_t)i2 * 2 * W;
    const uint64_t* dj = mdst + (int64_t)j2 * 2 * W;
    int acc = 0;
    for (int w = 0; w < W; ++w) {
        if (!het_only) acc += popc64((si[w] | sj[w]) & (di[w] | dj[w]));
        acc += popc64((si[W + w] | sj[W + w]) ^ (di[W + w] | dj[W + w]));
    }
    return acc;
}

// MASKS = false: delta comes from the precomputed matrix (t.delta, null when the transition has no colours:
// a warp-uniform predicate, so both cases share one code path); MASKS = true: masks on the fly.
// Gather for one destination pair (i',j') and RC consecutive layers r2 = r0 .. r0+RC-1 at once.  The
// in-edge decode, the candidate's flat source offset and its delta do not depend on r2, so they are
// computed once per candidate; the RC source loads of a candidate are independent (ILP).
// `load(idx)` returns the source-layer value at flat index (r*k + i)*k + j.
// best[rr] < 0 means layer r0+rr is dead (or beyond R).  code[rr] = e1 ordinal << 16 | e2 ordinal.
template <int RC, bool MASKS, class IdxT, class OffT, class Load>
DG_HD void relax_pair(const TransitionT<OffT>& t, Load load, int R, int r0, int i2, int j2,
                      int32_t (&best)[RC], uint32_t (&code)[RC]) {
    const int32_t a0 = (int32_t)t.in_off[i2], a1 = (int32_t)t.in_off[i2 + 1];
    const int32_t b0 = (int32_t)t.in_off[j2], b1 = (int32_t)t.in_off[j2 + 1];
    const IdxT kk = (IdxT)t.k * (IdxT)t.k;
    const bool has_matrix = !MASKS && t.delta != nullptr;
#pragma unroll
    for (int rr = 0; rr < RC; ++rr) { best[rr] = -1; code[rr] = 0xFFFFFFFFu; }
    for (int32_t e1 = a0; e1 < a1; ++e1) {
        const uint32_t x = t.in_edge[e1];
        const int i = (int)(x & 0xFFFFu), wu = (int)(x >> 16);
        const uint16_t* drow = t.delta + ((int64_t)(e1 - t.e1_base) * t.dstride - t.e2_base);
        for (int32_t e2 = b0; e2 < b1; ++e2) {
            const uint32_t y = t.in_edge[e2];
            const int j = (int)(y & 0xFFFFu), w = wu + (int)(y >> 16);
            const IdxT base = (IdxT)i * (IdxT)t.k + (IdxT)j;
            int d = 0;
            if (has_matrix) d = (int)drow[e2];
            if (MASKS) d = mask_delta(t.W, t.msrc, t.mdst, i, j, i2, j2);
            d <<= t.dshift;
            const uint32_t cd = ((uint32_t)(e1 - a0) << 16) | (uint32_t)(e2 - b0);
            // the RC source loads are issued unconditionally (layer index clamped into [0,R]) so that they
            // are independent and overlap; validity only gates the compare
            int32_t v[RC];
#pragma unroll
            for (int rr = 0; rr < RC; ++rr) {
                int r = r0 + rr - w;
                r = r < 0 ? 0 : (r > R ? R : r);
                v[rr] = load((IdxT)r * kk + base);
            }
#pragma unroll
            for (int rr = 0; rr < RC; ++rr) {
                const bool ok = (r0 + rr - w >= 0) && (r0 + rr <= R);
                const int32_t c = v[rr] + d;
                if (ok && c > best[rr]) { best[rr] = c; code[rr] = cd; }
            }
        }
    }
}

// Layers per work item.  One value for every task: the sweep kernel has to stay small (consecutive
// levels would otherwise keep switching between unrolled variants) and register-lean; 4 layers give the
// narrow levels enough items to spread over the CTA and still amortise the in-edge decode.
constexpr int DIP_RC = 4;
DG_HD int choose_rc(uint32_t, int, uint32_t) { return DIP_RC; }
// Layers per lane in the lane form (SweepShape::lane_rc).  Measured on B200 (MHC_4, R=18, 4 CTAs per sample):
// unpacked arithmetic 4 / 8 / 20 layers -> 283 / 321 / 564 ms (register pressure, long serial items); packed keys
// 4 / 5 / 8 / 10 layers -> 296 / 287 / 292 / 273 ms.  The planner takes LANE_RC_BIG with packed keys when there are
// at least that many layers, else LANE_RC_SMALL.
constexpr int LANE_RC_SMALL = 4, LANE_RC_BIG = 10;

// Packed keys.  When every DP value is known to stay below 2^21 (the planner bounds it by the sum over
// transitions of the colours present, DipPlan::value_bound), the layers hold value << KEY_SHIFT and the lane
// form keeps ONE word per layer: key = (value << 10) | (31 - e1 ordinal) << 5 | (31 - e2 ordinal).  The first
// strict maximum in (e1,e2) order is then a plain integer maximum (larger value, then smaller e1, then smaller
// e2), per lane and across the lanes of a destination column: one add + one max per layer instead of an add,
// three compares and two selects, and one shuffle instead of two in the segmented maximum.  Dead cells stay
// NEG_INF (NEG_INF + anything below 2^27 is still negative and never beats the initial -1).  shift == 0 selects
// the unpacked arithmetic (values of any size).
constexpr int KEY_SHIFT = 10, KEY_ORD_BITS = 5;
constexpr int64_t KEY_VALUE_LIMIT = (int64_t)1 << (31 - KEY_SHIFT);

// Same fold as oracle/ref_hook.h::dg_ref_level_done, for one live cell; the per-level checksum is the
// wrapping sum of these plus the FNV offset basis.
DG_HD uint64_t cell_fold(uint64_t flat, int32_t value, int pred_i, int pred_j) {
    uint64_t x = flat * 0x9E3779B97F4A7C15ull;
    x ^= (uint64_t)(uint32_t)value * 0xC2B2AE3D27D4EB4Full;
    x ^= ((uint64_t)(uint32_t)pred_i << 32 | (uint32_t)pred_j) * 0x165667B19E3779F9ull;
    x ^= x >> 29;
    return x * 0xBF58476D1CE4E5B9ull;
}
constexpr uint64_t FOLD_BASIS = 1469598103934665603ull;

// Slot of a multi-candidate cell in the level-program engine's code array (dp_prog.h: enumeration A = M x S1,
// B = S1 x M, C = M x M); vinfo = rank in class | class << 30 (0 S1, 1 M, 2 Z).
DG_HD int64_t multi_slot_of(uint32_t vi_i, uint32_t vi_j, uint32_t n1, uint32_t m) {
    const uint32_t ci = vi_i >> 30, cj = vi_j >> 30, ri = vi_i & 0x3FFFFFFFu, rj = vi_j & 0x3FFFFFFFu;
    if (ci == 1 && cj == 0) return (int64_t)ri * n1 + rj;
    if (ci == 0 && cj == 1) return (int64_t)m * n1 + (int64_t)rj * n1 + ri;
    return 2 * (int64_t)m * n1 + (int64_t)ri * m + rj;
}

// Arrays the traceback needs (device or host pointers).
struct TraceView {
    int32_t L, R;
    const int32_t* level_off;
    const int32_t* in_off;
    const uint32_t* in_edge;
    const int32_t* lvlW;
    const int64_t* msrc_off;
    const int64_t* mdst_off;
    const uint64_t* masks;
    const int64_t* pred_off;
    // level-program engine (dp_prog.h): codes exist for multi-candidate cells only, one u16 per (layer, slot), the
    // winner's ordinal in (e1,e2) order as 1023 - (code & 1023) (cells of more than 1024 candidates: the ordinal itself);
    // vinfo == nullptr selects the task-stream layout
    const uint32_t* vinfo = nullptr;    // [V] rank in class | class << 30
    const uint32_t* lvl_n1 = nullptr;   // [L] S1 vertices of the level
    const uint32_t* lvl_m = nullptr;    // [L] M vertices of the level
    int32_t RL = 0;                     // layers per level in the code array
    // [V] fast path of a step (device traceback): bit 31 set = the vertex has exactly ONE in-edge, whose `pos | w << 16`
    // sits in the low bits — a cell of two such vertices steps with two independent loads and no code (nullable)
    const uint32_t* vup = nullptr;
};

struct TraceState { int32_t r, i2, j2; };   // cell (r, i2, j2) of some level

// One step of the walk: from cell `s` of level l+1 to its winning source cell of level l.  Returns false
// when the cell is dead.  wu/wv = weights of the two edges taken, (i,j) = source positions.
template <class PredT>
DG_HD bool trace_step(const TraceView& v, const PredT* pred, int l, TraceState& s, int& wu, int& wv, int& i2_old, int& j2_old) {
    constexpr int SH = (sizeof(PredT) == 2) ? 8 : 16;
    constexpr uint32_t MK = (sizeof(PredT) == 2) ? 0xFFu : 0xFFFFu;
    const int32_t mid = v.level_off[l + 1];
    if (v.vinfo) {
        if (v.vup) {
            const uint32_t ux = v.vup[mid + s.i2], uy = v.vup[mid + s.j2];
            if ((ux & uy) >> 31) {
                wu = (int)((ux >> 16) & 0x7FFFu); wv = (int)((uy >> 16) & 0x7FFFu);
                i2_old = s.i2; j2_old = s.j2;
                s.r -= wu + wv; s.i2 = (int)(ux & 0xFFFFu); s.j2 = (int)(uy & 0xFFFFu);
                return s.r >= 0;
            }
        }
        const int32_t a0 = v.in_off[mid + s.i2], d1 = v.in_off[mid + s.i2 + 1] - a0;
        const int32_t b0 = v.in_off[mid + s.j2], d2 = v.in_off[mid + s.j2 + 1] - b0;
        if (d1 <= 0 || d2 <= 0) return false;
        int32_t ord = 0;
        if (d1 * d2 > 1) {
            const uint32_t n1 = v.lvl_n1[l + 1], m = v.lvl_m[l + 1];
            const int64_t slot = multi_slot_of(v.vinfo[mid + s.i2], v.vinfo[mid + s.j2], n1, m);
            const int64_t nm = 2 * (int64_t)m * n1 + (int64_t)m * m;
            const uint32_t raw = (uint32_t)pred[v.pred_off[l + 1] + (int64_t)s.r * nm + slot];
            ord = d1 * d2 > 1024 ? (int32_t)raw : (int32_t)(1023u - (raw & 1023u));     // (dp_prog.h: PROG_KEY_CAND)
            if (ord >= d1 * d2) ord = d1 * d2 - 1;       // (codes of dead cells are arbitrary; only live paths are ever followed)
        }
        const int32_t o1 = ord / d2, o2 = ord - o1 * d2;
        const uint32_t x = v.in_edge[a0 + o1], y = v.in_edge[b0 + o2];
        wu = (int)(x >> 16); wv = (int)(y >> 16);
        i2_old = s.i2; j2_old = s.j2;
        s.r -= wu + wv; s.i2 = (int)(x & 0xFFFFu); s.j2 = (int)(y & 0xFFFFu);
        return s.r >= 0;
    }
    const int32_t k2 = v.level_off[l + 2] - mid;
    const PredT raw = pred[v.pred_off[l + 1] + ((int64_t)s.r * k2 + s.i2) * k2 + s.j2];
    if (raw == (PredT) ~(PredT)0) return false;
    const uint32_t code = (uint32_t)raw;
    const int32_t e1 = v.in_off[mid + s.i2] + (int32_t)((code >> SH) & MK);
    const int32_t e2 = v.in_off[mid + s.j2] + (int32_t)(code & MK);
    const uint32_t x = v.in_edge[e1], y = v.in_edge[e2];
    wu = (int)(x >> 16); wv = (int)(y >> 16);
    i2_old = s.i2; j2_old = s.j2;
    s.r -= wu + wv; s.i2 = (int)(x & 0xFFFFu); s.j2 = (int)(y & 0xFFFFu);
    return true;
}

// Walks the predecessor codes from cell `s` of level l_hi down to level l_lo and emits what the
// reference keeps as linked lists (approximator.cpp:666-692, materialize_edges :757-764): every edge with
// weight > 0 on P1 (resp. P2), and the final edge into the sink level on both.  Edges come out
// newest-first into p1/p2 (capacity cap pairs).  Returns 0, or -1 on a dead cell, -2 on capacity overflow.
template <class PredT>
DG_HD int trace_segment(const TraceView& v, const PredT* pred, int l_hi, int l_lo, TraceState& s,
                        int32_t* p1, int32_t* n1_out, int32_t* p2, int32_t* n2_out, int cap, int32_t* s_het_out) {
    int n1 = 0, n2 = 0, s_het = 0;
    for (int l = l_hi - 1; l >= l_lo; --l) {
        const int32_t lo = v.level_off[l], mid = v.level_off[l + 1];
        int wu, wv, i2, j2;
        if (!trace_step<PredT>(v, pred, l, s, wu, wv, i2, j2)) return -1;
        if (l + 1 == v.L - 1) {   // both lists get the edge into the sink level (:684-692)
            if (n1 >= cap || n2 >= cap) return -2;
            p1[2 * n1] = lo + s.i2; p1[2 * n1 + 1] = mid + i2; ++n1;
            p2[2 * n2] = lo + s.j2; p2[2 * n2 + 1] = mid + j2; ++n2;
        }
        if (wu > 0) { if (n1 >= cap) return -2; p1[2 * n1] = lo + s.i2; p1[2 * n1 + 1] = mid + i2; ++n1; }
        if (wv > 0) { if (n2 >= cap) return -2; p2[2 * n2] = lo + s.j2; p2[2 * n2 + 1] = mid + j2; ++n2; }
        if (v.lvlW[l] > 0)
            s_het += mask_delta(v.lvlW[l], v.masks + v.msrc_off[l], v.masks + v.mdst_off[l], s.i2, s.j2, i2, j2, true);
    }
    *n1_out = n1; *n2_out = n2; *s_het_out = s_het;
    return 0;
}

}  // namespace dg
