// dp_cell.h — the per-cell arithmetic of the diploid DP, shared verbatim by the CUDA kernels
// (dp_diploid.cu) and by the CPU kernel-logic emulator used in the `-m "not gpu"` tests
// (tests/emu/dp_emu.cpp).  Nothing here is a fallback: the product path only ever runs it on the device.
//
// Reference semantics (src/approximator.cpp:627-701): destination cell (r2,i',j') of level l+1 receives
//     max over edges (i -w1-> i'), (j -w2-> j'), r = r2-w1-w2 >= 0, src(r,i,j) live
//         of  src(r,i,j).value + delta(i,j,i',j')
// ties broken towards smaller i, then smaller j (:657-659).  Packed as one unsigned 64-bit key
//     key = value << 32 | (0xFFFF - i) << 16 | (0xFFFF - j)
// so the winner is a plain integer max; key == 0 means "no live candidate" (cell stays NEG_INF, :568).
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

// One transition (level l -> l+1) as the kernels see it.  OffT = int32_t when in_off/in_edge point into
// the global in-edge CSR, uint16_t when they point into a packed per-transition record staged in
// shared memory (dp_prep.h: RecHeader), where offsets are relative to the record's own edge array.
template <class OffT>
struct TransitionT {
    int32_t k;                 // |level l|
    int32_t k2;                // |level l+1|
    int32_t W;                 // 64-bit mask words per set (0: no colours on either level)
    const OffT* in_off;        // k2+1 entries, indexed by destination position
    const uint32_t* in_edge;   // entry = source position | weight << 16
    const uint64_t* msrc;      // [k ][2W]  hom words then het words
    const uint64_t* mdst;      // [k2][2W]
};
using Transition = TransitionT<int32_t>;

// ---- packed per-transition records (staged through shared memory by the sweep kernel) ----------
// record = RecHeader | in_off2 u16[k2+1] | pad4 | in_edge u32[n_in] | pad8 | msrc u64[k*2W] | mdst u64[k2*2W] | pad16
// in_off2 is relative to the record's own in_edge array, so a record is self-contained.
enum : uint16_t {
    REC_WAIT = 1,        // grid-level wait (counter >= wait_target) before the transition
    REC_ARRIVE = 2,      // grid-level arrive after the transition
    REC_SRC_SMEM = 4,    // source layer lives in CTA 0's shared-memory tile
    REC_DST_SMEM = 8,    // destination layer goes to CTA 0's shared-memory tile
};
enum : uint8_t { MODE_FAST = 0, MODE_STAGED = 1, MODE_GLOBAL = 2 };

struct RecHeader {               // 32 bytes
    uint16_t k, k2, W, flags;
    uint32_t n_in, bytes;        // bytes: whole record, multiple of 16
    uint32_t P, wait_target;
    int64_t pred_off2;           // offset (cells) of level l+1's predecessor codes
};
static_assert(sizeof(RecHeader) == 32, "RecHeader layout");

// View of a packed record (global or shared memory) as a transition.
DG_HD void record_view(const uint8_t* rec, RecHeader& h, TransitionT<uint16_t>& tr) {
    h = *reinterpret_cast<const RecHeader*>(rec);
    tr.k = h.k; tr.k2 = h.k2; tr.W = h.W;
    size_t o = sizeof(RecHeader);
    tr.in_off = reinterpret_cast<const uint16_t*>(rec + o);
    o += (((size_t)h.k2 + 1) * 2 + 3) & ~(size_t)3;
    tr.in_edge = reinterpret_cast<const uint32_t*>(rec + o);
    o += (size_t)h.n_in * 4; o = (o + 7) & ~(size_t)7;
    tr.msrc = reinterpret_cast<const uint64_t*>(rec + o);
    tr.mdst = tr.msrc + (size_t)h.k * 2 * h.W;
}

// delta = |(Hom u1 ∪ Hom v1) ∩ (Hom u2 ∪ Hom v2)| + |(Het u1 ∪ Het v1) △ (Het u2 ∪ Het v2)|
// (approximator.cpp:614-619).  `het_only` returns just the second term (dp_entry::s_het, :662).
template <class OffT>
DG_HD int pair_delta(const TransitionT<OffT>& t, int i, int j, int i2, int j2, bool het_only = false) {
    const int W = t.W;
    if (W == 0) return 0;
    const uint64_t* si = t.msrc + (int64_t)i * 2 * W;
    const uint64_t* sj = t.msrc + (int64_t)j * 2 * W;
    const uint64_t* di = t.mdst + (int64_t)i2 * 2 * W;
    const uint64_t* dj = t.mdst + (int64_t)j2 * 2 * W;
    int acc = 0;
    for (int w = 0; w < W; ++w) {
        if (!het_only) acc += popc64((si[w] | sj[w]) & (di[w] | dj[w]));
        acc += popc64((si[W + w] | sj[W + w]) ^ (di[W + w] | dj[W + w]));
    }
    return acc;
}

DG_HD uint64_t pack_key(int32_t value, int i, int j) {
    return ((uint64_t)(uint32_t)value << 32) | ((uint64_t)(0xFFFFu - (uint32_t)i) << 16) | (uint64_t)(0xFFFFu - (uint32_t)j);
}
DG_HD int32_t key_value(uint64_t key) { return key ? (int32_t)(uint32_t)(key >> 32) : NEG_INF; }

// Gather for one destination cell.  `load(idx)` returns the source-layer value at flat index
// (r*k + i)*k + j.  Returns the winning key (0 = dead cell); code = e1 << 16 | e2.
template <class OffT, class Load>
DG_HD uint64_t relax_cell(const TransitionT<OffT>& t, Load load, int r2, int i2, int j2, uint32_t& code) {
    const int32_t a0 = (int32_t)t.in_off[i2], a1 = (int32_t)t.in_off[i2 + 1];
    const int32_t b0 = (int32_t)t.in_off[j2], b1 = (int32_t)t.in_off[j2 + 1];
    uint64_t best = 0;
    uint32_t best_code = 0xFFFFFFFFu;
    for (int32_t e1 = a0; e1 < a1; ++e1) {
        const uint32_t x = t.in_edge[e1];
        const int i = (int)(x & 0xFFFFu), wu = (int)(x >> 16);
        const int ra = r2 - wu;
        if (ra < 0) continue;
        for (int32_t e2 = b0; e2 < b1; ++e2) {
            const uint32_t y = t.in_edge[e2];
            const int j = (int)(y & 0xFFFFu), wv = (int)(y >> 16);
            const int r = ra - wv;
            if (r < 0) continue;
            const int32_t s = load(((int64_t)r * t.k + i) * t.k + j);
            if (s == NEG_INF) continue;
            const uint64_t key = pack_key(s + pair_delta(t, i, j, i2, j2), i, j);
            if (key > best) { best = key; best_code = ((uint32_t)(e1 - a0) << 16) | (uint32_t)(e2 - b0); }
        }
    }
    code = best_code;
    return best;
}

// Gather for one destination pair (i',j') and RC consecutive layers r2 = r0 .. r0+RC-1 at once
// ("pair-major" form used by the kernels).  The in-edge decode, the candidate's flat source offset, its
// tie-break bits and its colour delta do not depend on r2, so they are computed once per candidate and
// reused for all RC layers; the RC source loads of a candidate are independent (ILP).  For every layer
// the winner is the same lexicographic max as relax_cell().  best[rr] == 0 means layer r0+rr is dead
// (or beyond R).  code[rr] = e1 << 16 | e2.
template <int RC, bool HAS_MASK, class OffT, class Load>
DG_HD void relax_pair(const TransitionT<OffT>& t, Load load, int R, int r0, int i2, int j2,
                      uint64_t (&best)[RC], uint32_t (&code)[RC]) {
    const int32_t a0 = (int32_t)t.in_off[i2], a1 = (int32_t)t.in_off[i2 + 1];
    const int32_t b0 = (int32_t)t.in_off[j2], b1 = (int32_t)t.in_off[j2 + 1];
    const int64_t kk = (int64_t)t.k * t.k;
#pragma unroll
    for (int rr = 0; rr < RC; ++rr) { best[rr] = 0; code[rr] = 0xFFFFFFFFu; }
    for (int32_t e1 = a0; e1 < a1; ++e1) {
        const uint32_t x = t.in_edge[e1];
        const int i = (int)(x & 0xFFFFu), wu = (int)(x >> 16);
        for (int32_t e2 = b0; e2 < b1; ++e2) {
            const uint32_t y = t.in_edge[e2];
            const int j = (int)(y & 0xFFFFu), w = wu + (int)(y >> 16);
            const int64_t base = (int64_t)i * t.k + j;
            const uint64_t low = ((uint64_t)(0xFFFFu - (uint32_t)i) << 16) | (uint64_t)(0xFFFFu - (uint32_t)j);
            const int d = HAS_MASK ? pair_delta(t, i, j, i2, j2) : 0;
            const uint32_t cd = ((uint32_t)(e1 - a0) << 16) | (uint32_t)(e2 - b0);
#pragma unroll
            for (int rr = 0; rr < RC; ++rr) {
                const int r = r0 + rr - w;
                if (r >= 0 && r0 + rr <= R) {
                    const int32_t s = load((int64_t)r * kk + base);
                    if (s != NEG_INF) {
                        const uint64_t key = ((uint64_t)(uint32_t)(s + d) << 32) | low;
                        if (key > best[rr]) { best[rr] = key; code[rr] = cd; }
                    }
                }
            }
        }
    }
}

// Layers handled per work item: as many as possible (amortises the decode) while still giving every
// one of `nthreads` threads an item.
DG_HD int choose_rc(uint64_t npairs, int R, uint64_t nthreads) {
    int rc = 8;
    while (rc > 1 && npairs * (uint64_t)((R + rc) / rc) < nthreads) rc >>= 1;
    return rc;
}

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
};

// Walks the predecessor codes from the sink cell (r=R,0,0) back to level 0 and emits what the
// reference keeps as linked lists (approximator.cpp:666-692, materialize_edges :757-764):
// every edge with weight>0 on P1 (resp. P2), and the final edge into the sink level on both.
// Edges come out newest-first into p1/p2 (capacity cap pairs); the caller reverses them.
// Returns 0, or -1 if the sink cell is dead, -2 on capacity overflow.
template <class PredT>
DG_HD int traceback(const TraceView& v, const PredT* pred, int32_t sink_value,
                    int32_t* p1, int32_t* n1_out, int32_t* p2, int32_t* n2_out, int cap, int32_t* s_het_out) {
    int n1 = 0, n2 = 0, s_het = 0;
    *n1_out = 0; *n2_out = 0; *s_het_out = 0;
    if (sink_value == NEG_INF) return -1;
    constexpr int SH = (sizeof(PredT) == 2) ? 8 : 16;
    constexpr uint32_t MK = (sizeof(PredT) == 2) ? 0xFFu : 0xFFFFu;
    int r = v.R, i2 = 0, j2 = 0;
    for (int l = v.L - 2; l >= 0; --l) {
        const int32_t lo = v.level_off[l], mid = v.level_off[l + 1];
        const int32_t k = mid - lo, k2 = v.level_off[l + 2] - mid;
        const uint32_t code = (uint32_t)pred[v.pred_off[l + 1] + ((int64_t)r * k2 + i2) * k2 + j2];
        const int32_t e1 = v.in_off[mid + i2] + (int32_t)((code >> SH) & MK);
        const int32_t e2 = v.in_off[mid + j2] + (int32_t)(code & MK);
        const uint32_t x = v.in_edge[e1], y = v.in_edge[e2];
        const int i = (int)(x & 0xFFFFu), wu = (int)(x >> 16);
        const int j = (int)(y & 0xFFFFu), wv = (int)(y >> 16);
        if (l + 1 == v.L - 1) {   // both lists get the edge into the sink level (:684-692)
            if (n1 >= cap || n2 >= cap) return -2;
            p1[2 * n1] = lo + i; p1[2 * n1 + 1] = mid + i2; ++n1;
            p2[2 * n2] = lo + j; p2[2 * n2 + 1] = mid + j2; ++n2;
        }
        if (wu > 0) { if (n1 >= cap) return -2; p1[2 * n1] = lo + i; p1[2 * n1 + 1] = mid + i2; ++n1; }
        if (wv > 0) { if (n2 >= cap) return -2; p2[2 * n2] = lo + j; p2[2 * n2 + 1] = mid + j2; ++n2; }
        if (v.lvlW[l] > 0) {
            Transition t;
            t.k = k; t.k2 = k2; t.W = v.lvlW[l]; t.in_off = nullptr; t.in_edge = nullptr;
            t.msrc = v.masks + v.msrc_off[l]; t.mdst = v.masks + v.mdst_off[l];
            s_het += pair_delta(t, i, j, i2, j2, true);
        }
        r -= wu + wv; i2 = i; j2 = j;
    }
    *n1_out = n1; *n2_out = n2; *s_het_out = s_het;
    return 0;
}

}  // namespace dg
