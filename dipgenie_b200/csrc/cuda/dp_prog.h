// dp_prog.h — the LEVEL PROGRAM of the diploid DP sweep (engine "v4"), shared verbatim by the device builder
// (dp_sweep4.cu: prog_fill_kernel), the sweep kernel (dp_sweep4.cu: dip_sweep4_kernel) and the CPU kernel-logic
// emulator of the `-m "not gpu"` tests (tests/emu/dp_emu4.cpp).  Nothing here is a fallback: the product only runs
// it on the device.
//
// Reference semantics (src/approximator.cpp:627-701), gather form (dp_cell.h): destination cell (r2,i',j') of level
// l+1 takes the first strict maximum, in (e1,e2) order, over in-edges (i -w1-> i') x (j -w2-> j') of
// src(r2-w1-w2, i, j) + delta(i,j,i',j').  Everything in that sentence except the DP values is known from the graph:
// which source cell a candidate reads, by how many layers it is shifted, its pair score, its rank in the tie-break
// order.  The sweep is a chain of 10^5 dependent levels, so whatever does not depend on the values is hoisted out of
// the chain: a fully parallel pre-pass expands every transition into a flat program of descriptors, and the
// sweep kernel only streams descriptors and layers.
//
// A destination vertex is S1 (one in-edge), M (two or more) or Z (none).  The cells of level l+1 are
//   * copy cells   S1 x S1            one candidate: dst = src shifted by w, plus delta; NO predecessor code is stored
//                                     (the code is implied: ordinal 0);
//   * multi cells  A = M x S1, B = S1 x M, C = M x M, enumerated in that order ("slot" t = position in this
//                                     enumeration, closed form below); candidates in (e1,e2) lexicographic order;
//                                     a 16-bit predecessor code (the winner's ordinal) per layer is stored at
//                                     pred[pred_off + r * n_multi + t];
//   * dead cells   anything with a Z vertex: stored as DEAD.
// Multi cells with at least PROG_BIG_MIN candidates are also listed in `big`: the kernel gives them a whole warp
// (lanes over candidates, one REDUX.MAX per layer); the others are evaluated one thread per cell.
//
// Packed keys: layers hold value << 10 (dp_cell.h: KEY_SHIFT); a candidate's key is layer value + (delta << 10) +
// (1023 - ordinal), so the reference's winner (larger value, then smaller i, then smaller j = smaller ordinal) is
// a plain integer maximum.  Dead cells hold DEAD = INT32_MIN plus whatever multiples of 1024 their ancestors' deltas
// added: the planner only selects this engine when the sum over transitions of the colours present stays below
// 2^21, so a dead cell can never climb to 0 and a live one can never overflow — no liveness tests in the loops.
// Two permanently dead padding layers sit below layer 0 of every tile, so that a weighted candidate of a low layer
// (r2 - w < 0) reads DEAD without a range check.
#pragma once
#include <cstddef>
#include <cstdint>

#include "dp_cell.h"

namespace dg {

constexpr int32_t V4_DEAD = INT32_MIN;
constexpr int V4_SHIFT = KEY_SHIFT;                       // 10
constexpr uint32_t V4_ORD_MASK = (1u << V4_SHIFT) - 1u;   // 1023
constexpr uint32_t PROG_BIG_MIN = 12;                     // multi cells with this many candidates go to the warp form
constexpr uint32_t PROG_MAX_CAND = 1024;                  // candidates per cell the packed ordinal can hold
constexpr int PROG_COMPACT_K = 32;                        // compact descriptors: both levels at most this wide
constexpr uint32_t PROG_COMPACT_DELTA = 1023;             // ... and at most this many colours on the two levels

enum : uint32_t {
    PF_COMPACT = 1,      // compact descriptors (10-bit cell indices)
    PF_SRC_SMEM = 2,     // level l lives in CTA 0's shared-memory tile
    PF_DST_SMEM = 4,     // level l+1 goes to CTA 0's shared-memory tile
    PF_STAGED = 8,       // the whole program travels into the ring slot (else only dir entry + header)
    PF_WAIT = 16,        // wait for dir.wait_target arrivals before the level
    PF_ARRIVE = 32,      // arrive after the level
    PF_ALL_CTAS = 64,    // every CTA of the problem takes part (both layers in HBM/L2)
};

// Directory entry of a transition (16 bytes, copied into the slot in front of the program).
struct ProgDir {
    uint32_t off16;        // program offset in the program buffer, units of 16 bytes
    uint32_t stage_bytes;  // bytes the producer copies after the entry: the whole program or just the header
    uint32_t wait_target;
    uint32_t flags;
};
static_assert(sizeof(ProgDir) == 16, "ProgDir layout");

struct ProgHdr {           // 64 bytes
    uint16_t k, k2;
    uint32_t n_copy, n_multi, n_cand, n_big, n_dead;
    uint32_t max_n;        // most candidates of a thread-form multi cell (loop bound hint)
    uint32_t rsv;
    uint64_t pred_off;     // u16 elements: codes of level l+1 start here, layout [layer][slot]
    uint64_t rsv2;
    uint32_t off_cell, off_cand, off_big, off_dead;   // section offsets from the header's start (prog_layout; the copy section follows the header)
};
static_assert(sizeof(ProgHdr) == 64, "ProgHdr layout");

DG_HD size_t a16(size_t x) { return (x + 15) & ~(size_t)15; }

// Section offsets inside a level program.
struct ProgLayout {
    size_t copy, cell, cand, big, dead, end;
};
DG_HD ProgLayout prog_layout(bool compact, uint64_t n_copy, uint64_t n_multi, uint64_t n_cand, uint64_t n_big, uint64_t n_dead) {
    ProgLayout o;
    o.copy = sizeof(ProgHdr);
    o.cell = o.copy + a16(n_copy * (compact ? 4 : 16));
    o.cand = o.cell + a16(n_multi * (compact ? 8 : 16));
    o.big = o.cand + a16(n_cand * (compact ? 4 : 8));
    o.dead = o.big + a16(n_big * 4);
    o.end = o.dead + a16(n_dead * 4);
    return o;
}

// ---- descriptors -----------------------------------------------------------------------------------------------
// compact copy  u32 : src[0:10) dst[10:20) w[20:22) delta[22:32)
// compact cell  u32x2: dst[0:10) | n << 16 ; cand_off (level-local candidate index)
// compact cand  u32 : src[0:10) w[10:12) delta[12:32)
// wide copy     u32x4: src | w << 30, dst, delta, 0
// wide cell     u32x4: dst, n, cand_off, 0
// wide cand     u32x2: src | w << 30, delta
struct CopyDesc { uint32_t src, dst, w, delta; };
struct CellDesc { uint32_t dst, n, cand_off; };
struct CandDesc { uint32_t src, w, delta; };

DG_HD uint32_t pack_copy_c(const CopyDesc& d) { return d.src | (d.dst << 10) | (d.w << 20) | (d.delta << 22); }
DG_HD CopyDesc unpack_copy_c(uint32_t x) { return {x & 1023u, (x >> 10) & 1023u, (x >> 20) & 3u, x >> 22}; }
DG_HD uint32_t pack_cand_c(const CandDesc& d) { return d.src | (d.w << 10) | (d.delta << 12); }
DG_HD CandDesc unpack_cand_c(uint32_t x) { return {x & 1023u, (x >> 10) & 3u, x >> 12}; }

// ---- per-level class tables (built once per problem, prog_classify) ------------------------------------------------
// For the vertices of one level, in position order: S1 positions first, then M, then Z (cls_list), the rank of every
// vertex within its class (vrank), and for the M vertices the running sum of their in-degrees (mpre, m+1 entries at
// cls offset of the level + n1).
struct LevelClass {
    uint32_t k2, n1, m, z, dm;       // dm = sum of in-degrees over M
    const uint16_t* list;            // [k2]  S1 | M | Z positions
    const uint32_t* mpre;            // [m+1] prefix of in-degrees over M in rank order
};

struct ProgCounts { uint64_t n_copy, n_multi, n_cand, n_big, n_dead; uint32_t max_n; };

// in_off: in-edge CSR offsets of the level's vertices (k2+1 entries, absolute); counts are closed-form except n_big.
DG_HD ProgCounts prog_counts(const LevelClass& c) {
    ProgCounts o;
    const uint64_t n1 = c.n1, m = c.m, dm = c.dm, k2 = c.k2, live = n1 + m;
    o.n_copy = n1 * n1;
    o.n_multi = 2 * m * n1 + m * m;
    o.n_cand = 2 * dm * n1 + dm * dm;
    o.n_dead = k2 * k2 - live * live;
    // big cells: A/B runs of an M vertex with d >= BIG_MIN (n1 cells each, twice), C cells with d_a * d_b >= BIG_MIN.
    // O(m + BIG_MIN^2): ge[x] = M vertices with in-degree >= x, has[x] = some M vertex has in-degree exactly x.
    uint32_t ge[PROG_BIG_MIN + 1], has[PROG_BIG_MIN + 1];
    for (uint32_t x = 0; x <= PROG_BIG_MIN; ++x) { ge[x] = 0; has[x] = 0; }
    for (uint32_t a = 0; a < c.m; ++a) {
        const uint32_t da = c.mpre[a + 1] - c.mpre[a];
        if (da < PROG_BIG_MIN) has[da] = 1;
        ++ge[da < PROG_BIG_MIN ? da : PROG_BIG_MIN];
    }
    for (uint32_t x = PROG_BIG_MIN; x-- > 0;) ge[x] += ge[x + 1];
    uint64_t nb = 0;
    uint32_t mx = 0;
    for (uint32_t a = 0; a < c.m; ++a) {
        const uint32_t da = c.mpre[a + 1] - c.mpre[a];
        if (da >= PROG_BIG_MIN) { nb += 2 * n1 + m; continue; }
        nb += ge[(PROG_BIG_MIN + da - 1) / da];
    }
    for (uint32_t x = 2; x < PROG_BIG_MIN; ++x) {
        if (!has[x]) continue;
        if (n1 && x > mx) mx = x;
        for (uint32_t y = 2; y < PROG_BIG_MIN && x * y < PROG_BIG_MIN; ++y) if (has[y] && x * y > mx) mx = x * y;
    }
    o.n_big = nb;
    o.max_n = mx;
    return o;
}

// What the builder needs of a transition l -> l+1.
struct ProgLevelIn {
    uint32_t k, k2;
    const int32_t* in_off;      // &in_off[level_off[l+1]] : k2+1 absolute offsets
    const uint32_t* in_edge;    // whole array: entry = source position | weight << 16
    LevelClass cls;
    int32_t W;                  // mask words (0: no colours on the two levels)
    const uint64_t* msrc;
    const uint64_t* mdst;
};

DG_HD void in_edge_at(const ProgLevelIn& L, uint32_t pos, uint32_t e, uint32_t& src_pos, uint32_t& w) {
    const uint32_t x = L.in_edge[L.in_off[pos] + (int32_t)e];
    src_pos = x & 0xFFFFu; w = x >> 16;
}

// copy cell t in [0, n1^2)
DG_HD CopyDesc make_copy(const ProgLevelIn& L, uint64_t t) {
    const uint32_t n1 = L.cls.n1;
    const uint32_t a = (uint32_t)(t / n1), b = (uint32_t)(t - (uint64_t)a * n1);
    const uint32_t i2 = L.cls.list[a], j2 = L.cls.list[b];
    uint32_t i, wi, j, wj;
    in_edge_at(L, i2, 0, i, wi);
    in_edge_at(L, j2, 0, j, wj);
    CopyDesc d;
    d.src = i * L.k + j; d.dst = i2 * L.k2 + j2; d.w = wi + wj;
    d.delta = L.W ? (uint32_t)mask_delta(L.W, L.msrc, L.mdst, (int)i, (int)j, (int)i2, (int)j2) : 0u;
    return d;
}

// multi cell t in [0, 2 m n1 + m^2): destination pair, candidate count, first candidate, in-degree of j'
struct MultiCell { uint32_t i2, j2, n, d2; uint64_t cand_off; };
DG_HD MultiCell multi_cell(const ProgLevelIn& L, uint64_t t) {
    const uint64_t n1 = L.cls.n1, m = L.cls.m, dm = L.cls.dm;
    const uint16_t* S = L.cls.list;
    const uint16_t* M = L.cls.list + n1;
    const uint32_t* P = L.cls.mpre;
    MultiCell c;
    if (t < m * n1) {                              // A: M row x S1 column
        const uint32_t a = (uint32_t)(t / n1), b = (uint32_t)(t - (uint64_t)a * n1);
        const uint32_t da = P[a + 1] - P[a];
        c.i2 = M[a]; c.j2 = S[b]; c.n = da; c.d2 = 1;
        c.cand_off = (uint64_t)P[a] * n1 + (uint64_t)b * da;
    } else if (t < 2 * m * n1) {                   // B: S1 row x M column
        const uint64_t u = t - m * n1;
        const uint32_t b = (uint32_t)(u / n1), a = (uint32_t)(u - (uint64_t)b * n1);
        const uint32_t db = P[b + 1] - P[b];
        c.i2 = S[a]; c.j2 = M[b]; c.n = db; c.d2 = db;
        c.cand_off = dm * n1 + (uint64_t)P[b] * n1 + (uint64_t)a * db;
    } else {                                       // C: M x M
        const uint64_t u = t - 2 * m * n1;
        const uint32_t a = (uint32_t)(u / m), b = (uint32_t)(u - (uint64_t)a * m);
        const uint32_t da = P[a + 1] - P[a], db = P[b + 1] - P[b];
        c.i2 = M[a]; c.j2 = M[b]; c.n = da * db; c.d2 = db;
        c.cand_off = 2 * dm * n1 + (uint64_t)P[a] * dm + (uint64_t)da * P[b];
    }
    return c;
}

// candidate `ord` of a multi cell: (e1, e2) = (ord / d2, ord % d2)
DG_HD CandDesc make_cand(const ProgLevelIn& L, const MultiCell& c, uint32_t ord) {
    const uint32_t e1 = ord / c.d2, e2 = ord - e1 * c.d2;
    uint32_t i, wi, j, wj;
    in_edge_at(L, c.i2, e1, i, wi);
    in_edge_at(L, c.j2, e2, j, wj);
    CandDesc d;
    d.src = i * L.k + j; d.w = wi + wj;
    d.delta = L.W ? (uint32_t)mask_delta(L.W, L.msrc, L.mdst, (int)i, (int)j, (int)c.i2, (int)c.j2) : 0u;
    return d;
}

// dead cell x in [0, k2^2 - (n1+m)^2): rows of Z vertices in full, then the Z columns of the other rows
DG_HD uint32_t dead_cell(const ProgLevelIn& L, uint64_t x) {
    const uint64_t k2 = L.k2, z = L.cls.z, live = L.cls.n1 + L.cls.m;
    const uint16_t* Z = L.cls.list + live;
    if (x < z * k2) {
        const uint32_t a = (uint32_t)(x / k2), j2 = (uint32_t)(x - (uint64_t)a * k2);
        return (uint32_t)(Z[a] * k2 + j2);
    }
    const uint64_t u = x - z * k2;
    const uint32_t a = (uint32_t)(u / z), b = (uint32_t)(u - (uint64_t)a * z);
    return (uint32_t)(L.cls.list[a] * k2 + Z[b]);      // list[0 .. live) = the S1 and M positions
}

// Kernel geometry the directory is made for.
struct Sweep4Shape {
    int slog = 10;             // shared-memory layer stride = 1 << slog cells
    int kn = 32;               // levels at most this wide live in shared memory (kn * kn <= 1 << slog)
    int slot_bytes = 8192;     // ring slot (directory entry + program)
    int nslot = 4;             // ring depth
    int grid = 1;              // CTAs of the problem
};

}  // namespace dg
