// dp_prog.h — the LEVEL PROGRAM of the diploid DP sweep (engine "v4"), shared verbatim by the device builder
// (dp_sweep4.cuh: prog_fill_kernel), the sweep kernel (dp_sweep4.cuh: dip_sweep4_kernel) and the CPU kernel-logic
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
// IN-PLACE LAYERS.  The reference copies every cell of every level (dp_cur -> dp_next, :565-576, :704-706), although
// most of a pangenome level is a relabelling: a lane or dummy vertex with one weight-0 in-edge, no colours on either
// end.  Here a vertex owns a SLOT (row and column index of the layer tile) instead of its level-local position, and such
// a "passive" vertex inherits the slot of its predecessor: the cells of passive x passive pairs are the same memory
// in level l and l+1, their value is unchanged (delta = 0, no layer shift) — they are not touched and appear in no
// program.  Every other vertex of level l+1 gets a FRESH slot (one no vertex of level l holds), so every written cell
// has a fresh row or column and never aliases a cell that is read in the same transition: one tile, no ping-pong, no
// hazards.  Slots of level-l vertices that nobody inherits are recycled after the transition.  Tie-breaking is
// unaffected: the candidates of a cell are still enumerated in ascending source POSITION (i, then j).
// A level lives in the shared-memory tile of CTA 0 (at most `kn` slots) or in the HBM tile; a transition that changes
// the placement RELOCATES the level: all its cells are written (compact slots 0..k-1), passive ones included.
//
// A destination vertex is S1 (one in-edge), M (two or more) or Z (none); S1 splits into passive (S1p) and the rest
// (S1a).  The cells of level l+1 are
//   * copy cells   S1 x S1 minus S1p x S1p   one candidate: dst = src shifted by w, plus delta; NO predecessor code is
//                                     stored (the code is implied: ordinal 0);
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
constexpr uint32_t PROG_ALIGN = 64;                       // alignment of a transition's program in the program buffer (ProgDir::off64)
constexpr uint32_t PROG_BIG_MIN = 12;                     // multi cells with this many candidates go to the warp form
constexpr uint32_t PROG_KEY_CAND = 1024;                  // candidates per cell whose ordinal fits the packed key (code = low 16 key bits)
constexpr uint32_t PROG_MAX_CAND = 32768;                 // most candidates of a cell: above PROG_KEY_CAND the warp form keeps the ROUND
                                                          // (ordinal / 32) in the key, the winner's lane comes from a ballot, and the code
                                                          // is the plain ordinal (panels of up to 181 walks: a recombination x recombination
                                                          // cell has in-degree^2 candidates)
constexpr uint32_t PROG_COMPACT_DELTA = 1023;             // compact descriptors: at most this many colours on the two levels

enum : uint32_t {
    PF_COMPACT = 1,      // compact descriptors (10-bit cell indices: both levels in the shared-memory tile)
    PF_SRC_SMEM = 2,     // level l lives in CTA 0's shared-memory tile
    PF_DST_SMEM = 4,     // level l+1 lives in CTA 0's shared-memory tile
    PF_STAGED = 8,       // the whole program travels into the ring slot (else only dir entry + header)
    PF_WAIT = 16,        // wait for dir.wait_target arrivals before the level
    PF_ARRIVE = 32,      // arrive after the level
    PF_ALL_CTAS = 64,    // every CTA of the problem takes part (both levels in HBM/L2)
    PF_RELOCATE = 128,   // the level changes tiles: every cell is written (else in place: passive x passive cells are not)
};

// Directory entry of a transition (32 bytes, copied into the slot in front of the program).
struct ProgDir {
    uint32_t off64;        // program offset in the program buffer, units of 64 bytes (programs start 64-byte aligned: 256 GB)
    uint32_t stage_bytes;  // bytes the producer copies after the entry: the whole program or just the header
    uint32_t wait_target;
    uint32_t flags;
    int32_t level;         // transition level -> level + 1 (the timed directory skips the transitions with nothing to do)
    uint32_t prog16;       // size of the whole program, units of 16 bytes (the producer prefetches unstaged programs into L2)
    uint32_t rsv[2];
};
static_assert(sizeof(ProgDir) == 32, "ProgDir layout");

struct ProgHdr {           // 64 bytes
    uint16_t k, k2;
    uint32_t n_copy, n_multi, n_cand, n_big, n_dead;
    uint32_t max_n;        // most candidates of a thread-form multi cell (loop bound hint)
    uint32_t n_passive;    // S1p vertices of level l+1 (their pairs are not in the program)
    uint64_t pred_off;     // u16 elements: codes of level l+1 start here, layout [layer][slot]
    uint32_t n_mm;         // M x M cells (the last n_mm multi cells): the only ones that can exceed PROG_KEY_CAND candidates
    uint32_t n_giant;      // how many of them have more than PROG_KEY_CAND candidates (dp_sweep4.cuh: giant_cells)
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

// ---- descriptors (cell indices are SLOT based: slot(i) * stride + slot(j)) ------------------------------------------
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

// ---- per-level class tables (built once per problem, dp_plan4.cpp) -----------------------------------------------
// For the vertices of one level, positions in class order S1a | S1p | M | Z (each class in position order), and for the M
// vertices the running sum of their in-degrees.
struct LevelClass {
    uint32_t k2, n1, np, m, z, dm;   // n1 = all S1 vertices, np of them passive; dm = sum of in-degrees over M
    const uint16_t* list;            // [k2]  S1a | S1p | M | Z positions
    const uint32_t* mpre;            // [m+1] prefix of in-degrees over M in rank order
};

struct ProgCounts { uint64_t n_copy, n_multi, n_cand, n_big, n_dead; uint32_t max_n; };

// `relocate`: passive vertices are written like the others (np counts as 0).
DG_HD ProgCounts prog_counts(const LevelClass& c, bool relocate) {
    ProgCounts o;
    const uint64_t n1 = c.n1, np = relocate ? 0 : c.np, m = c.m, dm = c.dm, k2 = c.k2, live = n1 + m;
    o.n_copy = n1 * n1 - np * np;
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
    const uint16_t* slot_src;   // [k]  slot of every vertex of level l (by position)
    const uint16_t* slot_dst;   // [k2] ... of level l+1
    uint32_t stride_src, stride_dst;   // slots per row of the tile level l / l+1 lives in
    bool relocate;
    int32_t W;                  // mask words (0: no colours on the two levels)
    const uint64_t* msrc;
    const uint64_t* mdst;
};

DG_HD void in_edge_at(const ProgLevelIn& L, uint32_t pos, uint32_t e, uint32_t& src_pos, uint32_t& w) {
    const uint32_t x = L.in_edge[L.in_off[pos] + (int32_t)e];
    src_pos = x & 0xFFFFu; w = x >> 16;
}
DG_HD uint32_t src_cell(const ProgLevelIn& L, uint32_t i, uint32_t j) { return (uint32_t)L.slot_src[i] * L.stride_src + L.slot_src[j]; }
DG_HD uint32_t dst_cell(const ProgLevelIn& L, uint32_t i2, uint32_t j2) { return (uint32_t)L.slot_dst[i2] * L.stride_dst + L.slot_dst[j2]; }

// copy cell t in [0, n1^2 - np^2): the rows of the S1a vertices in full (all S1 columns), then the S1a columns of the S1p rows
DG_HD void copy_pair(const ProgLevelIn& L, uint64_t t, uint32_t& i2, uint32_t& j2) {
    const uint64_t n1 = L.cls.n1, np = L.relocate ? 0 : L.cls.np, na = n1 - np;
    uint32_t a, b;
    if (t < na * n1) { a = (uint32_t)(t / n1); b = (uint32_t)(t - (uint64_t)a * n1); }
    else { const uint64_t u = t - na * n1; const uint32_t r = (uint32_t)(u / na); a = (uint32_t)na + r; b = (uint32_t)(u - (uint64_t)r * na); }
    i2 = L.cls.list[a]; j2 = L.cls.list[b];
}
DG_HD CopyDesc make_copy(const ProgLevelIn& L, uint64_t t) {
    uint32_t i2, j2;
    copy_pair(L, t, i2, j2);
    uint32_t i, wi, j, wj;
    in_edge_at(L, i2, 0, i, wi);
    in_edge_at(L, j2, 0, j, wj);
    CopyDesc d;
    d.src = src_cell(L, i, j); d.dst = dst_cell(L, i2, j2); d.w = wi + wj;
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
    d.src = src_cell(L, i, j); d.w = wi + wj;
    d.delta = L.W ? (uint32_t)mask_delta(L.W, L.msrc, L.mdst, (int)i, (int)j, (int)c.i2, (int)c.j2) : 0u;
    return d;
}

// dead cell x in [0, k2^2 - (n1+m)^2): rows of Z vertices in full, then the Z columns of the other rows
DG_HD uint32_t dead_cell(const ProgLevelIn& L, uint64_t x) {
    const uint64_t k2 = L.k2, z = L.cls.z, live = L.cls.n1 + L.cls.m;
    const uint16_t* Z = L.cls.list + live;
    if (x < z * k2) {
        const uint32_t a = (uint32_t)(x / k2), j2 = (uint32_t)(x - (uint64_t)a * k2);
        return dst_cell(L, Z[a], j2);
    }
    const uint64_t u = x - z * k2;
    const uint32_t a = (uint32_t)(u / z), b = (uint32_t)(u - (uint64_t)a * z);
    return dst_cell(L, L.cls.list[a], Z[b]);      // list[0 .. live) = the S1 and M positions
}

// Passive pair x in [0, np^2) of an in-place transition (not in the program: the cell is the same memory in both
// levels): destination positions; the sources are the single in-edges' (checksum variant and emulator only).
DG_HD void passive_pair(const ProgLevelIn& L, uint64_t x, uint32_t& i2, uint32_t& j2) {
    const uint32_t np = L.cls.np, na = L.cls.n1 - np;
    const uint32_t a = (uint32_t)(x / np), b = (uint32_t)(x - (uint64_t)a * np);
    i2 = L.cls.list[na + a]; j2 = L.cls.list[na + b];
}

// Kernel geometry the directory is made for.
struct Sweep4Shape {
    int slog = 10;             // shared-memory layer stride = 1 << slog cells ...
    int stride = 0;            // ... or, when non-zero, this many cells (the kernel is instantiated for 1024, 680, 512, 256)
    int kn = 32;               // slots of the shared-memory tile (kn * kn <= stride)
    int cells() const { return stride ? stride : 1 << slog; }
    int slot_bytes = 8192;     // ring slot (directory entry + program)
    int nslot = 4;             // ring depth
    int grid = 1;              // CTAs of the problem
};

}  // namespace dg
