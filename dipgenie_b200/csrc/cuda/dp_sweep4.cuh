// dp_sweep4.cuh — device side of the level-program engine (v4) of the diploid DP: the program builder
// (prog_fill_kernel) and the persistent sweep kernel (dip_sweep4_kernel / dip_sweep4_many_kernel).  Included by
// dp_diploid.cu only (it uses that file's PTX helpers).  Format and semantics: dp_prog.h; host planning: dp_plan4.cpp.
//
// Replaces the relax loop of Approximator::diploid_dp_approximation_solver (reference src/approximator.cpp:627-701).
//
// Sweep kernel, per problem:
//   * CTA 0 keeps the two DP layers of every level at most `kn` wide in shared-memory tiles of FIXED layer stride
//     (ST cells: 1024, 680, 512 or 256): layer r of a cell is one LDS/STS with an immediate offset, no address arithmetic per layer;
//     wider levels live in HBM/L2 tiles and their transitions are shared by all CTAs of the problem (monotone
//     counter barrier); hand-over transitions between the two placements run on CTA 0;
//   * a producer warp streams every transition's directory entry and program into a ring of shared-memory slots
//     with TMA bulk copies (cp.async.bulk + mbarrier full/empty pairs), several levels ahead;
//   * the compute warps share a transition's work units round-robin: big cells (one warp per cell, lanes over
//     candidates, REDUX.MAX per layer), blocks of 32 multi cells (one thread per cell, candidates in a loop), blocks
//     of 32 copy cells (one thread per cell) and blocks of dead cells; every unit handles all layers, RC at a time
//     in registers;
//   * one named barrier per level; predecessor codes (u16 per layer and multi cell) are the only HBM stream.
#pragma once

#include "dg_common.cuh"
#include "dp_cell.h"
#include "dp_prog.h"

namespace dg {

// ---- builder -----------------------------------------------------------------------------------------------------
struct Fill4Args {
    int32_t l0, l1;                      // transitions [l0, l1) are written (a window of the program)
    const int32_t* level_off;
    const int32_t* in_off;
    const uint32_t* in_edge;
    const uint16_t* cls_list;
    const uint32_t* mpre;
    const int64_t* mpre_off;
    const uint32_t* lvl_n1; const uint32_t* lvl_np; const uint32_t* lvl_m; const uint32_t* lvl_z; const uint32_t* lvl_dm;
    const uint16_t* vslot;               // [V] slot of every vertex in its level's tile
    const uint8_t* lvl_dom;              // [L] 0 shared-memory tile, 1 HBM tile
    const uint32_t* tflags;              // [L-1] PF_COMPACT / PF_RELOCATE ... of the transition
    int32_t kn, hstride;                 // slots per row of the two tiles
    const int32_t* lvlW;
    const int64_t* msrc_off; const int64_t* mdst_off;
    const uint64_t* masks;
    const ProgHdr* hdr;
    const uint64_t* prog_off;            // byte offsets, absolute
    uint64_t prog_base;                  // byte offset of `prog` within the whole program (windowed building)
    uint8_t* prog;
};

// The dp_prog.h view of transition l from the device tables (builder and checksum variant).
template <class A>
__device__ __forceinline__ ProgLevelIn level_in(const A& a, int l) {
    ProgLevelIn in;
    const int32_t lo = a.level_off[l], mid = a.level_off[l + 1];
    in.k = (uint32_t)(mid - lo); in.k2 = (uint32_t)(a.level_off[l + 2] - mid);
    in.in_off = a.in_off + mid;
    in.in_edge = a.in_edge;
    in.cls.k2 = in.k2; in.cls.n1 = a.lvl_n1[l + 1]; in.cls.np = a.lvl_np[l + 1]; in.cls.m = a.lvl_m[l + 1]; in.cls.z = a.lvl_z[l + 1];
    in.cls.dm = a.lvl_dm[l + 1];
    in.cls.list = a.cls_list + mid;
    in.cls.mpre = a.mpre + a.mpre_off[l + 1];
    in.slot_src = a.vslot + lo; in.slot_dst = a.vslot + mid;
    in.stride_src = a.lvl_dom[l] == 0 ? (uint32_t)a.kn : (uint32_t)a.hstride;
    in.stride_dst = a.lvl_dom[l + 1] == 0 ? (uint32_t)a.kn : (uint32_t)a.hstride;
    in.relocate = (a.tflags[l] & PF_RELOCATE) != 0;
    in.W = a.lvlW[l];
    in.msrc = in.W ? a.masks + a.msrc_off[l] : nullptr;
    in.mdst = in.W ? a.masks + a.mdst_off[l] : nullptr;
    return in;
}

constexpr int FILL4_THREADS = 256;

// One CTA per transition (grid-stride).  The buffer is zeroed beforehand (section padding).
__global__ void __launch_bounds__(FILL4_THREADS) prog_fill_kernel(const Fill4Args a) {
    __shared__ uint32_t warp_cnt[FILL4_THREADS / 32];
    __shared__ uint32_t big_base;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int l = a.l0 + (int)blockIdx.x; l < a.l1; l += (int)gridDim.x) {
        const ProgHdr h = a.hdr[l];
        const bool compact = (a.tflags[l] & PF_COMPACT) != 0;
        const ProgLayout lay = prog_layout(compact, h.n_copy, h.n_multi, h.n_cand, h.n_big, h.n_dead);
        uint8_t* const out = a.prog + (a.prog_off[l] - a.prog_base);
        const ProgLevelIn in = level_in(a, l);
        if (tid < (int)(sizeof(ProgHdr) / 4)) reinterpret_cast<uint32_t*>(out)[tid] = reinterpret_cast<const uint32_t*>(&a.hdr[l])[tid];
        if (tid == 0) big_base = 0;
        __syncthreads();
        for (uint32_t t = (uint32_t)tid; t < h.n_copy; t += FILL4_THREADS) {
            const CopyDesc d = make_copy(in, t);
            if (compact) reinterpret_cast<uint32_t*>(out + lay.copy)[t] = pack_copy_c(d);
            else reinterpret_cast<uint4*>(out + lay.copy)[t] = make_uint4(d.src | (d.w << 30), d.dst, d.delta, 0u);
        }
        // multi cells in slot order, FILL4_THREADS at a time (the big list is an ordered compaction)
        for (uint32_t t0 = 0; t0 < h.n_multi; t0 += FILL4_THREADS) {
            const uint32_t t = t0 + (uint32_t)tid;
            bool is_big = false;
            if (t < h.n_multi) {
                const MultiCell c = multi_cell(in, t);
                const uint32_t dst = dst_cell(in, c.i2, c.j2);
                if (compact) reinterpret_cast<uint2*>(out + lay.cell)[t] = make_uint2(dst | (c.n << 16), (uint32_t)c.cand_off);
                else reinterpret_cast<uint4*>(out + lay.cell)[t] = make_uint4(dst, c.n, (uint32_t)c.cand_off, 0u);
                for (uint32_t o = 0; o < c.n; ++o) {
                    const CandDesc d = make_cand(in, c, o);
                    if (compact) reinterpret_cast<uint32_t*>(out + lay.cand)[c.cand_off + o] = pack_cand_c(d);
                    else reinterpret_cast<uint2*>(out + lay.cand)[c.cand_off + o] = make_uint2(d.src | (d.w << 30), d.delta);
                }
                is_big = c.n >= PROG_BIG_MIN;
            }
            if (h.n_big) {                                         // (uniform over the CTA)
                const uint32_t bal = __ballot_sync(0xFFFFFFFFu, is_big);
                if (lane == 0) warp_cnt[warp] = (uint32_t)__popc(bal);
                __syncthreads();
                uint32_t before = big_base;
                for (int w = 0; w < warp; ++w) before += warp_cnt[w];
                if (is_big) reinterpret_cast<uint32_t*>(out + lay.big)[before + (uint32_t)__popc(bal & ((1u << lane) - 1u))] = t;
                __syncthreads();
                if (tid == 0) { uint32_t s = 0; for (int w = 0; w < FILL4_THREADS / 32; ++w) s += warp_cnt[w]; big_base += s; }
                __syncthreads();
            }
        }
        for (uint32_t x = (uint32_t)tid; x < h.n_dead; x += FILL4_THREADS) reinterpret_cast<uint32_t*>(out + lay.dead)[x] = dead_cell(in, x);
        __syncthreads();
    }
}

// vup[v] = 1 << 31 | in_edge of v when v has exactly one in-edge, else 0 (dp_cell.h: TraceView::vup).
__global__ void vup_fill_kernel(const int32_t* __restrict__ in_off, const uint32_t* __restrict__ in_edge, int32_t V, uint32_t* __restrict__ vup) {
    for (int32_t v = (int32_t)(blockIdx.x * blockDim.x + threadIdx.x); v < V; v += (int32_t)(gridDim.x * blockDim.x)) {
        const int32_t a0 = in_off[v];
        vup[v] = in_off[v + 1] - a0 == 1 ? (0x80000000u | in_edge[a0]) : 0u;
    }
}

// Fills `n` int32 cells with DEAD (padding layers of the HBM tiles).
__global__ void fill_dead_kernel(int32_t* p, long long n) {
    for (long long x = (long long)blockIdx.x * blockDim.x + threadIdx.x; x < n; x += (long long)gridDim.x * blockDim.x) p[x] = V4_DEAD;
}

// ---- sweep -------------------------------------------------------------------------------------------------------

struct Sweep4Args {
    const ProgDir* dir;           // [L-1]
    const int32_t* wide_list;     // transitions every CTA takes part in
    int32_t n_trans, n_wide;
    const uint8_t* prog;
    int32_t* gtile;               // HBM tile, CELL-major: hstride^2 cells of gcs words — two dead padding layers, then the RL layers
    long long gcs;                //   (in place: dp_prog.h); a chunk of layers of a cell is 1-2 sectors whichever way the block runs
    uint16_t* pred;
    unsigned int* counter;        // [0] level arrivals, [1] first barrier that timed out (level + 1), sticky
    unsigned long long* level_sum;
    unsigned long long* level_live;
    int32_t* sink;                // [R+1] layer values (still shifted) of cell (0,0) of the last level
    int32_t R, nchunk;            // nchunk * RC layers are computed
    int32_t grid, ncw;            // CTAs of the problem, compute warps per CTA
    int32_t slot_bytes, nslot;    // ring geometry
    uint32_t m_nchunk;            // magic of nchunk (dp_cell.h: make_magic; 0 when nchunk == 1)
    int32_t use_l1;               // diagnostics switch for Lvl4::l1
    int32_t last_smem;            // the last level lives in the shared-memory tile
    uint32_t sink_cell;           // its cell (0,0)
    const Fill4Args* chk;         // checksum variant: the device tables the positions of a descriptor are decoded from
    uint32_t final_target;        // arrivals once the last transition is complete
    unsigned long long* prof;     // diagnostics (nullable): cycles of CTA 0 / thread 0 in [slot wait, level work, barrier], levels
    unsigned long long timeout_ns;
    unsigned long long* giant_key; // [GIANT_LIST_MAX][nchunk * RC] slices of a giant cell meet here (zero between levels)
    unsigned int* giant_cnt;       // [GIANT_LIST_MAX] slices arrived
};

__device__ __forceinline__ void l2_prefetch_bulk(const void* p, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bar_named(int id, int threads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory"); }

template <int IMM> __device__ __forceinline__ int32_t lds_imm(uint32_t a) {
    int32_t v;
    asm volatile("ld.shared.s32 %0, [%1+%2];" : "=r"(v) : "r"(a), "n"(IMM));
    return v;
}
template <int IMM> __device__ __forceinline__ void sts_imm(uint32_t a, int32_t v) {
    asm volatile("st.shared.s32 [%0+%1], %2;" ::"r"(a), "n"(IMM), "r"(v) : "memory");
}
// The same without `volatile`: the compiler may schedule these among the arithmetic of neighbouring candidates (software
// pipelining).  Safe here because a unit reads cells of level l (old slots) and writes cells of fresh slots only, and the
// addresses of every level come from descriptors that are themselves loaded after the level's barrier (volatile).
template <int IMM> __device__ __forceinline__ int32_t lds_imm_free(uint32_t a) {
    int32_t v;
    asm("ld.shared.s32 %0, [%1+%2];" : "=r"(v) : "r"(a), "n"(IMM));
    return v;
}
template <int ST, int RC, int Q = 0>
__device__ __forceinline__ void lds_layers_free(uint32_t a, int32_t (&v)[RC]) {
    if constexpr (Q < RC) { v[Q] = lds_imm_free<(Q * ST * 4)>(a); lds_layers_free<ST, RC, Q + 1>(a, v); }
}
template <int ST, int RC, int Q = 0>
__device__ __forceinline__ void lds_layers(uint32_t a, int32_t (&v)[RC]) {
    if constexpr (Q < RC) { v[Q] = lds_imm<(Q * ST * 4)>(a); lds_layers<ST, RC, Q + 1>(a, v); }
}
template <int ST, int RC, int Q = 0>
__device__ __forceinline__ void sts_layers(uint32_t a, const int32_t (&v)[RC]) {
    if constexpr (Q < RC) { sts_imm<(Q * ST * 4)>(a, v[Q]); sts_layers<ST, RC, Q + 1>(a, v); }
}

// What a compute warp knows about the transition it is working on.
struct Lvl4 {
    const uint8_t* copy_p; const uint8_t* cell_p; const uint8_t* cand_p; const uint8_t* big_p; const uint8_t* dead_p;   // generic pointers (slot or HBM)
    uint32_t k, k2, n_copy, n_multi, n_big, n_dead, n_mm;      // n_mm: M x M cells to scan for giants (0: the level has none)
    uint32_t src32, dst32;                // shared-memory tile: address of padding layer -2
    int32_t* gt;                          // HBM tile (cell-major): layer r of cell c at gt[c * cs + 2 + r]
    long long cs;
    uint16_t* pl;                         // codes of level l+1, [layer][slot]
    int level, R, nchunk, rc;
    uint32_t m_nchunk;                    // magic of nchunk (0: one chunk)
    bool l1;                              // HBM-tile loads may be served by L1 (the problem runs on one SM)
    const ProgLevelIn* in;                // checksum variant only
    unsigned long long* prof;             // diagnostics: non-null in lane 0 of warp 0 of CTA 0 when the run is profiled
};

template <bool COMPACT> __device__ __forceinline__ CopyDesc ld_copy(const uint8_t* p, uint32_t t) {
    if (COMPACT) return unpack_copy_c(reinterpret_cast<const uint32_t*>(p)[t]);
    const uint4 x = reinterpret_cast<const uint4*>(p)[t];
    return {x.x & 0x3FFFFFFFu, x.y, x.x >> 30, x.z};
}
template <bool COMPACT> __device__ __forceinline__ CellDesc ld_cell(const uint8_t* p, uint32_t t) {
    if (COMPACT) { const uint2 x = reinterpret_cast<const uint2*>(p)[t]; return {x.x & 1023u, x.x >> 16, x.y}; }
    const uint4 x = reinterpret_cast<const uint4*>(p)[t];
    return {x.x, x.y, x.z};
}
template <bool COMPACT> __device__ __forceinline__ CandDesc ld_cand(const uint8_t* p, uint32_t x) {
    if (COMPACT) return unpack_cand_c(reinterpret_cast<const uint32_t*>(p)[x]);
    const uint2 y = reinterpret_cast<const uint2*>(p)[x];
    return {y.x & 0x3FFFFFFFu, y.x >> 30, y.y};
}

// The raw descriptor words (compact: x only) and their decoding: raw words keep an item's whole candidate list in registers.
template <bool COMPACT> __device__ __forceinline__ uint2 ld_cand_raw(const uint8_t* p, uint32_t x) {
    if (COMPACT) return make_uint2(reinterpret_cast<const uint32_t*>(p)[x], 0u);
    return reinterpret_cast<const uint2*>(p)[x];
}
template <bool COMPACT> __device__ __forceinline__ CandDesc cand_of(uint2 y) {
    if (COMPACT) return unpack_cand_c(y.x);
    return {y.x & 0x3FFFFFFFu, y.x >> 30, y.y};
}

// RC consecutive layers r0 .. r0+RC-1 of source cell `src`, read w layers lower (padding layers are DEAD).
template <int ST, int RC, bool SS>
__device__ __forceinline__ void load_layers(const Lvl4& c, int r0, uint32_t src, uint32_t w, int32_t (&v)[RC]) {
    if (SS) {
        lds_layers<ST, RC>(c.src32 + (((uint32_t)(r0 + 2) - w) * (uint32_t)ST + src) * 4u, v);
    } else {
        const int32_t* p = c.gt + ((long long)src * c.cs + (long long)(r0 + 2 - (int)w));
        if (c.l1) {              // one CTA per problem: the cells were written by this SM, its L1 is coherent with them
#pragma unroll
            for (int q = 0; q < RC; ++q) v[q] = p[q];
        } else {
#pragma unroll
            for (int q = 0; q < RC; ++q) v[q] = __ldcg(p + q);
        }
    }
}
template <int ST, int RC, bool DS>
__device__ __forceinline__ void store_layers(const Lvl4& c, int r0, uint32_t dst, const int32_t (&v)[RC]) {
    if (DS) {
        sts_layers<ST, RC>(c.dst32 + ((uint32_t)(r0 + 2) * (uint32_t)ST + dst) * 4u, v);
    } else {
        int32_t* p = c.gt + ((long long)dst * c.cs + (long long)(r0 + 2));
        if constexpr (RC % 2 == 0) {          // r0 and RC even, cells 32-byte aligned: 8-byte stores
            int2* p2 = reinterpret_cast<int2*>(p);
            if (c.l1) {
#pragma unroll
                for (int q = 0; q < RC; q += 2) p2[q >> 1] = make_int2(v[q], v[q + 1]);
            } else {
#pragma unroll
                for (int q = 0; q < RC; q += 2) __stcg(p2 + (q >> 1), make_int2(v[q], v[q + 1]));
            }
        } else if (c.l1) {
#pragma unroll
            for (int q = 0; q < RC; ++q) p[q] = v[q];
        } else {
#pragma unroll
            for (int q = 0; q < RC; ++q) __stcg(p + q, v[q]);
        }
    }
}

// Checksum variant: the fold of oracle/ref_hook.h over (flat POSITION index, value, source positions) of a live cell.
struct Fold4 { unsigned long long sum, live; };
__device__ __forceinline__ void fold_pos(Fold4& f, int R, uint32_t k2, int r, uint32_t i2, uint32_t j2, int32_t val, uint32_t i, uint32_t j) {
    if (r > R || val < 0) return;
    ++f.live;
    f.sum += cell_fold(((uint64_t)r * k2 + i2) * k2 + j2, val >> V4_SHIFT, (int)i, (int)j);
}
__device__ __forceinline__ void fold_copy(Fold4& f, const ProgLevelIn& in, int R, uint32_t t, int r0, const int32_t* v, int n) {
    uint32_t i2, j2, i, wi, j, wj;
    copy_pair(in, t, i2, j2);
    in_edge_at(in, i2, 0, i, wi); in_edge_at(in, j2, 0, j, wj);
    for (int q = 0; q < n; ++q) fold_pos(f, R, in.k2, r0 + q, i2, j2, v[q], i, j);
}
// (`ord` != nullptr: the ordinals themselves — cells of more than PROG_KEY_CAND candidates)
__device__ __forceinline__ void fold_multi(Fold4& f, const ProgLevelIn& in, int R, uint32_t t, int r0, const int32_t* key, int n, const uint32_t* ord = nullptr) {
    const MultiCell mc = multi_cell(in, t);
    for (int q = 0; q < n; ++q) {
        const int32_t val = (int32_t)((uint32_t)key[q] & ~V4_ORD_MASK);
        if (val < 0 || r0 + q > R) continue;
        const uint32_t o = ord ? ord[q] : V4_ORD_MASK - ((uint32_t)key[q] & V4_ORD_MASK);
        const uint32_t e1 = o / mc.d2, e2 = o - e1 * mc.d2;
        uint32_t i, wi, j, wj;
        in_edge_at(in, mc.i2, e1, i, wi); in_edge_at(in, mc.j2, e2, j, wj);
        fold_pos(f, R, in.k2, r0 + q, mc.i2, mc.j2, val, i, j);
    }
}

// The generic path works on ITEMS = (cell, chunk of RC layers): a warp takes 32 items, so a block of cells is spread over
// nchunk times as many lanes and nothing loops over the chunks — every global access of an item (descriptor, candidate
// descriptors, layers) is one link of a short chain of L2 round trips, and the chain is the critical path of the level.
struct Item { uint32_t t; int r0; bool in; };
__device__ __forceinline__ Item item_of(const Lvl4& c, uint32_t blk, int lane, uint32_t n_cells) {
    const uint32_t x = blk * 32u + (uint32_t)lane;
    const uint32_t t = c.m_nchunk ? __umulhi(x, c.m_nchunk) : x;
    return {t, (int)(x - t * (uint32_t)c.nchunk) * c.rc, t < n_cells};
}

// 32 copy items: dst = src shifted by w layers, plus delta.
template <int ST, int RC, bool SS, bool DS, bool COMPACT, bool CHECK>
__device__ __forceinline__ void copy_block(const Lvl4& c, uint32_t blk, int lane, Fold4& f) {
    const Item it = item_of(c, blk, lane, c.n_copy);
    if (!it.in) return;
    const CopyDesc d = ld_copy<COMPACT>(c.copy_p, it.t);
    const int32_t add = (int32_t)(d.delta << V4_SHIFT);
    int32_t v[RC];
    load_layers<ST, RC, SS>(c, it.r0, d.src, d.w, v);
#pragma unroll
    for (int q = 0; q < RC; ++q) v[q] += add;
    store_layers<ST, RC, DS>(c, it.r0, d.dst, v);
    if (CHECK) fold_copy(f, *c.in, c.R, it.t, it.r0, v, RC);
}

// 32 multi items (cells of PROG_BIG_MIN candidates or more belong to
// the warp form).
template <int ST, int RC, bool SS, bool DS, bool COMPACT, bool CHECK>
__device__ __forceinline__ void multi_block(const Lvl4& c, uint32_t blk, int lane, Fold4& f) {
    const Item it = item_of(c, blk, lane, c.n_multi);
    CellDesc cd = {0u, 0u, 0u};
    if (it.in) cd = ld_cell<COMPACT>(c.cell_p, it.t);
    const uint32_t n = cd.n >= PROG_BIG_MIN ? 0u : cd.n;
    const uint32_t nmax = __reduce_max_sync(0xFFFFFFFFu, n);
    if (nmax == 0) return;
    int32_t key[RC];
#pragma unroll
    for (int q = 0; q < RC; ++q) key[q] = V4_DEAD;
    // candidates two at a time: the two descriptors are fetched first, then their 2 x RC layers (more at a time, or all the
    // descriptors first, spills at 96 registers and is slower: 342 against 299 ms for 256 resident MHC_4 samples)
    for (uint32_t o = 0; o < nmax; o += 2u) {
        const bool p0 = o < n, p1 = o + 1u < n;
        CandDesc e0 = {0u, 0u, 0u}, e1 = {0u, 0u, 0u};
        if (p0) e0 = ld_cand<COMPACT>(c.cand_p, cd.cand_off + o);
        if (p1) e1 = ld_cand<COMPACT>(c.cand_p, cd.cand_off + o + 1u);
        int32_t v0[RC], v1[RC];
#pragma unroll
        for (int q = 0; q < RC; ++q) { v0[q] = V4_DEAD; v1[q] = V4_DEAD; }
        if (p0) load_layers<ST, RC, SS>(c, it.r0, e0.src, e0.w, v0);
        if (p1) load_layers<ST, RC, SS>(c, it.r0, e1.src, e1.w, v1);
        const int32_t a0 = (int32_t)((e0.delta << V4_SHIFT) + (V4_ORD_MASK - o)), a1 = (int32_t)((e1.delta << V4_SHIFT) + (V4_ORD_MASK - o - 1u));
#pragma unroll
        for (int q = 0; q < RC; ++q) key[q] = max(key[q], max(v0[q] + a0, v1[q] + a1));
    }
    if (n) {
        uint16_t* pl = c.pl + ((size_t)it.r0 * c.n_multi + it.t);
        int32_t val[RC];
#pragma unroll
        for (int q = 0; q < RC; ++q) {
            val[q] = (int32_t)((uint32_t)key[q] & ~V4_ORD_MASK);
            pl[(size_t)q * c.n_multi] = (uint16_t)key[q];
        }
        store_layers<ST, RC, DS>(c, it.r0, cd.dst, val);
        if (CHECK) fold_multi(f, *c.in, c.R, it.t, it.r0, key, RC);
    }
}

// One (big cell, chunk) per warp: lanes over candidates, 32 at a time, then one REDUX.MAX per layer.  Cells of more than
// PROG_KEY_CAND candidates belong to giant_cells below.
template <int ST, int RC, bool SS, bool DS, bool COMPACT, bool CHECK>
__device__ __forceinline__ void big_cell(const Lvl4& c, uint32_t x, int lane, Fold4& f) {
    const uint32_t u = c.m_nchunk ? __umulhi(x, c.m_nchunk) : x;
    const int r0 = (int)(x - u * (uint32_t)c.nchunk) * c.rc;
    const uint32_t t = reinterpret_cast<const uint32_t*>(c.big_p)[u];
    const CellDesc cd = ld_cell<COMPACT>(c.cell_p, t);
    if (cd.n > PROG_KEY_CAND) return;
    int32_t key[RC];
#pragma unroll
    for (int q = 0; q < RC; ++q) key[q] = V4_DEAD;
    // (the next candidate's descriptor is fetched before this one's layers: two independent chains of loads)
    CandDesc e = {0u, 0u, 0u};
    if ((uint32_t)lane < cd.n) e = ld_cand<COMPACT>(c.cand_p, cd.cand_off + (uint32_t)lane);
    for (uint32_t o = (uint32_t)lane; o < cd.n; o += 32u) {
        CandDesc en = {0u, 0u, 0u};
        if (o + 32u < cd.n) en = ld_cand<COMPACT>(c.cand_p, cd.cand_off + o + 32u);
        const int32_t add = (int32_t)((e.delta << V4_SHIFT) + (V4_ORD_MASK - o));
        int32_t v[RC];
        load_layers<ST, RC, SS>(c, r0, e.src, e.w, v);
#pragma unroll
        for (int q = 0; q < RC; ++q) key[q] = max(key[q], v[q] + add);
        e = en;
    }
#pragma unroll
    for (int q = 0; q < RC; ++q) key[q] = __reduce_max_sync(0xFFFFFFFFu, key[q]);
    if (lane == 0) {
        uint16_t* pl = c.pl + ((size_t)r0 * c.n_multi + t);
        int32_t val[RC];
#pragma unroll
        for (int q = 0; q < RC; ++q) {
            val[q] = (int32_t)((uint32_t)key[q] & ~V4_ORD_MASK);
            pl[(size_t)q * c.n_multi] = (uint16_t)key[q];
        }
        store_layers<ST, RC, DS>(c, r0, cd.dst, val);
        if (CHECK) fold_multi(f, *c.in, c.R, t, r0, key, RC);
    }
}

// GIANT cells — more than PROG_KEY_CAND candidates: a recombination x recombination cell of a panel of more than 32 walks has
// in-degree^2 of them (8100 at 90 walks), and one warp walking them 32 at a time would be the critical path of the level.
// A giant cell is cut into SLICES, one per CTA (transitions shared by all CTAs: cell j of ng goes to the CTAs c with
// c mod ng == j; otherwise one slice), and within a CTA all compute warps take the slice together: slice s, warp w, lane x
// visits the ordinals o = round * (ns * S) + s * S + w * 32 + x (S = 32 * ncw), keeping the ROUND in the key's ordinal field —
// a lane's rounds ascend, so its key is (best value, earliest round); the warp maximum plus the lowest lane holding it give
// the warp's first maximum; the warps' (value, ordinal) pairs meet in shared memory and the slices' in a 64-bit global word
// per layer (atomicMax of value : ~ordinal), which the last slice to arrive turns into the cell (and clears).  The smallest
// ordinal among the best values wins at every stage: the reference's first strict maximum in (e1,e2) order
// (approximator.cpp:657-659), for up to PROG_MAX_CAND candidates.  The code is the plain ordinal.  M x M cells are the last
// n_mm multi cells; every CTA derives the same ordered list of the giant ones.  Uniform control flow within a CTA, named
// barrier 2 over its compute warps.
constexpr int GIANT_LIST_MAX = 1024;
struct GiantShared {
    int2 part[16][10];              // [warp][layer of the chunk]: (value, ordinal)
    uint32_t list[GIANT_LIST_MAX];  // M x M indices of the giant cells, ascending
    uint32_t wcnt[16];
    uint32_t last;
};
__device__ __forceinline__ unsigned long long giant_pack(int32_t value, uint32_t ord) {
    return ((unsigned long long)((uint32_t)value ^ 0x80000000u) << 32) | (unsigned long long)(0xFFFFFFFFu - ord);
}

template <int ST, int RC, bool SS, bool DS, bool COMPACT, bool CHECK>
__device__ __forceinline__ void giant_cells(const Lvl4& c, GiantShared& gs, uint32_t cta, uint32_t n_ctas, int warp, int ncw, int lane, Fold4& f,
                                            unsigned long long* gkey, unsigned int* gcnt, int RL) {
    static_assert(RC <= 10, "GiantShared::part");
    const uint32_t S = 32u * (uint32_t)ncw, CT = S, tid = (uint32_t)warp * 32u + (uint32_t)lane;
    const uint32_t t_mm = c.n_multi - c.n_mm;
    // ---- the ordered list of giant cells ----
    uint32_t ng = 0;
    for (uint32_t x0 = 0; x0 < c.n_mm; x0 += CT) {
        const uint32_t x = x0 + tid;
        const bool is = x < c.n_mm && ld_cell<COMPACT>(c.cell_p, t_mm + x).n > PROG_KEY_CAND;
        const uint32_t bal = __ballot_sync(0xFFFFFFFFu, is);
        if (lane == 0) gs.wcnt[warp] = (uint32_t)__popc(bal);
        bar_named(2, (int)CT);
        uint32_t before = ng, all = 0;
        for (int w = 0; w < ncw; ++w) { const uint32_t n = gs.wcnt[w]; if (w < warp) before += n; all += n; }
        if (is) { const uint32_t idx = before + (uint32_t)__popc(bal & ((1u << lane) - 1u)); if (idx < (uint32_t)GIANT_LIST_MAX) gs.list[idx] = x; }
        ng += all;
        bar_named(2, (int)CT);
    }
    if (ng > (uint32_t)GIANT_LIST_MAX) ng = (uint32_t)GIANT_LIST_MAX;       // (the planner refuses levels with more: dp_plan4.cpp)
    if (ng == 0) return;
    const bool split = n_ctas >= ng && n_ctas > 1;
    for (uint32_t j = split ? cta % ng : cta; j < ng; j += split ? ng : n_ctas) {
        const uint32_t slice = split ? cta / ng : 0u, ns = split ? (n_ctas - j + ng - 1u) / ng : 1u;
        const uint32_t t = t_mm + gs.list[j];
        const CellDesc cd = ld_cell<COMPACT>(c.cell_p, t);
        const uint32_t step = ns * S, first = slice * S + tid;
        for (int ch = 0; ch < c.nchunk; ++ch) {
            const int r0 = ch * RC;
            int32_t key[RC];
#pragma unroll
            for (int q = 0; q < RC; ++q) key[q] = V4_DEAD;
            // two rounds per iteration, the descriptors of the next two fetched before this pair's layers
            const CandDesc none = {0u, 0u, 0u};
            CandDesc e0 = first < cd.n ? ld_cand<COMPACT>(c.cand_p, cd.cand_off + first) : none;
            CandDesc e1 = first + step < cd.n ? ld_cand<COMPACT>(c.cand_p, cd.cand_off + first + step) : none;
            uint32_t round = 0;
            for (uint32_t o = first; o < cd.n; o += 2u * step, round += 2u) {
                const CandDesc n0 = o + 2u * step < cd.n ? ld_cand<COMPACT>(c.cand_p, cd.cand_off + o + 2u * step) : none;
                const CandDesc n1 = o + 3u * step < cd.n ? ld_cand<COMPACT>(c.cand_p, cd.cand_off + o + 3u * step) : none;
                const bool p1 = o + step < cd.n;
                int32_t v0[RC], v1[RC];
#pragma unroll
                for (int q = 0; q < RC; ++q) v1[q] = V4_DEAD;
                load_layers<ST, RC, SS>(c, r0, e0.src, e0.w, v0);
                if (p1) load_layers<ST, RC, SS>(c, r0, e1.src, e1.w, v1);
                const int32_t a0 = (int32_t)((e0.delta << V4_SHIFT) + (V4_ORD_MASK - round)), a1 = (int32_t)((e1.delta << V4_SHIFT) + (V4_ORD_MASK - round - 1u));
#pragma unroll
                for (int q = 0; q < RC; ++q) key[q] = max(key[q], max(v0[q] + a0, v1[q] + a1));
                e0 = n0; e1 = n1;
            }
#pragma unroll
            for (int q = 0; q < RC; ++q) {
                const int32_t m = __reduce_max_sync(0xFFFFFFFFu, key[q]);
                const uint32_t who = __ballot_sync(0xFFFFFFFFu, key[q] == m);
                if (lane == 0) {
                    // (a warp without candidates reports ordinal ~0 with a value no candidate can fall below)
                    const uint32_t ord = (uint32_t)warp * 32u + slice * S < cd.n
                                             ? (V4_ORD_MASK - ((uint32_t)m & V4_ORD_MASK)) * step + slice * S + (uint32_t)warp * 32u + (uint32_t)(__ffs((int)who) - 1)
                                             : 0xFFFFFFFEu;
                    gs.part[warp][q] = make_int2((int32_t)((uint32_t)m & ~V4_ORD_MASK), (int)ord);
                }
            }
            bar_named(2, (int)CT);
            if (warp == 0 && lane < RC) {
                int2 best = gs.part[0][lane];
                for (int w = 1; w < ncw; ++w) {
                    const int2 y = gs.part[w][lane];
                    if (y.x > best.x || (y.x == best.x && (uint32_t)y.y < (uint32_t)best.y)) best = y;
                }
                if (ns > 1u) {
                    if ((uint32_t)best.y != 0xFFFFFFFEu) atomicMax(gkey + (size_t)j * (size_t)RL + (size_t)(r0 + lane), giant_pack(best.x, (uint32_t)best.y));
                    __threadfence();
                } else {
                    c.pl[(size_t)(r0 + lane) * c.n_multi + t] = (uint16_t)best.y;
                    if (DS) sts_s32(c.dst32 + ((uint32_t)(r0 + lane + 2) * (uint32_t)ST + cd.dst) * 4u, best.x);
                    else if (c.l1) c.gt[(long long)cd.dst * c.cs + (long long)(r0 + lane + 2)] = best.x;
                    else __stcg(c.gt + ((long long)cd.dst * c.cs + (long long)(r0 + lane + 2)), best.x);
                    if (CHECK) {
                        const int32_t k1 = best.x;
                        const uint32_t o1 = (uint32_t)best.y;
                        fold_multi(f, *c.in, c.R, t, r0 + lane, &k1, 1, &o1);
                    }
                }
            }
            bar_named(2, (int)CT);
        }
        if (ns > 1u) {
            // the last slice to arrive owns the cell: every slice's maxima are in the global words by then
            if (tid == 0) {
                __threadfence();
                gs.last = atomicAdd(gcnt + j, 1u) == ns - 1u ? 1u : 0u;
                __threadfence();
            }
            bar_named(2, (int)CT);
            if (gs.last) {
                for (int r = (int)tid; r < c.nchunk * RC; r += (int)CT) {
                    const unsigned long long k = atomicExch(gkey + (size_t)j * (size_t)RL + (size_t)r, 0ull);
                    const int32_t val = (int32_t)((uint32_t)(k >> 32) ^ 0x80000000u);
                    const uint32_t ord = 0xFFFFFFFFu - (uint32_t)k;
                    c.pl[(size_t)r * c.n_multi + t] = (uint16_t)ord;
                    if (DS) sts_s32(c.dst32 + ((uint32_t)(r + 2) * (uint32_t)ST + cd.dst) * 4u, val);
                    else __stcg(c.gt + ((long long)cd.dst * c.cs + (long long)(r + 2)), val);
                    if (CHECK) fold_multi(f, *c.in, c.R, t, r, &val, 1, &ord);
                }
                if (tid == 0) gcnt[j] = 0u;
            }
            bar_named(2, (int)CT);
        }
    }
}

template <int ST, int RC, bool DS>
__device__ __forceinline__ void dead_block(const Lvl4& c, uint32_t blk, int lane) {
    const Item it = item_of(c, blk, lane, c.n_dead);
    if (!it.in) return;
    const uint32_t dst = reinterpret_cast<const uint32_t*>(c.dead_p)[it.t];
    int32_t v[RC];
#pragma unroll
    for (int q = 0; q < RC; ++q) v[q] = V4_DEAD;
    store_layers<ST, RC, DS>(c, it.r0, dst, v);
}

// The work units of one transition, dealt round-robin to the (global) warps gw, gw + gstride, ...: heaviest first.
template <int ST, int RC, bool SS, bool DS, bool COMPACT, bool CHECK>
__device__ __forceinline__ void run_level(const Lvl4& c, uint32_t gw, uint32_t gstride, int lane, Fold4& f) {
    const uint32_t nch = (uint32_t)c.nchunk;
    const uint32_t nmb = (c.n_multi * nch + 31u) >> 5, ncb = (c.n_copy * nch + 31u) >> 5, ndb = (c.n_dead * nch + 31u) >> 5;
    const uint32_t e0 = c.n_big * nch, e1 = e0 + nmb, e2 = e1 + ncb, e3 = e2 + ndb;
    for (uint32_t u = gw; u < e3; u += gstride) {
        const long long t0 = c.prof ? clock64() : 0;
        if (u < e0) big_cell<ST, RC, SS, DS, COMPACT, CHECK>(c, u, lane, f);
        else if (u < e1) multi_block<ST, RC, SS, DS, COMPACT, CHECK>(c, u - e0, lane, f);
        else if (u < e2) copy_block<ST, RC, SS, DS, COMPACT, CHECK>(c, u - e1, lane, f);
        else dead_block<ST, RC, DS>(c, u - e2, lane);
        if (c.prof) c.prof[u < e0 ? 12 : (u < e1 ? 13 : (u < e2 ? 14 : 15))] += (unsigned long long)(clock64() - t0);
    }
}

// The giant cells of a transition (header: n_giant > 0; never a compact one, dp_plan4.cpp), before its other units.  A function
// of its own, with its own view of the transition: what it needs in registers (and spills) stays out of the paths every
// level takes — inlined into run_generic it cost the 256-sample MHC_4 sweep 299 -> 380 ms without ever running there.
template <int ST, int RC, bool CHECK>
__device__ __noinline__ ulonglong2 run_giants(const Sweep4Args& a, const uint8_t* sb, uint32_t tile32, int l, int cta, int warp, int lane) {
    const ProgDir d = *reinterpret_cast<const ProgDir*>(sb);
    const ProgHdr h = *reinterpret_cast<const ProgHdr*>(sb + sizeof(ProgDir));
    const uint32_t flags = d.flags;
    Lvl4 c;
    c.k = h.k; c.k2 = h.k2; c.n_copy = h.n_copy; c.n_multi = h.n_multi; c.n_big = h.n_big; c.n_dead = h.n_dead; c.n_mm = h.n_mm;
    const uint8_t* const pb = (flags & PF_STAGED) ? sb + sizeof(ProgDir) : a.prog + (size_t)d.off64 * PROG_ALIGN;
    c.copy_p = pb + sizeof(ProgHdr); c.cell_p = pb + h.off_cell; c.cand_p = pb + h.off_cand; c.big_p = pb + h.off_big; c.dead_p = pb + h.off_dead;
    c.level = l; c.R = a.R; c.nchunk = a.nchunk; c.rc = RC; c.m_nchunk = a.m_nchunk; c.l1 = a.grid == 1 && a.use_l1;
    c.src32 = tile32; c.dst32 = tile32;
    c.gt = a.gtile; c.cs = a.gcs;
    c.pl = a.pred + h.pred_off;
    c.prof = nullptr;
    ProgLevelIn in;
    if (CHECK) in = level_in(*a.chk, l);
    c.in = &in;
    const bool all = (flags & PF_ALL_CTAS) != 0;
    const uint32_t gc = all ? (uint32_t)cta : 0u, gn = all ? (uint32_t)a.grid : 1u;
    const bool ss = (flags & PF_SRC_SMEM) != 0, ds = (flags & PF_DST_SMEM) != 0;
    __shared__ GiantShared giant_sh;
    Fold4 f = {0ull, 0ull};
    const int RL = a.nchunk * RC;
    if (ss && ds) giant_cells<ST, RC, true, true, false, CHECK>(c, giant_sh, gc, gn, warp, a.ncw, lane, f, a.giant_key, a.giant_cnt, RL);
    else if (ss) giant_cells<ST, RC, true, false, false, CHECK>(c, giant_sh, gc, gn, warp, a.ncw, lane, f, a.giant_key, a.giant_cnt, RL);
    else if (ds) giant_cells<ST, RC, false, true, false, CHECK>(c, giant_sh, gc, gn, warp, a.ncw, lane, f, a.giant_key, a.giant_cnt, RL);
    else giant_cells<ST, RC, false, false, false, CHECK>(c, giant_sh, gc, gn, warp, a.ncw, lane, f, a.giant_key, a.giant_cnt, RL);
    return make_ulonglong2(f.sum, f.live);
}

// Everything but the staged compact transitions (below): hand-overs between the placements, HBM-resident levels,
// wide-format programs, programs read in place.  Out of line, with a handful of scalar arguments: the narrow loop keeps
// its own register allocation.  `sb` = the ring slot (directory entry, then the header and, if staged, the program).
// GIANT: the kernel variant for problems that have giant cells (in-degrees above 32).  The others run a variant without that
// code: its mere presence — never executed — cost the 256-sample MHC_4 sweep 295 -> 344 ms (code layout, registers of the callers).
template <int ST, int RC, bool CHECK, bool GIANT>
__device__ __noinline__ ulonglong2 run_generic(const Sweep4Args& a, const uint8_t* sb, uint32_t tile32, int l, int cta, int warp, int lane) {
    Fold4 f = {0ull, 0ull};
    // giant cells first, before anything of this function is live in registers
    if constexpr (GIANT) if (reinterpret_cast<const ProgHdr*>(sb + sizeof(ProgDir))->n_giant) {
        const bool prof = a.prof != nullptr && cta == 0 && warp == 0 && lane == 0;
        const long long tg0 = prof ? clock64() : 0;
        const ulonglong2 fg = run_giants<ST, RC, CHECK>(a, sb, tile32, l, cta, warp, lane);
        f.sum = fg.x; f.live = fg.y;
        if (prof) a.prof[11] += (unsigned long long)(clock64() - tg0);
    }
    const ProgDir d = *reinterpret_cast<const ProgDir*>(sb);
    const ProgHdr h = *reinterpret_cast<const ProgHdr*>(sb + sizeof(ProgDir));
    const uint32_t flags = d.flags;
    Lvl4 c;
    c.k = h.k; c.k2 = h.k2; c.n_copy = h.n_copy; c.n_multi = h.n_multi; c.n_big = h.n_big; c.n_dead = h.n_dead; c.n_mm = h.n_giant ? h.n_mm : 0u;
    const uint8_t* const pb = (flags & PF_STAGED) ? sb + sizeof(ProgDir) : a.prog + (size_t)d.off64 * PROG_ALIGN;
    c.copy_p = pb + sizeof(ProgHdr); c.cell_p = pb + h.off_cell; c.cand_p = pb + h.off_cand; c.big_p = pb + h.off_big; c.dead_p = pb + h.off_dead;
    c.level = l; c.R = a.R; c.nchunk = a.nchunk; c.rc = RC; c.m_nchunk = a.m_nchunk; c.l1 = a.grid == 1 && a.use_l1;
    c.src32 = tile32; c.dst32 = tile32;
    c.gt = a.gtile; c.cs = a.gcs;
    c.pl = a.pred + h.pred_off;
    c.prof = (a.prof != nullptr && cta == 0 && warp == 0 && lane == 0) ? a.prof : nullptr;
    ProgLevelIn in;
    if (CHECK) in = level_in(*a.chk, l);
    c.in = &in;
    const bool all = (flags & PF_ALL_CTAS) != 0;
    const uint32_t gw = all ? (uint32_t)(cta * a.ncw + warp) : (uint32_t)warp;
    const uint32_t gs = all ? (uint32_t)(a.grid * a.ncw) : (uint32_t)a.ncw;
    const bool ss = (flags & PF_SRC_SMEM) != 0, ds = (flags & PF_DST_SMEM) != 0;
    if (flags & PF_COMPACT) run_level<ST, RC, true, true, true, CHECK>(c, gw, gs, lane, f);
    else if (ss && ds) run_level<ST, RC, true, true, false, CHECK>(c, gw, gs, lane, f);
    else if (ss) run_level<ST, RC, true, false, false, CHECK>(c, gw, gs, lane, f);
    else if (ds) run_level<ST, RC, false, true, false, CHECK>(c, gw, gs, lane, f);
    else run_level<ST, RC, false, false, false, CHECK>(c, gw, gs, lane, f);
    return make_ulonglong2(f.sum, f.live);
}

// ---- the narrow loop: compact program staged in the slot, both layers in shared memory -------------------------------
// Everything is a 32-bit shared-window address; a unit is (kind, block of 32 cells or one big cell, chunk of RC layers), so
// that the dozen units of a typical level spread over as many warps and the level's critical path is one chunk of one block.
struct Fast4 {
    uint32_t copy32, cell32, cand32, big32, dead32;
    uint32_t n_copy, n_multi, n_big, n_dead;
    uint32_t src32, dst32;        // tiles: address of padding layer -2
    uint16_t* pl;                 // codes of level l+1
    int R;
    const ProgLevelIn* in;        // checksum variant only
};

__device__ __forceinline__ uint2 lds_v2(uint32_t a) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a));
    return v;
}

template <int ST, int RC, bool CHECK>
__device__ __forceinline__ void fast_copy(const Fast4& c, uint32_t blk, int r0, int lane, Fold4& f) {
    const uint32_t t = blk * 32u + (uint32_t)lane;
    const bool active = t < c.n_copy;
    const uint32_t x = active ? lds_u32(c.copy32 + 4u * t) : 0u;
    const uint32_t src = x & 1023u, dst = (x >> 10) & 1023u, w = (x >> 20) & 3u;
    const int32_t add = (int32_t)((x >> 22) << V4_SHIFT);
    int32_t v[RC];
    lds_layers<ST, RC>(c.src32 + (((uint32_t)(r0 + 2) - w) * (uint32_t)ST + src) * 4u, v);
#pragma unroll
    for (int q = 0; q < RC; ++q) v[q] += add;
    if (active) {
        sts_layers<ST, RC>(c.dst32 + ((uint32_t)(r0 + 2) * (uint32_t)ST + dst) * 4u, v);
        if (CHECK) fold_copy(f, *c.in, c.R, t, r0, v, RC);
    }
}

// A candidate descriptor that changes nothing: the last cell of every layer is never a slot pair (kn^2 < ST) and is
// kept DEAD, so lanes that have run out of candidates walk on branch-free.
template <int ST> __device__ __forceinline__ constexpr uint32_t dead_cand() { return (uint32_t)ST - 1u; }

template <int ST, int RC, bool CHECK>
__device__ __forceinline__ void fast_multi(const Fast4& c, uint32_t blk, int r0, int lane, Fold4& f) {
    const uint32_t t = blk * 32u + (uint32_t)lane;
    uint2 cd = make_uint2(0u, 0u);
    if (t < c.n_multi) cd = lds_v2(c.cell32 + 8u * t);
    const uint32_t dst = cd.x & 1023u;
    uint32_t n = cd.x >> 16;
    if (n >= PROG_BIG_MIN) n = 0u;                                   // the warp form's
    const uint32_t nmax = __reduce_max_sync(0xFFFFFFFFu, n);
    if (nmax == 0u) return;
    const uint32_t ca = c.cand32 + 4u * cd.y;
    const uint32_t base = c.src32 + ((uint32_t)(r0 + 2) * (uint32_t)(ST * 4));
    int32_t key[RC];
#pragma unroll
    for (int q = 0; q < RC; ++q) key[q] = V4_DEAD;
    // two candidates per round, branch-free; the descriptors of the next round are fetched before this round's layers
    uint32_t x0 = lds_u32(ca), x1 = lds_u32(ca + 4u);
    for (uint32_t o = 0; o < nmax; o += 2u) {
        const uint32_t e0 = o < n ? x0 : dead_cand<ST>(), e1 = o + 1u < n ? x1 : dead_cand<ST>();
        x0 = lds_u32(ca + 4u * (o + 2u)); x1 = lds_u32(ca + 4u * (o + 3u));      // (may run past the cell's list: unused then)
        const int32_t add0 = (int32_t)(((e0 >> 12) << V4_SHIFT) + (V4_ORD_MASK - o));
        const int32_t add1 = (int32_t)(((e1 >> 12) << V4_SHIFT) + (V4_ORD_MASK - o - 1u));
        int32_t v0[RC], v1[RC];
        lds_layers_free<ST, RC>(base + ((e0 & 1023u) << 2) - (((e0 >> 10) & 3u) * (uint32_t)(ST * 4)), v0);
        lds_layers_free<ST, RC>(base + ((e1 & 1023u) << 2) - (((e1 >> 10) & 3u) * (uint32_t)(ST * 4)), v1);
#pragma unroll
        for (int q = 0; q < RC; ++q) key[q] = max(key[q], max(v0[q] + add0, v1[q] + add1));
    }
    if (n) {
        uint16_t* pl = c.pl + ((size_t)r0 * c.n_multi + t);
        int32_t val[RC];
#pragma unroll
        for (int q = 0; q < RC; ++q) {
            val[q] = (int32_t)((uint32_t)key[q] & ~V4_ORD_MASK);
            pl[(size_t)q * c.n_multi] = (uint16_t)key[q];
        }
        sts_layers<ST, RC>(c.dst32 + ((uint32_t)(r0 + 2) * (uint32_t)ST + dst) * 4u, val);
        if (CHECK) fold_multi(f, *c.in, c.R, t, r0, key, RC);
    }
}

template <int ST, int RC, bool CHECK>
__device__ __forceinline__ void fast_big(const Fast4& c, uint32_t u, int r0, int lane, Fold4& f) {
    const uint32_t t = lds_u32(c.big32 + 4u * u);
    const uint2 cd = lds_v2(c.cell32 + 8u * t);
    const uint32_t dst = cd.x & 1023u, n = cd.x >> 16;
    const uint32_t ca = c.cand32 + 4u * cd.y;
    const uint32_t base = c.src32 + ((uint32_t)(r0 + 2) * (uint32_t)(ST * 4));
    int32_t key[RC];
#pragma unroll
    for (int q = 0; q < RC; ++q) key[q] = V4_DEAD;
    for (uint32_t o = (uint32_t)lane; o < n; o += 32u) {
        const uint32_t e = lds_u32(ca + 4u * o);
        const int32_t add = (int32_t)(((e >> 12) << V4_SHIFT) + (V4_ORD_MASK - o));
        int32_t v[RC];
        lds_layers_free<ST, RC>(base + ((e & 1023u) << 2) - (((e >> 10) & 3u) * (uint32_t)(ST * 4)), v);
#pragma unroll
        for (int q = 0; q < RC; ++q) key[q] = max(key[q], v[q] + add);
    }
#pragma unroll
    for (int q = 0; q < RC; ++q) key[q] = __reduce_max_sync(0xFFFFFFFFu, key[q]);
    if (lane == 0) {
        uint16_t* pl = c.pl + ((size_t)r0 * c.n_multi + t);
        int32_t val[RC];
#pragma unroll
        for (int q = 0; q < RC; ++q) {
            val[q] = (int32_t)((uint32_t)key[q] & ~V4_ORD_MASK);
            pl[(size_t)q * c.n_multi] = (uint16_t)key[q];
        }
        sts_layers<ST, RC>(c.dst32 + ((uint32_t)(r0 + 2) * (uint32_t)ST + dst) * 4u, val);
        if (CHECK) fold_multi(f, *c.in, c.R, t, r0, key, RC);
    }
}

template <int ST, int RC>
__device__ __forceinline__ void fast_dead(const Fast4& c, uint32_t blk, int r0, int lane) {
    const uint32_t x = blk * 32u + (uint32_t)lane;
    if (x >= c.n_dead) return;
    const uint32_t dst = lds_u32(c.dead32 + 4u * x);
    int32_t v[RC];
#pragma unroll
    for (int q = 0; q < RC; ++q) v[q] = V4_DEAD;
    sts_layers<ST, RC>(c.dst32 + ((uint32_t)(r0 + 2) * (uint32_t)ST + dst) * 4u, v);
}

// Spin until the monotone arrival counter reaches `target`; gives up after timeout_ns or when another wait of the
// problem has already failed (counter[1] != 0: sticky), so that a lost CTA can never hang the GPU.
__device__ __noinline__ bool wait_counter4(unsigned int* counter, unsigned int target, unsigned long long timeout_ns, int level) {
    const unsigned long long t0 = global_ns();
    bool ok = true;
    for (unsigned int it = 0;; ++it) {
        unsigned int v;
        asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
        if (v >= target) break;
        __nanosleep(40);
        if ((it & 255u) == 255u) {
            unsigned int err;
            asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(err) : "l"(counter + 1) : "memory");
            if (err != 0u || global_ns() - t0 > timeout_ns) { atomicCAS(counter + 1, 0u, (unsigned int)level + 1u); ok = false; break; }
        }
    }
    asm volatile("fence.acq_rel.gpu;" ::: "memory");
    return ok;
}

template <int ST, int RC, bool CHECK, bool GIANT>
__device__ __forceinline__ void sweep4_body(const Sweep4Args& a, const int cta) {
    extern __shared__ __align__(128) uint8_t smem4[];
    const int RL = a.nchunk * RC;
    const uint32_t tile_bytes = (uint32_t)(RL + 2) * (uint32_t)(ST * 4);
    uint8_t* const slots = smem4;
    uint8_t* const tiles = slots + (size_t)a.nslot * a.slot_bytes;
    uint64_t* const full = reinterpret_cast<uint64_t*>(tiles + (size_t)tile_bytes);
    uint64_t* const empty = full + a.nslot;
    uint64_t* const failw = empty + a.nslot;               // set once a barrier of the problem has timed out
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ncw = a.ncw, CT = ncw * 32, NS = a.nslot;
    const int32_t n_my = cta == 0 ? a.n_trans : a.n_wide;

    if (tid == 0) {
        for (int s = 0; s < NS; ++s) { mbar_init(smem_u32(full + s), 1); mbar_init(smem_u32(empty + s), (uint32_t)ncw); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        *failw = 0ull;
    }
    if (cta == 0) {
        // two dead padding layers below layer 0 of the tile; level 0 (one vertex, slot 0): every layer starts at 0 (:535)
        int32_t* const t0 = reinterpret_cast<int32_t*>(tiles);
        for (int x = tid; x < 2 * ST; x += blockDim.x) t0[x] = V4_DEAD;
        for (int r = tid; r < RL; r += blockDim.x) {
            t0[(r + 2) * ST] = r <= a.R ? 0 : V4_DEAD;
            t0[(r + 2) * ST + (int)dead_cand<ST>()] = V4_DEAD;       // the cell no slot pair maps to (fast_multi)
        }
    }
    __syncthreads();

    if (tid >= CT) {
        // ---- producer warp: lanes fetch 32 directory entries at a time, lane 0 issues the copies ----
        int slot = 0;
        uint32_t use_parity = 0u;                           // parity of the previous use of the slot (first round: no wait)
        bool first_round = true;
        for (int32_t i0 = 0; i0 < n_my; i0 += 32) {
            int32_t lv = 0;
            uint4 f = make_uint4(0u, 0u, 0u, 0u);
            uint32_t p16 = 0u;
            if (i0 + lane < n_my) {
                lv = cta == 0 ? i0 + lane : __ldg(a.wide_list + i0 + lane);
                f = __ldg(reinterpret_cast<const uint4*>(a.dir + lv));
                p16 = __ldg(&a.dir[lv].prog16);
            }
            const int cnt = min(32, n_my - i0);
            for (int j = 0; j < cnt; ++j) {
                const int32_t l = __shfl_sync(0xFFFFFFFFu, lv, j);
                const uint32_t off64 = __shfl_sync(0xFFFFFFFFu, f.x, j), bytes = __shfl_sync(0xFFFFFFFFu, f.y, j);
                const uint32_t fl = __shfl_sync(0xFFFFFFFFu, f.w, j), all16 = __shfl_sync(0xFFFFFFFFu, p16, j);
                if (lane == 0) {
                    // a program too large for a slot is read in place by the compute warps: bring it into L2 now, NS levels
                    // ahead of its use (its first touch is otherwise a chain of cold DRAM reads inside the level)
                    // (a transition shared by all CTAs: CTA c brings in the c-th part)
                    if (!(fl & PF_STAGED)) {
                        const uint32_t parts = (fl & PF_ALL_CTAS) ? (uint32_t)a.grid : 1u;
                        const uint32_t per16 = (all16 + parts - 1u) / parts, lo16 = min(all16, per16 * (uint32_t)(parts > 1u ? cta : 0));
                        const uint32_t n16 = min(min(all16 - lo16, per16), 16384u);
                        if (n16) l2_prefetch_bulk(a.prog + (size_t)off64 * PROG_ALIGN + (size_t)lo16 * 16, n16 * 16u);
                    }
                    if (!first_round) mbar_wait(smem_u32(empty + slot), use_parity);
                    const uint32_t bar = smem_u32(full + slot);
                    const uint32_t dst = smem_u32(slots + (size_t)slot * a.slot_bytes);
                    mbar_expect_tx(bar, (uint32_t)sizeof(ProgDir) + bytes);
                    bulk_g2s(dst, a.dir + l, (uint32_t)sizeof(ProgDir), bar);
                    bulk_g2s(dst + (uint32_t)sizeof(ProgDir), a.prog + (size_t)off64 * PROG_ALIGN, bytes, bar);
                }
                if (++slot == NS) { slot = 0; if (first_round) first_round = false; else use_parity ^= 1u; }
            }
        }
        return;
    }

    // ---- compute warps ----
    const uint32_t slots32 = smem_u32(slots), tiles32 = smem_u32(tiles), full32 = smem_u32(full), empty32 = smem_u32(empty);
    const uint32_t fail32 = smem_u32(failw);
    bool failed = false;                    // a barrier of this problem timed out: stop working, keep the ring moving
    const bool profiling = a.prof != nullptr && cta == 0 && tid == 0;
    unsigned long long pc_wait = 0, pc_work = 0, pc_bar = 0;
    long long tk0 = 0, tk1 = 0, tk2 = 0;
    uint32_t slot = 0, parity = 0;
    for (int32_t i = 0; i < n_my; ++i) {
        if (profiling) tk0 = clock64();
        const uint32_t sb32 = slots32 + slot * (uint32_t)a.slot_bytes;
        mbar_wait(full32 + 8u * slot, parity);
        if (profiling) tk1 = clock64();
        const uint4 d = lds_v4(sb32);                       // directory entry
        const uint32_t flags = d.w;
        const int l = lds_s32(sb32 + 16u);                  // (the timed directory skips the transitions with nothing to do)
        if (a.grid > 1 && (flags & PF_WAIT) && !failed) {
            const long long tw0 = profiling ? clock64() : 0;
            if (tid == 0 && !wait_counter4(a.counter, d.z, a.timeout_ns, l)) sts_s32(fail32, 1);
            if (profiling) a.prof[10] += (unsigned long long)(clock64() - tw0);
            bar_named(1, CT);
            failed = lds_s32(fail32) != 0;
        }
        Fold4 f = {0ull, 0ull};
        constexpr uint32_t FAST = PF_COMPACT | PF_STAGED;
        if (!failed) {
            if ((flags & FAST) == FAST) {
                const uint4 h0 = lds_v4(sb32 + 32u), h1 = lds_v4(sb32 + 48u), h2 = lds_v4(sb32 + 64u), h3 = lds_v4(sb32 + 80u);
                Fast4 c;
                const uint32_t pb32 = sb32 + (uint32_t)sizeof(ProgDir);
                c.copy32 = pb32 + (uint32_t)sizeof(ProgHdr); c.cell32 = pb32 + h3.x; c.cand32 = pb32 + h3.y; c.big32 = pb32 + h3.z; c.dead32 = pb32 + h3.w;
                c.n_copy = h0.y; c.n_multi = h0.z; c.n_big = h1.x; c.n_dead = h1.y;
                c.src32 = tiles32; c.dst32 = tiles32;
                c.pl = a.pred + (((unsigned long long)h2.y << 32) | h2.x);
                c.R = a.R;
                ProgLevelIn in;
                if (CHECK) in = level_in(*a.chk, l);
                c.in = &in;
                const uint32_t nch = (uint32_t)a.nchunk;
                // (two lanes per multi cell — 16 cells per unit, candidates split by parity, SHFL.BFLY combine — was tried: the
                // shorter units do not pay for their doubled count: 301 against 289 ms for 256 resident MHC_4 samples)
                const uint32_t nmb = (c.n_multi + 31u) >> 5;
                const uint32_t e0 = c.n_big * nch, e1 = e0 + nmb * nch, e2 = e1 + ((c.n_copy + 31u) >> 5) * nch,
                               e3 = e2 + ((c.n_dead + 31u) >> 5) * nch;
                long long tu0 = 0;
                if (profiling) { tu0 = clock64(); a.prof[18] += (unsigned long long)(tu0 - tk1); }
                for (uint32_t u = (uint32_t)warp; u < e3; u += (uint32_t)ncw) {
                    if (profiling) { a.prof[20] += 1; a.prof[21 + (u < e0 ? 0 : (u < e1 ? 1 : (u < e2 ? 2 : 3)))] += 1; }
                    // unit -> (kind, block, chunk): chunks of one block are neighbours
                    const uint32_t lo = u < e0 ? 0u : (u < e1 ? e0 : (u < e2 ? e1 : e2));
                    const uint32_t v = u - lo, blk = a.m_nchunk ? __umulhi(v, a.m_nchunk) : v, ch = v - blk * nch;
                    const int r0 = (int)ch * RC;
                    if (u < e0) fast_big<ST, RC, CHECK>(c, blk, r0, lane, f);
                    else if (u < e1) fast_multi<ST, RC, CHECK>(c, blk, r0, lane, f);
                    else if (u < e2) fast_copy<ST, RC, CHECK>(c, blk, r0, lane, f);
                    else fast_dead<ST, RC>(c, blk, r0, lane);
                }
                if (profiling) a.prof[19] += (unsigned long long)(clock64() - tu0);
            } else {
                const ulonglong2 fs = run_generic<ST, RC, CHECK, GIANT>(a, slots + (size_t)slot * a.slot_bytes, tiles32, l, cta, warp, lane);
                f.sum = fs.x; f.live = fs.y;
            }
        }
        __syncwarp();
        if (profiling) tk2 = clock64();
        if (lane == 0) mbar_arrive(empty32 + 8u * slot);       // this warp is done with the slot
        bar_named(1, CT);                                       // the level is whole within this CTA
        if (a.grid > 1 && (flags & PF_ARRIVE) && tid == 0) red_release_add_u32(a.counter, 1u);
        if (profiling) {
            const long long tk3 = clock64();
            pc_wait += tk1 - tk0; pc_work += tk2 - tk1; pc_bar += tk3 - tk2;
            const int cls = (flags & FAST) == FAST ? 0 : ((flags & PF_ALL_CTAS) ? 2 : 1);      // narrow loop / hand-over and other / HBM to HBM
            a.prof[4 + 2 * cls] += 1; a.prof[5 + 2 * cls] += (unsigned long long)(tk3 - tk0);
        }
        if (CHECK && !failed && cta == 0) {
            // the passive pairs of level l+1: untouched memory, the same cells as in level l (dp_prog.h); every CTA-0
            // thread folds its share (HBM-resident cells of this level were completed before this transition's wait)
            const ProgLevelIn in = level_in(*a.chk, l);
            if (!in.relocate) {
                const uint64_t npp = (uint64_t)in.cls.np * in.cls.np;
                const bool dsm = (flags & PF_DST_SMEM) != 0;
                for (uint64_t x = (uint64_t)tid; x < npp; x += (uint64_t)CT) {
                    uint32_t i2, j2, i, wi, j, wj;
                    passive_pair(in, x, i2, j2);
                    in_edge_at(in, i2, 0, i, wi); in_edge_at(in, j2, 0, j, wj);
                    const uint32_t dc = dst_cell(in, i2, j2);
                    for (int r = 0; r <= a.R; ++r) {
                        const int32_t v = dsm ? lds_s32(tiles32 + ((uint32_t)(r + 2) * (uint32_t)ST + dc) * 4u) : __ldcg(a.gtile + (long long)dc * a.gcs + 2 + r);
                        fold_pos(f, a.R, in.k2, r, i2, j2, v, i, j);
                    }
                }
            }
        }
        if (CHECK) {
            for (int o = 16; o > 0; o >>= 1) {
                f.sum += __shfl_down_sync(0xFFFFFFFFu, f.sum, o);
                f.live += __shfl_down_sync(0xFFFFFFFFu, f.live, o);
            }
            if (lane == 0 && f.live) {
                atomicAdd(a.level_sum + l + 1, f.sum);
                atomicAdd(a.level_live + l + 1, f.live);
            }
        }
        if (++slot == (uint32_t)NS) { slot = 0; parity ^= 1u; }
    }
    if (profiling) { a.prof[0] = pc_wait; a.prof[1] = pc_work; a.prof[2] = pc_bar; a.prof[3] = (unsigned long long)n_my; }
    // the sink: layers of cell (0,0) of the last level (the traceback starts from layer R, :774-776)
    if (cta == 0) {
        if (!a.last_smem) {                                     // last level in HBM (DipGenie's sink level is one vertex: never there)
            if (a.grid > 1 && tid == 0 && !failed) wait_counter4(a.counter, a.final_target, a.timeout_ns, a.n_trans);
            bar_named(1, CT);
        }
        for (int r = tid; r <= a.R; r += CT) {
            int32_t v;
            if (a.last_smem) v = lds_s32(tiles32 + ((uint32_t)(r + 2) * (uint32_t)ST + a.sink_cell) * 4u);
            else v = __ldcg(a.gtile + (long long)a.sink_cell * a.gcs + 2 + r);
            a.sink[r] = v;
        }
    }
}

template <int ST, int RC, bool CHECK, bool GIANT>
__global__ void __launch_bounds__(544, 1) dip_sweep4_kernel(const __grid_constant__ Sweep4Args a) {
    sweep4_body<ST, RC, CHECK, GIANT>(a, (int)blockIdx.x);
}

// Many independent problems in ONE launch (dg_dip_run_many): CTA b works on problem cta_map[b].x as its local CTA
// cta_map[b].y; the problem's arguments are copied to shared memory once.
template <int ST, int RC, bool GIANT>
__global__ void __launch_bounds__(544, 1) dip_sweep4_many_kernel(const Sweep4Args* __restrict__ all, const int2* __restrict__ cta_map) {
    __shared__ Sweep4Args sa;
    const int2 who = cta_map[blockIdx.x];
    static_assert(sizeof(Sweep4Args) % 4 == 0, "word copy");
    const uint32_t* src = reinterpret_cast<const uint32_t*>(all + who.x);
    uint32_t* dst = reinterpret_cast<uint32_t*>(&sa);
    for (uint32_t x = threadIdx.x; x < sizeof(Sweep4Args) / 4; x += blockDim.x) dst[x] = src[x];
    __syncthreads();
    sweep4_body<ST, RC, false, GIANT>(sa, who.y);
}

constexpr size_t sweep4_smem_bytes(int stride, int RL, int slot_bytes, int nslot) {
    return (size_t)nslot * (size_t)slot_bytes + (size_t)(RL + 2) * (size_t)stride * 4 + (2 * (size_t)nslot + 1) * 8;
}

}  // namespace dg
