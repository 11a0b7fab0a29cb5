// dp_diploid.cu — B200 diploid recombination-constrained DP (include/dipgenie_cuda.h: dg_dp_diploid, dg_dip_*).
//
// Replaces Approximator::diploid_dp_approximation_solver's sweep and edge-list recovery
// (reference src/approximator.cpp:532-716, :757-785).  Design:
//   * gather form: every destination cell (r2,i',j') takes the first strict maximum over
//     in-edges(i') x in-edges(j') in ascending source order — no locks, no atomics, order-free, and the
//     winner is exactly the reference's (:657-659);
//   * K4 `dip_delta_kernel`: the pair scores of every (e1,e2) in-edge pair of a coloured transition
//     (:604-624) are evaluated once, fully parallel, into a u16 matrix D_l[e1][e2];
//   * K5 `dip_sweep_kernel`: one persistent cooperative launch interprets per-CTA task streams
//     (dp_cell.h: TaskHdr).  Narrow transitions run on CTA 0 alone with both (R+1) x k x k int32 layers in
//     shared memory and one named barrier per level; wide ones are split by destination row over P CTAs
//     with the layers in HBM/L2 and closed by a monotone-counter grid barrier.  A producer warp streams
//     every task's header, in-edge record and matrix rows into a 16-slot shared-memory ring with TMA bulk
//     copies (cp.async.bulk + mbarrier full/empty pairs), far enough ahead to hide HBM latency;
//   * only a 2-byte predecessor code per cell is streamed out to HBM;
//   * K7 traceback: checkpoint every 128 levels; all cells of a checkpoint level walk back to the
//     previous checkpoint in parallel (`dip_anc_kernel`), one thread hops checkpoint to checkpoint from the
//     sink cell (`dip_hop_kernel`), the segments of the winning path are walked in parallel
//     (`dip_seg_kernel`) and concatenated (`dip_merge_kernel`).
#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <condition_variable>
#include <memory>
#include <memory_resource>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#if defined(_OPENMP)
#include <omp.h>
#endif

#include "dg_common.cuh"
#include "dp_cell.h"
#include "dp_plan4.h"
#include "dp_prep.h"

namespace dg {

// ---- kernel geometry -------------------------------------------------------------------------
constexpr int DIP_NCW = 15;                    // compute warps per CTA (16 warps with the producer: 4 per SM sub-partition, 128 registers each)
constexpr int DIP_CT = DIP_NCW * 32;           // compute threads per CTA
constexpr int DIP_THREADS = DIP_CT + 32;       // + one producer warp (task prefetch via bulk async copies)
constexpr int DIP_NSLOT = 16;                  // task ring depth
constexpr int DIP_SLOT_BYTES = 4096;
constexpr int DIP_TILE_CELLS = 16384;          // int32 cells per shared-memory layer tile (x2)
constexpr size_t DIP_SMEM_BYTES = (size_t)DIP_NSLOT * DIP_SLOT_BYTES + 2 * (size_t)DIP_TILE_CELLS * 4 + 2 * DIP_NSLOT * 8 +
                                  (size_t)LANE_SCRATCH_ENTRIES * 8;
constexpr int DIP_TRACE_T = 128;               // levels between traceback checkpoints

constexpr int DG_MAX_PEERS = 8;

struct SweepArgs {
    const TaskHdr* tasks;
    const int64_t* task_begin;
    const uint8_t* records;
    const uint16_t* delta;
    const int64_t* delta_off;
    const int32_t* level_off;
    const int32_t* in_off;
    const uint32_t* in_edge;
    const int32_t* lvlW;
    const int64_t* msrc_off;
    const int64_t* mdst_off;
    const uint64_t* masks;
    int32_t* tile0;
    int32_t* tile1;
    void* pred;
    unsigned int* counter;
    unsigned long long* level_sum;    // [L] (CHECK only)
    unsigned long long* level_live;   // [L]
    unsigned long long* prof;         // [32] phase cycle counters of CTA 0 / thread 0 (nullable)
    int32_t R;
    int32_t shift;                    // layers hold value << shift (dp_cell.h: packed keys), 0 or KEY_SHIFT
    // row-sharded sweep over `world` GPUs (1 = off): the peers' layer tiles, predecessor codes and counters, mapped
    // through CUDA IPC (entry `rank` = this GPU's own); counter[0] = level arrivals, counter[1] = CTAs that left
    int32_t world, rank;
    int32_t* peer_tile0[DG_MAX_PEERS];
    int32_t* peer_tile1[DG_MAX_PEERS];
    uint8_t* peer_pred[DG_MAX_PEERS];
    unsigned int* peer_counter[DG_MAX_PEERS];
    uint32_t exit_target;             // world x CTAs per rank
    unsigned long long timeout_ns;    // a barrier wait longer than this fails the run (counter[2] = level + 1)
};

// ---- PTX helpers -----------------------------------------------------------------------------
__device__ __forceinline__ void red_release_add_u32(unsigned int* p, unsigned int v) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// Spin until the monotone arrival counter reaches `target`.  Polls are relaxed loads with a short fixed
// sleep (the pollers are the idle CTAs of a narrow stretch: a few dozen, not a whole grid); one acquire
// fence orders the layer reads after the last poll.
__device__ __forceinline__ void wait_counter(const unsigned int* counter, unsigned int target) {
    for (;;) {
        unsigned int v;
        asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
        if (v >= target) break;
        __nanosleep(40);
    }
    asm volatile("fence.acq_rel.gpu;" ::: "memory");
}
// The same across GPUs (row-sharded sweep): system-scope release / acquire on peer-mapped counters.
__device__ __forceinline__ void red_release_sys_add_u32(unsigned int* p, unsigned int v) {
    asm volatile("red.release.sys.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// Gives up after `timeout_ns` (a peer that never launched or died must not hang this GPU): returns false, and the
// caller flags the run as failed.
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// `failed` = the run's error word (counter[2] of this GPU): non-zero once any wait of any rank has timed out (the rank that
// gives up writes it into every peer's word too), after which nobody waits any more — the failure is sticky, so a dead
// peer costs ONE timeout, not one per level.
__device__ __noinline__ bool wait_counter_sys(const unsigned int* counter, unsigned int target, unsigned long long timeout_ns,
                                              const unsigned int* failed) {
    const unsigned long long t0 = global_ns();
    bool ok = true;
    for (unsigned int it = 0;; ++it) {
        unsigned int v;
        asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
        if (v >= target) break;
        __nanosleep(40);
        if ((it & 255u) == 255u) {
            unsigned int f;
            asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(f) : "l"(failed) : "memory");
            if (f != 0u || global_ns() - t0 > timeout_ns) { ok = false; break; }
        }
    }
    asm volatile("fence.acq_rel.sys;" ::: "memory");
    return ok;
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// TMA 1-D bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "DG_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DG_DONE_%=;\n"
        "bra DG_WAIT_%=;\n"
        "DG_DONE_%=:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bar_compute() { asm volatile("bar.sync 1, %0;" ::"n"(DIP_CT) : "memory"); }

// ---- K4: pair-score matrices -------------------------------------------------------------------
struct DeltaArgs {
    const int32_t* delta_list;
    int32_t n_list;
    const int32_t* level_off;
    const int32_t* in_off;
    const uint32_t* in_edge;
    const uint16_t* in_dst;
    const int32_t* lvlW;
    const int64_t* msrc_off;
    const int64_t* mdst_off;
    const uint64_t* masks;
    const int64_t* delta_off;
    uint16_t* delta;
};

// D_l[e1][e2] = |(Hom u1 ∪ Hom v1) ∩ (Hom u2 ∪ Hom v2)| + |(Het u1 ∪ Het v1) △ (Het u2 ∪ Het v2)| for the
// edge pair e1 = (u1 -> u2), e2 = (v1 -> v2) entering level l+1 (approximator.cpp:604-624), one CTA per
// coloured transition, all transitions in one launch.
// One transition's matrix, its elements dealt to `nthr` threads starting at `thr`.
__device__ __forceinline__ void delta_matrix(const DeltaArgs& a, int l, uint32_t thr, uint32_t nthr) {
    const int32_t mid = a.level_off[l + 1];
    const int32_t e0 = a.in_off[mid];
    const uint32_t n_in = (uint32_t)(a.in_off[a.level_off[l + 2]] - e0);
    const int W = a.lvlW[l];
    const uint64_t* msrc = a.masks + a.msrc_off[l];
    const uint64_t* mdst = a.masks + a.mdst_off[l];
    uint16_t* D = a.delta + a.delta_off[l];
    const uint32_t total = n_in * n_in;
    for (uint32_t x = thr; x < total; x += nthr) {
        const uint32_t e1 = x / n_in, e2 = x - e1 * n_in;
        const uint32_t ea = __ldg(a.in_edge + e0 + e1), eb = __ldg(a.in_edge + e0 + e2);
        const int i2 = (int)__ldg(a.in_dst + e0 + e1), j2 = (int)__ldg(a.in_dst + e0 + e2);
        D[x] = (uint16_t)mask_delta(W, msrc, mdst, (int)(ea & 0xFFFFu), (int)(eb & 0xFFFFu), i2, j2);
    }
}

// Small matrices (a few hundred elements: nearly all of them on real panels) go one per WARP, so that eight times as
// many chains of dependent index loads are in flight; matrices of more than DELTA_WARP_MAX elements one per block.
constexpr uint32_t DELTA_WARP_MAX = 2048;
__device__ __forceinline__ uint32_t delta_elems_of(const DeltaArgs& a, int l) {
    const uint32_t n_in = (uint32_t)(a.in_off[a.level_off[l + 2]] - a.in_off[a.level_off[l + 1]]);
    return n_in * n_in;
}
__global__ void __launch_bounds__(256) dip_delta_kernel(const DeltaArgs a) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    for (int t = blockIdx.x * wpb + warp; t < a.n_list; t += gridDim.x * wpb) {
        const int l = a.delta_list[t];
        if (delta_elems_of(a, l) <= DELTA_WARP_MAX) delta_matrix(a, l, (uint32_t)lane, 32u);
    }
    for (int t = blockIdx.x; t < a.n_list; t += gridDim.x) {
        const int l = a.delta_list[t];
        if (delta_elems_of(a, l) > DELTA_WARP_MAX) delta_matrix(a, l, threadIdx.x, blockDim.x);
    }
}

// ---- K5: the sweep ------------------------------------------------------------------------------
struct CellIO {
    const int32_t* src;        // source layer (shared or global)
    int32_t* dst;              // destination layer (shared or global)
    uint8_t* pl;               // predecessor codes of the destination level
    bool src_smem, dst_smem;
};

// Work item = (destination pair (i',j'), chunk of RC layers); pairs vary fastest over the threads so that
// the stores of a warp are contiguous in every layer; when a task has fewer pairs than threads, `groups`
// thread groups share the pairs and deal the chunks round-robin.  SMEM: both layers are shared-memory
// tiles (the common case inside a narrow stretch): plain LDS/STS and 32-bit indices; otherwise placement
// is a run-time flag and HBM layers are read/written with L2-only (.cg) accesses, since other SMs produce
// and consume them.  CHECK / PRED32 are kernel-level template flags so that the timed kernel carries
// neither the checksum fold nor the 32-bit code path.
template <class OffT, int RC, bool MASKS, bool SMEM, bool CHECK, bool PRED32>
__device__ __forceinline__ void sweep_items(const TransitionT<OffT>& t, const CellIO& io, int R, const TaskHdr& h, int tid,
                                            unsigned long long& hsum, unsigned long long& hlive) {
    using IdxT = typename std::conditional<SMEM, int32_t, int64_t>::type;
    const uint32_t k2 = h.k2, npairs = (uint32_t)(h.i1 - h.i0) * k2;
    const uint32_t nchunk = (uint32_t)(R + RC) / RC;
    const uint32_t groups = h.groups;
    uint32_t g = 0, p = (uint32_t)tid;
    if (groups > 1) {
        g = h.m_pairs ? div_magic((uint32_t)tid, h.m_pairs) : (uint32_t)tid;
        p = (uint32_t)tid - g * npairs;
        if (g >= groups) return;
    }
    const int32_t* __restrict__ src = io.src;
    const bool ssm = io.src_smem;
    auto load = [src, ssm](IdxT idx) -> int32_t {
        if (SMEM) return src[idx];
        return ssm ? src[idx] : __ldcg(src + idx);
    };
    const IdxT kk2 = (IdxT)k2 * (IdxT)k2;
    for (uint32_t pair = p; pair < npairs; pair += DIP_CT) {
        const uint32_t ir = h.m_k2 ? div_magic(pair, h.m_k2) : pair;
        const uint32_t j2 = pair - ir * k2, i2 = (uint32_t)h.i0 + ir;
        const IdxT cell0 = (IdxT)i2 * (IdxT)k2 + (IdxT)j2;
        for (uint32_t chunk = g; chunk < nchunk; chunk += groups) {
            const int r0 = (int)chunk * RC;
            int32_t best[RC];
            uint32_t code[RC];
            relax_pair<RC, MASKS, IdxT>(t, load, R, r0, (int)i2, (int)j2, best, code);
#pragma unroll
            for (int rr = 0; rr < RC; ++rr) {
                const int r2 = r0 + rr;
                if (r2 <= R) {
                    const IdxT c = (IdxT)r2 * kk2 + cell0;
                    const bool live = best[rr] >= 0;
                    const int32_t v = live ? best[rr] : NEG_INF;
                    if (SMEM || io.dst_smem) io.dst[c] = v; else __stcg(io.dst + c, v);
                    if (PRED32) reinterpret_cast<uint32_t*>(io.pl)[c] = code[rr];
                    else reinterpret_cast<uint16_t*>(io.pl)[c] = (uint16_t)(((code[rr] >> 16) << 8) | (code[rr] & 0xFFu));   // dead: 0xFFFF
                    if (CHECK && live) {
                        const int pi = (int)(t.in_edge[(int32_t)t.in_off[i2] + (int32_t)(code[rr] >> 16)] & 0xFFFFu);
                        const int pj = (int)(t.in_edge[(int32_t)t.in_off[j2] + (int32_t)(code[rr] & 0xFFFFu)] & 0xFFFFu);
                        ++hlive;
                        hsum += cell_fold((uint64_t)c, v >> t.dshift, pi, pj);
                    }
                }
            }
        }
    }
}

template <class OffT, bool MASKS, bool SMEM, bool CHECK, bool PRED32>
__device__ __forceinline__ void sweep_rc(const TransitionT<OffT>& t, const CellIO& io, int R, const TaskHdr& h, int tid,
                                         unsigned long long& hsum, unsigned long long& hlive) {
    sweep_items<OffT, DIP_RC, MASKS, SMEM, CHECK, PRED32>(t, io, R, h, tid, hsum, hlive);
}

template <class OffT, bool SMEM, bool CHECK, bool PRED32>
__device__ __forceinline__ void sweep_dm(const TransitionT<OffT>& t, const CellIO& io, int R, const TaskHdr& h, int tid,
                                         unsigned long long& hsum, unsigned long long& hlive) {
    if (!SMEM && (h.flags & TK_DELTA_MASKS)) sweep_rc<OffT, true, false, CHECK, PRED32>(t, io, R, h, tid, hsum, hlive);
    else sweep_rc<OffT, false, SMEM, CHECK, PRED32>(t, io, R, h, tid, hsum, hlive);
}

// ---- the lane form, written against 32-bit shared-window addresses --------------------------------
__device__ __forceinline__ uint32_t lds_u16(uint32_t a) { uint32_t v; asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint32_t lds_u32(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ int32_t lds_s32(uint32_t a) { int32_t v; asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint4 lds_v4(uint32_t a) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts_s32(uint32_t a, int32_t v) { asm volatile("st.shared.s32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }

__device__ __forceinline__ int32_t ldcg_s32(const int32_t* p) { return __ldcg(p); }

// Lane form of a task (dp_cell.h).  SMEM: both layers are shared-memory tiles of this CTA (32-bit shared-window
// addresses, LDS/STS only); otherwise both layers are in HBM/L2 (ld/st.global.cg: other SMs produce and consume
// them).  The record and the staged rows of the pair-score matrix always sit in the task slot.
struct LaneTask {
    uint32_t sb32;           // slot base (shared window)
    uint32_t src32, dst32;   // SMEM: tile addresses
    const int32_t* gsrc;     // !SMEM: layers in global memory
    int32_t* gdst;
    uint8_t* pl;
    uint32_t k, k2, i0, i1, n_in, rec_bytes, skew, nblk, rp, nrg, m_nblk, m_nrg, m_nin, n_witems, rounds, bstart_off;
    uint32_t n_long, long_off;       // TK_LONG: destinations with more than 32 in-edges, their positions in the record
    unsigned long long* scratch;     //          64-bit combine words [(row - i0) * n_long + g][layer]
    const uint16_t* gdelta;          // pair-score matrix of the transition in global memory when its rows are not staged (else null)
    const uint8_t* grec;             // TK_REC_GLOBAL: the record, read in place (too big for a slot); else null
    bool staged;
};

// Record reads of the lane form: `off` = byte offset inside the record.  RG: in place from global memory (read-only
// path), else from the task slot.
template <bool RG>
__device__ __forceinline__ uint32_t rec16(const LaneTask& t, uint32_t off) {
    if (RG) return (uint32_t)__ldg(reinterpret_cast<const uint16_t*>(t.grec + off));
    return lds_u16(t.sb32 + (uint32_t)sizeof(TaskHdr) + off);
}
template <bool RG>
__device__ __forceinline__ uint32_t rec32(const LaneTask& t, uint32_t off) {
    if (RG) return __ldg(reinterpret_cast<const uint32_t*>(t.grec + off));
    return lds_u32(t.sb32 + (uint32_t)sizeof(TaskHdr) + off);
}

struct LaneProf { unsigned long long items, setup, loop, reduce, store, iters; };

// The lane-form view of the task staged in slot `sb32` (header words as laid out in dp_cell.h: TaskHdr).
template <bool PRED32>
__device__ __forceinline__ void fill_lane_task(const SweepArgs& a, uint32_t sb32, uint32_t tiles32, unsigned long long* scratch,
                                               const uint4& h1, const uint4& h2, LaneTask& lt) {
    const uint4 h3 = lds_v4(sb32 + 48u), h4 = lds_v4(sb32 + 64u), h5 = lds_v4(sb32 + 80u);
    const uint32_t flags = h1.y;
    const int l = (int)h1.x;
    lt.sb32 = sb32;
    const uint32_t odd = (uint32_t)l & 1u;
    lt.src32 = tiles32 + odd * ((uint32_t)DIP_TILE_CELLS * 4u);
    lt.dst32 = tiles32 + (odd ^ 1u) * ((uint32_t)DIP_TILE_CELLS * 4u);
    lt.gsrc = odd ? a.tile1 : a.tile0;
    lt.gdst = odd ? a.tile0 : a.tile1;
    const unsigned long long pred_off2 = ((unsigned long long)h3.y << 32) | h3.x;
    lt.pl = reinterpret_cast<uint8_t*>(a.pred) + ((size_t)pred_off2 << (PRED32 ? 2 : 1));
    lt.k = h2.x & 0xFFFFu; lt.k2 = h2.x >> 16; lt.i0 = h2.y & 0xFFFFu; lt.i1 = h2.y >> 16; lt.n_in = h2.z;
    lt.rec_bytes = lds_u32(sb32 + 4u); lt.skew = h1.w;
    lt.nblk = h4.y & 0xFFFFu; lt.rp = h4.y >> 16; lt.nrg = h4.z; lt.m_nblk = h4.w;
    lt.m_nrg = h5.x; lt.m_nin = h5.y; lt.n_witems = h5.z; lt.rounds = h5.w;
    lt.bstart_off = lds_u32(sb32 + 96u);
    lt.n_long = 0; lt.long_off = 0; lt.scratch = scratch;
    if (flags & TK_LONG) { lt.n_long = lds_u32(sb32 + 100u); lt.long_off = lds_u32(sb32 + 104u); }
    lt.staged = (flags & TK_DELTA_STAGED) != 0;
    lt.gdelta = ((flags & TK_DELTA) && !lt.staged) ? a.delta + __ldg(a.delta_off + l) : nullptr;
    lt.grec = (flags & TK_REC_GLOBAL) ? a.records + (size_t)lds_u32(sb32) * 16 : nullptr;
}

// Unpacked arithmetic (value and code in separate registers): problems whose DP values may exceed the packed key
// (shift == 0) and, with shift == KEY_SHIFT, the tasks of levels that have destinations with more than 32 in-edges
// (TK_LONG: slice blocks are reduced over the whole warp and meet in the scratch words).  Out of line: the packed
// loop keeps its own register allocation.
template <int RC, bool SMEM, bool CHECK, bool PRED32, bool PROF, bool RG>
__device__ __noinline__ ulonglong2 lane_task(const SweepArgs& a, uint32_t sb32, uint32_t tiles32, unsigned long long* scratch, int warp, int lane) {
    unsigned long long hsum = 0, hlive = 0;
    LaneProf lp = {0, 0, 0, 0, 0, 0};
    LaneTask t;
    fill_lane_task<PRED32>(a, sb32, tiles32, scratch, lds_v4(sb32 + 16u), lds_v4(sb32 + 32u), t);
    const int R = a.R, shift = a.shift;
    const uint32_t slot32 = t.sb32 + (uint32_t)sizeof(TaskHdr);      // staged pair-score rows follow the staged record
    const uint32_t edge_o = (uint32_t)rec_edge_offset((int)t.k2);    // byte offsets inside the record
    const uint32_t dstp_o = (uint32_t)rec_dst_offset((int)t.k2, t.n_in);
    const uint32_t bst_o = t.bstart_off;
    const uint32_t k = t.k, k2 = t.k2, n_in = t.n_in;
    const uint32_t kk = k * k, kk2 = k2 * k2;
    // lane -> (row slot, in-edge within the block)
    uint32_t rs = 0, el = (uint32_t)lane;
    if (t.nblk == 1) {
        rs = t.m_nin ? __umulhi((uint32_t)lane, t.m_nin) : (uint32_t)lane;
        el = (uint32_t)lane - rs * n_in;
    }
    uint32_t delta32 = 0;
    if (t.staged) delta32 = slot32 + t.rec_bytes + 2u * t.skew - 2u * rec16<RG>(t, 2u * t.i0) * n_in;   // row of in-edge e1 at + 2*e1*n_in
    for (uint32_t wi = (uint32_t)warp; wi < t.n_witems; wi += DIP_NCW) {
        long long c0 = 0, c1 = 0, c2 = 0, c3 = 0;
        if (PROF) c0 = clock64();
        const uint32_t q = t.m_nblk ? __umulhi(wi, t.m_nblk) : wi;
        const uint32_t b = wi - q * t.nblk;
        const uint32_t chunk = t.m_nrg ? __umulhi(q, t.m_nrg) : q;
        const uint32_t rg = q - chunk * t.nrg;
        const uint32_t bs = rec16<RG>(t, bst_o + 2u * b), be = rec16<RG>(t, bst_o + 2u * b + 2u);
        const uint32_t row = t.i0 + rg * t.rp + rs, e2 = bs + el;
        const bool valid = rs < t.rp && row < t.i1 && e2 < be;
        const uint32_t rowc = valid ? row : t.i0, e2c = valid ? e2 : bs;
        const uint32_t a0 = rec16<RG>(t, 2u * rowc);
        uint32_t a1 = rec16<RG>(t, 2u * rowc + 2u);
        if (!valid) a1 = a0;
        const uint32_t y = rec32<RG>(t, edge_o + 4u * e2c);
        const uint32_t j2 = rec16<RG>(t, dstp_o + 2u * e2c);
        const uint32_t s0 = rec16<RG>(t, 2u * j2), s1 = rec16<RG>(t, 2u * j2 + 2u);
        const uint32_t pos = e2c - s0, seg = s1 - s0;
        const uint32_t j = y & 0xFFFFu;
        const int wv = (int)(y >> 16);
        const int r0 = (int)chunk * RC;
        int32_t best[RC];
        uint32_t code[RC];
#pragma unroll
        for (int rr = 0; rr < RC; ++rr) { best[rr] = -1; code[rr] = 0xFFFFFFFFu; }
        if (PROF) c1 = clock64();
        for (uint32_t e1 = a0; e1 < a1; ++e1) {
            if (PROF) ++lp.iters;
            const uint32_t x = rec32<RG>(t, edge_o + 4u * e1);
            const uint32_t base = (x & 0xFFFFu) * k + j;
            const int w = (int)(x >> 16) + wv;
            int d = 0;
            if (t.staged) d = (int)lds_u16(delta32 + 2u * (e1 * n_in + e2c));
            else if (t.gdelta) d = (int)__ldg(t.gdelta + ((size_t)e1 * n_in + e2c));
            d <<= shift;
            const uint32_t cd = ((e1 - a0) << 16) | pos;
            // loads in batches of LB independent accesses (layer index clamped into [0,R]; validity gates the compare)
            constexpr int LB = (RC % 10 == 0) ? 10 : ((RC % 8 == 0) ? 8 : RC);
#pragma unroll
            for (int b0 = 0; b0 < RC; b0 += LB) {
                int32_t v[LB];
#pragma unroll
                for (int q = 0; q < LB; ++q) {
                    int r = r0 + b0 + q - w;
                    r = r < 0 ? 0 : (r > R ? R : r);
                    if (SMEM) v[q] = lds_s32(t.src32 + 4u * ((uint32_t)r * kk + base));
                    else v[q] = ldcg_s32(t.gsrc + ((size_t)r * kk + base));
                }
#pragma unroll
                for (int q = 0; q < LB; ++q) {
                    const int rr = b0 + q;
                    const bool ok = (r0 + rr - w >= 0) && (r0 + rr <= R);
                    const int32_t c = v[q] + d;
                    if (ok && c > best[rr]) { best[rr] = c; code[rr] = cd; }
                }
            }
        }
        if (PROF) c2 = clock64();
        // segmented maximum over the lanes of one destination column: value desc, then code asc; a slice block (one
        // destination with more than 32 in-edges) is reduced over all its lanes
        const bool slice = t.n_long && __any_sync(0xFFFFFFFFu, valid && seg > 32u);
        const uint32_t nrd = slice ? 5u : t.rounds;
        for (uint32_t rd = 0, off = 1; rd < nrd; ++rd, off <<= 1) {
            const bool partner = slice ? (valid && el + off < be - bs) : (pos + off < seg);
#pragma unroll
            for (int rr = 0; rr < RC; ++rr) {
                const int32_t ob = __shfl_down_sync(0xFFFFFFFFu, best[rr], off);
                const uint32_t oc = __shfl_down_sync(0xFFFFFFFFu, code[rr], off);
                if (partner && (ob > best[rr] || (ob == best[rr] && oc < code[rr]))) { best[rr] = ob; code[rr] = oc; }
            }
        }
        if (PROF) c3 = clock64();
        if (slice) {
            if (valid && el == 0) {              // the slice's partial maxima meet the other slices in the scratch words
                uint32_t g = 0;
                while (g + 1 < t.n_long && rec16<RG>(t, t.long_off + 2u * g) != j2) ++g;
                unsigned long long* sc = t.scratch + ((size_t)(row - t.i0) * t.n_long + g) * (uint32_t)(R + 1);
#pragma unroll
                for (int rr = 0; rr < RC; ++rr)
                    if (r0 + rr <= R && best[rr] >= 0)
                        atomicMax(sc + (r0 + rr), ((unsigned long long)((uint32_t)best[rr] + 1u) << 32) | (unsigned long long)(0xFFFFFFFFu - code[rr]));
            }
        } else if (valid && pos == 0) {
            const uint32_t cell0 = row * k2 + j2;
#pragma unroll
            for (int rr = 0; rr < RC; ++rr) {
                const int r2 = r0 + rr;
                if (r2 <= R) {
                    const bool live = best[rr] >= 0;
                    const int32_t val = live ? best[rr] : NEG_INF;
                    size_t c;
                    if (SMEM) { const uint32_t c32 = (uint32_t)r2 * kk2 + cell0; sts_s32(t.dst32 + 4u * c32, val); c = c32; }
                    else { c = (size_t)r2 * kk2 + cell0; __stcg(t.gdst + c, val); }
                    if (PRED32) reinterpret_cast<uint32_t*>(t.pl)[c] = code[rr];
                    else reinterpret_cast<uint16_t*>(t.pl)[c] = (uint16_t)(((code[rr] >> 16) << 8) | (code[rr] & 0xFFu));   // dead: 0xFFFF
                    if (CHECK && live) {
                        const int pi = (int)(rec32<RG>(t, edge_o + 4u * (a0 + (code[rr] >> 16))) & 0xFFFFu);
                        const int pj = (int)(rec32<RG>(t, edge_o + 4u * (s0 + (code[rr] & 0xFFFFu))) & 0xFFFFu);
                        ++hlive;
                        hsum += cell_fold((uint64_t)c, val >> shift, pi, pj);
                    }
                }
            }
        }
        if (PROF) { const long long c4 = clock64(); lp.items += 1; lp.setup += c1 - c0; lp.loop += c2 - c1; lp.reduce += c3 - c2; lp.store += c4 - c3; }
    }
    return make_ulonglong2(hsum, hlive);
}

// The same with packed keys (dp_cell.h): one word per layer, one add + one max per candidate layer, one shuffle
// per layer and round.  Only the first chunk of layers runs a checked loop (layer index below 0 for weighted
// edges); the tiles are padded to whole chunks, so the layers above R that the last chunk loads are in bounds
// (and never stored).
template <int RC, bool SMEM, bool CHECK, bool PRED32, bool PROF, bool RG>
__device__ __forceinline__ void lane_task_packed(const LaneTask& t, int R, int warp, int lane,
                                                 unsigned long long& hsum, unsigned long long& hlive, LaneProf& lp) {
    constexpr uint32_t ORD_MASK = (1u << (2 * KEY_ORD_BITS)) - 1u, ORD_ONE = (1u << KEY_ORD_BITS) - 1u;
    const uint32_t slot32 = t.sb32 + (uint32_t)sizeof(TaskHdr);      // staged pair-score rows follow the staged record
    const uint32_t edge_o = (uint32_t)rec_edge_offset((int)t.k2);    // byte offsets inside the record
    const uint32_t dstp_o = (uint32_t)rec_dst_offset((int)t.k2, t.n_in);
    const uint32_t bst_o = t.bstart_off;
    const uint32_t k = t.k, k2 = t.k2, n_in = t.n_in;
    const uint32_t kk = k * k, kk2 = k2 * k2;
    uint32_t rs = 0, el = (uint32_t)lane;
    if (t.nblk == 1) {
        rs = t.m_nin ? __umulhi((uint32_t)lane, t.m_nin) : (uint32_t)lane;
        el = (uint32_t)lane - rs * n_in;
    }
    uint32_t delta32 = 0;
    if (t.staged) delta32 = slot32 + t.rec_bytes + 2u * t.skew - 2u * rec16<RG>(t, 2u * t.i0) * n_in;
    // what depends on the lane's in-edge only; with a single block (the common case) it is the same for every
    // warp item of the task and is computed once
    uint32_t bs = 0, be = 0, e2c = 0, j2 = 0, s0 = 0, pos = 0, seg = 1, j = 0;
    int wv = 0;
    bool lane_ok = false;
    auto lane_setup = [&](uint32_t b) {
        bs = rec16<RG>(t, bst_o + 2u * b); be = rec16<RG>(t, bst_o + 2u * b + 2u);
        const uint32_t e2 = bs + el;
        lane_ok = e2 < be;
        e2c = lane_ok ? e2 : bs;
        const uint32_t y = rec32<RG>(t, edge_o + 4u * e2c);
        j2 = rec16<RG>(t, dstp_o + 2u * e2c);
        s0 = rec16<RG>(t, 2u * j2);
        const uint32_t s1 = rec16<RG>(t, 2u * j2 + 2u);
        pos = e2c - s0; seg = s1 - s0;
        j = y & 0xFFFFu; wv = (int)(y >> 16);
    };
    if (t.nblk == 1) lane_setup(0);
    for (uint32_t wi = (uint32_t)warp; wi < t.n_witems; wi += DIP_NCW) {
        long long c0 = 0, c1 = 0, c2 = 0, c3 = 0;
        if (PROF) c0 = clock64();
        uint32_t q = wi;
        if (t.nblk != 1) {
            q = t.m_nblk ? __umulhi(wi, t.m_nblk) : wi;
            lane_setup(wi - q * t.nblk);
        }
        const uint32_t chunk = t.m_nrg ? __umulhi(q, t.m_nrg) : q;
        const uint32_t rg = q - chunk * t.nrg;
        const uint32_t row = t.i0 + rg * t.rp + rs;
        const bool valid = lane_ok && rs < t.rp && row < t.i1;
        const uint32_t rowc = valid ? row : t.i0;
        const uint32_t a0 = rec16<RG>(t, 2u * rowc);
        uint32_t a1 = rec16<RG>(t, 2u * rowc + 2u);
        if (!valid) a1 = a0;
        const int r0 = (int)chunk * RC;
        const bool edge_chunk = (r0 == 0);     // warp-uniform; layers above R (last chunk) are loaded inside the padded tile and never stored
        int32_t key[RC];
#pragma unroll
        for (int rr = 0; rr < RC; ++rr) key[rr] = -1;
        uint32_t ordbits = ((ORD_ONE) << KEY_ORD_BITS) | (ORD_ONE - pos);    // (31 - e1 ordinal) << 5 | (31 - e2 ordinal), e1 ordinal 0
        if (PROF) c1 = clock64();
        for (uint32_t e1 = a0; e1 < a1; ++e1, ordbits -= (1u << KEY_ORD_BITS)) {
            if (PROF) ++lp.iters;
            const uint32_t x = rec32<RG>(t, edge_o + 4u * e1);
            const uint32_t base = (x & 0xFFFFu) * k + j;
            const int w = (int)(x >> 16) + wv;
            uint32_t dp = ordbits;
            if (t.staged) dp += lds_u16(delta32 + 2u * (e1 * n_in + e2c)) << KEY_SHIFT;
            else if (t.gdelta) dp += (uint32_t)__ldg(t.gdelta + ((size_t)e1 * n_in + e2c)) << KEY_SHIFT;
            int32_t v[RC];
            if (!edge_chunk) {
                if (SMEM) {
                    const uint32_t a = t.src32 + 4u * ((uint32_t)(r0 - w) * kk + base);
#pragma unroll
                    for (int rr = 0; rr < RC; ++rr) v[rr] = lds_s32(a + 4u * (uint32_t)rr * kk);
                } else {
                    const int32_t* a = t.gsrc + ((size_t)(r0 - w) * kk + base);
#pragma unroll
                    for (int rr = 0; rr < RC; ++rr) v[rr] = ldcg_s32(a + (size_t)rr * kk);
                }
#pragma unroll
                for (int rr = 0; rr < RC; ++rr) key[rr] = max(key[rr], v[rr] + (int32_t)dp);
            } else {
#pragma unroll
                for (int rr = 0; rr < RC; ++rr) {
                    int r = r0 + rr - w;
                    r = r < 0 ? 0 : (r > R ? R : r);
                    if (SMEM) v[rr] = lds_s32(t.src32 + 4u * ((uint32_t)r * kk + base));
                    else v[rr] = ldcg_s32(t.gsrc + ((size_t)r * kk + base));
                }
#pragma unroll
                for (int rr = 0; rr < RC; ++rr) {
                    const bool ok = (r0 + rr - w >= 0) && (r0 + rr <= R);
                    key[rr] = max(key[rr], ok ? v[rr] + (int32_t)dp : -1);
                }
            }
        }
        if (PROF) c2 = clock64();
        // segmented maximum over the lanes of one destination column (keys of different lanes never tie)
        for (uint32_t rd = 0, off = 1; rd < t.rounds; ++rd, off <<= 1) {
            const bool partner = pos + off < seg;
#pragma unroll
            for (int rr = 0; rr < RC; ++rr) {
                const int32_t ok = __shfl_down_sync(0xFFFFFFFFu, key[rr], off);
                key[rr] = max(key[rr], partner ? ok : -1);
            }
        }
        if (PROF) c3 = clock64();
        if (valid && pos == 0) {
            const uint32_t cell0 = row * k2 + j2;
#pragma unroll
            for (int rr = 0; rr < RC; ++rr) {
                const int r2 = r0 + rr;
                if (r2 <= R) {
                    const bool live = key[rr] >= 0;
                    const int32_t val = live ? (int32_t)((uint32_t)key[rr] & ~ORD_MASK) : NEG_INF;
                    const uint32_t inv = ~(uint32_t)key[rr] & ORD_MASK;                 // e1 ordinal << 5 | e2 ordinal
                    const uint32_t o1 = inv >> KEY_ORD_BITS, o2 = inv & ORD_ONE;
                    size_t c;
                    if (SMEM) { const uint32_t c32 = (uint32_t)r2 * kk2 + cell0; sts_s32(t.dst32 + 4u * c32, val); c = c32; }
                    else { c = (size_t)r2 * kk2 + cell0; __stcg(t.gdst + c, val); }
                    if (PRED32) reinterpret_cast<uint32_t*>(t.pl)[c] = live ? ((o1 << 16) | o2) : 0xFFFFFFFFu;
                    else reinterpret_cast<uint16_t*>(t.pl)[c] = live ? (uint16_t)((o1 << 8) | o2) : (uint16_t)0xFFFFu;
                    if (CHECK && live) {
                        const int pi = (int)(rec32<RG>(t, edge_o + 4u * (a0 + o1)) & 0xFFFFu);
                        const int pj = (int)(rec32<RG>(t, edge_o + 4u * (s0 + o2)) & 0xFFFFu);
                        ++hlive;
                        hsum += cell_fold((uint64_t)c, val >> KEY_SHIFT, pi, pj);
                    }
                }
            }
        }
        if (PROF) { const long long c4 = clock64(); lp.items += 1; lp.setup += c1 - c0; lp.loop += c2 - c1; lp.reduce += c3 - c2; lp.store += c4 - c3; }
    }
}

// Second pass of a TK_LONG task: every scratch word (row, long destination, layer) becomes its cell.
template <bool CHECK, bool PRED32>
__device__ __noinline__ ulonglong2 long_finalize(const SweepArgs& a, uint32_t sb32, uint32_t tiles32, unsigned long long* scratch, bool smem_layers, int tid) {
    unsigned long long hsum = 0, hlive = 0;
    LaneTask t;
    fill_lane_task<PRED32>(a, sb32, tiles32, scratch, lds_v4(sb32 + 16u), lds_v4(sb32 + 32u), t);
    const int R = a.R;
    const uint32_t RP1 = (uint32_t)R + 1u, G = t.n_long;
    const uint32_t total = (t.i1 - t.i0) * G * RP1;
    const uint32_t edge_o = (uint32_t)rec_edge_offset((int)t.k2);
    const uint32_t kk2 = t.k2 * t.k2;
    // (not a hot loop: the record is read through a run-time switch between the slot and its place in global memory)
    auto r16 = [&](uint32_t off) { return t.grec ? rec16<true>(t, off) : rec16<false>(t, off); };
    auto r32 = [&](uint32_t off) { return t.grec ? rec32<true>(t, off) : rec32<false>(t, off); };
    for (uint32_t idx = (uint32_t)tid; idx < total; idx += DIP_CT) {
        const unsigned long long K = t.scratch[idx];
        const uint32_t q = idx / RP1, r2 = idx - q * RP1, rowrel = q / G, g = q - rowrel * G;
        const uint32_t row = t.i0 + rowrel, j2 = r16(t.long_off + 2u * g);
        const bool live = K != 0ull;
        const int32_t val = live ? (int32_t)((uint32_t)(K >> 32) - 1u) : NEG_INF;
        const uint32_t code = live ? 0xFFFFFFFFu - (uint32_t)K : 0xFFFFFFFFu;
        const size_t c = (size_t)r2 * kk2 + (size_t)row * t.k2 + j2;
        if (smem_layers) sts_s32(t.dst32 + 4u * (uint32_t)c, val); else __stcg(t.gdst + c, val);
        if (PRED32) reinterpret_cast<uint32_t*>(t.pl)[c] = code;
        else reinterpret_cast<uint16_t*>(t.pl)[c] = (uint16_t)(((code >> 16) << 8) | (code & 0xFFu));   // dead: 0xFFFF
        if (CHECK && live) {
            const uint32_t a0 = r16(2u * row), s0 = r16(2u * j2);
            const int pi = (int)(r32(edge_o + 4u * (a0 + (code >> 16))) & 0xFFFFu);
            const int pj = (int)(r32(edge_o + 4u * (s0 + (code & 0xFFFFu))) & 0xFFFFu);
            ++hlive;
            hsum += cell_fold((uint64_t)c, val >> KEY_SHIFT, pi, pj);
        }
    }
    return make_ulonglong2(hsum, hlive);
}

// The pair form (everything the lane form does not take): layers in HBM/L2 (or a hand-over between the two placements),
// records staged or read in place, pair scores staged / in place / popcounted on the fly.  Kept out of line
// so that the narrow loop gets its own register allocation and stays small.
template <bool CHECK, bool PRED32>
__device__ __noinline__ ulonglong2 generic_task(const SweepArgs& a, const uint8_t* sb, int32_t* tileS0, int32_t* tileS1, int tid) {
    unsigned long long hsum = 0, hlive = 0;
    const TaskHdr h = *reinterpret_cast<const TaskHdr*>(sb);
    const uint32_t rec_bytes = h.rec_bytes;
    CellIO io;
    io.pl = reinterpret_cast<uint8_t*>(a.pred) + ((size_t)h.pred_off2 << (PRED32 ? 2 : 1));
    io.src_smem = (h.flags & TK_SRC_SMEM) != 0; io.dst_smem = (h.flags & TK_DST_SMEM) != 0;
    const int l = h.level;
    const uint32_t flags = h.flags;
    const bool ssm = io.src_smem, dsm = io.dst_smem;
    io.src = ssm ? ((l & 1) ? tileS1 : tileS0) : ((l & 1) ? a.tile1 : a.tile0);
    io.dst = dsm ? ((l & 1) ? tileS0 : tileS1) : ((l & 1) ? a.tile0 : a.tile1);
    if (!(flags & TK_REC_GLOBAL)) {
        TransitionT<uint16_t> tr;
        tr.k = h.k; tr.k2 = h.k2;
        tr.in_off = reinterpret_cast<const uint16_t*>(sb + sizeof(TaskHdr));
        tr.in_edge = reinterpret_cast<const uint32_t*>(sb + sizeof(TaskHdr) + rec_edge_offset(h.k2));
        tr.delta = nullptr; tr.dstride = h.n_in; tr.e1_base = 0; tr.e2_base = 0; tr.dshift = a.shift;
        tr.W = 0; tr.msrc = nullptr; tr.mdst = nullptr;
        if (flags & TK_DELTA_STAGED) {
            tr.delta = reinterpret_cast<const uint16_t*>(sb + sizeof(TaskHdr) + rec_bytes) + h.delta_skew;
            tr.e1_base = (int32_t)tr.in_off[h.i0];
        } else if (flags & TK_DELTA) {
            tr.delta = a.delta + __ldg(a.delta_off + l);
        } else if (flags & TK_DELTA_MASKS) {
            tr.W = __ldg(a.lvlW + l); tr.msrc = a.masks + __ldg(a.msrc_off + l); tr.mdst = a.masks + __ldg(a.mdst_off + l);
        }
        sweep_dm<uint16_t, false, CHECK, PRED32>(tr, io, a.R, h, tid, hsum, hlive);
    } else {
        const int32_t mid = __ldg(a.level_off + l + 1);
        const int32_t ebase = __ldg(a.in_off + mid);
        TransitionT<int32_t> tr;
        tr.k = h.k; tr.k2 = h.k2;
        tr.in_off = a.in_off + mid;
        tr.in_edge = a.in_edge;
        tr.delta = nullptr; tr.dstride = __ldg(a.in_off + __ldg(a.level_off + l + 2)) - ebase;
        tr.e1_base = ebase; tr.e2_base = ebase; tr.dshift = a.shift;
        tr.W = 0; tr.msrc = nullptr; tr.mdst = nullptr;
        if (flags & TK_DELTA) tr.delta = a.delta + __ldg(a.delta_off + l);
        else if (flags & TK_DELTA_MASKS) {
            tr.W = __ldg(a.lvlW + l); tr.msrc = a.masks + __ldg(a.msrc_off + l); tr.mdst = a.masks + __ldg(a.mdst_off + l);
        }
        sweep_dm<int32_t, false, CHECK, PRED32>(tr, io, a.R, h, tid, hsum, hlive);
    }
    return make_ulonglong2(hsum, hlive);
}

// Row-sharded sweep: copies rows [i0, i1) of every layer of level l+1 — values from this GPU's tile, predecessor codes
// from its code array — to the same places on every peer GPU (plain stores through the NVLink peer mappings).
template <bool PRED32>
__device__ __noinline__ void push_rows(const SweepArgs& a, int l, uint32_t k2, uint32_t i0, uint32_t i1,
                                       unsigned long long pred_off2, int tid) {
    const uint32_t kk2 = k2 * k2, seg = (i1 - i0) * k2, first = i0 * k2;
    const bool odd = ((l + 1) & 1) != 0;
    const int32_t* const src = odd ? a.tile1 : a.tile0;
    const uint8_t* const psrc = reinterpret_cast<const uint8_t*>(a.pred);
    const uint32_t total = ((uint32_t)a.R + 1u) * seg;          // (layer, cell of the row range), one index
#pragma unroll 4
    for (uint32_t idx = (uint32_t)tid; idx < total; idx += DIP_CT) {
        const uint32_t r = idx / seg, x = idx - r * seg;
        const size_t cell = (size_t)r * kk2 + first + x;
        const int32_t v = __ldcg(src + cell);
        uint32_t code;
        if (PRED32) code = __ldcg(reinterpret_cast<const uint32_t*>(psrc) + pred_off2 + cell);
        else code = __ldcg(reinterpret_cast<const uint16_t*>(psrc) + pred_off2 + cell);
        for (int q = 0; q < a.world; ++q) {
            if (q == a.rank) continue;
            (odd ? a.peer_tile1[q] : a.peer_tile0[q])[cell] = v;
            if (PRED32) reinterpret_cast<uint32_t*>(a.peer_pred[q])[pred_off2 + cell] = code;
            else reinterpret_cast<uint16_t*>(a.peer_pred[q])[pred_off2 + cell] = (uint16_t)code;
        }
    }
}

template <bool CHECK, bool PRED32, bool PROF>
__device__ __forceinline__ void sweep_body(const SweepArgs& a, const int cta) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* const slots = smem;
    int32_t* const tileS0 = reinterpret_cast<int32_t*>(smem + (size_t)DIP_NSLOT * DIP_SLOT_BYTES);
    int32_t* const tileS1 = tileS0 + DIP_TILE_CELLS;
    uint64_t* const full = reinterpret_cast<uint64_t*>(tileS1 + DIP_TILE_CELLS);
    uint64_t* const empty = full + DIP_NSLOT;
    unsigned long long* const scratch = reinterpret_cast<unsigned long long*>(empty + DIP_NSLOT);   // TK_LONG tasks

    const int tid = threadIdx.x, lane = tid & 31;
    const int32_t n = (int32_t)(a.task_begin[cta + 1] - a.task_begin[cta]);
    const TaskHdr* const my_tasks = a.tasks + a.task_begin[cta];
    uint8_t* const pred = reinterpret_cast<uint8_t*>(a.pred);
    constexpr int pshift = PRED32 ? 2 : 1;

    if (tid == 0) {
        for (int s = 0; s < DIP_NSLOT; ++s) { mbar_init(smem_u32(full + s), 1); mbar_init(smem_u32(empty + s), DIP_NCW); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int x = tid; x <= a.R; x += DIP_THREADS) tileS0[x] = 0;   // level 0: all R+1 layers start at 0 (approximator.cpp:535)
    __syncthreads();

    if (tid >= DIP_CT) {
        // ---- producer warp: lanes fetch 32 task descriptors at a time, lane 0 issues the copies ----
        for (int32_t t0 = 0; t0 < n; t0 += 32) {
            uint4 f = make_uint4(0u, 0u, 0u, 0u);
            if (t0 + lane < n) f = __ldg(reinterpret_cast<const uint4*>(my_tasks + t0 + lane));
            const int cnt = min(32, n - t0);
            for (int j = 0; j < cnt; ++j) {
                const uint32_t rec_off16 = __shfl_sync(0xFFFFFFFFu, f.x, j), rec_bytes = __shfl_sync(0xFFFFFFFFu, f.y, j);
                const uint32_t d_off16 = __shfl_sync(0xFFFFFFFFu, f.z, j), d_bytes = __shfl_sync(0xFFFFFFFFu, f.w, j);
                if (lane == 0) {
                    const int32_t t = t0 + j;
                    const int slot = t % DIP_NSLOT, use = t / DIP_NSLOT;
                    if (use > 0) mbar_wait(smem_u32(empty + slot), (uint32_t)((use - 1) & 1));
                    const uint32_t bar = smem_u32(full + slot);
                    const uint32_t dst = smem_u32(slots + (size_t)slot * DIP_SLOT_BYTES);
                    mbar_expect_tx(bar, (uint32_t)sizeof(TaskHdr) + rec_bytes + d_bytes);
                    bulk_g2s(dst, my_tasks + t, (uint32_t)sizeof(TaskHdr), bar);
                    if (rec_bytes) bulk_g2s(dst + (uint32_t)sizeof(TaskHdr), a.records + (size_t)rec_off16 * 16, rec_bytes, bar);
                    if (d_bytes) bulk_g2s(dst + (uint32_t)sizeof(TaskHdr) + rec_bytes, reinterpret_cast<const uint8_t*>(a.delta) + (size_t)d_off16 * 16, d_bytes, bar);
                }
            }
        }
        return;
    }

    // ---- compute warps ----
    __shared__ int s_failed;                 // row-sharded sweep: a cross-GPU wait has timed out, stop working (the ring keeps moving)
    if (tid == 0) s_failed = 0;
    bar_compute();
    const bool profiling = PROF && cta == 0 && tid == 0;
    unsigned long long pc0 = 0, pc1 = 0, pc2 = 0, pc3 = 0, pc4 = 0, pc5 = 0, pc6 = 0, pc7 = 0, pc8 = 0, pc9 = 0;
    long long tk0 = 0, tk1 = 0, tk2 = 0, tk3 = 0, tk4 = 0;
    LaneProf lp = {0, 0, 0, 0, 0, 0};
    // shared-window addresses, computed once
    const uint32_t slots32 = smem_u32(slots), tiles32 = smem_u32(tileS0), full32 = smem_u32(full), empty32 = smem_u32(empty);
    const int warp = tid >> 5;
    for (int32_t t = 0; t < n; ++t) {
        const uint32_t slot = (uint32_t)t % DIP_NSLOT;
        const uint32_t sb32 = slots32 + slot * DIP_SLOT_BYTES;
        const uint8_t* const sb = slots + (size_t)slot * DIP_SLOT_BYTES;
        if (profiling) tk0 = clock64();
        mbar_wait(full32 + 8u * slot, ((uint32_t)t / DIP_NSLOT) & 1u);
        if (profiling) tk1 = clock64();
        // exec part of the header (16-byte shared loads; the lane-form words only when needed)
        const uint4 h1 = lds_v4(sb32 + 16u), h2 = lds_v4(sb32 + 32u);
        const uint32_t flags = h1.y;
        const int l = (int)h1.x;
        if (flags & TK_WAIT) {
            if (tid == 0) {
                if (a.world > 1) {
                    if (!s_failed && !wait_counter_sys(a.counter, h1.z, a.timeout_ns, a.counter + 2)) {
                        s_failed = 1;
                        if (atomicCAS(a.counter + 2, 0u, (unsigned int)l + 1u) == 0u) {
                            printf("dip_sweep_kernel: rank %d CTA %d gave up at level %d: %u of %u arrivals\n", a.rank, cta, l, *(volatile unsigned int*)a.counter, h1.z);
                            for (int q = 0; q < a.world; ++q)          // tell the peers: nobody waits for this run any more
                                if (q != a.rank) atomicCAS(a.peer_counter[q] + 2, 0u, (unsigned int)l + 1u);
                        }
                    }
                } else wait_counter(a.counter, h1.z);
            }
            bar_compute();
        }
        const bool failed = a.world > 1 && s_failed != 0;      // (set before a block barrier: the same for every thread of the task)
        if (profiling) tk2 = clock64();
        unsigned long long hsum = 0, hlive = 0;
        const bool ssm = (flags & TK_SRC_SMEM) != 0, dsm = (flags & TK_DST_SMEM) != 0;
        const bool is_long = (flags & TK_LONG) != 0;
        if (is_long) {                               // slices of a long destination meet in the scratch words: zero them
            const uint32_t rows = (h2.y >> 16) - (h2.y & 0xFFFFu);
            const uint32_t total = rows * lds_u32(sb32 + 100u) * ((uint32_t)a.R + 1u);
            for (uint32_t x = (uint32_t)tid; x < total; x += DIP_CT) scratch[x] = 0ull;
            bar_compute();
        }
        if ((uint32_t)(tid & ~31) < h2.w && !failed) {          // this warp owns work of the task
            if (flags & TK_LANES) {
                if (a.shift && !is_long) {
                    LaneTask lt;
                    fill_lane_task<PRED32>(a, sb32, tiles32, scratch, h1, h2, lt);
                    const uint32_t rc = lds_u32(sb32 + 64u) & 0xFFFFu;
                    const bool big = rc == (uint32_t)LANE_RC_BIG;
                    if (ssm) {
                        if (big) lane_task_packed<LANE_RC_BIG, true, CHECK, PRED32, PROF, false>(lt, a.R, warp, lane, hsum, hlive, lp);
                        else lane_task_packed<LANE_RC_SMALL, true, CHECK, PRED32, PROF, false>(lt, a.R, warp, lane, hsum, hlive, lp);
                    } else if (flags & TK_REC_GLOBAL) {      // record read in place (wide panels)
                        if (big) lane_task_packed<LANE_RC_BIG, false, CHECK, PRED32, PROF, true>(lt, a.R, warp, lane, hsum, hlive, lp);
                        else lane_task_packed<LANE_RC_SMALL, false, CHECK, PRED32, PROF, true>(lt, a.R, warp, lane, hsum, hlive, lp);
                    } else {
                        if (big) lane_task_packed<LANE_RC_BIG, false, CHECK, PRED32, PROF, false>(lt, a.R, warp, lane, hsum, hlive, lp);
                        else lane_task_packed<LANE_RC_SMALL, false, CHECK, PRED32, PROF, false>(lt, a.R, warp, lane, hsum, hlive, lp);
                    }
                } else {
                    const ulonglong2 hs = ssm ? lane_task<LANE_RC_SMALL, true, CHECK, PRED32, false, false>(a, sb32, tiles32, scratch, warp, lane)
                                          : (flags & TK_REC_GLOBAL) ? lane_task<LANE_RC_SMALL, false, CHECK, PRED32, false, true>(a, sb32, tiles32, scratch, warp, lane)
                                                                    : lane_task<LANE_RC_SMALL, false, CHECK, PRED32, false, false>(a, sb32, tiles32, scratch, warp, lane);
                    hsum = hs.x; hlive = hs.y;
                }
            } else {
                const ulonglong2 hs = generic_task<CHECK, PRED32>(a, sb, tileS0, tileS1, tid);
                hsum = hs.x; hlive = hs.y;
            }
        }
        if (is_long) {                               // second pass: every scratch word becomes its cell
            bar_compute();
            const ulonglong2 hs = long_finalize<CHECK, PRED32>(a, sb32, tiles32, scratch, ssm, tid);
            hsum += hs.x; hlive += hs.y;
        }
        if (profiling) tk3 = clock64();
        uint32_t pw = 0;                                        // TK_PUSH: header words read before the slot is handed back
        unsigned long long push_pred_off = 0;
        uint32_t arr_target = 0, arr_n = 0;
        if ((flags & TK_ARRIVE) && a.world > 1) { arr_target = lds_u32(sb32 + 112u); arr_n = lds_u32(sb32 + 116u); }
        if (flags & TK_PUSH) {
            pw = lds_u32(sb32 + 108u);
            push_pred_off = ((unsigned long long)lds_u32(sb32 + 52u) << 32) | lds_u32(sb32 + 48u);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty32 + 8u * slot);       // this warp is done with the slot
        if (flags & TK_BAR) bar_compute();                     // the destination rows of this CTA are whole
        if ((flags & TK_PUSH) && !failed) {
            // row-sharded sweep: this CTA's rows of layer l+1 (values and predecessor codes) go to every peer over NVLink
            push_rows<PRED32>(a, l, h2.x >> 16, pw & 0xFFFFu, pw >> 16, push_pred_off, tid);
            __threadfence_system();
            bar_compute();
        }
        if ((flags & TK_ARRIVE) && tid == 0) {
            if (a.world > 1) {
                // count on this GPU first; the rank's last CTA of the level forwards all of them at once (one remote
                // atomic per rank and level instead of one per CTA: remote atomics on one word serialise)
                unsigned int old;
                asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], 1;" : "=r"(old) : "l"(a.counter + 3) : "memory");
                if (old + 1u == arr_target)
                    for (int q = 0; q < a.world; ++q) red_release_sys_add_u32(a.peer_counter[q], arr_n);
            }
            else red_release_add_u32(a.counter, 1u);
        }
        if (profiling) {
            tk4 = clock64();
            if (ssm || dsm) { pc0 += 1; pc1 += tk1 - tk0; pc2 += tk2 - tk1; pc3 += tk3 - tk2; pc4 += tk4 - tk3; }
            else { pc5 += 1; pc6 += tk1 - tk0; pc7 += tk2 - tk1; pc8 += tk3 - tk2; pc9 += tk4 - tk3; }
        }
        if (CHECK) {
            for (int o = 16; o > 0; o >>= 1) {
                hsum += __shfl_down_sync(0xFFFFFFFFu, hsum, o);
                hlive += __shfl_down_sync(0xFFFFFFFFu, hlive, o);
            }
            if (lane == 0 && hlive) {
                atomicAdd(a.level_sum + l + 1, hsum);
                atomicAdd(a.level_live + l + 1, hlive);
            }
        }
    }
    if (a.world > 1) {
        // every CTA of every rank reports that all its pushes are out; CTA 0 keeps the kernel alive until all have,
        // so that what follows on the stream (traceback) sees the complete predecessor codes
        __threadfence_system();
        bar_compute();
        if (tid == 0) {
            for (int q = 0; q < a.world; ++q) red_release_sys_add_u32(a.peer_counter[q] + 1, 1u);
            if (cta == 0 && !s_failed && !wait_counter_sys(a.counter + 1, a.exit_target, a.timeout_ns, a.counter + 2)) {
                if (atomicCAS(a.counter + 2, 0u, 0x7FFFFFFFu) == 0u)
                    printf("dip_sweep_kernel: rank %d gave up at the exit barrier: %u of %u CTAs\n", a.rank, *(volatile unsigned int*)(a.counter + 1), a.exit_target);
            }
        }
    }
    if (profiling) {
        a.prof[0] = pc0; a.prof[1] = pc1; a.prof[2] = pc2; a.prof[3] = pc3; a.prof[4] = pc4; a.prof[5] = 0;
        a.prof[6] = pc5; a.prof[7] = pc6; a.prof[8] = pc7; a.prof[9] = pc8; a.prof[10] = pc9; a.prof[11] = 0;
        a.prof[12] = lp.items; a.prof[13] = lp.setup; a.prof[14] = lp.loop; a.prof[15] = lp.reduce; a.prof[16] = lp.store; a.prof[17] = lp.iters;
    }
}

template <bool CHECK, bool PRED32, bool PROF>
__global__ void __launch_bounds__(DIP_THREADS, 1) dip_sweep_kernel(const __grid_constant__ SweepArgs a) {
    sweep_body<CHECK, PRED32, PROF>(a, (int)blockIdx.x);
}

// Many independent problems in ONE launch (dg_dip_run_many): CTA b works on problem cta_map[b].x as its local CTA
// cta_map[b].y.  A GPU runs at most 32 streams' kernels side by side (hardware work queues), a grid has no such
// limit: one or two CTAs per sample fill all SMs with samples.  The problem's arguments are copied to shared memory once.
template <bool PRED32>
__global__ void __launch_bounds__(DIP_THREADS, 1) dip_sweep_many_kernel(const SweepArgs* __restrict__ all, const int2* __restrict__ cta_map) {
    __shared__ SweepArgs sa;
    const int2 who = cta_map[blockIdx.x];
    static_assert(sizeof(SweepArgs) % 4 == 0, "word copy");
    const uint32_t* src = reinterpret_cast<const uint32_t*>(all + who.x);
    uint32_t* dst = reinterpret_cast<uint32_t*>(&sa);
    for (uint32_t x = threadIdx.x; x < sizeof(SweepArgs) / 4; x += blockDim.x) dst[x] = src[x];
    __syncthreads();
    sweep_body<false, PRED32, false>(sa, who.y);
}

}  // namespace dg

#include "dp_sweep4.cuh"

namespace dg {

// ---- K7: traceback ------------------------------------------------------------------------------
struct TraceOut {           // device-side result block
    int32_t rc, value, s_het, n1, n2, pad[3];
};

struct TraceArgs {
    TraceView v;
    const void* pred;
    const int32_t* cp;        // [M+1] checkpoint levels, cp[0] = L-1 (sink level), descending, cp[M] = 0
    const int64_t* aoff;      // [M+1] prefix of (R+1)*k^2 over checkpoint levels cp[0..M-1]
    int32_t M;
    int32_t* anc;             // [aoff[M]] flat cell at cp[m+1] that cell x of cp[m] descends from (-1: dead)
    int32_t* path_cell;       // [M] cell of the winning path at cp[m] (-1: dead)
    int32_t* seg_p1;          // [M][2*cap]
    int32_t* seg_p2;
    int32_t* seg_n;           // [M][4]: n1, n2, s_het, rc
    const int32_t* sink_tile;
    const int32_t* sink_v4;   // level-program engine: the sink cell's layers [R+1] (else null)
    int cap;
    int shift;               // layers hold value << shift
    TraceOut* out;
    int32_t* p1;
    int32_t* p2;
};

DG_HD void cell_to_state(int64_t cell, int32_t k, TraceState& s) {
    const int64_t kk = (int64_t)k * k;
    s.r = (int32_t)(cell / kk);
    const int32_t rem = (int32_t)(cell - (int64_t)s.r * kk);
    s.i2 = rem / k; s.j2 = rem - s.i2 * k;
}

template <class PredT>
__device__ __forceinline__ void anc_body(const TraceArgs& a) {
    const int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= a.aoff[a.M]) return;
    int lo = 0, hi = a.M - 1;                 // largest m with aoff[m] <= x
    while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (a.aoff[mid] <= x) lo = mid; else hi = mid - 1; }
    const int m = lo;
    const int l_hi = a.cp[m], l_lo = a.cp[m + 1];
    TraceState s;
    cell_to_state(x - a.aoff[m], a.v.level_off[l_hi + 1] - a.v.level_off[l_hi], s);
    const PredT* pred = reinterpret_cast<const PredT*>(a.pred);
    bool live = true;
    for (int l = l_hi - 1; l >= l_lo && live; --l) {
        int wu, wv, i2, j2;
        live = trace_step<PredT>(a.v, pred, l, s, wu, wv, i2, j2);
    }
    const int32_t k = a.v.level_off[l_lo + 1] - a.v.level_off[l_lo];
    a.anc[x] = live ? (int32_t)(((int64_t)s.r * k + s.i2) * k + s.j2) : -1;
}
template <class PredT>
__global__ void __launch_bounds__(256) dip_anc_kernel(const TraceArgs a) { anc_body<PredT>(a); }
// The same for many problems in one launch: blockIdx.y = problem (blocks beyond a problem's cells leave at once).
template <class PredT>
__global__ void __launch_bounds__(256) dip_anc_many_kernel(const TraceArgs* __restrict__ all) { anc_body<PredT>(all[blockIdx.y]); }

__device__ __forceinline__ void hop_body(const TraceArgs& a) {
    const int64_t ks = a.v.level_off[a.v.L] - a.v.level_off[a.v.L - 1];
    const int32_t value = a.sink_v4 ? __ldcg(a.sink_v4 + a.v.R)
                                    : __ldcg(a.sink_tile + (int64_t)a.v.R * ks * ks);   // cell (r=R,0,0) of the last level (:730, :775)
    a.out->value = value < 0 ? NEG_INF : (value >> a.shift);
    int64_t cur = (value < 0) ? -1 : (int64_t)a.v.R * ks * ks;
    for (int m = 0; m < a.M; ++m) {
        a.path_cell[m] = (int32_t)cur;
        if (cur >= 0) cur = a.anc[a.aoff[m] + cur];
    }
}
__global__ void dip_hop_kernel(const TraceArgs a) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    hop_body(a);
}
__global__ void dip_hop_many_kernel(const TraceArgs* __restrict__ all, int n) {       // one thread per problem
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) hop_body(all[i]);
}

template <class PredT>
__device__ __forceinline__ void seg_body(const TraceArgs& a) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= a.M) return;
    int32_t* sn = a.seg_n + 4 * m;
    sn[0] = sn[1] = sn[2] = 0; sn[3] = -1;
    const int32_t cell = a.path_cell[m];
    if (cell < 0) return;
    const int l_hi = a.cp[m], l_lo = a.cp[m + 1];
    TraceState s;
    cell_to_state(cell, a.v.level_off[l_hi + 1] - a.v.level_off[l_hi], s);
    int32_t n1 = 0, n2 = 0, sh = 0;
    const int rc = trace_segment<PredT>(a.v, reinterpret_cast<const PredT*>(a.pred), l_hi, l_lo, s,
                                        a.seg_p1 + (size_t)m * 2 * a.cap, &n1, a.seg_p2 + (size_t)m * 2 * a.cap, &n2, a.cap, &sh);
    sn[0] = n1; sn[1] = n2; sn[2] = sh; sn[3] = rc;
}
template <class PredT>
__global__ void __launch_bounds__(128) dip_seg_kernel(const TraceArgs a) { seg_body<PredT>(a); }
template <class PredT>
__global__ void __launch_bounds__(128) dip_seg_many_kernel(const TraceArgs* __restrict__ all) { seg_body<PredT>(all[blockIdx.y]); }

__device__ __forceinline__ void merge_body(const TraceArgs& a) {
    int n1 = 0, n2 = 0, sh = 0, rc = 0;
    for (int m = 0; m < a.M && rc == 0; ++m) {      // m = 0 is nearest the sink: lists come out newest-first
        const int32_t* sn = a.seg_n + 4 * m;
        if (sn[3] != 0) { rc = sn[3]; break; }
        if (n1 + sn[0] > a.cap || n2 + sn[1] > a.cap) { rc = -2; break; }
        for (int x = 0; x < 2 * sn[0]; ++x) a.p1[2 * n1 + x] = a.seg_p1[(size_t)m * 2 * a.cap + x];
        for (int x = 0; x < 2 * sn[1]; ++x) a.p2[2 * n2 + x] = a.seg_p2[(size_t)m * 2 * a.cap + x];
        n1 += sn[0]; n2 += sn[1]; sh += sn[2];
    }
    if (rc == -1) { n1 = 0; n2 = 0; sh = 0; }
    a.out->rc = rc; a.out->s_het = sh; a.out->n1 = n1; a.out->n2 = n2;
}
__global__ void dip_merge_kernel(const TraceArgs a) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    merge_body(a);
}
__global__ void dip_merge_many_kernel(const TraceArgs* __restrict__ all, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) merge_body(all[i]);
}

}  // namespace dg

using namespace dg;

// Page-locked staging block of one planned problem (batch path): the plan's big arrays are written into it.
struct PlanStaging {
    dg_ctx* ctx = nullptr;
    dg_ctx::Pinned block = {nullptr, 0};
    std::unique_ptr<std::pmr::monotonic_buffer_resource> res;
    // A free block of at least `bytes` from the context's pool, else a new one; on failure the plan simply lives on the heap.
    void acquire(dg_ctx* c, size_t bytes) {
        ctx = c;
        {
            std::lock_guard<std::mutex> lk(c->pinned_mu);
            for (size_t x = 0; x < c->pinned_free.size(); ++x)
                if (c->pinned_free[x].bytes >= bytes) { block = c->pinned_free[x]; c->pinned_free.erase(c->pinned_free.begin() + (long)x); break; }
        }
        if (!block.p) {
            cudaSetDevice(c->device);
            void* q = nullptr;
            const cudaError_t e = cudaHostAlloc(&q, bytes, cudaHostAllocDefault);
            if (e == cudaSuccess) block = {q, bytes};
            else {
                (void)cudaGetLastError();
                if (getenv("DG_TIMING")) fprintf(stderr, "batch: no page-locked block of %zu bytes (%s): this plan goes to the heap\n", bytes, cudaGetErrorString(e));
            }
        }
        if (block.p) res.reset(new std::pmr::monotonic_buffer_resource(block.p, block.bytes, std::pmr::new_delete_resource()));
    }
    std::pmr::memory_resource* resource() { return res ? res.get() : std::pmr::get_default_resource(); }
    void release() {                     // only once nothing allocated from the block is alive or in flight
        res.reset();
        if (block.p) { std::lock_guard<std::mutex> lk(ctx->pinned_mu); ctx->pinned_free.push_back(block); }
        block = {nullptr, 0};
    }
    ~PlanStaging() { release(); }
};

struct dg_dip {
    explicit dg_dip(std::unique_ptr<PlanStaging> st = nullptr)
        : staging(std::move(st)), plan(staging ? staging->resource() : std::pmr::get_default_resource()),
          p4(staging ? staging->resource() : std::pmr::get_default_resource()) {}
    std::unique_ptr<PlanStaging> staging;   // declared before `plan`: outlives the plan's arrays
    cudaStream_t stream = nullptr;   // the context's stream, or one of the batch streams
    bool cooperative = true;         // batch slots use plain launches (see dip_run_impl)
    int shift = 0;                   // KEY_SHIFT when every DP value provably stays below 2^21 (packed keys), else 0
    int lane_rc = LANE_RC_SMALL;     // layers per lane of the lane form (tile padding follows it)
    std::vector<int32_t> h_cp;       // traceback checkpoints (host copies until uploaded)
    std::vector<int64_t> h_aoff;
    DipPlan plan;                 // host copy (small arrays kept for stats; big ones released after upload)
    Plan4 p4;                     // level-program engine: its planning (same storage rules)
    int pred_bytes = 2;
    int grid = 1;
    // row-sharded over `world` GPUs (dg_dip_create_sharded): peer mappings of tile0, tile1, pred, counter
    int world = 1, rank = 0;
    void* peer[4][DG_MAX_PEERS] = {};
    bool attached = false, armed = false, ipc_opened = false;
    int M = 0;                    // traceback segments
    int64_t anc_cells = 0;
    DevBuf<TaskHdr> tasks;
    DevBuf<int64_t> task_begin;
    DevBuf<uint8_t> records;
    DevBuf<uint16_t> delta;
    DevBuf<int64_t> delta_off;
    DevBuf<int32_t> delta_list;
    DevBuf<int32_t> level_off, in_off, lvlW;
    DevBuf<uint32_t> in_edge;
    DevBuf<uint16_t> in_dst;
    DevBuf<uint64_t> masks;
    DevBuf<int64_t> msrc_off, mdst_off, pred_off;
    DevBuf<int32_t> tile0, tile1;
    DevBuf<uint8_t> pred;
    DevBuf<unsigned int> counter;
    DevBuf<unsigned long long> level_sum, level_live, prof;
    bool want_prof = false;
    DevBuf<int32_t> cp, anc, path_cell, seg_p1, seg_p2, seg_n;
    DevBuf<int64_t> aoff;
    DevBuf<TraceOut> tout;
    DevBuf<int32_t> p1, p2;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    float delta_ms = 0.f, sweep_ms = 0.f, trace_ms = 0.f, plan_ms = 0.f, upload_ms = 0.f;
    float fused_ms = 0.f;            // > 0: the last run's sweep was part of a fused launch of this duration (dg_dip_run_many)
    float group_trace_ms = 0.f;      // > 0: ... and so was its traceback (no per-problem events were recorded)
    int launches = 0;
    bool ran = false, checks = false;
    uint64_t device_bytes = 0;
    uint64_t h2d_bytes = 0;          // host arrays copied to the device by dg_dip_create
    // level-program engine (dp_prog.h, dp_sweep4.cuh); v4 == false: the task-stream engine above
    bool v4 = false;
    int v4_ncw = 12;
    float build_ms = 0.f;            // prog_fill_kernel
    DevBuf<ProgDir> v4_dir, v4_dir_full;        // the timed directory (idle transitions skipped) and the complete one (checksums)
    DevBuf<ProgHdr> v4_hdr;
    DevBuf<uint64_t> v4_prog_off;
    DevBuf<int32_t> v4_wide, v4_wide_full, v4_sink;
    DevBuf<unsigned long long> v4_gkey;          // giant cells: where the CTAs' slices meet (dp_sweep4.cuh: giant_cells); null without such cells
    DevBuf<unsigned int> v4_gcnt;
    DevBuf<uint32_t> v4_vup;                     // traceback fast path (dp_cell.h: TraceView::vup)
    DevBuf<uint8_t> v4_prog, v4_dom;
    DevBuf<uint16_t> v4_pred, v4_cls, v4_vslot;
    DevBuf<uint32_t> v4_vinfo, v4_mpre, v4_n1, v4_np, v4_m, v4_z, v4_dm, v4_tflags;
    DevBuf<int64_t> v4_mpre_off;
    DevBuf<Fill4Args> v4_tables;                 // the builder's view of the device tables (also read by the checksum variant)
    // batch path: dg_dip_create's device half returns without waiting for the uploads; the plan's page-locked arrays are
    // handed back once `ev_up` (recorded behind the last copy and the program builder) has completed
    bool async_create = false, release_pending = false;
    cudaEvent_t ev_up = nullptr;
    ~dg_dip() {
        if ((release_pending || async_create) && stream) cudaStreamSynchronize(stream);      // (copies out of the plan's arrays may still be in flight)
        if (ev_up) cudaEventDestroy(ev_up);
        for (auto& e : ev) if (e) cudaEventDestroy(e);
        if (ipc_opened)
            for (int a = 0; a < 4; ++a)
                for (int q = 0; q < world; ++q) if (q != rank && peer[a][q]) cudaIpcCloseMemHandle(peer[a][q]);
    }
};

static const void* sweep_fn(bool pred32, bool check, bool prof = false) {
    if (prof && !check) return pred32 ? (const void*)dip_sweep_kernel<false, true, true> : (const void*)dip_sweep_kernel<false, false, true>;
    if (pred32) return check ? (const void*)dip_sweep_kernel<true, true, false> : (const void*)dip_sweep_kernel<false, true, false>;
    return check ? (const void*)dip_sweep_kernel<true, false, false> : (const void*)dip_sweep_kernel<false, false, false>;
}

// ---- level-program engine: kernel variants and geometry ----
// Kernel variants: shared-memory layer stride (cells) x layers per register chunk; problems with giant cells (an in-degree
// above 32: dp_sweep4.cuh) always take stride 1024 and the variants that carry that code.
#define DG_S4_VARIANTS(X) X(1024, 10) X(1024, 5) X(680, 10) X(680, 5) X(512, 10) X(512, 5) X(256, 10)
static const void* sweep4_fn(int stride, int rc, bool check, bool giant) {
    if (giant) {
        if (stride == 1024 && rc == 10) return check ? (const void*)dip_sweep4_kernel<1024, 10, true, true> : (const void*)dip_sweep4_kernel<1024, 10, false, true>;
        if (stride == 1024 && rc == 5) return check ? (const void*)dip_sweep4_kernel<1024, 5, true, true> : (const void*)dip_sweep4_kernel<1024, 5, false, true>;
        return nullptr;
    }
#define X(ST, RC) if (stride == ST && rc == RC) return check ? (const void*)dip_sweep4_kernel<ST, RC, true, false> : (const void*)dip_sweep4_kernel<ST, RC, false, false>;
    DG_S4_VARIANTS(X)
#undef X
    return nullptr;
}
static const void* sweep4_many_fn(int stride, int rc, bool giant) {
    if (giant) {
        if (stride == 1024 && rc == 10) return (const void*)dip_sweep4_many_kernel<1024, 10, true>;
        if (stride == 1024 && rc == 5) return (const void*)dip_sweep4_many_kernel<1024, 5, true>;
        return nullptr;
    }
#define X(ST, RC) if (stride == ST && rc == RC) return (const void*)dip_sweep4_many_kernel<ST, RC, false>;
    DG_S4_VARIANTS(X)
#undef X
    return nullptr;
}
constexpr size_t S4_SMEM_MAX = 226 * 1024;   // 227 KB per CTA, less the fused kernel's static argument block
// Layer chunk and shared-memory layer stride for R: the widest stride whose two tiles of RL + 2 layers fit beside the
// slot ring.  False: no variant fits (R too large): the task-stream engine takes the problem.
// `packed`: the problem shares the GPU with other resident problems, one CTA each (batch slots): a shared-memory tile of
// stride 680 (26 slots), small ring and few warps, so that two CTAs fit an SM with room left for L1 — the sweep of
// one problem is a chain of dependent levels that leaves its SM mostly idle, a second problem fills the gaps (measured
// on B200, MHC_4, R = 18, fused launch of 256 problems, two per SM: stride 512 316 ms, 680 300 ms, 1024 317 ms; one per
// SM, 144 problems: 262 ms).  Otherwise the problem has SMs to itself: stride 1024
// (31 slots).  Strides: 1024 (31 slots), 680 (26), 512 (22), 256 (15); DG_V4_STRIDE / DG_V4_SLOG override the choice.
static bool sweep4_shape(int R, int grid, bool packed, bool giant, Sweep4Shape& sh, int& rc, int& ncw) {
    rc = 10;
    if (const char* e = getenv("DG_V4_RC")) rc = atoi(e) == 5 ? 5 : 10;
    sh.slot_bytes = packed ? 4096 : 8192;
    if (const char* e = getenv("DG_V4_SLOT")) sh.slot_bytes = std::max(256, atoi(e) / 16 * 16);
    sh.nslot = 4;
    if (const char* e = getenv("DG_V4_NSLOT")) sh.nslot = std::max(2, std::min(16, atoi(e)));
    ncw = packed || giant ? 8 : 12;            // (wide panels, 90 walks: 8 / 12 / 16 compute warps 618 / 636 / 661 ms per sweep)
    if (const char* e = getenv("DG_V4_NCW")) ncw = std::max(1, std::min(16, atoi(e)));
    sh.grid = std::max(1, grid);
    const int RL = (R + rc) / rc * rc;
    static const int strides[4] = {1024, 680, 512, 256}, slots[4] = {31, 26, 22, 15};      // kn^2 < stride: the last cell of a layer stays DEAD
    int first = packed && !giant ? 1 : 0;
    if (const char* e = getenv("DG_V4_SLOG")) { const int sl = std::max(8, std::min(10, atoi(e))); first = sl == 10 ? 0 : (sl == 9 ? 2 : 3); }
    if (const char* e = getenv("DG_V4_STRIDE")) { const int st = atoi(e); for (int x = 0; x < 4; ++x) if (strides[x] == st) first = x; }
    for (int x = first; x < 4; ++x) {
        if (giant && x > 0) break;                     // (giant cells: stride 1024 only)
        if (rc != 10 && strides[x] == 256) break;
        if (sweep4_smem_bytes(strides[x], RL, sh.slot_bytes, sh.nslot) <= S4_SMEM_MAX) {
            sh.stride = strides[x]; sh.kn = slots[x];
            sh.slog = 10;      // (unused once stride is set)
            return true;
        }
    }
    return false;
}

static double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// What the kernels' geometry and the device allow (queried once per call that plans problems).
struct DipLimits {
    int max_grid = 1;
    int64_t delta_budget = 0;
};

static int dip_limits(dg_ctx* ctx, DipLimits& lim) {
    DG_CUDA(ctx, cudaSetDevice(ctx->device));
    for (int v = 0; v < 8; ++v)
        DG_CUDA(ctx, cudaFuncSetAttribute(sweep_fn(v & 1, v & 2, v & 4), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DIP_SMEM_BYTES));
    // Load every kernel a run launches now: with CUDA's lazy module loading the first launch of a kernel can wait for
    // the device to drain, and a sweep that spins on a peer (row-sharded problems) never drains by itself.
    {
        cudaFuncAttributes fa;
        const void* fns[] = {(const void*)dip_delta_kernel, (const void*)dip_anc_kernel<uint16_t>, (const void*)dip_anc_kernel<uint32_t>,
                             (const void*)dip_hop_kernel, (const void*)dip_seg_kernel<uint16_t>, (const void*)dip_seg_kernel<uint32_t>,
                             (const void*)dip_merge_kernel, (const void*)dip_sweep_many_kernel<false>, (const void*)dip_sweep_many_kernel<true>,
                             (const void*)prog_fill_kernel, (const void*)fill_dead_kernel, (const void*)dip_anc_many_kernel<uint16_t>,
                             (const void*)dip_hop_many_kernel, (const void*)dip_seg_many_kernel<uint16_t>, (const void*)dip_merge_many_kernel};
        for (const void* f : fns) DG_CUDA(ctx, cudaFuncGetAttributes(&fa, f));
    }
    int per_sm = 0;
    DG_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, sweep_fn(false, false), DIP_THREADS, DIP_SMEM_BYTES));
    if (per_sm < 1) return fail(ctx, DG_ERR_CUDA, "dg_dip_create: sweep kernel cannot be resident");
    lim.max_grid = ctx->sm_count * per_sm;
    size_t free_b = 0, total_b = 0;
    DG_CUDA(ctx, cudaMemGetInfo(&free_b, &total_b));
    lim.delta_budget = (int64_t)(free_b / 8);
    return DG_OK;
}

// Host half of dg_dip_create: no CUDA calls, safe to run for several problems on several host threads.
static bool dip_plan_host(const DipGraphView& g, const DipLimits& lim, int grid_cap, dg_dip* d) {
    const double t_plan0 = now_ms();
    if (!build_dip_plan(g, d->plan)) return false;
    DipPlan& p = d->plan;
    const int L = p.L;
    d->pred_bytes = (p.max_indeg <= 255) ? 2 : 4;
    d->shift = (p.value_bound < KEY_VALUE_LIMIT && !getenv("DG_NO_PACK")) ? KEY_SHIFT : 0;
    // grid: enough CTAs that the widest transition leaves about 16 candidates per thread, at most one
    // co-resident wave
    uint64_t widest_cand = 0;
    for (int l = 0; l + 1 < L; ++l) {
        const uint64_t n_in = (uint64_t)(p.in_off[p.level_off[l + 2]] - p.in_off[p.level_off[l + 1]]);
        widest_cand = std::max(widest_cand, (uint64_t)(p.R + 1) * n_in * n_in);
    }
    uint64_t want = (widest_cand + (uint64_t)DIP_CT * 16 - 1) / ((uint64_t)DIP_CT * 16);
    if (const char* e = getenv("DG_DIP_GRID")) want = (uint64_t)std::max(1, atoi(e));   // diagnostics / tuning
    if (grid_cap > 0) want = std::min<uint64_t>(want, (uint64_t)grid_cap);
    SweepShape shape;
    shape.grid = (int)std::min<uint64_t>(std::max<uint64_t>(want, 1), (uint64_t)lim.max_grid);
    shape.threads = DIP_CT;
    shape.tile_cells = DIP_TILE_CELLS;
    shape.slot_bytes = DIP_SLOT_BYTES;
    shape.delta_budget = lim.delta_budget;
    shape.lane_rc = (d->shift != 0 && p.R + 1 >= LANE_RC_BIG) ? LANE_RC_BIG : LANE_RC_SMALL;
    d->lane_rc = shape.lane_rc;
    shape.allow_long = d->shift != 0 && !getenv("DG_NO_LONG");     // slice blocks live in the packed-key kernel
    shape.replicas = d->world; shape.rank = d->rank;
    d->v4 = false;
    if (d->world == 1 && !getenv("DG_ENGINE_V3") && !getenv("DG_NO_PACK")) {
        Sweep4Shape s4;
        int rc = 10;
        std::string why = "no kernel variant for this R";
        if (sweep4_shape(p.R, shape.grid, !d->cooperative && shape.grid == 1, p.max_indeg > 32, s4, rc, d->v4_ncw) && plan4_build(p, g, s4, rc, d->p4, why)) d->v4 = true;
        else {
            d->p4.release_arrays();           // (whatever the failed attempt left in the page-locked block: it is handed back before d dies)
            if (getenv("DG_TIMING")) fprintf(stderr, "dg_dip: task-stream engine (%s)\n", why.c_str());
        }
    }
    if (d->v4) {
        d->grid = d->p4.full.wide_list.empty() ? 1 : shape.grid;      // no HBM-resident transition: CTA 0 does everything
        d->pred_bytes = 2; d->shift = KEY_SHIFT;
    } else {
        plan_tasks(p, shape);
        d->grid = 1;
        for (int c = 0; c < p.grid; ++c) if (p.task_begin[(size_t)c + 1] > p.task_begin[c]) d->grid = c + 1;
        if (d->world > 1) d->grid = p.grid;            // every rank launches the same grid (the exit barrier counts CTAs)
    }
    // traceback checkpoints: cp[0] = sink level, then the narrowest level about every DIP_TRACE_T levels, down to level 0
    d->h_cp = choose_checkpoints(p.level_off, DIP_TRACE_T);
    d->M = (int)d->h_cp.size() - 1;
    d->h_aoff.assign((size_t)d->M + 1, 0);
    for (int m = 0; m < d->M; ++m) {
        const int64_t k = p.level_off[d->h_cp[m] + 1] - p.level_off[d->h_cp[m]];
        d->h_aoff[(size_t)m + 1] = d->h_aoff[m] + (int64_t)(p.R + 1) * k * k;
    }
    d->anc_cells = d->h_aoff[d->M];
    d->plan_ms = (float)(now_ms() - t_plan0);
    return true;
}

static int dip_create_device(dg_ctx* ctx, dg_dip* d);

static int dip_create_impl(dg_ctx* ctx, const DipGraphView& g, dg_dip** out, cudaStream_t stream = nullptr, int grid_cap = 0) {
    *out = nullptr;
    DipLimits lim;
    if (int rc = dip_limits(ctx, lim)) return rc;
    std::unique_ptr<dg_dip> d(new dg_dip());
    d->stream = stream ? stream : ctx->stream;
    d->cooperative = stream == nullptr;
    if (!dip_plan_host(g, lim, grid_cap, d.get())) return fail(ctx, DG_ERR_ARG, "dg_dip_create: %s", d->plan.error.c_str());
    if (int rc = dip_create_device(ctx, d.get())) return rc;
    *out = d.release();
    return DG_OK;
}

// Hands the plan's big host arrays (and their page-locked block) back: only once the copies out of them have completed.
static void dip4_release_host(dg_dip* d) {
    DipPlan& p = d->plan;
    p.in_edge.clear(); p.in_edge.shrink_to_fit();
    p.in_dst.clear(); p.in_dst.shrink_to_fit();
    p.masks.clear(); p.masks.shrink_to_fit();
    p.in_off.clear(); p.in_off.shrink_to_fit();
    d->p4.release_arrays();
    if (d->staging) d->staging->release();
    d->release_pending = false;
}

// Device half of the level-program engine: the O(V) tables go up, the O(sum E^2) program is written by the device.
static int dip4_create_device(dg_ctx* ctx, dg_dip* d) {
    DG_CUDA(ctx, cudaSetDevice(ctx->device));
    DipPlan& p = d->plan;
    Plan4& q = d->p4;
    const int L = p.L;
    const double t_up0 = now_ms();
    cudaStream_t s = d->stream;
    {
        const int smem = (int)sweep4_smem_bytes(q.shape.cells(), q.RL, q.shape.slot_bytes, q.shape.nslot);
        for (int c = 0; c < 2; ++c) DG_CUDA(ctx, cudaFuncSetAttribute(sweep4_fn(q.shape.cells(), q.rc, c != 0, q.max_giant > 0), cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        DG_CUDA(ctx, cudaFuncSetAttribute(sweep4_many_fn(q.shape.cells(), q.rc, q.max_giant > 0), cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        // (no shared-memory carveout preference: the generic path's descriptor and spill traffic wants the L1 that two packed
        //  CTAs of ~61 KB leave — with the carveout forced to 100 % the same launch took 636 instead of 416 ms)
    }
    DG_CUDA(ctx, d->level_off.upload(p.level_off.data(), p.level_off.size(), s));
    DG_CUDA(ctx, d->in_off.upload(p.in_off.data(), p.in_off.size(), s));
    DG_CUDA(ctx, d->in_edge.upload(p.in_edge.data(), p.in_edge.size(), s));
    DG_CUDA(ctx, d->lvlW.upload(p.lvlW.data(), p.lvlW.size(), s));
    DG_CUDA(ctx, d->masks.upload(p.masks.data(), p.masks.size(), s));
    DG_CUDA(ctx, d->msrc_off.upload(p.msrc_off.data(), p.msrc_off.size(), s));
    DG_CUDA(ctx, d->mdst_off.upload(p.mdst_off.data(), p.mdst_off.size(), s));
    DG_CUDA(ctx, d->pred_off.upload(q.pred_off.data(), q.pred_off.size(), s));
    DG_CUDA(ctx, d->cp.upload(d->h_cp.data(), d->h_cp.size(), s));
    DG_CUDA(ctx, d->aoff.upload(d->h_aoff.data(), d->h_aoff.size(), s));
    DG_CUDA(ctx, d->v4_dir.upload(q.timed.dir.data(), q.timed.dir.size(), s));
    DG_CUDA(ctx, d->v4_dir_full.upload(q.full.dir.data(), q.full.dir.size(), s));
    DG_CUDA(ctx, d->v4_hdr.upload(q.hdr.data(), q.hdr.size(), s));
    DG_CUDA(ctx, d->v4_prog_off.upload(q.prog_off.data(), q.prog_off.size(), s));
    DG_CUDA(ctx, d->v4_wide.upload(q.timed.wide_list.data(), q.timed.wide_list.size(), s));
    DG_CUDA(ctx, d->v4_wide_full.upload(q.full.wide_list.data(), q.full.wide_list.size(), s));
    DG_CUDA(ctx, d->v4_vslot.upload(q.vslot.data(), q.vslot.size(), s));
    DG_CUDA(ctx, d->v4_dom.upload(q.lvl_dom.data(), q.lvl_dom.size(), s));
    DG_CUDA(ctx, d->v4_tflags.upload(q.tflags.data(), q.tflags.size(), s));
    DG_CUDA(ctx, d->v4_np.upload(q.lvl_np.data(), q.lvl_np.size(), s));
    DG_CUDA(ctx, d->v4_cls.upload(q.cls_list.data(), q.cls_list.size(), s));
    DG_CUDA(ctx, d->v4_vinfo.upload(q.vinfo.data(), q.vinfo.size(), s));
    DG_CUDA(ctx, d->v4_mpre.upload(q.mpre.data(), q.mpre.size(), s));
    DG_CUDA(ctx, d->v4_mpre_off.upload(q.mpre_off.data(), q.mpre_off.size(), s));
    DG_CUDA(ctx, d->v4_n1.upload(q.lvl_n1.data(), q.lvl_n1.size(), s));
    DG_CUDA(ctx, d->v4_m.upload(q.lvl_m.data(), q.lvl_m.size(), s));
    DG_CUDA(ctx, d->v4_z.upload(q.lvl_z.data(), q.lvl_z.size(), s));
    DG_CUDA(ctx, d->v4_dm.upload(q.lvl_dm.data(), q.lvl_dm.size(), s));
    d->h2d_bytes = d->level_off.bytes() + d->in_off.bytes() + d->in_edge.bytes() + d->lvlW.bytes() + d->masks.bytes() + d->msrc_off.bytes() +
                   d->mdst_off.bytes() + d->pred_off.bytes() + d->cp.bytes() + d->aoff.bytes() + d->v4_dir.bytes() + d->v4_dir_full.bytes() +
                   d->v4_hdr.bytes() + d->v4_prog_off.bytes() + d->v4_wide.bytes() + d->v4_wide_full.bytes() + d->v4_vslot.bytes() +
                   d->v4_dom.bytes() + d->v4_tflags.bytes() + d->v4_np.bytes() + d->v4_cls.bytes() + d->v4_vinfo.bytes() + d->v4_mpre.bytes() +
                   d->v4_mpre_off.bytes() + d->v4_n1.bytes() + d->v4_m.bytes() + d->v4_z.bytes() + d->v4_dm.bytes();
    DG_CUDA(ctx, d->v4_prog.alloc((size_t)q.prog_bytes + 16, s));
    DG_CUDA(ctx, d->v4_pred.alloc((size_t)q.pred_elems + 8, s));
    DG_CUDA(ctx, d->v4_sink.alloc((size_t)p.R + 1, s));
    if (q.max_giant > 0) {
        DG_CUDA(ctx, d->v4_gkey.alloc((size_t)GIANT_LIST_MAX * (size_t)q.RL, s));
        DG_CUDA(ctx, d->v4_gcnt.alloc((size_t)GIANT_LIST_MAX, s));
    }
    DG_CUDA(ctx, d->tile0.alloc((size_t)std::max<int64_t>(q.gtile_cells, 1), s));
    DG_CUDA(ctx, d->counter.alloc(4, s));
    DG_CUDA(ctx, d->level_sum.alloc((size_t)L, s));
    DG_CUDA(ctx, d->level_live.alloc((size_t)L, s));
    DG_CUDA(ctx, d->prof.alloc(32, s));
    DG_CUDA(ctx, cudaMemsetAsync(d->prof.p, 0, 32 * 8, s));
    DG_CUDA(ctx, d->anc.alloc((size_t)d->anc_cells, s));
    DG_CUDA(ctx, d->path_cell.alloc((size_t)std::max(d->M, 1), s));
    const size_t cap = (size_t)p.R + 2;
    DG_CUDA(ctx, d->seg_p1.alloc((size_t)std::max(d->M, 1) * 2 * cap, s));
    DG_CUDA(ctx, d->seg_p2.alloc((size_t)std::max(d->M, 1) * 2 * cap, s));
    DG_CUDA(ctx, d->seg_n.alloc((size_t)std::max(d->M, 1) * 4, s));
    DG_CUDA(ctx, d->tout.alloc(1, s));
    DG_CUDA(ctx, d->p1.alloc(2 * cap, s));
    DG_CUDA(ctx, d->p2.alloc(2 * cap, s));
    for (auto& e : d->ev) DG_CUDA(ctx, cudaEventCreate(&e));
    // the program: zero (section padding), then one CTA per transition
    DG_CUDA(ctx, cudaEventRecord(d->ev[0], s));
    DG_CUDA(ctx, cudaMemsetAsync(d->v4_prog.p, 0, (size_t)q.prog_bytes + 16, s));
    DG_CUDA(ctx, d->v4_vup.alloc((size_t)p.V, s));
    vup_fill_kernel<<<ctx->sm_count * 4, 256, 0, s>>>(d->in_off.p, d->in_edge.p, p.V, d->v4_vup.p);
    if (!q.full.wide_list.empty() + q.n_relocate > 0)      // the two padding layers of every cell of the HBM tile (the sweep only writes layers >= 0)
        fill_dead_kernel<<<ctx->sm_count * 2, 256, 0, s>>>(d->tile0.p, (long long)q.gtile_cells);
    Fill4Args fa;
    fa.l0 = 0; fa.l1 = L - 1; fa.level_off = d->level_off.p; fa.in_off = d->in_off.p; fa.in_edge = d->in_edge.p;
    fa.cls_list = d->v4_cls.p; fa.mpre = d->v4_mpre.p; fa.mpre_off = d->v4_mpre_off.p;
    fa.lvl_n1 = d->v4_n1.p; fa.lvl_np = d->v4_np.p; fa.lvl_m = d->v4_m.p; fa.lvl_z = d->v4_z.p; fa.lvl_dm = d->v4_dm.p;
    fa.vslot = d->v4_vslot.p; fa.lvl_dom = d->v4_dom.p; fa.tflags = d->v4_tflags.p; fa.kn = q.shape.kn; fa.hstride = q.hstride;
    fa.lvlW = d->lvlW.p; fa.msrc_off = d->msrc_off.p; fa.mdst_off = d->mdst_off.p; fa.masks = d->masks.p;
    fa.hdr = d->v4_hdr.p; fa.prog_off = d->v4_prog_off.p; fa.prog_base = 0; fa.prog = d->v4_prog.p;
    DG_CUDA(ctx, d->v4_tables.alloc(1, s));
    DG_CUDA(ctx, cudaMemcpyAsync(d->v4_tables.p, &fa, sizeof fa, cudaMemcpyHostToDevice, s));
    prog_fill_kernel<<<std::min(L - 1, ctx->sm_count * 16), FILL4_THREADS, 0, s>>>(fa);
    DG_CUDA(ctx, cudaGetLastError());
    DG_CUDA(ctx, cudaEventRecord(d->ev[1], s));
    if (d->async_create) {
        DG_CUDA(ctx, cudaEventCreateWithFlags(&d->ev_up, cudaEventDisableTiming));
        DG_CUDA(ctx, cudaEventRecord(d->ev_up, s));
        d->release_pending = true;
        d->build_ms = 0.f;               // (ev[0] / ev[1] are recorded again by the run that follows at once)
    } else {
        DG_CUDA(ctx, cudaStreamSynchronize(s));
        DG_CUDA(ctx, cudaEventElapsedTime(&d->build_ms, d->ev[0], d->ev[1]));
    }
    d->upload_ms = (float)(now_ms() - t_up0);
    d->device_bytes = d->level_off.bytes() + d->in_off.bytes() + d->in_edge.bytes() + d->lvlW.bytes() + d->masks.bytes() +
                      d->msrc_off.bytes() + d->mdst_off.bytes() + d->pred_off.bytes() + d->v4_dir.bytes() + d->v4_hdr.bytes() +
                      d->v4_prog_off.bytes() + d->v4_cls.bytes() + d->v4_vinfo.bytes() + d->v4_mpre.bytes() + d->v4_prog.bytes() +
                      d->v4_pred.bytes() + d->tile0.bytes() + d->v4_vup.bytes() + d->v4_gkey.bytes() + d->v4_vslot.bytes() + d->v4_dir_full.bytes() + d->level_sum.bytes() + d->level_live.bytes() +
                      d->anc.bytes() + d->seg_p1.bytes() + d->seg_p2.bytes();
    if (!d->async_create) dip4_release_host(d);
    return DG_OK;
}

// Device half: allocations from the stream-ordered pool and H2D copies on the problem's stream.
static int dip_create_device(dg_ctx* ctx, dg_dip* d) {
    if (d->v4) return dip4_create_device(ctx, d);
    DG_CUDA(ctx, cudaSetDevice(ctx->device));
    DipPlan& p = d->plan;
    const int L = p.L;
    const std::vector<int32_t>& cp = d->h_cp;
    const std::vector<int64_t>& aoff = d->h_aoff;
    const double t_up0 = now_ms();
    cudaStream_t s = d->stream;
    DG_CUDA(ctx, d->tasks.upload(p.tasks.data(), p.tasks.size(), s));
    DG_CUDA(ctx, d->task_begin.upload(p.task_begin.data(), p.task_begin.size(), s));
    DG_CUDA(ctx, d->records.upload(p.records.data(), p.records.size(), s));
    DG_CUDA(ctx, d->delta.alloc((size_t)p.delta_elems, s));
    DG_CUDA(ctx, d->delta_off.upload(p.delta_off.data(), p.delta_off.size(), s));
    DG_CUDA(ctx, d->delta_list.upload(p.delta_list.data(), p.delta_list.size(), s));
    DG_CUDA(ctx, d->level_off.upload(p.level_off.data(), p.level_off.size(), s));
    DG_CUDA(ctx, d->in_off.upload(p.in_off.data(), p.in_off.size(), s));
    DG_CUDA(ctx, d->in_edge.upload(p.in_edge.data(), p.in_edge.size(), s));
    DG_CUDA(ctx, d->in_dst.upload(p.in_dst.data(), p.in_dst.size(), s));
    DG_CUDA(ctx, d->lvlW.upload(p.lvlW.data(), p.lvlW.size(), s));
    DG_CUDA(ctx, d->masks.upload(p.masks.data(), p.masks.size(), s));
    DG_CUDA(ctx, d->msrc_off.upload(p.msrc_off.data(), p.msrc_off.size(), s));
    DG_CUDA(ctx, d->mdst_off.upload(p.mdst_off.data(), p.mdst_off.size(), s));
    DG_CUDA(ctx, d->pred_off.upload(p.pred_off.data(), p.pred_off.size(), s));
    DG_CUDA(ctx, d->cp.upload(cp.data(), cp.size(), s));
    DG_CUDA(ctx, d->aoff.upload(aoff.data(), aoff.size(), s));
    d->h2d_bytes = d->tasks.bytes() + d->task_begin.bytes() + d->records.bytes() + d->delta_off.bytes() + d->delta_list.bytes() +
                   d->level_off.bytes() + d->in_off.bytes() + d->in_edge.bytes() + d->in_dst.bytes() + d->lvlW.bytes() + d->masks.bytes() +
                   d->msrc_off.bytes() + d->mdst_off.bytes() + d->pred_off.bytes() + d->cp.bytes() + d->aoff.bytes();
    // (layers rounded up to whole lane-form chunks: the last chunk loads, but never stores, layers above R)
    const uint64_t widest = (uint64_t)(p.R + d->lane_rc) * (uint64_t)p.kmax * (uint64_t)p.kmax;
    const size_t tile = (size_t)std::max<uint64_t>(widest, (uint64_t)(p.R + 1));
    if (d->world > 1) {          // what the peers map through CUDA IPC cannot come from the stream-ordered pool
        DG_CUDA(ctx, d->tile0.alloc(tile));
        DG_CUDA(ctx, d->tile1.alloc(tile));
        DG_CUDA(ctx, d->pred.alloc((size_t)p.pred_off[L] * (size_t)d->pred_bytes));
        DG_CUDA(ctx, d->counter.alloc(4));
    } else {
        DG_CUDA(ctx, d->tile0.alloc(tile, s));
        DG_CUDA(ctx, d->tile1.alloc(tile, s));
        DG_CUDA(ctx, d->pred.alloc((size_t)p.pred_off[L] * (size_t)d->pred_bytes, s));
        DG_CUDA(ctx, d->counter.alloc(4, s));
    }
    DG_CUDA(ctx, d->level_sum.alloc((size_t)L, s));
    DG_CUDA(ctx, d->level_live.alloc((size_t)L, s));
    DG_CUDA(ctx, d->prof.alloc(32, s));
    DG_CUDA(ctx, d->anc.alloc((size_t)d->anc_cells, s));
    DG_CUDA(ctx, d->path_cell.alloc((size_t)std::max(d->M, 1), s));
    const size_t cap = (size_t)p.R + 2;
    DG_CUDA(ctx, d->seg_p1.alloc((size_t)std::max(d->M, 1) * 2 * cap, s));
    DG_CUDA(ctx, d->seg_p2.alloc((size_t)std::max(d->M, 1) * 2 * cap, s));
    DG_CUDA(ctx, d->seg_n.alloc((size_t)std::max(d->M, 1) * 4, s));
    DG_CUDA(ctx, d->tout.alloc(1, s));
    DG_CUDA(ctx, d->p1.alloc(2 * cap, s));
    DG_CUDA(ctx, d->p2.alloc(2 * cap, s));
    for (auto& e : d->ev) DG_CUDA(ctx, cudaEventCreate(&e));
    const double t_issued = now_ms();
    DG_CUDA(ctx, cudaStreamSynchronize(s));
    d->upload_ms = (float)(now_ms() - t_up0);
    if (getenv("DG_TIMING_UPLOAD")) fprintf(stderr, "dip_create_device: allocations + copies issued in %.2f ms, synchronized after %.2f ms\n", t_issued - t_up0, now_ms() - t_up0);
    d->device_bytes = d->tasks.bytes() + d->task_begin.bytes() + d->records.bytes() + d->delta.bytes() + d->delta_off.bytes() +
                      d->delta_list.bytes() + d->level_off.bytes() + d->in_off.bytes() + d->in_edge.bytes() + d->in_dst.bytes() +
                      d->lvlW.bytes() + d->masks.bytes() + d->msrc_off.bytes() + d->mdst_off.bytes() + d->pred_off.bytes() +
                      d->tile0.bytes() + d->tile1.bytes() + d->pred.bytes() + d->level_sum.bytes() + d->level_live.bytes() +
                      d->anc.bytes() + d->seg_p1.bytes() + d->seg_p2.bytes();
    // the big host arrays are no longer needed
    p.in_edge.clear(); p.in_edge.shrink_to_fit();
    p.in_dst.clear(); p.in_dst.shrink_to_fit();
    p.records.clear(); p.records.shrink_to_fit();
    p.masks.clear(); p.masks.shrink_to_fit();
    p.in_off.clear(); p.in_off.shrink_to_fit();
    p.tasks.clear(); p.tasks.shrink_to_fit();
    if (d->staging) d->staging->release();     // (the copies were synchronized above)
    return DG_OK;
}

static int dip_run_pre(dg_ctx* ctx, dg_dip* d, bool check) {
    const DipPlan& p = d->plan;
    cudaStream_t s = d->stream;
    if (d->world > 1) {
        // counters and level 0 were reset by dg_dip_shard_arm, before the ranks' host barrier: a peer may arrive on
        // this GPU's counter before this launch
        if (!d->attached || !d->armed) return fail(ctx, DG_ERR_ARG, "dg_dip_run: sharded problem needs dg_dip_ipc_attach and dg_dip_shard_arm first");
        if (check) return fail(ctx, DG_ERR_ARG, "dg_dip_run: level checksums are not available on a sharded problem");
        d->armed = false;
    } else if (d->v4) {
        DG_CUDA(ctx, cudaMemsetAsync(d->counter.p, 0, 4 * sizeof(unsigned int), s));            // (level 0 is set up by the kernel)
        if (d->v4_gkey.p) {          // (clean after a complete run; a run that timed out may have left slices behind)
            DG_CUDA(ctx, cudaMemsetAsync(d->v4_gkey.p, 0, d->v4_gkey.bytes(), s));
            DG_CUDA(ctx, cudaMemsetAsync(d->v4_gcnt.p, 0, d->v4_gcnt.bytes(), s));
        }
    } else {
        DG_CUDA(ctx, cudaMemsetAsync(d->counter.p, 0, 2 * sizeof(unsigned int), s));
        DG_CUDA(ctx, cudaMemsetAsync(d->tile0.p, 0, (size_t)(p.R + 1) * sizeof(int32_t), s));   // dp_cur.assign(R+1, {0,0}) :535
    }
    if (check) {
        std::vector<unsigned long long> basis((size_t)p.L, FOLD_BASIS);
        DG_CUDA(ctx, cudaMemcpyAsync(d->level_sum.p, basis.data(), basis.size() * 8, cudaMemcpyHostToDevice, s));
        DG_CUDA(ctx, cudaMemsetAsync(d->level_live.p, 0, (size_t)p.L * 8, s));
        DG_CUDA(ctx, cudaStreamSynchronize(s));   // basis is a stack-lifetime staging buffer
    }
    d->launches = 0;
    d->group_trace_ms = 0.f;
    DG_CUDA(ctx, cudaEventRecord(d->ev[0], s));
    if (!d->v4 && !p.delta_list.empty()) {
        DeltaArgs da;
        da.delta_list = d->delta_list.p; da.n_list = (int32_t)p.delta_list.size(); da.level_off = d->level_off.p;
        da.in_off = d->in_off.p; da.in_edge = d->in_edge.p; da.in_dst = d->in_dst.p; da.lvlW = d->lvlW.p;
        da.msrc_off = d->msrc_off.p; da.mdst_off = d->mdst_off.p; da.masks = d->masks.p; da.delta_off = d->delta_off.p;
        da.delta = d->delta.p;
        const int blocks = (int)std::min<size_t>(p.delta_list.size(), (size_t)ctx->sm_count * 16);
        dip_delta_kernel<<<blocks, 256, 0, s>>>(da);
        ++d->launches;
        DG_CUDA(ctx, cudaGetLastError());
    }
    DG_CUDA(ctx, cudaEventRecord(d->ev[1], s));
    return DG_OK;
}

static void fill_sweep_args(const dg_dip* d, SweepArgs& a) {
    const DipPlan& p = d->plan;
    a.tasks = d->tasks.p; a.task_begin = d->task_begin.p; a.records = d->records.p; a.delta = d->delta.p;
    a.delta_off = d->delta_off.p; a.level_off = d->level_off.p; a.in_off = d->in_off.p; a.in_edge = d->in_edge.p;
    a.lvlW = d->lvlW.p; a.msrc_off = d->msrc_off.p; a.mdst_off = d->mdst_off.p; a.masks = d->masks.p;
    a.tile0 = d->tile0.p; a.tile1 = d->tile1.p; a.pred = d->pred.p; a.counter = d->counter.p;
    a.level_sum = d->level_sum.p; a.level_live = d->level_live.p;
    a.prof = d->want_prof ? d->prof.p : nullptr;
    a.R = p.R; a.shift = d->shift;
    a.world = d->world; a.rank = d->rank; a.exit_target = (uint32_t)(d->world * d->grid);
    a.timeout_ns = 10000ull * 1000000ull;
    if (const char* e = getenv("DG_SHARD_TIMEOUT_MS")) a.timeout_ns = (unsigned long long)std::max(1, atoi(e)) * 1000000ull;
    for (int q = 0; q < DG_MAX_PEERS; ++q) {
        a.peer_tile0[q] = (int32_t*)d->peer[0][q]; a.peer_tile1[q] = (int32_t*)d->peer[1][q];
        a.peer_pred[q] = (uint8_t*)d->peer[2][q]; a.peer_counter[q] = (unsigned int*)d->peer[3][q];
    }
}

static void fill_sweep4_args(const dg_dip* d, Sweep4Args& a, bool check) {
    const DipPlan& p = d->plan;
    const Plan4& q = d->p4;
    const Plan4Dir& dir = check ? q.full : q.timed;         // the checksum variant folds every level, idle transitions included
    a.dir = check ? d->v4_dir_full.p : d->v4_dir.p; a.wide_list = check ? d->v4_wide_full.p : d->v4_wide.p;
    a.n_trans = dir.n; a.n_wide = (int32_t)dir.wide_list.size();
    a.prog = d->v4_prog.p; a.gtile = d->tile0.p; a.gcs = (long long)q.gcs;
    a.pred = d->v4_pred.p; a.counter = d->counter.p; a.level_sum = d->level_sum.p; a.level_live = d->level_live.p;
    a.sink = d->v4_sink.p; a.R = p.R; a.nchunk = q.nchunk; a.grid = d->grid; a.ncw = d->v4_ncw;
    a.slot_bytes = q.shape.slot_bytes; a.nslot = q.shape.nslot; a.m_nchunk = make_magic((uint32_t)q.nchunk);
    a.last_smem = q.last_smem ? 1 : 0; a.sink_cell = q.sink_cell;
    a.use_l1 = getenv("DG_V4_NO_L1") ? 0 : 1;
    a.chk = check ? d->v4_tables.p : nullptr;
    a.final_target = dir.final_target;
    a.prof = d->want_prof ? d->prof.p : nullptr;
    a.giant_key = d->v4_gkey.p; a.giant_cnt = d->v4_gcnt.p;
    a.timeout_ns = 10000ull * 1000000ull;
    if (const char* e = getenv("DG_SHARD_TIMEOUT_MS")) a.timeout_ns = (unsigned long long)std::max(1, atoi(e)) * 1000000ull;
}

static void fill_trace_args(const dg_dip* d, TraceArgs& ta) {
    const DipPlan& p = d->plan;
    TraceView& v = ta.v;
    v.L = p.L; v.R = p.R; v.level_off = d->level_off.p; v.in_off = d->in_off.p; v.in_edge = d->in_edge.p;
    v.lvlW = d->lvlW.p; v.msrc_off = d->msrc_off.p; v.mdst_off = d->mdst_off.p; v.masks = d->masks.p;
    v.pred_off = d->pred_off.p;
    ta.pred = d->pred.p; ta.cp = d->cp.p; ta.aoff = d->aoff.p; ta.M = d->M; ta.anc = d->anc.p; ta.path_cell = d->path_cell.p;
    ta.seg_p1 = d->seg_p1.p; ta.seg_p2 = d->seg_p2.p; ta.seg_n = d->seg_n.p;
    ta.sink_tile = ((p.L - 1) & 1) ? d->tile1.p : d->tile0.p;
    ta.sink_v4 = nullptr;
    if (d->v4) {
        ta.pred = d->v4_pred.p; ta.sink_v4 = d->v4_sink.p;
        v.vinfo = d->v4_vinfo.p; v.lvl_n1 = d->v4_n1.p; v.lvl_m = d->v4_m.p; v.RL = d->p4.RL;
        v.vup = getenv("DG_NO_VUP") ? nullptr : d->v4_vup.p;
    }
    ta.cap = p.R + 2; ta.shift = d->shift; ta.out = d->tout.p; ta.p1 = d->p1.p; ta.p2 = d->p2.p;
}

template <class PredT>
static int dip_run_post(dg_ctx* ctx, dg_dip* d, bool check) {
    cudaStream_t s = d->stream;
    DG_CUDA(ctx, cudaEventRecord(d->ev[2], s));
    TraceArgs ta;
    fill_trace_args(d, ta);
    if (d->anc_cells > 0) {
        const unsigned blocks = (unsigned)((d->anc_cells + 255) / 256);
        dip_anc_kernel<PredT><<<blocks, 256, 0, s>>>(ta);
        ++d->launches;
    }
    dip_hop_kernel<<<1, 32, 0, s>>>(ta);
    ++d->launches;
    if (d->M > 0) {
        dip_seg_kernel<PredT><<<(d->M + 127) / 128, 128, 0, s>>>(ta);
        ++d->launches;
    }
    dip_merge_kernel<<<1, 32, 0, s>>>(ta);
    ++d->launches;
    DG_CUDA(ctx, cudaGetLastError());
    DG_CUDA(ctx, cudaEventRecord(d->ev[3], s));
    d->ran = true; d->checks = check;
    return DG_OK;
}


template <class PredT>
static int dip_run_impl(dg_ctx* ctx, dg_dip* d, bool check) {
    if (int rc = dip_run_pre(ctx, d, check)) return rc;
    d->fused_ms = 0.f; d->group_trace_ms = 0.f;
    const DipPlan& p = d->plan;
    cudaStream_t s = d->stream;
    if (d->v4) {
        Sweep4Args a4;
        fill_sweep4_args(d, a4, check);
        void* args[] = {(void*)&a4};
        const Plan4& q = d->p4;
        const void* fn = sweep4_fn(q.shape.cells(), q.rc, check, q.max_giant > 0);
        const size_t smem = sweep4_smem_bytes(q.shape.cells(), q.RL, q.shape.slot_bytes, q.shape.nslot);
        const dim3 block((unsigned)(d->v4_ncw + 1) * 32u);
        if (d->cooperative) DG_CUDA(ctx, cudaLaunchCooperativeKernel(fn, dim3(d->grid), block, args, smem, s));
        else DG_CUDA(ctx, cudaLaunchKernel(fn, dim3(d->grid), block, args, smem, s));
        ++d->launches;
        return dip_run_post<PredT>(ctx, d, check);
    }
    SweepArgs a;
    fill_sweep_args(d, a);
    if (p.L > 1) {
        void* args[] = {(void*)&a};
        const void* fn = sweep_fn(sizeof(PredT) == 4, check, d->want_prof);
        // A lone problem is launched cooperatively (the driver guarantees that its CTAs are co-resident, which the
        // counter barrier needs).  Batch slots use plain launches: B200 runs at most 8 cooperative grids at a time
        // (measured: 12 slots took two waves), and the batch scheduler already keeps slots x CTAs within the SM
        // count with one CTA per SM (197 KB of shared memory each), so every grid becomes resident as a whole.
        if (d->cooperative) DG_CUDA(ctx, cudaLaunchCooperativeKernel(fn, dim3(d->grid), dim3(DIP_THREADS), args, DIP_SMEM_BYTES, s));
        else DG_CUDA(ctx, cudaLaunchKernel(fn, dim3(d->grid), dim3(DIP_THREADS), args, DIP_SMEM_BYTES, s));
        ++d->launches;
    }
    return dip_run_post<PredT>(ctx, d, check);
}

static int batch_stream(dg_ctx* ctx, int slot, cudaStream_t* out) {
    while ((int)ctx->batch_streams.size() <= slot) {
        cudaStream_t s = nullptr;
        DG_CUDA(ctx, cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
        ctx->batch_streams.push_back(s);
    }
    *out = ctx->batch_streams[(size_t)slot];
    return DG_OK;
}

// Events of one dg_dip_run_many call: destroyed on every way out (the DG_CUDA early returns included).
namespace {
struct EventBag {
    std::vector<cudaEvent_t> all;
    cudaError_t make(cudaEvent_t* e, unsigned flags = cudaEventDefault) {
        const cudaError_t r = cudaEventCreateWithFlags(e, flags);
        if (r == cudaSuccess) all.push_back(*e);
        return r;
    }
    ~EventBag() { for (cudaEvent_t e : all) cudaEventDestroy(e); }
};
}  // namespace

extern "C" {

int dg_dip_create(dg_ctx* ctx, int32_t n_levels, const int32_t* level_off, const int64_t* adj_off,
                  const int32_t* adj_dst, const uint8_t* adj_w, const int64_t* col_off, const int32_t* col_val,
                  const uint8_t* colour_is_hom, int32_t n_colours, int32_t R, dg_dip** out) {
    if (!ctx || !out) return DG_ERR_ARG;
    DipGraphView g;
    g.n_levels = n_levels; g.level_off = level_off; g.adj_off = adj_off; g.adj_dst = adj_dst; g.adj_w = adj_w;
    g.col_off = col_off; g.col_val = col_val; g.colour_is_hom = colour_is_hom; g.n_colours = n_colours; g.R = R;
    return dip_create_impl(ctx, g, out);
}

int dg_dip_run(dg_ctx* ctx, dg_dip* d, uint32_t flags) {
    if (!ctx || !d) return DG_ERR_ARG;
    DG_CUDA(ctx, cudaSetDevice(ctx->device));
    const bool check = (flags & 1u) != 0;
    d->want_prof = (flags & 2u) != 0;
    return d->pred_bytes == 2 ? dip_run_impl<uint16_t>(ctx, d, check) : dip_run_impl<uint32_t>(ctx, d, check);
}

int dg_dip_result(dg_ctx* ctx, dg_dip* d, int32_t* sink_value, int32_t* sink_s_het, int32_t* p1_edges, int32_t* n_p1,
                  int32_t* p2_edges, int32_t* n_p2) {
    if (!ctx || !d || !d->ran) return fail(ctx, DG_ERR_ARG, "dg_dip_result: dg_dip_run has not been called");
    DG_CUDA(ctx, cudaSetDevice(ctx->device));
    const int cap = d->plan.R + 2;
    TraceOut t;
    std::vector<int32_t> a((size_t)2 * cap), b((size_t)2 * cap);
    cudaStream_t s = d->stream;
    unsigned int shard_err = 0;
    if (d->world > 1) DG_CUDA(ctx, cudaMemcpyAsync(&shard_err, d->counter.p + 2, sizeof shard_err, cudaMemcpyDeviceToHost, s));
    DG_CUDA(ctx, cudaMemcpyAsync(&t, d->tout.p, sizeof t, cudaMemcpyDeviceToHost, s));
    DG_CUDA(ctx, cudaMemcpyAsync(a.data(), d->p1.p, a.size() * 4, cudaMemcpyDeviceToHost, s));
    DG_CUDA(ctx, cudaMemcpyAsync(b.data(), d->p2.p, b.size() * 4, cudaMemcpyDeviceToHost, s));
    DG_CUDA(ctx, cudaStreamSynchronize(s));
    if (shard_err)
        return fail(ctx, DG_ERR_CUDA, "dg_dip_result: row-sharded sweep of rank %d timed out at a cross-GPU barrier (%s %u): a peer is not running",
                    d->rank, shard_err == 0x7FFFFFFFu ? "exit barrier, code" : "level", shard_err == 0x7FFFFFFFu ? shard_err : shard_err - 1);
    if (d->group_trace_ms > 0.f) { d->delta_ms = 0.f; d->sweep_ms = d->fused_ms; d->trace_ms = d->group_trace_ms; }
    else {
        DG_CUDA(ctx, cudaEventElapsedTime(&d->delta_ms, d->ev[0], d->ev[1]));
        DG_CUDA(ctx, cudaEventElapsedTime(&d->sweep_ms, d->ev[1], d->ev[2]));
        if (d->fused_ms > 0.f) d->sweep_ms = d->fused_ms;      // (its own events also cover the wait for the other problems' pair scores)
        DG_CUDA(ctx, cudaEventElapsedTime(&d->trace_ms, d->ev[2], d->ev[3]));
    }
    if (t.rc == -2) return fail(ctx, DG_ERR_CAPACITY, "dg_dip_result: more than R+2 recorded edges on a path");
    if (sink_value) *sink_value = t.value;
    if (sink_s_het) *sink_s_het = t.s_het;
    if (n_p1) *n_p1 = t.n1;
    if (n_p2) *n_p2 = t.n2;
    if (p1_edges) for (int x = 0; x < t.n1; ++x) { p1_edges[2 * x] = a[2 * (t.n1 - 1 - x)]; p1_edges[2 * x + 1] = a[2 * (t.n1 - 1 - x) + 1]; }
    if (p2_edges) for (int x = 0; x < t.n2; ++x) { p2_edges[2 * x] = b[2 * (t.n2 - 1 - x)]; p2_edges[2 * x + 1] = b[2 * (t.n2 - 1 - x) + 1]; }
    return DG_OK;
}

int dg_dip_checksums(dg_ctx* ctx, dg_dip* d, uint64_t* level_checksum, uint64_t* level_live) {
    if (!ctx || !d || !d->ran || !d->checks) return fail(ctx, DG_ERR_ARG, "dg_dip_checksums: run with flags bit0 first");
    DG_CUDA(ctx, cudaSetDevice(ctx->device));
    DG_CUDA(ctx, cudaStreamSynchronize(d->stream));
    DG_CUDA(ctx, cudaMemcpy(level_checksum, d->level_sum.p, (size_t)d->plan.L * 8, cudaMemcpyDeviceToHost));
    DG_CUDA(ctx, cudaMemcpy(level_live, d->level_live.p, (size_t)d->plan.L * 8, cudaMemcpyDeviceToHost));
    return DG_OK;
}

int dg_dip_stats(dg_ctx* ctx, dg_dip* d, dg_dip_stats_t* out) {
    if (!d || !out) return DG_ERR_ARG;
    memset(out, 0, sizeof *out);
    out->cell_updates = d->plan.cell_updates; out->cells = d->plan.cells; out->algo_bytes = d->plan.algo_bytes;
    out->device_bytes = d->device_bytes;
    out->n_levels = d->plan.L; out->n_vertices = d->plan.V; out->max_width = d->plan.kmax;
    out->max_indegree = d->plan.max_indeg; out->mask_words_max = d->plan.Wmax; out->grid_ctas = d->grid;
    out->pred_bytes = d->pred_bytes; out->launches = d->launches;
    out->sweep_ms = d->sweep_ms; out->traceback_ms = d->trace_ms;
    out->delta_ms = d->delta_ms; out->plan_ms = d->plan_ms; out->upload_ms = d->upload_ms;
    out->n_narrow = (int32_t)d->plan.n_narrow; out->n_wide = (int32_t)d->plan.n_wide;
    out->n_tasks = d->plan.task_begin.empty() ? 0 : (int64_t)d->plan.task_begin.back();
    out->delta_bytes = (uint64_t)d->plan.delta_elems * 2;
    out->engine = d->v4 ? 4 : 3;
    out->h2d_bytes = d->h2d_bytes;
    if (d->v4) {
        out->n_narrow = (int32_t)d->p4.n_smem_trans; out->n_wide = (int32_t)d->p4.timed.wide_list.size();
        out->n_tasks = (int64_t)d->p4.timed.n;
        out->cells_written = d->p4.cells_written * (uint64_t)(d->plan.R + 1); out->n_relocate = (int32_t)d->p4.n_relocate;
        out->prog_bytes = d->p4.prog_bytes; out->code_bytes = (uint64_t)d->p4.pred_elems * 2;
        out->build_ms = d->build_ms;
    } else {
        out->code_bytes = (uint64_t)d->plan.pred_off[(size_t)d->plan.L] * (uint64_t)d->pred_bytes;
    }
    (void)ctx;
    return DG_OK;
}

int dg_dip_profile(dg_ctx* ctx, dg_dip* d, uint64_t* out24) {
    if (!ctx || !d || !d->ran || !d->want_prof) return fail(ctx, DG_ERR_ARG, "dg_dip_profile: run with flags bit1 first");
    DG_CUDA(ctx, cudaSetDevice(ctx->device));
    DG_CUDA(ctx, cudaStreamSynchronize(d->stream));
    memset(out24, 0, 24 * 8);
    DG_CUDA(ctx, cudaMemcpy(out24, d->prof.p, 24 * 8, cudaMemcpyDeviceToHost));
    return DG_OK;
}

int dg_dip_debug_program(dg_ctx* ctx, dg_dip* d, uint8_t* out, uint64_t cap, uint64_t* bytes) {
    if (!ctx || !d || !bytes) return DG_ERR_ARG;
    *bytes = d->v4 ? d->p4.prog_bytes : 0;
    if (!d->v4 || !out) return DG_OK;
    if (cap < d->p4.prog_bytes) return fail(ctx, DG_ERR_CAPACITY, "dg_dip_debug_program: %llu bytes needed", (unsigned long long)d->p4.prog_bytes);
    DG_CUDA(ctx, cudaSetDevice(ctx->device));
    DG_CUDA(ctx, cudaStreamSynchronize(d->stream));
    DG_CUDA(ctx, cudaMemcpy(out, d->v4_prog.p, (size_t)d->p4.prog_bytes, cudaMemcpyDeviceToHost));
    return DG_OK;
}

void dg_dip_destroy(dg_ctx* ctx, dg_dip* d) {
    if (ctx) cudaSetDevice(ctx->device);
    delete d;
}

int dg_dp_diploid(dg_ctx* ctx, int32_t n_levels, const int32_t* level_off, const int64_t* adj_off,
                  const int32_t* adj_dst, const uint8_t* adj_w, const int64_t* col_off, const int32_t* col_val,
                  const uint8_t* colour_is_hom, int32_t n_colours, int32_t R, int32_t* sink_value,
                  int32_t* sink_s_het, int32_t* p1_edges, int32_t* n_p1, int32_t* p2_edges, int32_t* n_p2) {
    dg_dip* d = nullptr;
    const bool timing = getenv("DG_TIMING") != nullptr;     // diagnostics: phase times of the one-shot call on stderr
    const double t0 = now_ms();
    int rc = dg_dip_create(ctx, n_levels, level_off, adj_off, adj_dst, adj_w, col_off, col_val, colour_is_hom, n_colours, R, &d);
    if (rc) return rc;
    const double t1 = now_ms();
    rc = dg_dip_run(ctx, d, 0);
    const double t2 = now_ms();
    if (!rc) rc = dg_dip_result(ctx, d, sink_value, sink_s_het, p1_edges, n_p1, p2_edges, n_p2);
    const double t3 = now_ms();
    const float plan_ms = d->plan_ms, upload_ms = d->upload_ms, sweep_ms = d->sweep_ms, trace_ms = d->trace_ms;
    dg_dip_destroy(ctx, d);
    if (timing)
        fprintf(stderr, "dg_dp_diploid: create %.1f ms (plan %.1f, alloc+upload %.1f) launch %.1f wait+result %.1f (sweep %.1f trace %.1f) destroy %.1f\n",
                t1 - t0, plan_ms, upload_ms, t2 - t1, t3 - t2, sweep_ms, trace_ms, now_ms() - t3);
    return rc;
}

int dg_dp_diploid_batch(dg_ctx* ctx, int32_t n, const dg_dip_input_t* in, dg_dip_output_t* out, int32_t max_concurrent,
                        int32_t ctas_per_sample) {
    if (!ctx || n < 0 || (n > 0 && (!in || !out))) return DG_ERR_ARG;
    if (n == 0) return DG_OK;
    DipLimits lim;
    if (int r = dip_limits(ctx, lim)) return r;
    const int grid = ctas_per_sample > 0 ? ctas_per_sample : 4;
    int K = std::max(1, std::min(ctx->sm_count / grid, 32));    // slots x CTAs must fit the SMs (plain launches, spin barrier); 32 work queues
    if (max_concurrent > 0) K = std::min(K, (int)max_concurrent);
    K = std::min(K, (int)n);
    { cudaStream_t last = nullptr; if (int r = batch_stream(ctx, K - 1, &last)) return r; }

    // planning (host only) runs ahead on a few worker threads, each with its share of the cores; this thread
    // uploads, launches and collects in sample order
    int hw = (int)std::max(1u, std::thread::hardware_concurrency());
    if (const char* e = getenv("DG_HOST_THREADS")) hw = std::max(1, atoi(e));   // this process's share of the host cores (one process per GPU)
    // Eight planners of two threads.  Measured on the GPU box (22 MHC_4 samples, 16 cores, ms per call): 8 x 2 threads 310 / 309 /
    // 309; 11 x 1 thread 327 / 325 / 346 / 318 / 545; 16 x 1 thread 299 / 632 / 301 — more planners keep more page-locked plan
    // blocks in flight, and a call that has to pin another one pays 0.1-0.3 s for it (DG_PLAN_WORKERS overrides).
    // With fewer cores than that (several ranks sharing a host: DG_HOST_THREADS) the cores go to planners first, threads
    // second: the planner's OpenMP regions scale poorly (MHC_4: 1 / 2 / 4 threads 70 / 118 / 62 ms here, 2 / 16 threads 40 / 42 ms
    // on the GPU box), sample-parallelism does.
    int W = std::max(1, std::min({(int)n, 8, hw}));
    if (const char* e = getenv("DG_PLAN_WORKERS")) W = std::max(1, std::min({atoi(e), (int)n, hw}));
    const int lookahead = W + 2;
    const bool pinned = !getenv("DG_NO_PINNED_PLAN");
    const bool timing = getenv("DG_TIMING") != nullptr;
    const double t_batch0 = now_ms();
    std::vector<std::unique_ptr<dg_dip>> planned((size_t)n);
    std::vector<int> state((size_t)n, 0);                       // 0 pending, 1 planned, -1 failed
    std::vector<std::string> perr((size_t)n);
    std::mutex mu;
    std::condition_variable cv;
    int next = 0, consumed = 0;
    auto worker = [&]() {
        for (;;) {
            int i;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return next >= n || next < consumed + lookahead; });
                if (next >= n) return;
                i = next++;
            }
#if defined(_OPENMP)
            // (a fixed share: handing the idle cores of the last wave to its planners oversubscribes while the wave before
            // still runs — OpenMP's waiting threads spin — and cost 90 ms per 22-sample call)
            omp_set_num_threads(std::max(1, hw / W));
#endif
            const dg_dip_input_t& x = in[i];
            // the plan is written straight into page-locked memory: about 60 bytes per vertex on the pangenome
            // panels (an array that does not fit any more goes to the heap and is copied from there)
            std::unique_ptr<PlanStaging> st;
            if (pinned && x.n_levels > 0 && x.level_off) {
                st.reset(new PlanStaging());
                const size_t V = (size_t)x.level_off[x.n_levels];
                st->acquire(ctx, ((size_t)120 * V + ((size_t)6 << 20) + ((size_t)1 << 20) - 1) >> 20 << 20);
            }
            std::unique_ptr<dg_dip> d(new dg_dip(std::move(st)));
            d->stream = ctx->batch_streams[(size_t)(i % K)];
            d->cooperative = false;
            DipGraphView g;
            g.n_levels = x.n_levels; g.level_off = x.level_off; g.adj_off = x.adj_off; g.adj_dst = x.adj_dst; g.adj_w = x.adj_w;
            g.col_off = x.col_off; g.col_val = x.col_val; g.colour_is_hom = x.colour_is_hom; g.n_colours = x.n_colours; g.R = x.R;
            bool ok = x.R + 2 <= DG_BATCH_MAX_EDGES;
            std::string err = ok ? "" : "R + 2 > DG_BATCH_MAX_EDGES";
            if (ok) { ok = dip_plan_host(g, lim, grid, d.get()); if (!ok) err = d->plan.error; }
            if (timing) fprintf(stderr, "batch: sample %d planned at %.1f ms (plan %.1f ms)\n", i, now_ms() - t_batch0, d->plan_ms);
            {
                std::lock_guard<std::mutex> lk(mu);
                if (ok) planned[(size_t)i] = std::move(d); else perr[(size_t)i] = err;
                state[(size_t)i] = ok ? 1 : -1;
            }
            cv.notify_all();
        }
    };
    std::vector<std::thread> pool;
    for (int w = 0; w < W; ++w) pool.emplace_back(worker);

    std::vector<dg_dip*> slot((size_t)K, nullptr);
    std::vector<int32_t> owner((size_t)K, -1);
    int rc = DG_OK;
    std::vector<dg_dip*> uploading;          // launched, copies out of the plan's page-locked block possibly still in flight
    auto collect = [&](int k) {
        if (!slot[(size_t)k]) return;
        for (size_t x = 0; x < uploading.size(); ++x)
            if (uploading[x] == slot[(size_t)k]) { uploading.erase(uploading.begin() + (long)x); break; }      // (its destructor waits and releases)
        dg_dip_output_t& o = out[owner[(size_t)k]];
        const int r = dg_dip_result(ctx, slot[(size_t)k], &o.sink_value, &o.sink_s_het, o.p1_edges, &o.n_p1, o.p2_edges, &o.n_p2);
        o.status = r;
        if (r && !rc) rc = r;
        dg_dip_destroy(ctx, slot[(size_t)k]);
        slot[(size_t)k] = nullptr;
    };
    for (int32_t i = 0; i < n; ++i) {
        std::unique_ptr<dg_dip> d;
        int st;
        {
            std::unique_lock<std::mutex> lk(mu);
            cv.wait(lk, [&] { return state[(size_t)i] != 0; });
            st = state[(size_t)i];
            d = std::move(planned[(size_t)i]);
            consumed = i + 1;
        }
        cv.notify_all();
        memset(&out[i], 0, sizeof out[i]);
        if (st < 0) {
            out[i].status = fail(ctx, perr[(size_t)i] == "R + 2 > DG_BATCH_MAX_EDGES" ? DG_ERR_CAPACITY : DG_ERR_ARG,
                                 "dg_dp_diploid_batch: sample %d: %s", (int)i, perr[(size_t)i].c_str());
            if (!rc) rc = out[i].status;
            continue;
        }
        const int k = i % K;
        collect(k);                               // the slot's previous sample (its kernels ran while the host planned others)
        // Admission by memory: a wide-panel sample needs gigabytes (programs + codes); wait for earlier samples to finish
        // (oldest first) until this one fits, instead of failing its allocation.
        if (d->v4) {
            const size_t need = (size_t)d->p4.prog_bytes + (size_t)d->p4.pred_elems * 2 + (size_t)d->p4.gtile_cells * 4 +
                                (size_t)d->plan.V * 128 + ((size_t)256 << 20);
            for (int back = K - 1; back >= 1; --back) {
                size_t free_b = 0, total_b = 0;
                if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess || free_b >= need) break;
                const int kk = (int)((i + K - back) % K);          // slots in launch order, oldest first
                if (slot[(size_t)kk]) { collect(kk); dg_release_cached_memory(ctx); }
            }
        }
        // (the level-program engine's device half does not wait for its copies here: 22 x 5 ms of waiting were a third of the
        //  call; the page-locked plan block goes back to the pool when the copies are through)
        for (size_t x = 0; x < uploading.size();) {
            if (cudaEventQuery(uploading[x]->ev_up) == cudaSuccess) { dip4_release_host(uploading[x]); uploading.erase(uploading.begin() + (long)x); }
            else ++x;
        }
        d->async_create = d->v4 && d->staging != nullptr;
        int r = dip_create_device(ctx, d.get());     // (uploads stay on this thread: a pageable H2D issued by a worker on a
                                                     //  busy slot stream blocks that worker until the slot's sweep ends)
        if (!r) r = dg_dip_run(ctx, d.get(), 0);
        if (timing) fprintf(stderr, "batch: sample %d launched at %.1f ms (alloc+upload %.1f ms)\n", (int)i, now_ms() - t_batch0, d->upload_ms);
        if (r) { out[i].status = r; if (!rc) rc = r; continue; }
        if (d->release_pending) uploading.push_back(d.get());
        slot[(size_t)k] = d.release(); owner[(size_t)k] = i;
    }
    for (std::thread& t : pool) t.join();
    for (int k = 0; k < K; ++k) collect(k);
    if (timing) fprintf(stderr, "batch: %d samples done at %.1f ms\n", (int)n, now_ms() - t_batch0);
    return rc;
}

// ---- row-sharded diploid DP over several GPUs of one node (one process per GPU) --------------------------------
// Every rank plans the same graph for world x ctas global CTAs and keeps its share of the wide transitions' rows;
// the sweep kernels of all ranks run at the same time and exchange rows and barrier arrivals through NVLink peer
// mappings (no NCCL call on the data path: the exchange is fused into the sweep kernel).
int dg_dip_create_sharded(dg_ctx* ctx, int32_t n_levels, const int32_t* level_off, const int64_t* adj_off,
                          const int32_t* adj_dst, const uint8_t* adj_w, const int64_t* col_off, const int32_t* col_val,
                          const uint8_t* colour_is_hom, int32_t n_colours, int32_t R, int32_t rank, int32_t world,
                          int32_t ctas, dg_dip** out) {
    if (!ctx || !out || world < 1 || world > DG_MAX_PEERS || rank < 0 || rank >= world) return DG_ERR_ARG;
    *out = nullptr;
    DipLimits lim;
    if (int rc = dip_limits(ctx, lim)) return rc;
    DipGraphView g;
    g.n_levels = n_levels; g.level_off = level_off; g.adj_off = adj_off; g.adj_dst = adj_dst; g.adj_w = adj_w;
    g.col_off = col_off; g.col_val = col_val; g.colour_is_hom = colour_is_hom; g.n_colours = n_colours; g.R = R;
    std::unique_ptr<dg_dip> d(new dg_dip());
    d->stream = ctx->stream;
    d->cooperative = true;
    d->world = world; d->rank = rank;
    if (!dip_plan_host(g, lim, ctas > 0 ? ctas : lim.max_grid, d.get())) return fail(ctx, DG_ERR_ARG, "dg_dip_create_sharded: %s", d->plan.error.c_str());
    if (int rc = dip_create_device(ctx, d.get())) return rc;
    d->peer[0][rank] = d->tile0.p; d->peer[1][rank] = d->tile1.p; d->peer[2][rank] = d->pred.p; d->peer[3][rank] = d->counter.p;
    *out = d.release();
    return DG_OK;
}

int dg_dip_ipc_export(dg_ctx* ctx, dg_dip* d, uint8_t* handles) {
    if (!ctx || !d || !handles) return DG_ERR_ARG;
    DG_CUDA(ctx, cudaSetDevice(ctx->device));
    static_assert(sizeof(cudaIpcMemHandle_t) == DG_IPC_HANDLE_BYTES, "handle size");
    void* ptrs[4] = {d->tile0.p, d->tile1.p, d->pred.p, d->counter.p};
    for (int a = 0; a < 4; ++a) {
        cudaIpcMemHandle_t h;
        DG_CUDA(ctx, cudaIpcGetMemHandle(&h, ptrs[a]));
        memcpy(handles + (size_t)a * DG_IPC_HANDLE_BYTES, &h, sizeof h);
    }
    return DG_OK;
}

int dg_dip_ipc_attach(dg_ctx* ctx, dg_dip* d, const uint8_t* all_handles) {
    if (!ctx || !d || !all_handles || d->world < 2 || d->attached) return DG_ERR_ARG;
    DG_CUDA(ctx, cudaSetDevice(ctx->device));
    for (int q = 0; q < d->world; ++q) {
        if (q == d->rank) continue;
        for (int a = 0; a < 4; ++a) {
            cudaIpcMemHandle_t h;
            memcpy(&h, all_handles + ((size_t)q * 4 + (size_t)a) * DG_IPC_HANDLE_BYTES, sizeof h);
            DG_CUDA(ctx, cudaIpcOpenMemHandle(&d->peer[a][q], h, cudaIpcMemLazyEnablePeerAccess));
        }
    }
    d->attached = true; d->ipc_opened = true;
    return DG_OK;
}

// The ranks of one sharded problem as sibling problems of ONE process on ONE GPU (tests, single-GPU boxes): the
// "peers" are the siblings' buffers themselves; every sibling gets a stream of its own and a plain launch, and the
// caller runs them together (world x ctas CTAs must fit the GPU, as for dg_dip_run_many).
int dg_dip_attach_in_process(dg_ctx* ctx, dg_dip** all, int32_t world) {
    if (!ctx || !all || world < 2 || world > DG_MAX_PEERS) return DG_ERR_ARG;
    DG_CUDA(ctx, cudaSetDevice(ctx->device));
    int64_t ctas = 0;
    for (int q = 0; q < world; ++q) {
        if (!all[q] || all[q]->world != world || all[q]->rank != q || all[q]->attached) return DG_ERR_ARG;
        ctas += all[q]->grid;
    }
    if (ctas > ctx->sm_count) return fail(ctx, DG_ERR_CAPACITY, "dg_dip_attach_in_process: %lld sweep CTAs do not fit %d SMs", (long long)ctas, ctx->sm_count);
    for (int q = 0; q < world; ++q) {
        dg_dip* d = all[q];
        cudaStream_t s = nullptr;
        if (int rc = batch_stream(ctx, q, &s)) return rc;
        d->stream = s; d->cooperative = false;
        for (int t = 0; t < world; ++t) {
            d->peer[0][t] = all[t]->tile0.p; d->peer[1][t] = all[t]->tile1.p;
            d->peer[2][t] = all[t]->pred.p; d->peer[3][t] = all[t]->counter.p;
        }
        d->armed = false;
    }
    for (int q = 0; q < world; ++q) all[q]->attached = true;
    return DG_OK;
}

int dg_dip_shard_arm(dg_ctx* ctx, dg_dip* d) {
    if (!ctx || !d || d->world < 2) return DG_ERR_ARG;
    DG_CUDA(ctx, cudaSetDevice(ctx->device));
    DG_CUDA(ctx, cudaMemsetAsync(d->counter.p, 0, 4 * sizeof(unsigned int), d->stream));
    DG_CUDA(ctx, cudaMemsetAsync(d->tile0.p, 0, (size_t)(d->plan.R + 1) * sizeof(int32_t), d->stream));
    DG_CUDA(ctx, cudaStreamSynchronize(d->stream));
    d->armed = true;
    return DG_OK;
}

int dg_dip_create_slot(dg_ctx* ctx, int32_t n_levels, const int32_t* level_off, const int64_t* adj_off,
                       const int32_t* adj_dst, const uint8_t* adj_w, const int64_t* col_off, const int32_t* col_val,
                       const uint8_t* colour_is_hom, int32_t n_colours, int32_t R, int32_t slot, int32_t ctas, dg_dip** out) {
    if (!ctx || !out || slot < 0 || slot >= 1024) return DG_ERR_ARG;
    DG_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t s = nullptr;
    if (int rc = batch_stream(ctx, slot, &s)) return rc;
    DipGraphView g;
    g.n_levels = n_levels; g.level_off = level_off; g.adj_off = adj_off; g.adj_dst = adj_dst; g.adj_w = adj_w;
    g.col_off = col_off; g.col_val = col_val; g.colour_is_hom = colour_is_hom; g.n_colours = n_colours; g.R = R;
    return dip_create_impl(ctx, g, out, s, ctas > 0 ? ctas : 8);
}

int dg_dip_run_many(dg_ctx* ctx, dg_dip** ds, int32_t n, float* wall_ms) {
    if (!ctx || n < 0 || (n > 0 && !ds)) return DG_ERR_ARG;
    DG_CUDA(ctx, cudaSetDevice(ctx->device));
    EventBag bag;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    DG_CUDA(ctx, bag.make(&e0));
    DG_CUDA(ctx, bag.make(&e1));
    std::vector<cudaEvent_t> done((size_t)n, nullptr);
    int rc = DG_OK;
    int64_t ctas = 0;
    bool spins = false;                 // some problem's CTAs wait on one another: the whole group must be co-resident
    for (int32_t i = 0; i < n; ++i) if (ds[i] && !ds[i]->cooperative) { ctas += ds[i]->grid; spins = spins || ds[i]->grid > 1 || !ds[i]->v4; }
    if (spins && ctas > ctx->sm_count) return fail(ctx, DG_ERR_CAPACITY, "dg_dip_run_many: %lld sweep CTAs do not fit %d SMs", (long long)ctas, ctx->sm_count);
    // One fused sweep launch for all problems when they allow it (plain-launch slots, same code width, no sharding):
    // the sweeps then do not need a hardware work queue each, so more than 32 samples can be resident together.
    bool fused = n >= 2 && !getenv("DG_NO_FUSED_MANY");
    for (int32_t i = 0; i < n && fused; ++i) {
        fused = ds[i] && !ds[i]->cooperative && ds[i]->world == 1 && ds[i]->plan.L > 1 && ds[i]->pred_bytes == ds[0]->pred_bytes &&
                ds[i]->v4 == ds[0]->v4;
        if (fused && ds[i]->v4)
            fused = ds[i]->p4.shape.cells() == ds[0]->p4.shape.cells() && (ds[i]->p4.max_giant > 0) == (ds[0]->p4.max_giant > 0) && ds[i]->p4.rc == ds[0]->p4.rc && ds[i]->p4.RL == ds[0]->p4.RL &&
                    ds[i]->p4.shape.slot_bytes == ds[0]->p4.shape.slot_bytes && ds[i]->p4.shape.nslot == ds[0]->p4.shape.nslot &&
                    ds[i]->v4_ncw == ds[0]->v4_ncw;
    }
    DG_CUDA(ctx, cudaEventRecord(e0, ctx->stream));
    if (fused && ds[0]->v4) {
        // Level-program problems: resets, ONE sweep launch and ONE traceback launch set for the whole group, all on the
        // context's stream (no per-problem streams, events or launches: with hundreds of resident problems those cost
        // more than the kernels they order).
        std::vector<Sweep4Args> h_args((size_t)n);
        std::vector<TraceArgs> h_ta((size_t)n);
        std::vector<int2> h_map;
        int64_t max_anc = 0;
        int max_M = 0;
        for (int32_t i = 0; i < n; ++i) {
            dg_dip* d = ds[i];
            d->want_prof = false; d->launches = 0;
            DG_CUDA(ctx, cudaMemsetAsync(d->counter.p, 0, 4 * sizeof(unsigned int), ctx->stream));
            fill_sweep4_args(d, h_args[(size_t)i], false);
            fill_trace_args(d, h_ta[(size_t)i]);
            for (int c = 0; c < d->grid; ++c) h_map.push_back(make_int2(i, c));
            max_anc = std::max(max_anc, d->anc_cells); max_M = std::max(max_M, d->M);
        }
        DevBuf<Sweep4Args> d_args;
        DevBuf<TraceArgs> d_ta;
        DevBuf<int2> d_map;
        DG_CUDA(ctx, d_args.upload(h_args.data(), h_args.size(), ctx->stream));
        DG_CUDA(ctx, d_ta.upload(h_ta.data(), h_ta.size(), ctx->stream));
        DG_CUDA(ctx, d_map.upload(h_map.data(), h_map.size(), ctx->stream));
        cudaEvent_t f0 = nullptr, swept = nullptr;
        DG_CUDA(ctx, bag.make(&f0));
        DG_CUDA(ctx, bag.make(&swept));
        const Plan4& q = ds[0]->p4;
        const void* fn = sweep4_many_fn(q.shape.cells(), q.rc, q.max_giant > 0);
        const size_t smem = sweep4_smem_bytes(q.shape.cells(), q.RL, q.shape.slot_bytes, q.shape.nslot);
        const Sweep4Args* pa = d_args.p;
        const int2* pm = d_map.p;
        void* args[] = {(void*)&pa, (void*)&pm};
        if (getenv("DG_TIMING")) {
            int per_sm = 0;
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, (ds[0]->v4_ncw + 1) * 32, smem);
            fprintf(stderr, "dg_dip_run_many: %zu CTAs of %d threads, %zu bytes of shared memory: %d per SM\n", h_map.size(), (ds[0]->v4_ncw + 1) * 32, smem, per_sm);
        }
        DG_CUDA(ctx, cudaEventRecord(f0, ctx->stream));
        DG_CUDA(ctx, cudaLaunchKernel(fn, dim3((unsigned)h_map.size()), dim3((unsigned)(ds[0]->v4_ncw + 1) * 32u), args, smem, ctx->stream));
        DG_CUDA(ctx, cudaEventRecord(swept, ctx->stream));
        const TraceArgs* pt = d_ta.p;
        if (max_anc > 0) dip_anc_many_kernel<uint16_t><<<dim3((unsigned)((max_anc + 255) / 256), (unsigned)n), 256, 0, ctx->stream>>>(pt);
        dip_hop_many_kernel<<<(n + 31) / 32, 32, 0, ctx->stream>>>(pt, n);
        if (max_M > 0) dip_seg_many_kernel<uint16_t><<<dim3((unsigned)((max_M + 127) / 128), (unsigned)n), 128, 0, ctx->stream>>>(pt);
        dip_merge_many_kernel<<<(n + 31) / 32, 32, 0, ctx->stream>>>(pt, n);
        DG_CUDA(ctx, cudaGetLastError());
        DG_CUDA(ctx, cudaEventRecord(e1, ctx->stream));
        DG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));      // (the argument vectors are read by the copies until here)
        float ms = 0.f, fm = 0.f, tm = 0.f;
        DG_CUDA(ctx, cudaEventElapsedTime(&ms, e0, e1));
        DG_CUDA(ctx, cudaEventElapsedTime(&fm, f0, swept));
        DG_CUDA(ctx, cudaEventElapsedTime(&tm, swept, e1));
        if (wall_ms) *wall_ms = ms;
        for (int32_t i = 0; i < n; ++i) { ds[i]->fused_ms = fm; ds[i]->group_trace_ms = tm; ds[i]->ran = true; ds[i]->checks = false; ds[i]->launches = i == 0 ? 5 : 0; }
        return DG_OK;
    }
    if (fused) {
        const bool v4 = ds[0]->v4;
        std::vector<SweepArgs> h_args(v4 ? 0 : (size_t)n);
        std::vector<Sweep4Args> h_args4(v4 ? (size_t)n : 0);
        std::vector<int2> h_map;
        for (int32_t i = 0; i < n; ++i) {
            dg_dip* d = ds[i];
            d->want_prof = false;
            DG_CUDA(ctx, cudaStreamWaitEvent(d->stream, e0, 0));
            if (int r = dip_run_pre(ctx, d, false)) { rc = r; break; }           // resets + pair scores, on the problem's stream
            DG_CUDA(ctx, bag.make(&done[(size_t)i], cudaEventDisableTiming));
            DG_CUDA(ctx, cudaEventRecord(done[(size_t)i], d->stream));
            DG_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, done[(size_t)i], 0));
            if (v4) fill_sweep4_args(d, h_args4[(size_t)i], false); else fill_sweep_args(d, h_args[(size_t)i]);
            for (int c = 0; c < d->grid; ++c) h_map.push_back(make_int2(i, c));
        }
        DevBuf<SweepArgs> d_args;
        DevBuf<Sweep4Args> d_args4;
        DevBuf<int2> d_map;
        cudaEvent_t swept = nullptr, f0 = nullptr;
        if (!rc) {
            if (v4) DG_CUDA(ctx, d_args4.upload(h_args4.data(), h_args4.size(), ctx->stream));
            else DG_CUDA(ctx, d_args.upload(h_args.data(), h_args.size(), ctx->stream));
            DG_CUDA(ctx, d_map.upload(h_map.data(), h_map.size(), ctx->stream));
            const int2* pm = d_map.p;
            DG_CUDA(ctx, bag.make(&f0));
            if (v4) {
                const Plan4& q = ds[0]->p4;
                const void* fn = sweep4_many_fn(q.shape.cells(), q.rc, q.max_giant > 0);
                const size_t smem = sweep4_smem_bytes(q.shape.cells(), q.RL, q.shape.slot_bytes, q.shape.nslot);
                const Sweep4Args* pa = d_args4.p;
                void* args[] = {(void*)&pa, (void*)&pm};
                DG_CUDA(ctx, cudaEventRecord(f0, ctx->stream));
                DG_CUDA(ctx, cudaLaunchKernel(fn, dim3((unsigned)h_map.size()), dim3((unsigned)(ds[0]->v4_ncw + 1) * 32u), args, smem, ctx->stream));
            } else {
                const bool p32 = ds[0]->pred_bytes == 4;
                const void* fn = p32 ? (const void*)dip_sweep_many_kernel<true> : (const void*)dip_sweep_many_kernel<false>;
                DG_CUDA(ctx, cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DIP_SMEM_BYTES));
                const SweepArgs* pa = d_args.p;
                void* args[] = {(void*)&pa, (void*)&pm};
                DG_CUDA(ctx, cudaEventRecord(f0, ctx->stream));
                DG_CUDA(ctx, cudaLaunchKernel(fn, dim3((unsigned)h_map.size()), dim3(DIP_THREADS), args, DIP_SMEM_BYTES, ctx->stream));
            }
            DG_CUDA(ctx, bag.make(&swept));
            DG_CUDA(ctx, cudaEventRecord(swept, ctx->stream));
            for (int32_t i = 0; i < n && !rc; ++i) {
                dg_dip* d = ds[i];
                DG_CUDA(ctx, cudaStreamWaitEvent(d->stream, swept, 0));
                if (i == 0) ++d->launches;                                      // the one fused launch is counted once
                rc = d->pred_bytes == 2 ? dip_run_post<uint16_t>(ctx, d, false) : dip_run_post<uint32_t>(ctx, d, false);
                if (rc) break;
                DG_CUDA(ctx, cudaEventRecord(done[(size_t)i], d->stream));
                DG_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, done[(size_t)i], 0));
            }
        }
        DG_CUDA(ctx, cudaEventRecord(e1, ctx->stream));
        DG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));      // (the argument vectors are read by the copies until here)
        float ms = 0.f;
        if (!rc) DG_CUDA(ctx, cudaEventElapsedTime(&ms, e0, e1));
        if (wall_ms) *wall_ms = ms;
        if (!rc && f0 && swept) {
            float fm = 0.f;
            DG_CUDA(ctx, cudaEventElapsedTime(&fm, f0, swept));
            for (int32_t i = 0; i < n; ++i) ds[i]->fused_ms = fm;
        }
        return rc;
    }
    for (int32_t i = 0; i < n && !rc; ++i) {
        if (!ds[i]) { rc = DG_ERR_ARG; break; }
        DG_CUDA(ctx, cudaStreamWaitEvent(ds[i]->stream, e0, 0));
        rc = dg_dip_run(ctx, ds[i], 0);
        if (rc) break;
        DG_CUDA(ctx, bag.make(&done[(size_t)i], cudaEventDisableTiming));
        DG_CUDA(ctx, cudaEventRecord(done[(size_t)i], ds[i]->stream));
        DG_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, done[(size_t)i], 0));
    }
    DG_CUDA(ctx, cudaEventRecord(e1, ctx->stream));
    DG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    float ms = 0.f;
    if (!rc) DG_CUDA(ctx, cudaEventElapsedTime(&ms, e0, e1));
    if (wall_ms) *wall_ms = ms;
    return rc;
}

}  // extern "C"
