// dp_diploid.cu — B200 diploid recombination-constrained DP (include/dipgenie_cuda.h: dg_dp_diploid, dg_dip_*).
//
// Replaces Approximator::diploid_dp_approximation_solver's sweep and edge-list recovery
// (reference src/approximator.cpp:532-716, :757-785).  Design:
//   * gather form: one thread per destination cell (r2,i',j') takes the lexicographic max of
//     (value, -i, -j) over in-edges(i') x in-edges(j') — no locks, no atomics, order-free, and the
//     winner is exactly the reference's (:657-659);
//   * the (R+1) x k x k int32 score layers of two consecutive levels ping-pong between two HBM
//     buffers that stay L2-resident; only a 2-byte predecessor code per cell is streamed out;
//   * one persistent cooperative kernel walks all L-1 transitions.  Narrow transitions run on CTA 0
//     alone (block barrier only); wide ones are spread over P_l CTAs and closed by a monotone-counter
//     grid barrier (no reset, no last-arriver logic: transition l is complete when the counter reaches
//     the precomputed prefix sum of arrivals);
//   * a single-thread traceback kernel turns predecessor codes into the two recombination-edge lists.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <vector>

#include "dg_common.cuh"
#include "dp_cell.h"
#include "dp_prep.h"

namespace dg {

// ---- kernel geometry -------------------------------------------------------------------------
constexpr int DIP_CT = 480;                    // compute threads per CTA (15 warps; 16 warps total -> 128 regs/thread)
constexpr int DIP_THREADS = DIP_CT + 32;       // + one producer warp (record prefetch via bulk async copies)
constexpr int DIP_CELLS_PER_CTA = DIP_CT * 4;
constexpr int DIP_STAGES = 4;                  // record ring depth
constexpr int DIP_STAGE_BYTES = 16384;
constexpr int DIP_TILE_CELLS = 16384;          // int32 cells per shared-memory layer tile (x2)
constexpr int DIP_QUEUE = 8;                   // look-ahead queue of (level, stage) entries
constexpr size_t DIP_SMEM_BYTES = (size_t)DIP_STAGES * DIP_STAGE_BYTES + 2 * (size_t)DIP_TILE_CELLS * 4 +
                                  DIP_STAGES * 8 + DIP_QUEUE * 8 + 16;

constexpr uint32_t CTL_WAIT = 1u;     // grid-level wait before the transition
constexpr uint32_t CTL_ARRIVE = 2u;   // grid-level arrive after the transition

struct __align__(16) LevelCtl {       // MODE_GLOBAL transitions read this instead of a staged record, 64 bytes
    int32_t voff2;                    // first vertex of level l+1
    int32_t k, k2, W;
    int32_t P;
    uint32_t wait_target;
    uint32_t flags;
    int32_t pad0;
    int64_t msrc_off, mdst_off, pred_off2;
    int64_t pad1;
};

struct __align__(16) LevelIdx {       // one per transition, scanned by every CTA's producer lane, 16 bytes
    int64_t rec_off;                  // byte offset of the packed record (-1: MODE_GLOBAL)
    uint32_t rec_bytes;
    uint16_t P;
    uint16_t mode;
};

struct SweepArgs {
    const LevelIdx* idx;
    const LevelCtl* ctl;
    const uint8_t* records;
    const int32_t* in_off;
    const uint32_t* in_edge;
    const uint64_t* masks;
    int32_t* tile0;
    int32_t* tile1;
    void* pred;
    unsigned int* counter;
    unsigned long long* level_sum;    // [L] (CHECK only)
    unsigned long long* level_live;   // [L]
    unsigned long long* prof;         // [32] phase cycle counters of CTA 0 / thread 0 (nullable)
    int32_t l_begin, l_end, R;
    int32_t pred32, check;
};

// ---- PTX helpers -----------------------------------------------------------------------------
__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release_add_u32(unsigned int* p, unsigned int v) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// Spin until the monotone arrival counter reaches `target`.  Polls are relaxed loads (no L1 invalidate
// per poll) with exponential back-off, so that idle CTAs — most of the grid during a narrow stretch —
// do not hammer the counter's L2 slice; one acquire fence orders the layer reads after the last poll.
__device__ __forceinline__ void wait_counter(const unsigned int* counter, unsigned int target) {
    unsigned int ns = 32;
    for (;;) {
        unsigned int v;
        asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
        if (v >= target) break;
        __nanosleep(ns);
        if (ns < 2048u) ns <<= 1;
    }
    asm volatile("fence.acq_rel.gpu;" ::: "memory");
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
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

// ---- the cell loop ---------------------------------------------------------------------------
struct CellIO {
    const int32_t* src;        // source layer (shared or global)
    int32_t* dst;              // destination layer (shared or global)
    uint8_t* pl;               // predecessor codes of the destination level
    bool src_smem, dst_smem, pred32, check;
};

// Work item = (destination pair (i',j'), chunk of RC layers); pairs vary fastest so that the stores of a
// warp are contiguous in every layer.  SMEM: both layers are shared-memory tiles (the common case inside
// a narrow stretch) and plain LDS/STS are used; otherwise placement is a run-time flag and HBM layers
// are read/written with L2-only (.cg) accesses, since other SMs produce and consume them.
template <class OffT, int RC, bool HAS_MASK, bool SMEM>
__device__ __forceinline__ void sweep_pairs(const TransitionT<OffT>& t, const CellIO& io, int R, uint64_t first,
                                            uint64_t stride, unsigned long long& hsum, unsigned long long& hlive) {
    const uint32_t k2 = (uint32_t)t.k2, npairs = k2 * k2;
    const uint32_t nchunk = (uint32_t)(R + RC) / RC;
    const uint64_t nitems = (uint64_t)npairs * nchunk;
    const int32_t* __restrict__ src = io.src;
    const bool ssm = io.src_smem;
    auto load = [src, ssm](int64_t idx) -> int32_t {
        if (SMEM) return src[idx];
        return ssm ? src[idx] : __ldcg(src + idx);
    };
    for (uint64_t x = first; x < nitems; x += stride) {
        uint32_t chunk, pair;
        if (nitems <= 0xFFFFFFFFull) { chunk = (uint32_t)x / npairs; pair = (uint32_t)x - chunk * npairs; }
        else { chunk = (uint32_t)(x / npairs); pair = (uint32_t)(x - (uint64_t)chunk * npairs); }
        const uint32_t i2 = pair / k2, j2 = pair - i2 * k2;
        const int r0 = (int)chunk * RC;
        uint64_t best[RC];
        uint32_t code[RC];
        relax_pair<RC, HAS_MASK>(t, load, R, r0, (int)i2, (int)j2, best, code);
#pragma unroll
        for (int rr = 0; rr < RC; ++rr) {
            const int r2 = r0 + rr;
            if (r2 <= R) {
                const uint64_t c = (uint64_t)r2 * npairs + pair;
                const uint64_t key = best[rr];
                const int32_t v = key_value(key);
                if (SMEM || io.dst_smem) io.dst[c] = v; else __stcg(io.dst + c, v);
                if (io.pred32) reinterpret_cast<uint32_t*>(io.pl)[c] = key ? code[rr] : 0xFFFFFFFFu;
                else reinterpret_cast<uint16_t*>(io.pl)[c] = key ? (uint16_t)(((code[rr] >> 16) << 8) | (code[rr] & 0xFFu)) : (uint16_t)0xFFFFu;
                if (io.check && key) {
                    ++hlive;
                    hsum += cell_fold(c, v, 0xFFFF - (int)((key >> 16) & 0xFFFF), 0xFFFF - (int)(key & 0xFFFF));
                }
            }
        }
    }
}

template <class OffT, bool SMEM>
__device__ __forceinline__ void sweep_dispatch(const TransitionT<OffT>& t, const CellIO& io, int R, uint64_t first,
                                               uint64_t stride, uint64_t nthreads, unsigned long long& hsum,
                                               unsigned long long& hlive) {
    const int rc = choose_rc((uint64_t)t.k2 * t.k2, R, nthreads);
    if (t.W > 0) {
        switch (rc) {
            case 8: sweep_pairs<OffT, 8, true, SMEM>(t, io, R, first, stride, hsum, hlive); break;
            case 4: sweep_pairs<OffT, 4, true, SMEM>(t, io, R, first, stride, hsum, hlive); break;
            case 2: sweep_pairs<OffT, 2, true, SMEM>(t, io, R, first, stride, hsum, hlive); break;
            default: sweep_pairs<OffT, 1, true, SMEM>(t, io, R, first, stride, hsum, hlive); break;
        }
    } else {
        switch (rc) {
            case 8: sweep_pairs<OffT, 8, false, SMEM>(t, io, R, first, stride, hsum, hlive); break;
            case 4: sweep_pairs<OffT, 4, false, SMEM>(t, io, R, first, stride, hsum, hlive); break;
            case 2: sweep_pairs<OffT, 2, false, SMEM>(t, io, R, first, stride, hsum, hlive); break;
            default: sweep_pairs<OffT, 1, false, SMEM>(t, io, R, first, stride, hsum, hlive); break;
        }
    }
}

// Producer lane state: scans the level index for this CTA's next transitions, issues the bulk copies
// of their records into free ring stages and publishes (level, stage) entries in the look-ahead queue.
struct Producer {
    int pf;            // scan cursor (level)
    int n_written;     // queue entries published so far
    int n_issued;      // records issued so far
    bool ended;
};

__global__ void __launch_bounds__(DIP_THREADS, 1) dip_sweep_kernel(const SweepArgs a) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* const stage_base = smem;
    int32_t* const tileS0 = reinterpret_cast<int32_t*>(smem + (size_t)DIP_STAGES * DIP_STAGE_BYTES);
    int32_t* const tileS1 = tileS0 + DIP_TILE_CELLS;
    uint64_t* const mbar = reinterpret_cast<uint64_t*>(tileS1 + DIP_TILE_CELLS);
    int2* const queue = reinterpret_cast<int2*>(mbar + DIP_STAGES);

    const int cta = blockIdx.x, tid = threadIdx.x;
    const bool is_compute = tid < DIP_CT;
    const bool is_producer = tid == DIP_CT;          // lane 0 of the extra warp
    uint8_t* const pred = reinterpret_cast<uint8_t*>(a.pred);
    const int pshift = a.pred32 ? 2 : 1;

    if (tid == 0) {
        for (int s = 0; s < DIP_STAGES; ++s) mbar_init(smem_u32(mbar + s), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int x = tid; x <= a.R; x += DIP_THREADS) tileS0[x] = 0;   // level 0: all R+1 layers start at 0 (approximator.cpp:535)
    __syncthreads();

    Producer pr;
    pr.pf = a.l_begin; pr.n_written = 0; pr.n_issued = 0; pr.ended = false;
    auto produce = [&](int t_cur, int rec_consumed) {
        // publish entries for iterations < t_cur + DIP_QUEUE - 1, as far as ring stages are free
        while (!pr.ended && pr.n_written < t_cur + DIP_QUEUE - 1) {
            while (pr.pf < a.l_end && cta >= (int)__ldg(&a.idx[pr.pf].P)) ++pr.pf;
            if (pr.pf >= a.l_end) {
                queue[pr.n_written % DIP_QUEUE] = make_int2(-1, -1);
                ++pr.n_written; pr.ended = true;
                break;
            }
            const int4 raw = __ldg(reinterpret_cast<const int4*>(a.idx + pr.pf));
            LevelIdx li;
            memcpy(&li, &raw, sizeof li);
            int stage = -1;
            if (li.rec_off >= 0) {
                if (pr.n_issued - rec_consumed >= DIP_STAGES) break;      // ring full: retry after the next level
                stage = pr.n_issued % DIP_STAGES;
                const uint32_t bar = smem_u32(mbar + stage);
                mbar_expect_tx(bar, li.rec_bytes);
                bulk_g2s(smem_u32(stage_base + (size_t)stage * DIP_STAGE_BYTES), a.records + li.rec_off, li.rec_bytes, bar);
                ++pr.n_issued;
            }
            queue[pr.n_written % DIP_QUEUE] = make_int2(pr.pf, stage);
            ++pr.n_written; ++pr.pf;
        }
    };
    if (is_producer) produce(0, 0);

    int t = 0, rc = 0;                               // iteration index, records consumed so far
    const bool profiling = a.prof != nullptr && cta == 0 && tid == 0;
    unsigned long long pc[24];
    if (profiling) for (int x = 0; x < 24; ++x) pc[x] = 0;
    long long tk0 = 0, tk1 = 0, tk2 = 0, tk3 = 0, tk4 = 0, tk5 = 0;
    for (;;) {
        if (profiling) tk0 = clock64();
        __syncthreads();                             // iteration t-1 complete: its layer is whole, its stage is free
        if (profiling) tk1 = clock64();
        const int2 e = queue[t % DIP_QUEUE];
        const int l = e.x, stage = e.y;
        if (l < 0) break;
        if (is_producer) produce(t + 1, rc);
        if (is_compute) {
            unsigned long long hsum = 0, hlive = 0;
            if (stage >= 0) {
                const uint8_t* rec = stage_base + (size_t)stage * DIP_STAGE_BYTES;
                mbar_wait(smem_u32(mbar + stage), (uint32_t)((rc / DIP_STAGES) & 1));
                if (profiling) tk2 = clock64();
                RecHeader h;
                TransitionT<uint16_t> tr;
                record_view(rec, h, tr);
                if (h.flags & REC_WAIT) {
                    if (tid == 0) wait_counter(a.counter, h.wait_target);
                    bar_compute();
                }
                if (profiling) tk3 = clock64();
                const bool ssm = (h.flags & REC_SRC_SMEM) != 0, dsm = (h.flags & REC_DST_SMEM) != 0;
                CellIO io;
                io.src = ssm ? ((l & 1) ? tileS1 : tileS0) : ((l & 1) ? a.tile1 : a.tile0);
                io.dst = dsm ? ((l & 1) ? tileS0 : tileS1) : ((l & 1) ? a.tile0 : a.tile1);
                io.pl = pred + ((size_t)h.pred_off2 << pshift);
                io.src_smem = ssm; io.dst_smem = dsm; io.pred32 = a.pred32 != 0; io.check = a.check != 0;
                const uint64_t first = (uint64_t)cta * DIP_CT + tid, stride = (uint64_t)h.P * DIP_CT;
                if (ssm && dsm) {
                    CellIO ios = io;
                    ios.src = (l & 1) ? tileS1 : tileS0;
                    ios.dst = (l & 1) ? tileS0 : tileS1;
                    sweep_dispatch<uint16_t, true>(tr, ios, a.R, first, stride, stride, hsum, hlive);
                } else {
                    sweep_dispatch<uint16_t, false>(tr, io, a.R, first, stride, stride, hsum, hlive);
                }
                if (profiling) tk4 = clock64();
                if (h.flags & REC_ARRIVE) {
                    bar_compute();
                    if (tid == 0) red_release_add_u32(a.counter, 1u);
                }
                if (profiling) {
                    tk5 = clock64();
                    const int m = (ssm || dsm) ? 0 : 1;       // 0: shared-memory layers, 1: staged record + HBM layers
                    pc[m * 6 + 0] += 1; pc[m * 6 + 1] += tk1 - tk0; pc[m * 6 + 2] += tk2 - tk1;
                    pc[m * 6 + 3] += tk3 - tk2; pc[m * 6 + 4] += tk4 - tk3; pc[m * 6 + 5] += tk5 - tk4;
                }
            } else {
                if (profiling) tk2 = clock64();
                const LevelCtl* __restrict__ cg = a.ctl + l;
                const uint32_t flags = __ldg(&cg->flags);
                if (flags & CTL_WAIT) {
                    if (tid == 0) wait_counter(a.counter, __ldg(&cg->wait_target));
                    bar_compute();
                }
                if (profiling) tk3 = clock64();
                Transition tr;
                tr.k = __ldg(&cg->k); tr.k2 = __ldg(&cg->k2); tr.W = __ldg(&cg->W);
                tr.in_off = a.in_off + __ldg(&cg->voff2);
                tr.in_edge = a.in_edge;
                tr.msrc = a.masks + __ldg(&cg->msrc_off);
                tr.mdst = a.masks + __ldg(&cg->mdst_off);
                CellIO io;
                io.src = (l & 1) ? a.tile1 : a.tile0;
                io.dst = (l & 1) ? a.tile0 : a.tile1;
                io.pl = pred + ((size_t)__ldg(&cg->pred_off2) << pshift);
                io.src_smem = false; io.dst_smem = false; io.pred32 = a.pred32 != 0; io.check = a.check != 0;
                const uint64_t first = (uint64_t)cta * DIP_CT + tid, stride = (uint64_t)__ldg(&cg->P) * DIP_CT;
                sweep_dispatch<int32_t, false>(tr, io, a.R, first, stride, stride, hsum, hlive);
                if (profiling) tk4 = clock64();
                if (flags & CTL_ARRIVE) {
                    bar_compute();
                    if (tid == 0) red_release_add_u32(a.counter, 1u);
                }
                if (profiling) {
                    tk5 = clock64();
                    pc[12] += 1; pc[13] += tk1 - tk0; pc[14] += tk2 - tk1; pc[15] += tk3 - tk2; pc[16] += tk4 - tk3; pc[17] += tk5 - tk4;
                }
            }
            if (a.check) {
                for (int o = 16; o > 0; o >>= 1) {
                    hsum += __shfl_down_sync(0xFFFFFFFFu, hsum, o);
                    hlive += __shfl_down_sync(0xFFFFFFFFu, hlive, o);
                }
                if ((tid & 31) == 0 && hlive) {
                    atomicAdd(a.level_sum + l + 1, hsum);
                    atomicAdd(a.level_live + l + 1, hlive);
                }
            }
        }
        if (stage >= 0) ++rc;
        ++t;
    }
    if (profiling) for (int x = 0; x < 24; ++x) a.prof[x] = pc[x];
}

struct TraceOut {           // device-side result block
    int32_t rc, value, s_het, n1, n2, pad[3];
};

template <class PredT>
__global__ void dip_traceback_kernel(TraceView v, const PredT* pred, const int32_t* sink_tile, int cap,
                                     TraceOut* out, int32_t* p1, int32_t* p2) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const int64_t ks = v.level_off[v.L] - v.level_off[v.L - 1];
    const int32_t value = __ldcg(sink_tile + (int64_t)v.R * ks * ks);   // cell (r=R,0,0) of the last level (:730, :775)
    int32_t n1 = 0, n2 = 0, s_het = 0;
    const int rc = traceback<PredT>(v, pred, value, p1, &n1, p2, &n2, cap, &s_het);
    out->rc = rc; out->value = value; out->s_het = s_het; out->n1 = n1; out->n2 = n2;
}

}  // namespace dg

using namespace dg;

struct dg_dip {
    DipPlan plan;                 // host copy (small arrays kept for stats; big ones released after upload)
    int pred_bytes = 2;
    int grid = 1;
    DevBuf<LevelCtl> ctl;
    DevBuf<LevelIdx> idx;
    DevBuf<uint8_t> records;
    DevBuf<int32_t> level_off, in_off, lvlW;
    DevBuf<uint32_t> in_edge;
    DevBuf<uint64_t> masks;
    DevBuf<int64_t> msrc_off, mdst_off, pred_off;
    DevBuf<int32_t> tile0, tile1;
    DevBuf<uint8_t> pred;
    DevBuf<unsigned int> counter;
    DevBuf<unsigned long long> level_sum, level_live, prof;
    bool want_prof = false;
    DevBuf<TraceOut> tout;
    DevBuf<int32_t> p1, p2;
    cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
    float sweep_ms = 0.f, trace_ms = 0.f;
    int launches = 0;
    bool ran = false, checks = false;
    uint64_t device_bytes = 0;
    ~dg_dip() { for (auto& e : ev) if (e) cudaEventDestroy(e); }
};

static int dip_create_impl(dg_ctx* ctx, const DipGraphView& g, dg_dip** out) {
    *out = nullptr;
    DG_CUDA(ctx, cudaSetDevice(ctx->device));
    std::unique_ptr<dg_dip> d(new dg_dip());
    if (!build_dip_plan(g, d->plan)) return fail(ctx, DG_ERR_ARG, "dg_dip_create: %s", d->plan.error.c_str());
    DipPlan& p = d->plan;
    const int L = p.L;
    d->pred_bytes = (p.max_indeg <= 255) ? 2 : 4;

    // grid: enough CTAs for the widest transition, at most one co-resident wave (cooperative launch)
    const void* fnc = (const void*)dip_sweep_kernel;
    DG_CUDA(ctx, cudaFuncSetAttribute(fnc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DIP_SMEM_BYTES));
    int per_sm = 0;
    DG_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fnc, DIP_THREADS, DIP_SMEM_BYTES));
    if (per_sm < 1) return fail(ctx, DG_ERR_CUDA, "dg_dip_create: sweep kernel cannot be resident");
    int max_grid = ctx->sm_count * per_sm;
    if (const char* e = getenv("DG_DIP_MAX_GRID")) max_grid = std::max(1, std::min(max_grid, atoi(e)));   // diagnostics
    const uint64_t widest = (uint64_t)(p.R + 1) * (uint64_t)p.kmax * (uint64_t)p.kmax;
    const uint64_t want = (widest + DIP_CELLS_PER_CTA - 1) / DIP_CELLS_PER_CTA;
    SweepShape shape;
    shape.grid = (int)std::min<uint64_t>(std::max<uint64_t>(want, 1), (uint64_t)max_grid);
    shape.cells_per_cta = DIP_CELLS_PER_CTA;
    shape.tile_cells = DIP_TILE_CELLS;
    shape.stage_bytes = DIP_STAGE_BYTES;
    plan_sweep(p, shape);
    int gmax = 1;
    for (int l = 0; l + 1 < L; ++l) gmax = std::max(gmax, p.P[l]);
    d->grid = gmax;

    std::vector<LevelCtl> ctl((size_t)std::max(L - 1, 1));
    std::vector<LevelIdx> idx((size_t)std::max(L - 1, 1));
    memset(ctl.data(), 0, ctl.size() * sizeof(LevelCtl));
    memset(idx.data(), 0, idx.size() * sizeof(LevelIdx));
    for (int l = 0; l + 1 < L; ++l) {
        LevelCtl& c = ctl[l];
        c.voff2 = p.level_off[l + 1];
        c.k = p.level_off[l + 1] - p.level_off[l];
        c.k2 = p.level_off[l + 2] - p.level_off[l + 1];
        c.W = p.lvlW[l];
        c.P = p.P[l];
        c.flags = 0;
        if (p.flags[l] & REC_WAIT) { c.flags |= CTL_WAIT; c.wait_target = p.bar_target[l - 1]; }
        if (p.flags[l] & REC_ARRIVE) c.flags |= CTL_ARRIVE;
        c.msrc_off = p.msrc_off[l]; c.mdst_off = p.mdst_off[l]; c.pred_off2 = p.pred_off[l + 1];
        LevelIdx& x = idx[l];
        x.rec_off = p.rec_off[l];
        x.rec_bytes = 0;
        if (p.rec_off[l] >= 0) {
            RecHeader h;
            memcpy(&h, p.records.data() + p.rec_off[l], sizeof h);
            x.rec_bytes = h.bytes;
        }
        x.P = (uint16_t)p.P[l];
        x.mode = p.mode[l];
    }

    cudaStream_t s = ctx->stream;
    DG_CUDA(ctx, d->ctl.upload(ctl.data(), ctl.size(), s));
    DG_CUDA(ctx, d->idx.upload(idx.data(), idx.size(), s));
    DG_CUDA(ctx, d->records.upload(p.records.data(), p.records.size(), s));
    DG_CUDA(ctx, d->level_off.upload(p.level_off.data(), p.level_off.size(), s));
    DG_CUDA(ctx, d->in_off.upload(p.in_off.data(), p.in_off.size(), s));
    DG_CUDA(ctx, d->in_edge.upload(p.in_edge.data(), p.in_edge.size(), s));
    DG_CUDA(ctx, d->lvlW.upload(p.lvlW.data(), p.lvlW.size(), s));
    DG_CUDA(ctx, d->masks.upload(p.masks.data(), p.masks.size(), s));
    DG_CUDA(ctx, d->msrc_off.upload(p.msrc_off.data(), p.msrc_off.size(), s));
    DG_CUDA(ctx, d->mdst_off.upload(p.mdst_off.data(), p.mdst_off.size(), s));
    DG_CUDA(ctx, d->pred_off.upload(p.pred_off.data(), p.pred_off.size(), s));
    const size_t tile = (size_t)std::max<uint64_t>(widest, (uint64_t)(p.R + 1));
    DG_CUDA(ctx, d->tile0.alloc(tile));
    DG_CUDA(ctx, d->tile1.alloc(tile));
    DG_CUDA(ctx, d->pred.alloc((size_t)p.pred_off[L] * (size_t)d->pred_bytes));
    DG_CUDA(ctx, d->counter.alloc(1));
    DG_CUDA(ctx, d->level_sum.alloc((size_t)L));
    DG_CUDA(ctx, d->level_live.alloc((size_t)L));
    DG_CUDA(ctx, d->prof.alloc(32));
    DG_CUDA(ctx, d->tout.alloc(1));
    DG_CUDA(ctx, d->p1.alloc((size_t)2 * (p.R + 2)));
    DG_CUDA(ctx, d->p2.alloc((size_t)2 * (p.R + 2)));
    for (auto& e : d->ev) DG_CUDA(ctx, cudaEventCreate(&e));
    DG_CUDA(ctx, cudaStreamSynchronize(s));
    d->device_bytes = d->idx.bytes() + d->records.bytes() + d->ctl.bytes() + d->level_off.bytes() + d->in_off.bytes() + d->in_edge.bytes() + d->lvlW.bytes() +
                      d->masks.bytes() + d->msrc_off.bytes() + d->mdst_off.bytes() + d->pred_off.bytes() + d->tile0.bytes() +
                      d->tile1.bytes() + d->pred.bytes() + d->level_sum.bytes() + d->level_live.bytes();
    // the big host arrays are no longer needed
    std::vector<uint32_t>().swap(p.in_edge);
    std::vector<uint8_t>().swap(p.records);
    std::vector<uint64_t>().swap(p.masks);
    std::vector<int32_t>().swap(p.in_off);
    *out = d.release();
    return DG_OK;
}

template <class PredT>
static int dip_run_impl(dg_ctx* ctx, dg_dip* d, bool check) {
    const DipPlan& p = d->plan;
    cudaStream_t s = ctx->stream;
    DG_CUDA(ctx, cudaMemsetAsync(d->counter.p, 0, sizeof(unsigned int), s));
    DG_CUDA(ctx, cudaMemsetAsync(d->tile0.p, 0, (size_t)(p.R + 1) * sizeof(int32_t), s));   // dp_cur.assign(R+1, {0,0}) :535
    if (check) {
        std::vector<unsigned long long> basis((size_t)p.L, FOLD_BASIS);
        DG_CUDA(ctx, cudaMemcpyAsync(d->level_sum.p, basis.data(), basis.size() * 8, cudaMemcpyHostToDevice, s));
        DG_CUDA(ctx, cudaMemsetAsync(d->level_live.p, 0, (size_t)p.L * 8, s));
        DG_CUDA(ctx, cudaStreamSynchronize(s));   // basis is a stack-lifetime staging buffer
    }
    SweepArgs a;
    a.idx = d->idx.p; a.records = d->records.p;
    a.ctl = d->ctl.p; a.in_off = d->in_off.p; a.in_edge = d->in_edge.p; a.masks = d->masks.p;
    a.tile0 = d->tile0.p; a.tile1 = d->tile1.p; a.pred = d->pred.p; a.counter = d->counter.p;
    a.level_sum = d->level_sum.p; a.level_live = d->level_live.p;
    a.prof = d->want_prof ? d->prof.p : nullptr;
    a.l_begin = 0; a.l_end = p.L - 1; a.R = p.R;
    a.pred32 = sizeof(PredT) == 4 ? 1 : 0; a.check = check ? 1 : 0;
    d->launches = 0;
    DG_CUDA(ctx, cudaEventRecord(d->ev[0], s));
    if (p.L > 1) {
        void* args[] = {(void*)&a};
        const void* fn = (const void*)dip_sweep_kernel;
        DG_CUDA(ctx, cudaLaunchCooperativeKernel(fn, dim3(d->grid), dim3(DIP_THREADS), args, DIP_SMEM_BYTES, s));
        ++d->launches;
    }
    DG_CUDA(ctx, cudaEventRecord(d->ev[1], s));
    TraceView v;
    v.L = p.L; v.R = p.R; v.level_off = d->level_off.p; v.in_off = d->in_off.p; v.in_edge = d->in_edge.p;
    v.lvlW = d->lvlW.p; v.msrc_off = d->msrc_off.p; v.mdst_off = d->mdst_off.p; v.masks = d->masks.p;
    v.pred_off = d->pred_off.p;
    const int32_t* sink_tile = ((p.L - 1) & 1) ? d->tile1.p : d->tile0.p;
    dip_traceback_kernel<PredT><<<1, 32, 0, s>>>(v, reinterpret_cast<const PredT*>(d->pred.p), sink_tile, p.R + 2,
                                                 d->tout.p, d->p1.p, d->p2.p);
    ++d->launches;
    DG_CUDA(ctx, cudaGetLastError());
    DG_CUDA(ctx, cudaEventRecord(d->ev[2], s));
    d->ran = true; d->checks = check;
    return DG_OK;
}

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
    cudaStream_t s = ctx->stream;
    DG_CUDA(ctx, cudaMemcpyAsync(&t, d->tout.p, sizeof t, cudaMemcpyDeviceToHost, s));
    DG_CUDA(ctx, cudaMemcpyAsync(a.data(), d->p1.p, a.size() * 4, cudaMemcpyDeviceToHost, s));
    DG_CUDA(ctx, cudaMemcpyAsync(b.data(), d->p2.p, b.size() * 4, cudaMemcpyDeviceToHost, s));
    DG_CUDA(ctx, cudaStreamSynchronize(s));
    DG_CUDA(ctx, cudaEventElapsedTime(&d->sweep_ms, d->ev[0], d->ev[1]));
    DG_CUDA(ctx, cudaEventElapsedTime(&d->trace_ms, d->ev[1], d->ev[2]));
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
    DG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
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
    (void)ctx;
    return DG_OK;
}

int dg_dip_profile(dg_ctx* ctx, dg_dip* d, uint64_t* out24) {
    if (!ctx || !d || !d->ran || !d->want_prof) return fail(ctx, DG_ERR_ARG, "dg_dip_profile: run with flags bit1 first");
    DG_CUDA(ctx, cudaSetDevice(ctx->device));
    DG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    DG_CUDA(ctx, cudaMemcpy(out24, d->prof.p, 24 * 8, cudaMemcpyDeviceToHost));
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
    int rc = dg_dip_create(ctx, n_levels, level_off, adj_off, adj_dst, adj_w, col_off, col_val, colour_is_hom, n_colours, R, &d);
    if (rc) return rc;
    rc = dg_dip_run(ctx, d, 0);
    if (!rc) rc = dg_dip_result(ctx, d, sink_value, sink_s_het, p1_edges, n_p1, p2_edges, n_p2);
    dg_dip_destroy(ctx, d);
    return rc;
}

}  // extern "C"
