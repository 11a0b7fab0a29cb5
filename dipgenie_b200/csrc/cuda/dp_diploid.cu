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
#include <cstring>
#include <memory>
#include <vector>

#include "dg_common.cuh"
#include "dp_cell.h"
#include "dp_prep.h"

namespace dg {

constexpr int DIP_THREADS = 256;
constexpr int DIP_CELLS_PER_CTA = DIP_THREADS * 8;

constexpr uint32_t CTL_WAIT = 1u;     // grid-level wait before the transition
constexpr uint32_t CTL_ARRIVE = 2u;   // grid-level arrive after the transition

struct __align__(16) LevelCtl {       // one per transition l (level l -> l+1), 64 bytes
    int32_t voff2;                    // first vertex of level l+1
    int32_t k, k2, W;
    int32_t P;
    uint32_t wait_target;
    uint32_t flags;
    int32_t pad0;
    int64_t msrc_off, mdst_off, pred_off2;
    int64_t pad1;
};

struct SweepArgs {
    const LevelCtl* ctl;
    const int32_t* in_off;
    const uint32_t* in_edge;
    const uint64_t* masks;
    int32_t* tile0;
    int32_t* tile1;
    void* pred;
    unsigned int* counter;
    unsigned long long* level_sum;    // [L] (CHECK only)
    unsigned long long* level_live;   // [L]
    int32_t l_begin, l_end, R;
};

__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release_add_u32(unsigned int* p, unsigned int v) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

template <class PredT, bool CHECK>
__global__ void __launch_bounds__(DIP_THREADS) dip_sweep_kernel(const SweepArgs a) {
    const int cta = blockIdx.x, tid = threadIdx.x;
    PredT* __restrict__ pred = reinterpret_cast<PredT*>(a.pred);
    constexpr int SH = (sizeof(PredT) == 2) ? 8 : 16;

    for (int l = a.l_begin; l < a.l_end; ++l) {
        const LevelCtl* __restrict__ cg = a.ctl + l;
        const int P = __ldg(&cg->P);
        if (cta >= P) continue;                       // not a participant of this transition
        const uint32_t flags = __ldg(&cg->flags);
        if (flags & CTL_WAIT) {
            if (tid == 0) {
                const unsigned int target = __ldg(&cg->wait_target);
                while (ld_acquire_u32(a.counter) < target) __nanosleep(20);
            }
        }
        __syncthreads();

        Transition t;
        t.k = __ldg(&cg->k);
        t.k2 = __ldg(&cg->k2);
        t.W = __ldg(&cg->W);
        t.in_off = a.in_off + __ldg(&cg->voff2);
        t.in_edge = a.in_edge;
        t.msrc = a.masks + __ldg(&cg->msrc_off);
        t.mdst = a.masks + __ldg(&cg->mdst_off);
        const int32_t* __restrict__ src = (l & 1) ? a.tile1 : a.tile0;
        int32_t* __restrict__ dst = (l & 1) ? a.tile0 : a.tile1;
        PredT* __restrict__ pl = pred + __ldg(&cg->pred_off2);
        auto load = [src](int64_t idx) { return __ldcg(src + idx); };

        const uint32_t kk = (uint32_t)t.k2 * (uint32_t)t.k2;
        const uint64_t ncell = (uint64_t)(a.R + 1) * kk;
        unsigned long long hsum = 0, hlive = 0;
        if (ncell <= 0x7FFFFFFFull) {
            const uint32_t n32 = (uint32_t)ncell, stride = (uint32_t)P * DIP_THREADS;
            for (uint32_t c = (uint32_t)cta * DIP_THREADS + tid; c < n32; c += stride) {
                const uint32_t r2 = c / kk, rem = c - r2 * kk;
                const uint32_t i2 = rem / (uint32_t)t.k2, j2 = rem - i2 * (uint32_t)t.k2;
                uint32_t code;
                const uint64_t key = relax_cell(t, load, (int)r2, (int)i2, (int)j2, code);
                __stcg(dst + c, key_value(key));
                pl[c] = key ? (PredT)(((code >> 16) << SH) | (code & 0xFFFFu)) : (PredT) ~(PredT)0;
                if (CHECK && key) {
                    ++hlive;
                    hsum += cell_fold(c, key_value(key), 0xFFFF - (int)((key >> 16) & 0xFFFF), 0xFFFF - (int)(key & 0xFFFF));
                }
            }
        } else {
            const uint64_t stride = (uint64_t)P * DIP_THREADS;
            for (uint64_t c = (uint64_t)cta * DIP_THREADS + tid; c < ncell; c += stride) {
                const uint64_t r2 = c / kk, rem = c - r2 * kk;
                const uint32_t i2 = (uint32_t)(rem / (uint32_t)t.k2), j2 = (uint32_t)(rem - (uint64_t)i2 * (uint32_t)t.k2);
                uint32_t code;
                const uint64_t key = relax_cell(t, load, (int)r2, (int)i2, (int)j2, code);
                __stcg(dst + c, key_value(key));
                pl[c] = key ? (PredT)(((code >> 16) << SH) | (code & 0xFFFFu)) : (PredT) ~(PredT)0;
                if (CHECK && key) {
                    ++hlive;
                    hsum += cell_fold(c, key_value(key), 0xFFFF - (int)((key >> 16) & 0xFFFF), 0xFFFF - (int)(key & 0xFFFF));
                }
            }
        }
        if (CHECK) {
            for (int o = 16; o > 0; o >>= 1) {
                hsum += __shfl_down_sync(0xFFFFFFFFu, hsum, o);
                hlive += __shfl_down_sync(0xFFFFFFFFu, hlive, o);
            }
            if ((tid & 31) == 0 && hlive) {
                atomicAdd(a.level_sum + l + 1, hsum);
                atomicAdd(a.level_live + l + 1, hlive);
            }
        }
        if (flags & CTL_ARRIVE) {
            __syncthreads();
            if (tid == 0) red_release_add_u32(a.counter, 1u);
        }
    }
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
    DevBuf<int32_t> level_off, in_off, lvlW;
    DevBuf<uint32_t> in_edge;
    DevBuf<uint64_t> masks;
    DevBuf<int64_t> msrc_off, mdst_off, pred_off;
    DevBuf<int32_t> tile0, tile1;
    DevBuf<uint8_t> pred;
    DevBuf<unsigned int> counter;
    DevBuf<unsigned long long> level_sum, level_live;
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

    // grid: enough CTAs for the widest transition, at most one co-resident wave
    int per_sm = 0;
    if (d->pred_bytes == 2) DG_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, dip_sweep_kernel<uint16_t, false>, DIP_THREADS, 0));
    else DG_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, dip_sweep_kernel<uint32_t, false>, DIP_THREADS, 0));
    if (per_sm < 1) return fail(ctx, DG_ERR_CUDA, "dg_dip_create: sweep kernel cannot be resident");
    const int max_grid = ctx->sm_count * std::min(per_sm, 2);
    const uint64_t widest = (uint64_t)(p.R + 1) * (uint64_t)p.kmax * (uint64_t)p.kmax;
    const uint64_t want = (widest + DIP_CELLS_PER_CTA - 1) / DIP_CELLS_PER_CTA;
    d->grid = (int)std::min<uint64_t>(std::max<uint64_t>(want, 1), (uint64_t)max_grid);
    plan_participants(p, d->grid, DIP_CELLS_PER_CTA);

    std::vector<LevelCtl> ctl((size_t)std::max(L - 1, 1));
    memset(ctl.data(), 0, ctl.size() * sizeof(LevelCtl));
    for (int l = 0; l + 1 < L; ++l) {
        LevelCtl& c = ctl[l];
        c.voff2 = p.level_off[l + 1];
        c.k = p.level_off[l + 1] - p.level_off[l];
        c.k2 = p.level_off[l + 2] - p.level_off[l + 1];
        c.W = p.lvlW[l];
        c.P = p.P[l];
        c.flags = 0;
        if (l > 0 && p.bar_edge[l - 1]) { c.flags |= CTL_WAIT; c.wait_target = p.bar_target[l - 1]; }
        if (p.bar_edge[l]) c.flags |= CTL_ARRIVE;
        c.msrc_off = p.msrc_off[l]; c.mdst_off = p.mdst_off[l]; c.pred_off2 = p.pred_off[l + 1];
    }

    cudaStream_t s = ctx->stream;
    DG_CUDA(ctx, d->ctl.upload(ctl.data(), ctl.size(), s));
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
    DG_CUDA(ctx, d->tout.alloc(1));
    DG_CUDA(ctx, d->p1.alloc((size_t)2 * (p.R + 2)));
    DG_CUDA(ctx, d->p2.alloc((size_t)2 * (p.R + 2)));
    for (auto& e : d->ev) DG_CUDA(ctx, cudaEventCreate(&e));
    DG_CUDA(ctx, cudaStreamSynchronize(s));
    d->device_bytes = d->ctl.bytes() + d->level_off.bytes() + d->in_off.bytes() + d->in_edge.bytes() + d->lvlW.bytes() +
                      d->masks.bytes() + d->msrc_off.bytes() + d->mdst_off.bytes() + d->pred_off.bytes() + d->tile0.bytes() +
                      d->tile1.bytes() + d->pred.bytes() + d->level_sum.bytes() + d->level_live.bytes();
    // the big host arrays are no longer needed
    std::vector<uint32_t>().swap(p.in_edge);
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
    a.ctl = d->ctl.p; a.in_off = d->in_off.p; a.in_edge = d->in_edge.p; a.masks = d->masks.p;
    a.tile0 = d->tile0.p; a.tile1 = d->tile1.p; a.pred = d->pred.p; a.counter = d->counter.p;
    a.level_sum = d->level_sum.p; a.level_live = d->level_live.p;
    a.l_begin = 0; a.l_end = p.L - 1; a.R = p.R;
    d->launches = 0;
    DG_CUDA(ctx, cudaEventRecord(d->ev[0], s));
    if (p.L > 1) {
        void* args[] = {(void*)&a};
        const void* fn = check ? (const void*)dip_sweep_kernel<PredT, true> : (const void*)dip_sweep_kernel<PredT, false>;
        DG_CUDA(ctx, cudaLaunchCooperativeKernel(fn, dim3(d->grid), dim3(DIP_THREADS), args, 0, s));
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
