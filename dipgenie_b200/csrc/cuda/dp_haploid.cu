// dp_haploid.cu — B200 haploid recombination-constrained DP (include/dipgenie_cuda.h: dg_dp_haploid, dg_hap_*).
//
// Replaces the push-relaxation loop and the (R+1) tracebacks of Approximator::dp_approximation_solver
// (reference src/approximator.cpp:50-102, :141-151).  Design:
//   * gather form over longest-path levels.  The reference visits u in Kahn order, r ascending, edges in
//     adjacency order and writes dp[v][r+w] only on a strict improvement (:60), so cell (v,r2) ends with the
//     maximum candidate and the back pointer of the FIRST candidate reaching it in the order
//     (u ascending, r ascending, adjacency order).  Here every cell scans its in-edges, pre-sorted on the
//     host in exactly that order, and keeps the first strict maximum; the table starts at 0 and a
//     candidate that does not beat 0 leaves the cell without a predecessor (:50-52) — reproduced.
//   * vertices are renumbered level-major ("slots") so that one level is one contiguous run of the score
//     table [slot][r]; a level only reads earlier levels, so one block barrier per level is all the
//     synchronisation the persistent sweep needs.  The winner's in-edge (source slot | weight, 4 bytes) is
//     the only predecessor state kept, so a traceback step is one dependent load.
//   * R+1 independent traceback chains (one CTA each), then a fully parallel distinct-colour count per
//     chain (bitmaps + popcount), which is what the reference's best_r rule consumes (:116-136; the
//     floating-point rule itself stays on the host).
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <numeric>
#include <vector>

#include "dg_common.cuh"

namespace dg {

constexpr int HAP_THREADS = 1024;
constexpr uint32_t HAP_W_SHIFT = 30;            // in_edge = source slot | weight << 30
constexpr uint32_t HAP_SLOT_MASK = (1u << HAP_W_SHIFT) - 1;
constexpr uint32_t HAP_NO_PRED = 0xFFFFFFFFu;  // slots stay below 2^30 - 1, so this is never a real in-edge

struct HapSweepArgs {
    const int32_t* level_off;   // [L+1] slot ranges
    const int64_t* in_off;      // [n+1] by slot
    const uint32_t* in_edge;    // source slot | w << 30, reference visiting order
    const int32_t* ncol;        // [n] by slot: |color[v]|
    int32_t* dp;                // [n][R+1] by slot
    uint32_t* pred;             // [n][R+1] winner's in-edge (source slot | w << 30), HAP_NO_PRED = none
    int32_t L, R;
    int32_t ring;               // slots held by the shared-memory score ring (power of two)
};

// One persistent CTA walks the levels; level 0 (no in-edges) keeps the initial 0 / no-pred state.
// A level costs one dependent chain in-edge offsets -> in-edge -> source score.  Everything but the score is
// independent of the DP, so every thread fetches the offsets and the colour count of its cell two levels ahead
// and its first in-edge one level ahead: after the barrier only the score load (shared-memory ring) is left.
struct HapCell {
    int32_t s, r2, add;
    int64_t e0, e1;
    uint32_t x0;
    bool has;
};
// stage A: which cell, its in-edge range and colour count (loads that depend on nothing)
__device__ __forceinline__ HapCell hap_stage_a(const HapSweepArgs& a, int l, int S, int c) {
    HapCell p;
    p.has = false; p.s = 0; p.r2 = 0; p.add = 0; p.e0 = 0; p.e1 = 0; p.x0 = 0;
    if (l >= a.L) return p;
    const int lo = __ldg(a.level_off + l), hi = __ldg(a.level_off + l + 1);
    if (c >= (hi - lo) * S) return p;
    p.has = true;
    p.s = lo + c / S; p.r2 = c - (c / S) * S;
    p.e0 = __ldg(a.in_off + p.s); p.e1 = __ldg(a.in_off + p.s + 1);
    p.add = __ldg(a.ncol + p.s);
    return p;
}
// stage B: the first in-edge (depends on stage A's offsets, issued one level later)
__device__ __forceinline__ void hap_stage_b(const HapSweepArgs& a, HapCell& p) {
    if (p.has && p.e0 < p.e1) p.x0 = __ldg(a.in_edge + p.e0);
}
__device__ __forceinline__ HapCell hap_prefetch(const HapSweepArgs& a, int l, int S, int c) {
    HapCell p = hap_stage_a(a, l, S, c);
    hap_stage_b(a, p);
    return p;
}
// The scores of the most recent `ring` slots (slots are level-major, so these are the last few hundred levels)
// also live in a shared-memory ring: almost every in-edge comes from there (29-cycle LDS instead of an L2 round
// trip on the per-level critical chain); older sources are read from the table in HBM/L2.  `hi` = end of the level
// being written: entries below hi - ring are being overwritten by this level and must not be read from the ring.
__device__ __forceinline__ void hap_cell(const HapSweepArgs& a, int S, const HapCell& p, int32_t* ring, int hi, int valid_from) {
    int32_t best = 0;
    uint32_t code = HAP_NO_PRED;
    const int mask = a.ring - 1, oldest = max(hi - a.ring, valid_from);
    for (int64_t e = p.e0; e < p.e1; ++e) {
        const uint32_t x = (e == p.e0) ? p.x0 : __ldg(a.in_edge + e);
        const int r = p.r2 - (int)(x >> HAP_W_SHIFT);
        if (r >= 0) {
            const int src = (int)(x & HAP_SLOT_MASK);
            const int32_t v = (src >= oldest) ? ring[(src & mask) * S + r] : a.dp[(int64_t)src * S + r];
            const int32_t cand = v + p.add;
            if (cand > best) { best = cand; code = x; }        // the winner's in-edge itself: source slot | weight << 30
        }
    }
    ring[(p.s & mask) * S + p.r2] = best;
    a.dp[(int64_t)p.s * S + p.r2] = best;
    a.pred[(int64_t)p.s * S + p.r2] = code;
}

__global__ void __launch_bounds__(HAP_THREADS, 1) hap_sweep_kernel(const HapSweepArgs a) {
    extern __shared__ int32_t hap_ring[];
    const int S = a.R + 1, T = (int)blockDim.x;     // narrow panels run with fewer warps: the per-level barrier is cheaper
    {   // level 0 keeps the initial 0 state (:50)
        const int hi0 = __ldg(a.level_off + 1);
        for (int c = threadIdx.x; c < min(hi0, a.ring) * S; c += T) hap_ring[c] = 0;
    }
    int valid_from = (__ldg(a.level_off + 1) > a.ring) ? __ldg(a.level_off + 1) : 0;   // slots >= valid_from (and young enough) are in the ring
    __syncthreads();
    // two-deep software pipeline over the levels: offsets of level l+2 and the first in-edge of level l+1 are
    // requested while level l is computed, so each of the two dependent graph loads has a whole level to arrive
    HapCell cur = hap_prefetch(a, 1, S, threadIdx.x);
    HapCell nxt = hap_stage_a(a, 2, S, threadIdx.x);
    for (int l = 1; l < a.L; ++l) {
        const HapCell nn = hap_stage_a(a, l + 2, S, threadIdx.x);
        hap_stage_b(a, nxt);
        const int lo = __ldg(a.level_off + l), hi = __ldg(a.level_off + l + 1);
        if (hi - lo > a.ring) {          // a level wider than the ring cannot use it (its own stores would collide)
            for (int c = threadIdx.x; c < (hi - lo) * S; c += T) {
                HapCell p = hap_prefetch(a, l, S, c);
                int32_t best = 0;
                uint32_t code = HAP_NO_PRED;
                for (int64_t e = p.e0; e < p.e1; ++e) {
                    const uint32_t x = __ldg(a.in_edge + e);
                    const int r = p.r2 - (int)(x >> HAP_W_SHIFT);
                    if (r >= 0) {
                        const int32_t cand = a.dp[(int64_t)(x & HAP_SLOT_MASK) * S + r] + p.add;
                        if (cand > best) { best = cand; code = x; }
                    }
                }
                a.dp[(int64_t)p.s * S + p.r2] = best;
                a.pred[(int64_t)p.s * S + p.r2] = code;
            }
            valid_from = hi;
        } else if (cur.has) {
            // sources older than hi - ring, and everything older than the last level wider than the ring, come from HBM/L2
            hap_cell(a, S, cur, hap_ring, hi, valid_from);
            const int cells = (hi - lo) * S;
            for (int c = threadIdx.x + T; c < cells; c += T) hap_cell(a, S, hap_prefetch(a, l, S, c), hap_ring, hi, valid_from);
        }
        __syncthreads();
        cur = nxt;
        nxt = nn;
    }
}

// Chain r: follow the predecessors from the last vertex (n-1) in layer r until a cell without one (:141-151);
// slots are written sink-first into path[r][0..len).  One dependent load per step.
__global__ void hap_trace_kernel(const uint32_t* pred, int32_t sink_slot, int32_t R, int32_t cap, int32_t* path, int32_t* path_len) {
    if (threadIdx.x != 0) return;
    const int r0 = blockIdx.x, S = R + 1;
    int32_t* out = path + (int64_t)r0 * cap;
    int32_t s = sink_slot, r = r0, len = 0;
    for (;;) {
        if (len < cap) out[len] = s;
        ++len;
        const uint32_t x = __ldcg(pred + (int64_t)s * S + r);
        if (x == HAP_NO_PRED) break;
        s = (int32_t)(x & HAP_SLOT_MASK);
        r -= (int)(x >> HAP_W_SHIFT);
    }
    path_len[r0] = len;
}

// Distinct colours on chain r (:77-99): set bit (r, colour) for every colour of every path vertex.
__global__ void hap_colour_mark_kernel(const int32_t* path, const int32_t* path_len, int32_t cap, const int64_t* col_off,
                                       const int32_t* col_val, int64_t words_per_r, unsigned int* bits) {
    const int r = blockIdx.y;
    const int len = min(path_len[r], cap);
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < len; t += gridDim.x * blockDim.x) {
        const int s = path[(int64_t)r * cap + t];
        for (int64_t c = col_off[s]; c < col_off[s + 1]; ++c) {
            const int32_t col = col_val[c];
            atomicOr(bits + (int64_t)r * words_per_r + (col >> 5), 1u << (col & 31));
        }
    }
}
__global__ void hap_colour_count_kernel(const unsigned int* bits, int64_t words_per_r, int32_t* colours_by_r) {
    const int r = blockIdx.x;
    int acc = 0;
    for (int64_t w = threadIdx.x; w < words_per_r; w += blockDim.x) acc += __popc(bits[(int64_t)r * words_per_r + w]);
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xFFFFFFFFu, acc, o);
    __shared__ int part[32];
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        int tot = 0;
        for (int i = 0; i < (int)(blockDim.x + 31) / 32; ++i) tot += part[i];
        colours_by_r[r] = tot;
    }
}

}  // namespace dg

using namespace dg;

struct dg_hap {
    int32_t n = 0, R = 0, L = 0, n_colours = 0, max_width = 0, max_indeg = 0;
    int64_t nE = 0;
    int32_t cap = 0, sink_slot = 0;
    int64_t words_per_r = 1;
    std::vector<int32_t> vtx_of_slot;
    DevBuf<int32_t> level_off, ncol, dp, path, path_len, colours;
    DevBuf<int64_t> in_off, col_off;
    DevBuf<uint32_t> in_edge;
    DevBuf<int32_t> col_val;
    DevBuf<uint32_t> pred;
    DevBuf<unsigned int> bits;
    cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
    float sweep_ms = 0.f, trace_ms = 0.f;
    int launches = 0;
    bool ran = false;
    uint64_t device_bytes = 0;
    ~dg_hap() { for (auto& e : ev) if (e) cudaEventDestroy(e); }
};

extern "C" {

int dg_hap_create(dg_ctx* ctx, int32_t n, const int64_t* adj_off, const int32_t* adj_dst, const uint8_t* adj_w,
                  const int64_t* col_off, const int32_t* col_val, int32_t n_colours, int32_t R, dg_hap** out) {
    if (!ctx || !out) return DG_ERR_ARG;
    *out = nullptr;
    if (n <= 0 || R < 0 || !adj_off || !col_off) return fail(ctx, DG_ERR_ARG, "dg_hap_create: empty graph or negative R");
    if ((uint32_t)n > HAP_SLOT_MASK) return fail(ctx, DG_ERR_ARG, "dg_hap_create: more than 2^30 vertices");
    DG_CUDA(ctx, cudaSetDevice(ctx->device));
    std::unique_ptr<dg_hap> d(new dg_hap());
    d->n = n; d->R = R; d->n_colours = n_colours;
    const int64_t nE = adj_off[n];
    d->nE = nE;
    // longest-path levels (edges go forward in the Kahn order the reference hands to the DP, :1256)
    std::vector<int32_t> level((size_t)n, 0);
    for (int32_t u = 0; u < n; ++u)
        for (int64_t e = adj_off[u]; e < adj_off[u + 1]; ++e) {
            const int32_t v = adj_dst[e];
            if (v <= u || v >= n) return fail(ctx, DG_ERR_ARG, "dg_hap_create: edge %d->%d is not forward in topological order", u, v);
            if (adj_w[e] > 3) return fail(ctx, DG_ERR_ARG, "dg_hap_create: edge weight %d > 3", (int)adj_w[e]);
            level[v] = std::max(level[v], level[u] + 1);
        }
    int32_t L = 0;
    for (int32_t v = 0; v < n; ++v) L = std::max(L, level[v] + 1);
    d->L = L;
    std::vector<int32_t> level_off((size_t)L + 1, 0);
    for (int32_t v = 0; v < n; ++v) ++level_off[level[v] + 1];
    for (int32_t l = 0; l < L; ++l) { d->max_width = std::max(d->max_width, level_off[l + 1]); level_off[l + 1] += level_off[l]; }
    std::vector<int32_t> slot((size_t)n), cursor(level_off.begin(), level_off.end() - 1);
    d->vtx_of_slot.resize((size_t)n);
    for (int32_t v = 0; v < n; ++v) { slot[v] = cursor[level[v]]++; d->vtx_of_slot[slot[v]] = v; }
    // in-edge CSR by destination slot.  Filling in (u ascending, adjacency order) and then moving, within
    // one (destination, u) run, the larger weights first gives the reference's visiting order
    // (u, then r = r2 - w ascending, then adjacency order).
    std::vector<int64_t> in_off((size_t)n + 1, 0);
    for (int64_t e = 0; e < nE; ++e) ++in_off[slot[adj_dst[e]] + 1];
    for (int32_t s = 0; s < n; ++s) { d->max_indeg = std::max<int64_t>(d->max_indeg, in_off[s + 1]); in_off[s + 1] += in_off[s]; }
    std::vector<uint32_t> in_edge((size_t)nE);
    {
        std::vector<int64_t> cur(in_off.begin(), in_off.end() - 1);
        for (int32_t u = 0; u < n; ++u)
            for (int64_t e = adj_off[u]; e < adj_off[u + 1]; ++e)
                in_edge[cur[slot[adj_dst[e]]]++] = (uint32_t)slot[u] | ((uint32_t)adj_w[e] << HAP_W_SHIFT);
        for (int32_t s = 0; s < n; ++s) {
            int64_t a = in_off[s];
            const int64_t b = in_off[s + 1];
            while (a < b) {
                int64_t c = a + 1;
                while (c < b && (in_edge[c] & HAP_SLOT_MASK) == (in_edge[a] & HAP_SLOT_MASK)) ++c;
                if (c - a > 1)
                    std::stable_sort(in_edge.begin() + a, in_edge.begin() + c,
                                     [](uint32_t x, uint32_t y) { return (x >> HAP_W_SHIFT) > (y >> HAP_W_SHIFT); });
                a = c;
            }
        }
    }
    std::vector<int32_t> ncol((size_t)n);
    std::vector<int64_t> col_off_s((size_t)n + 1, 0);
    for (int32_t s = 0; s < n; ++s) {
        const int32_t v = d->vtx_of_slot[s];
        ncol[s] = (int32_t)(col_off[v + 1] - col_off[v]);
        col_off_s[s + 1] = col_off_s[s] + ncol[s];
    }
    std::vector<int32_t> col_val_s((size_t)col_off_s[n]);
    for (int32_t s = 0; s < n; ++s) {
        const int32_t v = d->vtx_of_slot[s];
        for (int64_t c = col_off[v]; c < col_off[v + 1]; ++c) {
            if (col_val[c] < 0 || col_val[c] >= n_colours) return fail(ctx, DG_ERR_ARG, "dg_hap_create: colour id %d out of range", col_val[c]);
            col_val_s[col_off_s[s] + (c - col_off[v])] = col_val[c];
        }
    }
    d->sink_slot = slot[n - 1];
    d->cap = L + 1;
    d->words_per_r = std::max<int64_t>(1, ((int64_t)n_colours + 31) / 32);

    cudaStream_t st = ctx->stream;
    const size_t cells = (size_t)n * (size_t)(R + 1);
    DG_CUDA(ctx, d->level_off.upload(level_off.data(), level_off.size(), st));
    DG_CUDA(ctx, d->in_off.upload(in_off.data(), in_off.size(), st));
    DG_CUDA(ctx, d->in_edge.upload(in_edge.data(), in_edge.size(), st));
    DG_CUDA(ctx, d->ncol.upload(ncol.data(), ncol.size(), st));
    DG_CUDA(ctx, d->col_off.upload(col_off_s.data(), col_off_s.size(), st));
    DG_CUDA(ctx, d->col_val.upload(col_val_s.data(), col_val_s.size(), st));
    DG_CUDA(ctx, d->dp.alloc(cells));
    DG_CUDA(ctx, d->pred.alloc(cells));
    DG_CUDA(ctx, d->path.alloc((size_t)(R + 1) * (size_t)d->cap));
    DG_CUDA(ctx, d->path_len.alloc((size_t)R + 1));
    DG_CUDA(ctx, d->colours.alloc((size_t)R + 1));
    DG_CUDA(ctx, d->bits.alloc((size_t)(R + 1) * (size_t)d->words_per_r));
    for (auto& e : d->ev) DG_CUDA(ctx, cudaEventCreate(&e));
    DG_CUDA(ctx, cudaStreamSynchronize(st));
    d->device_bytes = d->level_off.bytes() + d->in_off.bytes() + d->in_edge.bytes() + d->ncol.bytes() + d->col_off.bytes() +
                      d->col_val.bytes() + d->dp.bytes() + d->pred.bytes() + d->path.bytes() + d->bits.bytes();
    *out = d.release();
    return DG_OK;
}

int dg_hap_run(dg_ctx* ctx, dg_hap* d) {
    if (!ctx || !d) return DG_ERR_ARG;
    DG_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const size_t cells = (size_t)d->n * (size_t)(d->R + 1);
    d->launches = 0;
    DG_CUDA(ctx, cudaEventRecord(d->ev[0], st));
    DG_CUDA(ctx, cudaMemsetAsync(d->dp.p, 0, cells * sizeof(int32_t), st));            // :50
    DG_CUDA(ctx, cudaMemsetAsync(d->pred.p, 0xFF, cells * sizeof(uint32_t), st));       // :51-52
    DG_CUDA(ctx, cudaMemsetAsync(d->bits.p, 0, d->bits.bytes(), st));
    HapSweepArgs a;
    a.level_off = d->level_off.p; a.in_off = d->in_off.p; a.in_edge = d->in_edge.p; a.ncol = d->ncol.p;
    a.dp = d->dp.p; a.pred = d->pred.p; a.L = d->L; a.R = d->R;
    // score ring: as many slots (power of two) as fit ~160 KB of shared memory
    int ring = 4096;
    while (ring > 1 && (size_t)ring * (size_t)(d->R + 1) * 4 > (size_t)160 * 1024) ring >>= 1;
    a.ring = ring;
    const size_t ring_bytes = (size_t)ring * (size_t)(d->R + 1) * 4;
    DG_CUDA(ctx, cudaFuncSetAttribute((const void*)hap_sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ring_bytes));
    const int threads = ((int64_t)d->max_width * (d->R + 1) <= 512) ? 256 : HAP_THREADS;
    hap_sweep_kernel<<<1, threads, ring_bytes, st>>>(a);
    ++d->launches;
    DG_CUDA(ctx, cudaGetLastError());
    DG_CUDA(ctx, cudaEventRecord(d->ev[1], st));
    hap_trace_kernel<<<d->R + 1, 32, 0, st>>>(d->pred.p, d->sink_slot, d->R, d->cap, d->path.p, d->path_len.p);
    ++d->launches;
    DG_CUDA(ctx, cudaGetLastError());
    const int bx = std::max(1, std::min(64, (d->cap + 255) / 256));
    hap_colour_mark_kernel<<<dim3(bx, d->R + 1), 256, 0, st>>>(d->path.p, d->path_len.p, d->cap, d->col_off.p, d->col_val.p,
                                                             d->words_per_r, d->bits.p);
    ++d->launches;
    DG_CUDA(ctx, cudaGetLastError());
    hap_colour_count_kernel<<<d->R + 1, 256, 0, st>>>(d->bits.p, d->words_per_r, d->colours.p);
    ++d->launches;
    DG_CUDA(ctx, cudaGetLastError());
    DG_CUDA(ctx, cudaEventRecord(d->ev[2], st));
    d->ran = true;
    return DG_OK;
}

int dg_hap_result(dg_ctx* ctx, dg_hap* d, int32_t* colours_by_r, int32_t* path_len) {
    if (!ctx || !d || !d->ran) return fail(ctx, DG_ERR_ARG, "dg_hap_result: dg_hap_run has not been called");
    DG_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    std::vector<int32_t> len((size_t)d->R + 1);
    if (colours_by_r) DG_CUDA(ctx, cudaMemcpyAsync(colours_by_r, d->colours.p, ((size_t)d->R + 1) * 4, cudaMemcpyDeviceToHost, st));
    DG_CUDA(ctx, cudaMemcpyAsync(len.data(), d->path_len.p, ((size_t)d->R + 1) * 4, cudaMemcpyDeviceToHost, st));
    DG_CUDA(ctx, cudaStreamSynchronize(st));
    DG_CUDA(ctx, cudaEventElapsedTime(&d->sweep_ms, d->ev[0], d->ev[1]));
    DG_CUDA(ctx, cudaEventElapsedTime(&d->trace_ms, d->ev[1], d->ev[2]));
    for (int r = 0; r <= d->R; ++r) {
        if (len[r] > d->cap) return fail(ctx, DG_ERR_CAPACITY, "dg_hap_result: path %d longer than the level count", r);
        if (path_len) path_len[r] = len[r];
    }
    return DG_OK;
}

int dg_hap_path(dg_ctx* ctx, dg_hap* d, int32_t r, int32_t* path, int32_t cap, int32_t* len_out) {
    if (!ctx || !d || !d->ran || r < 0 || r > d->R) return fail(ctx, DG_ERR_ARG, "dg_hap_path: bad arguments");
    DG_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    int32_t len = 0;
    DG_CUDA(ctx, cudaMemcpyAsync(&len, d->path_len.p + r, 4, cudaMemcpyDeviceToHost, st));
    DG_CUDA(ctx, cudaStreamSynchronize(st));
    if (len_out) *len_out = len;
    if (len > cap || len > d->cap) return fail(ctx, DG_ERR_CAPACITY, "dg_hap_path: path of %d vertices does not fit capacity %d", len, cap);
    std::vector<int32_t> tmp((size_t)len);
    DG_CUDA(ctx, cudaMemcpyAsync(tmp.data(), d->path.p + (size_t)r * d->cap, (size_t)len * 4, cudaMemcpyDeviceToHost, st));
    DG_CUDA(ctx, cudaStreamSynchronize(st));
    for (int32_t t = 0; t < len; ++t) path[t] = d->vtx_of_slot[tmp[len - 1 - t]];   // source first (:153), vertex ids
    return DG_OK;
}

int dg_hap_stats(dg_ctx* ctx, dg_hap* d, dg_hap_stats_t* out) {
    if (!d || !out) return DG_ERR_ARG;
    memset(out, 0, sizeof *out);
    out->cell_updates = (uint64_t)(d->R + 1) * (uint64_t)d->nE;
    out->cells = (uint64_t)(d->R + 1) * (uint64_t)d->n;
    out->algo_bytes = (uint64_t)(d->R + 1) * (8ull * (uint64_t)d->n + 8ull * (uint64_t)d->nE);
    out->device_bytes = d->device_bytes;
    out->n_levels = d->L; out->n_vertices = d->n; out->max_width = d->max_width; out->max_indegree = d->max_indeg;
    out->launches = d->launches; out->sweep_ms = d->sweep_ms; out->traceback_ms = d->trace_ms;
    (void)ctx;
    return DG_OK;
}

void dg_hap_destroy(dg_ctx* ctx, dg_hap* d) {
    if (ctx) cudaSetDevice(ctx->device);
    delete d;
}

int dg_dp_haploid(dg_ctx* ctx, int32_t n, const int64_t* adj_off, const int32_t* adj_dst, const uint8_t* adj_w,
                  const int64_t* col_off, const int32_t* col_val, int32_t n_colours, int32_t R, int32_t* colours_by_r,
                  int64_t* path_off, int32_t** paths) {
    if (!ctx || !colours_by_r || !path_off || !paths) return DG_ERR_ARG;
    *paths = nullptr;
    dg_hap* d = nullptr;
    int rc = dg_hap_create(ctx, n, adj_off, adj_dst, adj_w, col_off, col_val, n_colours, R, &d);
    if (rc) return rc;
    std::vector<int32_t> len((size_t)R + 1);
    rc = dg_hap_run(ctx, d);
    if (!rc) rc = dg_hap_result(ctx, d, colours_by_r, len.data());
    if (!rc) {
        path_off[0] = 0;
        for (int r = 0; r <= R; ++r) path_off[r + 1] = path_off[r] + len[r];
        int32_t* buf = (int32_t*)malloc(std::max<size_t>(1, (size_t)path_off[R + 1]) * sizeof(int32_t));
        if (!buf) rc = fail(ctx, DG_ERR_NOMEM, "dg_dp_haploid: host allocation failed");
        for (int r = 0; r <= R && !rc; ++r) rc = dg_hap_path(ctx, d, r, buf + path_off[r], len[r], nullptr);
        if (rc) free(buf); else *paths = buf;
    }
    dg_hap_destroy(ctx, d);
    return rc;
}

}  // extern "C"
