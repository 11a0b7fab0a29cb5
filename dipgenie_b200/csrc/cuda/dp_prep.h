// dp_prep.h — host-side planning for the diploid DP sweep (no CUDA types; also compiled into the
// CPU kernel-logic emulator under tests/emu/).
//
// Turns the levelized ExpandedGraph (reference: src/ExpandedGraph.hpp:16-26 after
// strict_bfs_levelize_and_reorder :269-409, the input of
// Approximator::diploid_dp_approximation_solver, src/approximator.cpp:362) into the *gather* form
// the kernels consume:
//   * in-edge CSR per destination vertex, entries (source position in previous level, weight),
//     ascending source position — the reference's scatter loop over (i, j, e1, e2)
//     (approximator.cpp:627-701) becomes "every destination cell takes the first strict maximum over
//     in-edges(i') x in-edges(j')", which needs no locks and is order-free;
//   * per-transition colour bit-masks over the local colour universe of levels l and l+1, split
//     hom/het with colour_is_hom (approximator.cpp:431-453), so that the reference's two 4-way sorted
//     merges per edge pair (inter_size_union2x2 / symdiff_size_union2x2, :269-311, :604-624) become
//     popcounts:  delta = popc((Hs[i]|Hs[j]) & (Hd[i']|Hd[j'])) + popc((Ts[i]|Ts[j]) ^ (Td[i']|Td[j']));
//     a device kernel evaluates them once per (e1,e2) into the pair-score matrix of the transition;
//   * one task stream per CTA (dp_cell.h: TaskHdr) with the monotone barrier targets of the
//     persistent sweep.
#pragma once
#include <cstdint>
#include <memory_resource>
#include <string>
#include <vector>

#include "dp_cell.h"

namespace dg {

struct DipGraphView {
    int32_t n_levels = 0;
    const int32_t* level_off = nullptr;   // [L+1]
    const int64_t* adj_off = nullptr;     // [V+1]
    const int32_t* adj_dst = nullptr;     // [E]
    const uint8_t* adj_w = nullptr;       // [E]
    const int64_t* col_off = nullptr;     // [V+1]
    const int32_t* col_val = nullptr;
    const uint8_t* colour_is_hom = nullptr;
    int32_t n_colours = 0;
    int32_t R = 0;
};

constexpr uint32_t IN_POS_MASK = 0xFFFFu;   // in_edge = pos | (w << 16)
constexpr int IN_W_SHIFT = 16;

struct SweepShape {              // kernel geometry the plan is made for
    int grid = 1;                // CTAs in the cooperative grid
    int threads = 512;           // compute threads per CTA
    int tile_cells = 16384;      // int32 cells per shared-memory layer tile (two tiles)
    int slot_bytes = 4096;       // bytes per task slot in shared memory (header + record + delta slice)
    int64_t delta_budget = (int64_t)8 << 30;   // bytes of pair-score matrices to materialise at most
    int delta_max_in = 4096;     // widest in-edge count that still gets a matrix
    int lane_rc = LANE_RC_SMALL; // layers per lane in the lane form (LANE_RC_SMALL or LANE_RC_BIG)
    bool allow_long = false;     // lane form may take destinations with more than 32 in-edges (needs packed keys' kernel)
    // Row-sharded sweep over `replicas` GPUs (one process each): the plan is made for replicas x grid global CTAs and
    // keeps the streams of this rank's (global CTA c -> rank c % replicas, local CTA c / replicas).  Narrow transitions
    // run on the local CTA 0 of every rank; wide transitions are row-split over all global CTAs and every CTA pushes
    // its rows to the peers (TK_PUSH); every arrival is broadcast to all ranks' counters.
    int replicas = 1;
    int rank = 0;
};

struct DipPlan {
    // The arrays that go to the device wholesale take their storage from `res` (default: the heap).  The batch
    // entry point hands every planner a block of page-locked host memory, so that the plan is written once, into
    // memory the copy engine reads directly.
    explicit DipPlan(std::pmr::memory_resource* res = std::pmr::get_default_resource())
        : in_off(res), in_edge(res), in_dst(res), masks(res), records(res), tasks(res) {}
    int32_t L = 0, V = 0, R = 0;
    int32_t kmax = 0, max_indeg = 0, Wmax = 0;
    int64_t n_in = 0;
    std::vector<int32_t> level_off;    // [L+1]
    std::pmr::vector<int32_t> in_off;  // [V+1]
    std::pmr::vector<uint32_t> in_edge;  // [n_in]   source position | weight << 16
    std::pmr::vector<uint16_t> in_dst;  // [n_in]   destination position of the in-edge (for the delta kernel)
    std::vector<int32_t> lvlW;         // [L]   64-bit mask words of transition l (0 = no colours)
    std::vector<int64_t> msrc_off;     // [L]   offset (u64 words) of level-l source masks
    std::vector<int64_t> mdst_off;     // [L]   offset of level-(l+1) destination masks
    std::pmr::vector<uint64_t> masks;
    std::vector<int64_t> pred_off;     // [L+1] offset of level l's predecessor codes
    // --- task plan (plan_tasks) ---
    std::vector<int32_t> P;            // [L]   CTAs taking part in transition l
    std::vector<uint32_t> bar_target;  // [L]   arrivals that must be visible once transition l is complete
    std::vector<uint8_t> bar_edge;     // [L]   1 = a grid-level barrier follows transition l
    std::vector<uint8_t> narrow;       // [L]   1 = both layers of transition l fit CTA 0's shared-memory tiles
    std::vector<int64_t> rec_off;      // [L]   byte offset of transition l's record in `records` (-1: none)
    std::pmr::vector<uint8_t> records; // packed records (16-byte aligned)
    std::vector<int64_t> delta_off;    // [L]   u16-element offset of transition l's pair-score matrix (-1: none)
    std::vector<int32_t> delta_list;   // transitions that own a matrix, ascending
    int64_t delta_elems = 0;           // total u16 elements (every matrix start is 16-byte aligned)
    std::pmr::vector<TaskHdr> tasks;   // all CTAs' streams, concatenated in CTA order
    std::vector<int64_t> task_begin;   // [grid+1]
    int32_t grid = 1;
    int64_t n_narrow = 0, n_wide = 0, n_tasks_global = 0, n_tasks_masks = 0;
    // accounting
    uint64_t cell_updates = 0;         // U = (R+1) * sum_l E_l^2
    uint64_t cells = 0;                // C = (R+1) * sum_{l>=1} k_l^2
    uint64_t algo_bytes = 0;           // B = (R+1) * sum_l (4 k_l^2 + 5 k_{l+1}^2)
    int64_t value_bound = 0;           // no DP value exceeds this: sum over transitions of the distinct colours present
    std::string error;
};

// Builds the gather-form arrays; returns false and sets plan.error on malformed input.
bool build_dip_plan(const DipGraphView& g, DipPlan& plan);

// Traceback checkpoints (dp_diploid.cu: dip_anc_kernel walks EVERY cell of a checkpoint level back to the previous
// checkpoint, so narrow checkpoint levels are cheap ones): cp[0] = L-1 (sink level), then, about every T levels,
// the narrowest level of the window [cur - 3T/2, cur - T/2]; the list ends with level 0.
std::vector<int32_t> choose_checkpoints(const std::vector<int32_t>& level_off, int T);

// Compiles the sweep into per-CTA task streams for the given kernel geometry: narrow/wide placement of
// every layer, row partition of wide transitions, slot-sized sub-tasks, barrier schedule, packed records
// and the layout of the pair-score matrices.
void plan_tasks(DipPlan& plan, const SweepShape& shape);

}  // namespace dg
