// dp_prep.h — host-side planning for the diploid DP sweep (no CUDA types; also compiled into the
// CPU kernel-logic emulator under tests/emu/).
//
// Turns the levelized ExpandedGraph (reference: src/ExpandedGraph.hpp:16-26 after
// strict_bfs_levelize_and_reorder :269-409, the input of
// Approximator::diploid_dp_approximation_solver, src/approximator.cpp:362) into the *gather* form
// the kernels consume:
//   * in-edge CSR per destination vertex, entries (source position in previous level, weight),
//     ascending source position — the reference's scatter loop over (i, j, e1, e2)
//     (approximator.cpp:627-701) becomes "every destination cell takes the lexicographic max of
//     (value, -i, -j) over in-edges(i') x in-edges(j')", which needs no locks and is order-free;
//   * per-transition colour bit-masks over the local colour universe of levels l and l+1, split
//     hom/het with colour_is_hom (approximator.cpp:431-453), so that the reference's two 4-way sorted
//     merges per edge pair (inter_size_union2x2 / symdiff_size_union2x2, :269-311, :604-624) become
//     popcounts:  delta = popc((Hs[i]|Hs[j]) & (Hd[i']|Hd[j'])) + popc((Ts[i]|Ts[j]) ^ (Td[i']|Td[j']));
//   * the per-transition participant count and the monotone barrier targets of the persistent sweep.
#pragma once
#include <cstdint>
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
    int cells_per_cta = 2048;    // destination cells one CTA takes per transition before more CTAs join
    int tile_cells = 16384;      // int32 cells per shared-memory layer tile (two tiles)
    int stage_bytes = 16384;     // bytes per record stage in shared memory
};

struct DipPlan {
    int32_t L = 0, V = 0, R = 0;
    int32_t kmax = 0, max_indeg = 0, Wmax = 0;
    int64_t n_in = 0;
    std::vector<int32_t> level_off;    // [L+1]
    std::vector<int32_t> in_off;       // [V+1]
    std::vector<uint32_t> in_edge;     // [n_in]
    std::vector<int32_t> lvlW;         // [L]   64-bit mask words of transition l (0 = no colours)
    std::vector<int64_t> msrc_off;     // [L]   offset (u64 words) of level-l source masks
    std::vector<int64_t> mdst_off;     // [L]   offset of level-(l+1) destination masks
    std::vector<uint64_t> masks;
    std::vector<int64_t> pred_off;     // [L+1] offset of level l's predecessor codes
    std::vector<int32_t> P;            // [L]   CTAs taking part in transition l
    std::vector<uint32_t> bar_target;  // [L]   arrivals that must be visible once transition l is complete
    std::vector<uint8_t> bar_edge;     // [L]   1 = a grid-level barrier follows transition l
    std::vector<uint8_t> mode;         // [L]   MODE_* of transition l
    std::vector<uint16_t> flags;       // [L]   REC_* of transition l
    std::vector<int64_t> rec_off;      // [L]   byte offset of transition l's record in `records` (-1: none)
    std::vector<uint8_t> records;      // packed records of all FAST/STAGED transitions (16-byte aligned)
    int64_t n_fast = 0, n_staged = 0, n_global = 0;
    // accounting
    uint64_t cell_updates = 0;         // U = (R+1) * sum_l E_l^2
    uint64_t cells = 0;                // C = (R+1) * sum_{l>=1} k_l^2
    uint64_t algo_bytes = 0;           // B = (R+1) * sum_l (4 k_l^2 + 5 k_{l+1}^2)
    std::string error;
};

// Builds everything except P / bar_*; returns false and sets plan.error on malformed input.
bool build_dip_plan(const DipGraphView& g, DipPlan& plan);

// Chooses, for the given kernel geometry, the mode (FAST: CTA 0 alone with both layers in shared
// memory; STAGED: P CTAs, metadata staged through shared memory, layers in HBM/L2; GLOBAL: metadata
// read in place) and the participants of every transition, derives the monotone-counter barrier
// schedule, and packs the records.
void plan_sweep(DipPlan& plan, const SweepShape& shape);

}  // namespace dg
