// ref_hook.h — TEST INFRASTRUCTURE ONLY.
// Declarations of the dump hooks that oracle/Makefile splices (with sed, into a
// scratch copy under oracle/_ref/instr/, never into the repo) into the
// reference's src/approximator.cpp so that the *reference's own* intermediate
// products (the expanded graph handed to the DP, the per-level DP state
// checksums, the recombination-edge lists, the haploid path) can be written to a
// DGD1 file and compared with this repo's host glue, oracle port and CUDA path.
// The hooks only read reference state; the instrumented binary's FASTA output
// is checked to be byte-identical to the unmodified binary's (tests/test_ref_pin.py).
#pragma once
#include <cstdint>
#include <utility>
#include <vector>

class ExpandedGraph;

void dg_ref_dump_haploid_input(const ExpandedGraph& g, int R);
void dg_ref_dump_diploid_input(const ExpandedGraph& g, int R, const std::vector<bool>& color_homo_bv);
void dg_ref_dump_haploid_result(const std::vector<int>& colors_by_r, int best_r,
                                const std::vector<int>& path, const std::vector<int>& path_original);
void dg_ref_dump_diploid_result(const std::vector<std::pair<int, int>>& p1_edges,
                                const std::vector<std::pair<int, int>>& p2_edges, int value, int s_het);
void dg_ref_level_checksum_push(int level, uint64_t checksum, uint64_t n_live);

// Per-level checksum of the DP layer just produced (reference: approximator.cpp:704-706,
// right after dp_cur.swap(dp_next)).  FNV-style fold of (flat index, value, pred_i, pred_j)
// over live cells; the oracle port computes the same fold so the first divergent level
// can be located without dumping 0.4 G cells.
template <class Buf>
inline void dg_ref_level_done(int level, const Buf& buf, int32_t neg_inf) {
    uint64_t h = 1469598103934665603ull, live = 0;
    for (std::size_t t = 0; t < buf.size(); ++t) {
        const auto& e = buf[t];
        if (e.value == neg_inf) continue;
        ++live;
        uint64_t x = (uint64_t)t * 0x9E3779B97F4A7C15ull;
        x ^= (uint64_t)(uint32_t)e.value * 0xC2B2AE3D27D4EB4Full;
        x ^= ((uint64_t)(uint32_t)e.pred_i << 32 | (uint32_t)e.pred_j) * 0x165667B19E3779F9ull;
        x ^= x >> 29;
        h += x * 0xBF58476D1CE4E5B9ull;
    }
    dg_ref_level_checksum_push(level, h, live);
}
