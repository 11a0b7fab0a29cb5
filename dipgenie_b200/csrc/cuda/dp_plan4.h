// dp_plan4.h — host-side planning of the level-program sweep (engine v4; see dp_prog.h).  No CUDA types: also
// compiled into the CPU kernel-logic emulator under tests/emu/.
//
// From the gather-form plan of dp_prep.h (in-edge CSR by destination in ascending source position, colour masks)
// this computes what is O(V): the slot of every vertex (passive vertices inherit their predecessor's, the others get
// a fresh one; dp_prog.h), the placement of every level (shared-memory tile of CTA 0 or HBM tile), the
// S1a/S1p/M/Z class tables, the closed-form descriptor counts, the program and predecessor-code offsets, and two
// directories with their barrier schedules: every transition (the checksum variant needs them all) and the transitions
// that have something to do (the timed variant).  The O(sum E^2) part — the descriptors themselves — is written by the
// device (dp_sweep4.cuh: prog_fill_kernel) with the per-descriptor functions of dp_prog.h; prog_fill_level_host runs
// the same functions serially for the emulator and for the test that compares the device-built program byte for byte.
#pragma once
#include <cstdint>
#include <memory_resource>
#include <string>
#include <vector>

#include "dp_prep.h"
#include "dp_prog.h"

namespace dg {

struct Plan4Dir {
    explicit Plan4Dir(std::pmr::memory_resource* res) : dir(res) {}
    std::pmr::vector<ProgDir> dir;       // transitions in level order
    std::vector<int32_t> wide_list;      // indices into `dir` of the transitions every CTA takes part in
    uint32_t final_target = 0;           // barrier arrivals once the last transition is complete
    int32_t n = 0;                       // dir.size() (kept when the arrays' page-locked storage has been handed back)
};

struct Plan4 {
    // The arrays that go to the device wholesale take their storage from `res` (the batch entry point hands every planner
    // a block of page-locked host memory, like DipPlan).
    explicit Plan4(std::pmr::memory_resource* res = std::pmr::get_default_resource())
        : cls_list(res), vinfo(res), vslot(res), lvl_dom(res), mpre(res), mpre_off(res), lvl_n1(res), lvl_np(res), lvl_m(res), lvl_z(res),
          lvl_dm(res), hdr(res), tflags(res), full(res), timed(res), prog_off(res), pred_off(res) {}
    Sweep4Shape shape;
    int32_t L = 0, R = 0, rc = 10, nchunk = 1, RL = 0;   // RL = nchunk * rc layers are computed (>= R+1)
    std::pmr::vector<uint16_t> cls_list;      // [V]   per level: S1a | S1p | M | Z positions
    std::pmr::vector<uint32_t> vinfo;         // [V]   rank in class | class << 30 (S1 ranks over S1a | S1p)
    std::pmr::vector<uint16_t> vslot;         // [V]   slot of the vertex in its level's tile
    std::pmr::vector<uint8_t> lvl_dom;        // [L]   0: shared-memory tile, 1: HBM tile
    std::pmr::vector<uint32_t> mpre;          // per level m+1 entries: prefix of in-degrees over M
    std::pmr::vector<int64_t> mpre_off;       // [L]
    std::pmr::vector<uint32_t> lvl_n1, lvl_np, lvl_m, lvl_z, lvl_dm;   // [L]
    std::pmr::vector<ProgHdr> hdr;            // [L-1] header of transition l
    std::pmr::vector<uint32_t> tflags;        // [L-1] PF_COMPACT / SRC_SMEM / DST_SMEM / RELOCATE of transition l
    Plan4Dir full, timed;
    std::pmr::vector<uint64_t> prog_off;      // [L]   byte offset of transition l's program (prog_off[L-1] = total)
    std::pmr::vector<int64_t> pred_off;       // [L+1] u16 elements: codes of level l
    uint64_t prog_bytes = 0;
    int64_t pred_elems = 0;
    int32_t hstride = 1;                 // slots per row of the HBM tile
    int32_t gcs = 0;                     // words per cell of the HBM tile (cell-major: 2 dead padding layers, RL layers, padded to 32 bytes)
    int64_t gtile_cells = 0;             // words of the HBM tile: hstride^2 cells of gcs words
    int64_t n_smem_trans = 0;            // transitions with both levels in shared memory
    int64_t n_relocate = 0, n_skipped = 0;
    uint64_t cells_written = 0, cells_total = 0;   // destination cells in the programs / of the levels (per layer)
    uint32_t max_cand = 0;
    uint32_t max_giant = 0;              // most cells of more than PROG_KEY_CAND candidates in one level
    // Hands the storage of every array back (after the upload; the scalars and wide lists stay).
    void release_arrays() {
        auto drop = [](auto& v) { v.clear(); v.shrink_to_fit(); };
        drop(cls_list); drop(vinfo); drop(vslot); drop(lvl_dom); drop(mpre); drop(mpre_off); drop(lvl_n1); drop(lvl_np); drop(lvl_m);
        drop(lvl_z); drop(lvl_dm); drop(hdr); drop(tflags); drop(full.dir); drop(timed.dir); drop(prog_off); drop(pred_off);
    }
    uint32_t sink_cell = 0;              // cell (0,0) of the last level in its tile
    bool last_smem = true;               // ... which is the shared-memory one
};

// Returns false (with `why`) when the problem is outside what the packed-key level program covers; the caller then
// takes the task-stream engine.  `out` must be freshly constructed (with the memory resource its arrays shall use).
bool plan4_build(const DipPlan& p, const DipGraphView& g, const Sweep4Shape& shape, int rc, Plan4& out, std::string& why);

// The view of transition l the dp_prog.h functions take.
ProgLevelIn plan4_level_in(const DipPlan& p, const Plan4& q, int l);

// Writes transition l's program (header + sections) at `out` (prog_off[l+1] - prog_off[l] bytes).
void prog_fill_level_host(const DipPlan& p, const Plan4& q, int l, uint8_t* out);

}  // namespace dg
