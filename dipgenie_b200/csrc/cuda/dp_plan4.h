// dp_plan4.h — host-side planning of the level-program sweep (engine v4; see dp_prog.h).  No CUDA types: also
// compiled into the CPU kernel-logic emulator under tests/emu/.
//
// From the gather-form plan of dp_prep.h (in-edge CSR by destination in ascending source position, colour masks)
// this computes what is O(V): the S1/M/Z class tables of every level, the closed-form descriptor counts, the
// program and predecessor-code offsets, the placement of every layer (shared memory of CTA 0 or HBM/L2) and the
// barrier schedule.  The O(sum E^2) part — the descriptors themselves — is written by the device
// (dp_sweep4.cu: prog_fill_kernel) with the per-descriptor functions of dp_prog.h; prog_fill_level_host runs the same
// functions serially for the emulator and for the test that compares the device-built program byte for byte.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "dp_prep.h"
#include "dp_prog.h"

namespace dg {

struct Plan4 {
    Sweep4Shape shape;
    int32_t L = 0, R = 0, rc = 10, nchunk = 1, RL = 0;   // RL = nchunk * rc layers are computed (>= R+1)
    std::vector<uint16_t> cls_list;      // [V]   per level: S1 | M | Z positions
    std::vector<uint32_t> vinfo;         // [V]   rank in class | class << 30
    std::vector<uint32_t> mpre;          // per level m+1 entries: prefix of in-degrees over M
    std::vector<int64_t> mpre_off;       // [L]
    std::vector<uint32_t> lvl_n1, lvl_m, lvl_z, lvl_dm;   // [L]
    std::vector<ProgHdr> hdr;            // [L-1] header of transition l
    std::vector<ProgDir> dir;            // [L-1]
    std::vector<uint64_t> prog_off;      // [L]   byte offset of transition l's program (prog_off[L-1] = total)
    std::vector<int64_t> pred_off;       // [L+1] u16 elements: codes of level l
    std::vector<int32_t> wide_list;      // transitions every CTA takes part in
    uint64_t prog_bytes = 0;
    int64_t pred_elems = 0;
    int64_t gpad = 0, gtile_cells = 0;   // HBM tile: gpad dead cells, then RL layers of the widest HBM-resident level
    int64_t n_smem_trans = 0;            // transitions with both layers in shared memory
    uint32_t max_cand = 0;
    uint32_t final_target = 0;           // barrier arrivals once the last transition is complete
};

// Returns false (with `why`) when the problem is outside what the packed-key level program covers; the caller then
// takes the task-stream engine.
bool plan4_build(const DipPlan& p, const Sweep4Shape& shape, int rc, Plan4& out, std::string& why);

// The view of transition l the dp_prog.h functions take.
ProgLevelIn plan4_level_in(const DipPlan& p, const Plan4& q, int l);

// Writes transition l's program (header + sections) at `out` (prog_off[l+1] - prog_off[l] bytes).
void prog_fill_level_host(const DipPlan& p, const Plan4& q, int l, uint8_t* out);

}  // namespace dg
