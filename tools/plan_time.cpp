// plan_time.cpp — host-only timing of the diploid sweep planner (build_dip_plan + plan_tasks) on a dumped level graph.
//   g++ -O3 -fopenmp -std=c++17 -I dipgenie_b200/csrc/cuda tools/plan_time.cpp dipgenie_b200/csrc/cuda/dp_prep.cpp -o /tmp/plan_time
//   python tools/plan_time.py tests/golden/mhc4_chm13_dipin.npz   (dumps the arrays, builds and runs this)
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "dp_prep.h"
#include "dp_plan4.h"
#include <string>
using namespace dg;
template <class T> static std::vector<T> rd(const char* dir, const char* name) {
    char path[512]; snprintf(path, sizeof path, "%s/%s.bin", dir, name);
    FILE* f = fopen(path, "rb"); if (!f) { perror(path); exit(1); }
    fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
    std::vector<T> v((size_t)n / sizeof(T)); if (fread(v.data(), 1, (size_t)n, f) != (size_t)n) exit(1); fclose(f); return v;
}
static double now() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main(int argc, char** argv) {
    const char* dir = argv[1]; const int R = argc > 2 ? atoi(argv[2]) : 18; const int reps = argc > 3 ? atoi(argv[3]) : 5;
    auto level_off = rd<int32_t>(dir, "level_off"); auto adj_off = rd<int64_t>(dir, "adj_off"); auto adj_dst = rd<int32_t>(dir, "adj_dst");
    auto adj_w = rd<uint8_t>(dir, "adj_w"); auto col_off = rd<int64_t>(dir, "col_off"); auto col_val = rd<int32_t>(dir, "col_val");
    auto hom = rd<uint8_t>(dir, "colour_is_hom");
    DipGraphView g; g.n_levels = (int32_t)level_off.size() - 1; g.level_off = level_off.data(); g.adj_off = adj_off.data();
    g.adj_dst = adj_dst.data(); g.adj_w = adj_w.data(); g.col_off = col_off.data(); g.col_val = col_val.data();
    g.colour_is_hom = hom.data(); g.n_colours = (int32_t)hom.size(); g.R = R;
    for (int r = 0; r < reps; ++r) {
        const double t0 = now();
        DipPlan p; if (!build_dip_plan(g, p)) { fprintf(stderr, "%s\n", p.error.c_str()); return 1; }
        const double t1 = now();
        SweepShape sh; sh.grid = 6; sh.threads = 480; sh.lane_rc = LANE_RC_BIG; sh.allow_long = true;
        plan_tasks(p, sh);
        const double t2 = now();
        size_t bytes = p.tasks.size() * sizeof(TaskHdr) + p.records.size() + p.masks.size() * 8 + p.in_edge.size() * 4 + p.in_dst.size() * 2 +
                       p.in_off.size() * 4 + (p.msrc_off.size() + p.mdst_off.size() + p.pred_off.size() + p.delta_off.size()) * 8;
        printf("build %.1f ms  tasks %.1f ms  total %.1f ms | tasks %.1f MB records %.1f MB masks %.1f MB in_* %.1f MB all %.1f MB\n", t1 - t0, t2 - t1, t2 - t0,
               p.tasks.size() * sizeof(TaskHdr) / 1e6, p.records.size() / 1e6, p.masks.size() * 8 / 1e6,
               (p.in_edge.size() * 4 + p.in_dst.size() * 2 + p.in_off.size() * 4) / 1e6, bytes / 1e6);
        {   // the level-program engine's host plan (what dg_dip_create runs for engine 4)
            const double t4 = now();
            Sweep4Shape s4; s4.stride = 680; s4.kn = 26; s4.slot_bytes = 4096; s4.grid = 1;
            Plan4 q4; std::string why;
            const bool ok = plan4_build(p, g, s4, 10, q4, why);
            printf("  plan4_build %.1f ms (%s) program %.1f MB codes %.1f MB\n", now() - t4, ok ? "ok" : why.c_str(), q4.prog_bytes / 1e6, q4.pred_elems * 2 / 1e6);
        }
        const double t3 = now();
        { DipPlan q; std::swap(q, p); }
        printf("  free %.1f ms\n", now() - t3);
    }
}
