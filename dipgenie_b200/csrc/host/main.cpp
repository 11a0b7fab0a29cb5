// main.cpp — the drop-in front end: same flags and defaults as the reference's main()
// (src/main.cpp:39-111: -g -r -o required; -t threads, -p ploidy {1,2}, -R recombination limit, -k, -w, -T
// threshold; usage + exit 1 when an input is missing, message + exit 0 for an unknown ploidy, :159-162),
// host glue from libdipgenie_host.so, hot path on the GPU through the C ABI of libdipgenie_cuda.so.
// There is no CPU fallback: without a usable CUDA device the program exits with an error.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../../include/dipgenie_cuda.h"
#include "pipeline.h"

static void usage(FILE* f) {
    fprintf(f, "Usage: dipgenie [options]\n");
    fprintf(f, "  -g FILE   pangenome graph, GFA v1.1 with W lines (optionally gzip)   [required]\n");
    fprintf(f, "  -r FILE   reads, FASTA/FASTQ (optionally gzip)                       [required]\n");
    fprintf(f, "  -o FILE   output FASTA                                               [required]\n");
    fprintf(f, "  -t INT    host threads [4]\n");
    fprintf(f, "  -p INT    ploidy, 1 or 2 [2]\n");
    fprintf(f, "  -R INT    recombination limit [18]\n");
    fprintf(f, "  -k INT    k-mer size [31]      -w INT   minimizer window [25]\n");
    fprintf(f, "  -T FLOAT  shared-anchor threshold [1.0]\n");
    fprintf(f, "  --device INT   CUDA device [0]      --quiet   no progress lines\n");
    fprintf(f, "  (the reference's other flags are accepted with its own arity and ignored: -x -d -c -l -s -m -P -a -q -H -N take a\n");
    fprintf(f, "   value, -D -S do not; they steer its ILP branch and debug output only)\n");
    fprintf(f, "  -B FILE   batch: one job per line, \"graph<TAB>reads<TAB>out\" (then -g/-r/-o are not needed); the\n");
    fprintf(f, "            diploid DPs of all jobs run side by side on the GPU\n");
}

int main(int argc, char** argv) {
    dgh::Options o;
    int device = 0;
    std::string batch_file;
    for (int i = 1; i < argc; ++i) {
        const char* a = argv[i];
        if (!strcmp(a, "--version")) { puts("dipgenie-b200 1.0"); return 0; }
        if (!strcmp(a, "-h")) { usage(stdout); return 0; }
        if (!strcmp(a, "--quiet")) { o.verbose = false; continue; }
        if (!strcmp(a, "--device")) { if (i + 1 >= argc) { usage(stderr); return 1; } device = atoi(argv[++i]); continue; }
        if (a[0] != '-' || !a[1]) continue;
        const char f = a[1];
        if (f == 'D' || f == 'S') continue;                                         // argument-less in the reference's opt_str (src/main.cpp:39)
        const char* v = a[2] ? a + 2 : (i + 1 < argc ? argv[++i] : nullptr);       // "-R18" and "-R 18" both parse, like ketopt
        if (!v) { usage(stderr); return 1; }
        switch (f) {
            case 'g': o.gfa = v; break;
            case 'r': o.reads = v; break;
            case 'o': o.out = v; break;
            case 't': o.threads = atoi(v); break;
            case 'p': o.ploidy = atoi(v); break;
            case 'R': o.R = atoi(v); break;
            case 'k': o.k = atoi(v); break;
            case 'w': o.w = atoi(v); break;
            case 'T': o.threshold = (float)atof(v); break;
            case 'B': batch_file = v; break;
            default: break;                                                         // the reference's -x -d -c -l -s -m -P -a -q -H -N: value consumed, ignored
        }
    }
    if (batch_file.empty() && (o.gfa.empty() || o.reads.empty() || o.out.empty())) { usage(stderr); return 1; }
    if (o.ploidy != 1 && o.ploidy != 2) { fprintf(stderr, "Ploidy must be 1 or 2\n"); return 0; }
    auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t0 = now();
    dg_ctx* ctx = dg_create(device);
    if (o.verbose) fprintf(stderr, "[M::main] CUDA context ready after %.3f sec\n", now() - t0);
    if (!ctx) { fprintf(stderr, "dipgenie: %s\n", dg_last_error(nullptr)); return 1; }
    dgh::Backend be;
    be.ctx = ctx;
    be.sketch_reads = [](void* c, const uint8_t* b, const uint64_t* off, uint32_t n, int k, int w, uint64_t** sp, uint32_t** rc, uint64_t* ns) {
        return dg_sketch_reads((dg_ctx*)c, b, off, n, k, w, sp, rc, ns); };
    be.index_walks = [](void* c, const uint8_t* sb, const uint64_t* so, uint32_t ns, const int32_t* wv, const uint64_t* wo, uint32_t nw,
                        const int32_t* tom, int k, int w, const uint64_t* sp, uint64_t nsp, uint64_t* nm, uint64_t** ho, uint32_t** hs,
                        uint64_t** vo, int32_t** hv) {
        return dg_index_walks((dg_ctx*)c, sb, so, ns, wv, wo, nw, tom, k, w, sp, nsp, nm, ho, hs, vo, hv); };
    be.dp_haploid = [](void* c, int32_t n, const int64_t* ao, const int32_t* ad, const uint8_t* aw, const int64_t* co, const int32_t* cv,
                       int32_t nc, int32_t R, int32_t* cb, int64_t* po, int32_t** paths) {
        return dg_dp_haploid((dg_ctx*)c, n, ao, ad, aw, co, cv, nc, R, cb, po, paths); };
    be.dp_diploid = [](void* c, int32_t L, const int32_t* lo, const int64_t* ao, const int32_t* ad, const uint8_t* aw, const int64_t* co,
                       const int32_t* cv, const uint8_t* hom, int32_t nc, int32_t R, int32_t* val, int32_t* sh, int32_t* p1, int32_t* n1,
                       int32_t* p2, int32_t* n2) {
        return dg_dp_diploid((dg_ctx*)c, L, lo, ao, ad, aw, co, cv, hom, nc, R, val, sh, p1, n1, p2, n2); };
    be.dp_diploid_batch = [](void* c, int32_t n, const dg_dip_input_t* in, dg_dip_output_t* out, int32_t mc, int32_t cps) {
        return dg_dp_diploid_batch((dg_ctx*)c, n, in, out, mc, cps); };
    be.free_array = [](void* p) { dg_free(p); };
    be.last_error = [](void* c) { return dg_last_error((dg_ctx*)c); };
    int rc = 0;
    if (!batch_file.empty()) {
        std::vector<dgh::Options> jobs;
        FILE* f = fopen(batch_file.c_str(), "r");
        if (!f) { fprintf(stderr, "dipgenie: cannot open %s\n", batch_file.c_str()); dg_destroy(ctx); return 1; }
        char line[8192];
        while (fgets(line, sizeof line, f)) {
            char g[4096], r[4096], out[4096];
            if (sscanf(line, "%4095s %4095s %4095s", g, r, out) != 3) continue;
            dgh::Options j = o;
            j.gfa = g; j.reads = r; j.out = out; j.verbose = false;
            jobs.push_back(j);
        }
        fclose(f);
        std::vector<dgh::RunSummary> sums;
        std::vector<std::string> errs;
        const double tb = now();
        rc = dgh::run_batch(jobs, be, sums, errs) ? 1 : 0;
        for (size_t i = 0; i < jobs.size(); ++i) {
            if (!errs[i].empty()) fprintf(stderr, "dipgenie: job %zu (%s): %s\n", i, jobs[i].out.c_str(), errs[i].c_str());
            else if (o.verbose) fprintf(stderr, "job %zu\t%s\tDP value %d\tr %d/%d\tbp %lld/%lld\n", i, jobs[i].out.c_str(), sums[i].dp_value, sums[i].r1, sums[i].r2, (long long)sums[i].len1, (long long)sums[i].len2);
        }
        if (o.verbose) fprintf(stderr, "[M::main] %zu jobs in %.3f sec (%.2f samples/s)\n", jobs.size(), now() - tb, jobs.size() / (now() - tb));
    } else {
        dgh::RunSummary sum;
        std::string err;
        rc = dgh::run_pipeline(o, be, sum, err);
        if (rc) fprintf(stderr, "dipgenie: %s\n", err.c_str());
    }
    const double t1 = now();
    dg_destroy(ctx);
    if (o.verbose) fprintf(stderr, "[M::main] released the device after %.3f sec; total %.3f sec\n", now() - t1, now() - t0);
    return rc;
}
