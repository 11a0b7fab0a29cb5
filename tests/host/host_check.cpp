// tests/host/host_check.cpp — TEST INFRASTRUCTURE ONLY.
// Runs the product's host glue (libdipgenie_host.so: GFA/read parsing, panel model, anchor filter + classifier,
// graph expansion, levelization, stitching, FASTA) end to end on a CPU-only box by binding the pipeline's
// device-stage table to the oracle (oracle/liboracle.so) instead of libdipgenie_cuda.so, so that the FASTA it
// writes can be compared byte for byte with the reference's.  The product CLI (csrc/host/main.cpp) never links this.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../dipgenie_b200/csrc/host/pipeline.h"

extern "C" {
int dgo_sketch_reads(const uint8_t*, const uint64_t*, uint32_t, int, int, uint64_t**, uint32_t**, uint64_t*, uint64_t**, uint64_t**);
int dgo_index_walks(const uint8_t*, const uint64_t*, uint32_t, const int32_t*, const uint64_t*, uint32_t, const int32_t*, int, int,
                    const uint64_t*, uint64_t, uint64_t*, uint64_t**, uint32_t**, uint64_t**, int32_t**, uint64_t**, uint64_t**);
int dgo_dp_haploid(int32_t, const int64_t*, const int32_t*, const uint8_t*, const int64_t*, const int32_t*, int32_t, int32_t, int32_t*,
                   int64_t*, int32_t*, int64_t);
int dgo_dp_diploid(int32_t, const int32_t*, const int64_t*, const int32_t*, const uint8_t*, const int64_t*, const int32_t*, const uint8_t*,
                   int32_t, int32_t, int32_t*, int32_t*, int32_t*, int32_t*, int32_t*, int32_t*, uint64_t*, uint64_t*);
}

int main(int argc, char** argv) {
    dgh::Options o;
    o.verbose = getenv("HOST_CHECK_VERBOSE") != nullptr;      // the stage timeline on stderr (diagnostics)
    for (int i = 1; i + 1 < argc; i += 2) {
        const char f = argv[i][1];
        const char* v = argv[i + 1];
        switch (f) {
            case 'g': o.gfa = v; break; case 'r': o.reads = v; break; case 'o': o.out = v; break;
            case 't': o.threads = atoi(v); break; case 'p': o.ploidy = atoi(v); break; case 'R': o.R = atoi(v); break;
            case 'k': o.k = atoi(v); break; case 'w': o.w = atoi(v); break; case 'T': o.threshold = (float)atof(v); break;
            default: break;
        }
    }
    dgh::Backend be;
    be.sketch_reads = [](void*, const uint8_t* b, const uint64_t* off, uint32_t n, int k, int w, uint64_t** sp, uint32_t** rc, uint64_t* ns) {
        return dgo_sketch_reads(b, off, n, k, w, sp, rc, ns, nullptr, nullptr); };
    be.index_walks = [](void*, const uint8_t* sb, const uint64_t* so, uint32_t ns, const int32_t* wv, const uint64_t* wo, uint32_t nw,
                        const int32_t* tom, int k, int w, const uint64_t* sp, uint64_t nsp, uint64_t* nm, uint64_t** ho, uint32_t** hs,
                        uint64_t** vo, int32_t** hv) {
        return dgo_index_walks(sb, so, ns, wv, wo, nw, tom, k, w, sp, nsp, nm, ho, hs, vo, hv, nullptr, nullptr); };
    be.dp_haploid = [](void*, int32_t n, const int64_t* ao, const int32_t* ad, const uint8_t* aw, const int64_t* co, const int32_t* cv,
                       int32_t nc, int32_t R, int32_t* cb, int64_t* po, int32_t** paths) {
        const int64_t cap = (int64_t)n * (R + 1);
        *paths = (int32_t*)malloc((size_t)(cap > 0 ? cap : 1) * 4);
        return dgo_dp_haploid(n, ao, ad, aw, co, cv, nc, R, cb, po, *paths, cap); };
    be.dp_diploid = [](void*, int32_t L, const int32_t* lo, const int64_t* ao, const int32_t* ad, const uint8_t* aw, const int64_t* co,
                       const int32_t* cv, const uint8_t* hom, int32_t nc, int32_t R, int32_t* val, int32_t* sh, int32_t* p1, int32_t* n1,
                       int32_t* p2, int32_t* n2) {
        return dgo_dp_diploid(L, lo, ao, ad, aw, co, cv, hom, nc, R, val, sh, p1, n1, p2, n2, nullptr, nullptr); };
    be.free_array = [](void* p) { free(p); };
    be.last_error = [](void*) { return "oracle stage failed"; };
    dgh::RunSummary s;
    std::string err;
    const int rc = dgh::run_pipeline(o, be, s, err);
    if (rc) { fprintf(stderr, "host_check: %s\n", err.c_str()); return rc; }
    printf("{\"spectrum\": %lld, \"n_hom\": %lld, \"n_het\": %lld, \"dp_value\": %d, \"r1\": %d, \"r2\": %d, \"best_r\": %d, \"len1\": %lld, \"len2\": %lld}\n",
           (long long)s.spectrum, (long long)s.n_hom, (long long)s.n_het, s.dp_value, s.r1, s.r2, s.best_r, (long long)s.len1, (long long)s.len2);
    return 0;
}
