// pipeline.h — the DipGenie run as the reference's main() drives it (src/main.cpp:24-209), with the hot-path
// stages behind a table of C-ABI entry points (include/dipgenie_cuda.h).  The CLI binds the table to
// libdipgenie_cuda.so; the CPU test harness (tests/host/host_check.cpp, test infrastructure) binds the same
// table to the oracle so that the host glue can be checked against the reference without a GPU.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../../include/dipgenie_cuda.h"

namespace dgh {

struct Options {                       // defaults of src/main.cpp:44-59 and src/options.cpp:7-16
    std::string gfa, reads, out;
    int threads = 4, ploidy = 2, R = 18, k = 31, w = 25;
    float threshold = 1.0f;
    bool verbose = true;
};

struct Backend {                       // signatures of include/dipgenie_cuda.h (ctx is opaque here)
    void* ctx = nullptr;
    int (*sketch_reads)(void*, const uint8_t*, const uint64_t*, uint32_t, int, int, uint64_t**, uint32_t**, uint64_t*) = nullptr;
    int (*index_walks)(void*, const uint8_t*, const uint64_t*, uint32_t, const int32_t*, const uint64_t*, uint32_t,
                       const int32_t*, int, int, const uint64_t*, uint64_t, uint64_t*, uint64_t**, uint32_t**, uint64_t**,
                       int32_t**) = nullptr;
    int (*dp_haploid)(void*, int32_t, const int64_t*, const int32_t*, const uint8_t*, const int64_t*, const int32_t*, int32_t,
                      int32_t, int32_t*, int64_t*, int32_t**) = nullptr;
    int (*dp_diploid)(void*, int32_t, const int32_t*, const int64_t*, const int32_t*, const uint8_t*, const int64_t*,
                      const int32_t*, const uint8_t*, int32_t, int32_t, int32_t*, int32_t*, int32_t*, int32_t*, int32_t*,
                      int32_t*) = nullptr;
    int (*dp_diploid_batch)(void*, int32_t, const dg_dip_input_t*, dg_dip_output_t*, int32_t, int32_t) = nullptr;   // optional
    void (*free_array)(void*) = nullptr;
    const char* (*last_error)(void*) = nullptr;
};

struct RunSummary {                    // integer checkpoints of the reference's log
    int64_t spectrum = 0, n_hom = 0, n_het = 0;
    int32_t dp_value = 0, r1 = 0, r2 = 0, best_r = 0;
    int64_t len1 = 0, len2 = 0;
};

// Returns 0, or the reference's exit status for the failure (usage/IO errors 1).  Progress lines go to stderr.
int run_pipeline(const Options& o, const Backend& be, RunSummary& sum, std::string& err);

// Several runs in one process; diploid DPs are batched on the GPU when the backend offers dg_dp_diploid_batch.
// Returns the number of failed jobs (errs[i] non-empty for those).
int run_batch(const std::vector<Options>& jobs, const Backend& be, std::vector<RunSummary>& sums, std::vector<std::string>& errs);

}  // namespace dgh
