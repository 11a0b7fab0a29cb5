// dg_common.cuh — context object and error plumbing shared by the .cu translation units.
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <mutex>
#include <string>
#include <vector>

#include "../../../include/dipgenie_cuda.h"

struct dg_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    std::vector<cudaStream_t> batch_streams;   // dg_dp_diploid_batch: one per concurrently resident sample
    // page-locked staging blocks of the batch planners (kept for the life of the context: pinning is slow)
    struct Pinned { void* p; size_t bytes; };
    std::vector<Pinned> pinned_free;
    std::mutex pinned_mu;
    std::string err;
    dg_sketch_stats_t sketch_stats = {};
};

namespace dg {

inline int fail(dg_ctx* ctx, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (ctx) {
        static std::mutex mu;                  // batch workers may fail concurrently
        std::lock_guard<std::mutex> lk(mu);
        ctx->err = buf;
    }
    return code;
}

#define DG_CUDA(ctx, expr)                                                                          \
    do {                                                                                            \
        cudaError_t e__ = (expr);                                                                   \
        if (e__ != cudaSuccess)                                                                     \
            return dg::fail((ctx), e__ == cudaErrorMemoryAllocation ? DG_ERR_NOMEM : DG_ERR_CUDA,   \
                            "%s:%d %s: %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e__));    \
    } while (0)

// Owning device buffer (freed on destruction).  With a stream the memory comes from the device's
// stream-ordered pool (cudaMallocAsync; dg_create raises the pool's release threshold, so the GiB-sized
// buffers of one problem are recycled by the next instead of going back to the driver: cudaFree of ~1 GB
// cost 25-1100 ms per problem on B200); without one it is a plain cudaMalloc.
template <class T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    cudaStream_t s_ = nullptr;
    bool async_ = false;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { release(); }
    void release() {
        if (!p) return;
        if (async_) cudaFreeAsync(p, s_); else cudaFree(p);
        p = nullptr;
    }
    cudaError_t alloc(size_t count) {
        release();
        n = count; async_ = false;
        return cudaMalloc((void**)&p, (count ? count : 1) * sizeof(T));
    }
    cudaError_t alloc(size_t count, cudaStream_t s) {
        release();
        n = count; s_ = s; async_ = true;
        return cudaMallocAsync((void**)&p, (count ? count : 1) * sizeof(T), s);
    }
    cudaError_t upload(const T* h, size_t count, cudaStream_t s) {
        cudaError_t e = alloc(count, s);
        if (e != cudaSuccess || count == 0) return e;
        return cudaMemcpyAsync(p, h, count * sizeof(T), cudaMemcpyHostToDevice, s);
    }
    size_t bytes() const { return n * sizeof(T); }
};

}  // namespace dg
