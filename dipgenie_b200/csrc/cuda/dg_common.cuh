// dg_common.cuh — context object and error plumbing shared by the .cu translation units.
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <string>

#include "../../../include/dipgenie_cuda.h"

struct dg_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
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
    if (ctx) ctx->err = buf;
    return code;
}

#define DG_CUDA(ctx, expr)                                                                          \
    do {                                                                                            \
        cudaError_t e__ = (expr);                                                                   \
        if (e__ != cudaSuccess)                                                                     \
            return dg::fail((ctx), e__ == cudaErrorMemoryAllocation ? DG_ERR_NOMEM : DG_ERR_CUDA,   \
                            "%s:%d %s: %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e__));    \
    } while (0)

// Owning device buffer (freed on destruction).
template <class T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t count) {
        if (p) { cudaFree(p); p = nullptr; }
        n = count;
        return cudaMalloc((void**)&p, (count ? count : 1) * sizeof(T));
    }
    cudaError_t upload(const T* h, size_t count, cudaStream_t s) {
        cudaError_t e = alloc(count);
        if (e != cudaSuccess || count == 0) return e;
        return cudaMemcpyAsync(p, h, count * sizeof(T), cudaMemcpyHostToDevice, s);
    }
    size_t bytes() const { return n * sizeof(T); }
};

}  // namespace dg
