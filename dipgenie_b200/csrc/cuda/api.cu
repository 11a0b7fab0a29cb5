// api.cu — context lifecycle of libdipgenie_cuda.so (include/dipgenie_cuda.h).
#include <cstdlib>

#include "dg_common.cuh"

extern "C" {

dg_ctx* dg_create(int device) {
    // Batch slots are long-running persistent kernels on separate streams: with the default of 8 hardware
    // work queues, streams 9+ share a queue with a resident sweep and wait for it (measured on B200: 12 slots
    // ran as two waves).  Only effective if set before the process creates its CUDA context; hosts that
    // initialise CUDA earlier (PyTorch) must export it themselves.
    setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0);
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || device < 0 || device >= n) return nullptr;
    if (cudaSetDevice(device) != cudaSuccess) return nullptr;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return nullptr;
    dg_ctx* ctx = new dg_ctx();
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete ctx;
        return nullptr;
    }
    // keep freed blocks in the stream-ordered pool (see DevBuf)
    cudaMemPool_t pool = nullptr;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess && pool) {
        unsigned long long keep = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    return ctx;
}

void dg_destroy(dg_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    for (cudaStream_t s : ctx->batch_streams) cudaStreamDestroy(s);
    for (auto& b : ctx->pinned_free) cudaFreeHost(b.p);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char* dg_last_error(dg_ctx* ctx) { return ctx ? ctx->err.c_str() : "no context (dg_create failed: no usable CUDA device)"; }

void dg_free(void* p) { free(p); }

int dg_release_cached_memory(dg_ctx* ctx) {
    if (!ctx) return DG_ERR_ARG;
    DG_CUDA(ctx, cudaSetDevice(ctx->device));
    DG_CUDA(ctx, cudaDeviceSynchronize());
    cudaMemPool_t pool = nullptr;
    DG_CUDA(ctx, cudaDeviceGetDefaultMemPool(&pool, ctx->device));
    DG_CUDA(ctx, cudaMemPoolTrimTo(pool, 0));
    return DG_OK;
}

int dg_device_info(dg_ctx* ctx, int* sm_count, size_t* free_bytes, size_t* total_bytes) {
    if (!ctx) return DG_ERR_ARG;
    DG_CUDA(ctx, cudaSetDevice(ctx->device));
    size_t f = 0, t = 0;
    DG_CUDA(ctx, cudaMemGetInfo(&f, &t));
    if (sm_count) *sm_count = ctx->sm_count;
    if (free_bytes) *free_bytes = f;
    if (total_bytes) *total_bytes = t;
    return DG_OK;
}

}  // extern "C"
