// Context management, memory staging and small utilities of the mkidgpu C ABI.
#include <stdarg.h>

#include "common.cuh"

static std::string g_init_error;

int mkid_fail(mkid_ctx *ctx, int code, const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf; else g_init_error = buf;
    return code;
}

extern "C" const char *mkid_version(void) { return "mkidgpu 0.1 (sm_100a)"; }

extern "C" int mkid_init(int device, mkid_ctx **out) {
    if (!out) return mkid_fail(nullptr, MKID_EINVAL, "mkid_init: out is NULL");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return mkid_fail(nullptr, MKID_ENODEV, "no CUDA device (%s); there is no CPU fallback",
                         e != cudaSuccess ? cudaGetErrorString(e) : "count=0");
    if (device < 0 || device >= n) return mkid_fail(nullptr, MKID_EINVAL, "device %d out of range (%d)", device, n);
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess)
        return mkid_fail(nullptr, MKID_ECUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
    if (prop.major != 10)
        return mkid_fail(nullptr, MKID_ENODEV, "device %d is sm_%d%d; this library is built for sm_100a only",
                         device, prop.major, prop.minor);
    if ((e = cudaSetDevice(device)) != cudaSuccess)
        return mkid_fail(nullptr, MKID_ECUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
    mkid_ctx *ctx = new mkid_ctx();
    ctx->device = device;
    ctx->num_sms = prop.multiProcessorCount;
    if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess) {
        delete ctx;
        return mkid_fail(nullptr, MKID_ECUDA, "cudaStreamCreate: %s", cudaGetErrorString(e));
    }
    for (int i = 0; i < MKID_NUM_EVENTS; ++i) cudaEventCreate(&ctx->events[i]);
    *out = ctx;
    return MKID_OK;
}

extern "C" void mkid_destroy(mkid_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (int i = 0; i < 16; ++i) if (ctx->scratch[i]) cudaFree(ctx->scratch[i]);
    if (ctx->l2_flush) cudaFree(ctx->l2_flush);
    if (ctx->merge_dev) cudaFree(ctx->merge_dev);
    for (int i = 0; i < 2; ++i) if (ctx->dec_priv[i]) cudaFree(ctx->dec_priv[i]);
    for (int i = 0; i < MKID_NUM_EVENTS; ++i) if (ctx->events[i]) cudaEventDestroy(ctx->events[i]);
    for (int i = 0; i < 8; ++i) if (ctx->dbg_events[i]) cudaEventDestroy(ctx->dbg_events[i]);
    for (int i = 0; i < 4; ++i) { if (ctx->up_done[i]) cudaEventDestroy(ctx->up_done[i]); if (ctx->up_free[i]) cudaEventDestroy(ctx->up_free[i]); }
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
}

extern "C" const char *mkid_last_error(mkid_ctx *ctx) { return ctx ? ctx->err.c_str() : g_init_error.c_str(); }

extern "C" int mkid_sync(mkid_ctx *ctx) {
    MKID_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return MKID_OK;
}

extern "C" void *mkid_stream(mkid_ctx *ctx) { return (void *)ctx->stream; }
extern "C" int64_t mkid_launch_count(mkid_ctx *ctx) { return ctx->launches; }

extern "C" int mkid_event_record(mkid_ctx *ctx, int slot) {
    MKID_REQUIRE(ctx, slot >= 0 && slot < MKID_NUM_EVENTS, "event slot out of range");
    MKID_CUDA(ctx, cudaEventRecord(ctx->events[slot], ctx->stream));
    return MKID_OK;
}

extern "C" int mkid_event_elapsed_ms(mkid_ctx *ctx, int a, int b, float *ms) {
    MKID_REQUIRE(ctx, a >= 0 && a < MKID_NUM_EVENTS && b >= 0 && b < MKID_NUM_EVENTS && ms, "bad event args");
    MKID_CUDA(ctx, cudaEventSynchronize(ctx->events[b]));
    MKID_CUDA(ctx, cudaEventElapsedTime(ms, ctx->events[a], ctx->events[b]));
    return MKID_OK;
}

extern "C" int mkid_host_alloc(mkid_ctx *ctx, size_t bytes, void **out) {
    MKID_CUDA(ctx, cudaSetDevice(ctx->device));
    MKID_CUDA(ctx, cudaHostAlloc(out, bytes, cudaHostAllocDefault));
    return MKID_OK;
}
extern "C" int mkid_host_free(mkid_ctx *ctx, void *p) { MKID_CUDA(ctx, cudaFreeHost(p)); return MKID_OK; }
extern "C" int mkid_dev_alloc(mkid_ctx *ctx, size_t bytes, void **out) {
    MKID_CUDA(ctx, cudaSetDevice(ctx->device));
    MKID_CUDA(ctx, cudaMalloc(out, bytes));
    return MKID_OK;
}
extern "C" int mkid_dev_free(mkid_ctx *ctx, void *p) { MKID_CUDA(ctx, cudaFree(p)); return MKID_OK; }
extern "C" int mkid_memcpy(mkid_ctx *ctx, void *dst, const void *src, size_t bytes) {
    MKID_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, ctx->stream));
    return MKID_OK;
}
// Double-buffered uploads on a second stream: the copy of batch k+1 runs while batch k is processed.
extern "C" int mkid_upload_async(mkid_ctx *ctx, void *dst_dev, const void *src_host, size_t bytes, int slot) {
    if (!ctx) return MKID_EINVAL;
    MKID_REQUIRE(ctx, dst_dev && src_host && slot >= 0 && slot < 4 && mkid_is_device_ptr(dst_dev), "upload_async: bad argument");
    MKID_CUDA(ctx, cudaSetDevice(ctx->device));
    if (!ctx->copy_stream) MKID_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    if (!ctx->up_done[slot]) {
        MKID_CUDA(ctx, cudaEventCreateWithFlags(&ctx->up_done[slot], cudaEventDisableTiming));
        MKID_CUDA(ctx, cudaEventCreateWithFlags(&ctx->up_free[slot], cudaEventDisableTiming));
    }
    // the previous consumer of this slot (mkid_upload_consumed) must be done before the buffer is overwritten
    if (ctx->up_free_valid[slot]) MKID_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->up_free[slot], 0));
    MKID_CUDA(ctx, cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, ctx->copy_stream));
    MKID_CUDA(ctx, cudaEventRecord(ctx->up_done[slot], ctx->copy_stream));
    return MKID_OK;
}
extern "C" int mkid_upload_wait(mkid_ctx *ctx, int slot) {
    if (!ctx) return MKID_EINVAL;
    MKID_REQUIRE(ctx, slot >= 0 && slot < 4 && ctx->up_done[slot], "upload_wait: no upload in this slot");
    MKID_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->up_done[slot], 0));      // asynchronous: orders the context stream
    return MKID_OK;
}
extern "C" int mkid_upload_consumed(mkid_ctx *ctx, int slot) {
    if (!ctx) return MKID_EINVAL;
    MKID_REQUIRE(ctx, slot >= 0 && slot < 4 && ctx->up_free[slot], "upload_consumed: no upload in this slot");
    MKID_CUDA(ctx, cudaEventRecord(ctx->up_free[slot], ctx->stream));
    ctx->up_free_valid[slot] = true;
    return MKID_OK;
}
extern "C" int mkid_memset(mkid_ctx *ctx, void *dst, int value, size_t bytes) {
    if (mkid_is_device_ptr(dst)) {
        MKID_CUDA(ctx, cudaMemsetAsync(dst, value, bytes, ctx->stream));
    } else {
        MKID_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        memset(dst, value, bytes);
    }
    return MKID_OK;
}

extern "C" int mkid_flush_l2(mkid_ctx *ctx) {
    const size_t bytes = 256u << 20;   // > 126 MB L2
    if (!ctx->l2_flush) {
        MKID_CUDA(ctx, cudaMalloc(&ctx->l2_flush, bytes));
        ctx->l2_flush_bytes = bytes;
    }
    MKID_CUDA(ctx, cudaMemsetAsync(ctx->l2_flush, 0x5a, bytes, ctx->stream));
    return MKID_OK;
}

int mkid_scratch(mkid_ctx *ctx, int slot, size_t bytes, void **out) {
    if (bytes == 0) bytes = 16;
    if (ctx->scratch_bytes[slot] < bytes) {
        if (ctx->scratch[slot]) {
            MKID_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            MKID_CUDA(ctx, cudaFree(ctx->scratch[slot]));
            ctx->scratch[slot] = nullptr;
            ctx->scratch_bytes[slot] = 0;
        }
        size_t cap = bytes + bytes / 4 + 256;
        cudaError_t e = cudaMalloc(&ctx->scratch[slot], cap);
        if (e != cudaSuccess) return mkid_fail(ctx, MKID_ENOMEM, "scratch alloc of %zu bytes failed: %s", cap, cudaGetErrorString(e));
        ctx->scratch_bytes[slot] = cap;
    }
    *out = ctx->scratch[slot];
    return MKID_OK;
}

bool mkid_is_device_ptr(const void *p) {
    cudaPointerAttributes a;
    cudaError_t e = cudaPointerGetAttributes(&a, p);
    if (e != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

int mkid_stage_in(mkid_ctx *ctx, const void *p, size_t bytes, int slot, const void **dev) {
    if (!p) { *dev = nullptr; return MKID_OK; }
    if (mkid_is_device_ptr(p)) { *dev = p; return MKID_OK; }
    void *d = nullptr;
    int rc = mkid_scratch(ctx, slot, bytes, &d);
    if (rc) return rc;
    MKID_CUDA(ctx, cudaMemcpyAsync(d, p, bytes, cudaMemcpyHostToDevice, ctx->stream));
    *dev = d;
    return MKID_OK;
}

int mkid_stage_out(mkid_ctx *ctx, void *p, size_t bytes, int slot, bool accumulate, void **dev) {
    if (!p) { *dev = nullptr; return MKID_OK; }
    if (mkid_is_device_ptr(p)) { *dev = p; return MKID_OK; }
    void *d = nullptr;
    int rc = mkid_scratch(ctx, slot, bytes, &d);
    if (rc) return rc;
    if (accumulate) MKID_CUDA(ctx, cudaMemcpyAsync(d, p, bytes, cudaMemcpyHostToDevice, ctx->stream));
    *dev = d;
    return MKID_OK;
}

int mkid_stage_out_finish(mkid_ctx *ctx, void *p, size_t bytes, void *dev) {
    if (!p || p == dev) return MKID_OK;
    MKID_CUDA(ctx, cudaMemcpyAsync(p, dev, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    MKID_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return MKID_OK;
}
