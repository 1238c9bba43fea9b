// Shared context + helpers for the mkidgpu C ABI (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <string>
#include <vector>

#include "../../include/mkidgpu.h"

#define MKID_NUM_EVENTS 64

struct mkid_ctx {
    int device = 0;
    int num_sms = 148;
    cudaStream_t stream = nullptr;
    cudaEvent_t events[MKID_NUM_EVENTS] = {};
    std::string err;
    int64_t launches = 0;
    // grow-on-demand scratch (device) and staging buffers
    void *scratch[16] = {};
    size_t scratch_bytes[16] = {};
    void *l2_flush = nullptr;
    // last segment table uploaded by the decode path (skips re-upload when unchanged)
    std::vector<char> dec_meta_host;
    void *dec_meta_dev = nullptr;
    std::vector<char> dec_ranges_host, dec_key;
    // device buffers that hold the cached tables: private to the decode path (scratch slots are shared)
    void *dec_priv[2] = {};
    cudaEvent_t dbg_events[8] = {};     // MKID_DEC_TIMING switch
    // double-buffered host -> device uploads that overlap the compute stream (mkid_upload_async)
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t up_done[4] = {}, up_free[4] = {};
    bool up_free_valid[4] = {};
    size_t dec_priv_bytes[2] = {};
    void *dec_ranges_dev = nullptr;
    size_t l2_flush_bytes = 0;
    // segment table + per-key counts of mkid_merge_words_dev (private: the scratch slots are shared between entry points)
    void *merge_dev = nullptr;
    size_t merge_bytes = 0;
    std::vector<char> merge_host;
};

enum { SCR_IN = 0, SCR_IN1, SCR_IN2, SCR_IN3, SCR_OUT0, SCR_OUT1, SCR_OUT2, SCR_OUT3, SCR_STATE, SCR_META, SCR_AUX0, SCR_AUX1, SCR_AUX2, SCR_AUX3, SCR_AUX4, SCR_AUX5 };

int mkid_fail(mkid_ctx *ctx, int code, const char *fmt, ...);

#define MKID_CUDA(ctx, call)                                                          \
    do {                                                                              \
        cudaError_t e__ = (call);                                                     \
        if (e__ != cudaSuccess)                                                       \
            return mkid_fail((ctx), MKID_ECUDA, "%s failed: %s (%s:%d)", #call,       \
                             cudaGetErrorString(e__), __FILE__, __LINE__);            \
    } while (0)

#define MKID_CHECK_LAUNCH(ctx)                                                        \
    do {                                                                              \
        (ctx)->launches++;                                                            \
        cudaError_t e__ = cudaGetLastError();                                         \
        if (e__ != cudaSuccess)                                                       \
            return mkid_fail((ctx), MKID_ECUDA, "kernel launch failed: %s (%s:%d)",   \
                             cudaGetErrorString(e__), __FILE__, __LINE__);            \
    } while (0)

#define MKID_REQUIRE(ctx, cond, msg)                                                  \
    do {                                                                              \
        if (!(cond)) return mkid_fail((ctx), MKID_EINVAL, "%s (%s:%d)", msg, __FILE__, __LINE__); \
    } while (0)

// scratch buffer of at least `bytes` (contents undefined after growth)
int mkid_scratch(mkid_ctx *ctx, int slot, size_t bytes, void **out);
bool mkid_is_device_ptr(const void *p);

// Returns a device pointer for `p`: `p` itself when it is device memory, else a copy in
// scratch slot `slot` (async H2D on the ctx stream).
int mkid_stage_in(mkid_ctx *ctx, const void *p, size_t bytes, int slot, const void **dev);
// For outputs: device pointer to write to (p itself or scratch); `accumulate` => host
// contents are copied in first.  mkid_stage_out_finish copies back when p is host memory.
int mkid_stage_out(mkid_ctx *ctx, void *p, size_t bytes, int slot, bool accumulate, void **dev);
int mkid_stage_out_finish(mkid_ctx *ctx, void *p, size_t bytes, void *dev);

// ---- mbarrier + 1-D TMA bulk copy (cp.async.bulk, SASS UBLKCP) helpers
__device__ __forceinline__ uint32_t mk_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mk_mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mk_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mk_mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mk_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mk_mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(mk_smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void mk_bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     mk_smem_u32(dst)), "l"(src), "r"(bytes), "r"(mk_smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ uint4 ld_stream_u4(const uint4 *p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
