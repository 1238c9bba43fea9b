// The one collective of the path: the sum of the per-pixel products (counts [sec][pixel], histograms [pixel][bin]) over
// the GPUs of a box, through NCCL on the context's stream.  In the reference one process adds every roach's photons
// into photon_counts[sec][roach * NPIXELS_PER_ROACH + adr] (PacketMaster.c:371-381); here every rank decodes the
// boards / file chunks it owns into its own copy of the arrays and the copies are summed once.  Integer sums are order
// independent, so the result is bit-identical at any GPU count; the 2500-event cap quirk (PacketMaster.c:373-380) is
// applied afterwards by mkid_counts_cap.
//
// NCCL is bound at run time (dlopen of libnccl.so.2: the copy a host process such as PyTorch has already loaded, else
// the system one), so that single-GPU users of the library need no NCCL at all.  Only NCCL's stable C entry points are
// used; their prototypes are restated here.
#include <dlfcn.h>

#include "common.cuh"

namespace {

typedef struct { char internal[128]; } nccl_unique_id;       // ncclUniqueId
typedef void *nccl_comm;                                        // ncclComm_t
enum { NCCL_UINT32 = 3, NCCL_SUM = 0 };                         // ncclUint32, ncclSum

struct NcclApi {
    void *so = nullptr;
    int (*GetUniqueId)(nccl_unique_id *) = nullptr;
    int (*CommInitRank)(nccl_comm *, int, nccl_unique_id, int) = nullptr;
    int (*CommDestroy)(nccl_comm) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, nccl_comm, cudaStream_t) = nullptr;
    int (*Reduce)(const void *, void *, size_t, int, int, int, nccl_comm, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    int (*GetVersion)(int *) = nullptr;
    std::string err;
};

NcclApi *nccl_api() {
    static NcclApi api;
    if (api.so || !api.err.empty()) return &api;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *n : names) {
        api.so = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (api.so) break;
    }
    if (!api.so) { api.err = std::string("libnccl.so.2 not found: ") + dlerror(); return &api; }
    auto sym = [&](const char *n) -> void * {
        void *p = dlsym(api.so, n);
        if (!p && api.err.empty()) api.err = std::string("NCCL symbol missing: ") + n;
        return p;
    };
    api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
    api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
    api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
    api.AllReduce = (decltype(api.AllReduce))sym("ncclAllReduce");
    api.Reduce = (decltype(api.Reduce))sym("ncclReduce");
    api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
    api.GetVersion = (decltype(api.GetVersion))sym("ncclGetVersion");
    return &api;
}

#define MKID_NCCL(ctx, api, call)                                                                       \
    do {                                                                                                \
        int r__ = (call);                                                                               \
        if (r__ != 0) return mkid_fail((ctx), MKID_ENCCL, "%s failed: %s", #call, (api)->GetErrorString(r__)); \
    } while (0)

}  // namespace

extern "C" int mkid_nccl_version(mkid_ctx *ctx, int32_t *version) {
    NcclApi *a = nccl_api();
    if (!a->err.empty()) return mkid_fail(ctx, MKID_ENCCL, "%s", a->err.c_str());
    if (!version) return mkid_fail(ctx, MKID_EINVAL, "nccl_version: NULL");
    int v = 0;
    MKID_NCCL(ctx, a, a->GetVersion(&v));
    *version = v;
    return MKID_OK;
}

extern "C" int mkid_nccl_unique_id(mkid_ctx *ctx, uint8_t id_out[128]) {
    NcclApi *a = nccl_api();
    if (!a->err.empty()) return mkid_fail(ctx, MKID_ENCCL, "%s", a->err.c_str());
    if (!id_out) return mkid_fail(ctx, MKID_EINVAL, "nccl_unique_id: NULL");
    nccl_unique_id id;
    MKID_NCCL(ctx, a, a->GetUniqueId(&id));
    memcpy(id_out, id.internal, 128);
    return MKID_OK;
}

extern "C" int mkid_nccl_init(mkid_ctx *ctx, const uint8_t id[128], int32_t n_ranks, int32_t rank, void **comm_out) {
    if (!ctx) return MKID_EINVAL;
    MKID_REQUIRE(ctx, id && comm_out && n_ranks >= 1 && rank >= 0 && rank < n_ranks, "nccl_init: bad argument");
    NcclApi *a = nccl_api();
    if (!a->err.empty()) return mkid_fail(ctx, MKID_ENCCL, "%s", a->err.c_str());
    MKID_CUDA(ctx, cudaSetDevice(ctx->device));
    nccl_unique_id uid;
    memcpy(uid.internal, id, 128);
    nccl_comm comm = nullptr;
    MKID_NCCL(ctx, a, a->CommInitRank(&comm, n_ranks, uid, rank));
    *comm_out = comm;
    return MKID_OK;
}

extern "C" int mkid_nccl_destroy(mkid_ctx *ctx, void *comm) {
    if (!comm) return MKID_OK;
    NcclApi *a = nccl_api();
    if (!a->err.empty()) return mkid_fail(ctx, MKID_ENCCL, "%s", a->err.c_str());
    if (ctx) { cudaSetDevice(ctx->device); cudaStreamSynchronize(ctx->stream); }
    MKID_NCCL(ctx, a, a->CommDestroy((nccl_comm)comm));
    return MKID_OK;
}

extern "C" int mkid_hist_allreduce(mkid_ctx *ctx, void *comm, uint32_t *buf, size_t n) {
    if (!ctx) return MKID_EINVAL;
    MKID_REQUIRE(ctx, comm && buf && n > 0, "hist_allreduce: bad argument");
    MKID_REQUIRE(ctx, mkid_is_device_ptr(buf), "hist_allreduce: the products must be in device memory");
    NcclApi *a = nccl_api();
    if (!a->err.empty()) return mkid_fail(ctx, MKID_ENCCL, "%s", a->err.c_str());
    MKID_CUDA(ctx, cudaSetDevice(ctx->device));
    MKID_NCCL(ctx, a, a->AllReduce(buf, buf, n, NCCL_UINT32, NCCL_SUM, (nccl_comm)comm, ctx->stream));
    ctx->launches++;
    return MKID_OK;
}

extern "C" int mkid_hist_reduce(mkid_ctx *ctx, void *comm, uint32_t *buf, size_t n, int32_t root) {
    if (!ctx) return MKID_EINVAL;
    MKID_REQUIRE(ctx, comm && buf && n > 0 && root >= 0, "hist_reduce: bad argument");
    MKID_REQUIRE(ctx, mkid_is_device_ptr(buf), "hist_reduce: the products must be in device memory");
    NcclApi *a = nccl_api();
    if (!a->err.empty()) return mkid_fail(ctx, MKID_ENCCL, "%s", a->err.c_str());
    MKID_CUDA(ctx, cudaSetDevice(ctx->device));
    MKID_NCCL(ctx, a, a->Reduce(buf, buf, n, NCCL_UINT32, NCCL_SUM, root, (nccl_comm)comm, ctx->stream));
    ctx->launches++;
    return MKID_OK;
}

// ---- ordering against streams the library does not own (torch's current stream, a caller's stream)
extern "C" int mkid_wait_stream(mkid_ctx *ctx, void *ext_stream) {
    if (!ctx) return MKID_EINVAL;
    cudaEvent_t ev;
    MKID_CUDA(ctx, cudaSetDevice(ctx->device));
    MKID_CUDA(ctx, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    MKID_CUDA(ctx, cudaEventRecord(ev, (cudaStream_t)ext_stream));
    MKID_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ev, 0));
    MKID_CUDA(ctx, cudaEventDestroy(ev));
    return MKID_OK;
}

extern "C" int mkid_stream_wait_ctx(mkid_ctx *ctx, void *ext_stream) {
    if (!ctx) return MKID_EINVAL;
    cudaEvent_t ev;
    MKID_CUDA(ctx, cudaSetDevice(ctx->device));
    MKID_CUDA(ctx, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    MKID_CUDA(ctx, cudaEventRecord(ev, ctx->stream));
    MKID_CUDA(ctx, cudaStreamWaitEvent((cudaStream_t)ext_stream, ev, 0));
    MKID_CUDA(ctx, cudaEventDestroy(ev));
    return MKID_OK;
}

// the context's stream waits for event `slot` of ANOTHER context (recorded there with mkid_event_record): two contexts on
// one GPU form a two-stage pipeline (channelizer kernel of batch k + 1 under the detection / decode of batch k)
extern "C" int mkid_stream_wait_event(mkid_ctx *ctx, mkid_ctx *owner, int32_t slot) {
    if (!ctx) return MKID_EINVAL;
    MKID_REQUIRE(ctx, owner && slot >= 0 && slot < MKID_NUM_EVENTS, "stream_wait_event: bad argument");
    MKID_REQUIRE(ctx, owner->device == ctx->device, "stream_wait_event: both contexts must be on the same device");
    MKID_CUDA(ctx, cudaSetDevice(ctx->device));
    MKID_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, owner->events[slot], 0));
    return MKID_OK;
}
