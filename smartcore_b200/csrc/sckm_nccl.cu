// sckm_nccl.cu -- NCCL binding by dlopen (no link-time dependency, so the library loads on a
// box without NCCL and shares the copy torch already mapped when there is one).
//
// The only data-path collectives of the k-means hot path:
//   * per Lloyd iteration: ONE all-reduce (sum, f64) of [k*d sums | k counts | inertia];
//   * per kmeans++ pass: an all-gather of one f64 per rank (D^2 totals) and an all-reduce (sum, u64)
//     of the chosen seed row published by its owner (zeros elsewhere: a broadcast whose root is
//     only known on the device).
#include "sckm_common.cuh"
#include <dlfcn.h>
#include <cstdlib>

namespace sckm {

// minimal NCCL ABI (stable across 2.x)
typedef struct { char internal[128]; } ncclUniqueId_t;
typedef int ncclResult_t_;
enum { ncclSum_ = 0 };
enum { ncclUint8_ = 1, ncclUint64_ = 5, ncclFloat64_ = 8 };

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t_ (*GetUniqueId)(ncclUniqueId_t*) = nullptr;
    ncclResult_t_ (*CommInitRank)(void**, int, ncclUniqueId_t, int) = nullptr;
    ncclResult_t_ (*CommDestroy)(void*) = nullptr;
    ncclResult_t_ (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    ncclResult_t_ (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t_) = nullptr;
    std::string load_error;
};

static NcclApi* api() {
    static NcclApi a;
    static bool tried = false;
    if (tried) return &a;
    tried = true;
    const char* env = getenv("SCKM_NCCL_LIB");
    const char* names[] = {env, "libnccl.so.2", "libnccl.so"};
    // prefer a copy that is already mapped (torch's bundled NCCL) so that one process never runs two
    for (const char* nm : names) {
        if (!nm) continue;
        a.handle = dlopen(nm, RTLD_NOW | RTLD_NOLOAD);
        if (a.handle) break;
    }
    for (const char* nm : names) {
        if (a.handle) break;
        if (!nm) continue;
        a.handle = dlopen(nm, RTLD_NOW | RTLD_LOCAL);
    }
    if (!a.handle) { a.load_error = std::string("cannot dlopen libnccl.so.2: ") + (dlerror() ? dlerror() : "?"); return &a; }
#define SYM(field, name)                                                       \
    a.field = reinterpret_cast<decltype(a.field)>(dlsym(a.handle, name));      \
    if (!a.field) { a.load_error = std::string("missing NCCL symbol ") + name; a.handle = nullptr; return &a; }
    SYM(GetUniqueId, "ncclGetUniqueId")
    SYM(CommInitRank, "ncclCommInitRank")
    SYM(CommDestroy, "ncclCommDestroy")
    SYM(AllReduce, "ncclAllReduce")
    SYM(AllGather, "ncclAllGather")
    SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
    return &a;
}

#define SCKM_NCCL(ctx, call)                                                                       \
    do {                                                                                           \
        ncclResult_t_ _r = (call);                                                                 \
        if (_r != 0)                                                                               \
            return fail((ctx), SCKM_ERR_NCCL, "%s failed: %s", #call, api()->GetErrorString(_r));  \
    } while (0)

int nccl_unique_id(sckm_ctx* ctx, void* id128) {
    NcclApi* a = api();
    if (!a->handle) return fail(ctx, SCKM_ERR_NCCL, "%s", a->load_error.c_str());
    ncclUniqueId_t id;
    SCKM_NCCL(ctx, a->GetUniqueId(&id));
    memcpy(id128, &id, 128);
    return SCKM_OK;
}

int nccl_init_rank(sckm_ctx* ctx, int nranks, int rank, const void* id128) {
    if (nranks < 1 || rank < 0 || rank >= nranks) return fail(ctx, SCKM_ERR_INVALID, "bad rank %d of %d", rank, nranks);
    if (ctx->nccl_comm) return fail(ctx, SCKM_ERR_STATE, "communicator already initialised");
    if (nranks > 1) {
        NcclApi* a = api();
        if (!a->handle) return fail(ctx, SCKM_ERR_NCCL, "%s", a->load_error.c_str());
        ncclUniqueId_t id;
        memcpy(&id, id128, 128);
        SCKM_CUDA(ctx, cudaSetDevice(ctx->device));
        SCKM_NCCL(ctx, a->CommInitRank(&ctx->nccl_comm, nranks, id, rank));
    }
    ctx->nranks = nranks; ctx->rank = rank;
    if (ctx->d_totals) { cudaFree(ctx->d_totals); ctx->d_totals = nullptr; }
    SCKM_CUDA(ctx, cudaMalloc((void**)&ctx->d_totals, sizeof(double) * (size_t)std::max(nranks, 1)));
    SCKM_CUDA(ctx, cudaMemset(ctx->d_totals, 0, sizeof(double) * (size_t)std::max(nranks, 1)));
    if (nranks > 1) {
        // NCCL sets up its channels lazily on the first collective of each kind (~1 s on 8 GPUs): pay that here,
        // not inside the first kmeans++ pass / Lloyd step.  The buffer holds zeros, so the results are zeros again.
        SCKM_TRY(nccl_allgather_f64(ctx, ctx->d_totals + rank, ctx->d_totals));
        SCKM_TRY(nccl_allreduce_f64(ctx, ctx->d_totals, 1));
        SCKM_TRY(nccl_allreduce_u64(ctx, (unsigned long long*)ctx->d_totals, 1));
        SCKM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return SCKM_OK;
}

void nccl_destroy(sckm_ctx* ctx) {
    if (ctx->nccl_comm && api()->handle) api()->CommDestroy(ctx->nccl_comm);
    ctx->nccl_comm = nullptr;
}

int nccl_allreduce_f64(sckm_ctx* ctx, double* buf, size_t count) {
    if (ctx->nranks <= 1) return SCKM_OK;
    SCKM_NCCL(ctx, api()->AllReduce(buf, buf, count, ncclFloat64_, ncclSum_, ctx->nccl_comm, ctx->stream));
    ctx->launches++;
    return SCKM_OK;
}

int nccl_allreduce_u64(sckm_ctx* ctx, unsigned long long* buf, size_t count) {
    if (ctx->nranks <= 1) return SCKM_OK;
    SCKM_NCCL(ctx, api()->AllReduce(buf, buf, count, ncclUint64_, ncclSum_, ctx->nccl_comm, ctx->stream));
    ctx->launches++;
    return SCKM_OK;
}

int nccl_allgather_f64(sckm_ctx* ctx, const double* send1, double* recv) {
    if (ctx->nranks <= 1) return SCKM_OK;
    SCKM_NCCL(ctx, api()->AllGather(send1, recv, 1, ncclFloat64_, ctx->nccl_comm, ctx->stream));
    ctx->launches++;
    return SCKM_OK;
}

int nccl_allgather_bytes(sckm_ctx* ctx, const void* send, void* recv, size_t bytes_per_rank) {
    if (ctx->nranks <= 1) return SCKM_OK;
    SCKM_NCCL(ctx, api()->AllGather(send, recv, bytes_per_rank, ncclUint8_, ctx->nccl_comm, ctx->stream));
    ctx->launches++;
    return SCKM_OK;
}

}  // namespace sckm
