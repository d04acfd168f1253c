// sckm_peer.cu -- the per-step all-reduce of a multi-GPU Lloyd loop, done by the reduce and finalize kernels themselves
// over peer memory.
//
// The step of `KMeans::fit` (kmeans.rs:294-310) ends, on G GPUs, with the sum over the ranks of one small vector
// [k*d sums | k counts | inertia] (133 KB at config C3, 1.05 MB at C4/C5) followed by centroids = sums / counts and the
// stop rule.  Through NCCL that is reduce_partials -> ncclAllReduce -> finalize: three launches and a protocol of its own
// for a payload that one NVLink hop moves in a microsecond.  Here every rank owns a receive area that all the other
// ranks map (cudaIpc between processes -- the torchrun form -- or plain peer access inside one process -- the
// sckm_ctx_create_multi form), and the all-reduce disappears into the two kernels that surround it (sckm_kernels.cu):
//
//   reduce_partials_kernel   the thread that holds an element of the rank's vector stores it into EVERY rank's receive
//                            area (G stores over NVSwitch, fire and forget), each value in a 16-byte cell
//                            {lo32, tag, hi32, tag} whose 8-byte halves arrive atomically -- the tag is the flag
//   finalize_kernel          each thread reads the G cells of ITS elements from its OWN memory (spinning on the tags of
//                            the few that are still in flight), adds them in rank order 0..G-1 and carries on with
//                            the division / norms / stop rule as on one GPU.
//
// One-shot: G * payload * 2 bytes leave every rank per step (2.1 MB at C3, 17 MB at C4/C5 with G = 8), nothing is
// forwarded, there is no second phase, no flag round trip and no launch: the critical path is ONE one-way NVLink hop
// between the end of the reduce and the start of the finalize.  The sum order is fixed, so all ranks get bit-identical
// centroids and inertia (the stop rule must fire on the same iteration everywhere) and a run is reproducible.
// Two halves (exchange number & 1) suffice: a rank can only reach the reduce of exchange e+2 after ITS finalize of e+1
// has read every rank's e+1 cells, which each rank sends only after its finalize of e has completed (stream order), so
// nobody is still reading the half that e+2 overwrites.  Skipped iterations (the loop already stopped: every kernel
// returns at once, on every rank alike) send nothing; the NCCL barrier at the start of each loop covers the gap they
// leave in that argument.
//
// If any rank cannot map any other (no peer access, IPC refused), all ranks agree -- through NCCL -- to stay on the
// NCCL path.  SCKM_PEER_ALLREDUCE=0 forces that.
#include "sckm_common.cuh"
#include <algorithm>
#include <cstdlib>
#include <random>
#include <unistd.h>

namespace sckm {

constexpr size_t kPeerMinCap = 1u << 15;      // cells per (half, source): covers config C3 (16.6K); C4 / C5 grow once

struct PeerXchg {
    bool ready = false;         // every rank mapped every rank: finalize_kernel does the all-reduce
    bool given_up = false;      // the ranks agreed to stay on NCCL (decided once per context)
    char* base = nullptr;       // this rank's receive area: [2 halves][G sources][cap cells of 16 bytes]
    bool exported = false;      // `base` was handed to another process (cudaIpc)
    size_t cap = 0;             // cells per (half, source)
    unsigned long long epoch = 0;
    std::vector<void*> opened;  // cudaIpc mappings of the other ranks' areas
    // device-resident pieces: [G] pointers to the ranks' receive areas, local words {unused, sync, err}
    uint4** d_recv = nullptr;
    unsigned long long* d_words = nullptr;
};

struct PeerInfo {               // what every rank tells every other
    unsigned long long pid, token, ptr;
    int dev, ok, have_handle, pad;
    cudaIpcMemHandle_t handle;
};

static unsigned long long process_token() {
    static const unsigned long long t = ((unsigned long long)std::random_device{}() << 32) ^ std::random_device{}();
    return t;
}

static void peer_unmap(sckm_ctx* ctx, bool free_exported) {
    PeerXchg* p = ctx->peer;
    for (void* q : p->opened) cudaIpcCloseMemHandle(q);
    p->opened.clear();
    // a buffer another process may still have mapped is left to the driver (freeing it under an open mapping is undefined)
    if (p->base && (free_exported || !p->exported)) cudaFree(p->base);
    p->base = nullptr; p->exported = false; p->cap = 0; p->ready = false;
    cudaGetLastError();
}

// one u64 through NCCL: sum of `mine` over the ranks (also the barrier of this file)
static int agree(sckm_ctx* ctx, unsigned long long mine, unsigned long long* sum) {
    PeerXchg* p = ctx->peer;
    SCKM_CUDA(ctx, cudaMemcpyAsync(p->d_words + 1, &mine, 8, cudaMemcpyHostToDevice, ctx->stream));
    SCKM_TRY(nccl_allreduce_u64(ctx, p->d_words + 1, 1));
    unsigned long long got = 0;
    SCKM_CUDA(ctx, cudaMemcpyAsync(&got, p->d_words + 1, 8, cudaMemcpyDeviceToHost, ctx->stream));
    SCKM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (sum) *sum = got;
    return SCKM_OK;
}

// Collective.  Between the collectives nothing returns early on a LOCAL failure: the failure travels in `ok` and all
// ranks take the same exit.
static int peer_build(sckm_ctx* ctx, size_t pk) {
    PeerXchg* p = ctx->peer;
    const int G = ctx->nranks;
    if (!p->d_words) {
        SCKM_CUDA(ctx, cudaMalloc((void**)&p->d_words, 64));
        SCKM_CUDA(ctx, cudaMemsetAsync(p->d_words, 0, 64, ctx->stream));
        SCKM_CUDA(ctx, cudaMalloc((void**)&p->d_recv, sizeof(void*) * (size_t)G));
    }
    if (p->base) {                      // growing: every rank drops its mappings, then (barrier) its own buffer
        SCKM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        for (void* q : p->opened) cudaIpcCloseMemHandle(q);
        p->opened.clear();
        SCKM_TRY(agree(ctx, 0, nullptr));
        peer_unmap(ctx, /*free_exported=*/true);
    }
    int ok = 1;
    const size_t cap = std::max(kPeerMinCap, (pk + 1023) / 1024 * 1024);
    const size_t bytes = 2 * (size_t)G * cap * sizeof(uint4);      // zeroed: tag 0 never matches an exchange
    if (cudaMalloc((void**)&p->base, bytes) != cudaSuccess || cudaMemset(p->base, 0, bytes) != cudaSuccess) {
        cudaGetLastError();
        if (p->base) cudaFree(p->base);
        p->base = nullptr; ok = 0;
    }
    PeerInfo mine;
    memset(&mine, 0, sizeof(mine));
    mine.pid = (unsigned long long)getpid(); mine.token = process_token(); mine.ptr = (unsigned long long)(uintptr_t)p->base;
    mine.dev = ctx->device; mine.ok = ok;
    if (p->base && cudaIpcGetMemHandle(&mine.handle, p->base) == cudaSuccess) mine.have_handle = 1;
    cudaGetLastError();
    std::vector<PeerInfo> all((size_t)G);
    char* d_x = nullptr;
    SCKM_CUDA(ctx, cudaMalloc((void**)&d_x, sizeof(PeerInfo) * (size_t)(G + 1)));
    int rc = SCKM_OK;
    if (cudaMemcpyAsync(d_x, &mine, sizeof(PeerInfo), cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) rc = SCKM_ERR_CUDA;
    if (rc == SCKM_OK) rc = nccl_allgather_bytes(ctx, d_x, d_x + sizeof(PeerInfo), sizeof(PeerInfo));
    if (rc == SCKM_OK && (cudaMemcpyAsync(all.data(), d_x + sizeof(PeerInfo), sizeof(PeerInfo) * (size_t)G, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
                          cudaStreamSynchronize(ctx->stream) != cudaSuccess)) rc = SCKM_ERR_CUDA;
    cudaFree(d_x);
    if (rc != SCKM_OK) return rc == SCKM_ERR_CUDA ? fail(ctx, rc, "peer exchange setup: %s", cudaGetErrorString(cudaGetLastError())) : rc;
    std::vector<char*> maps((size_t)G, nullptr);
    for (int r = 0; r < G && ok; r++) {
        const PeerInfo& o = all[(size_t)r];
        if (!o.ok) { ok = 0; break; }
        if (r == ctx->rank) { maps[(size_t)r] = p->base; continue; }
        if (o.pid == mine.pid && o.token == mine.token) {          // same process: plain peer access
            int can = 0;
            if (o.dev == ctx->device) can = 1;
            else if (cudaDeviceCanAccessPeer(&can, ctx->device, o.dev) == cudaSuccess && can) {
                const cudaError_t e = cudaDeviceEnablePeerAccess(o.dev, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) can = 0;
            }
            cudaGetLastError();
            if (!can) { ok = 0; break; }
            maps[(size_t)r] = (char*)(uintptr_t)o.ptr;
        } else {                                                   // another process of this node: cudaIpc
            void* q = nullptr;
            if (!o.have_handle || cudaIpcOpenMemHandle(&q, o.handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = 0; break; }
            p->opened.push_back(q);
            maps[(size_t)r] = (char*)q;
        }
    }
    for (int r = 0; r < G; r++)
        if (r != ctx->rank && !(all[(size_t)r].pid == mine.pid && all[(size_t)r].token == mine.token) && mine.have_handle) p->exported = true;
    unsigned long long bad = 0;
    SCKM_TRY(agree(ctx, ok ? 0ull : 1ull, &bad));
    if (bad) {                           // somebody could not map somebody: everyone stays on NCCL, for good
        peer_unmap(ctx, /*free_exported=*/false);
        p->given_up = true;
        if (getenv("SCKM_PEER_TRACE")) fprintf(stderr, "[sckm] rank %d: peer all-reduce unavailable (%llu rank(s) failed to map), using NCCL\n", ctx->rank, bad);
        return SCKM_OK;
    }
    SCKM_CUDA(ctx, cudaMemcpyAsync(p->d_recv, maps.data(), sizeof(void*) * (size_t)G, cudaMemcpyHostToDevice, ctx->stream));
    SCKM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    p->cap = cap;
    p->ready = true;
    if (getenv("SCKM_PEER_TRACE")) fprintf(stderr, "[sckm] rank %d: peer all-reduce over %d ranks, %zu cells per source (%.1f MB receive area)\n", ctx->rank, G, cap, bytes / 1048576.0);
    return SCKM_OK;
}

int peer_prepare(sckm_ctx* ctx, size_t pk, bool* use) {
    *use = false;
    if (ctx->nranks <= 1) return SCKM_OK;
    if (const char* e = getenv("SCKM_PEER_ALLREDUCE")) if (atoi(e) == 0) return SCKM_OK;
    if (!ctx->peer) ctx->peer = new PeerXchg();
    PeerXchg* p = ctx->peer;
    if (p->given_up) return SCKM_OK;
    SCKM_CUDA(ctx, cudaSetDevice(ctx->device));
    if (!(p->ready && p->cap >= pk)) SCKM_TRY(peer_build(ctx, pk));
    // loop-start barrier: no rank is still reading a half written by an earlier loop (see the header)
    if (p->ready) SCKM_TRY(nccl_allreduce_u64(ctx, p->d_words + 1, 1));
    *use = p->ready;
    return SCKM_OK;
}

void peer_next(sckm_ctx* ctx) { ctx->peer->epoch++; }

PeerArgs peer_args(const sckm_ctx* ctx) {
    const PeerXchg* p = ctx->peer;
    PeerArgs a;
    a.recv = p->d_recv; a.err = (unsigned int*)(p->d_words + 2); a.packed_out = ctx->d_packed; a.cap = p->cap;
    a.tag = (uint32_t)(p->epoch % 0xFFFFFFFFull) + 1u; a.half = (uint32_t)(p->epoch & 1ull);
    a.G = (uint32_t)ctx->nranks; a.rank = (uint32_t)ctx->rank;
    return a;
}

int peer_check(sckm_ctx* ctx) {
    if (!ctx->peer || !ctx->peer->ready) return SCKM_OK;
    unsigned int err = 0;
    SCKM_CUDA(ctx, cudaMemcpyAsync(&err, ctx->peer->d_words + 2, 4, cudaMemcpyDeviceToHost, ctx->stream));
    SCKM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (err) return fail(ctx, SCKM_ERR_STATE, "peer all-reduce: the partial sums of a rank did not arrive within the time limit");
    return SCKM_OK;
}

void peer_destroy(sckm_ctx* ctx) {
    if (!ctx->peer) return;
    PeerXchg* p = ctx->peer;
    peer_unmap(ctx, /*free_exported=*/false);
    cudaFree(p->d_recv); cudaFree(p->d_words);
    cudaGetLastError();
    delete p;
    ctx->peer = nullptr;
}

}  // namespace sckm
