// sckm_ingest.cu -- host <-> device transfer engine of the drop-in boundary.
//
// The reference's containers are ordinary heap memory: `DenseMatrix.values: Vec<T>` (src/linalg/basic/matrix.rs:27-32)
// and the label vector `Vec<TY>` allocated in predict (src/cluster/kmeans.rs:329).  Such pageable memory cannot be
// DMA'd directly; a plain cudaMemcpy bounces it through one driver-owned staging buffer on ONE host thread
// (measured 7-12 GB/s on the B200 box), far below what PCIe Gen5 x16 delivers from pinned memory (~53 GB/s).
//
// Engine: T <= 8 host threads, each with its own CUDA stream and two pinned 4 MiB buffers, pinned by that thread on
// first use; a transfer gets one lane per 128 MB (pinning is slow, ~1.5 GB/s) plus every lane that is already pinned.  Thread t takes chunks
// t, t+T, t+2T, ... of the transfer; for every chunk it memcpy()s pageable -> pinned (host DRAM bandwidth, in
// parallel across threads) and queues the DMA pinned -> device on its stream, reusing a buffer only after the event
// of its previous DMA has completed.  Device -> host runs the same ring backwards.  Pinned or registered caller
// memory (cudaPointerGetAttributes) skips the engine: one DMA straight from the caller's buffer.
//
// The caller's pointer is only dereferenced for the duration of the call (ownership rule of the C ABI).
#include "sckm_common.cuh"
#include <algorithm>
#include <atomic>
#include <thread>
#include <pthread.h>
#include <sched.h>

namespace sckm {

namespace {

constexpr size_t kChunk = 4u << 20;          // bytes per DMA
constexpr size_t kStagedMin = 32u << 20;     // below this a plain cudaMemcpy is as fast as spinning up the ring
constexpr int kMaxThreads = 16;

struct Lane {
    cudaStream_t stream = nullptr;
    void* buf[2] = {nullptr, nullptr};
    cudaEvent_t ev[2] = {nullptr, nullptr};
};

}  // namespace

struct StagePool {
    int device = 0;
    int nthreads = 0;
    Lane lane[kMaxThreads];
    bool have_cpus = false;          // CPUs of the NUMA node the device hangs off (sysfs); staging threads run there so
    cpu_set_t cpus;                  // that the pinned buffers (first touch) and the memcpy traffic stay node-local

    ~StagePool() {
        cudaSetDevice(device);
        for (int t = 0; t < nthreads; t++) {
            for (int b = 0; b < 2; b++) {
                if (lane[t].ev[b]) cudaEventDestroy(lane[t].ev[b]);
                if (lane[t].buf[b]) cudaFreeHost(lane[t].buf[b]);
            }
            if (lane[t].stream) cudaStreamDestroy(lane[t].stream);
        }
    }
};

static int pool_threads(const sckm_ctx* ctx) {
    if (const char* e = getenv("SCKM_INGEST_THREADS")) return std::max(1, std::min(kMaxThreads, atoi(e)));
    const unsigned hc = std::thread::hardware_concurrency();
    int t = (int)std::max(2u, std::min<unsigned>(8u, hc ? hc / 2 : 4));
    if (ctx->ingest_max_threads > 0) t = std::min(t, ctx->ingest_max_threads);   // several devices share the host cores
    return t;
}

// lanes are pinned on first use, by the thread that drives them (see the sizing rule in staged_copy)
static bool lane_ready(Lane& l) {
    if (l.stream) return true;
    bool ok = cudaStreamCreateWithFlags(&l.stream, cudaStreamNonBlocking) == cudaSuccess;
    for (int b = 0; b < 2 && ok; b++)
        ok = cudaHostAlloc(&l.buf[b], kChunk, cudaHostAllocDefault) == cudaSuccess &&
             cudaEventCreateWithFlags(&l.ev[b], cudaEventDisableTiming) == cudaSuccess;
    if (!ok) cudaGetLastError();
    return ok;
}

// cpulist of the device's NUMA node: /sys/bus/pci/devices/<bus id>/numa_node -> /sys/devices/system/node/nodeN/cpulist
static bool device_numa_cpus(int device, cpu_set_t* out) {
    char bus[32] = {0};
    if (cudaDeviceGetPCIBusId(bus, sizeof(bus), device) != cudaSuccess) { cudaGetLastError(); return false; }
    for (char* p = bus; *p; p++) if (*p >= 'A' && *p <= 'Z') *p = (char)(*p - 'A' + 'a');
    char path[160];
    snprintf(path, sizeof(path), "/sys/bus/pci/devices/%s/numa_node", bus);
    FILE* f = fopen(path, "r");
    if (!f) return false;
    int node = -1;
    const int got = fscanf(f, "%d", &node);
    fclose(f);
    if (got != 1 || node < 0) return false;
    snprintf(path, sizeof(path), "/sys/devices/system/node/node%d/cpulist", node);
    f = fopen(path, "r");
    if (!f) return false;
    char list[4096] = {0};
    const bool ok = fgets(list, sizeof(list), f) != nullptr;
    fclose(f);
    if (!ok) return false;
    CPU_ZERO(out);
    int n = 0;
    for (const char* p = list; *p;) {                            // "0-31,64-95"
        if (*p < '0' || *p > '9') { p++; continue; }
        char* end = nullptr;
        long a = strtol(p, &end, 10), b = a;
        if (*end == '-') b = strtol(end + 1, &end, 10);
        for (long c = a; c <= b && c < CPU_SETSIZE; c++) { CPU_SET((int)c, out); n++; }
        p = end;
    }
    // only keep CPUs this process may run on (cgroup / taskset)
    cpu_set_t allowed;
    if (sched_getaffinity(0, sizeof(allowed), &allowed) == 0) {
        n = 0;
        for (int c = 0; c < CPU_SETSIZE; c++) { if (CPU_ISSET(c, out) && !CPU_ISSET(c, &allowed)) CPU_CLR(c, out); if (CPU_ISSET(c, out)) n++; }
    }
    return n > 0;
}

static StagePool* get_pool(sckm_ctx* ctx) {
    if (!ctx->stage_pool) {
        ctx->stage_pool = new StagePool();
        ctx->stage_pool->device = ctx->device;
        ctx->stage_pool->nthreads = pool_threads(ctx);
        // bind only when the node offers every lane a CPU of its own among those this process may use (a container that
        // is confined to a few CPUs of the node would otherwise pile all lanes onto them)
        if (!getenv("SCKM_INGEST_NO_NUMA") && device_numa_cpus(ctx->device, &ctx->stage_pool->cpus))
            ctx->stage_pool->have_cpus = CPU_COUNT(&ctx->stage_pool->cpus) >= std::max(4, 2 * ctx->stage_pool->nthreads);
        if (getenv("SCKM_TRACE"))
            fprintf(stderr, "[sckm] ingest: device %d, %d lanes, NUMA binding %s (%d usable CPUs on the device's node)\n", ctx->device,
                    ctx->stage_pool->nthreads, ctx->stage_pool->have_cpus ? "on" : "off", ctx->stage_pool->have_cpus ? CPU_COUNT(&ctx->stage_pool->cpus) : 0);
    }
    return ctx->stage_pool;
}

void ingest_destroy(sckm_ctx* ctx) {
    delete ctx->stage_pool;
    ctx->stage_pool = nullptr;
    if (ctx->copy_stream) { cudaStreamDestroy(ctx->copy_stream); ctx->copy_stream = nullptr; }
}

static bool is_pageable(const void* host) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, host) != cudaSuccess) { cudaGetLastError(); return true; }
    return a.type == cudaMemoryTypeUnregistered;
}

// dst/src: exactly one of them is device memory (to_device selects which).  Blocking; on return the bytes have
// landed and the host buffer is no longer referenced.
static int staged_copy(sckm_ctx* ctx, void* dst, const void* src, size_t bytes, bool to_device, bool sync_first) {
    if (bytes == 0) return SCKM_OK;
    // whatever produced the device side (or still reads it) is ordered on the context's stream; a caller that
    // knows the device buffer is idle (double buffering) may skip the wait so the transfer overlaps its kernels
    if (sync_first) SCKM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    const void* host = to_device ? src : dst;
    const cudaMemcpyKind kind = to_device ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost;
    // Lanes that are already pinned are free to use; pinning a new one costs ~5 ms (8 MiB at ~1.5 GB/s), which only a
    // lane that will carry >= 128 MB repays -- of this transfer or of the operation it belongs to (ctx->ingest_hint).
    int T = 0;
    StagePool* pool = nullptr;
    if (bytes >= kStagedMin && is_pageable(host) && !getenv("SCKM_INGEST_DIRECT")) {
        pool = get_pool(ctx);
        int ready = 0;
        while (ready < pool->nthreads && pool->lane[ready].stream) ready++;
        const size_t scope = std::max(bytes, ctx->ingest_hint);
        const int want = getenv("SCKM_INGEST_FORCE") ? pool->nthreads      // tests: exercise the ring on small arrays
                         : (int)std::min<size_t>((size_t)pool->nthreads, scope / ((size_t)128 << 20));
        const size_t nchunks = (bytes + kChunk - 1) / kChunk;
        T = (int)std::min<size_t>(nchunks, (size_t)std::max(ready, want));
    }
    if (T == 0) {
        if (sync_first) {
            SCKM_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, kind, ctx->stream));
            SCKM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        } else {                                               // not behind the kernels queued on ctx->stream
            if (!ctx->copy_stream) SCKM_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
            SCKM_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, kind, ctx->copy_stream));
            SCKM_CUDA(ctx, cudaStreamSynchronize(ctx->copy_stream));
        }
        return SCKM_OK;
    }
    const size_t nchunks = (bytes + kChunk - 1) / kChunk;
    std::atomic<int> err{(int)cudaSuccess};
    auto work = [&](int t) {
        Lane& l = pool->lane[t];
        cudaError_t e = cudaSetDevice(pool->device);
        if (e == cudaSuccess && !lane_ready(l)) e = cudaErrorMemoryAllocation;
        int slot = 0;
        bool used[2] = {false, false};
        size_t pend_off[2] = {0, 0}, pend_len[2] = {0, 0};
        for (size_t c = (size_t)t; c < nchunks && e == cudaSuccess && err.load(std::memory_order_relaxed) == (int)cudaSuccess;
             c += (size_t)T, slot ^= 1) {
            const size_t off = c * kChunk, len = std::min(kChunk, bytes - off);
            if (used[slot]) {                                   // the buffer's previous DMA must have drained
                e = cudaEventSynchronize(l.ev[slot]);
                if (e != cudaSuccess) break;
                if (!to_device) memcpy((char*)dst + pend_off[slot], l.buf[slot], pend_len[slot]);
            }
            if (to_device) {
                memcpy(l.buf[slot], (const char*)src + off, len);
                e = cudaMemcpyAsync((char*)dst + off, l.buf[slot], len, kind, l.stream);
            } else {
                e = cudaMemcpyAsync(l.buf[slot], (const char*)src + off, len, kind, l.stream);
                pend_off[slot] = off; pend_len[slot] = len;
            }
            if (e == cudaSuccess) e = cudaEventRecord(l.ev[slot], l.stream);
            used[slot] = true;
        }
        if (e == cudaSuccess) e = cudaStreamSynchronize(l.stream);
        if (e == cudaSuccess && !to_device)                     // drain what is still parked in the pinned buffers
            for (int s = 0; s < 2; s++) {
                const int b = slot ^ s;                         // oldest first
                if (used[b] && pend_len[b]) { memcpy((char*)dst + pend_off[b], l.buf[b], pend_len[b]); pend_len[b] = 0; }
            }
        if (e != cudaSuccess) err.store((int)e);
    };
    // all lanes run on threads of their own (not the caller's): they bind themselves to the device's NUMA node
    auto bound_work = [&](int t) {
        if (pool->have_cpus) pthread_setaffinity_np(pthread_self(), sizeof(cpu_set_t), &pool->cpus);
        work(t);
    };
    std::vector<std::thread> th;
    th.reserve(T);
    for (int t = 0; t < T; t++) th.emplace_back(bound_work, t);
    for (auto& x : th) x.join();
    if (err.load() != (int)cudaSuccess)
        return fail(ctx, SCKM_ERR_CUDA, "staged %s copy failed: %s", to_device ? "H2D" : "D2H",
                    cudaGetErrorString((cudaError_t)err.load()));
    return SCKM_OK;
}

int copy_to_device(sckm_ctx* ctx, void* dst_dev, const void* src_host, size_t bytes, bool sync_first) {
    return staged_copy(ctx, dst_dev, src_host, bytes, true, sync_first);
}

int copy_to_host(sckm_ctx* ctx, void* dst_host, const void* src_dev, size_t bytes) {
    return staged_copy(ctx, dst_host, src_dev, bytes, false, true);
}

}  // namespace sckm
