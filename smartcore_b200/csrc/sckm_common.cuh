// sckm_common.cuh -- shared declarations of libsmartcore_kmeans_cuda.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstddef>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/smartcore_kmeans_cuda.h"

namespace sckm {

constexpr int kKppBlockRows = 1024;   // fixed summation unit of the kmeans++ D^2 array

// pitch (in doubles) of one partial slot [k*d sums | k counts | inertia]: 128-byte aligned rows
inline size_t slot_pitch(size_t pk) { return (pk + 15) / 16 * 16; }

struct MultiExt;    // sckm_multi.cu
struct PeerXchg;    // sckm_peer.cu
struct StagePool;   // pinned staging ring of the host<->device transfer engine (sckm_ingest.cu)

// Device-resident state of the loop of KMeans::fit (kmeans.rs:294-310).  The stop rule `if distortion <= dist { break }`
// (kmeans.rs:305-309) is evaluated on the device by the finalize kernel of each iteration, so the host can enqueue
// several iterations without reading the inertia back: once `done_at` is set, every kernel of a LATER iteration
// returns at once and the state (labels, centroids, sizes) stays that of the iteration that broke the loop.
struct LoopState {
    double distortion;             // last strictly smaller inertia (f64::MAX before the first step, kmeans.rs:273)
    unsigned long long done_at;    // iteration (1-based) whose stop test fired; 0 = still running
    unsigned long long iters;      // clustering steps executed so far
    unsigned long long honor_stop; // 0: fixed number of steps (sckm_lloyd_iterate)
};

// By-value argument of reduce_partials_kernel and finalize_kernel: the one-shot all-reduce of the step's packed vector
// over peer memory (sckm_peer.cu).  G == 0: nothing to exchange, the kernels use `packed` as they always did.
// A cell is 16 bytes {lo32(value), tag, hi32(value), tag}: each 8-byte half travels atomically, so a reader that finds
// both tags equal to this exchange's tag holds the value -- no separate flag, no fence between data and flag.
struct PeerArgs {
    uint4* const* recv;                  // [G] every rank's receive area, mapped into this device:
                                         //     recv[r][(half * G + src) * cap + e] = element e of rank src's vector, sent to rank r
    unsigned int* err;                   // local: set when a wait ran into its time limit
    double* packed_out;                  // local d_packed: receives the all-reduced vector
    unsigned long long cap;              // cells per (half, source rank)
    uint32_t tag;                        // tag of this exchange (never 0, the value of fresh memory)
    uint32_t half;                       // exchange number & 1
    uint32_t G, rank;
};

}  // namespace sckm

struct sckm_ctx {
    int device = 0;
    int num_sms = 0;
    int smem_optin = 0;          // max dynamic shared memory per CTA
    cudaStream_t stream = nullptr;
    std::string err;
    int assign_kernel = SCKM_ASSIGN_AUTO;
    uint64_t launches = 0;
    // communicator (nullptr/1 rank when single GPU)
    void* nccl_comm = nullptr;
    int nranks = 1, rank = 0;
    // workspaces, grown on demand
    double* d_centroids = nullptr;   // [k*d] current centroids
    double* d_cnorm = nullptr;       // [k] ||c - mu||^2 (GEMM-form kernels), [k] = their max
    double* d_mu = nullptr;          // [d] centring shift of the GEMM-form kernels (see launch_cnorm); zeros = no shift
    size_t cap_mu = 0;
    bool mu_zero = true;             // d_mu currently holds zeros
    bool mu_requested = false;       // what the last launch_cnorm was asked for (centred kernels vs raw norms)
    bool center_on = false;          // ... and what it decided: the tile kernels subtract mu (CENTER instantiation)
    bool packed_centered = false;    // d_packed sums of the last step are sums of (x - mu), not of x
    double* d_packed = nullptr;      // [k*d sums | k counts | inertia]
    double* d_partials = nullptr;    // [P][k*d + k + 1] per-CTA/warp partial sums (deterministic)
    size_t cap_centroids = 0, cap_packed = 0, cap_cnorm = 0, cap_size = 0, cap_seeds = 0, cap_partials = 0;
    int64_t* d_size = nullptr;       // [k]
    double* d_blocksum = nullptr;    // kmeans++: per-1024-row-block sums of D^2
    size_t cap_blocks = 0;
    double* d_totals = nullptr;      // [nranks] rank totals of D^2
    void* d_seedrow = nullptr;       // [d elements of TX, padded] + global index (8 B)
    size_t cap_seedrow = 0;
    int64_t* d_seeds = nullptr;      // [k] chosen global rows
    void* d_seedtab = nullptr;       // [k][d] of TX: the chosen seed rows (kmeans++ pruning)
    double* d_skiptab = nullptr;     // [k] pruning thresholds of the current pass
    size_t cap_seedtab = 0, cap_skiptab = 0;
    uint32_t* d_surv = nullptr;      // kmeans++: rows that survived the pruning test of the current pass
    size_t cap_surv = 0;
    unsigned* d_kppctr = nullptr;    // [0] survivor count, [1] block-sum CTAs finished
    float* d_tshift = nullptr;       // kmeans++ screening: f32(new seed - seed 0) [d]
    double* d_tshift_err = nullptr;  // ... and the length of what that rounding dropped
    size_t cap_tshift = 0;
    sckm::LoopState* d_loop = nullptr;      // stop-rule state of the running Lloyd loop
    uint32_t loop_it = 0;                   // iteration being enqueued (1-based); 0 = outside a loop (kernels never skip)
    double* d_inertia_trace = nullptr;      // [cap_trace] inertia of every step of the running loop
    size_t cap_trace = 0;
    unsigned long long* d_flags = nullptr;  // [0] rows marked as near-ties in the current step; rest: scratch
    bool cnorm_valid = false;        // d_cnorm matches d_centroids
    uint64_t ws_k = 0, ws_d = 0;     // shape the centroid workspaces currently hold
    uint32_t partial_slots_used = 0; // slots written by the last fused assignment launch
    float* d_tc5 = nullptr;          // tcgen05 path: centroid hi | lo parts (TF32) and -||c||^2/2 in f32
    size_t cap_tc5 = 0;
    void* d_flush = nullptr;         // L2 flush buffer
    size_t flush_bytes = 0;
    double* h_pinned = nullptr;      // small pinned scratch (>= 64 doubles)
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    sckm::PeerXchg* peer = nullptr;          // peer-memory exchange buffers of the fused all-reduce (multi-rank only)
    int allreduce_path = 0;                  // SCKM_ALLREDUCE_* of the last Lloyd loop
    bool peer_step = false;                  // the step being enqueued sums over the ranks inside reduce_partials / finalize
    sckm::MultiExt* multi = nullptr;         // per-device contexts of a multi-GPU context (sckm_ctx_create_multi)
    int ingest_max_threads = 0;              // cap on the staging threads of this context (0 = default)
    double fit_times[6] = {0, 0, 0, 0, 0, 0}; // last sckm_kmeans_fit: upload, kmeans++ + means, loop, download, total [s], devices
    size_t ingest_hint = 0;                  // bytes the current multi-transfer operation will move in total (0 = unknown)
    cudaStream_t copy_stream = nullptr;      // transfers that must not queue behind the kernels on `stream`
    sckm::StagePool* stage_pool = nullptr;   // pinned ring for transfers from/to pageable memory (lanes pinned on first use)
};

struct sckm_dataset {
    sckm_ctx* ctx = nullptr;
    void* x = nullptr;               // row-major [n][d] of dtype
    uint64_t n = 0, d = 0;
    int dtype = SCKM_F64;
    uint64_t row_offset = 0, n_global = 0;
    uint32_t* labels = nullptr;      // [n]
    double* mind = nullptr;          // [n] kmeans++ D^2 / per-point min distance
    uint64_t* labels64 = nullptr;    // lazily allocated widening buffer for usize downloads
    float* x32 = nullptr;            // f32 shadow of an f64 X (tcgen05 ranking only; lazily built)
    uint16_t* kpp_shadow = nullptr;  // kmeans++ only: bf16(x - seed 0) [n][d], built by the first pass, freed after the last
    float* kpp_shadow_err = nullptr; // ... and ||(x - seed 0) - shadow row|| rounded up [n]
    bool have_labels = false;
    size_t elem() const { return dtype == SCKM_F32 ? 4 : 8; }
};

namespace sckm {

int fail(sckm_ctx* ctx, int code, const char* fmt, ...);

// Transient device buffers (datasets, staging images, the kmeans++ shadow) come from the device's stream-ordered
// memory pool, configured at context creation to KEEP what is freed: releasing a multi-GB block to the driver and
// mapping it again on the next fit costs anything from 1 to 500 ms on this box (measured: cudaFree of the 1.3 GB
// shadow), stream-ordered frees are microseconds.  Allocation and release are ordered on ctx->stream.
inline cudaError_t dev_alloc(sckm_ctx* ctx, void** p, size_t bytes) {
    *p = nullptr;
    cudaError_t e = cudaMallocAsync(p, bytes ? bytes : 1, ctx->stream);
    if (e != cudaSuccess) { cudaGetLastError(); *p = nullptr; }
    return e;
}
inline void dev_free(sckm_ctx* ctx, void* p) {
    if (p) cudaFreeAsync(p, ctx->stream);
}
// hand the pool's idle blocks back to the driver (context teardown; a plain cudaMalloc that ran out of memory)
inline void dev_pool_trim(sckm_ctx* ctx) {
    cudaMemPool_t pool = nullptr;
    if (cudaDeviceGetDefaultMemPool(&pool, ctx->device) == cudaSuccess) {
        cudaStreamSynchronize(ctx->stream);
        cudaMemPoolTrimTo(pool, 0);
    }
    cudaGetLastError();
}
// long-lived workspaces use plain cudaMalloc; if that fails while the pool sits on idle memory, trim and retry once
inline cudaError_t ws_malloc(sckm_ctx* ctx, void** p, size_t bytes) {
    cudaError_t e = cudaMalloc(p, bytes);
    if (e == cudaErrorMemoryAllocation) { cudaGetLastError(); dev_pool_trim(ctx); e = cudaMalloc(p, bytes); }
    return e;
}

// the (state, iteration) pair every kernel of a Lloyd step receives: it returns at once when the loop already stopped
#define SCKM_LOOP_ARGS(ctx) ((ctx)->loop_it ? (ctx)->d_loop : nullptr), (ctx)->loop_it
#ifdef __CUDACC__
__device__ __forceinline__ bool loop_done(const LoopState* st, uint32_t it) {
    return st != nullptr && st->done_at != 0ull && st->done_at < (unsigned long long)it;
}
#endif

#ifdef __CUDACC__
// Programmatic dependent launch (PDL): a kernel launched with launch_pdl() may be scheduled while its predecessor on the
// stream is still draining; it must call pdl_wait() before touching anything the predecessor wrote (the wait returns
// once that grid has completed and its writes are visible).  For the tiny kernels around a 40 us assignment launch
// (refine, reduce, finalize: config C2) this hides most of the launch latency of each boundary.  A kernel launched the
// ordinary way sees pdl_wait() as a no-op.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
#endif

#define SCKM_CUDA(ctx, call)                                                               \
    do {                                                                                   \
        cudaError_t _e = (call);                                                           \
        if (_e != cudaSuccess)                                                             \
            return sckm::fail((ctx), SCKM_ERR_CUDA, "%s failed: %s (%s:%d)", #call,        \
                              cudaGetErrorString(_e), __FILE__, __LINE__);                 \
    } while (0)

#define SCKM_TRY(expr)                   \
    do {                                 \
        int _rc = (expr);                \
        if (_rc != SCKM_OK) return _rc;  \
    } while (0)

// ---- NCCL (sckm_nccl.cu) ----
int nccl_unique_id(sckm_ctx* ctx, void* id128);
int nccl_init_rank(sckm_ctx* ctx, int nranks, int rank, const void* id128);
void nccl_destroy(sckm_ctx* ctx);
int nccl_allreduce_f64(sckm_ctx* ctx, double* buf, size_t count);
int nccl_allreduce_u64(sckm_ctx* ctx, unsigned long long* buf, size_t count);
int nccl_allgather_f64(sckm_ctx* ctx, const double* send1, double* recv);  // 1 double per rank
int nccl_allgather_bytes(sckm_ctx* ctx, const void* send, void* recv, size_t bytes_per_rank);

// ---- fused one-shot all-reduce over peer memory (sckm_peer.cu) ----
int peer_prepare(sckm_ctx* ctx, size_t pk, bool* use);   // collective: map every rank's exchange buffer; *use = false: stay on NCCL
void peer_next(sckm_ctx* ctx);                // start the next exchange (once per step, before its reduce launch)
PeerArgs peer_args(const sckm_ctx* ctx);      // kernel argument of the reduce and finalize launches of the current exchange
int peer_check(sckm_ctx* ctx);                // after a stream synchronisation: did every wait of the loop come through
void peer_destroy(sckm_ctx* ctx);

// ---- host <-> device transfers (sckm_ingest.cu): blocking, pageable memory goes through a threaded pinned ring ----
int copy_to_device(sckm_ctx* ctx, void* dst_dev, const void* src_host, size_t bytes, bool sync_first = true);
int copy_to_host(sckm_ctx* ctx, void* dst_host, const void* src_dev, size_t bytes);
void ingest_destroy(sckm_ctx* ctx);

// ---- pieces of the API layer shared with the multi-GPU driver (sckm_api.cu) ----
int dataset_alloc(sckm_ctx* ctx, uint64_t n, uint64_t d, int dtype, uint64_t row_offset, uint64_t n_global, sckm_dataset** out);
int upload_rows(sckm_dataset* ds, const void* host, uint64_t host_rows, uint64_t lo, int column_major);
int download_labels(sckm_dataset* ds, void* out, int width);
int lloyd_loop(sckm_dataset* ds, uint64_t k, uint64_t max_iter, bool honor_stop, double* centroids_inout, int64_t* size_out,
               double* distortion_out, int64_t* iters_out, double* inertia_trace, float* ms_trace, float* assign_ms_trace);
int fit_upload(sckm_ctx* ctx, const void* x_host, uint64_t host_rows, uint64_t lo, uint64_t n_local, uint64_t d, int dtype,
               int column_major, sckm_dataset** out);
int fit_reserve(sckm_dataset* ds, uint64_t k);
int fit_compute(sckm_dataset* ds, uint64_t k, uint64_t max_iter, uint64_t first_index, const double* uniforms,
                int64_t* size_out, double* centroids_out, double* distortion_out, int64_t* iters_out, double* phase_s);
int predict_rows(sckm_ctx* ctx, const void* x_host, uint64_t host_rows, uint64_t lo, uint64_t n, uint64_t d, int dtype,
                 int column_major, const double* centroids, uint64_t k, void* labels_out, int width);

// ---- kernel launchers (sckm_kernels.cu) ----
int launch_transpose(sckm_ctx* ctx, const void* src_colmajor, void* dst_rowmajor, uint64_t n, uint64_t d, int dtype);
int launch_blobs(sckm_ctx* ctx, void* x, int dtype, uint64_t row0, uint64_t nrows, uint64_t d,
                 uint64_t n_centers, uint64_t seed);
// kmeans++ pass: D^2 refresh against the seed row in ctx->d_seedrow, label = `label` where improved;
// writes per-1024-row block sums to ctx->d_blocksum and the rank total to ctx->d_totals[rank].
int launch_kpp_refresh(sckm_dataset* ds, uint32_t label, bool first_pass, bool prune, bool want_sums = true);
int launch_kpp_seedtab(sckm_dataset* ds, uint32_t slot);
// pick the next seed: cutoff = u * sum(totals); the owning rank locates the row and publishes it
// (row + global index) in ctx->d_seedrow; other ranks publish zeros (all-reduced by the caller).
// inject_row >= 0 bypasses sampling.  Stores the global row in ctx->d_seeds[slot] (owner only; zero elsewhere).
int launch_kpp_select(sckm_dataset* ds, double u, int64_t inject_row, uint32_t slot);
// direct-form assignment against ctx->d_centroids (labels + mind)
int launch_assign_direct(sckm_dataset* ds, uint64_t k);
int launch_assign_direct_raw(sckm_ctx* ctx, const void* x, int dtype, uint64_t n, uint64_t d, uint64_t k,
                             uint32_t* labels, double* mind);
// deterministic per-label sums/counts/inertia of the local rows into ctx->d_packed
int launch_update(sckm_dataset* ds, uint64_t k, bool with_inertia);
int launch_reduce_partials(sckm_ctx* ctx, uint32_t slots, size_t pk);
// centroids = sums / counts (guarded: keep old when count == 0; unguarded for the initial means)
int launch_finalize(sckm_ctx* ctx, uint64_t k, uint64_t d, bool guarded);
int launch_loop_init(sckm_ctx* ctx, uint64_t max_iter, bool honor_stop);
int launch_labels_widen(sckm_ctx* ctx, const uint32_t* in, uint64_t* out, uint64_t n);
int measure_peaks(sckm_ctx* ctx, double* out3);

int ensure_workspace(sckm_ctx* ctx, uint64_t k, uint64_t d, size_t partial_slots);

}  // namespace sckm
