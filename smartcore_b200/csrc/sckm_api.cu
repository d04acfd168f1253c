// sckm_api.cu -- C-ABI entry points (include/smartcore_kmeans_cuda.h) and host orchestration.
//
// The driver logic mirrors KMeans::fit (src/cluster/kmeans.rs:254-323): kmeans++ labels ->
// per-label means -> loop { clustering step; centroid update; stop rule }.  Everything below runs
// on the context's own CUDA stream; the stop rule `if distortion <= dist { break }` (kmeans.rs:305) is
// evaluated on the device, the host reads a 32-byte loop state back once per batch of iterations.
#include "sckm_common.cuh"
#include "sckm_blobs.cuh"
#include <cfloat>
#include <algorithm>
#include <cstdlib>
#include <chrono>

namespace sckm {
const char* create_error_text();
int launch_assign_dmma(sckm_dataset* ds, uint64_t k);         // sckm_dmma.cu
bool dmma_supported(const sckm_dataset* ds, uint64_t k);      // sckm_dmma.cu
uint32_t dmma_partial_slots(const sckm_ctx* ctx);             // sckm_dmma.cu
int launch_predict_dmma(sckm_dataset* ds, uint64_t k);        // sckm_dmma.cu
int knn_search(sckm_dataset* ds, const void* queries_host, uint64_t nq, uint64_t k, int64_t* idx_out, double* dist_out);   // sckm_knn.cu
int radius_search(sckm_dataset* ds, const void* queries_host, uint64_t nq, double radius, int64_t* counts_out,
                  const int64_t* offsets_host, uint64_t total, int64_t* idx_out, double* dist_out);               // sckm_knn.cu
int launch_contingency(sckm_ctx* ctx, const uint32_t* d_a, const uint32_t* d_b, uint64_t n, uint64_t na, uint64_t nb,
                       unsigned long long* d_out);            // sckm_metrics.cu
int launch_assign_stream(sckm_dataset* ds, uint64_t k);       // sckm_stream.cu
bool stream_supported(const sckm_dataset* ds, uint64_t k);    // sckm_stream.cu
int launch_assign_tc5(sckm_dataset* ds, uint64_t k);          // sckm_tc5.cu
bool tc5_supported(const sckm_dataset* ds, uint64_t k);       // sckm_tc5.cu
bool tc5_auto(const sckm_dataset* ds, uint64_t k);            // sckm_tc5.cu
void multi_destroy(sckm_ctx* ctx);                            // sckm_multi.cu
bool multi_shards(const sckm_ctx* ctx, uint64_t n, std::vector<uint64_t>* bounds);   // sckm_multi.cu
int multi_kmeans_fit(sckm_ctx* ctx, const void* x_host, uint64_t n, uint64_t d, int dtype, int column_major, uint64_t k,
                     uint64_t max_iter, uint64_t first_index, const double* uniforms, void* labels_out, int width,
                     int64_t* size_out, double* centroids_out, double* distortion_out, int64_t* iters_out);   // sckm_multi.cu
int multi_predict(sckm_ctx* ctx, const void* x_host, uint64_t n, uint64_t d, int dtype, int column_major,
                  const double* centroids, uint64_t k, void* labels_out, int width);   // sckm_multi.cu
uint64_t multi_launch_count(const sckm_ctx* ctx);             // sckm_multi.cu
}
using namespace sckm;

extern "C" {

int sckm_abi_version(void) { return SCKM_ABI_VERSION; }

int sckm_ctx_create(int device, sckm_ctx** out) {
    if (!out) return fail(nullptr, SCKM_ERR_INVALID, "sckm_ctx_create: out is NULL");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(nullptr, SCKM_ERR_CUDA, "no CUDA device available (%s); this library has no CPU fallback",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    if (device < 0 || device >= count) return fail(nullptr, SCKM_ERR_INVALID, "device %d out of range (0..%d)", device, count - 1);
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess)
        return fail(nullptr, SCKM_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
    if (prop.major != 10)
        return fail(nullptr, SCKM_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a (B200) only",
                    device, prop.major, prop.minor);
    sckm_ctx* ctx = new sckm_ctx();
    ctx->device = device;
    ctx->num_sms = prop.multiProcessorCount;
    ctx->smem_optin = (int)prop.sharedMemPerBlockOptin;
#define CREATE_CUDA(call)                                                                         \
    do { cudaError_t _e = (call); if (_e != cudaSuccess) {                                        \
        int rc = fail(nullptr, SCKM_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(_e));           \
        sckm_ctx_destroy(ctx); return rc; } } while (0)
    CREATE_CUDA(cudaSetDevice(device));
    CREATE_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    {   // stream-ordered pool of this device: never hand freed blocks back to the driver behind our back (see dev_alloc)
        cudaMemPool_t pool = nullptr;
        unsigned long long keep = ~0ull;
        CREATE_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
        CREATE_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
    }
    CREATE_CUDA(cudaEventCreate(&ctx->ev0));
    CREATE_CUDA(cudaEventCreate(&ctx->ev1));
    CREATE_CUDA(cudaMalloc((void**)&ctx->d_flags, 64));
    CREATE_CUDA(cudaMemset(ctx->d_flags, 0, 64));
    CREATE_CUDA(cudaMalloc((void**)&ctx->d_totals, sizeof(double)));
    CREATE_CUDA(cudaMemset(ctx->d_totals, 0, sizeof(double)));
    CREATE_CUDA(cudaMallocHost((void**)&ctx->h_pinned, 256 * sizeof(double)));
#undef CREATE_CUDA
    *out = ctx;
    return SCKM_OK;
}

void sckm_ctx_destroy(sckm_ctx* ctx) {
    if (!ctx) return;
    multi_destroy(ctx);                       // the per-device contexts of a multi-GPU context, if any
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    peer_destroy(ctx);
    nccl_destroy(ctx);
    ingest_destroy(ctx);
    if (ctx->stream) dev_pool_trim(ctx);
    cudaFree(ctx->d_centroids); cudaFree(ctx->d_cnorm); cudaFree(ctx->d_packed); cudaFree(ctx->d_partials);
    cudaFree(ctx->d_size); cudaFree(ctx->d_blocksum); cudaFree(ctx->d_totals); cudaFree(ctx->d_seedrow);
    cudaFree(ctx->d_seeds); cudaFree(ctx->d_seedtab); cudaFree(ctx->d_skiptab); cudaFree(ctx->d_flags); cudaFree(ctx->d_surv); cudaFree(ctx->d_kppctr); cudaFree(ctx->d_tshift); cudaFree(ctx->d_tshift_err); cudaFree(ctx->d_flush); cudaFree(ctx->d_tc5); cudaFree(ctx->d_loop); cudaFree(ctx->d_inertia_trace); cudaFree(ctx->d_mu);
    if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char* sckm_last_error(const sckm_ctx* ctx) { return ctx ? ctx->err.c_str() : create_error_text(); }

int sckm_ctx_set_assign_kernel(sckm_ctx* ctx, int which) {
    if (!ctx) return SCKM_ERR_INVALID;
    if (which < SCKM_ASSIGN_AUTO || which > SCKM_ASSIGN_TC5) return fail(ctx, SCKM_ERR_INVALID, "unknown assign kernel %d", which);
    ctx->assign_kernel = which;
    return SCKM_OK;
}

uint64_t sckm_ctx_launch_count(const sckm_ctx* ctx) { return ctx ? ctx->launches + multi_launch_count(ctx) : 0; }

int sckm_comm_unique_id(sckm_ctx* ctx, void* id128) {
    if (!ctx || !id128) return SCKM_ERR_INVALID;
    return nccl_unique_id(ctx, id128);
}
int sckm_comm_init_rank(sckm_ctx* ctx, int nranks, int rank, const void* id128) {
    if (!ctx || (nranks > 1 && !id128)) return SCKM_ERR_INVALID;
    SCKM_CUDA(ctx, cudaSetDevice(ctx->device));
    return nccl_init_rank(ctx, nranks, rank, id128);
}

// ---- dataset ------------------------------------------------------------------------------
}  // extern "C"
namespace sckm {
int dataset_alloc(sckm_ctx* ctx, uint64_t n, uint64_t d, int dtype, uint64_t row_offset, uint64_t n_global,
                  sckm_dataset** out) {
    if (!ctx || !out) return SCKM_ERR_INVALID;
    *out = nullptr;
    if (dtype != SCKM_F32 && dtype != SCKM_F64) return fail(ctx, SCKM_ERR_INVALID, "dtype must be SCKM_F32 or SCKM_F64");
    if (d == 0 || d > (1u << 20)) return fail(ctx, SCKM_ERR_INVALID, "d=%llu out of range", (unsigned long long)d);
    if (n > 0xFFFFFFFFull * 16) return fail(ctx, SCKM_ERR_INVALID, "n too large");
    if (n_global == 0) n_global = n;
    if (row_offset + n > n_global) return fail(ctx, SCKM_ERR_INVALID, "rows [%llu,%llu) exceed n_global=%llu",
        (unsigned long long)row_offset, (unsigned long long)(row_offset + n), (unsigned long long)n_global);
    SCKM_CUDA(ctx, cudaSetDevice(ctx->device));
    sckm_dataset* ds = new sckm_dataset();
    ds->ctx = ctx; ds->n = n; ds->d = d; ds->dtype = dtype; ds->row_offset = row_offset; ds->n_global = n_global;
    const size_t nn = std::max<uint64_t>(n, 1);
    cudaError_t e;
    if ((e = dev_alloc(ctx, &ds->x, nn * d * ds->elem())) != cudaSuccess ||
        (e = dev_alloc(ctx, (void**)&ds->labels, nn * sizeof(uint32_t))) != cudaSuccess ||
        (e = dev_alloc(ctx, (void**)&ds->mind, nn * sizeof(double))) != cudaSuccess) {
        sckm_dataset_destroy(ds);
        return fail(ctx, SCKM_ERR_CUDA, "cudaMalloc for %llu x %llu dataset failed: %s", (unsigned long long)n,
                    (unsigned long long)d, cudaGetErrorString(e));
    }
    *out = ds;
    return SCKM_OK;
}

// Rows [lo, lo + ds->n) of a host matrix of `host_rows` rows -> ds->x (row-major on the device).  Row-major host
// memory (DenseMatrix::new(.., false), matrix.rs:187-206): one contiguous block.  Column-major (from_2d_array,
// matrix.rs:215-237): the shard is d strided runs of ds->n elements; they land as a [d][n] image (every run through
// the pinned ring) and are transposed on the device.
int upload_rows(sckm_dataset* ds, const void* host, uint64_t host_rows, uint64_t lo, int column_major) {
    sckm_ctx* ctx = ds->ctx;
    const size_t elem = ds->elem(), bytes = (size_t)ds->n * ds->d * elem;
    if (!bytes) return SCKM_OK;
    if (!column_major) return copy_to_device(ctx, ds->x, (const char*)host + (size_t)lo * ds->d * elem, bytes);
    void* tmp = nullptr;
    if (dev_alloc(ctx, &tmp, bytes) != cudaSuccess)
        return fail(ctx, SCKM_ERR_CUDA, "cudaMalloc of the %zu-byte column-major staging image failed", bytes);
    int rc = SCKM_OK;
    if (lo == 0 && host_rows == ds->n) {
        rc = copy_to_device(ctx, tmp, host, bytes);                           // the whole image is one block
    } else {
        const size_t run = (size_t)ds->n * elem;
        ctx->ingest_hint = bytes;
        for (uint64_t c = 0; c < ds->d && rc == SCKM_OK; c++)
            rc = copy_to_device(ctx, (char*)tmp + c * run, (const char*)host + ((size_t)c * host_rows + lo) * elem, run, c == 0);
        ctx->ingest_hint = 0;
    }
    if (rc == SCKM_OK) rc = launch_transpose(ctx, tmp, ds->x, ds->n, ds->d, ds->dtype);
    dev_free(ctx, tmp);
    cudaStreamSynchronize(ctx->stream);
    return rc;
}
}  // namespace sckm
extern "C" {

int sckm_dataset_upload(sckm_ctx* ctx, const void* host, uint64_t n_local, uint64_t d, int dtype,
                        int column_major, uint64_t row_offset, uint64_t n_global, sckm_dataset** out) {
    if (!ctx) return SCKM_ERR_INVALID;
    if (!host && n_local) return fail(ctx, SCKM_ERR_INVALID, "host pointer is NULL");
    sckm_dataset* ds = nullptr;
    SCKM_TRY(dataset_alloc(ctx, n_local, d, dtype, row_offset, n_global, &ds));
    const int rc = upload_rows(ds, host, n_local, 0, column_major);
    if (rc != SCKM_OK) { sckm_dataset_destroy(ds); return rc; }
    *out = ds;
    return SCKM_OK;
}

int sckm_dataset_generate_blobs(sckm_ctx* ctx, uint64_t n_local, uint64_t d, uint64_t n_centers,
                                uint64_t seed, int dtype, uint64_t row_offset, uint64_t n_global,
                                sckm_dataset** out) {
    if (!ctx) return SCKM_ERR_INVALID;
    if (n_centers == 0) return fail(ctx, SCKM_ERR_INVALID, "n_centers must be >= 1");
    sckm_dataset* ds = nullptr;
    SCKM_TRY(dataset_alloc(ctx, n_local, d, dtype, row_offset, n_global, &ds));
    int rc = n_local ? launch_blobs(ctx, ds->x, dtype, row_offset, n_local, d, n_centers, seed) : SCKM_OK;
    if (rc == SCKM_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess) rc = fail(ctx, SCKM_ERR_CUDA, "blob generation failed");
    if (rc != SCKM_OK) { sckm_dataset_destroy(ds); return rc; }
    *out = ds;
    return SCKM_OK;
}

int sckm_blobs_fill_host(void* out, int dtype, uint64_t row0, uint64_t nrows, uint64_t d,
                         uint64_t n_centers, uint64_t seed) {
    if (!out || n_centers == 0 || (dtype != SCKM_F32 && dtype != SCKM_F64)) return SCKM_ERR_INVALID;
    for (uint64_t r = 0; r < nrows; r++)
        for (uint64_t c = 0; c < d; c++) {
            double v = blob_value(seed, n_centers, row0 + r, c);
            if (dtype == SCKM_F32) ((float*)out)[r * d + c] = (float)v; else ((double*)out)[r * d + c] = v;
        }
    return SCKM_OK;
}

int sckm_dataset_download_rows(sckm_dataset* ds, uint64_t local_row0, uint64_t nrows, void* host_out) {
    if (!ds || !host_out) return SCKM_ERR_INVALID;
    sckm_ctx* ctx = ds->ctx;
    if (local_row0 + nrows > ds->n) return fail(ctx, SCKM_ERR_INVALID, "row range out of bounds");
    SCKM_CUDA(ctx, cudaSetDevice(ctx->device));
    SCKM_CUDA(ctx, cudaMemcpyAsync(host_out, (const char*)ds->x + local_row0 * ds->d * ds->elem(),
                                   nrows * ds->d * ds->elem(), cudaMemcpyDeviceToHost, ctx->stream));
    SCKM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SCKM_OK;
}

void sckm_dataset_destroy(sckm_dataset* ds) {
    if (!ds) return;
    if (ds->ctx) { cudaSetDevice(ds->ctx->device); cudaStreamSynchronize(ds->ctx->stream); }
    if (ds->ctx) {
        sckm_ctx* c = ds->ctx;
        dev_free(c, ds->x); dev_free(c, ds->labels); dev_free(c, ds->mind); dev_free(c, ds->labels64); dev_free(c, ds->x32);
        dev_free(c, ds->kpp_shadow); dev_free(c, ds->kpp_shadow_err);
        cudaStreamSynchronize(c->stream);
    }
    delete ds;
}

// ---- kmeans++ -----------------------------------------------------------------------------
static void kpp_shadow_free(sckm_dataset* ds) {
    dev_free(ds->ctx, ds->kpp_shadow);
    dev_free(ds->ctx, ds->kpp_shadow_err);
    ds->kpp_shadow = nullptr; ds->kpp_shadow_err = nullptr;
}

static int ensure_kpp(sckm_dataset* ds, uint64_t k) {
    sckm_ctx* ctx = ds->ctx;
    SCKM_TRY(ensure_workspace(ctx, k, ds->d, 0));
    const size_t nb = (ds->n + kKppBlockRows - 1) / kKppBlockRows + 1;
    if (nb > ctx->cap_blocks) {
        if (ctx->d_blocksum) { SCKM_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); cudaFree(ctx->d_blocksum); ctx->d_blocksum = nullptr; }
        SCKM_CUDA(ctx, cudaMalloc((void**)&ctx->d_blocksum, nb * sizeof(double)));
        ctx->cap_blocks = nb;
    }
    if (ds->n > ctx->cap_surv && ds->n < 0xFFFFFFFFull) {
        if (ctx->d_surv) { SCKM_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); cudaFree(ctx->d_surv); ctx->d_surv = nullptr; ctx->cap_surv = 0; }
        SCKM_CUDA(ctx, ws_malloc(ctx, (void**)&ctx->d_surv, ds->n * sizeof(uint32_t)));
        ctx->cap_surv = ds->n;
    }
    if (!ctx->d_kppctr) {
        SCKM_CUDA(ctx, cudaMalloc((void**)&ctx->d_kppctr, 2 * sizeof(unsigned)));
        SCKM_CUDA(ctx, cudaMemsetAsync(ctx->d_kppctr, 0, 2 * sizeof(unsigned), ctx->stream));
    }
    if (ds->d > ctx->cap_tshift) {
        if (ctx->d_tshift) { SCKM_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); cudaFree(ctx->d_tshift); ctx->d_tshift = nullptr; }
        SCKM_CUDA(ctx, cudaMalloc((void**)&ctx->d_tshift, ds->d * sizeof(float)));
        ctx->cap_tshift = ds->d;
    }
    if (!ctx->d_tshift_err) SCKM_CUDA(ctx, cudaMalloc((void**)&ctx->d_tshift_err, sizeof(double)));
    const size_t tab_bytes = (size_t)k * ds->d * ds->elem();
    if (tab_bytes > ctx->cap_seedtab) {
        if (ctx->d_seedtab) { SCKM_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); cudaFree(ctx->d_seedtab); ctx->d_seedtab = nullptr; }
        SCKM_CUDA(ctx, cudaMalloc(&ctx->d_seedtab, tab_bytes));
        ctx->cap_seedtab = tab_bytes;
    }
    if (k > ctx->cap_skiptab) {
        if (ctx->d_skiptab) { SCKM_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); cudaFree(ctx->d_skiptab); ctx->d_skiptab = nullptr; }
        SCKM_CUDA(ctx, cudaMalloc((void**)&ctx->d_skiptab, k * sizeof(double)));
        ctx->cap_skiptab = k;
    }
    const size_t seed_bytes = ((size_t)ds->d * ds->elem() + 7) / 8 * 8 + 8;
    if (seed_bytes > ctx->cap_seedrow) {
        if (ctx->d_seedrow) { SCKM_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); cudaFree(ctx->d_seedrow); ctx->d_seedrow = nullptr; }
        SCKM_CUDA(ctx, cudaMalloc(&ctx->d_seedrow, seed_bytes));
        ctx->cap_seedrow = seed_bytes;
    } else if (seed_bytes < ctx->cap_seedrow) {
        // the select kernel derives the index slot from the buffer size: keep it exact
        SCKM_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); cudaFree(ctx->d_seedrow); ctx->d_seedrow = nullptr;
        SCKM_CUDA(ctx, cudaMalloc(&ctx->d_seedrow, seed_bytes));
        ctx->cap_seedrow = seed_bytes;
    }
    return SCKM_OK;
}

int sckm_kmeanspp(sckm_dataset* ds, uint64_t k, uint64_t first_index, const double* uniforms,
                  const int64_t* inject_rows, int64_t* seed_rows_out) {
    if (!ds) return SCKM_ERR_INVALID;
    sckm_ctx* ctx = ds->ctx;
    if (k < 1) return fail(ctx, SCKM_ERR_INVALID, "k must be >= 1");
    if (ds->n_global == 0) return fail(ctx, SCKM_ERR_INVALID, "empty dataset");
    if (!inject_rows && k > 1 && !uniforms) return fail(ctx, SCKM_ERR_INVALID, "uniforms is NULL");
    if (!inject_rows && first_index >= ds->n_global) return fail(ctx, SCKM_ERR_INVALID, "first_index out of range");
    if (inject_rows)
        for (uint64_t j = 0; j < k; j++)
            if (inject_rows[j] < 0 || (uint64_t)inject_rows[j] >= ds->n_global)
                return fail(ctx, SCKM_ERR_INVALID, "inject_rows[%llu] out of range", (unsigned long long)j);
    SCKM_CUDA(ctx, cudaSetDevice(ctx->device));
    const bool trace = getenv("SCKM_TRACE") != nullptr;     // host-side phase times on stderr (diagnostics)
    auto now = []() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t0 = now();
    SCKM_TRY(ensure_kpp(ds, k));
    const double t1 = now();
    const size_t seed_words = ctx->cap_seedrow / 8;
    // seed 0: the row drawn by gen_range (kmeans.rs:358-362)
    SCKM_TRY(launch_kpp_select(ds, 0.0, inject_rows ? inject_rows[0] : (int64_t)first_index, 0));
    SCKM_TRY(nccl_allreduce_u64(ctx, (unsigned long long*)ctx->d_seedrow, seed_words));
    // Triangle-inequality pruning: a row whose D^2 to its nearest seed s is <= ||s_new - s||^2 / 4 cannot improve, so
    // it is neither read nor touched.  Exact (the skipped rows are exactly rows the reference would leave unchanged).
    const bool prune = getenv("SCKM_KPP_NOPRUNE") == nullptr && (size_t)k * sizeof(double) <= 64 * 1024;
    // bf16 shadow for the screening test of the pruned passes (16-byte rows of 8 bf16, enough passes to repay the
    // build, memory permitting): see kpp_prune_kernel.  Absent shadow = exact path only, same results.
    kpp_shadow_free(ds);
    if (prune && k >= 8 && ds->d % 8 == 0 && ds->n && ds->n < 0xFFFFFFFFull && !getenv("SCKM_KPP_NOSHADOW")) {
        if (dev_alloc(ctx, (void**)&ds->kpp_shadow, ds->n * ds->d * sizeof(uint16_t)) != cudaSuccess ||
            dev_alloc(ctx, (void**)&ds->kpp_shadow_err, ds->n * sizeof(float)) != cudaSuccess) {
            kpp_shadow_free(ds);
        }
    }
    const double t2 = now();
    SCKM_TRY(launch_kpp_seedtab(ds, 0));
    for (uint64_t j = 1; j < k; j++) {
        SCKM_TRY(launch_kpp_refresh(ds, (uint32_t)(j - 1), j == 1, prune));
        if (ctx->nranks > 1) SCKM_TRY(nccl_allgather_f64(ctx, ctx->d_totals + ctx->rank, ctx->d_totals));
        SCKM_TRY(launch_kpp_select(ds, inject_rows ? 0.0 : uniforms[j - 1], inject_rows ? inject_rows[j] : -1, (uint32_t)j));
        SCKM_TRY(nccl_allreduce_u64(ctx, (unsigned long long*)ctx->d_seedrow, seed_words));
        SCKM_TRY(launch_kpp_seedtab(ds, (uint32_t)j));
    }
    SCKM_TRY(launch_kpp_refresh(ds, (uint32_t)(k - 1), k == 1, prune, /*want_sums=*/false));  // final pass, label k-1 (kmeans.rs:399-410)
    if (seed_rows_out) {
        if (ctx->nranks > 1) SCKM_TRY(nccl_allreduce_u64(ctx, (unsigned long long*)ctx->d_seeds, k));
        SCKM_CUDA(ctx, cudaMemcpyAsync(seed_rows_out, ctx->d_seeds, k * sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
    }
    const double t3 = now();
    SCKM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    const double t4 = now();
    kpp_shadow_free(ds);
    if (trace)
        fprintf(stderr, "[sckm] kmeans++ k=%llu: workspaces %.2f ms, seed0+shadow alloc %.2f ms, enqueue %.2f ms, drain %.2f ms, shadow free %.2f ms\n",
                (unsigned long long)k, t1 - t0, t2 - t1, t3 - t2, t4 - t3, now() - t4);
    ds->have_labels = true;
    return SCKM_OK;
}

// ---- shared pieces of the Lloyd driver -------------------------------------------------------
static int pick_assign(const sckm_dataset* ds, uint64_t k) {
    const sckm_ctx* ctx = ds->ctx;
    int which = ctx->assign_kernel;
    if (which == SCKM_ASSIGN_AUTO)
        which = tc5_auto(ds, k) ? SCKM_ASSIGN_TC5 : stream_supported(ds, k) ? SCKM_ASSIGN_STREAM
                : dmma_supported(ds, k) ? SCKM_ASSIGN_DMMA : SCKM_ASSIGN_DIRECT;
    if (which == SCKM_ASSIGN_TC5 && !tc5_supported(ds, k)) which = dmma_supported(ds, k) ? SCKM_ASSIGN_DMMA : SCKM_ASSIGN_DIRECT;
    if (which == SCKM_ASSIGN_DMMA && !dmma_supported(ds, k)) which = SCKM_ASSIGN_DIRECT;
    if (which == SCKM_ASSIGN_STREAM && !stream_supported(ds, k)) which = SCKM_ASSIGN_DIRECT;
    return which;
}

// one clustering step on the device: labels, packed = all-reduced [sums | counts | inertia]
static int clustering_step(sckm_dataset* ds, uint64_t k, cudaEvent_t ev_a0 = nullptr, cudaEvent_t ev_a1 = nullptr) {
    sckm_ctx* ctx = ds->ctx;
    const int which = pick_assign(ds, k);
    if (ev_a0) SCKM_CUDA(ctx, cudaEventRecord(ev_a0, ctx->stream));
    if (which == SCKM_ASSIGN_DMMA || which == SCKM_ASSIGN_STREAM || which == SCKM_ASSIGN_TC5) {
        if (which == SCKM_ASSIGN_DMMA) SCKM_TRY(launch_assign_dmma(ds, k));   // assignment + fused partial sums
        else if (which == SCKM_ASSIGN_TC5) SCKM_TRY(launch_assign_tc5(ds, k));
        else SCKM_TRY(launch_assign_stream(ds, k));
        if (ev_a1) SCKM_CUDA(ctx, cudaEventRecord(ev_a1, ctx->stream));
        SCKM_TRY(launch_reduce_partials(ctx, ctx->partial_slots_used, (size_t)k * ds->d + k + 1));
    } else {
        ctx->packed_centered = false;
        SCKM_TRY(launch_assign_direct(ds, k));
        if (ev_a1) SCKM_CUDA(ctx, cudaEventRecord(ev_a1, ctx->stream));
        SCKM_TRY(launch_update(ds, k, true));
    }
    // multi-GPU: the sum over the ranks -- inside the finalize launch that follows when the loop runs the peer exchange
    if (!ctx->peer_step) SCKM_TRY(nccl_allreduce_f64(ctx, ctx->d_packed, (size_t)k * ds->d + k + 1));
    ds->have_labels = true;
    return SCKM_OK;
}

static int check_k(sckm_dataset* ds, uint64_t k) {
    if (k < 1 || k > 0x7FFFFFFFull) return fail(ds->ctx, SCKM_ERR_INVALID, "k=%llu out of range", (unsigned long long)k);
    return SCKM_OK;
}

int sckm_init_centroids(sckm_dataset* ds, uint64_t k, double* centroids_out, int64_t* size_out) {
    if (!ds) return SCKM_ERR_INVALID;
    sckm_ctx* ctx = ds->ctx;
    SCKM_TRY(check_k(ds, k));
    if (!ds->have_labels) return fail(ctx, SCKM_ERR_STATE, "no labels on the dataset: run sckm_kmeanspp first");
    SCKM_CUDA(ctx, cudaSetDevice(ctx->device));
    SCKM_TRY(launch_update(ds, k, false));
    SCKM_TRY(nccl_allreduce_f64(ctx, ctx->d_packed, (size_t)k * ds->d + k + 1));
    SCKM_TRY(launch_finalize(ctx, k, ds->d, /*guarded=*/false));
    ctx->cnorm_valid = false;     // the first step of the fit picks its centring shift from these centroids (launch_cnorm)
    if (centroids_out) SCKM_CUDA(ctx, cudaMemcpyAsync(centroids_out, ctx->d_centroids, k * ds->d * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (size_out) SCKM_CUDA(ctx, cudaMemcpyAsync(size_out, ctx->d_size, k * sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
    SCKM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SCKM_OK;
}

int sckm_lloyd_step(sckm_dataset* ds, const double* centroids, uint64_t k, double* sums_out,
                    int64_t* counts_out, double* inertia_out) {
    if (!ds || !centroids) return SCKM_ERR_INVALID;
    sckm_ctx* ctx = ds->ctx;
    SCKM_TRY(check_k(ds, k));
    SCKM_CUDA(ctx, cudaSetDevice(ctx->device));
    SCKM_TRY(ensure_workspace(ctx, k, ds->d, 0));
    const size_t kd = (size_t)k * ds->d;
    SCKM_CUDA(ctx, cudaMemcpyAsync(ctx->d_centroids, centroids, kd * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    ctx->cnorm_valid = false;
    SCKM_TRY(clustering_step(ds, k));
    std::vector<double> packed(kd + k + 1), mu;
    SCKM_CUDA(ctx, cudaMemcpyAsync(packed.data(), ctx->d_packed, packed.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (ctx->packed_centered && sums_out) {
        mu.resize(ds->d);
        SCKM_CUDA(ctx, cudaMemcpyAsync(mu.data(), ctx->d_mu, ds->d * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    }
    SCKM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (sums_out) memcpy(sums_out, packed.data(), kd * sizeof(double));
    if (!mu.empty())                                   // the tile kernel summed x - mu: sum(x) = sum(x - mu) + count * mu
        for (uint64_t c = 0; c < k; c++)
            for (uint64_t j = 0; j < ds->d; j++) sums_out[c * ds->d + j] += packed[kd + c] * mu[j];
    if (counts_out) for (uint64_t c = 0; c < k; c++) counts_out[c] = (int64_t)packed[kd + c];
    if (inertia_out) *inertia_out = packed[kd + k];
    return SCKM_OK;
}

// The loop of KMeans::fit (kmeans.rs:294-310).  The stop rule lives on the device (finalize_kernel + LoopState), so
// the host enqueues BATCHES of iterations and reads the 32-byte state back once per batch: kernels of iterations past
// the one that broke the loop return at once, which leaves labels / centroids / sizes exactly as the reference's
// `break` does.  The batch size adapts to the measured iteration time: a long iteration (config C3: 10 ms) is followed
// by a read-back every time (nothing speculative is ever enqueued), a short one (config C2: ~50 us) only every few
// iterations, so the host round trip (~10-15 us) stops being a quarter of the step.  The batch size is a pure function
// of the shape (never of a clock): every rank of a multi-GPU fit must enqueue the same sequence of collectives.
// honor_stop = false (bench): all iterations in one batch, no read-back in between.
}  // extern "C"
namespace sckm {
int lloyd_loop(sckm_dataset* ds, uint64_t k, uint64_t max_iter, bool honor_stop, double* centroids_inout,
               int64_t* size_out, double* distortion_out, int64_t* iters_out, double* inertia_trace,
               float* ms_trace, float* assign_ms_trace) {
    sckm_ctx* ctx = ds->ctx;
    SCKM_TRY(check_k(ds, k));
    if (max_iter == 0) return fail(ctx, SCKM_ERR_INVALID, "max_iter must be >= 1");
    if (max_iter > 0xFFFFFFF0ull) return fail(ctx, SCKM_ERR_INVALID, "max_iter too large");
    SCKM_CUDA(ctx, cudaSetDevice(ctx->device));
    SCKM_TRY(ensure_workspace(ctx, k, ds->d, 0));
    const size_t kd = (size_t)k * ds->d;
    if (centroids_inout) {
        SCKM_CUDA(ctx, cudaMemcpyAsync(ctx->d_centroids, centroids_inout, kd * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        ctx->cnorm_valid = false;
    }
    bool peers = false;
    SCKM_TRY(peer_prepare(ctx, kd + k + 1, &peers));  // multi-GPU: map the ranks' exchange buffers (first loop / larger k*d)
    ctx->allreduce_path = ctx->nranks <= 1 ? SCKM_ALLREDUCE_NONE : peers ? SCKM_ALLREDUCE_PEER : SCKM_ALLREDUCE_NCCL;
    SCKM_TRY(launch_loop_init(ctx, max_iter, honor_stop));
    std::vector<cudaEvent_t> evs, evs_a;
    if (assign_ms_trace) {
        evs_a.resize(2 * max_iter);
        for (auto& e : evs_a) SCKM_CUDA(ctx, cudaEventCreate(&e));
    }
    // ms_trace without assign_ms_trace: the loop is timed as a whole -- two events, none between the steps (an event record
    // is a stream operation of its own: three per step cost ~8 us of a ~50 us step at config C2) -- and every slot of
    // ms_trace carries the mean
    const bool whole = ms_trace && !assign_ms_trace;
    if (ms_trace) {
        evs.resize(whole ? 2 : max_iter + 1);
        for (auto& e : evs) SCKM_CUDA(ctx, cudaEventCreate(&e));
        SCKM_CUDA(ctx, cudaEventRecord(evs[0], ctx->stream));
    }
    LoopState* h_state = reinterpret_cast<LoopState*>(ctx->h_pinned);
    uint64_t batch = max_iter;
    if (honor_stop) {
        // ~400 us of estimated device work between read-backs, at most 8 iterations enqueued ahead of the stop test
        const double n_loc = (double)((ds->n_global + ctx->nranks - 1) / std::max(ctx->nranks, 1));
        const double us_hbm = n_loc * ((double)ds->d * ds->elem() + 4.0) / 5e6;                  // ~5 TB/s
        const double us_fp = 2.0 * n_loc * (double)k * (double)ds->d / (ds->dtype == SCKM_F32 ? 150e6 : 25e6);   // TFLOP/s -> flop/us
        batch = (uint64_t)std::max(1.0, std::min(8.0, 400.0 / (std::max(us_hbm, us_fp) + 15.0)));
    }
    if (const char* e = getenv("SCKM_LLOYD_BATCH")) batch = std::max<uint64_t>(1, strtoull(e, nullptr, 10));   // tests / tuning
    int rc = SCKM_OK;
    uint64_t it = 0;
    h_state->done_at = 0; h_state->iters = 0; h_state->distortion = DBL_MAX;
    while (it < max_iter && rc == SCKM_OK) {
        const uint64_t nb = std::min(batch, max_iter - it);
        for (uint64_t b = 0; b < nb && rc == SCKM_OK; b++) {
            it++;
            ctx->loop_it = (uint32_t)it;
            if (peers) { peer_next(ctx); ctx->peer_step = true; }
            if (assign_ms_trace) rc = clustering_step(ds, k, evs_a[2 * (it - 1)], evs_a[2 * (it - 1) + 1]);
            else rc = clustering_step(ds, k);                                      // bbd.clustering(...)        kmeans.rs:296
            if (rc == SCKM_OK) rc = launch_finalize(ctx, k, ds->d, /*guarded=*/true);  // centroids = sums / size + stop rule  kmeans.rs:297-309
            ctx->peer_step = false;
            if (rc == SCKM_OK && ms_trace && !whole && cudaEventRecord(evs[it], ctx->stream) != cudaSuccess) rc = fail(ctx, SCKM_ERR_CUDA, "cudaEventRecord failed");
        }
        ctx->loop_it = 0;
        if (rc != SCKM_OK) break;
        if (whole && it >= max_iter && cudaEventRecord(evs[1], ctx->stream) != cudaSuccess) { rc = fail(ctx, SCKM_ERR_CUDA, "cudaEventRecord failed"); break; }
        if (cudaMemcpyAsync(h_state, ctx->d_loop, sizeof(LoopState), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
            cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
            rc = fail(ctx, SCKM_ERR_CUDA, "Lloyd loop failed: %s", cudaGetErrorString(cudaGetLastError()));
            break;
        }
        if (h_state->done_at) break;
    }
    ctx->loop_it = 0;
    const int64_t iters = (int64_t)h_state->iters;
    const double distortion = h_state->distortion;
    if (rc == SCKM_OK) {
        if (centroids_inout && cudaMemcpyAsync(centroids_inout, ctx->d_centroids, kd * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess) rc = SCKM_ERR_CUDA;
        if (size_out && cudaMemcpyAsync(size_out, ctx->d_size, k * sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess) rc = SCKM_ERR_CUDA;
        if (inertia_trace && iters > 0 &&
            cudaMemcpyAsync(inertia_trace, ctx->d_inertia_trace, (size_t)iters * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess) rc = SCKM_ERR_CUDA;
        if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) rc = SCKM_ERR_CUDA;
        if (rc != SCKM_OK) rc = fail(ctx, SCKM_ERR_CUDA, "Lloyd loop download failed: %s", cudaGetErrorString(cudaGetLastError()));
        if (rc == SCKM_OK && peers) rc = peer_check(ctx);
    } else {
        cudaStreamSynchronize(ctx->stream);
    }
    if (rc == SCKM_OK && ms_trace && !whole)
        for (int64_t i = 0; i < iters; i++) cudaEventElapsedTime(&ms_trace[i], evs[i], evs[i + 1]);
    if (rc == SCKM_OK && whole && iters > 0) {
        float total = 0.f;
        cudaEventElapsedTime(&total, evs[0], evs[1]);
        for (int64_t i = 0; i < iters; i++) ms_trace[i] = total / (float)iters;
    }
    if (rc == SCKM_OK && assign_ms_trace)
        for (int64_t i = 0; i < iters; i++) cudaEventElapsedTime(&assign_ms_trace[i], evs_a[2 * i], evs_a[2 * i + 1]);
    for (auto& e : evs) cudaEventDestroy(e);
    for (auto& e : evs_a) cudaEventDestroy(e);
    if (rc != SCKM_OK) return rc;
    if (distortion_out) *distortion_out = distortion;
    if (iters_out) *iters_out = iters;
    return SCKM_OK;
}
}  // namespace sckm
extern "C" {

int sckm_lloyd_fit(sckm_dataset* ds, uint64_t k, uint64_t max_iter, double* centroids_inout,
                   int64_t* size_out, double* distortion_out, int64_t* iters_out) {
    if (!ds || !centroids_inout) return SCKM_ERR_INVALID;
    return lloyd_loop(ds, k, max_iter, true, centroids_inout, size_out, distortion_out, iters_out, nullptr, nullptr, nullptr);
}

int sckm_lloyd_iterate(sckm_dataset* ds, uint64_t k, uint64_t n_iters, double* centroids_inout,
                       int64_t* size_out, double* inertia_out, float* ms_per_iter_out, float* assign_ms_out) {
    if (!ds || !centroids_inout) return SCKM_ERR_INVALID;
    return lloyd_loop(ds, k, n_iters, false, centroids_inout, size_out, nullptr, nullptr, inertia_out, ms_per_iter_out,
                      assign_ms_out);
}

}  // extern "C"
namespace sckm {
int download_labels(sckm_dataset* ds, void* out, int width) {
    sckm_ctx* ctx = ds->ctx;
    const uint64_t n = ds->n;
    if (width == 4) {
        return copy_to_host(ctx, out, ds->labels, n * 4);
    }
    if (width != 8) return fail(ctx, SCKM_ERR_INVALID, "label width must be 4 or 8");
    if (!ds->labels64) SCKM_CUDA(ctx, dev_alloc(ctx, (void**)&ds->labels64, std::max<uint64_t>(n, 1) * 8));
    SCKM_TRY(launch_labels_widen(ctx, ds->labels, ds->labels64, n));
    return copy_to_host(ctx, out, ds->labels64, n * 8);
}
}  // namespace sckm
extern "C" {

int sckm_labels_download(sckm_dataset* ds, void* out, int width) {
    if (!ds || !out) return SCKM_ERR_INVALID;
    if (!ds->have_labels) return fail(ds->ctx, SCKM_ERR_STATE, "no labels on the dataset yet");
    SCKM_CUDA(ds->ctx, cudaSetDevice(ds->ctx->device));
    return download_labels(ds, out, width);
}

int sckm_mindist_download(sckm_dataset* ds, double* out) {
    if (!ds || !out) return SCKM_ERR_INVALID;
    sckm_ctx* ctx = ds->ctx;
    SCKM_CUDA(ctx, cudaSetDevice(ctx->device));
    return copy_to_host(ctx, out, ds->mind, ds->n * sizeof(double));
}

// ---- predict --------------------------------------------------------------------------------
// KMeans::predict (kmeans.rs:327-352) for any n: X streams through two device buffers of <= ~256 MB, the upload of
// chunk i+1 (threaded pinned ring, sckm_ingest.cu) overlapping the kernels of chunk i, so device memory stays
// bounded and a prediction set need not fit in HBM.  Labels are those of the direct form for every row.
static int predict_chunk(sckm_dataset* ds, uint64_t k) {
    sckm_ctx* ctx = ds->ctx;
    // large k: rank on the DMMA tiles and re-decide near-ties exactly (same labels as the direct form, much faster);
    // otherwise the direct-form kernel
    if (dmma_supported(ds, k) && k >= 32 && ctx->assign_kernel != SCKM_ASSIGN_DIRECT) return launch_predict_dmma(ds, k);
    return launch_assign_direct_raw(ctx, ds->x, ds->dtype, ds->n, ds->d, k, ds->labels, nullptr);
}

}  // extern "C"
namespace sckm {
// rows [lo, lo + n) of a host matrix with `host_rows` rows (the pitch of a column-major image) -> labels_out[0..n)
int predict_rows(sckm_ctx* ctx, const void* x_host, uint64_t host_rows, uint64_t lo, uint64_t n, uint64_t d, int dtype,
                 int column_major, const double* centroids, uint64_t k, void* labels_out, int width) {
    if (n == 0) return SCKM_OK;
    SCKM_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t elem = dtype == SCKM_F32 ? 4 : 8, row_bytes = (size_t)d * elem;
    uint64_t chunk_rows = std::max<uint64_t>(4096, ((size_t)256 << 20) / std::max<size_t>(row_bytes, 1));
    if (const char* e = getenv("SCKM_PREDICT_CHUNK_ROWS")) chunk_rows = std::max<uint64_t>(1, strtoull(e, nullptr, 10));   // tests
    chunk_rows = std::min(chunk_rows, n);
    const uint64_t nchunks = (n + chunk_rows - 1) / chunk_rows;

    sckm_dataset* buf[2] = {nullptr, nullptr};
    void* cm_tmp = nullptr;                                   // column-major image of one chunk
    int rc = dataset_alloc(ctx, chunk_rows, d, dtype, 0, chunk_rows, &buf[0]);
    if (rc == SCKM_OK && nchunks > 1) rc = dataset_alloc(ctx, chunk_rows, d, dtype, 0, chunk_rows, &buf[1]);
    if (rc == SCKM_OK && column_major && dev_alloc(ctx, &cm_tmp, chunk_rows * row_bytes) != cudaSuccess) {
        rc = fail(ctx, SCKM_ERR_CUDA, "cudaMalloc of the column-major staging image failed");
    }
    if (rc == SCKM_OK) rc = ensure_workspace(ctx, k, d, 0);
    if (rc == SCKM_OK && cudaMemcpyAsync(ctx->d_centroids, centroids, k * d * sizeof(double), cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess)
        rc = fail(ctx, SCKM_ERR_CUDA, "centroid upload failed");
    ctx->cnorm_valid = false;

    // chunk c covers rows [lo + c*chunk_rows, ...): row-major = one contiguous block; column-major = d strided pieces,
    // landed as a [d][rows] image (2-D copy, stream-ordered) and transposed on the device
    auto upload = [&](uint64_t c, bool first) -> int {
        sckm_dataset* ds = buf[c & 1];
        const uint64_t r0 = lo + c * chunk_rows, rows = std::min(chunk_rows, lo + n - r0);
        ds->n = rows;
        if (!column_major) return copy_to_device(ctx, ds->x, (const char*)x_host + r0 * row_bytes, rows * row_bytes, first);
        SCKM_CUDA(ctx, cudaMemcpy2DAsync(cm_tmp, rows * elem, (const char*)x_host + r0 * elem, host_rows * elem, rows * elem, d,
                                         cudaMemcpyHostToDevice, ctx->stream));
        return launch_transpose(ctx, cm_tmp, ds->x, rows, d, dtype);
    };
    ctx->ingest_hint = (size_t)n * row_bytes;                 // the chunks belong to one transfer of this size
    if (rc == SCKM_OK) rc = upload(0, true);
    for (uint64_t c = 0; c < nchunks && rc == SCKM_OK; c++) {
        sckm_dataset* ds = buf[c & 1];
        rc = predict_chunk(ds, k);                                            // asynchronous on ctx->stream
        if (rc == SCKM_OK && c + 1 < nchunks) rc = upload(c + 1, false);      // other buffer: idle since its labels came back
        if (rc == SCKM_OK) rc = download_labels(ds, (char*)labels_out + c * chunk_rows * (size_t)width, width);
    }
    ctx->ingest_hint = 0;
    dev_free(ctx, cm_tmp);
    cudaStreamSynchronize(ctx->stream);
    sckm_dataset_destroy(buf[0]);
    sckm_dataset_destroy(buf[1]);
    return rc;
}
}  // namespace sckm
extern "C" {

int sckm_predict(sckm_ctx* ctx, const void* x_host, uint64_t n, uint64_t d, int dtype,
                 int column_major, const double* centroids, uint64_t k, void* labels_out, int width) {
    if (!ctx) return SCKM_ERR_INVALID;
    if ((!x_host || !labels_out) && n) return fail(ctx, SCKM_ERR_INVALID, "NULL buffer");
    if (!centroids || k < 1) return fail(ctx, SCKM_ERR_INVALID, "no centroids");
    if (width != 4 && width != 8) return fail(ctx, SCKM_ERR_INVALID, "label width must be 4 or 8");
    if (dtype != SCKM_F32 && dtype != SCKM_F64) return fail(ctx, SCKM_ERR_INVALID, "dtype must be SCKM_F32 or SCKM_F64");
    if (d == 0 || d > (1u << 20)) return fail(ctx, SCKM_ERR_INVALID, "d=%llu out of range", (unsigned long long)d);
    if (multi_shards(ctx, n, nullptr))                                        // rows sharded over the context's devices
        return multi_predict(ctx, x_host, n, d, dtype, column_major, centroids, k, labels_out, width);
    return predict_rows(ctx, x_host, n, 0, n, d, dtype, column_major, centroids, k, labels_out, width);
}

// ---- cluster quality: contingency table --------------------------------------------------------
static int contingency_common(sckm_ctx* ctx, const uint32_t* a_host, const uint32_t* d_b, const uint32_t* b_host, uint64_t n,
                              uint64_t na, uint64_t nb, bool allreduce, int64_t* out) {
    if (!out || na == 0 || nb == 0) return fail(ctx, SCKM_ERR_INVALID, "empty contingency table");
    if (na > (1u << 20) || nb > (1u << 20) || na * nb > (1ull << 26)) return fail(ctx, SCKM_ERR_INVALID, "contingency table too large");
    if (n && (!a_host || (!d_b && !b_host))) return fail(ctx, SCKM_ERR_INVALID, "NULL id array");
    SCKM_CUDA(ctx, cudaSetDevice(ctx->device));
    const uint64_t ncell = na * nb;
    uint32_t *d_a = nullptr, *d_b2 = nullptr;
    unsigned long long* d_out = nullptr;
    int rc = SCKM_OK;
    auto cleanup = [&]() { dev_free(ctx, d_a); dev_free(ctx, d_b2); dev_free(ctx, d_out); };
    if (dev_alloc(ctx, (void**)&d_a, std::max<uint64_t>(n, 1) * 4) != cudaSuccess ||
        dev_alloc(ctx, (void**)&d_out, (ncell + 1) * 8) != cudaSuccess ||
        (!d_b && dev_alloc(ctx, (void**)&d_b2, std::max<uint64_t>(n, 1) * 4) != cudaSuccess)) {
        cleanup();
        return fail(ctx, SCKM_ERR_CUDA, "cudaMalloc for the contingency table failed");
    }
    rc = copy_to_device(ctx, d_a, a_host, n * 4);
    if (rc == SCKM_OK && !d_b) rc = copy_to_device(ctx, d_b2, b_host, n * 4);
    if (rc == SCKM_OK) rc = launch_contingency(ctx, d_a, d_b ? d_b : d_b2, n, na, nb, d_out);
    if (rc == SCKM_OK && allreduce) rc = nccl_allreduce_u64(ctx, d_out, ncell + 1);
    std::vector<unsigned long long> h(ncell + 1);
    if (rc == SCKM_OK && (cudaMemcpyAsync(h.data(), d_out, (ncell + 1) * 8, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
                          cudaStreamSynchronize(ctx->stream) != cudaSuccess))
        rc = fail(ctx, SCKM_ERR_CUDA, "contingency download failed: %s", cudaGetErrorString(cudaGetLastError()));
    cleanup();
    if (rc != SCKM_OK) return rc;
    if (h[ncell]) return fail(ctx, SCKM_ERR_INVALID, "%llu rows carry an id outside [0,%llu) x [0,%llu)", h[ncell],
                              (unsigned long long)na, (unsigned long long)nb);
    for (uint64_t i = 0; i < ncell; i++) out[i] = (int64_t)h[i];
    return SCKM_OK;
}

int sckm_contingency(sckm_dataset* ds, const uint32_t* class_ids_host, uint64_t n_classes, uint64_t k, int64_t* out) {
    if (!ds) return SCKM_ERR_INVALID;
    if (!ds->have_labels) return fail(ds->ctx, SCKM_ERR_STATE, "no labels on the dataset yet");
    return contingency_common(ds->ctx, class_ids_host, ds->labels, nullptr, ds->n, n_classes, k, true, out);
}

int sckm_contingency_host(sckm_ctx* ctx, const uint32_t* a_host, const uint32_t* b_host, uint64_t n, uint64_t na,
                          uint64_t nb, int64_t* out) {
    if (!ctx) return SCKM_ERR_INVALID;
    return contingency_common(ctx, a_host, nullptr, b_host, n, na, nb, false, out);
}

// ---- batched k-nearest neighbours -----------------------------------------------------------------
int sckm_knn(sckm_dataset* ds, const void* queries_host, uint64_t nq, uint64_t k, int64_t* idx_out, double* dist_out) {
    if (!ds) return SCKM_ERR_INVALID;
    if (nq && (!queries_host || !idx_out || !dist_out)) return fail(ds->ctx, SCKM_ERR_INVALID, "NULL buffer");
    return knn_search(ds, queries_host, nq, k, idx_out, dist_out);
}

int sckm_radius_count(sckm_dataset* ds, const void* queries_host, uint64_t nq, double radius, int64_t* counts_out) {
    if (!ds) return SCKM_ERR_INVALID;
    if (nq && (!queries_host || !counts_out)) return fail(ds->ctx, SCKM_ERR_INVALID, "NULL buffer");
    return radius_search(ds, queries_host, nq, radius, counts_out, nullptr, 0, nullptr, nullptr);
}

int sckm_radius_fill(sckm_dataset* ds, const void* queries_host, uint64_t nq, double radius, const int64_t* offsets,
                     uint64_t total, int64_t* idx_out, double* dist_out) {
    if (!ds) return SCKM_ERR_INVALID;
    if (nq && (!queries_host || !offsets || (total && (!idx_out || !dist_out)))) return fail(ds->ctx, SCKM_ERR_INVALID, "NULL buffer");
    return radius_search(ds, queries_host, nq, radius, nullptr, offsets, total, idx_out, dist_out);
}

// ---- whole fit from host buffers ----------------------------------------------------------------
}  // extern "C"
namespace sckm {
// KMeans::fit on the rows [lo, lo + n_local) of the host matrix, as rank ctx->rank of ctx->nranks (1 rank: the whole
// matrix).  Phase 1 (allocation + upload) and phase 2 (kmeans++, means, loop, download) are separate calls so that a
// multi-GPU driver can stop every rank before the first collective when one of them could not get its memory.
int fit_upload(sckm_ctx* ctx, const void* x_host, uint64_t host_rows, uint64_t lo, uint64_t n_local, uint64_t d, int dtype,
               int column_major, sckm_dataset** out) {
    SCKM_CUDA(ctx, cudaSetDevice(ctx->device));
    sckm_dataset* ds = nullptr;
    SCKM_TRY(dataset_alloc(ctx, n_local, d, dtype, lo, host_rows, &ds));
    const int rc = upload_rows(ds, x_host, host_rows, lo, column_major);
    if (rc != SCKM_OK) { sckm_dataset_destroy(ds); return rc; }
    *out = ds;
    return SCKM_OK;
}
// Every device allocation the compute phase will need (kmeans++ scratch, centroid workspaces, the partial slots of the
// assignment kernel this shape selects), made up front: in a multi-GPU fit a rank that ran out of memory in the middle
// of the loop would leave the others waiting in a collective, so all allocation failures must surface before the join.
int fit_reserve(sckm_dataset* ds, uint64_t k) {
    sckm_ctx* ctx = ds->ctx;
    SCKM_CUDA(ctx, cudaSetDevice(ctx->device));
    SCKM_TRY(check_k(ds, k));
    SCKM_TRY(ensure_kpp(ds, k));
    size_t slots = (size_t)ctx->num_sms * 16;                     // direct path: update_partial_kernel / update_given_kernel
    switch (pick_assign(ds, k)) {
        case SCKM_ASSIGN_DMMA: slots = dmma_partial_slots(ctx); break;
        case SCKM_ASSIGN_TC5: slots = std::max<size_t>(slots, (size_t)ctx->num_sms * 8); break;
        case SCKM_ASSIGN_STREAM: slots = std::max<size_t>(slots, (size_t)ctx->num_sms * 4 * 8); break;
        default: break;
    }
    slots = std::max<size_t>(slots, dmma_partial_slots(ctx));      // the initial means of a large fit use the tile-layout update
    return ensure_workspace(ctx, k, ds->d, slots);
}
int fit_compute(sckm_dataset* ds, uint64_t k, uint64_t max_iter, uint64_t first_index, const double* uniforms,
                int64_t* size_out, double* centroids_out, double* distortion_out, int64_t* iters_out, double* phase_s) {
    sckm_ctx* ctx = ds->ctx;
    SCKM_CUDA(ctx, cudaSetDevice(ctx->device));
    auto now = []() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t0 = now();
    int rc = sckm_kmeanspp(ds, k, first_index, uniforms, nullptr, nullptr);
    if (rc == SCKM_OK) rc = sckm_init_centroids(ds, k, nullptr, nullptr);
    const double t1 = now();
    if (rc == SCKM_OK) rc = lloyd_loop(ds, k, max_iter, true, nullptr, size_out, distortion_out, iters_out, nullptr, nullptr, nullptr);
    if (rc == SCKM_OK && centroids_out &&
        (cudaMemcpyAsync(centroids_out, ctx->d_centroids, k * ds->d * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
         cudaStreamSynchronize(ctx->stream) != cudaSuccess))
        rc = fail(ctx, SCKM_ERR_CUDA, "centroid download failed");
    if (phase_s) { phase_s[0] = t1 - t0; phase_s[1] = now() - t1; }
    return rc;
}
}  // namespace sckm
extern "C" {

int sckm_kmeans_fit(sckm_ctx* ctx, const void* x_host, uint64_t n, uint64_t d, int dtype,
                    int column_major, uint64_t k, uint64_t max_iter, uint64_t first_index,
                    const double* uniforms, void* labels_out, int width, int64_t* size_out,
                    double* centroids_out, double* distortion_out, int64_t* iters_out) {
    if (!ctx) return SCKM_ERR_INVALID;
    if (!centroids_out) return fail(ctx, SCKM_ERR_INVALID, "centroids_out is NULL");
    if (n == 0) return fail(ctx, SCKM_ERR_INVALID, "empty input");
    if (!x_host) return fail(ctx, SCKM_ERR_INVALID, "host pointer is NULL");
    if (labels_out && width != 4 && width != 8) return fail(ctx, SCKM_ERR_INVALID, "label width must be 4 or 8");
    if (multi_shards(ctx, n, nullptr))                                        // rows sharded over the context's devices
        return multi_kmeans_fit(ctx, x_host, n, d, dtype, column_major, k, max_iter, first_index, uniforms, labels_out,
                                width, size_out, centroids_out, distortion_out, iters_out);
    auto now = []() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t0 = now();
    sckm_dataset* ds = nullptr;
    SCKM_TRY(fit_upload(ctx, x_host, n, 0, n, d, dtype, column_major, &ds));
    const double t1 = now();
    double ph[2] = {0, 0};
    int rc = fit_compute(ds, k, max_iter, first_index, uniforms, size_out, centroids_out, distortion_out, iters_out, ph);
    const double t2 = now();
    if (rc == SCKM_OK && labels_out) rc = download_labels(ds, labels_out, width);
    sckm_dataset_destroy(ds);
    const double t3 = now();
    ctx->fit_times[0] = t1 - t0; ctx->fit_times[1] = ph[0]; ctx->fit_times[2] = ph[1]; ctx->fit_times[3] = t3 - t2;
    ctx->fit_times[4] = t3 - t0; ctx->fit_times[5] = 1.0;
    return rc;
}

// The per-rank form of sckm_kmeans_fit for one-process-per-GPU deployments (torchrun): this rank holds rows
// [row_offset, row_offset + n_local) of the n_global x d matrix, the context is joined to the others by
// sckm_comm_init_rank.  Same outputs; labels_out covers this rank's rows.
int sckm_kmeans_fit_shard(sckm_ctx* ctx, const void* x_local, uint64_t n_local, uint64_t d, int dtype, int column_major,
                          uint64_t row_offset, uint64_t n_global, uint64_t k, uint64_t max_iter, uint64_t first_index,
                          const double* uniforms, void* labels_out, int width, int64_t* size_out, double* centroids_out,
                          double* distortion_out, int64_t* iters_out) {
    if (!ctx) return SCKM_ERR_INVALID;
    if (!centroids_out) return fail(ctx, SCKM_ERR_INVALID, "centroids_out is NULL");
    if (n_global == 0) return fail(ctx, SCKM_ERR_INVALID, "empty input");
    if (!x_local && n_local) return fail(ctx, SCKM_ERR_INVALID, "host pointer is NULL");
    if (labels_out && width != 4 && width != 8) return fail(ctx, SCKM_ERR_INVALID, "label width must be 4 or 8");
    auto now = []() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t0 = now();
    sckm_dataset* ds = nullptr;
    SCKM_TRY(sckm_dataset_upload(ctx, x_local, n_local, d, dtype, column_major, row_offset, n_global, &ds));
    const double t1 = now();
    double ph[2] = {0, 0};
    int rc = fit_compute(ds, k, max_iter, first_index, uniforms, size_out, centroids_out, distortion_out, iters_out, ph);
    const double t2 = now();
    if (rc == SCKM_OK && labels_out) rc = download_labels(ds, labels_out, width);
    sckm_dataset_destroy(ds);
    const double t3 = now();
    ctx->fit_times[0] = t1 - t0; ctx->fit_times[1] = ph[0]; ctx->fit_times[2] = ph[1]; ctx->fit_times[3] = t3 - t2;
    ctx->fit_times[4] = t3 - t0; ctx->fit_times[5] = (double)ctx->nranks;
    return rc;
}

int sckm_ctx_last_fit_times(const sckm_ctx* ctx, double* out6) {
    if (!ctx || !out6) return SCKM_ERR_INVALID;
    for (int i = 0; i < 6; i++) out6[i] = ctx->fit_times[i];
    return SCKM_OK;
}

// ---- measurement --------------------------------------------------------------------------------
int sckm_device_peaks(sckm_ctx* ctx, double* out3) {
    if (!ctx || !out3) return SCKM_ERR_INVALID;
    SCKM_CUDA(ctx, cudaSetDevice(ctx->device));
    return measure_peaks(ctx, out3);
}

int sckm_flush_l2(sckm_ctx* ctx) {
    if (!ctx) return SCKM_ERR_INVALID;
    SCKM_CUDA(ctx, cudaSetDevice(ctx->device));
    if (!ctx->d_flush) {
        ctx->flush_bytes = (size_t)256 << 20;  // > 126 MB L2
        SCKM_CUDA(ctx, cudaMalloc(&ctx->d_flush, ctx->flush_bytes));
    }
    SCKM_CUDA(ctx, cudaMemsetAsync(ctx->d_flush, 0, ctx->flush_bytes, ctx->stream));
    return SCKM_OK;
}

}  // extern "C"
