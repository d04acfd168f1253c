// sckm_multi.cu -- ONE process, ONE context, every GPU of the box: what `KMeans::fit(&x, params)` /
// `predict` (src/cluster/kmeans.rs:254, :327) get when the drop-in boundary is called by a single-threaded
// smartcore program.  The reference has no notion of devices; the G devices stay invisible to the caller
// (threading row).
//
// A multi-GPU context is an ordinary single-device context (device dev_ids[0]: every dataset-level entry point works
// on it unchanged) plus G joined per-device contexts, rank r = dev_ids[r], each with its own stream, workspaces,
// pinned staging ring and NCCL communicator (ncclCommInitRank from G threads of this process).  The whole-matrix
// calls shard the caller's ONE host buffer in contiguous row blocks (aligned to the 1024-row kmeans++ summation
// block, the same split as smartcore_b200/dist.py) and run the single-rank code path of sckm_api.cu on G host
// threads -- exactly what G torchrun ranks would execute, minus the processes:
//   sckm_kmeans_fit : per rank  upload(shard) | kmeans++ | means | Lloyd loop | labels -> caller's slice
//                     collectives per Lloyd step: one all-reduce of [k*d sums | k counts | inertia] (sckm_nccl.cu);
//                     every rank evaluates the stop rule on the same all-reduced inertia, so all break together.
//   sckm_predict    : per rank  its slice of the rows, no collective.
// The upload phase also reserves every workspace the compute phase needs and ends with a join: if any rank failed to
// get its memory, every rank stops BEFORE the first collective (a rank that never enqueues its all-reduce would hang
// the others).
#include "sckm_common.cuh"
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdlib>
#include <thread>

namespace sckm {

struct MultiExt {
    std::vector<sckm_ctx*> dev;        // rank r -> its single-device context (communicator joined)
    uint64_t min_rows_per_dev = 32768; // below this a device's share is not worth a collective per step
};

static double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

// run fn(r) for every rank on its own host thread (rank 0 on the caller's); returns the first non-zero status
template <typename F> static int on_all(const MultiExt* m, F&& fn) {
    const int G = (int)m->dev.size();
    std::vector<int> rc(G, SCKM_OK);
    std::vector<std::thread> th;
    th.reserve(G);
    for (int r = 1; r < G; r++) th.emplace_back([&, r]() { rc[r] = fn(r); });
    rc[0] = fn(0);
    for (auto& t : th) t.join();
    for (int r = 0; r < G; r++) if (rc[r] != SCKM_OK) return rc[r];
    return SCKM_OK;
}

// first error text among the ranks -> the parent context
static int adopt_error(sckm_ctx* ctx, int rc) {
    if (rc == SCKM_OK) return rc;
    for (size_t r = 0; r < ctx->multi->dev.size(); r++)
        if (!ctx->multi->dev[r]->err.empty()) return fail(ctx, rc, "device %d (rank %zu): %s", ctx->multi->dev[r]->device, r, ctx->multi->dev[r]->err.c_str());
    return fail(ctx, rc, "multi-GPU call failed (status %d)", rc);
}

// Row bounds of the G shards: bounds[r] .. bounds[r+1].  false = run on the primary device alone (single-device
// context, too few rows for every device to get a worthwhile, non-empty share).
bool multi_shards(const sckm_ctx* ctx, uint64_t n, std::vector<uint64_t>* bounds) {
    if (!ctx || !ctx->multi) return false;
    const uint64_t G = ctx->multi->dev.size();
    if (G < 2) return false;
    uint64_t min_rows = ctx->multi->min_rows_per_dev;
    if (const char* e = getenv("SCKM_MULTI_MIN_ROWS")) min_rows = strtoull(e, nullptr, 10);   // tests: shard small inputs too
    uint64_t per = (n + G - 1) / G;
    per = (per + kKppBlockRows - 1) / kKppBlockRows * kKppBlockRows;
    if (per * (G - 1) >= n) return false;                         // the last rank(s) would be empty
    if (n / G < min_rows) return false;
    if (bounds) {
        bounds->resize(G + 1);
        for (uint64_t r = 0; r <= G; r++) (*bounds)[r] = std::min(n, r * per);
    }
    return true;
}

uint64_t multi_launch_count(const sckm_ctx* ctx) {
    uint64_t s = 0;
    if (ctx && ctx->multi) for (sckm_ctx* c : ctx->multi->dev) s += c->launches;
    return s;
}

void multi_destroy(sckm_ctx* ctx) {
    if (!ctx || !ctx->multi) return;
    MultiExt* m = ctx->multi;
    ctx->multi = nullptr;
    // communicators are torn down by their own threads (ncclCommDestroy of one rank may wait for its peers)
    std::vector<std::thread> th;
    for (sckm_ctx* c : m->dev) th.emplace_back([c]() { sckm_ctx_destroy(c); });
    for (auto& t : th) t.join();
    delete m;
}

int multi_kmeans_fit(sckm_ctx* ctx, const void* x_host, uint64_t n, uint64_t d, int dtype, int column_major, uint64_t k,
                     uint64_t max_iter, uint64_t first_index, const double* uniforms, void* labels_out, int width,
                     int64_t* size_out, double* centroids_out, double* distortion_out, int64_t* iters_out) {
    MultiExt* m = ctx->multi;
    std::vector<uint64_t> b;
    if (!multi_shards(ctx, n, &b)) return fail(ctx, SCKM_ERR_STATE, "multi_kmeans_fit on an unsharded input");
    const int G = (int)m->dev.size();
    for (sckm_ctx* c : m->dev) c->err.clear();
    std::vector<sckm_dataset*> ds(G, nullptr);
    const double t0 = now_s();
    // phase 1: every rank lands its rows (G staging rings work on disjoint slices of the caller's buffer)
    int rc = on_all(m, [&](int r) {
        const int rc_up = fit_upload(m->dev[r], x_host, n, b[r], b[r + 1] - b[r], d, dtype, column_major, &ds[r]);
        return rc_up != SCKM_OK ? rc_up : fit_reserve(ds[r], k);      // every allocation of the compute phase, before the join
    });
    const double t1 = now_s();
    // phase 2: the single-rank driver on every rank; collectives inside keep the ranks in lock step
    std::vector<double> ph(2 * G, 0.0);
    std::vector<int64_t> iters(G, 0);
    std::vector<double> dist(G, 0.0);
    if (rc == SCKM_OK)
        rc = on_all(m, [&](int r) {
            return fit_compute(ds[r], k, max_iter, first_index, uniforms, r == 0 ? size_out : nullptr, r == 0 ? centroids_out : nullptr,
                               &dist[r], &iters[r], &ph[2 * r]);
        });
    const double t2 = now_s();
    if (rc == SCKM_OK)
        for (int r = 1; r < G; r++)
            if (iters[r] != iters[0] || !(dist[r] == dist[0] || (dist[r] != dist[r] && dist[0] != dist[0]))) {
                m->dev[r]->err = "ranks disagree on the stop rule";
                rc = SCKM_ERR_STATE;
            }
    // phase 3: labels of every shard straight into the caller's slice
    if (rc == SCKM_OK && labels_out)
        rc = on_all(m, [&](int r) {
            if (cudaSetDevice(m->dev[r]->device) != cudaSuccess) return (int)SCKM_ERR_CUDA;
            return download_labels(ds[r], (char*)labels_out + b[r] * (size_t)width, width);
        });
    on_all(m, [&](int r) { if (ds[r]) { cudaSetDevice(m->dev[r]->device); sckm_dataset_destroy(ds[r]); } return (int)SCKM_OK; });
    const double t3 = now_s();
    if (rc != SCKM_OK) return adopt_error(ctx, rc);
    if (distortion_out) *distortion_out = dist[0];
    if (iters_out) *iters_out = iters[0];
    double kpp = 0, loop = 0;
    for (int r = 0; r < G; r++) { kpp = std::max(kpp, ph[2 * r]); loop = std::max(loop, ph[2 * r + 1]); }
    ctx->fit_times[0] = t1 - t0; ctx->fit_times[1] = kpp; ctx->fit_times[2] = loop; ctx->fit_times[3] = t3 - t2;
    ctx->fit_times[4] = t3 - t0; ctx->fit_times[5] = (double)G;
    return SCKM_OK;
}

int multi_predict(sckm_ctx* ctx, const void* x_host, uint64_t n, uint64_t d, int dtype, int column_major,
                  const double* centroids, uint64_t k, void* labels_out, int width) {
    MultiExt* m = ctx->multi;
    std::vector<uint64_t> b;
    if (!multi_shards(ctx, n, &b)) return fail(ctx, SCKM_ERR_STATE, "multi_predict on an unsharded input");
    for (sckm_ctx* c : m->dev) c->err.clear();
    const int rc = on_all(m, [&](int r) {
        return predict_rows(m->dev[r], x_host, n, b[r], b[r + 1] - b[r], d, dtype, column_major, centroids, k,
                            (char*)labels_out + b[r] * (size_t)width, width);
    });
    return adopt_error(ctx, rc);
}

}  // namespace sckm

using namespace sckm;

extern "C" {

int sckm_ctx_create_multi(int n_dev, const int* dev_ids, sckm_ctx** out) {
    if (!out) return fail(nullptr, SCKM_ERR_INVALID, "sckm_ctx_create_multi: out is NULL");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(nullptr, SCKM_ERR_CUDA, "no CUDA device available (%s); this library has no CPU fallback",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    std::vector<int> ids;
    if (n_dev <= 0) { for (int i = 0; i < count; i++) ids.push_back(i); }      // every visible device
    else for (int i = 0; i < n_dev; i++) ids.push_back(dev_ids ? dev_ids[i] : i);
    for (size_t i = 0; i < ids.size(); i++) {
        if (ids[i] < 0 || ids[i] >= count) return fail(nullptr, SCKM_ERR_INVALID, "device %d out of range (0..%d)", ids[i], count - 1);
        for (size_t j = 0; j < i; j++) if (ids[j] == ids[i]) return fail(nullptr, SCKM_ERR_INVALID, "device %d listed twice", ids[i]);
    }
    sckm_ctx* ctx = nullptr;
    SCKM_TRY(sckm_ctx_create(ids[0], &ctx));
    if (ids.size() == 1) { *out = ctx; return SCKM_OK; }                        // one device: nothing to join
    const int G = (int)ids.size();
    MultiExt* m = new MultiExt();
    ctx->multi = m;
    const unsigned hc = std::thread::hardware_concurrency();
    for (int r = 0; r < G; r++) {
        sckm_ctx* c = nullptr;
        const int rc = sckm_ctx_create(ids[r], &c);
        if (rc != SCKM_OK) {
            const std::string why = sckm_last_error(nullptr);
            sckm_ctx_destroy(ctx);
            return fail(nullptr, rc, "device %d: %s", ids[r], why.c_str());
        }
        c->ingest_max_threads = (int)std::max(2u, std::min(8u, (hc ? hc : 16u) / (2u * (unsigned)G)));
        m->dev.push_back(c);
    }
    unsigned char id[128];
    int rc = sckm_comm_unique_id(m->dev[0], id);
    if (rc == SCKM_OK) rc = on_all(m, [&](int r) { return sckm_comm_init_rank(m->dev[r], G, r, id); });
    if (rc != SCKM_OK) {
        std::string why = "NCCL communicator setup failed";
        for (sckm_ctx* c : m->dev) if (!c->err.empty()) { why = c->err; break; }
        sckm_ctx_destroy(ctx);
        return fail(nullptr, rc, "%s", why.c_str());
    }
    *out = ctx;
    return SCKM_OK;
}

int sckm_ctx_allreduce_path(const sckm_ctx* ctx) {
    if (!ctx) return SCKM_ALLREDUCE_NONE;
    // a multi-GPU context reports what its devices did in the last sharded fit (an unsharded one ran on the context itself)
    if (ctx->multi && ctx->fit_times[5] > 1.0) return ctx->multi->dev[0]->allreduce_path;
    return ctx->allreduce_path;
}

int sckm_ctx_device_count(const sckm_ctx* ctx) {
    if (!ctx) return 0;
    return ctx->multi ? (int)ctx->multi->dev.size() : 1;
}

}  // extern "C"
