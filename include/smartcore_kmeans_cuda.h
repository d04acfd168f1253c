/* =============================================================================
 * smartcore_kmeans_cuda.h -- C ABI of libsmartcore_kmeans_cuda.so
 *
 * B200 (sm_100a) implementation of smartcore v0.4.0's k-means hot path.  This is
 * the drop-in boundary: plain pointers and sizes, opaque handles, int status
 * codes.  It is what a Rust `src/gpu/ffi.rs` (cargo feature `cuda`) binds; see
 * INTEGRATION.md for the binding and the edit to src/cluster/kmeans.rs.
 *
 * Reference interfaces replaced (paths relative to the smartcore source tree):
 *   KMeans::fit              src/cluster/kmeans.rs:254-323
 *   KMeans::predict          src/cluster/kmeans.rs:327-352
 *   KMeans::kmeans_plus_plus src/cluster/kmeans.rs:354-413
 *   BBDTree::clustering      src/algorithm/neighbour/bbd_tree.rs:62-163
 *   Euclidian::squared_distance  src/metrics/distance/euclidian.rs:51-66
 *
 * Conventions
 *   * Every function returns SCKM_OK (0) or an error code; the message is kept
 *     per context (sckm_last_error).  No C++ exception crosses this boundary.
 *   * The caller owns every host buffer for the duration of the call only; the
 *     library keeps no host pointer after return.  Device memory lives behind
 *     the opaque handles and is released by the matching *_destroy.
 *   * A context is not thread-safe.  Multi-GPU comes in two forms: ONE context over
 *     all devices of the box (sckm_ctx_create_multi: the drop-in form, the caller
 *     stays single-threaded), or one single-device context per process / host
 *     thread joined by sckm_comm_init_rank (torchrun).  Either way rows are sharded
 *     across ranks and the only per-iteration exchange is one all-reduce of
 *     [k*d sums | k counts | inertia] (f64).
 *   * The RNG stays on the host (kmeans.rs:355, src/rand_custom.rs:8-33): the
 *     caller draws `first_index = rng.gen_range(0..n)` and the k-1 uniforms
 *     `rng.gen::<f64>()` (they do not depend on the data) and passes them in.
 *   * There is no CPU fallback: every compute entry point fails with
 *     SCKM_ERR_CUDA when no sm_100 device is usable.
 * ============================================================================= */
#ifndef SMARTCORE_KMEANS_CUDA_H
#define SMARTCORE_KMEANS_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SCKM_ABI_VERSION 3

typedef struct sckm_ctx sckm_ctx;         /* device + stream + workspaces (+ NCCL communicator) */
typedef struct sckm_dataset sckm_dataset; /* this rank's rows of X, labels y and D^2 array, on device */

enum { SCKM_F32 = 0, SCKM_F64 = 1 };      /* element type TX of X */

enum {
    SCKM_OK = 0,
    SCKM_ERR_INVALID = 1, /* bad argument (null pointer, k > n, d mismatch, ...) */
    SCKM_ERR_CUDA = 2,    /* CUDA runtime / driver error, or no usable device */
    SCKM_ERR_NCCL = 3,    /* NCCL not loadable or collective failed */
    SCKM_ERR_STATE = 4    /* call sequence error (e.g. lloyd before labels exist) */
};

/* Which Lloyd assignment kernel to run.  AUTO picks by shape. */
enum {
    SCKM_ASSIGN_AUTO = 0,
    SCKM_ASSIGN_DIRECT = 1, /* direct-difference form, bit-exact distances (euclidian.rs:56-63) */
    SCKM_ASSIGN_DMMA = 2,   /* ||x||^2 - 2 X.C^T + ||c||^2 on FP64 DMMA tiles + exact near-tie refine */
    SCKM_ASSIGN_STREAM = 3, /* small k (<16), d <= 32: one HBM pass does assignment and update */
    SCKM_ASSIGN_TC5 = 4     /* f32 data: tcgen05 (TMA + TMEM) ranking, 3xFP16 for d <= 32, 3xTF32 for d <= 64, + exact f64 decision */
};

int sckm_abi_version(void);

/* ---- context ------------------------------------------------------------- */
int  sckm_ctx_create(int device, sckm_ctx** out);
void sckm_ctx_destroy(sckm_ctx* ctx);
/* ONE context over several devices of this box, driven from the caller's single thread -- what a smartcore program
 * calling KMeans::fit / predict (kmeans.rs:254, :327) gets: the devices stay invisible to the caller.
 * n_dev <= 0: every visible device; dev_ids nullable (0..n_dev-1).  One device: identical to sckm_ctx_create.
 * The result behaves as a context on dev_ids[0] for every dataset-level entry point; the whole-matrix calls
 * sckm_kmeans_fit and sckm_predict shard the caller's host buffer in contiguous row blocks over the devices (one host
 * thread, stream, staging ring and NCCL communicator per device; one all-reduce of k*d+k+1 doubles per Lloyd step).
 * Inputs too small for every device to get >= 32768 rows run on dev_ids[0] alone. */
int  sckm_ctx_create_multi(int n_dev, const int* dev_ids, sckm_ctx** out);
/* Devices behind this context (1 for sckm_ctx_create). */
int  sckm_ctx_device_count(const sckm_ctx* ctx);
/* Wall-clock phases of the last sckm_kmeans_fit on this context, seconds: out6 = {upload, kmeans++ + initial means,
 * Lloyd loop, label download, total, devices used} (max over devices per phase). */
int  sckm_ctx_last_fit_times(const sckm_ctx* ctx, double* out6);
/* How the last Lloyd loop of this context summed its per-step vector over the ranks: SCKM_ALLREDUCE_NONE (one rank),
 * SCKM_ALLREDUCE_NCCL, or SCKM_ALLREDUCE_PEER -- the one-shot sum over peer memory inside the finalize kernel
 * (csrc/sckm_peer.cu; the default whenever every rank can map every other; SCKM_PEER_ALLREDUCE=0 turns it off). */
#define SCKM_ALLREDUCE_NONE 0
#define SCKM_ALLREDUCE_NCCL 1
#define SCKM_ALLREDUCE_PEER 2
int  sckm_ctx_allreduce_path(const sckm_ctx* ctx);
/* Last error text of this context (or of the failed sckm_ctx_create when ctx == NULL). */
const char* sckm_last_error(const sckm_ctx* ctx);
/* Force an assignment kernel (SCKM_ASSIGN_*); default AUTO. */
int  sckm_ctx_set_assign_kernel(sckm_ctx* ctx, int which);
/* Kernels launched by this context since creation (for bench.py's gpu_launches). */
uint64_t sckm_ctx_launch_count(const sckm_ctx* ctx);

/* ---- multi-GPU (NCCL, loaded with dlopen; one rank per context) ------------ */
/* Fill a 128-byte ncclUniqueId on rank 0; ship it to the other ranks by any means. */
int sckm_comm_unique_id(sckm_ctx* ctx, void* id128);
int sckm_comm_init_rank(sckm_ctx* ctx, int nranks, int rank, const void* id128);

/* ---- dataset: rows [row_offset, row_offset + n_local) of a global n_global x d matrix -- */
/* Replaces the container reads of DenseMatrix (src/linalg/basic/matrix.rs:367-381,
 * 391-404): `column_major` != 0 means host[c * n_local + r], else host[r * d + c].
 * Device storage is always row-major. */
int sckm_dataset_upload(sckm_ctx* ctx, const void* host, uint64_t n_local, uint64_t d, int dtype,
                        int column_major, uint64_t row_offset, uint64_t n_global, sckm_dataset** out);
/* Synthetic Gaussian blobs generated on device (recipe of make_blobs,
 * src/dataset/generator.rs:10-48, with a counter-based RNG so that any row can be
 * regenerated on the host bit-for-bit: sckm_blobs_fill_host). */
int sckm_dataset_generate_blobs(sckm_ctx* ctx, uint64_t n_local, uint64_t d, uint64_t n_centers,
                                uint64_t seed, int dtype, uint64_t row_offset, uint64_t n_global,
                                sckm_dataset** out);
/* Host twin of the generator (pure CPU, no device needed): rows [row0, row0+nrows) row-major. */
int sckm_blobs_fill_host(void* out, int dtype, uint64_t row0, uint64_t nrows, uint64_t d,
                         uint64_t n_centers, uint64_t seed);
int sckm_dataset_download_rows(sckm_dataset* ds, uint64_t local_row0, uint64_t nrows, void* host_out);
void sckm_dataset_destroy(sckm_dataset* ds);

/* ---- kmeans++ labels (kmeans.rs:354-413) ----------------------------------- */
/* first_index: global row drawn by rng.gen_range(0..n); uniforms[k-1]: the gen::<f64>() draws.
 * inject_rows (nullable, k entries): bypass the D^2 sampling and use these global rows as seeds
 * (parity harness).  On return the dataset holds the labels y (nearest chosen seed) and
 * seed_rows_out (nullable) the k chosen global rows. */
int sckm_kmeanspp(sckm_dataset* ds, uint64_t k, uint64_t first_index, const double* uniforms,
                  const int64_t* inject_rows, int64_t* seed_rows_out);

/* ---- initial centroids = per-label means (kmeans.rs:275-292) ---------------- */
int sckm_init_centroids(sckm_dataset* ds, uint64_t k, double* centroids_out, int64_t* size_out);

/* ---- one Lloyd step == BBDTree::clustering (bbd_tree.rs:62-87) --------------- */
/* in: centroids[k*d]; out: sums[k*d], counts[k], *inertia (w.r.t. the INPUT centroids); the
 * dataset's labels are overwritten.  All-reduced over ranks when a communicator is set. */
int sckm_lloyd_step(sckm_dataset* ds, const double* centroids, uint64_t k, double* sums_out,
                    int64_t* counts_out, double* inertia_out);

/* ---- the loop of KMeans::fit (kmeans.rs:294-310) with the reference stop rule ---- */
/* centroids_inout: initial centroids in, final out.  size_out[k], *distortion_out,
 * *iters_out = number of clustering steps executed. */
int sckm_lloyd_fit(sckm_dataset* ds, uint64_t k, uint64_t max_iter, double* centroids_inout,
                   int64_t* size_out, double* distortion_out, int64_t* iters_out);
/* Same loop with the stop rule disabled (exactly n_iters steps) and device timing:
 * ms_per_iter_out[n_iters] (nullable) are CUDA-event times of each whole iteration on the context's
 * stream, assign_ms_out[n_iters] (nullable) those of the assignment kernel alone (the dominant
 * kernel, for the roofline).  With ms_per_iter_out but WITHOUT assign_ms_out the loop is timed as a whole --
 * two events, none between the steps, which is how a fit runs -- and every slot carries the mean.
 * inertia_out (nullable) gets n_iters values. */
int sckm_lloyd_iterate(sckm_dataset* ds, uint64_t k, uint64_t n_iters, double* centroids_inout,
                       int64_t* size_out, double* inertia_out, float* ms_per_iter_out,
                       float* assign_ms_out);

/* Labels of this rank's rows; width = 4 (uint32_t) or 8 (uint64_t, Rust usize). */
int sckm_labels_download(sckm_dataset* ds, void* out, int width);
/* The D^2 array of kmeans++; after a Lloyd step, the per-point min distance (direct-form and tile kernels only:
 * the streaming kernel for k < 16 does not materialise it). */
int sckm_mindist_download(sckm_dataset* ds, double* out);

/* ---- predict (kmeans.rs:327-352): direct form in f64, strict <, lowest index wins ---- */
int sckm_predict(sckm_ctx* ctx, const void* x_host, uint64_t n, uint64_t d, int dtype,
                 int column_major, const double* centroids, uint64_t k, void* labels_out, int width);

/* ---- whole KMeans::fit from host buffers (what the Rust fit() calls) ----------- */
/* Validation (k >= 2, max_iter >= 1) stays with the caller so the reference's messages are
 * produced before any device work.  labels_out: n x width bytes. */
int sckm_kmeans_fit(sckm_ctx* ctx, const void* x_host, uint64_t n, uint64_t d, int dtype,
                    int column_major, uint64_t k, uint64_t max_iter, uint64_t first_index,
                    const double* uniforms, void* labels_out, int width, int64_t* size_out,
                    double* centroids_out, double* distortion_out, int64_t* iters_out);

/* The per-rank form for one-process-per-GPU deployments: this rank passes rows [row_offset, row_offset + n_local)
 * of the n_global x d matrix (x_local points at ITS rows; column_major: a [d][n_local] image), the context having
 * been joined to the other ranks by sckm_comm_init_rank.  size / centroids / distortion / iters are the global
 * results on every rank; labels_out covers this rank's rows. */
int sckm_kmeans_fit_shard(sckm_ctx* ctx, const void* x_local, uint64_t n_local, uint64_t d, int dtype,
                          int column_major, uint64_t row_offset, uint64_t n_global, uint64_t k, uint64_t max_iter,
                          uint64_t first_index, const double* uniforms, void* labels_out, int width,
                          int64_t* size_out, double* centroids_out, double* distortion_out, int64_t* iters_out);

/* ---- cluster quality (src/metrics/cluster_helpers.rs:7-25 contingency_matrix) --------
 * out[n_classes][k] (row-major) = number of rows with class id c and cluster label j, counted on the device.
 * sckm_contingency: cluster labels = the dataset's resident labels (after a fit / Lloyd step); class_ids_host holds
 * this rank's n_local dense class ids in [0, n_classes).  With a communicator the table is summed over ranks.
 * sckm_contingency_host: both id arrays come from the host (any two labelings of n rows).
 * An id outside its range is an error (SCKM_ERR_INVALID).  Entropy / mutual information / homogeneity,
 * completeness and V-measure (cluster_hcv.rs:36-55) follow from the table on the host. */
int sckm_contingency(sckm_dataset* ds, const uint32_t* class_ids_host, uint64_t n_classes, uint64_t k,
                     int64_t* out);
int sckm_contingency_host(sckm_ctx* ctx, const uint32_t* a_host, const uint32_t* b_host, uint64_t n,
                          uint64_t na, uint64_t nb, int64_t* out);

/* ---- batched k-nearest neighbours (src/algorithm/neighbour/linear_search.rs:52-84 with Euclidian::distance) ----
 * For each of the nq query rows (host, row-major, the dataset's element type) the k nearest rows of this rank's
 * resident dataset: idx_out[nq][k] = global row indices, dist_out[nq][k] = Euclidian::distance values (bit-identical
 * to the reference's arithmetic), ascending by (distance, index).  1 <= k <= n, else SCKM_ERR_INVALID with the
 * reference's message; any d (rows that are not multiples of 16 bytes take an element-wise staging path) and any k
 * (k > 64 runs ceil(k/64) passes over the rows).  The reference returns the same neighbours in its heap's internal
 * order and resolves EXACT ties at the k-th distance by that heap's layout; here the lowest indices win.  Rows at NaN
 * or +inf distance are never neighbours (linear_search.rs:62-76: only `d < INFINITY` enters the heap): the unfilled
 * places carry idx -1 / dist +inf (the reference returns fewer tuples). */
int sckm_knn(sckm_dataset* ds, const void* queries_host, uint64_t nq, uint64_t k, int64_t* idx_out,
             double* dist_out);

/* LinearKNNSearch::find_radius (linear_search.rs:89-110): every row with Euclidian distance <= radius, in ascending
 * row order, for nq queries.  The result is ragged, hence two calls: sckm_radius_count fills counts_out[nq]; the caller
 * builds offsets[nq] = exclusive prefix of the counts, sizes idx_out / dist_out to total = sum(counts) and calls
 * sckm_radius_fill, which writes query q's neighbours at [offsets[q], offsets[q] + counts[q]).  radius <= 0 is
 * SCKM_ERR_INVALID with the reference's message.  sckm_radius_fill recomputes the counts: offsets/total that do not
 * match them (another query set, radius or dataset) are SCKM_ERR_INVALID and nothing is written outside a query's
 * slot [offsets[q], offsets[q+1]) (the last slot ends at total). */
int sckm_radius_count(sckm_dataset* ds, const void* queries_host, uint64_t nq, double radius, int64_t* counts_out);
int sckm_radius_fill(sckm_dataset* ds, const void* queries_host, uint64_t nq, double radius,
                     const int64_t* offsets, uint64_t total, int64_t* idx_out, double* dist_out);

/* ---- measurement helpers ------------------------------------------------------- */
/* out[0] = HBM copy GB/s (read+write), out[1] = FP64 DFMA TFLOP/s, out[2] = FP64 DMMA TFLOP/s,
 * measured now on the context's device with CUDA events (micro-kernels, ~100 ms). */
int sckm_device_peaks(sckm_ctx* ctx, double* out3);
/* Write > L2-size bytes so the next timed launch starts with a cold L2. */
int sckm_flush_l2(sckm_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* SMARTCORE_KMEANS_CUDA_H */
