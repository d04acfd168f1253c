// sckm_knn.cu -- batched brute-force k-nearest-neighbour search over the rows of a resident dataset
// (SURVEY.md section 8(f) rank 2b): the batched form of LinearKNNSearch::find
// (src/algorithm/neighbour/linear_search.rs:52-84) with Euclidian::distance
// (src/metrics/distance/euclidian.rs:51-76: squared_distance(..).sqrt()) as the metric, i.e. what KNNClassifier /
// KNNRegressor / DBSCAN ask of it when built with Distances::euclidian().
//
// Arithmetic is the reference's: (a-b) and its square in TX, widened, sequential f64 sum, then sqrt -- every
// returned distance is bit-identical to Euclidian::distance.  Candidates are ranked by (distance, row index):
// the k smallest, ascending.  The reference returns the same neighbours in the internal order of its HeapSelection
// (heap_select.rs:62-83) and, among rows whose distance ties EXACTLY with the k-th smallest, the survivor depends on
// that heap's layout; callers sort or aggregate the result (linear_search.rs:151-160), so the contract here is:
// identical distance multiset always, identical index set whenever the k-th distance is not tied.
//
//   knn_tile_kernel : grid (row chunks, tiles of 8 queries).  A warp stages 32 rows at a time (16-byte cp.async
//                     into a padded slab, one lane = one row), computes its row's distance to the 8 queries held
//                     in shared memory (broadcast reads) and offers it to the warp's private top-k list of each
//                     query: ballot of the lanes that beat the current k-th entry, warp-parallel sorted insertion
//                     (insertion point = popcount of the entries smaller than the candidate).
//   knn_merge_kernel: one CTA per query selects the k smallest of all partial lists by k rounds of "smallest
//                     entry greater than the previous pick" -- a pure reduction, hence deterministic.
//   Any d and any k <= n are accepted, as in the reference: rows that are not multiples of 16 bytes are staged
//   element by element (same padded slab, same arithmetic), and k > 64 runs ceil(k / 64) passes, each collecting
//   the next 64 entries above the (distance, index) of the previous pass's last pick.
#include "sckm_common.cuh"
#include <algorithm>
#include <cfloat>

namespace sckm {

constexpr int KNN_WARPS = 4;
constexpr int KNN_TQ = 8;          // queries per CTA
constexpr int KNN_MAXK = 64;       // two list slots per lane

__device__ __forceinline__ void knn_cp_async16(void* smem_dst, const void* gmem_src) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src));
}
__device__ __forceinline__ void knn_cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
}
// (a-b)^2 with the reference's rounding: subtract and multiply in TX, widen to f64 (no FMA)
__device__ __forceinline__ double knn_sqdiff(double a, double b) { const double r = __dsub_rn(a, b); return __dmul_rn(r, r); }
__device__ __forceinline__ double knn_sqdiff(float a, float b) { const float r = __fsub_rn(a, b); return (double)__fmul_rn(r, r); }

// Stage `nrows` consecutive rows (one contiguous block starting at `src`) into a warp's slab, one padded slab row per
// data row.  Rows that are multiples of 16 bytes go as 16-byte cp.async chunks (consecutive lanes, consecutive chunks);
// any other row length element by element with plain loads (the block need not even be 16-byte aligned then).
template <typename T>
__device__ __forceinline__ void knn_stage_rows(unsigned char* slab, const T* src, uint32_t nrows, uint32_t d, uint32_t pitch16, int lane) {
    const uint32_t row_bytes = d * sizeof(T);
    if (row_bytes % 16 == 0) {
        const uint32_t cpr = row_bytes / 16, total = nrows * cpr;
        const unsigned char* s8 = reinterpret_cast<const unsigned char*>(src);
        for (uint32_t c = lane; c < total; c += 32) {
            const uint32_t r = c / cpr, qq = c - r * cpr;
            knn_cp_async16(slab + ((size_t)r * pitch16 + qq) * 16, s8 + (size_t)c * 16);
        }
        knn_cp_async_wait_all();
    } else {
        const uint32_t total = nrows * d;
        for (uint32_t e = lane; e < total; e += 32) {
            const uint32_t r = e / d, j = e - r * d;
            reinterpret_cast<T*>(slab + (size_t)r * pitch16 * 16)[j] = __ldg(src + e);
        }
    }
    __syncwarp();
}

// (distance, index) lexicographic order
__device__ __forceinline__ bool knn_less(double da, uint32_t ia, double db, uint32_t ib) {
    return da < db || (da == db && ia < ib);
}

template <typename T>
__global__ void __launch_bounds__(KNN_WARPS * 32)
knn_tile_kernel(const T* __restrict__ x, uint64_t n, uint32_t d, const T* __restrict__ queries, uint32_t nq, uint32_t k,
                uint32_t pitch16, uint64_t rows_per_cta, const double* __restrict__ lb_dist, const uint32_t* __restrict__ lb_idx,
                double* __restrict__ part_dist, uint32_t* __restrict__ part_idx) {
    extern __shared__ __align__(16) unsigned char smem_k[];
    const uint32_t row_bytes = d * sizeof(T);
    const size_t qbytes = ((size_t)KNN_TQ * row_bytes + 15) / 16 * 16;
    T* qbuf = reinterpret_cast<T*>(smem_k);                                          // [TQ][d]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char* slab = smem_k + qbytes + (size_t)warp * 32 * pitch16 * 16;       // this warp's 32 staged rows
    double* ld = reinterpret_cast<double*>(smem_k + qbytes + (size_t)KNN_WARPS * 32 * pitch16 * 16) +
                 (size_t)warp * KNN_TQ * KNN_MAXK;                                   // [TQ][64] distances, ascending
    uint32_t* li = reinterpret_cast<uint32_t*>(smem_k + qbytes + (size_t)KNN_WARPS * 32 * pitch16 * 16 +
                                               (size_t)KNN_WARPS * KNN_TQ * KNN_MAXK * sizeof(double)) +
                   (size_t)warp * KNN_TQ * KNN_MAXK;                                 // [TQ][64] row indices

    const uint32_t q0 = blockIdx.y * KNN_TQ;
    const uint32_t tq = min((uint32_t)KNN_TQ, nq - q0);
    for (uint32_t e = threadIdx.x; e < tq * d; e += blockDim.x) qbuf[e] = queries[(size_t)q0 * d + e];
    __syncthreads();
    uint32_t len[KNN_TQ];                                                            // list lengths (warp-uniform)
    double lbd[KNN_TQ]; uint32_t lbi[KNN_TQ];                                        // only entries ABOVE this bound compete
#pragma unroll
    for (int q = 0; q < KNN_TQ; q++) {
        len[q] = 0;
        lbd[q] = (uint32_t)q < tq ? lb_dist[q0 + q] : -1.0;
        lbi[q] = (uint32_t)q < tq ? lb_idx[q0 + q] : 0u;
    }

    const uint64_t r_begin = (uint64_t)blockIdx.x * rows_per_cta, r_end = min(n, r_begin + rows_per_cta);
    for (uint64_t row0 = r_begin + (uint64_t)warp * 32; row0 < r_end; row0 += KNN_WARPS * 32) {
        const uint32_t nrows = (uint32_t)min((uint64_t)32, r_end - row0);
        knn_stage_rows<T>(slab, x + row0 * d, nrows, d, pitch16, lane);
        double dq[KNN_TQ];
#pragma unroll
        for (int q = 0; q < KNN_TQ; q++) dq[q] = DBL_MAX;
        if (lane < nrows) {
            const T* xr = reinterpret_cast<const T*>(slab + (size_t)lane * pitch16 * 16);
            double s[KNN_TQ];
#pragma unroll
            for (int q = 0; q < KNN_TQ; q++) s[q] = 0.0;
            for (uint32_t j = 0; j < d; j++) {                   // feature-major: the row element is read once, every
                const T xv = xr[j];                              // query keeps its own sequential f64 sum
#pragma unroll
                for (int q = 0; q < KNN_TQ; q++)
                    if ((uint32_t)q < tq) s[q] = __dadd_rn(s[q], knn_sqdiff(xv, qbuf[(size_t)q * d + j]));
            }
#pragma unroll
            for (int q = 0; q < KNN_TQ; q++)
                if ((uint32_t)q < tq) dq[q] = __dsqrt_rn(s[q]);  // Euclidian::distance
        }
        const uint32_t my_idx = (uint32_t)(row0 + lane);
#pragma unroll
        for (int q = 0; q < KNN_TQ; q++) {
            if ((uint32_t)q >= tq) continue;                                         // warp-uniform
            double* lq = ld + q * KNN_MAXK;
            uint32_t* iq = li + q * KNN_MAXK;
            // lanes whose row beats the current k-th entry.  NaN and +inf distances never do: the reference's heap
            // starts out full of INFINITY entries and only `d < datum.distance` replaces one (linear_search.rs:62-76)
            bool cand = lane < nrows && dq[q] < INFINITY && knn_less(lbd[q], lbi[q], dq[q], my_idx);
            if (cand && len[q] == k) cand = knn_less(dq[q], my_idx, lq[k - 1], iq[k - 1]);
            unsigned ball = __ballot_sync(0xffffffffu, cand);
            while (ball) {
                const int src = __ffs(ball) - 1;
                ball &= ball - 1;
                const double cd = __shfl_sync(0xffffffffu, dq[q], src);
                const uint32_t ci = (uint32_t)(row0 + src);
                // entries this lane owns: positions lane and lane + 32
                const bool h0 = (uint32_t)lane < len[q], h1 = (uint32_t)lane + 32 < len[q];
                const double e0d = h0 ? lq[lane] : 0.0, e1d = h1 ? lq[lane + 32] : 0.0;
                const uint32_t e0i = h0 ? iq[lane] : 0u, e1i = h1 ? iq[lane + 32] : 0u;
                const unsigned s0 = __ballot_sync(0xffffffffu, h0 && knn_less(e0d, e0i, cd, ci));
                const unsigned s1 = __ballot_sync(0xffffffffu, h1 && knn_less(e1d, e1i, cd, ci));
                const uint32_t pos = __popc(s0) + __popc(s1);                        // entries smaller than the candidate
                if (pos >= k) continue;                                              // an earlier insertion raised the bar
                __syncwarp();                                                        // everyone has read its entries
                if (h1 && (uint32_t)lane + 32 >= pos && (uint32_t)lane + 33 < k) { lq[lane + 33] = e1d; iq[lane + 33] = e1i; }
                if (h0 && (uint32_t)lane >= pos && (uint32_t)lane + 1 < k) { lq[lane + 1] = e0d; iq[lane + 1] = e0i; }
                __syncwarp();                                                        // shifts done before the insert lands
                if (lane == 0) { lq[pos] = cd; iq[pos] = ci; }
                len[q] = min(len[q] + 1u, k);
                __syncwarp();
            }
        }
        __syncwarp();                                                                // slab free for the next group
    }
    // this warp's lists -> global partials [query][list][k], padded with (+inf, 0xffffffff)
    const uint32_t nlists = gridDim.x * KNN_WARPS, mylist = blockIdx.x * KNN_WARPS + warp;
    __syncwarp();
#pragma unroll
    for (int q = 0; q < KNN_TQ; q++) {
        if ((uint32_t)q >= tq) continue;
        double* od = part_dist + ((size_t)(q0 + q) * nlists + mylist) * k;
        uint32_t* oi = part_idx + ((size_t)(q0 + q) * nlists + mylist) * k;
        for (uint32_t p = lane; p < k; p += 32) {
            const bool has = p < len[q];
            od[p] = has ? ld[q * KNN_MAXK + p] : INFINITY;
            oi[p] = has ? li[q * KNN_MAXK + p] : 0xffffffffu;
        }
    }
}

// one CTA per query: k rounds of "smallest (distance, index) greater than the previous pick" over all partial entries
__global__ void __launch_bounds__(256)
knn_merge_kernel(const double* __restrict__ part_dist, const uint32_t* __restrict__ part_idx, uint32_t nentries, uint32_t k,
                 uint64_t row_offset, uint32_t out_stride, uint32_t out_off, double* __restrict__ lb_dist, uint32_t* __restrict__ lb_idx,
                 long long* __restrict__ idx_out, double* __restrict__ dist_out) {
    __shared__ double sd[8];
    __shared__ uint32_t si[8];
    __shared__ double pick_d;
    __shared__ uint32_t pick_i;
    const double* pd = part_dist + (size_t)blockIdx.x * nentries;
    const uint32_t* pi = part_idx + (size_t)blockIdx.x * nentries;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double prev_d = -1.0;                                         // distances are >= 0; (-1, 0) precedes everything
    uint32_t prev_i = 0;
    for (uint32_t r = 0; r < k; r++) {
        double bd = INFINITY; uint32_t bi = 0xffffffffu;          // the padding entry is the identity
        for (uint32_t e = threadIdx.x; e < nentries; e += blockDim.x) {
            const double dd = pd[e]; const uint32_t ii = pi[e];
            if (knn_less(prev_d, prev_i, dd, ii) && knn_less(dd, ii, bd, bi)) { bd = dd; bi = ii; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double od = __shfl_xor_sync(0xffffffffu, bd, o);
            const uint32_t oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (knn_less(od, oi, bd, bi)) { bd = od; bi = oi; }
        }
        if (lane == 0) { sd[warp] = bd; si[warp] = bi; }
        __syncthreads();
        if (threadIdx.x == 0) {
            double d0 = sd[0]; uint32_t i0 = si[0];
            for (int w = 1; w < 8; w++) if (knn_less(sd[w], si[w], d0, i0)) { d0 = sd[w]; i0 = si[w]; }
            pick_d = d0; pick_i = i0;
            idx_out[(size_t)blockIdx.x * out_stride + out_off + r] = i0 == 0xffffffffu ? -1ll : (long long)(row_offset + i0);
            dist_out[(size_t)blockIdx.x * out_stride + out_off + r] = d0;
            // the last pick of this pass is the bound of the next one (k > 64); the padding entry (+inf, 2^32-1) ends the search
            if (r + 1 == k) { lb_dist[blockIdx.x] = d0; lb_idx[blockIdx.x] = i0; }
        }
        __syncthreads();
        prev_d = pick_d; prev_i = pick_i;
        __syncthreads();
    }
}

__global__ void knn_bound_init_kernel(double* __restrict__ lb_dist, uint32_t* __restrict__ lb_idx, uint32_t nq) {
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q < nq) { lb_dist[q] = -1.0; lb_idx[q] = 0u; }            // distances are >= 0: (-1, 0) precedes every entry
}

template <typename T>
static int knn_t(sckm_dataset* ds, const T* d_queries, uint64_t nq, uint64_t k, long long* d_idx, double* d_dist) {
    sckm_ctx* ctx = ds->ctx;
    const uint32_t d = (uint32_t)ds->d;
    const uint32_t row_bytes = d * sizeof(T), pitch16 = ((row_bytes + 15) / 16) | 1;
    const size_t qbytes = ((size_t)KNN_TQ * row_bytes + 15) / 16 * 16;
    const size_t smem = qbytes + (size_t)KNN_WARPS * 32 * pitch16 * 16 + (size_t)KNN_WARPS * KNN_TQ * KNN_MAXK * (sizeof(double) + sizeof(uint32_t));
    if (smem > (size_t)ctx->smem_optin) return fail(ctx, SCKM_ERR_INVALID, "d=%u too large for the k-NN tile kernel", d);
    const unsigned qtiles = (unsigned)((nq + KNN_TQ - 1) / KNN_TQ);
    // enough CTAs to fill the GPU a few times over, but no more row chunks than that needs: every chunk adds
    // KNN_WARPS partial lists per query to the merge
    const uint64_t groups = (ds->n + KNN_WARPS * 32 - 1) / (KNN_WARPS * 32);
    const uint64_t want_ctas = (uint64_t)ctx->num_sms * 4;
    uint64_t chunks = std::max<uint64_t>(1, std::min<uint64_t>(groups, (want_ctas + qtiles - 1) / qtiles));
    const uint64_t rows_per_cta = (groups + chunks - 1) / chunks * (KNN_WARPS * 32);
    chunks = (ds->n + rows_per_cta - 1) / rows_per_cta;
    const uint32_t nlists = (uint32_t)chunks * KNN_WARPS;
    const uint32_t kpass = (uint32_t)std::min<uint64_t>(k, KNN_MAXK);          // list length of one pass
    double* part_dist = nullptr; uint32_t* part_idx = nullptr; double* lb_dist = nullptr; uint32_t* lb_idx = nullptr;
    auto cleanup = [&]() { dev_free(ctx, part_dist); dev_free(ctx, part_idx); dev_free(ctx, lb_dist); dev_free(ctx, lb_idx); };
    if (dev_alloc(ctx, (void**)&part_dist, (size_t)nq * nlists * kpass * sizeof(double)) != cudaSuccess ||
        dev_alloc(ctx, (void**)&part_idx, (size_t)nq * nlists * kpass * sizeof(uint32_t)) != cudaSuccess ||
        dev_alloc(ctx, (void**)&lb_dist, nq * sizeof(double)) != cudaSuccess ||
        dev_alloc(ctx, (void**)&lb_idx, nq * sizeof(uint32_t)) != cudaSuccess) {
        cleanup();
        return fail(ctx, SCKM_ERR_CUDA, "cudaMalloc for the k-NN partial lists failed");
    }
    auto kern = knn_tile_kernel<T>;
    int rc = SCKM_OK;
    if (smem > 48 * 1024 && cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        rc = fail(ctx, SCKM_ERR_CUDA, "cudaFuncSetAttribute failed: %s", cudaGetErrorString(cudaGetLastError()));
    if (rc == SCKM_OK) {
        knn_bound_init_kernel<<<(unsigned)((nq + 255) / 256), 256, 0, ctx->stream>>>(lb_dist, lb_idx, (uint32_t)nq);
        ctx->launches++;
        // k <= 64: one pass.  Larger k: pass p collects entries [64 p, 64 p + kk) of the sorted result, i.e. the kk
        // smallest (distance, index) pairs above the last pick of pass p - 1; the distances are recomputed per pass.
        for (uint64_t done = 0; done < k; done += KNN_MAXK) {
            const uint32_t kk = (uint32_t)std::min<uint64_t>(KNN_MAXK, k - done);
            kern<<<dim3((unsigned)chunks, qtiles), KNN_WARPS * 32, smem, ctx->stream>>>((const T*)ds->x, ds->n, d, d_queries, (uint32_t)nq,
                kk, pitch16, rows_per_cta, lb_dist, lb_idx, part_dist, part_idx);
            knn_merge_kernel<<<(unsigned)nq, 256, 0, ctx->stream>>>(part_dist, part_idx, nlists * kk, kk, ds->row_offset, (uint32_t)k,
                                                                  (uint32_t)done, lb_dist, lb_idx, d_idx, d_dist);
            ctx->launches += 2;
        }
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) rc = fail(ctx, SCKM_ERR_CUDA, "k-NN kernel launch failed: %s", cudaGetErrorString(e));
    }
    cleanup();
    return rc;
}

// ---- find_radius (linear_search.rs:89-110): every row with distance <= radius, in ascending row order ----
// Same staging and arithmetic as knn_tile_kernel.  A warp owns a CONTIGUOUS run of rows of its CTA's chunk, so the
// per-(query, chunk, warp) segments concatenate in row order.  Pass 1 (FILL = false) counts the hits of every segment,
// radius_scan_kernel turns the counts into segment offsets (+ the caller's per-query offsets), pass 2 recomputes the
// same distances and writes (index, distance) at segment offset + running count + rank within the ballot.
template <typename T, bool FILL>
__global__ void __launch_bounds__(KNN_WARPS * 32)
radius_tile_kernel(const T* __restrict__ x, uint64_t n, uint32_t d, const T* __restrict__ queries, uint32_t nq, double radius,
                   uint32_t pitch16, uint64_t rows_per_cta, uint32_t* __restrict__ seg_cnt, const unsigned long long* __restrict__ seg_off,
                   const unsigned long long* __restrict__ slot_end, uint64_t row_offset, long long* __restrict__ idx_out,
                   double* __restrict__ dist_out) {
    extern __shared__ __align__(16) unsigned char smem_k[];
    const uint32_t row_bytes = d * sizeof(T);
    const size_t qbytes = ((size_t)KNN_TQ * row_bytes + 15) / 16 * 16;
    T* qbuf = reinterpret_cast<T*>(smem_k);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char* slab = smem_k + qbytes + (size_t)warp * 32 * pitch16 * 16;
    const uint32_t q0 = blockIdx.y * KNN_TQ;
    const uint32_t tq = min((uint32_t)KNN_TQ, nq - q0);
    for (uint32_t e = threadIdx.x; e < tq * d; e += blockDim.x) qbuf[e] = queries[(size_t)q0 * d + e];
    __syncthreads();
    const uint32_t nseg = gridDim.x * KNN_WARPS, myseg = blockIdx.x * KNN_WARPS + warp;
    unsigned long long run[KNN_TQ];                                 // hits so far in this segment (warp-uniform)
    unsigned long long lim[KNN_TQ];                                 // end of the caller's slot for the query: never write past it
#pragma unroll
    for (int q = 0; q < KNN_TQ; q++) {
        run[q] = (FILL && (uint32_t)q < tq) ? seg_off[(size_t)(q0 + q) * nseg + myseg] : 0ull;
        lim[q] = (FILL && (uint32_t)q < tq) ? slot_end[q0 + q] : 0ull;
    }

    const uint64_t rows_per_warp = rows_per_cta / KNN_WARPS;        // rows_per_cta is a multiple of 32 * KNN_WARPS
    const uint64_t w_begin = min(n, (uint64_t)blockIdx.x * rows_per_cta + (uint64_t)warp * rows_per_warp);
    const uint64_t w_end = min(n, w_begin + rows_per_warp);
    const unsigned lt = (1u << lane) - 1u;
    for (uint64_t row0 = w_begin; row0 < w_end; row0 += 32) {
        const uint32_t nrows = (uint32_t)min((uint64_t)32, w_end - row0);
        knn_stage_rows<T>(slab, x + row0 * d, nrows, d, pitch16, lane);
        double s[KNN_TQ];
#pragma unroll
        for (int q = 0; q < KNN_TQ; q++) s[q] = 0.0;
        if (lane < nrows) {
            const T* xr = reinterpret_cast<const T*>(slab + (size_t)lane * pitch16 * 16);
            for (uint32_t j = 0; j < d; j++) {
                const T xv = xr[j];
#pragma unroll
                for (int q = 0; q < KNN_TQ; q++)
                    if ((uint32_t)q < tq) s[q] = __dadd_rn(s[q], knn_sqdiff(xv, qbuf[(size_t)q * d + j]));
            }
        }
#pragma unroll
        for (int q = 0; q < KNN_TQ; q++) {
            if ((uint32_t)q >= tq) continue;
            const double dist = __dsqrt_rn(s[q]);
            const bool hit = lane < nrows && dist <= radius;       // `d <= radius` (linear_search.rs:101); NaN: no
            const unsigned ball = __ballot_sync(0xffffffffu, hit);
            if (FILL && hit) {
                const unsigned long long o = run[q] + __popc(ball & lt);
                if (o < lim[q]) {                                   // (a mismatch is reported by radius_scan_kernel)
                    idx_out[o] = (long long)(row_offset + row0 + lane);
                    dist_out[o] = dist;
                }
            }
            run[q] += __popc(ball);
        }
        __syncwarp();
    }
    if (!FILL && lane == 0)
#pragma unroll
        for (int q = 0; q < KNN_TQ; q++)
            if ((uint32_t)q < tq) seg_cnt[(size_t)(q0 + q) * nseg + myseg] = (uint32_t)run[q];
}

// per query: exclusive prefix of its segment counts (+ base offset when given), total into counts_out
// Fill mode (base != nullptr): the caller's slot of query q is [base[q], base[q+1]) (the last one ends at `total`);
// slot_end[q] receives its end, and a recomputed count that does not fill the slot exactly -- offsets that belong to
// another query set, another radius or another dataset -- raises *mismatch (the fill kernel never writes past a slot).
__global__ void radius_scan_kernel(const uint32_t* __restrict__ seg_cnt, uint32_t nq, uint32_t nseg,
                                   const long long* __restrict__ base, unsigned long long total,
                                   unsigned long long* __restrict__ seg_off, unsigned long long* __restrict__ slot_end,
                                   unsigned long long* __restrict__ mismatch, long long* __restrict__ counts_out) {
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    unsigned long long run = base ? (unsigned long long)base[q] : 0ull;
    const unsigned long long start = run;
    for (uint32_t sgm = 0; sgm < nseg; sgm++) {
        if (seg_off) seg_off[(size_t)q * nseg + sgm] = run;
        run += seg_cnt[(size_t)q * nseg + sgm];
    }
    if (counts_out) counts_out[q] = (long long)(run - start);
    if (base) {
        const unsigned long long end = q + 1 < nq ? (unsigned long long)base[q + 1] : total;
        slot_end[q] = end;
        if (run != end) atomicAdd(mismatch, 1ull);
    }
}


// fill == false: counts_out[nq]; fill == true: offsets[nq] (exclusive prefix of the counts, caller-computed) and
// idx_out / dist_out of sum(counts) entries
template <typename T>
static int radius_t(sckm_dataset* ds, const T* d_queries, uint64_t nq, double radius, bool fill, const long long* d_offsets,
                    uint64_t total, long long* d_counts, long long* d_idx, double* d_dist) {
    sckm_ctx* ctx = ds->ctx;
    const uint32_t d = (uint32_t)ds->d;
    const uint32_t row_bytes = d * sizeof(T), pitch16 = ((row_bytes + 15) / 16) | 1;
    const size_t qbytes = ((size_t)KNN_TQ * row_bytes + 15) / 16 * 16;
    const size_t smem = qbytes + (size_t)KNN_WARPS * 32 * pitch16 * 16;
    if (smem > (size_t)ctx->smem_optin) return fail(ctx, SCKM_ERR_INVALID, "d=%u too large for the radius kernel", d);
    const unsigned qtiles = (unsigned)((nq + KNN_TQ - 1) / KNN_TQ);
    const uint64_t groups = (ds->n + KNN_WARPS * 32 - 1) / (KNN_WARPS * 32);
    const uint64_t want_ctas = (uint64_t)ctx->num_sms * 4;
    uint64_t chunks = std::max<uint64_t>(1, std::min<uint64_t>(groups, (want_ctas + qtiles - 1) / qtiles));
    const uint64_t rows_per_cta = (groups + chunks - 1) / chunks * (KNN_WARPS * 32);
    chunks = (ds->n + rows_per_cta - 1) / rows_per_cta;
    const uint32_t nseg = (uint32_t)chunks * KNN_WARPS;
    uint32_t* seg_cnt = nullptr; unsigned long long *seg_off = nullptr, *slot_end = nullptr;   // slot_end[nq] | mismatch count
    if (dev_alloc(ctx, (void**)&seg_cnt, (size_t)nq * nseg * sizeof(uint32_t)) != cudaSuccess ||
        (fill && (dev_alloc(ctx, (void**)&seg_off, (size_t)nq * nseg * sizeof(unsigned long long)) != cudaSuccess ||
                  dev_alloc(ctx, (void**)&slot_end, (nq + 1) * sizeof(unsigned long long)) != cudaSuccess))) {
        dev_free(ctx, seg_cnt); dev_free(ctx, seg_off); dev_free(ctx, slot_end);
        return fail(ctx, SCKM_ERR_CUDA, "cudaMalloc for the radius segments failed");
    }
    int rc = SCKM_OK;
    auto count_kern = radius_tile_kernel<T, false>;
    auto fill_kern = radius_tile_kernel<T, true>;
    if (smem > 48 * 1024 && (cudaFuncSetAttribute(count_kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ||
                             cudaFuncSetAttribute(fill_kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess))
        rc = fail(ctx, SCKM_ERR_CUDA, "cudaFuncSetAttribute failed: %s", cudaGetErrorString(cudaGetLastError()));
    if (rc == SCKM_OK) {
        const dim3 grid((unsigned)chunks, qtiles);
        count_kern<<<grid, KNN_WARPS * 32, smem, ctx->stream>>>((const T*)ds->x, ds->n, d, d_queries, (uint32_t)nq, radius, pitch16,
                                                              rows_per_cta, seg_cnt, nullptr, nullptr, ds->row_offset, nullptr, nullptr);
        if (fill) cudaMemsetAsync(slot_end + nq, 0, sizeof(unsigned long long), ctx->stream);
        radius_scan_kernel<<<(unsigned)((nq + 127) / 128), 128, 0, ctx->stream>>>(seg_cnt, (uint32_t)nq, nseg, fill ? d_offsets : nullptr,
                                                                                  total, seg_off, slot_end, fill ? slot_end + nq : nullptr,
                                                                                  fill ? nullptr : d_counts);
        ctx->launches += 2;
        if (fill) {
            fill_kern<<<grid, KNN_WARPS * 32, smem, ctx->stream>>>((const T*)ds->x, ds->n, d, d_queries, (uint32_t)nq, radius, pitch16,
                                                                 rows_per_cta, nullptr, seg_off, slot_end, ds->row_offset, d_idx, d_dist);
            ctx->launches++;
        }
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) rc = fail(ctx, SCKM_ERR_CUDA, "radius kernel launch failed: %s", cudaGetErrorString(e));
        if (rc == SCKM_OK && fill) {
            unsigned long long bad = 0;
            if (cudaMemcpyAsync(&bad, slot_end + nq, sizeof(bad), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
                cudaStreamSynchronize(ctx->stream) != cudaSuccess)
                rc = fail(ctx, SCKM_ERR_CUDA, "radius search failed: %s", cudaGetErrorString(cudaGetLastError()));
            else if (bad)
                rc = fail(ctx, SCKM_ERR_INVALID, "sckm_radius_fill: offsets/total do not match the counts of these queries "
                          "(%llu of %llu queries differ): run sckm_radius_count on the same dataset, queries and radius first",
                          bad, (unsigned long long)nq);
        }
    }
    dev_free(ctx, seg_cnt); dev_free(ctx, seg_off); dev_free(ctx, slot_end);
    return rc;
}

// counts_out != nullptr: pass 1 only.  Otherwise offsets_host[nq] + total entries: pass 1 + scan + pass 2.
int radius_search(sckm_dataset* ds, const void* queries_host, uint64_t nq, double radius, int64_t* counts_out,
                  const int64_t* offsets_host, uint64_t total, int64_t* idx_out, double* dist_out) {
    sckm_ctx* ctx = ds->ctx;
    if (!(radius > 0.0)) return fail(ctx, SCKM_ERR_INVALID, "radius should be > 0");                     // linear_search.rs:90-95
    if (nq == 0) return SCKM_OK;
    if (nq > 65535ull * KNN_TQ) return fail(ctx, SCKM_ERR_INVALID, "radius search: at most %d queries per call", 65535 * KNN_TQ);
    SCKM_CUDA(ctx, cudaSetDevice(ctx->device));
    const bool fill = counts_out == nullptr;
    if (fill)                                                       // slots must be ordered and inside [0, total]
        for (uint64_t q = 0; q < nq; q++) {
            const int64_t end = q + 1 < nq ? offsets_host[q + 1] : (int64_t)total;
            if (offsets_host[q] < 0 || offsets_host[q] > end || (uint64_t)end > total)
                return fail(ctx, SCKM_ERR_INVALID, "sckm_radius_fill: offsets[%llu] is not an exclusive prefix within total=%llu",
                            (unsigned long long)q, (unsigned long long)total);
        }
    const size_t qbytes = (size_t)nq * ds->d * ds->elem();
    void* d_q = nullptr; long long *d_cnt = nullptr, *d_off = nullptr, *d_idx = nullptr; double* d_dist = nullptr;
    auto cleanup = [&]() { dev_free(ctx, d_q); dev_free(ctx, d_cnt); dev_free(ctx, d_off); dev_free(ctx, d_idx); dev_free(ctx, d_dist);
                           cudaStreamSynchronize(ctx->stream); };
    bool ok = dev_alloc(ctx, &d_q, qbytes) == cudaSuccess;
    if (ok && !fill) ok = dev_alloc(ctx, (void**)&d_cnt, nq * sizeof(long long)) == cudaSuccess;
    if (ok && fill) ok = dev_alloc(ctx, (void**)&d_off, nq * sizeof(long long)) == cudaSuccess &&
                         dev_alloc(ctx, (void**)&d_idx, std::max<uint64_t>(total, 1) * sizeof(long long)) == cudaSuccess &&
                         dev_alloc(ctx, (void**)&d_dist, std::max<uint64_t>(total, 1) * sizeof(double)) == cudaSuccess;
    if (!ok) { cleanup(); return fail(ctx, SCKM_ERR_CUDA, "cudaMalloc for the radius search failed"); }
    int rc = copy_to_device(ctx, d_q, queries_host, qbytes);
    if (rc == SCKM_OK && fill) rc = copy_to_device(ctx, d_off, offsets_host, nq * sizeof(long long));
    if (rc == SCKM_OK)
        rc = ds->dtype == SCKM_F32 ? radius_t<float>(ds, (const float*)d_q, nq, radius, fill, d_off, total, d_cnt, d_idx, d_dist)
                                   : radius_t<double>(ds, (const double*)d_q, nq, radius, fill, d_off, total, d_cnt, d_idx, d_dist);
    if (rc == SCKM_OK && !fill) rc = copy_to_host(ctx, counts_out, d_cnt, nq * sizeof(long long));
    if (rc == SCKM_OK && fill && total) {
        rc = copy_to_host(ctx, idx_out, d_idx, total * sizeof(long long));
        if (rc == SCKM_OK) rc = copy_to_host(ctx, dist_out, d_dist, total * sizeof(double));
    }
    cleanup();
    return rc;
}

// queries: nq rows of the dataset's element type, row-major, on the HOST; outputs on the host: [nq][k]
int knn_search(sckm_dataset* ds, const void* queries_host, uint64_t nq, uint64_t k, int64_t* idx_out, double* dist_out) {
    sckm_ctx* ctx = ds->ctx;
    if (k < 1 || k > ds->n) return fail(ctx, SCKM_ERR_INVALID, "k should be >= 1 and <= length(data)");   // linear_search.rs:53-58
    if (ds->n >= 0xFFFFFFFFull) return fail(ctx, SCKM_ERR_INVALID, "k-NN: more than 2^32-2 rows per rank");
    if (nq == 0) return SCKM_OK;
    if (nq > 65535ull * KNN_TQ) return fail(ctx, SCKM_ERR_INVALID, "k-NN: at most %d queries per call", 65535 * KNN_TQ);
    SCKM_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t qbytes = (size_t)nq * ds->d * ds->elem();
    void* d_q = nullptr; long long* d_idx = nullptr; double* d_dist = nullptr;
    auto cleanup = [&]() { dev_free(ctx, d_q); dev_free(ctx, d_idx); dev_free(ctx, d_dist); cudaStreamSynchronize(ctx->stream); };
    if (dev_alloc(ctx, &d_q, qbytes) != cudaSuccess || dev_alloc(ctx, (void**)&d_idx, nq * k * sizeof(long long)) != cudaSuccess ||
        dev_alloc(ctx, (void**)&d_dist, nq * k * sizeof(double)) != cudaSuccess) {
        cleanup();
        return fail(ctx, SCKM_ERR_CUDA, "cudaMalloc for the k-NN queries failed");
    }
    int rc = copy_to_device(ctx, d_q, queries_host, qbytes);
    if (rc == SCKM_OK)
        rc = ds->dtype == SCKM_F32 ? knn_t<float>(ds, (const float*)d_q, nq, k, d_idx, d_dist)
                                   : knn_t<double>(ds, (const double*)d_q, nq, k, d_idx, d_dist);
    if (rc == SCKM_OK) rc = copy_to_host(ctx, idx_out, d_idx, nq * k * sizeof(long long));
    if (rc == SCKM_OK) rc = copy_to_host(ctx, dist_out, d_dist, nq * k * sizeof(double));
    cleanup();
    return rc;
}

}  // namespace sckm
