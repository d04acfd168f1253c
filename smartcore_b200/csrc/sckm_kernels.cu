// sckm_kernels.cu -- hand-written sm_100a kernels of the k-means hot path (first generation:
// exact direct-form kernels + deterministic update; the DMMA tile kernel lives in sckm_dmma.cu).
//
// Reference arithmetic being reproduced (paths relative to the smartcore tree):
//   Euclidian::squared_distance  src/metrics/distance/euclidian.rs:51-66
//       diff and square in TX, widened per element to f64, sequential f64 sum, never fused.
//   kmeans_plus_plus             src/cluster/kmeans.rs:354-413
//   predict                      src/cluster/kmeans.rs:327-352
//   BBDTree::clustering          src/algorithm/neighbour/bbd_tree.rs:62-163 (dense equivalent)
#include "sckm_common.cuh"
#include "sckm_blobs.cuh"
#include <cfloat>
#include <cstdarg>
#include <cuda_bf16.h>

namespace sckm {

// ------------------------------------------------------------------------------------------
// small device helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
}

// (a-b)^2 with the reference's rounding: subtract and multiply in TX, widen to f64 (no FMA)
__device__ __forceinline__ double sqdiff(double a, double b) {
    double r = __dsub_rn(a, b);
    return __dmul_rn(r, r);
}
__device__ __forceinline__ double sqdiff(float a, float b) {
    float r = __fsub_rn(a, b);
    return (double)__fmul_rn(r, r);
}

template <typename T> struct Vec16;  // 16-byte vector of T
template <> struct Vec16<double> { using type = double2; static constexpr int N = 2; };
template <> struct Vec16<float>  { using type = float4;  static constexpr int N = 4; };

// pitch (in 16-byte units) of a staged row: odd, so that 8 lanes reading 16 B at consecutive rows
// hit 8 distinct 16-byte bank groups (conflict-free LDS.128)
__host__ __device__ inline uint32_t slab_pitch16(uint32_t row_bytes) { return (row_bytes / 16) | 1; }

// ------------------------------------------------------------------------------------------
// warp-level staging: 32 consecutive rows (a contiguous block in HBM) -> per-warp smem slab
// with padded pitch, using 16-byte cp.async (coalesced: consecutive lanes, consecutive 16 B).
// ------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ void stage_rows_vec(const T* __restrict__ x, uint64_t row0, uint32_t nrows,
                                               uint32_t d, unsigned char* slab, uint32_t pitch16, int lane) {
    const uint32_t chunks_per_row = d * sizeof(T) / 16;
    const uint32_t total = nrows * chunks_per_row;
    const unsigned char* src = reinterpret_cast<const unsigned char*>(x + row0 * d);
    for (uint32_t c = lane; c < total; c += 32) {
        uint32_t r = c / chunks_per_row, q = c - r * chunks_per_row;
        cp_async16(slab + ((size_t)r * pitch16 + q) * 16, src + (size_t)c * 16);
    }
    cp_async_wait_all();
    __syncwarp();
}

// same, but rows whose bit is set in `skipmask` are not fetched at all (kmeans++ triangle-inequality pruning)
template <typename T>
__device__ __forceinline__ void stage_rows_vec_masked(const T* __restrict__ x, uint64_t row0, uint32_t nrows, uint32_t d,
                                                      unsigned char* slab, uint32_t pitch16, int lane, unsigned skipmask) {
    const uint32_t chunks_per_row = d * sizeof(T) / 16;
    const uint32_t total = nrows * chunks_per_row;
    const unsigned char* src = reinterpret_cast<const unsigned char*>(x + row0 * d);
    uint32_t r = lane / chunks_per_row, q = lane % chunks_per_row;
    const uint32_t step_r = 32 / chunks_per_row, step_q = 32 % chunks_per_row;
    for (uint32_t c = lane; c < total; c += 32) {
        if (!((skipmask >> r) & 1u)) cp_async16(slab + ((size_t)r * pitch16 + q) * 16, src + (size_t)c * 16);
        r += step_r; q += step_q;
        if (q >= chunks_per_row) { q -= chunks_per_row; r++; }
    }
    cp_async_wait_all();
    __syncwarp();
}

// ------------------------------------------------------------------------------------------
// K5: kmeans++ D^2 refresh (kmeans.rs:368-379 / 399-410) + per-1024-row block sums (kmeans.rs:381-384)
// One CTA per 1024-row block, 4 warps, each warp stages 32 rows at a time; one lane = one row.
// TX arithmetic for diff/square, f64 sequential accumulation over the features: bit-identical D^2.
// ------------------------------------------------------------------------------------------
constexpr int KPP_WARPS = 4;

template <typename T, bool VEC>
__global__ void __launch_bounds__(KPP_WARPS * 32)
kpp_refresh_kernel(const T* __restrict__ x, uint64_t n, uint32_t d, const T* __restrict__ seedrow,
                   double* __restrict__ mind, uint32_t* __restrict__ labels, uint32_t label, int first_pass,
                   double* __restrict__ blocksum, uint32_t pitch16, const double* __restrict__ skiptab, uint32_t ntab) {
    // skiptab[j] (nullable) = (1 - margin) * ||new seed - seed j||^2 / 4: a row whose current D^2 (to its nearest
    // seed j = labels[row]) is <= skiptab[j] cannot be closer to the new seed (triangle inequality), so the
    // reference's `if dist < d[i]` (kmeans.rs:375) is false for it and the row need not even be read.
    extern __shared__ __align__(16) unsigned char smem[];
    T* cent = reinterpret_cast<T*>(smem);                       // [d] (padded to 16 B)
    const uint32_t cent_bytes = (d * sizeof(T) + 15) / 16 * 16;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char* slab = smem + cent_bytes + (size_t)warp * 32 * pitch16 * 16;
    double* tab = reinterpret_cast<double*>(smem + cent_bytes + (size_t)KPP_WARPS * 32 * pitch16 * 16 * (VEC ? 1 : 0));
    __shared__ double warp_part[KPP_WARPS];
    const bool prune = skiptab != nullptr && !first_pass;

    for (uint32_t j = threadIdx.x; j < d; j += blockDim.x) cent[j] = seedrow[j];
    if (prune) for (uint32_t j = threadIdx.x; j < ntab; j += blockDim.x) tab[j] = skiptab[j];
    __syncthreads();

    const uint64_t block_row0 = (uint64_t)blockIdx.x * kKppBlockRows;
    double acc_rows = 0.0;  // D^2 values this thread is responsible for, in a fixed order
    constexpr int SEGS = kKppBlockRows / 32;                    // 32-row segments of the block

    if (VEC) {
        // ---- phase 1: decide per row (coalesced reads of D^2 / label), compact the surviving rows of the block.
        // Pruned rows only contribute their unchanged D^2 to the block sum; with clustered data most of the block
        // vanishes here after the first few seeds, and phase 2 runs dense warps over what is left. ----
        __shared__ uint16_t list[kKppBlockRows];
        __shared__ uint32_t seg_cnt[SEGS];
        unsigned ball[SEGS / KPP_WARPS];
#pragma unroll
        for (int t = 0; t < SEGS / KPP_WARPS; t++) {
            const uint32_t local = (uint32_t)(t * KPP_WARPS + warp) * 32 + lane;
            const uint64_t r = block_row0 + local;
            bool active = false;
            if (r < n) {
                active = true;
                if (first_pass) labels[r] = 0;
                else if (prune) {
                    const double old = mind[r];
                    if (old <= tab[labels[r]]) { active = false; acc_rows = __dadd_rn(acc_rows, old); }
                }
            }
            ball[t] = __ballot_sync(0xffffffffu, active);
            if (lane == 0) seg_cnt[t * KPP_WARPS + warp] = __popc(ball[t]);
        }
        __syncthreads();
        uint32_t total = 0;
        {
            uint32_t run = 0, mybase[SEGS / KPP_WARPS];
#pragma unroll
            for (int sgi = 0; sgi < SEGS; sgi++) {                 // exclusive prefix over the 32 segments
                const uint32_t c = seg_cnt[sgi];
#pragma unroll
                for (int t = 0; t < SEGS / KPP_WARPS; t++) if (sgi == t * KPP_WARPS + warp) mybase[t] = run;
                run += c;
            }
            total = run;
#pragma unroll
            for (int t = 0; t < SEGS / KPP_WARPS; t++)
                if ((ball[t] >> lane) & 1u)
                    list[mybase[t] + __popc(ball[t] & ((1u << lane) - 1u))] = (uint16_t)((t * KPP_WARPS + warp) * 32 + lane);
        }
        __syncthreads();
        // ---- phase 2: dense warps over the compacted rows; each row's 16-byte chunks are fetched individually ----
        const uint32_t cpr = d * sizeof(T) / 16;
        const uint32_t lane_e = lane / cpr, lane_q = lane % cpr, step_e = 32 / cpr, step_q = 32 % cpr;
        for (uint32_t base = warp * 32; base < total; base += KPP_WARPS * 32) {
            const uint32_t cnt = min(32u, total - base);
            {
                uint32_t e = lane_e, q = lane_q;
                const uint32_t nchunks = cnt * cpr;
                for (uint32_t c = lane; c < nchunks; c += 32) {
                    const uint64_t r = block_row0 + list[base + e];
                    cp_async16(slab + ((size_t)e * pitch16 + q) * 16,
                               reinterpret_cast<const unsigned char*>(x + r * d) + (size_t)q * 16);
                    e += step_e; q += step_q;
                    if (q >= cpr) { q -= cpr; e++; }
                }
                cp_async_wait_all();
                __syncwarp();
            }
            if (lane < cnt) {
                const uint64_t r = block_row0 + list[base + lane];
                using V = typename Vec16<T>::type;
                const V* xr = reinterpret_cast<const V*>(slab + (size_t)lane * pitch16 * 16);
                const V* cr = reinterpret_cast<const V*>(cent);
                const uint32_t nv = d / Vec16<T>::N;
                double dist = 0.0;
                for (uint32_t q = 0; q < nv; q++) {
                    V xv = xr[q], cv = cr[q];
                    const T* xe = reinterpret_cast<const T*>(&xv);
                    const T* ce = reinterpret_cast<const T*>(&cv);
#pragma unroll
                    for (int e = 0; e < Vec16<T>::N; e++) dist = __dadd_rn(dist, sqdiff(xe[e], ce[e]));
                }
                double old = first_pass ? DBL_MAX : mind[r];
                if (dist < old) { old = dist; mind[r] = dist; labels[r] = label; }
                else if (first_pass) mind[r] = old;
                acc_rows = __dadd_rn(acc_rows, old);
            }
            __syncwarp();
        }
    } else {
        for (int t = 0; t < SEGS / KPP_WARPS; t++) {
            const uint64_t row0 = block_row0 + (uint64_t)(t * KPP_WARPS + warp) * 32;
            if (row0 >= n) break;
            const uint32_t nrows = (uint32_t)min((uint64_t)32, n - row0);
            if (lane < nrows) {
                const uint64_t r = row0 + lane;
                double old = first_pass ? DBL_MAX : mind[r];
                const bool skip = prune && old <= tab[labels[r]];
                if (first_pass) labels[r] = 0;
                if (!skip) {
                    const T* xr = x + r * d;
                    double dist = 0.0;
                    for (uint32_t j = 0; j < d; j++) dist = __dadd_rn(dist, sqdiff(xr[j], cent[j]));
                    if (dist < old) { old = dist; mind[r] = dist; labels[r] = label; }
                    else if (first_pass) mind[r] = old;
                }
                acc_rows = __dadd_rn(acc_rows, old);
            }
        }
    }
    // fixed-order block reduction: lanes (xor tree), then warps 0..3 sequentially
    double v = acc_rows;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = __dadd_rn(v, __shfl_xor_sync(0xffffffffu, v, o));
    if (lane == 0) warp_part[warp] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < KPP_WARPS; w++) s = __dadd_rn(s, warp_part[w]);
        blocksum[blockIdx.x] = s;
    }
}

// ------------------------------------------------------------------------------------------
// kmeans++ pass, second generation (16-byte-aligned rows, n < 2^32): three streaming kernels whose cost is
//   prune   : 12 B per row            (D^2 + label -> survivor list, one atomic per 4096-row CTA)
//   compute : d*s + 12 B per SURVIVOR (dense warps over the compacted list, rows gathered with cp.async)
//   sums    : 8 B per row             (per-1024-row block sums in an order that does not depend on the pruning
//                                      pattern, total by the last CTA to finish)
// instead of one kernel whose every 1024-row CTA paid three barriers and four dependent global round trips whether
// or not any of its rows survived (~200 us per pass at 10M rows however few rows were left).
// ------------------------------------------------------------------------------------------
constexpr int KPPA_THREADS = 256, KPPA_ITERS = 16;              // 4096 rows per CTA, 512 per warp

// Screening of a row that survived the triangle test, on a bf16 SHADOW of X (built by the first pass): with
//   r = x - seed0 (exact), xs = bf16(r) stored, ex >= ||r - xs||, t = f32(seed - seed0), et >= ||(seed - seed0) - t||
// the triangle inequality gives  ||x - seed|| >= ||xs - t|| - ex - et.  The f32 evaluation a of ||xs - t|| carries a
// relative error < (d+2)*2^-24, covered by `shrink`.  If that lower bound, squared and reduced by `margin` (the
// rounding of the reference's own TX-arithmetic distance), is still >= D^2[row], then `dist < d[i]` (kmeans.rs:375)
// is false for the exact distance too and the 4x (f64) / 2x (f32) larger exact row is never read.  Rows it cannot
// exclude -- the ones that do improve, and the borderline -- go to kpp_compute_kernel unchanged: results are
// bit-identical with or without the shadow.
template <int NV8>                                              // NV8 = d/8 when known at compile time (all loads in flight), 0 = loop
__device__ __forceinline__ bool kpp_screen_excludes(const uint4* __restrict__ srow, uint32_t nv8, const float* __restrict__ t_s,
                                                    float ex, double et, double old, float shrink, double margin) {
    float a2 = 0.f;
    auto eat = [&](const uint4& v, uint32_t q) {
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int e = 0; e < 4; e++) {
            const float lo = __uint_as_float(w[e] << 16), hi = __uint_as_float(w[e] & 0xffff0000u);   // bf16 -> f32
            const float r0 = lo - t_s[q * 8 + 2 * e], r1 = hi - t_s[q * 8 + 2 * e + 1];
            a2 = fmaf(r0, r0, a2);
            a2 = fmaf(r1, r1, a2);
        }
    };
    if (NV8 > 0) {
        uint4 v[NV8 > 0 ? NV8 : 1];
#pragma unroll
        for (int q = 0; q < NV8; q++) v[q] = __ldg(srow + q);
#pragma unroll
        for (int q = 0; q < NV8; q++) eat(v[q], q);
    } else {
        for (uint32_t q = 0; q < nv8; q++) eat(__ldg(srow + q), q);
    }
    // f32 overflow (|x - seed0| beyond ~1e19 with f64 data) must not pass for "far away": the bound only counts
    // while its f32 evaluation stayed finite; NaN anywhere on the way makes every comparison below false
    if (!(a2 <= 3.0e38f)) return false;
    const double lb = (double)(sqrtf(a2) * shrink) - (double)ex - et;
    return lb > 0.0 && lb * lb * (1.0 - margin) >= old;
}

template <int NV8>
__global__ void __launch_bounds__(KPPA_THREADS)
kpp_prune_kernel(const double* __restrict__ mind, const uint32_t* __restrict__ labels, uint64_t n,
                 const double* __restrict__ skiptab, uint32_t ntab, uint32_t* __restrict__ surv,
                 unsigned* __restrict__ counter, const uint16_t* __restrict__ shadow, const float* __restrict__ shadow_err,
                 const float* __restrict__ tshift, const double* __restrict__ tshift_err, uint32_t d, double margin) {
    extern __shared__ __align__(16) double tab_s[];             // [ntab] (1-margin)*||new seed - seed j||^2/4 | [d] f32 t
    __shared__ uint32_t warp_cnt[KPPA_THREADS / 32];
    __shared__ uint32_t cta_base;
    __shared__ uint16_t list_s[KPPA_THREADS / 32][32 * KPPA_ITERS];   // per warp: rows that passed the triangle test
    float* t_s = reinterpret_cast<float*>(tab_s + ntab);
    for (uint32_t j = threadIdx.x; j < ntab; j += blockDim.x) tab_s[j] = skiptab[j];
    if (shadow) for (uint32_t j = threadIdx.x; j < d; j += blockDim.x) t_s[j] = tshift[j];
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint64_t tile0 = ((uint64_t)blockIdx.x * (KPPA_THREADS / 32) + warp) * (32 * KPPA_ITERS);
    uint16_t* list = list_s[warp];
    // ---- phase 1: triangle test, coalesced over the warp's 512 rows; the rows it cannot exclude are compacted ----
    uint32_t cnt = 0;
#pragma unroll
    for (int t = 0; t < KPPA_ITERS; t++) {
        const uint64_t r = tile0 + (uint64_t)t * 32 + lane;
        bool active = false;
        if (r < n) active = !(mind[r] <= tab_s[labels[r]]);      // NaN D^2 stays active, like `dist < d[i]` would see it
        const unsigned ball = __ballot_sync(0xffffffffu, active);
        if (active) list[cnt + __popc(ball & ((1u << lane) - 1u))] = (uint16_t)(t * 32 + lane);
        cnt += __popc(ball);
    }
    __syncwarp();
    // ---- phase 2: dense warps over the compacted rows: screening on the bf16 shadow (every lane busy, all of a
    // row's loads in flight at once) ----
    unsigned pass[KPPA_ITERS];
    uint32_t cnt2 = 0;
    const double et = shadow ? *tshift_err : 0.0;
    const float shrink = 1.0f - (float)(d + 2) * 1.2e-7f;       // > 2 x the f32 summation + sqrt error bound
#pragma unroll
    for (int t = 0; t < KPPA_ITERS; t++) {
        bool keep = false;
        if ((uint32_t)(t * 32) < cnt) {                          // warp-uniform
            const uint32_t i = t * 32 + lane;
            if (i < cnt) {
                keep = true;
                if (shadow) {
                    const uint64_t r = tile0 + list[i];
                    keep = !kpp_screen_excludes<NV8>(reinterpret_cast<const uint4*>(shadow + r * d), d / 8, t_s, shadow_err[r], et,
                                                     mind[r], shrink, margin);
                }
            }
        }
        pass[t] = __ballot_sync(0xffffffffu, keep);
        cnt2 += __popc(pass[t]);
    }
    if (lane == 0) warp_cnt[warp] = cnt2;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t total = 0;
        for (int w = 0; w < KPPA_THREADS / 32; w++) total += warp_cnt[w];
        cta_base = total ? atomicAdd(counter, total) : 0u;
    }
    __syncthreads();
    uint32_t base = cta_base;
    for (int w = 0; w < warp; w++) base += warp_cnt[w];
#pragma unroll
    for (int t = 0; t < KPPA_ITERS; t++) {
        if ((pass[t] >> lane) & 1u) surv[base + __popc(pass[t] & ((1u << lane) - 1u))] = (uint32_t)(tile0 + list[t * 32 + lane]);
        base += __popc(pass[t]);
    }
}

// D^2 of the listed rows (surv == nullptr: all rows) against the new seed; bit-identical arithmetic to
// Euclidian::squared_distance (TX diff/square, sequential f64 sum) and the update rule of kmeans.rs:368-379
template <typename T>
__global__ void __launch_bounds__(KPP_WARPS * 32)
kpp_compute_kernel(const T* __restrict__ x, uint64_t n, uint32_t d, const T* __restrict__ seedrow, double* __restrict__ mind,
                   uint32_t* __restrict__ labels, uint32_t label, int first_pass, const uint32_t* __restrict__ surv,
                   const unsigned* __restrict__ counter, uint32_t pitch16, uint16_t* __restrict__ shadow,
                   float* __restrict__ shadow_err) {
    extern __shared__ __align__(16) unsigned char smem[];
    T* cent = reinterpret_cast<T*>(smem);
    const uint32_t cent_bytes = (d * sizeof(T) + 15) / 16 * 16;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char* slab = smem + cent_bytes + (size_t)warp * 32 * pitch16 * 16;
    for (uint32_t j = threadIdx.x; j < d; j += blockDim.x) cent[j] = seedrow[j];
    __syncthreads();
    const uint64_t total = surv ? (uint64_t)*counter : n;
    const uint32_t cpr = d * sizeof(T) / 16;                    // 16-byte chunks per row
    const uint32_t lane_e = lane / cpr, lane_q = lane % cpr, step_e = 32 / cpr, step_q = 32 % cpr;
    using V = typename Vec16<T>::type;
    const uint32_t nv = d / Vec16<T>::N;
    for (uint64_t base = ((uint64_t)blockIdx.x * KPP_WARPS + warp) * 32; base < total; base += (uint64_t)gridDim.x * KPP_WARPS * 32) {
        const uint32_t cnt = (uint32_t)min((uint64_t)32, total - base);
        const uint64_t my_row = lane < cnt ? (surv ? (uint64_t)surv[base + lane] : base + lane) : 0;
        {
            uint32_t e = lane_e, q = lane_q;
            const uint32_t nchunks = cnt * cpr;
            for (uint32_t c0 = 0; c0 < nchunks; c0 += 32) {       // uniform trip count: every lane reaches the shuffle
                const uint64_t r = __shfl_sync(0xffffffffu, my_row, (int)(e & 31u));
                if (c0 + lane < nchunks)
                    cp_async16(slab + ((size_t)e * pitch16 + q) * 16, reinterpret_cast<const unsigned char*>(x + r * d) + (size_t)q * 16);
                e += step_e; q += step_q;
                if (q >= cpr) { q -= cpr; e++; }
            }
            cp_async_wait_all();
            __syncwarp();
        }
        if (lane < cnt) {
            const V* xr = reinterpret_cast<const V*>(slab + (size_t)lane * pitch16 * 16);
            const V* cr = reinterpret_cast<const V*>(cent);
            double dist = 0.0;
            for (uint32_t q = 0; q < nv; q++) {
                V xv = xr[q], cv = cr[q];
                const T* xe = reinterpret_cast<const T*>(&xv);
                const T* ce = reinterpret_cast<const T*>(&cv);
#pragma unroll
                for (int e = 0; e < Vec16<T>::N; e++) dist = __dadd_rn(dist, sqdiff(xe[e], ce[e]));
            }
            if (first_pass) { labels[my_row] = 0; mind[my_row] = dist < DBL_MAX ? dist : DBL_MAX; }
            else if (dist < mind[my_row]) { mind[my_row] = dist; labels[my_row] = label; }
            if (first_pass && shadow) {
                // bf16 shadow of (x - seed 0) and the exact length of what the rounding dropped (see kpp_prune_kernel);
                // d % 8 == 0 here: eight elements -> one 16-byte store
                uint4* srow = reinterpret_cast<uint4*>(shadow + my_row * d);
                const T* xs = reinterpret_cast<const T*>(xr);
                double e2 = 0.0;
                for (uint32_t g8 = 0; g8 < d / 8; g8++) {
                    uint32_t w[4];
#pragma unroll
                    for (int e = 0; e < 8; e += 2) {
                        const uint32_t j = g8 * 8 + e;
                        const double r0 = (double)xs[j] - (double)cent[j], r1 = (double)xs[j + 1] - (double)cent[j + 1];
                        const __nv_bfloat16 b0 = __float2bfloat16_rn((float)r0), b1 = __float2bfloat16_rn((float)r1);
                        const double q0 = r0 - (double)__bfloat162float(b0), q1 = r1 - (double)__bfloat162float(b1);
                        e2 = fma(q0, q0, e2);
                        e2 = fma(q1, q1, e2);
                        w[e / 2] = (uint32_t)__bfloat16_as_ushort(b0) | ((uint32_t)__bfloat16_as_ushort(b1) << 16);
                    }
                    srow[g8] = make_uint4(w[0], w[1], w[2], w[3]);
                }
                shadow_err[my_row] = __double2float_ru(sqrt(e2) * (1.0 + 1e-9));
            }
        }
        __syncwarp();
    }
}

// block sums of D^2 (one warp per 1024-row block: lane-strided sequential sums, xor tree) and, by the last CTA to
// finish, the rank total in the order of kpp_total_kernel -- both independent of which rows were pruned
__global__ void __launch_bounds__(256)
kpp_blocksum_kernel(const double* __restrict__ mind, uint64_t n, uint32_t nb, double* __restrict__ blocksum,
                    double* __restrict__ totals, int rank, unsigned* __restrict__ done) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t b = blockIdx.x * 8 + warp;
    if (b < nb) {
        const uint64_t r0 = (uint64_t)b * kKppBlockRows;
        double v[kKppBlockRows / 32];
#pragma unroll
        for (int i = 0; i < kKppBlockRows / 32; i++) {          // all loads in flight, then a fixed-order sum
            const uint64_t r = r0 + (uint64_t)i * 32 + lane;
            v[i] = r < n ? __ldcg(mind + r) : 0.0;
        }
        double s = 0.0;
#pragma unroll
        for (int i = 0; i < kKppBlockRows / 32; i++) s = __dadd_rn(s, v[i]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s = __dadd_rn(s, __shfl_xor_sync(0xffffffffu, s, o));
        if (lane == 0) blocksum[b] = s;
    }
    __shared__ bool is_last;
    __shared__ double sh[256];
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = atomicAdd(done, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    const uint32_t per = (nb + 255) / 256;
    const uint32_t b0 = min(nb, threadIdx.x * per), b1 = min(nb, b0 + per);
    double s = 0.0;
    for (uint32_t i = b0; i < b1; i++) s = __dadd_rn(s, __ldcg(blocksum + i));
    sh[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        double t = 0.0;
#pragma unroll
        for (int i = 0; i < 8; i++) t = __dadd_rn(t, sh[threadIdx.x * 8 + i]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t = __dadd_rn(t, __shfl_xor_sync(0xffffffffu, t, o));
        if (threadIdx.x == 0) { totals[rank] = t; *done = 0u; }
    }
}

// seedtab[slot] = the seed just published in seedrow; skiptab[j] = (1-margin) * ||seed_slot - seed_j||^2 / 4, j < slot
template <typename T>
__global__ void __launch_bounds__(256)
kpp_seedtab_kernel(const T* __restrict__ seedrow, T* __restrict__ seedtab, uint32_t d, uint32_t slot,
                   double* __restrict__ skiptab, double margin, float* __restrict__ tshift, double* __restrict__ tshift_err) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (blockIdx.x == 0) {
        for (uint32_t j = threadIdx.x; j < d; j += blockDim.x) seedtab[(size_t)slot * d + j] = seedrow[j];
        if (tshift) {
            // the new seed relative to seed 0 (the origin of the bf16 shadow), rounded to f32, and the length of what
            // the rounding dropped -- both enter the screening bound of kpp_prune_kernel
            __shared__ double part[8];
            double e2 = 0.0;
            for (uint32_t j = threadIdx.x; j < d; j += blockDim.x) {
                const double rd = slot ? (double)seedrow[j] - (double)seedtab[j] : 0.0;
                const float tf = (float)rd;
                tshift[j] = tf;
                const double r = rd - (double)tf;
                e2 = fma(r, r, e2);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) e2 += __shfl_xor_sync(0xffffffffu, e2, o);
            if (lane == 0) part[warp] = e2;
            __syncthreads();
            if (threadIdx.x == 0) {
                double t = 0.0;
                for (int w = 0; w < 8; w++) t += part[w];
                *tshift_err = sqrt(t) * (1.0 + 1e-9);
            }
        }
    }
    const uint32_t j = blockIdx.x * 8 + warp;
    if (j >= slot) return;
    double s = 0.0;
    for (uint32_t f = lane; f < d; f += 32) {
        const double r = (double)seedrow[f] - (double)seedtab[(size_t)j * d + f];
        s = fma(r, r, s);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) skiptab[j] = 0.25 * s * (1.0 - margin);
}

// rank total = fixed-order reduction of the block sums (one CTA)
__global__ void __launch_bounds__(1024) kpp_total_kernel(const double* __restrict__ blocksum, uint32_t nb,
                                                         double* __restrict__ totals, int rank) {
    __shared__ double sh[1024];
    const uint32_t per = (nb + 1023) / 1024;
    const uint32_t b0 = threadIdx.x * per, b1 = min(nb, b0 + per);
    double s = 0.0;
    for (uint32_t b = b0; b < b1; b++) s = __dadd_rn(s, blocksum[b]);
    sh[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x < 32) {  // 32 lanes x 32 chunk sums sequentially, then lane 0 sequentially
        double t = 0.0;
        for (int i = 0; i < 32; i++) t = __dadd_rn(t, sh[threadIdx.x * 32 + i]);
        sh[threadIdx.x * 32] = t;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int i = 0; i < 32; i++) t = __dadd_rn(t, sh[i * 32]);
        totals[rank] = t;
    }
}

// Select the next seed row (kmeans.rs:385-396).  Every rank runs it with the same (all-gathered)
// totals; only the owner walks its D^2 array: thread-chunk sums -> block -> row, always as a running
// `cost += ...; if cost >= cutoff break` like the reference, at three granularities.
// Publishes [row elements | global index] into seedbuf on the owner and zeros elsewhere.
template <typename T>
__global__ void __launch_bounds__(1024)
kpp_select_kernel(const T* __restrict__ x, uint64_t n, uint32_t d, uint64_t row_offset, uint64_t n_global,
                  const double* __restrict__ mind, const double* __restrict__ blocksum, uint32_t nb,
                  const double* __restrict__ totals, int nranks, int rank, double u, long long inject_row,
                  unsigned char* __restrict__ seedbuf, uint32_t seed_words, long long* __restrict__ seeds,
                  uint32_t slot) {
    __shared__ long long s_local;  // local row chosen on this rank, or -1
    unsigned long long* out = reinterpret_cast<unsigned long long*>(seedbuf);
    for (uint32_t w = threadIdx.x; w < seed_words; w += blockDim.x) out[w] = 0ull;
    if (threadIdx.x == 0) s_local = -1;
    __syncthreads();

    if (inject_row >= 0) {
        if (threadIdx.x == 0 && (uint64_t)inject_row >= row_offset && (uint64_t)inject_row < row_offset + n)
            s_local = inject_row - (long long)row_offset;
    } else {
        // global total and owner, sequential over ranks (same on every rank)
        double total = 0.0;
        for (int r = 0; r < nranks; r++) total = __dadd_rn(total, totals[r]);
        const double cutoff = __dmul_rn(u, total);
        double run = 0.0; int owner = -1;
        for (int r = 0; r < nranks; r++) {
            double nxt = __dadd_rn(run, totals[r]);
            if (nxt >= cutoff) { owner = r; break; }
            run = nxt;
        }
        if (owner < 0) {  // rounding left the running cost below the cutoff: clamp to the last row
            owner = nranks - 1;
            // the last rank with any rows owns the clamp
        }
        if (owner == rank && n > 0) {
            // first index whose running cost reaches the cutoff, at three granularities: 1024 thread chunks of block
            // sums -> the blocks of one chunk -> the 1024 rows of one block.  The running cost of every candidate is a
            // CTA-wide inclusive scan (fixed shape => deterministic), the pick the lowest index that reaches the cutoff.
            __shared__ double s_warp[32];
            __shared__ uint32_t s_first;
            __shared__ uint32_t s_block;
            __shared__ double s_cost;
            auto first_reaching = [&](double v, bool valid, double cost0, double* excl_out) -> uint32_t {
                // returns the lowest thread index t with cost0 + sum(v[0..t]) >= cutoff (1024 if none); every thread
                // receives the result; *excl_out (thread-local) = running cost BEFORE this thread's own value
                const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
                double inc = v;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const double up = __shfl_up_sync(0xffffffffu, inc, o);
                    if (lane >= o) inc = __dadd_rn(up, inc);
                }
                __syncthreads();
                if (lane == 31) s_warp[warp] = inc;
                if (threadIdx.x == 0) s_first = 1024u;
                __syncthreads();
                double wbase = cost0;
                for (int w = 0; w < warp; w++) wbase = __dadd_rn(wbase, s_warp[w]);
                const double incl = __dadd_rn(wbase, inc);
                const double prev = __shfl_up_sync(0xffffffffu, inc, 1);
                *excl_out = lane ? __dadd_rn(wbase, prev) : wbase;
                if (valid && incl >= cutoff) atomicMin(&s_first, threadIdx.x);
                __syncthreads();
                return s_first;
            };
            const uint32_t per = (nb + 1023) / 1024;
            const uint32_t b0 = min(nb, threadIdx.x * per), b1 = min(nb, b0 + per);
            double s = 0.0;
            for (uint32_t b = b0; b < b1; b++) s = __dadd_rn(s, blocksum[b]);
            double excl = 0.0;
            const uint32_t chunk = first_reaching(s, b0 < nb, run, &excl);
            if (threadIdx.x == (chunk < 1024u ? chunk : 0u)) {
                uint32_t blk = nb;                                  // not found -> clamp to the last row
                double cost = excl;
                if (chunk < 1024u) {
                    blk = b1 - 1;
                    for (uint32_t b = b0; b < b1; b++) {
                        const double nxt = __dadd_rn(cost, blocksum[b]);
                        if (nxt >= cutoff) { blk = b; break; }
                        cost = nxt;
                    }
                }
                s_block = blk; s_cost = cost;
            }
            __syncthreads();
            const uint32_t blk = s_block;
            if (blk >= nb) {
                if (threadIdx.x == 0) s_local = (long long)n - 1;
            } else {
                const uint64_t r0 = (uint64_t)blk * kKppBlockRows;
                const uint32_t cnt = (uint32_t)min((uint64_t)kKppBlockRows, n - r0);
                const double v = threadIdx.x < cnt ? mind[r0 + threadIdx.x] : 0.0;
                const uint32_t hit = first_reaching(v, threadIdx.x < cnt, s_cost, &excl);
                if (threadIdx.x == 0) s_local = (long long)(r0 + (hit < cnt ? hit : cnt - 1));
            }
        }
    }
    __syncthreads();
    const long long loc = s_local;
    if (loc >= 0) {
        T* orow = reinterpret_cast<T*>(seedbuf);
        for (uint32_t j = threadIdx.x; j < d; j += blockDim.x) orow[j] = x[(uint64_t)loc * d + j];
        if (threadIdx.x == 0) out[seed_words - 1] = (unsigned long long)(loc + (long long)row_offset);
    }
    if (threadIdx.x == 0) seeds[slot] = loc >= 0 ? loc + (long long)row_offset : 0;
    (void)n_global;
}

// ------------------------------------------------------------------------------------------
// K7 / direct Lloyd assignment: argmin_j squared_distance(row widened to f64, centroid_j), strict <,
// lowest index wins (kmeans.rs:334-347).  One lane = one row, rows staged per warp; centroids are
// streamed through shared memory in chunks and read as broadcasts.
// ------------------------------------------------------------------------------------------
constexpr int ASG_WARPS = 4;

template <typename T, bool VEC>
__global__ void __launch_bounds__(ASG_WARPS * 32)
assign_direct_kernel(const T* __restrict__ x, uint64_t n, uint32_t d, const double* __restrict__ centroids,
                     uint32_t k, uint32_t kc, uint32_t* __restrict__ labels, double* __restrict__ mind,
                     uint32_t pitch16, const LoopState* __restrict__ loop_st, uint32_t loop_it) {
    if (loop_done(loop_st, loop_it)) return;                              // the fit's stop rule already fired
    extern __shared__ __align__(16) unsigned char smem[];
    double* cbuf = reinterpret_cast<double*>(smem);                       // [kc][d]
    const size_t cbuf_bytes = ((size_t)kc * d * sizeof(double) + 15) / 16 * 16;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char* slab = smem + cbuf_bytes + (size_t)warp * 32 * pitch16 * 16;

    const uint64_t tile_rows = ASG_WARPS * 32;
    const uint64_t ntiles = (n + tile_rows - 1) / tile_rows;
    for (uint64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const uint64_t row0 = tile * tile_rows + (uint64_t)warp * 32;
        const uint32_t nrows = row0 < n ? (uint32_t)min((uint64_t)32, n - row0) : 0u;
        if (VEC && nrows) stage_rows_vec<T>(x, row0, nrows, d, slab, pitch16, lane);
        double best = DBL_MAX; uint32_t bi = 0;
        for (uint32_t c0 = 0; c0 < k; c0 += kc) {
            const uint32_t cn = min(kc, k - c0);
            __syncthreads();  // previous chunk fully consumed
            for (uint32_t e = threadIdx.x; e < cn * d; e += blockDim.x) cbuf[e] = centroids[(size_t)c0 * d + e];
            __syncthreads();
            if (lane < nrows) {
                for (uint32_t c = 0; c < cn; c++) {
                    const double* cr = cbuf + (size_t)c * d;
                    double dist = 0.0;
                    if (VEC) {
                        using V = typename Vec16<T>::type;
                        const V* xr = reinterpret_cast<const V*>(slab + (size_t)lane * pitch16 * 16);
                        const uint32_t nv = d / Vec16<T>::N;
                        for (uint32_t q = 0; q < nv; q++) {
                            V xv = xr[q];
                            const T* xe = reinterpret_cast<const T*>(&xv);
#pragma unroll
                            for (int e = 0; e < Vec16<T>::N; e++)
                                dist = __dadd_rn(dist, sqdiff((double)xe[e], cr[q * Vec16<T>::N + e]));
                        }
                    } else {
                        const T* xr = x + (row0 + lane) * d;
                        for (uint32_t j = 0; j < d; j++) dist = __dadd_rn(dist, sqdiff((double)xr[j], cr[j]));
                    }
                    if (dist < best) { best = dist; bi = c0 + c; }
                }
            }
        }
        if (lane < nrows) {
            labels[row0 + lane] = bi;
            if (mind) mind[row0 + lane] = best;
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------
// K4: deterministic per-label sums / counts / inertia.  CTA p owns a contiguous run of rows and a
// private partial [k*d | k | 1] in global memory (L2 resident); thread j owns feature j (j+blockDim, ...)
// and walks the rows in ascending order, so the summation order is fixed by (n, grid) alone.
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void update_partial_kernel(const T* __restrict__ x, uint64_t n, uint32_t d,
                                      const uint32_t* __restrict__ labels, const double* __restrict__ mind,
                                      uint32_t k, uint64_t rows_per_cta, double* __restrict__ partials,
                                      const LoopState* __restrict__ loop_st, uint32_t loop_it) {
    if (loop_done(loop_st, loop_it)) return;
    const size_t pk = (size_t)k * d + k + 1;
    double* part = partials + (size_t)blockIdx.x * ((pk + 15) / 16 * 16);
    const uint64_t r0 = (uint64_t)blockIdx.x * rows_per_cta;
    const uint64_t r1 = min(n, r0 + rows_per_cta);
    double inertia = 0.0;
    for (uint64_t r = r0; r < r1; r++) {
        const uint32_t lbl = labels[r];
        double* prow = part + (size_t)lbl * d;
        const T* xr = x + r * d;
        for (uint32_t j = threadIdx.x; j < d; j += blockDim.x) prow[j] = __dadd_rn(prow[j], (double)xr[j]);
        if (threadIdx.x == 0) {
            double* pc = part + (size_t)k * d + lbl;
            *pc = *pc + 1.0;
            if (mind) inertia = __dadd_rn(inertia, mind[r]);
        }
    }
    if (threadIdx.x == 0) part[pk - 1] = inertia;
}

// ---- one-shot all-reduce over peer memory inside reduce_partials_kernel / finalize_kernel (protocol: sckm_peer.cu) ----
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void st_cell(uint4* cell, double v, uint32_t tag) {      // two 8-byte halves {data32, tag}
    asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(cell), "r"((uint32_t)__double2loint(v)), "r"(tag),
                 "r"((uint32_t)__double2hiint(v)), "r"(tag) : "memory");
}
__device__ __forceinline__ uint4 ld_cell(const uint4* cell) {
    uint4 v;
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(cell) : "memory");
    return v;
}
// this rank's element e of the current exchange -> every rank's receive area
__device__ __forceinline__ void peer_send(const PeerArgs& pa, size_t e, double v) {
    const size_t off = ((size_t)pa.half * pa.G + pa.rank) * pa.cap + e;
    for (uint32_t r = 0; r < pa.G; r++) st_cell(pa.recv[r] + off, v, pa.tag);
}
// element e of the all-reduced vector: the G received copies added in rank order (identical on every rank)
__device__ __forceinline__ double peer_sum(const PeerArgs& pa, size_t e) {
    const uint4* mine = pa.recv[pa.rank] + (size_t)pa.half * pa.G * pa.cap + e;
    double s = 0.0;
    for (uint32_t r0 = 0; r0 < pa.G; r0 += 8) {
        uint4 v[8];
#pragma unroll
        for (uint32_t i = 0; i < 8; i++) if (r0 + i < pa.G) v[i] = ld_cell(mine + (size_t)(r0 + i) * pa.cap);
#pragma unroll
        for (uint32_t i = 0; i < 8; i++) {
            if (r0 + i >= pa.G) break;
            if (v[i].y != pa.tag || v[i].w != pa.tag) {                   // still in flight: wait for it (30 s: the rank is gone)
                const unsigned long long t0 = global_ns();
                do {
                    __nanosleep(20);
                    v[i] = ld_cell(mine + (size_t)(r0 + i) * pa.cap);
                    if (global_ns() - t0 > 30000000000ull) { atomicExch(pa.err, 1u); break; }
                } while (v[i].y != pa.tag || v[i].w != pa.tag);
            }
            s = __dadd_rn(s, __hiloint2double((int)v[i].z, (int)v[i].x));
        }
    }
    return s;
}

// packed[e] = sum over the partial slots in a FIXED order (8 slot groups summed sequentially by 8 thread rows,
// then the 8 group sums in order); the slots are zeroed for the next step.  blockDim = (32, 8); each thread owns
// two adjacent elements (16-byte accesses; slot rows are 128-byte aligned).
template <int GROUPS>
__global__ void __launch_bounds__(32 * GROUPS) reduce_partials_kernel(double* __restrict__ partials, uint32_t nslots, size_t pk,
                                                                     size_t pitch, double* __restrict__ packed,
                                                                     unsigned long long* __restrict__ nmarked,
                                                                     const LoopState* __restrict__ loop_st, uint32_t loop_it, const PeerArgs pa) {
    pdl_wait();
    if (loop_done(loop_st, loop_it)) return;
    __shared__ double2 sh[GROUPS][33];
    if (blockIdx.x == 0 && threadIdx.x == 0 && threadIdx.y == 0) *nmarked = 0ull;   // consumed by the refine pass of this step
    const size_t e = ((size_t)blockIdx.x * 32 + threadIdx.x) * 2;
    const uint32_t per = (nslots + GROUPS - 1) / GROUPS;
    const uint32_t p0 = min(nslots, threadIdx.y * per), p1 = min(nslots, p0 + per);
    double2 s = make_double2(0.0, 0.0);
    const double2 zero = make_double2(0.0, 0.0);
    if (e < pitch) {
        double2* q = reinterpret_cast<double2*>(partials + (size_t)p0 * pitch + e);
        const size_t step = pitch / 2;
        uint32_t p = p0;
        for (; p + 8 <= p1; p += 8) {            // 8 independent loads in flight, added in slot order
            double2 v[8];
#pragma unroll
            for (int i = 0; i < 8; i++) v[i] = __ldcg(q + (size_t)i * step);
#pragma unroll
            for (int i = 0; i < 8; i++) {
                s.x = __dadd_rn(s.x, v[i].x); s.y = __dadd_rn(s.y, v[i].y);
                __stcg(q + (size_t)i * step, zero);
            }
            q += 8 * step;
        }
        for (; p < p1; p++) {
            const double2 v = __ldcg(q);
            s.x = __dadd_rn(s.x, v.x); s.y = __dadd_rn(s.y, v.y);
            __stcg(q, zero); q += step;
        }
    }
    sh[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.y == 0) {
        double2 t = make_double2(0.0, 0.0);
#pragma unroll
        for (int i = 0; i < GROUPS; i++) { t.x = __dadd_rn(t.x, sh[i][threadIdx.x].x); t.y = __dadd_rn(t.y, sh[i][threadIdx.x].y); }
        if (pa.G != 0) {                          // multi-GPU loop: straight into every rank's receive area (sckm_peer.cu)
            if (e < pk) peer_send(pa, e, t.x);
            if (e + 1 < pk) peer_send(pa, e + 1, t.y);
        } else {
            if (e < pk) packed[e] = t.x;
            if (e + 1 < pk) packed[e + 1] = t.y;
        }
    }
}

// K6: centroids = sums / counts (kmeans.rs:288-292 unguarded, :297-303 guarded), sizes, ||c||^2.
// Inside a Lloyd loop (loop_st != nullptr) the kernel also applies the reference's stop rule to the device-resident
// state (kmeans.rs:305-309: compare, then store; the centroids of the breaking iteration ARE updated): CTA 0 records
// `done_at = it` when `distortion <= dist`, else `distortion = dist`.  The CTAs of this launch test `done_at < it`, so
// the write cannot hide work of the iteration that sets it; every kernel of a later iteration returns at once.
// centered != 0: the packed sums are sums of (x - mu) (tile kernel, see sckm_dmma.cu): centroid = sum / count + mu.
// The norms are always ||c - mu||^2 for the shift currently in `mu` (zeros when nothing is centred).
// pa.G != 0 (multi-GPU loop): `packed` is not read; the kernel first sums the ranks' vectors over peer memory
// (peer_sum: the cells reduce_partials_kernel of every rank sent here) and leaves the all-reduced vector in pa.packed_out.
__global__ void finalize_kernel(const double* __restrict__ packed, uint32_t k, uint32_t d, int guarded, int centered,
                                const double* __restrict__ mu, double* __restrict__ centroids, double* __restrict__ cnorm,
                                long long* __restrict__ size, LoopState* __restrict__ loop_st, uint32_t loop_it,
                                double* __restrict__ inertia_trace, const PeerArgs pa) {
    pdl_wait();
    if (loop_done(loop_st, loop_it)) return;
    const uint32_t c = blockIdx.x;
    const bool peers = pa.G != 0;
    const double cnt = peers ? peer_sum(pa, (size_t)k * d + c) : packed[(size_t)k * d + c];
    if (threadIdx.x == 0) {
        size[c] = (long long)cnt;
        if (peers) pa.packed_out[(size_t)k * d + c] = cnt;
    }
    for (uint32_t j = threadIdx.x; j < d; j += blockDim.x) {
        double sum;
        if (peers) { sum = peer_sum(pa, (size_t)c * d + j); pa.packed_out[(size_t)c * d + j] = sum; }
        else if (!guarded || cnt > 0.0) sum = packed[(size_t)c * d + j];
        else continue;
        if (!guarded || cnt > 0.0) {
            const double mean = __ddiv_rn(sum, cnt);
            centroids[(size_t)c * d + j] = centered ? __dadd_rn(mean, mu[j]) : mean;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0 && cnorm) {
        double s = 0.0;
        for (uint32_t j = 0; j < d; j++) { double v = centroids[(size_t)c * d + j] - mu[j]; s = fma(v, v, s); }
        cnorm[c] = s;
    }
    if (loop_st != nullptr && c == 0 && threadIdx.x == 0) {
        const double dist = peers ? peer_sum(pa, (size_t)k * d + k) : packed[(size_t)k * d + k];
        if (peers) pa.packed_out[(size_t)k * d + k] = dist;
        if (inertia_trace) inertia_trace[loop_it - 1] = dist;
        loop_st->iters = loop_it;
        if (loop_st->honor_stop) {
            if (loop_st->distortion <= dist) loop_st->done_at = loop_it;      // break (kmeans.rs:305-306)
            else loop_st->distortion = dist;                                  // kmeans.rs:307-308
        }
    }
}

__global__ void loop_init_kernel(LoopState* st, int honor_stop) {
    st->distortion = DBL_MAX; st->done_at = 0ull; st->iters = 0ull; st->honor_stop = honor_stop ? 1ull : 0ull;
}

// ------------------------------------------------------------------------------------------
// layout / generation / misc
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void transpose_kernel(const T* __restrict__ src, T* __restrict__ dst, uint64_t n, uint64_t d) {
    // src: column-major (d columns of n), dst: row-major n x d; 32x32 tiles through shared memory
    __shared__ T tile[32][33];
    const uint64_t r0 = (uint64_t)blockIdx.x * 32, c0 = (uint64_t)blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        uint64_t c = c0 + i, r = r0 + threadIdx.x;
        if (c < d && r < n) tile[i][threadIdx.x] = src[c * n + r];
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        uint64_t r = r0 + i, c = c0 + threadIdx.x;
        if (r < n && c < d) dst[r * d + c] = tile[threadIdx.x][i];
    }
}

template <typename T>
__global__ void blobs_kernel(T* __restrict__ x, uint64_t row0, uint64_t nrows, uint64_t d, uint64_t n_centers,
                             uint64_t seed) {
    const uint64_t total = nrows * d;
    for (uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t r = e / d, c = e - r * d;
        x[e] = (T)blob_value(seed, n_centers, row0 + r, c);
    }
}

__global__ void widen_labels_kernel(const uint32_t* __restrict__ in, unsigned long long* __restrict__ out, uint64_t n) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        out[i] = in[i];
}

// ------------------------------------------------------------------------------------------
// peak micro-kernels (roofline denominators measured on this very device)
// ------------------------------------------------------------------------------------------
__global__ void copy_kernel(const double2* __restrict__ a, double2* __restrict__ b, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) b[i] = a[i];
}
__global__ void __launch_bounds__(256) dfma_peak_kernel(double* out, int iters) {
    double a[8];
    const double m = 1.0000001, c = 1e-9 * threadIdx.x;
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) a[i] = fma(a[i], m, c);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += a[i];
    if (s == 12345.678) out[0] = s;
}
__global__ void __launch_bounds__(256) dmma_peak_kernel(double* out, int iters) {
    double a = threadIdx.x * 1e-3, b = threadIdx.x * 2e-3;
    double c[8][2];
#pragma unroll
    for (int i = 0; i < 8; i++) { c[i][0] = i; c[i][1] = -i; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += c[i][0] + c[i][1];
    if (s == 12345.678) out[0] = s;
}

// ------------------------------------------------------------------------------------------
// host side: error text, workspaces, launchers
// ------------------------------------------------------------------------------------------
static thread_local std::string g_create_err;
const char* create_error_text() { return g_create_err.c_str(); }

int fail(sckm_ctx* ctx, int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    if (ctx) ctx->err = buf; else g_create_err = buf;
    return code;
}

#define LAUNCH_CHECK(ctx)                                                                          \
    do {                                                                                           \
        (ctx)->launches++;                                                                         \
        cudaError_t _e = cudaGetLastError();                                                       \
        if (_e != cudaSuccess)                                                                     \
            return fail((ctx), SCKM_ERR_CUDA, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), \
                        __FILE__, __LINE__);                                                       \
    } while (0)

template <typename P> static int regrow(sckm_ctx* ctx, P** p, size_t* cap, size_t need_elems, size_t elem) {
    if (need_elems <= *cap && *p) return SCKM_OK;
    if (*p) { SCKM_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); SCKM_CUDA(ctx, cudaFree(*p)); *p = nullptr; }
    SCKM_CUDA(ctx, ws_malloc(ctx, (void**)p, need_elems * elem));
    SCKM_CUDA(ctx, cudaMemsetAsync(*p, 0, need_elems * elem, ctx->stream));
    *cap = need_elems;
    return SCKM_OK;
}

int ensure_workspace(sckm_ctx* ctx, uint64_t k, uint64_t d, size_t partial_slots) {
    const size_t kd = (size_t)k * d, pk = kd + k + 1;
    if (k != ctx->ws_k || d != ctx->ws_d) { ctx->cnorm_valid = false; ctx->ws_k = k; ctx->ws_d = d; }
    SCKM_TRY(regrow(ctx, &ctx->d_centroids, &ctx->cap_centroids, kd, sizeof(double)));
    SCKM_TRY(regrow(ctx, &ctx->d_packed, &ctx->cap_packed, pk, sizeof(double)));
    SCKM_TRY(regrow(ctx, &ctx->d_cnorm, &ctx->cap_cnorm, k + 1, sizeof(double)));  // [k] norms + max
    if (d > ctx->cap_mu || !ctx->d_mu) { SCKM_TRY(regrow(ctx, &ctx->d_mu, &ctx->cap_mu, d, sizeof(double))); ctx->mu_zero = true; ctx->cnorm_valid = false; }
    SCKM_TRY(regrow(ctx, &ctx->d_size, &ctx->cap_size, k, sizeof(int64_t)));
    SCKM_TRY(regrow(ctx, &ctx->d_seeds, &ctx->cap_seeds, k, sizeof(int64_t)));
    if (partial_slots)
        SCKM_TRY(regrow(ctx, &ctx->d_partials, &ctx->cap_partials, partial_slots * slot_pitch(pk), sizeof(double)));
    return SCKM_OK;
}

static bool vec_ok(uint64_t d, int dtype) { return (d * (dtype == SCKM_F32 ? 4 : 8)) % 16 == 0; }

int launch_transpose(sckm_ctx* ctx, const void* src, void* dst, uint64_t n, uint64_t d, int dtype) {
    dim3 grid((unsigned)((n + 31) / 32), (unsigned)((d + 31) / 32)), block(32, 8);
    if (dtype == SCKM_F32) transpose_kernel<float><<<grid, block, 0, ctx->stream>>>((const float*)src, (float*)dst, n, d);
    else transpose_kernel<double><<<grid, block, 0, ctx->stream>>>((const double*)src, (double*)dst, n, d);
    LAUNCH_CHECK(ctx);
    return SCKM_OK;
}

int launch_blobs(sckm_ctx* ctx, void* x, int dtype, uint64_t row0, uint64_t nrows, uint64_t d,
                 uint64_t n_centers, uint64_t seed) {
    const int grid = ctx->num_sms * 8;
    if (dtype == SCKM_F32) blobs_kernel<float><<<grid, 256, 0, ctx->stream>>>((float*)x, row0, nrows, d, n_centers, seed);
    else blobs_kernel<double><<<grid, 256, 0, ctx->stream>>>((double*)x, row0, nrows, d, n_centers, seed);
    LAUNCH_CHECK(ctx);
    return SCKM_OK;
}

template <typename K> static int set_smem(sckm_ctx* ctx, K kernel, size_t bytes) {
    if (bytes > 48 * 1024) SCKM_CUDA(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return SCKM_OK;
}

template <typename T>
static int kpp_refresh_t(sckm_dataset* ds, uint32_t label, bool first_pass, bool prune) {
    sckm_ctx* ctx = ds->ctx;
    const uint32_t d = (uint32_t)ds->d;
    const uint32_t nb = (uint32_t)((ds->n + kKppBlockRows - 1) / kKppBlockRows);
    const uint32_t row_bytes = d * sizeof(T), pitch16 = slab_pitch16(row_bytes);
    const size_t cent_bytes = ((size_t)row_bytes + 15) / 16 * 16;
    const uint32_t ntab = prune ? label : 0;                      // seeds 0 .. label-1 precede the new one
    const size_t tab_bytes = (size_t)ntab * sizeof(double);
    const size_t smem_vec = cent_bytes + (size_t)KPP_WARPS * 32 * pitch16 * 16 + tab_bytes;
    const bool vec = vec_ok(d, ds->dtype) && smem_vec <= (size_t)ctx->smem_optin;
    const double* tab = prune && ntab ? ctx->d_skiptab : nullptr;
    if (nb) {
        if (vec) {
            SCKM_TRY(set_smem(ctx, kpp_refresh_kernel<T, true>, smem_vec));
            kpp_refresh_kernel<T, true><<<nb, KPP_WARPS * 32, smem_vec, ctx->stream>>>(
                (const T*)ds->x, ds->n, d, (const T*)ctx->d_seedrow, ds->mind, ds->labels, label, first_pass ? 1 : 0,
                ctx->d_blocksum, pitch16, tab, ntab);
        } else {
            if (cent_bytes + tab_bytes > (size_t)ctx->smem_optin) return fail(ctx, SCKM_ERR_INVALID, "d=%u too large", d);
            SCKM_TRY(set_smem(ctx, kpp_refresh_kernel<T, false>, cent_bytes + tab_bytes));
            kpp_refresh_kernel<T, false><<<nb, KPP_WARPS * 32, cent_bytes + tab_bytes, ctx->stream>>>(
                (const T*)ds->x, ds->n, d, (const T*)ctx->d_seedrow, ds->mind, ds->labels, label, first_pass ? 1 : 0,
                ctx->d_blocksum, pitch16, tab, ntab);
        }
        LAUNCH_CHECK(ctx);
    }
    kpp_total_kernel<<<1, 1024, 0, ctx->stream>>>(ctx->d_blocksum, nb, ctx->d_totals, ctx->rank);
    LAUNCH_CHECK(ctx);
    return SCKM_OK;
}

template <typename T>
static int kpp_pass_t(sckm_dataset* ds, uint32_t label, bool first_pass, bool prune, bool want_sums) {
    sckm_ctx* ctx = ds->ctx;
    const uint32_t d = (uint32_t)ds->d;
    const uint32_t row_bytes = d * sizeof(T), pitch16 = slab_pitch16(row_bytes);
    const size_t cent_bytes = ((size_t)row_bytes + 15) / 16 * 16;
    const size_t smem = cent_bytes + (size_t)KPP_WARPS * 32 * pitch16 * 16;
    if (!vec_ok(d, ds->dtype) || smem > (size_t)ctx->smem_optin || ds->n >= 0xFFFFFFFFull || !ctx->d_surv ||
        getenv("SCKM_KPP_GEN1"))
        return kpp_refresh_t<T>(ds, label, first_pass, prune);          // first-generation single kernel
    const uint32_t nb = (uint32_t)((ds->n + kKppBlockRows - 1) / kKppBlockRows);
    if (ds->n) {
        const uint32_t ntab = prune && !first_pass ? label : 0;
        const uint32_t* surv = nullptr;
        const bool screen = ds->kpp_shadow != nullptr;
        if (ntab) {
            SCKM_CUDA(ctx, cudaMemsetAsync(ctx->d_kppctr, 0, sizeof(unsigned), ctx->stream));
            const unsigned grid = (unsigned)((ds->n + KPPA_THREADS * KPPA_ITERS - 1) / (KPPA_THREADS * KPPA_ITERS));
            const size_t psmem = (size_t)ntab * sizeof(double) + (screen ? (size_t)d * sizeof(float) : 0);
            const uint32_t nv8 = screen ? d / 8 : 0;
            auto prune_kern = nv8 == 1 ? kpp_prune_kernel<1> : nv8 == 2 ? kpp_prune_kernel<2> : nv8 == 4 ? kpp_prune_kernel<4>
                            : nv8 == 8 ? kpp_prune_kernel<8> : nv8 == 16 ? kpp_prune_kernel<16> : kpp_prune_kernel<0>;
            SCKM_TRY(set_smem(ctx, prune_kern, psmem));
            // rounding of the reference's own distance: (d+2) ulp of TX arithmetic, with two orders of magnitude to spare
            const double margin = (ds->dtype == SCKM_F32 ? 1.2e-7 : 2.3e-16) * 100.0 * (double)(d + 2);
            prune_kern<<<grid, KPPA_THREADS, psmem, ctx->stream>>>(ds->mind, ds->labels, ds->n, ctx->d_skiptab, ntab,
                ctx->d_surv, ctx->d_kppctr, screen ? ds->kpp_shadow : nullptr, ds->kpp_shadow_err, ctx->d_tshift,
                ctx->d_tshift_err, d, margin);
            LAUNCH_CHECK(ctx);
            surv = ctx->d_surv;
        }
        static thread_local size_t cached_smem = 0;              // one attribute + occupancy query per shape, not per pass
        static thread_local int per_sm = 0, cached_dev = -1;     // (function attributes are per device)
        if (cached_smem != smem || per_sm == 0 || cached_dev != ctx->device) {
            cached_dev = ctx->device;
            SCKM_TRY(set_smem(ctx, kpp_compute_kernel<T>, smem));
            SCKM_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kpp_compute_kernel<T>, KPP_WARPS * 32, smem));
            cached_smem = smem;
        }
        const uint64_t groups = (ds->n + KPP_WARPS * 32 - 1) / (KPP_WARPS * 32);
        const unsigned grid = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>(groups, (uint64_t)ctx->num_sms * std::max(per_sm, 1)));
        kpp_compute_kernel<T><<<grid, KPP_WARPS * 32, smem, ctx->stream>>>((const T*)ds->x, ds->n, d, (const T*)ctx->d_seedrow,
            ds->mind, ds->labels, label, first_pass ? 1 : 0, surv, ctx->d_kppctr, pitch16,
            first_pass ? ds->kpp_shadow : nullptr, ds->kpp_shadow_err);
        LAUNCH_CHECK(ctx);
    }
    if (want_sums) {
        if (nb) {
            kpp_blocksum_kernel<<<(nb + 7) / 8, 256, 0, ctx->stream>>>(ds->mind, ds->n, nb, ctx->d_blocksum, ctx->d_totals,
                                                                       ctx->rank, ctx->d_kppctr + 1);
        } else {
            kpp_total_kernel<<<1, 1024, 0, ctx->stream>>>(ctx->d_blocksum, 0, ctx->d_totals, ctx->rank);
        }
        LAUNCH_CHECK(ctx);
    }
    return SCKM_OK;
}

int launch_kpp_refresh(sckm_dataset* ds, uint32_t label, bool first_pass, bool prune, bool want_sums) {
    return ds->dtype == SCKM_F32 ? kpp_pass_t<float>(ds, label, first_pass, prune, want_sums)
                                 : kpp_pass_t<double>(ds, label, first_pass, prune, want_sums);
}

// remember the seed just published (slot) and build the pruning table against all earlier seeds
int launch_kpp_seedtab(sckm_dataset* ds, uint32_t slot) {
    sckm_ctx* ctx = ds->ctx;
    const unsigned grid = (unsigned)std::max<uint32_t>(1, (slot + 7) / 8);
    // margin: the bound is exact in real arithmetic; computed D^2 values carry <= d*eps relative rounding
    // (eps = 2^-24 for f32 element arithmetic, 2^-53 for f64), so leave orders of magnitude of slack
    if (ds->dtype == SCKM_F32)
        kpp_seedtab_kernel<float><<<grid, 256, 0, ctx->stream>>>((const float*)ctx->d_seedrow, (float*)ctx->d_seedtab,
                                                              (uint32_t)ds->d, slot, ctx->d_skiptab, 1e-3,
                                                              ds->kpp_shadow ? ctx->d_tshift : nullptr, ctx->d_tshift_err);
    else
        kpp_seedtab_kernel<double><<<grid, 256, 0, ctx->stream>>>((const double*)ctx->d_seedrow, (double*)ctx->d_seedtab,
                                                               (uint32_t)ds->d, slot, ctx->d_skiptab, 1e-9,
                                                               ds->kpp_shadow ? ctx->d_tshift : nullptr, ctx->d_tshift_err);
    LAUNCH_CHECK(ctx);
    return SCKM_OK;
}

int launch_kpp_select(sckm_dataset* ds, double u, int64_t inject_row, uint32_t slot) {
    sckm_ctx* ctx = ds->ctx;
    const uint32_t nb = (uint32_t)((ds->n + kKppBlockRows - 1) / kKppBlockRows);
    const uint32_t seed_words = (uint32_t)(ctx->cap_seedrow / 8);
    if (ds->dtype == SCKM_F32)
        kpp_select_kernel<float><<<1, 1024, 0, ctx->stream>>>((const float*)ds->x, ds->n, (uint32_t)ds->d, ds->row_offset,
            ds->n_global, ds->mind, ctx->d_blocksum, nb, ctx->d_totals, ctx->nranks, ctx->rank, u, (long long)inject_row,
            (unsigned char*)ctx->d_seedrow, seed_words, (long long*)ctx->d_seeds, slot);
    else
        kpp_select_kernel<double><<<1, 1024, 0, ctx->stream>>>((const double*)ds->x, ds->n, (uint32_t)ds->d, ds->row_offset,
            ds->n_global, ds->mind, ctx->d_blocksum, nb, ctx->d_totals, ctx->nranks, ctx->rank, u, (long long)inject_row,
            (unsigned char*)ctx->d_seedrow, seed_words, (long long*)ctx->d_seeds, slot);
    LAUNCH_CHECK(ctx);
    return SCKM_OK;
}

template <typename T>
static int assign_direct_t(sckm_ctx* ctx, const void* x, int dtype, uint64_t n, uint64_t d64, uint64_t k64,
                           uint32_t* labels, double* mind) {
    if (n == 0) return SCKM_OK;
    const uint32_t d = (uint32_t)d64, k = (uint32_t)k64;
    const uint32_t row_bytes = d * sizeof(T), pitch16 = slab_pitch16(row_bytes);
    // centroid chunk streamed through shared memory: <= 16 KB, at least one centroid
    const uint32_t kc = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(k, (16 * 1024) / ((uint64_t)d * sizeof(double))));
    const size_t cbuf_bytes = ((size_t)kc * d * sizeof(double) + 15) / 16 * 16;
    const size_t smem_vec = cbuf_bytes + (size_t)ASG_WARPS * 32 * pitch16 * 16;
    const bool vec = vec_ok(d, dtype) && smem_vec <= (size_t)ctx->smem_optin;
    const uint64_t ntiles = (n + ASG_WARPS * 32 - 1) / (ASG_WARPS * 32);
    const unsigned grid = (unsigned)std::min<uint64_t>(ntiles, (uint64_t)ctx->num_sms * 16);
    if (vec) {
        SCKM_TRY(set_smem(ctx, assign_direct_kernel<T, true>, smem_vec));
        assign_direct_kernel<T, true><<<grid, ASG_WARPS * 32, smem_vec, ctx->stream>>>(
            (const T*)x, n, d, ctx->d_centroids, k, kc, labels, mind, pitch16, SCKM_LOOP_ARGS(ctx));
    } else {
        if (cbuf_bytes > (size_t)ctx->smem_optin) return fail(ctx, SCKM_ERR_INVALID, "d=%u too large", d);
        SCKM_TRY(set_smem(ctx, assign_direct_kernel<T, false>, cbuf_bytes));
        assign_direct_kernel<T, false><<<grid, ASG_WARPS * 32, cbuf_bytes, ctx->stream>>>(
            (const T*)x, n, d, ctx->d_centroids, k, kc, labels, mind, pitch16, SCKM_LOOP_ARGS(ctx));
    }
    LAUNCH_CHECK(ctx);
    return SCKM_OK;
}

int launch_assign_direct_raw(sckm_ctx* ctx, const void* x, int dtype, uint64_t n, uint64_t d, uint64_t k,
                             uint32_t* labels, double* mind) {
    return dtype == SCKM_F32 ? assign_direct_t<float>(ctx, x, dtype, n, d, k, labels, mind)
                             : assign_direct_t<double>(ctx, x, dtype, n, d, k, labels, mind);
}

int launch_assign_direct(sckm_dataset* ds, uint64_t k) {
    return launch_assign_direct_raw(ds->ctx, ds->x, ds->dtype, ds->n, ds->d, k, ds->labels, ds->mind);
}

bool update_given_supported(const sckm_dataset* ds, uint64_t k);   // sckm_dmma.cu
int launch_update_given(sckm_dataset* ds, uint64_t k);             // sckm_dmma.cu

static uint32_t update_slots(const sckm_ctx* ctx, uint64_t n) {
    uint64_t p = (n + 63) / 64;
    uint64_t cap = (uint64_t)ctx->num_sms * 16;
    return (uint32_t)std::max<uint64_t>(1, std::min(p, cap));
}

int launch_update(sckm_dataset* ds, uint64_t k, bool with_inertia) {
    sckm_ctx* ctx = ds->ctx;
    ctx->packed_centered = false;                                // plain sums of x on every path below
    const uint32_t slots = update_slots(ctx, ds->n);
    SCKM_TRY(ensure_workspace(ctx, k, ds->d, slots));
    const size_t pk = (size_t)k * ds->d + k + 1;
    if (!with_inertia && ds->n >= 65536 && update_given_supported(ds, k))
        return launch_update_given(ds, k);                       // HBM-rate path for the initial means of a large fit
    const uint64_t rows_per_cta = (ds->n + slots - 1) / slots;
    const unsigned threads = (unsigned)std::min<uint64_t>(256, (ds->d + 31) / 32 * 32);
    if (ds->n) {
        if (ds->dtype == SCKM_F32)
            update_partial_kernel<float><<<slots, threads, 0, ctx->stream>>>((const float*)ds->x, ds->n, (uint32_t)ds->d,
                ds->labels, with_inertia ? ds->mind : nullptr, (uint32_t)k, rows_per_cta, ctx->d_partials, SCKM_LOOP_ARGS(ctx));
        else
            update_partial_kernel<double><<<slots, threads, 0, ctx->stream>>>((const double*)ds->x, ds->n, (uint32_t)ds->d,
                ds->labels, with_inertia ? ds->mind : nullptr, (uint32_t)k, rows_per_cta, ctx->d_partials, SCKM_LOOP_ARGS(ctx));
        LAUNCH_CHECK(ctx);
    }
    SCKM_TRY(launch_reduce_partials(ctx, slots, pk));
    return SCKM_OK;
}

int launch_reduce_partials(sckm_ctx* ctx, uint32_t slots, size_t pk) {
    const size_t pitch = slot_pitch(pk);
    const unsigned blocks = (unsigned)((pitch / 2 + 31) / 32);
    PeerArgs pa;
    memset(&pa, 0, sizeof(pa));
    if (ctx->peer_step) pa = peer_args(ctx);           // multi-GPU loop: the vector goes to every rank's receive area instead
    // small payload, or many slots to walk: more slot groups per block so that enough loads are in flight
    if (blocks < (unsigned)ctx->num_sms || slots >= 256)
        SCKM_CUDA(ctx, launch_pdl(reduce_partials_kernel<32>, dim3(blocks), dim3(32, 32), 0, ctx->stream, ctx->d_partials, slots, pk, pitch,
                                  ctx->d_packed, ctx->d_flags, SCKM_LOOP_ARGS(ctx), pa));
    else
        SCKM_CUDA(ctx, launch_pdl(reduce_partials_kernel<8>, dim3(blocks), dim3(32, 8), 0, ctx->stream, ctx->d_partials, slots, pk, pitch,
                                  ctx->d_packed, ctx->d_flags, SCKM_LOOP_ARGS(ctx), pa));
    LAUNCH_CHECK(ctx);
    return SCKM_OK;
}

int launch_finalize(sckm_ctx* ctx, uint64_t k, uint64_t d, bool guarded) {
    const unsigned threads = (unsigned)std::min<uint64_t>(256, (d + 31) / 32 * 32);
    PeerArgs pa;
    memset(&pa, 0, sizeof(pa));
    if (ctx->peer_step) pa = peer_args(ctx);           // the ranks' vectors sit in this rank's receive area: sum them here
    SCKM_CUDA(ctx, launch_pdl(finalize_kernel, dim3((unsigned)k), dim3(threads), 0, ctx->stream, (const double*)ctx->d_packed, (uint32_t)k,
                              (uint32_t)d, guarded ? 1 : 0, ctx->packed_centered ? 1 : 0, (const double*)ctx->d_mu, ctx->d_centroids,
                              ctx->d_cnorm, (long long*)ctx->d_size, SCKM_LOOP_ARGS(ctx), ctx->loop_it ? ctx->d_inertia_trace : (double*)nullptr, pa));
    LAUNCH_CHECK(ctx);
    ctx->cnorm_valid = ctx->cnorm_valid || !guarded;   // a guarded update keeps stale norms of empty clusters stale
    return SCKM_OK;
}

// start a Lloyd loop on the device: distortion = f64::MAX, nothing done (kmeans.rs:273); room for `max_iter` inertias
int launch_loop_init(sckm_ctx* ctx, uint64_t max_iter, bool honor_stop) {
    if (!ctx->d_loop) SCKM_CUDA(ctx, cudaMalloc((void**)&ctx->d_loop, sizeof(LoopState)));
    if (max_iter > ctx->cap_trace) {
        if (ctx->d_inertia_trace) { SCKM_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); cudaFree(ctx->d_inertia_trace); ctx->d_inertia_trace = nullptr; }
        SCKM_CUDA(ctx, cudaMalloc((void**)&ctx->d_inertia_trace, max_iter * sizeof(double)));
        ctx->cap_trace = max_iter;
    }
    loop_init_kernel<<<1, 1, 0, ctx->stream>>>(ctx->d_loop, honor_stop ? 1 : 0);
    LAUNCH_CHECK(ctx);
    return SCKM_OK;
}

int launch_labels_widen(sckm_ctx* ctx, const uint32_t* in, uint64_t* out, uint64_t n) {
    if (!n) return SCKM_OK;
    widen_labels_kernel<<<ctx->num_sms * 4, 256, 0, ctx->stream>>>(in, (unsigned long long*)out, n);
    LAUNCH_CHECK(ctx);
    return SCKM_OK;
}

int measure_peaks(sckm_ctx* ctx, double* out3) {
    float ms = 0;
    // HBM copy: 1 GiB read + 1 GiB write, best of 5
    const size_t bytes = (size_t)1 << 30;
    void *a = nullptr, *b = nullptr;
    SCKM_CUDA(ctx, cudaMalloc(&a, bytes)); SCKM_CUDA(ctx, cudaMalloc(&b, bytes));
    SCKM_CUDA(ctx, cudaMemsetAsync(a, 1, bytes, ctx->stream));
    double best = 0;
    for (int it = 0; it < 6; it++) {
        SCKM_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
        copy_kernel<<<ctx->num_sms * 16, 512, 0, ctx->stream>>>((const double2*)a, (double2*)b, bytes / 16);
        ctx->launches++;
        SCKM_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
        SCKM_CUDA(ctx, cudaEventSynchronize(ctx->ev1));
        SCKM_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
        if (it) best = std::max(best, 2.0 * bytes / (ms * 1e-3) / 1e9);
    }
    out3[0] = best;
    cudaFree(a); cudaFree(b);
    const int iters = 4096;
    for (int which = 0; which < 2; which++) {
        best = 0;
        for (int it = 0; it < 4; it++) {
            SCKM_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
            if (which == 0) dfma_peak_kernel<<<ctx->num_sms * 8, 256, 0, ctx->stream>>>((double*)ctx->d_flags, iters);
            else dmma_peak_kernel<<<ctx->num_sms * 8, 256, 0, ctx->stream>>>((double*)ctx->d_flags, iters);
            ctx->launches++;
            SCKM_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
            SCKM_CUDA(ctx, cudaEventSynchronize(ctx->ev1));
            SCKM_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
            const double threads = (double)ctx->num_sms * 8 * 256;
            const double flops = which == 0 ? threads * iters * 8 * 2.0            // 8 DFMA / thread / iter
                                            : (threads / 32) * iters * 8 * 512.0;  // 8 m8n8k4 / warp / iter
            if (it) best = std::max(best, flops / (ms * 1e-3) / 1e12);
        }
        out3[1 + which] = best;
    }
    return SCKM_OK;
}

}  // namespace sckm
