// sckm_dmma.cu -- K1: Lloyd assignment as a fused  -(||x||^2 - 2 x.c + ||c||^2)/2  tile kernel on
// the FP64 tensor path (DMMA.8x8x4 via mma.sync.m8n8k4.f64; tcgen05 has no f64 kind), with the
// argmin kept in registers and an exact direct-form re-decision for near-ties.
//
// Replaces the per-iteration work of BBDTree::clustering (src/algorithm/neighbour/bbd_tree.rs:62-163):
// label_i = argmin_j ||x_i - c_j||^2, lowest index on exact ties (bbd_tree.rs:101-111).
//
// Mapping
//   * One CTA per SM (persistent, 8 warps).  The centroid block C[bn][d] lives in shared memory with
//     a row pitch of d*8+32 bytes so the 8x4 B-fragment reads are bank-conflict free; for
//     k*d small enough (config C3: 256 x 64 -> 136 KB) it is loaded once per launch.
//   * Each warp owns slabs of 8*MT rows.  The rows are loaded from HBM straight into DMMA
//     A-fragment registers (each LDG.64 warp request = 8 fully used 32-byte sectors), so X is read
//     exactly once per iteration and never staged; warps run free of CTA barriers while the centroid
//     block is resident, so one warp's HBM latency hides under the other warps' DMMAs.
//   * Accumulators start at -(||x||^2 + ||c||^2)/2, DMMA adds x.c: acc = -dist^2/2; the epilogue
//     is max / second-max tracking only.
//   * Rows whose best/second gap is within 1e-10*(||x||^2 + max||c||^2) (>= 1e4 x the rounding
//     error bound of the GEMM form) are appended to a list and re-decided by refine_rows_kernel with
//     the reference's exact arithmetic (euclidian.rs:56-63), so labels equal the dense oracle.
#include "sckm_common.cuh"
#include <cfloat>
#include <algorithm>

namespace sckm {

#define LAUNCH_CHECK_D(ctx)                                                                        \
    do {                                                                                           \
        (ctx)->launches++;                                                                         \
        cudaError_t _e = cudaGetLastError();                                                       \
        if (_e != cudaSuccess)                                                                     \
            return fail((ctx), SCKM_ERR_CUDA, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), \
                        __FILE__, __LINE__);                                                       \
    } while (0)

constexpr int DMMA_WARPS = 8;
constexpr int DMMA_NT = 8;                       // n-tiles per sub-block: 64 centroids
constexpr double DMMA_TIE_REL = 1e-10;

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// KSTEPS: d padded to 4*KSTEPS; MT: m-tiles (8 rows each) per warp slab
template <int KSTEPS, int MT, typename TX>
__global__ void __launch_bounds__(DMMA_WARPS * 32, 1)
assign_dmma_kernel(const TX* __restrict__ x, uint64_t n, uint32_t d, const double* __restrict__ centroids,
                   const double* __restrict__ cnorm, uint32_t k, uint32_t bn, uint32_t* __restrict__ labels,
                   double* __restrict__ mind, unsigned long long* __restrict__ flag_count,
                   uint32_t* __restrict__ flag_rows) {
    constexpr int DP = KSTEPS * 4;                 // padded feature count
    constexpr int PITCH = DP + 4;                  // doubles per staged centroid row (pitch = d*8+32 B)
    constexpr int ROWS = 8 * MT;
    extern __shared__ __align__(16) double smem_d[];
    double* cbuf = smem_d;                         // [bn][PITCH]
    double* cn = smem_d + (size_t)bn * PITCH;      // [bn]  ||c||^2, +inf for padding columns
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    const double cmax = cnorm[k];                  // max_j ||c_j||^2 (written by cnorm_max_kernel)

    const uint64_t nslabs = (n + ROWS - 1) / ROWS;
    const uint64_t stride = (uint64_t)gridDim.x * DMMA_WARPS;
    const uint64_t rounds = (nslabs + stride - 1) / stride;
    const uint32_t nchunks = (k + bn - 1) / bn;
    const double* bbase = cbuf + (size_t)g * PITCH + t;

    for (uint64_t rd = 0; rd < rounds; rd++) {
        const uint64_t slab = rd * stride + (uint64_t)blockIdx.x * DMMA_WARPS + warp;
        const bool active = slab < nslabs;
        const uint64_t r0 = slab * ROWS;
        // ---- rows -> A fragments (registers), ||x||^2 ----
        double a[MT][KSTEPS];
        double xn[MT];
#pragma unroll
        for (int mt = 0; mt < MT; mt++) {
            const uint64_t row = r0 + mt * 8 + g;
            const bool rok = active && row < n;
            const TX* xr = x + row * d;
            double s = 0.0;
#pragma unroll
            for (int ks = 0; ks < KSTEPS; ks++) {
                const uint32_t col = ks * 4 + t;
                double v = 0.0;
                if (rok && col < d) v = (double)__ldg(xr + col);
                a[mt][ks] = v;
                s = fma(v, v, s);
            }
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            s += __shfl_xor_sync(0xffffffffu, s, 2);
            xn[mt] = s;
        }
        double best[MT], second[MT];
        uint32_t bidx[MT];
#pragma unroll
        for (int mt = 0; mt < MT; mt++) { best[mt] = -DBL_MAX; second[mt] = -DBL_MAX; bidx[mt] = 0; }

        for (uint32_t ch = 0; ch < nchunks; ch++) {
            const uint32_t c0 = ch * bn;
            if (nchunks > 1 || rd == 0) {
                // (re)load the centroid block; resident across rounds when it is the only one
                if (nchunks > 1) __syncthreads();
                for (uint32_t e = threadIdx.x; e < bn * DP; e += blockDim.x) {
                    const uint32_t r = e / DP, c = e - r * DP;
                    double v = 0.0;
                    if (c0 + r < k && c < d) v = centroids[(size_t)(c0 + r) * d + c];
                    cbuf[(size_t)r * PITCH + c] = v;
                }
                for (uint32_t r = threadIdx.x; r < bn; r += blockDim.x)
                    cn[r] = (c0 + r < k) ? cnorm[c0 + r] : INFINITY;
                __syncthreads();
            }
            const uint32_t cols = min(bn, k - c0);
            for (uint32_t ns = 0; ns < cols; ns += 8 * DMMA_NT) {
                double acc[MT][DMMA_NT][2];
#pragma unroll
                for (int nt = 0; nt < DMMA_NT; nt++) {
                    const double cn0 = cn[ns + nt * 8 + 2 * t], cn1 = cn[ns + nt * 8 + 2 * t + 1];
#pragma unroll
                    for (int mt = 0; mt < MT; mt++) {
                        acc[mt][nt][0] = -0.5 * (xn[mt] + cn0);
                        acc[mt][nt][1] = -0.5 * (xn[mt] + cn1);
                    }
                }
                const double* bp = bbase + (size_t)ns * PITCH;
#pragma unroll
                for (int ks = 0; ks < KSTEPS; ks++) {
                    double b[DMMA_NT];
#pragma unroll
                    for (int nt = 0; nt < DMMA_NT; nt++) b[nt] = bp[(size_t)nt * 8 * PITCH + ks * 4];
#pragma unroll
                    for (int mt = 0; mt < MT; mt++)
#pragma unroll
                        for (int nt = 0; nt < DMMA_NT; nt++) dmma884(acc[mt][nt][0], acc[mt][nt][1], a[mt][ks], b[nt]);
                }
                // epilogue: running max / second max of acc = -dist^2/2, ascending column order
#pragma unroll
                for (int nt = 0; nt < DMMA_NT; nt++)
#pragma unroll
                    for (int e = 0; e < 2; e++) {
                        const uint32_t col = c0 + ns + nt * 8 + 2 * t + e;
#pragma unroll
                        for (int mt = 0; mt < MT; mt++) {
                            const double v = acc[mt][nt][e];
                            second[mt] = fmax(second[mt], fmin(v, best[mt]));
                            const bool gt = v > best[mt];
                            bidx[mt] = gt ? col : bidx[mt];
                            best[mt] = fmax(best[mt], v);
                        }
                    }
            }
        }
        // ---- merge the 4 lanes that share a row, write out, flag near-ties ----
#pragma unroll
        for (int mt = 0; mt < MT; mt++) {
#pragma unroll
            for (int o = 1; o <= 2; o <<= 1) {
                const double ob = __shfl_xor_sync(0xffffffffu, best[mt], o);
                const double os = __shfl_xor_sync(0xffffffffu, second[mt], o);
                const uint32_t oi = __shfl_xor_sync(0xffffffffu, bidx[mt], o);
                const bool take = ob > best[mt] || (ob == best[mt] && oi < bidx[mt]);
                second[mt] = fmax(fmax(second[mt], os), fmin(best[mt], ob));
                bidx[mt] = take ? oi : bidx[mt];
                best[mt] = fmax(best[mt], ob);
            }
            const uint64_t row = r0 + mt * 8 + g;
            if (active && row < n && t == 0) {
                labels[row] = bidx[mt];
                mind[row] = fmax(0.0, -2.0 * best[mt]);
                const double gap = 2.0 * (best[mt] - second[mt]);
                if (!(gap > DMMA_TIE_REL * (xn[mt] + cmax))) {   // also catches NaN
                    const unsigned long long slot = atomicAdd(flag_count, 1ull);
                    flag_rows[slot] = (uint32_t)row;
                }
            }
        }
    }
}

// exact re-decision of the flagged rows: one warp per row, lane l scans centroids l, l+32, ... with the
// reference's arithmetic (widen to f64, diff, square, sequential sum, never fused), then a warp argmin
// with strict < and lowest index on ties (kmeans.rs:334-347 / bbd_tree.rs:101-111).
template <typename TX>
__global__ void __launch_bounds__(256)
refine_rows_kernel(const TX* __restrict__ x, uint32_t d, const double* __restrict__ centroids, uint32_t k,
                   const unsigned long long* __restrict__ flag_count, const uint32_t* __restrict__ flag_rows,
                   uint32_t* __restrict__ labels, double* __restrict__ mind) {
    const unsigned long long count = *flag_count;
    const int lane = threadIdx.x & 31;
    const unsigned long long warp_global = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const unsigned long long nwarps = ((unsigned long long)gridDim.x * blockDim.x) >> 5;
    for (unsigned long long f = warp_global; f < count; f += nwarps) {
        const uint64_t row = flag_rows[f];
        const TX* xr = x + row * d;
        double best = DBL_MAX; uint32_t bi = 0xffffffffu;
        for (uint32_t c = lane; c < k; c += 32) {
            const double* cr = centroids + (size_t)c * d;
            double dist = 0.0;
            for (uint32_t j = 0; j < d; j++) {
                const double r = __dsub_rn((double)xr[j], cr[j]);
                dist = __dadd_rn(dist, __dmul_rn(r, r));
            }
            if (dist < best) { best = dist; bi = c; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ob = __shfl_xor_sync(0xffffffffu, best, o);
            const uint32_t oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ob < best || (ob == best && oi < bi)) { best = ob; bi = oi; }
        }
        if (lane == 0) { labels[row] = bi == 0xffffffffu ? 0u : bi; mind[row] = best; }
    }
}

// cnorm[k] = max_j cnorm[j]  (single warp; k is small)
__global__ void cnorm_max_kernel(double* __restrict__ cnorm, uint32_t k) {
    double m = 0.0;
    for (uint32_t j = threadIdx.x; j < k; j += 32) m = fmax(m, cnorm[j]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (threadIdx.x == 0) cnorm[k] = m;
}

// ||c||^2 of the centroids currently in ctx->d_centroids (the finalize kernel also writes them, but
// sckm_lloyd_step uploads centroids from the host)
__global__ void cnorm_kernel(const double* __restrict__ centroids, uint32_t k, uint32_t d, double* __restrict__ cnorm) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= k) return;
    double s = 0.0;
    for (uint32_t j = 0; j < d; j++) { const double v = centroids[(size_t)c * d + j]; s = fma(v, v, s); }
    cnorm[c] = s;
}

bool dmma_supported(const sckm_dataset* ds, uint64_t k) {
    return ds->d >= 4 && ds->d <= 128 && k >= 16 && k <= (1u << 24) && ds->n < 0xFFFFFFFFull;
}

template <int KSTEPS, int MT, typename TX>
static int launch_t(sckm_dataset* ds, uint64_t k) {
    sckm_ctx* ctx = ds->ctx;
    constexpr int DP = KSTEPS * 4, PITCH = DP + 4;
    const size_t row_bytes = (size_t)PITCH * 8 + 8;                   // staged row + its norm
    uint32_t bn = (uint32_t)(((size_t)ctx->smem_optin - 1024) / row_bytes);
    bn = bn / 64 * 64;
    const uint32_t kpad = (uint32_t)((k + 63) / 64 * 64);
    if (bn >= kpad) bn = kpad;                                         // whole centroid set resident
    if (bn < 64) return fail(ctx, SCKM_ERR_INVALID, "shared memory too small for the DMMA tile");
    const size_t smem = (size_t)bn * row_bytes;
    auto kern = assign_dmma_kernel<KSTEPS, MT, TX>;
    SCKM_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const uint64_t nslabs = (ds->n + 8 * MT - 1) / (8 * MT);
    const unsigned grid = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((nslabs + DMMA_WARPS - 1) / DMMA_WARPS, ctx->num_sms));
    kern<<<grid, DMMA_WARPS * 32, smem, ctx->stream>>>((const TX*)ds->x, ds->n, (uint32_t)ds->d, ctx->d_centroids,
                                                      ctx->d_cnorm, (uint32_t)k, bn, ds->labels, ds->mind,
                                                      ctx->d_flags, ctx->d_flagrows);
    LAUNCH_CHECK_D(ctx);
    return SCKM_OK;
}

template <typename TX>
static int launch_by_d(sckm_dataset* ds, uint64_t k) {
    const uint64_t d = ds->d;
    if (d <= 16) return launch_t<4, 2, TX>(ds, k);
    if (d <= 32) return launch_t<8, 2, TX>(ds, k);
    if (d <= 64) return launch_t<16, 2, TX>(ds, k);
    return launch_t<32, 1, TX>(ds, k);
}

int launch_assign_dmma(sckm_dataset* ds, uint64_t k) {
    sckm_ctx* ctx = ds->ctx;
    if (!dmma_supported(ds, k)) return fail(ctx, SCKM_ERR_INVALID, "shape not supported by the DMMA kernel");
    if (ds->n == 0) return SCKM_OK;
    // near-tie list (one u32 per local row is the worst case: every row tied)
    if (ctx->cap_flagrows < ds->n) {
        if (ctx->d_flagrows) { SCKM_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); cudaFree(ctx->d_flagrows); ctx->d_flagrows = nullptr; }
        SCKM_CUDA(ctx, cudaMalloc((void**)&ctx->d_flagrows, ds->n * sizeof(uint32_t)));
        ctx->cap_flagrows = ds->n;
    }
    SCKM_CUDA(ctx, cudaMemsetAsync(ctx->d_flags, 0, sizeof(unsigned long long), ctx->stream));
    cnorm_kernel<<<(unsigned)((k + 127) / 128), 128, 0, ctx->stream>>>(ctx->d_centroids, (uint32_t)k, (uint32_t)ds->d, ctx->d_cnorm);
    LAUNCH_CHECK_D(ctx);
    cnorm_max_kernel<<<1, 32, 0, ctx->stream>>>(ctx->d_cnorm, (uint32_t)k);
    LAUNCH_CHECK_D(ctx);
    SCKM_TRY(ds->dtype == SCKM_F32 ? launch_by_d<float>(ds, k) : launch_by_d<double>(ds, k));
    const unsigned rgrid = (unsigned)ctx->num_sms * 2;
    if (ds->dtype == SCKM_F32)
        refine_rows_kernel<float><<<rgrid, 256, 0, ctx->stream>>>((const float*)ds->x, (uint32_t)ds->d, ctx->d_centroids,
            (uint32_t)k, ctx->d_flags, ctx->d_flagrows, ds->labels, ds->mind);
    else
        refine_rows_kernel<double><<<rgrid, 256, 0, ctx->stream>>>((const double*)ds->x, (uint32_t)ds->d, ctx->d_centroids,
            (uint32_t)k, ctx->d_flags, ctx->d_flagrows, ds->labels, ds->mind);
    LAUNCH_CHECK_D(ctx);
    return SCKM_OK;
}

}  // namespace sckm
