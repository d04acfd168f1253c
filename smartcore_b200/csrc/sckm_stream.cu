// sckm_stream.cu -- K3s: streaming Lloyd step for small k*d (config C2: 1M x 16, k = 8), where the step is
// HBM-bound (k/4 flop per byte < ~6): one pass over X does assignment AND the per-cluster update.
//
// Replaces BBDTree::clustering (src/algorithm/neighbour/bbd_tree.rs:62-163) for shapes the DMMA tile kernel
// does not take (k < 16).
//
//   * a warp owns batches of 32 consecutive rows; the rows go from HBM straight into DMMA A-fragment registers
//     (16-byte loads, every request = full 32-byte sectors; the feature order is permuted consistently in both
//     GEMM operands so a lane's elements are contiguous in memory) -- no staging, no shared memory in the loop;
//   * scores  x.c_j - ||c_j||^2/2  as 8x8x4 DMMAs against centroid B-fragments that stay in registers for the
//     whole launch; top-2 per row on integer keys; rows whose gap is within 1e-10*(||x||^2 + max||c||^2) are
//     re-decided exactly, in place, by the lane that owns the row (same rule as the DMMA tile kernel), so labels equal
//     the exact direct-form argmin;
//   * update: sums[cluster][feature] += onehot(label)^T . X, again as DMMAs whose accumulators live in registers
//     for the whole launch (the FP64 tensor path used as a wide adder with a fixed, hardware-defined order: no
//     atomics, no read-modify-write traffic); counts are integer adds; one store per CTA into its partial slot at
//     the end => bit-reproducible for a fixed (n, grid).
#include "sckm_common.cuh"
#include "sckm_tile.cuh"
#include <cfloat>
#include <algorithm>
#include <type_traits>

namespace sckm {

#define LAUNCH_CHECK_S(ctx)                                                                        \
    do {                                                                                           \
        (ctx)->launches++;                                                                         \
        cudaError_t _e = cudaGetLastError();                                                       \
        if (_e != cudaSuccess)                                                                     \
            return fail((ctx), SCKM_ERR_CUDA, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), \
                        __FILE__, __LINE__);                                                       \
    } while (0)

#ifndef SCKM_STREAM_WARPS
#define SCKM_STREAM_WARPS 8
#endif
constexpr int STREAM_WARPS = SCKM_STREAM_WARPS;
// Update scheme of the registers-direct path.  1 (default): lanes regroup as (cluster, feature slice) and add the rows of
// THEIR cluster with plain DADDs -- a 32-row batch costs ~max_c(rows of c) x (features per lane) adds instead of the
// 8*KT*NTU DMMAs of the one-hot GEMM (C2: ~32 DADD issue slots instead of 16 DMMAs = 256 clocks of the same FP64
// datapath) and the dependent chain per batch shrinks from 8 DMMA latencies to the adds of one cluster.
// 0: the one-hot GEMM (kept for A/B builds, tools/build_variant.sh, and used by the bulk-copy ring variant).
#ifndef SCKM_STREAM_GROUPED
#define SCKM_STREAM_GROUPED 1
#endif
#ifdef SCKM_STREAM_TRACE
// Diagnostic build only (tools/build_variant.sh trace "-DSCKM_STREAM_TRACE"): per-CTA timeline of the last launch --
// [cta][0] entry, [1] after the dependency wait, [2] loop start, [3] loop end, [5] after the CTA barrier, [4] exit (ns, globaltimer)
__device__ unsigned long long g_stream_trace[1024 * 8];
__device__ __forceinline__ unsigned long long gtimer() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define STREAM_TRACE(slot) do { if (threadIdx.x == 0 && blockIdx.x < 1024) g_stream_trace[blockIdx.x * 8 + (slot)] = gtimer(); } while (0)
#else
#define STREAM_TRACE(slot) do { } while (0)
#endif
#ifndef SCKM_STREAM_EARLY
#define SCKM_STREAM_EARLY 1      // first row loads before the dependency wait (0: after the prologue, A/B builds)
#endif
#ifndef SCKM_STREAM_UNROLL
#define SCKM_STREAM_UNROLL 1
#endif
constexpr int STREAM_MAX_K = 15;
constexpr double STREAM_TIE_REL = 1e-10;

int launch_refine_rows(sckm_dataset* ds, uint64_t k, size_t pk, unsigned grid_ctas);   // sckm_dmma.cu

// Feature (k-dimension) permutation used by both operands of the scoring GEMM so that a lane's KS elements of a
// row are VW-element vectors in memory (one 16-byte LDG per VW k-steps): the dot product does not care about the
// order of its terms, only that A and B agree.  VW = 1 is the textbook layout col = 4*ks + t.
template <int VW> __device__ __forceinline__ constexpr int kcol(int t, int ks) { return (ks / VW) * (4 * VW) + t * VW + (ks % VW); }

template <typename TX, int VW> struct VecLoad;
template <typename TX> struct VecLoad<TX, 1> {
    static __device__ __forceinline__ void ld(const TX* p, double (&o)[1]) { o[0] = (double)__ldg(p); }
    static __device__ __forceinline__ void lds(const TX* p, double (&o)[1]) { o[0] = (double)*p; }
};
template <> struct VecLoad<double, 2> {
    static __device__ __forceinline__ void ld(const double* p, double (&o)[2]) {
        const double2 v = __ldg(reinterpret_cast<const double2*>(p)); o[0] = v.x; o[1] = v.y; }
    static __device__ __forceinline__ void lds(const double* p, double (&o)[2]) {
        const double2 v = *reinterpret_cast<const double2*>(p); o[0] = v.x; o[1] = v.y; }
};
template <> struct VecLoad<float, 2> {
    static __device__ __forceinline__ void ld(const float* p, double (&o)[2]) {
        const float2 v = __ldg(reinterpret_cast<const float2*>(p)); o[0] = v.x; o[1] = v.y; }
    static __device__ __forceinline__ void lds(const float* p, double (&o)[2]) {
        const float2 v = *reinterpret_cast<const float2*>(p); o[0] = v.x; o[1] = v.y; }
};
template <> struct VecLoad<float, 4> {
    static __device__ __forceinline__ void ld(const float* p, double (&o)[4]) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(p)); o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w; }
    static __device__ __forceinline__ void lds(const float* p, double (&o)[4]) {
        const float4 v = *reinterpret_cast<const float4*>(p); o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w; }
};

// 32-byte vectors: one LDG.256 per lane (sm_100 and later).  A warp request then covers FULL 128-byte lines (8 rows x 4
// lanes x 32 B for the A fragments of a 16-double row) instead of half lines, which halves the L1 wavefronts of the row
// loads -- the data stage of the L1 is the busiest unit of this kernel (ncu: 45 % of its peak, everything else < 35 %).
// It did not pay (see launch_stream_vw): kept as an opt-in instantiation.
template <> struct VecLoad<double, 4> {
    static __device__ __forceinline__ void ld(const double* p, double (&o)[4]) {
        asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(o[0]), "=d"(o[1]), "=d"(o[2]), "=d"(o[3]) : "l"(p)); }
    static __device__ __forceinline__ void lds(const double* p, double (&o)[4]) {
        const double2 a = *reinterpret_cast<const double2*>(p), b = *reinterpret_cast<const double2*>(p + 2);
        o[0] = a.x; o[1] = a.y; o[2] = b.x; o[3] = b.y; }
};
template <> struct VecLoad<float, 8> {
    static __device__ __forceinline__ void ld(const float* p, double (&o)[8]) {
        float f[8];
        asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=f"(f[0]), "=f"(f[1]), "=f"(f[2]), "=f"(f[3]), "=f"(f[4]), "=f"(f[5]), "=f"(f[6]), "=f"(f[7]) : "l"(p));
#pragma unroll
        for (int i = 0; i < 8; i++) o[i] = f[i]; }
    static __device__ __forceinline__ void lds(const float* p, double (&o)[8]) {
        const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
        o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w; o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w; }
};

// ---- TMA ring (1-D bulk copies): PTX wrappers ----
__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void s_mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s_u32(bar)), "r"(count));
}
__device__ __forceinline__ void s_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void s_mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile("{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n }"
                 ::"r"(s_u32(bar)), "r"(parity) : "memory");
}
// one contiguous block global -> shared, completion counted in bytes on the mbarrier (SASS UBLKCP)
__device__ __forceinline__ void s_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(s_u32(dst)), "l"(src), "r"(bytes), "r"(s_u32(bar)) : "memory");
}

// Column of update n-tile `nt`, tile column `g`: a lane's VU n-tiles are VU consecutive features (one vector load)
template <int VU> __device__ __forceinline__ constexpr int ucol(int g, int nt) { return (nt / VU) * (8 * VU) + g * VU + (nt % VU); }

// KS: k-steps of the scoring GEMM (features padded to 4*KS); KT: 8-cluster tiles (k <= 8*KT); VW: see kcol;
// DFULL: d == 4*KS exactly (then every offset is a compile-time constant and full batches run without predicates)
// STAGES > 0 (needs DFULL): the 32 rows of a batch are ONE contiguous block in HBM, fetched by a single
// cp.async.bulk (TMA) into a per-warp ring of STAGES buffers -- one instruction per 32 rows, completion on an
// mbarrier, STAGES batches in flight per warp with no registers held -- and both GEMM operands (A: this batch's rows,
// B of the update: the rows again) are read from that buffer.  STAGES = 0: rows straight from HBM into registers.
template <int KS, int KT, int VW, bool DFULL, int STAGES, typename TX>
__global__ void __launch_bounds__(STREAM_WARPS * 32, (KS * KT <= 8) ? 2 : 1)
assign_stream_kernel(const TX* __restrict__ x, uint64_t n, uint32_t d_rt, const double* __restrict__ centroids,
                     const double* __restrict__ cnorm, const double* __restrict__ mu, uint32_t k, uint32_t* __restrict__ labels,
                     double* __restrict__ partials, size_t pk, unsigned long long* __restrict__ nmarked, uint32_t pf_ahead,
                     const LoopState* __restrict__ loop_st, uint32_t loop_it) {
    STREAM_TRACE(0);
    constexpr int NTU = KS / 2 > 0 ? KS / 2 : 1;         // feature n-tiles of the update GEMM (8 features each)
    constexpr int VU = VW < NTU ? VW : NTU;              // features per update load
    constexpr bool GROUPED = SCKM_STREAM_GROUPED != 0 && STAGES == 0;
    constexpr int LPC = 32 / (8 * KT);                   // grouped update: lanes per cluster (4 for k <= 8, 2 for k <= 16)
    constexpr int FPL = (4 * KS) / LPC;                  // ... and features per lane (the padded d split over them)
    constexpr int VG = VW < FPL ? VW : FPL;              // ... loaded VG at a time
    const uint32_t d = DFULL ? (uint32_t)(4 * KS) : d_rt;
    extern __shared__ __align__(16) double smem_s[];     // [warps][k*d] sums | [warps][16] counts | [warps] inertia
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    constexpr bool TMA = STAGES > 0;
    static_assert(!TMA || DFULL, "the bulk-copy ring needs d == 4*KS");
    const uint64_t nbatches = (n + 31) / 32;
    const uint64_t wglobal = (uint64_t)blockIdx.x * STREAM_WARPS + warp;
    const uint64_t nwarps = (uint64_t)gridDim.x * STREAM_WARPS;

    // rows -> A fragments straight from HBM (8 rows x full 32-byte sectors per request).  FULL: all 32 rows exist
    // and d == 4*KS, so there is nothing to predicate and every offset folds into the instruction.
    double a[4][KS];
    auto load_rows = [&](uint64_t row0, auto full_tag) {
        constexpr bool FULL = decltype(full_tag)::value;
        const TX* base = x + (row0 + g) * d + t * VW;
#pragma unroll
        for (int mt = 0; mt < 4; mt++) {
#pragma unroll
            for (int i = 0; i < KS / VW; i++) {
                double v[VW];
                bool have = true;
                if (FULL) {
                    VecLoad<TX, VW>::ld(base + (size_t)mt * 8 * d + i * 4 * VW, v);
                } else {
#pragma unroll
                    for (int e = 0; e < VW; e++) v[e] = 0.0;
                    have = row0 + mt * 8 + g < n && (uint32_t)(i * 4 * VW + t * VW) < d;
                    if (have) VecLoad<TX, VW>::ld(base + (size_t)mt * 8 * d + i * 4 * VW, v);
                }
#pragma unroll
                for (int e = 0; e < VW; e++) a[mt][i * VW + e] = have ? v[e] : 0.0;   // raw: batch() subtracts mu when it picks them up
            }
        }
    };
    // L2 prefetch of batch `b` (32 rows = 32 * d elements, contiguous): lane l touches the l-th 128-byte line (and the
    // following ones when the batch is longer than 4 KB)
    auto l2_prefetch_batch = [&](uint64_t b) {
        if (pf_ahead == 0 || b >= nbatches) return;
        const char* p0 = reinterpret_cast<const char*>(x + b * 32 * d);
        const size_t bytes = (size_t)min((uint64_t)32, n - b * 32) * d * sizeof(TX);
        for (size_t off = (size_t)lane * 128; off < bytes; off += 32 * 128)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(p0 + off));
    };
    auto load_any = [&](uint64_t row0) {
        if (DFULL && row0 + 32 <= n) load_rows(row0, std::true_type{});
        else load_rows(row0, std::false_type{});
    };
    // X is an input of the whole fit, not a product of the kernel before this one on the stream: the first batch (and the
    // L2 prefetches behind it) go out BEFORE the dependency wait and before the dependent loads of the prologue (stop
    // flag, norms, shift, centroid fragments: three round trips that cost 3 us of a 35 us kernel when they came first)
    if (SCKM_STREAM_EARLY && !TMA && wglobal < nbatches) {
        load_any(wglobal * 32);
        for (uint32_t i = 1; i < pf_ahead; i++) l2_prefetch_batch(wglobal + (uint64_t)i * nwarps);
    }
    pdl_wait();                                          // (launched with launch_pdl: the previous step's finalize may still be draining)
    STREAM_TRACE(1);
    // the fit's stop rule may already have fired (kmeans.rs:305): the flag is read here and tested after the loads of the
    // prologue have been issued, so that the whole prologue is one memory round trip, not one per dependent stage
    const bool stop_fired = loop_done(loop_st, loop_it);
    // max_j ||c_j - mu||^2 over the <= 15 centroids: every thread reads them itself (L1 broadcasts) -- no CTA barrier in
    // the prologue of a kernel whose whole run is ~40 us at config C2
    double cmax = 0.0;
    for (uint32_t j = 0; j < k; j++) cmax = fmax(cmax, __ldg(cnorm + j));
    const double tie_half = 0.5 * STREAM_TIE_REL;
    __shared__ __align__(8) uint64_t full_bar[TMA ? STREAM_WARPS : 1][TMA ? STAGES : 1];

    // Both operands of the scoring GEMM are centred on the fit's shift mu (see sckm_dmma.cu / launch_cnorm): rows become
    // x - mu as they arrive, the centroid fragments hold c - mu, cnorm holds ||c - mu||^2, so the cancellation error of
    // ||x||^2 - 2 x.c + ||c||^2 follows the spread of the data, not its distance from the origin.  The update GEMM reads
    // the rows again, uncentred: the sums this kernel stores are plain sums of x.
    double mu_f[KS];
#pragma unroll
    for (int ks = 0; ks < KS; ks++) { const uint32_t col = kcol<VW>(t, ks); mu_f[ks] = col < d ? mu[col] : 0.0; }
    // centroid B fragments and -||c - mu||^2/2, resident in registers for the whole launch
    double bc[KT][KS], hc[KT][2];
#pragma unroll
    for (int nt = 0; nt < KT; nt++) {
        const uint32_t c = nt * 8 + g;
#pragma unroll
        for (int ks = 0; ks < KS; ks++) {
            const uint32_t col = kcol<VW>(t, ks);
            bc[nt][ks] = (c < k && col < d) ? centroids[(size_t)c * d + col] - mu_f[ks] : 0.0;
        }
#pragma unroll
        for (int e = 0; e < 2; e++) {
            const uint32_t ce = nt * 8 + 2 * t + e;
            hc[nt][e] = ce < k ? -0.5 * cnorm[ce] : -INFINITY;
        }
    }
    if (stop_fired) return;
    double cu[KT][NTU][2];                               // per-cluster sums: cluster ct*8+g, feature ucol(2t+e, nt)
#pragma unroll
    for (int ct = 0; ct < KT; ct++)
#pragma unroll
        for (int nt = 0; nt < NTU; nt++) { cu[ct][nt][0] = 0.0; cu[ct][nt][1] = 0.0; }
    uint32_t cnt[KT];
#pragma unroll
    for (int ct = 0; ct < KT; ct++) cnt[ct] = 0;
    // grouped update: this lane adds features [fs*FPL, fs*FPL + FPL) of the rows assigned to cluster cq
    const uint32_t cq = (uint32_t)lane / LPC, fs = (uint32_t)lane % LPC;
    double su[GROUPED ? FPL : 1];
#pragma unroll
    for (int j = 0; j < (GROUPED ? FPL : 1); j++) su[j] = 0.0;
    uint32_t cnt_g = 0;
    double inertia = 0.0;                                // this lane's share (rows whose winning score it holds)

    // the same fragments out of a ring buffer (dense [32][d] image of the batch)
    auto load_rows_smem = [&](const TX* stage, uint64_t row0, auto full_tag) {
        constexpr bool FULL = decltype(full_tag)::value;
        const TX* base = stage + (size_t)g * d + t * VW;
#pragma unroll
        for (int mt = 0; mt < 4; mt++) {
#pragma unroll
            for (int i = 0; i < KS / VW; i++) {
                double v[VW];
#pragma unroll
                for (int e = 0; e < VW; e++) v[e] = 0.0;
                const bool have = FULL || row0 + mt * 8 + g < n;
                if (have) VecLoad<TX, VW>::lds(base + (size_t)mt * 8 * d + i * 4 * VW, v);
#pragma unroll
                for (int e = 0; e < VW; e++) a[mt][i * VW + e] = have ? v[e] - mu_f[i * VW + e] : 0.0;
            }
        }
    };

    // per-warp staging tile for the epilogue: [32 rows][8*KT scores | 4 partial norms | pad], pitch = odd multiple
    // of 16 bytes so that the row-per-lane 16-byte reads are conflict-free
    constexpr int RP = 8 * KT + 6;
    double* tw = smem_s + (size_t)warp * 32 * RP;

    // ring buffers behind the staging tiles (128-byte aligned), one set per warp
    TX* ring_w = nullptr;
    if (TMA) {
        unsigned char* ring0 = reinterpret_cast<unsigned char*>(smem_s) + (((size_t)STREAM_WARPS * 32 * RP * sizeof(double) + 127) / 128) * 128;
        ring_w = reinterpret_cast<TX*>(ring0 + (size_t)warp * (STAGES > 0 ? STAGES : 1) * 32 * d * sizeof(TX));
        if (lane == 0) {
#pragma unroll
            for (int s = 0; s < (STAGES > 0 ? STAGES : 1); s++) s_mbar_init(&full_bar[warp][s], 1);
            asm volatile("fence.mbarrier_init.release.cluster;");
        }
        __syncwarp();
    }
    auto ring_fill = [&](uint64_t b, int stage) {            // lane 0: fetch batch b into `stage`
        const uint64_t row0 = b * 32;
        const uint32_t bytes = (uint32_t)(min((uint64_t)32, n - row0) * d * sizeof(TX));
        s_mbar_expect_tx(&full_bar[warp][stage], bytes);
        s_bulk_g2s(ring_w + (size_t)stage * 32 * d, x + row0 * d, bytes, &full_bar[warp][stage]);
    };

    auto batch = [&](uint64_t b, uint32_t it, auto full_tag) {
        constexpr bool FULL = decltype(full_tag)::value;
        const uint64_t row0 = b * 32;
        const int stage = TMA ? (int)(it % (STAGES > 0 ? STAGES : 1)) : 0;
        const TX* stage_rows = TMA ? ring_w + (size_t)stage * 32 * d : nullptr;
        // (ring mode: the A fragments of this batch were read out of its ring buffer, centred, at the end of the previous turn)
        if (!TMA) {
#pragma unroll
            for (int mt = 0; mt < 4; mt++)
#pragma unroll
                for (int ks = 0; ks < KS; ks++) a[mt][ks] -= mu_f[ks];     // (rows past the end become -mu: their scores are never used)
        }
        // ---- scores x.c - ||c||^2/2 on the FP64 tensor path ----
        double acc[4][KT][2];
#pragma unroll
        for (int mt = 0; mt < 4; mt++)
#pragma unroll
            for (int nt = 0; nt < KT; nt++) { acc[mt][nt][0] = hc[nt][0]; acc[mt][nt][1] = hc[nt][1]; }
#pragma unroll
        for (int ks = 0; ks < KS; ks++)
#pragma unroll
            for (int mt = 0; mt < 4; mt++)
#pragma unroll
                for (int nt = 0; nt < KT; nt++) dmma884(acc[mt][nt][0], acc[mt][nt][1], a[mt][ks], bc[nt][ks]);
        // this lane's share of ||x||^2 for its 4 rows, and the scores, go through the staging tile so that the
        // epilogue runs one row per lane (no butterflies): lane L then owns row row0 + L
#pragma unroll
        for (int mt = 0; mt < 4; mt++) {
            double sq = 0.0;
#pragma unroll
            for (int ks = 0; ks < KS; ks++) sq = fma(a[mt][ks], a[mt][ks], sq);
            double* tr = tw + (mt * 8 + g) * RP;
            tr[8 * KT + t] = sq;
#pragma unroll
            for (int nt = 0; nt < KT; nt++)
                *reinterpret_cast<double2*>(tr + nt * 8 + 2 * t) = make_double2(acc[mt][nt][0], acc[mt][nt][1]);
        }
        __syncwarp();
        // the A registers are dead now.  Registers-direct mode: prefetch the next batch into them so its HBM latency
        // hides under epilogue + update.  Ring mode: read the update's B fragments (this batch's rows again, K = row)
        // out of the ring buffer now, so that their latency hides under the epilogue instead of stalling the DMMAs.
        if (!TMA && b + nwarps < nbatches) load_any((b + nwarps) * 32);
        // ... and pull the batches after that into L2 (one line per lane, no registers held): the kernel is latency-bound on
        // its row loads at four warps per scheduler, and an L2 hit costs a third of an HBM access
        if (!TMA) l2_prefetch_batch(b + (uint64_t)pf_ahead * nwarps);
        double xbe[TMA ? 8 : 1][NTU];
        if (TMA) {
            const TX* ubs = stage_rows + (size_t)t * d + g * VU;
#pragma unroll
            for (int ks = 0; ks < 8; ks++)
#pragma unroll
                for (int j = 0; j < NTU / VU; j++) {
                    double v[VU];
#pragma unroll
                    for (int e = 0; e < VU; e++) v[e] = 0.0;
                    if (FULL || row0 + 4 * ks + t < n) VecLoad<TX, VU>::lds(ubs + (size_t)ks * 4 * d + j * 8 * VU, v);
#pragma unroll
                    for (int e = 0; e < VU; e++) xbe[ks][j * VU + e] = v[e];
                }
        }
        double sc[8 * KT];
        double xn;
        {
            const double* tr = tw + lane * RP;
#pragma unroll
            for (int j = 0; j < 4 * KT; j++) {
                const double2 v = *reinterpret_cast<const double2*>(tr + 2 * j);
                sc[2 * j] = v.x; sc[2 * j + 1] = v.y;
            }
            const double2 p01 = *reinterpret_cast<const double2*>(tr + 8 * KT);
            const double2 p23 = *reinterpret_cast<const double2*>(tr + 8 * KT + 2);
            xn = (p01.x + p01.y) + (p23.x + p23.y);
        }
        __syncwarp();
        // ---- argmax of this lane's row (strict >: lowest index first); rows with another score within the tolerance
        // of the winner (exact ties and NaN included) are re-decided exactly below ----
        double best = sc[0]; uint32_t bi = 0;
#pragma unroll
        for (int j = 1; j < 8 * KT; j++) {
            const bool gt = sc[j] > best;
            bi = gt ? (uint32_t)j : bi;
            best = gt ? sc[j] : best;
        }
        const double thr = fma(-tie_half, xn + cmax, best);      // gap = 2*(best - v) <= tol  <=>  v >= best - tol/2
        uint32_t nnear = 0;
#pragma unroll
        for (int j = 0; j < 8 * KT; j++) nnear += (sc[j] >= thr) ? 1u : 0u;
        const bool valid = FULL || row0 + lane < n;
        bool ok = valid && nnear == 1u;                            // exactly the winner itself (NaN winner: 0)
        double dist = fma(-2.0, best, xn);
        dist = dist < 0.0 ? 0.0 : dist;
        // Near-ties are re-decided right here, exactly, by the lane that owns the row: the reference's arithmetic (widen,
        // diff, square, sequential sum, never fused; raw x and raw centroids), strict <, lowest index
        // (kmeans.rs:334-347 / bbd_tree.rs:101-111).  With k <= 15 centroids that is a few hundred instructions for a
        // rare row, and the step needs no separate refine launch (every launch boundary costs ~4 us of a ~45 us step).
        if (__any_sync(0xffffffffu, valid && !ok)) {
            if (valid && !ok) {
                const TX* xr = x + (row0 + lane) * d;
                double bestd = DBL_MAX; uint32_t bj = 0xffffffffu;
                for (uint32_t c = 0; c < k; c++) {
                    const double* cr = centroids + (size_t)c * d;
                    double dd = 0.0;
                    for (uint32_t j = 0; j < d; j++) {
                        const double r = __dsub_rn((double)xr[j], cr[j]);
                        dd = __dadd_rn(dd, __dmul_rn(r, r));
                    }
                    if (dd < bestd) { bestd = dd; bj = c; }
                }
                bi = bj == 0xffffffffu ? 0u : bj;                  // all distances NaN: the reference keeps cluster 0
                dist = bestd;
                ok = true;
            }
        }
        if (ok) inertia = __dadd_rn(inertia, dist);
        const uint32_t lab = ok ? bi : 0xffffffffu;
        if (valid) labels[row0 + lane] = lab;
        if (GROUPED) {
            // ---- update, grouped: which rows of the batch went to MY cluster (marked / invalid rows carry 0xffffffff
            // and match nothing), then add them in ascending row order -- a fixed order, no atomics.  The rows come out
            // of L1 (this batch was just loaded through it). ----
            unsigned mine = 0;
#pragma unroll
            for (uint32_t c = 0; c < 8u * KT; c++) {
                const unsigned m = __ballot_sync(0xffffffffu, lab == c);
                mine = c == cq ? m : mine;
            }
            cnt_g += __popc(mine);
            const int rounds = __reduce_max_sync(0xffffffffu, __popc(mine));
            const TX* gb = x + row0 * d + fs * FPL;
            // UR rounds per turn: their row loads are all issued before the first add, so a turn costs ONE L1 round trip
            // instead of UR dependent ones (ncu, one round per turn: 31 % of the kernel's stall samples sat on the add
            // that waits for its load).  Lanes that have run out of rows add 0.0.
            constexpr int UR = SCKM_STREAM_UNROLL;
            for (int r = 0; r < rounds; r += UR) {
                double v[UR][GROUPED ? FPL : 1];
#pragma unroll
                for (int u = 0; u < UR; u++) {
                    const bool act = mine != 0u;
                    const int row = act ? __ffs(mine) - 1 : 0;
                    mine &= mine - 1u;
#pragma unroll
                    for (int j = 0; j < FPL / VG; j++) {
                        double w[VG];
#pragma unroll
                        for (int e = 0; e < VG; e++) w[e] = 0.0;
                        if (act && (DFULL || fs * FPL + j * VG < d)) VecLoad<TX, VG>::ld(gb + (size_t)row * d + j * VG, w);
#pragma unroll
                        for (int e = 0; e < VG; e++) v[u][GROUPED ? j * VG + e : 0] = w[e];
                    }
                }
#pragma unroll
                for (int u = 0; u < UR; u++)
#pragma unroll
                    for (int e = 0; e < (GROUPED ? FPL : 1); e++) su[e] = __dadd_rn(su[e], v[u][e]);
            }
        } else {
        // ---- update: sums[cluster][feature] += onehot(label)^T . X as DMMAs accumulating in registers.
        // K dimension = the 32 rows (row 4*ks+t), A = one-hot of the labels, B = the rows again (L1 hits). ----
        const TX* ub = x + (row0 + t) * d + g * VU;
#pragma unroll
        for (int ks = 0; ks < 8; ks++) {
            const uint32_t lr = __shfl_sync(0xffffffffu, lab, 4 * ks + t);
            double xb[NTU];
            if (TMA) {
#pragma unroll
                for (int nt = 0; nt < NTU; nt++) xb[nt] = xbe[ks][nt];
            } else {
#pragma unroll
                for (int j = 0; j < NTU / VU; j++) {
                    double v[VU];
                    if (FULL) {
                        VecLoad<TX, VU>::ld(ub + (size_t)ks * 4 * d + j * 8 * VU, v);
                    } else {
#pragma unroll
                        for (int e = 0; e < VU; e++) v[e] = 0.0;
                        if (row0 + 4 * ks + t < n && (uint32_t)(j * 8 * VU + g * VU) < d)
                            VecLoad<TX, VU>::ld(ub + (size_t)ks * 4 * d + j * 8 * VU, v);
                    }
#pragma unroll
                    for (int e = 0; e < VU; e++) xb[j * VU + e] = v[e];
                }
            }
#pragma unroll
            for (int ct = 0; ct < KT; ct++) {
                const bool hit = lr == (uint32_t)(ct * 8 + g);
                cnt[ct] += hit ? 1u : 0u;
                const double oh = hit ? 1.0 : 0.0;
#pragma unroll
                for (int nt = 0; nt < NTU; nt++) dmma884(cu[ct][nt][0], cu[ct][nt][1], oh, xb[nt]);
            }
        }
        }
        if (TMA) {
            // every lane is done reading this buffer: hand it to the copy engine for the batch STAGES turns ahead
            __syncwarp();
            const uint64_t nb2 = b + (uint64_t)(STAGES > 0 ? STAGES : 1) * nwarps;
            if (lane == 0 && nb2 < nbatches) ring_fill(nb2, stage);
            // and pull the next batch's A fragments out of its buffer (landed long ago) while the update DMMAs drain
            const uint64_t b1 = b + nwarps;
            if (b1 < nbatches) {
                const int st1 = (int)((it + 1) % (STAGES > 0 ? STAGES : 1));
                s_mbar_wait(&full_bar[warp][st1], ((it + 1) / (STAGES > 0 ? STAGES : 1)) & 1u);
                const TX* rows1 = ring_w + (size_t)st1 * 32 * d;
                if (b1 * 32 + 32 <= n) load_rows_smem(rows1, b1 * 32, std::true_type{});
                else load_rows_smem(rows1, b1 * 32, std::false_type{});
            }
        }
    };

    if (!SCKM_STREAM_EARLY && !TMA && wglobal < nbatches) {
        load_any(wglobal * 32);
        for (uint32_t i = 1; i < pf_ahead; i++) l2_prefetch_batch(wglobal + (uint64_t)i * nwarps);
    }
    if (TMA) {
        if (lane == 0)
            for (int s = 0; s < (STAGES > 0 ? STAGES : 1); s++) {
                const uint64_t b = wglobal + (uint64_t)s * nwarps;
                if (b < nbatches) ring_fill(b, s);
            }
        if (wglobal < nbatches) {
            s_mbar_wait(&full_bar[warp][0], 0u);
            if (wglobal * 32 + 32 <= n) load_rows_smem(ring_w, wglobal * 32, std::true_type{});
            else load_rows_smem(ring_w, wglobal * 32, std::false_type{});
        }
    }
    uint32_t it = 0;
    STREAM_TRACE(2);
    for (uint64_t b = wglobal; b < nbatches; b += nwarps, it++) {
        if (DFULL && b * 32 + 32 <= n) batch(b, it, std::true_type{});
        else batch(b, it, std::false_type{});
    }
    STREAM_TRACE(3);
    __syncthreads();                                      // the staging tiles are reused by the combine below
    STREAM_TRACE(5);
    // ---- per-warp results -> shared memory, CTA combine in warp order, store this CTA's partial slot ----
    const uint32_t kd = k * d;
    double* s_sum = smem_s + (size_t)warp * kd;
    uint32_t* s_cnt = reinterpret_cast<uint32_t*>(smem_s + (size_t)STREAM_WARPS * kd) + warp * 16;
    double* s_in = smem_s + (size_t)STREAM_WARPS * kd + STREAM_WARPS * 8;
    if (GROUPED) {
#pragma unroll
        for (int j = 0; j < (GROUPED ? FPL : 1); j++) {
            const uint32_t f = fs * FPL + j;
            if (cq < k && f < d) s_sum[(size_t)cq * d + f] = su[j];
        }
        if (fs == 0 && cq < 16) s_cnt[cq] = cnt_g;
    } else {
#pragma unroll
    for (int ct = 0; ct < KT; ct++) {
        const uint32_t c = ct * 8 + g;
#pragma unroll
        for (int nt = 0; nt < NTU; nt++)
#pragma unroll
            for (int e = 0; e < 2; e++) {
                const uint32_t f = ucol<VU>(2 * t + e, nt);
                if (c < k && f < d) s_sum[(size_t)c * d + f] = cu[ct][nt][e];
            }
        uint32_t cc = cnt[ct];
        cc += __shfl_xor_sync(0xffffffffu, cc, 1);
        cc += __shfl_xor_sync(0xffffffffu, cc, 2);
        if (t == 0 && c < 16) s_cnt[c] = cc;
    }
    }
    double v = inertia;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = __dadd_rn(v, __shfl_xor_sync(0xffffffffu, v, o));
    if (lane == 0) s_in[warp] = v;
    __syncthreads();
    double* part = partials + (size_t)blockIdx.x * ((pk + 15) / 16 * 16);
    for (uint32_t e = threadIdx.x; e < kd; e += blockDim.x) {
        double tsum = 0.0;
#pragma unroll
        for (int w = 0; w < STREAM_WARPS; w++) tsum = __dadd_rn(tsum, smem_s[(size_t)w * kd + e]);
        part[e] = tsum;
    }
    if (threadIdx.x < k) {
        const uint32_t* cbase = reinterpret_cast<const uint32_t*>(smem_s + (size_t)STREAM_WARPS * kd);
        uint32_t c = 0;
#pragma unroll
        for (int w = 0; w < STREAM_WARPS; w++) c += cbase[w * 16 + threadIdx.x];
        part[(size_t)kd + threadIdx.x] = (double)c;
    }
    if (threadIdx.x == 0) {
        double tsum = 0.0;
#pragma unroll
        for (int w = 0; w < STREAM_WARPS; w++) tsum = __dadd_rn(tsum, s_in[w]);
        part[pk - 1] = tsum;
    }
    STREAM_TRACE(4);
}

#ifdef SCKM_STREAM_TRACE
extern "C" int sckm_debug_stream_trace(unsigned long long* out, int n_ctas) {
    return (int)cudaMemcpyFromSymbol(out, g_stream_trace, (size_t)n_ctas * 8 * sizeof(unsigned long long));
}
#endif

bool stream_supported(const sckm_dataset* ds, uint64_t k) {
    return k >= 1 && k <= STREAM_MAX_K && ds->d >= 1 && ds->d <= 32 && ds->n < 0xFFFFFFFFull;
}

template <int KS, int KT, int VW, bool DFULL, int STAGES, typename TX>
static int launch_stream_f(sckm_dataset* ds, uint64_t k, size_t pk, unsigned* grid_out) {
    sckm_ctx* ctx = ds->ctx;
    const uint32_t d = (uint32_t)ds->d;
    const size_t tiles = (size_t)STREAM_WARPS * 32 * (8 * KT + 6) * sizeof(double);               // staging tiles
    const size_t combine = ((size_t)STREAM_WARPS * k * d + STREAM_WARPS * 8 + STREAM_WARPS) * sizeof(double);
    const size_t ring = STAGES ? (size_t)STREAM_WARPS * STAGES * 32 * d * sizeof(TX) : 0;         // behind the tiles
    const size_t smem = std::max(combine, (tiles + 127) / 128 * 128 + ring);
    auto kern = assign_stream_kernel<KS, KT, VW, DFULL, STAGES, TX>;
    uint32_t pf_ahead = 3;                                     // batches ahead of the register prefetch that are pulled into L2
    if (const char* e = getenv("SCKM_STREAM_PF")) pf_ahead = (uint32_t)std::max(0, std::min(16, atoi(e)));
    if (smem > 48 * 1024) SCKM_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int ctas_per_sm = 0;
    SCKM_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, kern, STREAM_WARPS * 32, smem));
    ctas_per_sm = std::max(1, std::min(ctas_per_sm, 4));
    const uint64_t nbatches = (ds->n + 31) / 32;
    const unsigned grid = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((nbatches + STREAM_WARPS - 1) / STREAM_WARPS,
                                                                              (uint64_t)ctx->num_sms * ctas_per_sm));
    SCKM_CUDA(ctx, launch_pdl(kern, dim3(grid), dim3(STREAM_WARPS * 32), smem, ctx->stream, (const TX*)ds->x, ds->n, d, (const double*)ctx->d_centroids,
                              (const double*)ctx->d_cnorm, (const double*)ctx->d_mu, (uint32_t)k, ds->labels, ctx->d_partials, pk, ctx->d_flags,
                              pf_ahead, SCKM_LOOP_ARGS(ctx)));
    LAUNCH_CHECK_S(ctx);
    *grid_out = grid;
    return SCKM_OK;
}

template <int KS, int KT, int VW, typename TX>
static int launch_stream_t(sckm_dataset* ds, uint64_t k, size_t pk, unsigned* grid_out) {
    if (ds->d == 4 * KS) {
        // whole rows, 16-byte multiples (d = 4*KS, KS even): batches are contiguous 16-byte-aligned blocks, so the
        // TMA ring applies.  Measured at 10M x 16 k=8 on B200: ring 338 us (f64) / 258 us (f32), registers-direct
        // 316 us / 243 us.  With f32 rows (half the bytes) the kernel still takes 243 us: it is bound by the FP64
        // datapath (32 DMMAs + ~55 scalar FP64 instructions per 32 rows keep it ~70 % busy) and the latency of the
        // per-batch chain at 4 warps per scheduler, not by how the rows arrive; the ring adds an mbarrier wait and
        // shared-memory reads with 2-way (A) / 4-way (update B) bank conflicts from the dense 128-byte row pitch a
        // single bulk copy produces.  It therefore stays opt-in (SCKM_STREAM_TMA=1, covered by the parity tests).
        // An integer-key epilogue (as in the tile kernel) measured 316 / 254 us: no gain here, not kept.
        constexpr bool RING_OK = (4 * KS * sizeof(TX)) % 16 == 0;
        if (RING_OK && getenv("SCKM_STREAM_TMA")) {
            constexpr int STAGES = (32 * 4 * KS * sizeof(TX) <= 2048) ? 4 : 2;
            return launch_stream_f<KS, KT, VW, true, RING_OK ? STAGES : 0, TX>(ds, k, pk, grid_out);
        }
        return launch_stream_f<KS, KT, VW, true, 0, TX>(ds, k, pk, grid_out);
    }
    return launch_stream_f<KS, KT, VW, false, 0, TX>(ds, k, pk, grid_out);
}

template <int KS, int KT, typename TX>
static int launch_stream_vw(sckm_dataset* ds, uint64_t k, size_t pk, unsigned* grid_out) {
    constexpr int VMAX = (16 / sizeof(TX)) < KS ? (16 / sizeof(TX)) : KS;     // widest 16-byte vector that fits KS
    constexpr int V32 = 32 / sizeof(TX);                                       // 32-byte vectors (LDG.256) when they fit
    // Measured at C2 (bench.py, 1M x 16): 40.8 us with 16-byte loads, 51.6 us with 32-byte loads when the update takes one
    // round per turn (39.1 us when it takes four, SCKM_STREAM_UNROLL=4: no better than the 16-byte form) -- opt-in only
    if (V32 <= KS && ds->d % V32 == 0 && getenv("SCKM_STREAM_256"))
        return launch_stream_t<KS, KT, (V32 <= KS ? V32 : VMAX), TX>(ds, k, pk, grid_out);
    if (ds->d % VMAX == 0) return launch_stream_t<KS, KT, VMAX, TX>(ds, k, pk, grid_out);
    if (VMAX > 2 && ds->d % 2 == 0) return launch_stream_t<KS, KT, 2, TX>(ds, k, pk, grid_out);
    return launch_stream_t<KS, KT, 1, TX>(ds, k, pk, grid_out);
}

template <typename TX>
static int launch_stream_by_d(sckm_dataset* ds, uint64_t k, size_t pk, unsigned* grid_out) {
    const uint64_t d = ds->d;
    if (k <= 8) {
        if (d <= 8) return launch_stream_vw<2, 1, TX>(ds, k, pk, grid_out);
        if (d <= 16) return launch_stream_vw<4, 1, TX>(ds, k, pk, grid_out);
        return launch_stream_vw<8, 1, TX>(ds, k, pk, grid_out);
    }
    if (d <= 8) return launch_stream_vw<2, 2, TX>(ds, k, pk, grid_out);
    if (d <= 16) return launch_stream_vw<4, 2, TX>(ds, k, pk, grid_out);
    return launch_stream_vw<8, 2, TX>(ds, k, pk, grid_out);
}

int launch_cnorm(sckm_ctx* ctx, uint64_t k, uint64_t d, bool center);   // sckm_dmma.cu

// labels + per-warp partials (fused update); the caller reduces ctx->partial_slots_used slots
int launch_assign_stream(sckm_dataset* ds, uint64_t k) {
    sckm_ctx* ctx = ds->ctx;
    if (!stream_supported(ds, k)) return fail(ctx, SCKM_ERR_INVALID, "shape not supported by the streaming kernel");
    const size_t pk = (size_t)k * ds->d + k + 1;
    SCKM_TRY(ensure_workspace(ctx, k, ds->d, (size_t)ctx->num_sms * 4 * STREAM_WARPS));
    ctx->partial_slots_used = 0;
    if (ds->n == 0) return SCKM_OK;
    SCKM_TRY(launch_cnorm(ctx, k, ds->d, true));
    ctx->packed_centered = false;                                 // the update GEMM sums the rows as they are
    unsigned grid = 0;
    SCKM_TRY(ds->dtype == SCKM_F32 ? launch_stream_by_d<float>(ds, k, pk, &grid) : launch_stream_by_d<double>(ds, k, pk, &grid));
    ctx->partial_slots_used = grid;          // one slot per CTA; near-ties were re-decided inside the kernel: no refine launch
    return SCKM_OK;
}

}  // namespace sckm
