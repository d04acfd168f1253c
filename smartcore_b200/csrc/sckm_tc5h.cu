// sckm_tc5h.cu -- K2h: Lloyd assignment for f32 data with d <= 32 on the 5th-generation tensor cores, 3xFP16.
//
// Replaces the per-iteration work of BBDTree::clustering (src/algorithm/neighbour/bbd_tree.rs:62-163), like K2
// (sckm_tc5.cu), whose structure it keeps: one persistent warp-specialised CTA per SM, accumulators in tensor memory,
// one epilogue thread per row, primed chunk skipping, and the same exactness contract -- the tensor cores only RANK;
// rows whose best/second gap is within the error bound of that arithmetic are re-decided by refine_rows_kernel in the
// reference's f64 arithmetic, every row's distance to its centroid is recomputed in f64, sums are fixed-order f64.
//
// What changes is the arithmetic of the ranking:
//   * operands are FP16, not TF32: both have an 11-bit significand, so the split x = hi + lo carries the same 22 bits,
//     but kind::f16 multiplies 16 K-columns per instruction where kind::tf32 multiplies 8: half the tensor time and
//     half the shared-memory operand traffic for the same three products Xh.Ch + Xh.Cl + Xl.Ch.
//   * FP16 has a 5-bit exponent, so both operands are scaled by powers of two (exact): every row by 2^s_row, chosen
//     from the row's own largest element (the epilogue thread that owns the row does the split, so the scale is a
//     register of that thread), every centroid by 2^s_c from max ||c||^2.  Elements that fall below the FP16 normal
//     range after scaling lose at most 2^-38 of the operand's largest element: far inside the tie margin, which is
//     relative to ||x||^2 + max ||c||^2.  A row keeps its scale for all comparisons (a positive factor does not change
//     an argmax); best and second are taken back to real units in f64 for the tie test.
//   * -||c||^2/2 no longer costs an FADD per score in the epilogue: it is a rank-one term of the GEMM itself,
//     [2^s_row] x [-||c||^2/2 * 2^s_c], issued as one more MMA per tile (K = 16, BF16 operands: 8-bit exponent, the
//     norm split into three BF16 pieces = 24 bits).  The accumulator IS the scaled score; the chunk pre-pass is half an
//     FMNMX3 per score.  (SCKM_TC5H_NOFOLD=1 keeps the norm in the epilogue instead, one FFMA per score: A/B + a
//     fallback for the tests.)
//
// Shared memory per CTA (190 KB): raw f32 super-tile as TMA delivers it (two stages: the exact part at the end of a
// super-tile reads its rows again from there), FP16 image [row][hi 32 | lo 32] (128-byte rows, SWIZZLE_128B: hi and lo
// are K-offsets 0 / 64 B of ONE atom; two stages), centroid block in the same form (two stages), the two small BF16
// operands of the rank-one MMA, and the merge scratch of the two-threads-per-row variant.
// The accumulators of the two tiles of a super-tile are separate rings (their own full / empty barriers): a tile's
// epilogue warps start as soon as ITS seven MMAs are done and hand the columns back without waiting for the other tile.
// The MMA warp walks its loops as a whole warp with ONE elected lane issuing, loops unrolled, descriptors as (low, high)
// words: see the comment at the issuer.  How each of these steps was measured: DESIGN.md, K2h row; profiles/README.md.
#include "sckm_common.cuh"
#include "sckm_tile.cuh"
#include "sckm_umma.cuh"
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <cfloat>
#include <algorithm>
#include <cstdlib>

namespace sckm {

#define LAUNCH_CHECK_H(ctx)                                                                        \
    do {                                                                                           \
        (ctx)->launches++;                                                                         \
        cudaError_t _e = cudaGetLastError();                                                       \
        if (_e != cudaSuccess)                                                                     \
            return fail((ctx), SCKM_ERR_CUDA, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), \
                        __FILE__, __LINE__);                                                       \
    } while (0)

constexpr int H_BM = 128;                // rows per MMA tile (TMEM lanes)
constexpr int H_BN = 128;                // centroids per block
constexpr int H_TILES = 2;               // tiles per super-tile
constexpr int H_SLOT_WARPS = 8;          // partial slots per CTA: one per (tile, TMEM lane quadrant)
// EW epilogue warps (8 or 16): warps 0 X-producer, 1 MMA, 2..EW+1 epilogue, EW+2 C-producer.  With 16, a row's 128 columns
// of a block are shared by two threads (column parts) that merge their top-2 at the end of the super-tile.
constexpr double H_TIE_REL = 2e-5;       // as K2: >= 10x the error of (22-bit operands, FP32 accumulation)
constexpr int H_XMAX_EXP = 13;           // scaled operands lie below 2^(13+1) (rows) / 2^13 (centroids)
constexpr int H_ROW_FLOOR = 20;          // a row is never scaled as if it were smaller than 2^-20 of the centroids

struct alignas(1024) HSmem {             // dynamic shared memory image (base aligned to 1024 B)
    float xraw[2][H_TILES][H_BM * 32];             // raw rows as TMA delivers them (SWIZZLE_128B); read again by the exact part
    __half xs[2][H_TILES][H_BM * 64];              // [row][hi 32 | lo 32] scaled FP16, SWIZZLE_128B
    __half cs[2][H_BN * 64];                       // centroid block, same form
    __nv_bfloat16 xe[2][H_TILES][H_BM * 16];       // rank-one A operand: row r = {p, p, p, 0, ...}, p = 2^s_row (no swizzle)
    __nv_bfloat16 ce[2][H_BN * 16];                // rank-one B operand: centroid c = {h1, h2, h3, 0, ...} (no swizzle)
    uint64_t raw_full[2], raw_empty[2], x_ready[2], x_empty[2], c_full[2], c_empty[2];
    uint64_t t_full[2][H_TILES], t_empty[2][H_TILES];   // one accumulator ring per tile: its epilogue warps and the MMA warp
    float m_best[2][H_TILES][H_BM], m_second[2][H_TILES][H_BM];   // column-part merge scratch (EW == 16), double-buffered by super-tile
    uint32_t m_idx[2][H_TILES][H_BM];
    uint32_t tmem_base;
};

#ifdef SCKM_TC5H_TRACE
// Diagnostic build only (tools/build_variant.sh h5trace "-DSCKM_TC5H_TRACE"): clocks spent by the MMA warp and by epilogue
// warp 2 of every CTA in each of their waits, summed over the launch: [cta][0] MMA x_ready, [1] c_full, [2] t_empty,
// [3] MMA total, [4] epilogue t_full, [5] split, [6] exact part, [7] epilogue total
__device__ long long g_tc5h_trace[1024 * 8];
#define H_T0() const long long _t0 = clock64()
#define H_ACC(slot) do { if ((threadIdx.x & 31) == 0) g_tc5h_trace[blockIdx.x * 8 + (slot)] += clock64() - _t0; } while (0)
extern "C" int sckm_debug_tc5h_trace(long long* out, int n_ctas, int reset) {
    cudaError_t e = cudaMemcpyFromSymbol(out, g_tc5h_trace, (size_t)n_ctas * 8 * sizeof(long long));
    if (reset) { static long long z[1024 * 8]; cudaMemcpyToSymbol(g_tc5h_trace, z, sizeof(z)); }
    return (int)e;
}
#else
#define H_T0() do { } while (0)
#define H_ACC(slot) do { } while (0)
#endif

// exponent m with 2^m >= sqrt(cmax) (cmax = max ||c||^2), clamped; 0 when there is nothing to scale by
__host__ __device__ __forceinline__ int h_centroid_exp(double cmax) {
    if (!(cmax > 0.0) || cmax > 1.7e308) return 0;
    long long b;
#ifdef __CUDA_ARCH__
    b = __double_as_longlong(cmax);
#else
    memcpy(&b, &cmax, 8);
#endif
    const int ec = (int)((b >> 52) & 0x7ff) - 1023;       // cmax in [2^ec, 2^(ec+1))
    int m = (ec + 2) >> 1;                                 // floor((ec + 2) / 2) >= (ec + 1) / 2
    return m < -100 ? -100 : (m > 100 ? 100 : m);
}
__device__ __forceinline__ float h_pow2f(int e) { return __uint_as_float((uint32_t)(e + 127) << 23); }              // |e| <= 126
__device__ __forceinline__ double h_pow2d(int e) { return __longlong_as_double((long long)(e + 1023) << 52); }     // |e| <= 1022

// Per step: centroids (f64 [k][d]) -> scaled FP16 hi | lo image [kpad][64], an f32 copy [kpad][32] (priming), the scaled
// norm term -||c||^2/2 * 2^s_c in f32 [kpad] (-inf on padding rows) and its three BF16 pieces in the shared-memory layout
// of the rank-one operand (per 128-centroid block: 16 groups x 2 K-chunks x 8 rows x 8 elements).  One CTA per block.
__global__ void __launch_bounds__(256)
tc5h_prep_kernel(const double* __restrict__ centroids, const double* __restrict__ cnorm, uint32_t k, uint32_t d,
                 __half* __restrict__ chl, float* __restrict__ c32, float* __restrict__ hcn_sc, __nv_bfloat16* __restrict__ ce_g) {
    const double cmax = cta_max(cnorm, k);
    const int s_c = H_XMAX_EXP - h_centroid_exp(cmax);
    const float cscale = h_pow2f(s_c);
    const uint32_t r0 = blockIdx.x * H_BN;
    for (uint32_t e = threadIdx.x; e < H_BN * 32; e += blockDim.x) {
        const uint32_t r = r0 + e / 32, c = e % 32;
        const float v = (r < k && c < d) ? (float)centroids[(size_t)r * d + c] : 0.f;
        const float vs = v * cscale;
        const __half h = __float2half_rn(vs);
        const __half l = __float2half_rn(vs - __half2float(h));
        chl[(size_t)r * 64 + c] = h;
        chl[(size_t)r * 64 + 32 + c] = l;
        c32[(size_t)r * 32 + c] = v;
    }
    for (uint32_t e = threadIdx.x; e < H_BN; e += blockDim.x) {
        const uint32_t r = r0 + e;
        __nv_bfloat16 h1, h2 = __float2bfloat16_rn(0.f), h3 = h2;
        float v = -INFINITY;
        if (r < k) {
            v = (float)(-0.5 * cnorm[r] * (double)cscale);
            h1 = __float2bfloat16_rn(v);
            if (isfinite(v)) {
                const float r1 = v - __bfloat162float(h1);
                h2 = __float2bfloat16_rn(r1);
                h3 = __float2bfloat16_rn(r1 - __bfloat162float(h2));
            }
        } else {
            h1 = __float2bfloat16_rn(v);
        }
        hcn_sc[r] = v;
        __nv_bfloat16* blk = ce_g + (size_t)blockIdx.x * (H_BN * 16) + (e >> 3) * 128 + (e & 7) * 8;   // K-chunk 0 of row e
        const __nv_bfloat16 z = __float2bfloat16_rn(0.f);
        blk[0] = h1; blk[1] = h2; blk[2] = h3;
#pragma unroll
        for (int j = 3; j < 8; j++) blk[j] = z;
#pragma unroll
        for (int j = 0; j < 8; j++) blk[64 + j] = z;                                                   // K-chunk 1
    }
}

// FOLD: the norm term rides in the GEMM (rank-one BF16 MMA); else it is added per score in the epilogue.
// TXS: type of the rows used for the exact part (float = the data itself, double = f64 data ranked through an f32 shadow).
// KS: 16-column K-steps per product (1: d <= 16, 2: d <= 32) -- a template parameter so that the MMA warp's issue loop
// unrolls completely: with run-time loops and per-MMA descriptor arithmetic that ONE thread needed ~1850 clocks per
// centroid block for 14 MMAs the tensor pipe executes in ~900, and the whole CTA ran at its pace (bench/c5_trace_probe.py).
template <bool FOLD, int EW, int KS, typename TXS>
__global__ void __launch_bounds__((EW + 3) * 32, 1)
assign_tc5h_kernel(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapC,
                   const TXS* __restrict__ xsrc, uint64_t n, uint32_t d,
                   const double* __restrict__ centroids, const double* __restrict__ cnorm, const float* __restrict__ hcn_sc,
                   const float* __restrict__ c32, const __nv_bfloat16* __restrict__ ce_g, const uint32_t* prev_labels,
                   uint32_t k, uint32_t nblocks, uint32_t* labels, double* __restrict__ mind,
                   double* __restrict__ partials, size_t pk, unsigned long long* __restrict__ nmarked,
                   const LoopState* __restrict__ loop_st, uint32_t loop_it) {
    if (loop_done(loop_st, loop_it)) return;                          // the fit's stop rule already fired (kmeans.rs:305)
    constexpr int TSTAGE = H_TILES * H_BN;                            // TMEM columns per accumulator stage
    constexpr int CP = EW / 8;                                        // column parts per row (threads sharing a row)
    constexpr int COLS = H_BN / CP;                                   // columns per epilogue thread per block
    static_assert(EW == 8 || EW == 16, "8 or 16 epilogue warps");
    constexpr uint32_t IDESC_F16 = (1u << 4) | ((uint32_t)(H_BN >> 3) << 17) | ((uint32_t)(H_BM >> 4) << 24);   // F16 x F16 -> F32
    constexpr uint32_t IDESC_BF16 = IDESC_F16 | (1u << 7) | (1u << 10);                                         // BF16 x BF16 -> F32
    extern __shared__ unsigned char smem_raw[];
    HSmem& S = *reinterpret_cast<HSmem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint64_t rows_per_super = (uint64_t)H_TILES * H_BM;
    const uint64_t nsuper = (n + rows_per_super - 1) / rows_per_super;
    const double cmax = cta_max(cnorm, k);                            // max_j ||c_j||^2 (all threads take part)
    const int m_c = h_centroid_exp(cmax);
    const int s_c = H_XMAX_EXP - m_c;

    if (threadIdx.x == 0) {
        for (int s = 0; s < 2; s++) {
            mbar_init(&S.raw_full[s], 1); mbar_init(&S.raw_empty[s], H_SLOT_WARPS);
            mbar_init(&S.x_ready[s], EW); mbar_init(&S.x_empty[s], 1);
            mbar_init(&S.c_full[s], 1); mbar_init(&S.c_empty[s], 1);
            for (int m = 0; m < H_TILES; m++) { mbar_init(&S.t_full[s][m], 1); mbar_init(&S.t_empty[s][m], EW / H_TILES); }
        }
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (FOLD) {                                                       // K-chunk 1 of the rank-one A operand stays zero
        uint4* z = reinterpret_cast<uint4*>(&S.xe[0][0][0]);
        for (uint32_t i = threadIdx.x; i < sizeof(S.xe) / 16; i += blockDim.x) z[i] = make_uint4(0u, 0u, 0u, 0u);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    constexpr uint32_t TMEM_COLS = 2 * TSTAGE;                        // 512
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&S.tmem_base)), "r"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = S.tmem_base;

    if (warp == 0) {
        // ================= TMA producer: raw X super-tiles (two stages, freed when the super-tile's exact part is done) =================
        if (lane == 0) {
            uint32_t it = 0;
            for (uint64_t st = blockIdx.x; st < nsuper; st += gridDim.x, it++) {
                const int rs = it & 1;
                mbar_wait(&S.raw_empty[rs], ((it >> 1) & 1) ^ 1);
                mbar_expect_tx(&S.raw_full[rs], H_TILES * H_BM * 32 * 4);
                for (int m = 0; m < H_TILES; m++)
                    tma_load_2d(S.xraw[rs][m], &mapX, 0, (int)(st * rows_per_super + (uint64_t)m * H_BM), &S.raw_full[rs]);
            }
        }
    } else if (warp == EW + 2) {
        // ================= TMA producer: centroid blocks (FP16 image + the rank-one operand) =================
        if (lane == 0) {
            uint32_t j = 0;
            for (uint64_t st = blockIdx.x; st < nsuper; st += gridDim.x)
                for (uint32_t b = 0; b < nblocks; b++, j++) {
                    const int cs = j & 1; const uint32_t ph = (j >> 1) & 1;
                    mbar_wait(&S.c_empty[cs], ph ^ 1);
                    mbar_expect_tx(&S.c_full[cs], H_BN * 128 + (FOLD ? H_BN * 32 : 0));
                    tma_load_2d(S.cs[cs], &mapC, 0, (int)(b * H_BN), &S.c_full[cs]);
                    if (FOLD) bulk_load_1d(S.ce[cs], ce_g + (size_t)b * (H_BN * 16), H_BN * 32, &S.c_full[cs]);
                }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        // The WHOLE warp walks the loops and waits on the barriers; one elected lane issues.  Everything the MMAs take
        // (descriptor words, TMEM columns) is then warp-uniform and lives in uniform registers -- with `if (lane == 0)`
        // around the loops the compiler moved each operand into a uniform register through a per-MMA election loop
        // (~14 instructions and two branches per tcgen05.mma).
        {
            uint32_t it = 0, j = 0;
            const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
            // low / high words of the operand descriptors (see umma_desc_sw128 / umma_desc_nosw)
            constexpr uint32_t SW128_HI = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);
            constexpr uint32_t NOSW_HI = (uint32_t)(256 >> 4) | (1u << 14);
            const uint32_t x_lo0 = ((smem_u32(&S.xs[0][0][0]) >> 4) & 0x3FFF) | (1u << 16);
            const uint32_t c_lo0 = ((smem_u32(&S.cs[0][0]) >> 4) & 0x3FFF) | (1u << 16);
            const uint32_t xe_lo0 = ((smem_u32(&S.xe[0][0][0]) >> 4) & 0x3FFF) | ((uint32_t)(128 >> 4) << 16);
            const uint32_t ce_lo0 = ((smem_u32(&S.ce[0][0]) >> 4) & 0x3FFF) | ((uint32_t)(128 >> 4) << 16);
#ifdef SCKM_TC5H_TRACE
            const long long tm0 = clock64();
#endif
            for (uint64_t st = blockIdx.x; st < nsuper; st += gridDim.x, it++) {
                const int xs = it & 1; const uint32_t xph = (it >> 1) & 1;
                { H_T0(); mbar_wait(&S.x_ready[xs], xph); H_ACC(0); }   // scaled hi | lo image written by the epilogue warps
                asm volatile("tcgen05.fence::after_thread_sync;");
                for (uint32_t b = 0; b < nblocks; b++, j++) {
                    const int cs = j & 1; const uint32_t ph = (j >> 1) & 1;   // centroid stage == TMEM stage index
                    { H_T0(); mbar_wait(&S.c_full[cs], ph); H_ACC(1); }
                    // descriptors differ from stage to stage only in the 14-bit address field of their low word, and the
                    // K offsets (hi at byte 0 of the 128-byte row, lo at byte 64, 32 bytes per K-step; units of 16 B) never
                    // carry out of it: tiles are 1024-byte aligned
                    const uint32_t cLo = c_lo0 + (uint32_t)cs * (uint32_t)(sizeof(S.cs[0]) >> 4);
                    const uint32_t ceLo = ce_lo0 + (uint32_t)cs * (uint32_t)(sizeof(S.ce[0]) >> 4);
#pragma unroll
                    for (int m = 0; m < H_TILES; m++) {
                        { H_T0(); mbar_wait(&S.t_empty[cs][m], ph ^ 1); H_ACC(2); }
                        asm volatile("tcgen05.fence::after_thread_sync;");
                        const uint32_t tcol = tmem_u + (uint32_t)(cs * TSTAGE + m * H_BN);
                        const uint32_t xLo = x_lo0 + (uint32_t)(xs * H_TILES + m) * (uint32_t)(sizeof(S.xs[0][0]) >> 4);
                        if (elect_one()) {
                        // Xh.Ch + Xh.Cl + Xl.Ch
#pragma unroll
                        for (int prod = 0; prod < 3; prod++)
#pragma unroll
                            for (int ks = 0; ks < KS; ks++)
                                umma_f16_parts(tcol, xLo + (prod == 2 ? 4 : 0) + 2 * ks, SW128_HI, cLo + (prod == 1 ? 4 : 0) + 2 * ks, SW128_HI,
                                               IDESC_F16, (prod | ks) != 0);
                        if (FOLD) umma_f16_parts(tcol, xe_lo0 + (uint32_t)(xs * H_TILES + m) * (uint32_t)(sizeof(S.xe[0][0]) >> 4), NOSW_HI,
                                                 ceLo, NOSW_HI, IDESC_BF16, 1);
                        umma_commit(&S.t_full[cs][m]);               // this tile's accumulators are ready for its four epilogue warps
                        if (m == H_TILES - 1) {
                            umma_commit(&S.c_empty[cs]);             // centroid stage free once these MMAs have read it
                            if (b + 1 == nblocks) umma_commit(&S.x_empty[xs]);   // X stage free for the split of super-tile it + 2
                        }
                        }
                        __syncwarp();
                    }
                }
            }
#ifdef SCKM_TC5H_TRACE
            if (lane == 0) g_tc5h_trace[blockIdx.x * 8 + 3] += clock64() - tm0;
#endif
        }
    } else {
        // ================= epilogue warps: one thread per row =================
        const int ew = warp - 2;                                     // 0..EW-1
        const int m = (ew >> 2) / CP;                                // tile of the super-tile
        const int cp = (ew >> 2) % CP;                               // column part of the row this thread ranks
        const int q = warp & 3;                                      // TMEM lane quadrant this warp may access
        const int rloc = q * 32 + lane;                              // row within the tile
        double* part = partials + ((size_t)blockIdx.x * H_SLOT_WARPS + (m * 4 + (ew & 3))) * ((pk + 15) / 16 * 16);
        const float tie25 = 2.5f * (float)(0.5 * H_TIE_REL);
        const float cmax_f = (float)cmax, cscale = h_pow2f(s_c);
        // what the split leaves behind for the super-tile it prepared
        double xn_next = 0.0; float prime_next = -FLT_MAX, sf_next = 1.f; int srow_next = 0; bool force_next = false;
        // Split my row of super-tile `stn` (iteration `itn` of this CTA): scale by 2^s_row, write the FP16 hi | lo image and
        // the rank-one operand, take ||x||^2 in f64 and the priming bound, then hand the stage to the MMA warp and the raw
        // buffer back to the TMA producer.
        auto split_stage = [&](uint32_t itn, uint64_t stn) {
            const int xs = itn & 1;
            mbar_wait(&S.raw_full[xs], (itn >> 1) & 1);
            mbar_wait(&S.x_empty[xs], ((itn >> 1) & 1) ^ 1);          // the MMAs of super-tile itn - 2 have read this stage
            const float4* src = reinterpret_cast<const float4*>(S.xraw[xs][m] + rloc * 32);
            float4 v[8];
#pragma unroll
            for (int c = 0; c < 8; c++) v[c] = src[c ^ (rloc & 7)];   // logical 16-byte chunk c (the swizzle permutes chunks)
            float mx = 0.f; double xn = 0.0;
#pragma unroll
            for (int c = 0; c < 8; c++) {
                mx = fmaxf(fmaxf(mx, fabsf(v[c].x)), fabsf(v[c].y)); mx = fmaxf(fmaxf(mx, fabsf(v[c].z)), fabsf(v[c].w));
                xn = fma((double)v[c].x, (double)v[c].x, xn); xn = fma((double)v[c].y, (double)v[c].y, xn);
                xn = fma((double)v[c].z, (double)v[c].z, xn); xn = fma((double)v[c].w, (double)v[c].w, xn);
            }
            const int e_raw = (int)((__float_as_uint(mx) >> 23) & 0xffu) - 127;       // mx in [2^e, 2^(e+1)); 128 for inf
            int s_row = H_XMAX_EXP - max(e_raw, m_c - H_ROW_FLOOR);
            bool force = false;
            if (s_row > 126) { s_row = 126; force = true; }           // beyond what one f32 factor can carry: decide exactly
            const float sf = h_pow2f(s_row);
            // priming bound from the centroid the row had in the previous step, in scaled units
            const uint64_t rown = stn * rows_per_super + (uint64_t)m * H_BM + rloc;
            float prime = -FLT_MAX;
            if (prev_labels != nullptr) {
                prime = FLT_MAX;                                       // rows past the end never ask for a scan
                if (rown < n) {
                    prime = -FLT_MAX;
                    const uint32_t pl = prev_labels[rown];
                    if (pl < k) {
                        const float4* cr = reinterpret_cast<const float4*>(c32 + (size_t)pl * 32);
                        float dot = 0.f, xx = 0.f;
#pragma unroll
                        for (int c = 0; c < 8; c++) {
                            const float4 cv = __ldg(cr + c);
                            dot = fmaf(v[c].x, cv.x, dot); dot = fmaf(v[c].y, cv.y, dot); dot = fmaf(v[c].z, cv.z, dot); dot = fmaf(v[c].w, cv.w, dot);
                            xx = fmaf(v[c].x, v[c].x, xx); xx = fmaf(v[c].y, v[c].y, xx); xx = fmaf(v[c].z, v[c].z, xx); xx = fmaf(v[c].w, v[c].w, xx);
                        }
                        // 2.5 tie margins below the score of that centroid (the tensor cores' value differs from `dot` by
                        // ~1e-6 (xx + cmax), a tenth of one margin); units: 2^s_c, then 2^s_row
                        const float p = fmaf(dot - tie25 * (xx + cmax_f), cscale, __ldg(hcn_sc + pl)) * sf;
                        prime = p == p ? p : -FLT_MAX;                 // NaN centroid: no priming
                    }
                }
            }
            // scaled FP16 hi | lo: logical chunks 0..3 = hi (8 halves each), 4..7 = lo
            uint4* dst = reinterpret_cast<uint4*>(S.xs[xs][m] + rloc * 64);
#pragma unroll
            for (int c = 0; c < 4; c++) {
                const float a0 = v[2 * c].x * sf, a1 = v[2 * c].y * sf, a2 = v[2 * c].z * sf, a3 = v[2 * c].w * sf;
                const float a4 = v[2 * c + 1].x * sf, a5 = v[2 * c + 1].y * sf, a6 = v[2 * c + 1].z * sf, a7 = v[2 * c + 1].w * sf;
                const __half2 h0 = __floats2half2_rn(a0, a1), h1 = __floats2half2_rn(a2, a3), h2 = __floats2half2_rn(a4, a5), h3 = __floats2half2_rn(a6, a7);
                const float2 f0 = __half22float2(h0), f1 = __half22float2(h1), f2 = __half22float2(h2), f3 = __half22float2(h3);
                const __half2 l0 = __floats2half2_rn(a0 - f0.x, a1 - f0.y), l1 = __floats2half2_rn(a2 - f1.x, a3 - f1.y);
                const __half2 l2 = __floats2half2_rn(a4 - f2.x, a5 - f2.y), l3 = __floats2half2_rn(a6 - f3.x, a7 - f3.y);
                if (cp == 0)                                          // (two threads per row: one writes hi, the other lo)
                    dst[c ^ (rloc & 7)] = make_uint4(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1),
                                                     *reinterpret_cast<const uint32_t*>(&h2), *reinterpret_cast<const uint32_t*>(&h3));
                if (cp == CP - 1)
                    dst[(c + 4) ^ (rloc & 7)] = make_uint4(*reinterpret_cast<const uint32_t*>(&l0), *reinterpret_cast<const uint32_t*>(&l1),
                                                       *reinterpret_cast<const uint32_t*>(&l2), *reinterpret_cast<const uint32_t*>(&l3));
            }
            if (FOLD && cp == 0) {                                    // K-chunk 0 of my row: {p, p, p, 0, 0, 0, 0, 0}, p = 2^s_row in BF16
                const uint32_t pb = (uint32_t)(s_row + 127) << 7;
                reinterpret_cast<uint4*>(S.xe[xs][m])[(rloc >> 3) * 16 + (rloc & 7)] = make_uint4(pb | (pb << 16), pb, 0u, 0u);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic writes -> visible to the MMA (async proxy)
            __syncwarp();
            if (lane == 0) mbar_arrive(&S.x_ready[xs]);
            xn_next = xn; prime_next = prime; sf_next = sf; srow_next = s_row; force_next = force;
        };
        uint32_t it = 0, j = 0;
#ifdef SCKM_TC5H_TRACE
        const long long tep0 = clock64();
#endif
        if (blockIdx.x < nsuper) split_stage(0, blockIdx.x);
        const uint32_t split_at = nblocks > 4 ? 4u : nblocks - 1;    // late enough for the next raw tile to have landed
        for (uint64_t st = blockIdx.x; st < nsuper; st += gridDim.x, it++) {
            const int xs = it & 1;
            const uint64_t row = st * rows_per_super + (uint64_t)m * H_BM + rloc;
            const bool valid = row < n;
            const double xn = xn_next; const float prime = prime_next, sf = sf_next; const int s_row = srow_next; const bool force = force_next;
            // ---- running top-2 over all centroid blocks, in this row's scaled units ----
            float best = -FLT_MAX, second = -FLT_MAX; uint32_t bi = 0;
            for (uint32_t b = 0; b < nblocks; b++, j++) {
                const int ts = j & 1; const uint32_t ph = (j >> 1) & 1;
                // split the NEXT super-tile as soon as this one is under way, so the MMA warp never waits for it
                if (b == split_at && st + gridDim.x < nsuper) {
#ifdef SCKM_TC5H_TRACE
                    const long long ts0 = clock64();
#endif
                    split_stage(it + 1, st + gridDim.x);
#ifdef SCKM_TC5H_TRACE
                    if (threadIdx.x == 64) g_tc5h_trace[blockIdx.x * 8 + 5] += clock64() - ts0;
#endif
                }
#ifdef SCKM_TC5H_TRACE
                const long long tw0 = clock64();
#endif
                mbar_wait(&S.t_full[ts][m], ph);
#ifdef SCKM_TC5H_TRACE
                if (threadIdx.x == 64) g_tc5h_trace[blockIdx.x * 8 + 4] += clock64() - tw0;
#endif
                asm volatile("tcgen05.fence::after_thread_sync;");
                const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(ts * TSTAGE + m * H_BN + cp * COLS);
                const uint32_t col0 = b * H_BN + cp * COLS;
                const float4* h_gl = reinterpret_cast<const float4*>(hcn_sc + col0);
                uint32_t va[32], vb[32];
                auto consume = [&](const uint32_t (&v)[32], int c0) {
                    float s[32];
                    if (FOLD) {
#pragma unroll
                        for (int e = 0; e < 32; e++) s[e] = __uint_as_float(v[e]);
                    } else {
#pragma unroll
                        for (int u = 0; u < 8; u++) {
                            const float4 hv = __ldg(h_gl + (c0 >> 2) + u);
                            s[4 * u + 0] = fmaf(hv.x, sf, __uint_as_float(v[4 * u + 0])); s[4 * u + 1] = fmaf(hv.y, sf, __uint_as_float(v[4 * u + 1]));
                            s[4 * u + 2] = fmaf(hv.z, sf, __uint_as_float(v[4 * u + 2])); s[4 * u + 3] = fmaf(hv.w, sf, __uint_as_float(v[4 * u + 3]));
                        }
                    }
                    // Can this chunk change any row's best or second, or come within the tie margin of a winner?  The maxima
                    // of its four 8-column groups (four independent chains: the warp waits on latency here) answer that for
                    // the chunk, and then group by group: in the steady state ONE column of a chunk is above a row's bar
                    // (the row's own centroid), so the five-instruction top-2 update runs over 8 columns, not 32.
                    float mg[4];
#pragma unroll
                    for (int g = 0; g < 4; g++) {
                        mg[g] = fmaxf(fmaxf(s[8 * g], s[8 * g + 1]), s[8 * g + 2]);
                        mg[g] = fmaxf(fmaxf(mg[g], s[8 * g + 3]), s[8 * g + 4]);
                        mg[g] = fmaxf(fmaxf(mg[g], s[8 * g + 5]), s[8 * g + 6]);
                        mg[g] = fmaxf(mg[g], s[8 * g + 7]);
                    }
                    const float bar = fmaxf(second, prime);            // columns at or below it can neither win nor come within the margin
                    if (!__any_sync(0xffffffffu, fmaxf(fmaxf(mg[0], mg[1]), fmaxf(mg[2], mg[3])) > bar)) return;
#pragma unroll
                    for (int g = 0; g < 4; g++) {
                        if (!__any_sync(0xffffffffu, mg[g] > bar)) continue;
#pragma unroll
                        for (int e = 8 * g; e < 8 * g + 8; e++) {
                            const float sc = s[e];
                            const bool gt = sc > best;
                            second = fmaxf(second, gt ? best : sc);
                            bi = gt ? (col0 + c0 + e) : bi;
                            best = fmaxf(best, sc);
                        }
                    }
                };
                auto release_stage = [&]() {                           // all of this stage's columns are in registers
                    asm volatile("tcgen05.fence::before_thread_sync;");
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&S.t_empty[ts][m]);
                };
                if (COLS == 64) {
                    // both chunks of my part at once: the columns go back to the MMA warp before the first is ranked
                    tmem_ld32(taddr, va);
                    tmem_ld32(taddr + 32, vb);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    release_stage();
                    consume(va, 0);
                    consume(vb, 32);
                } else {
                    // software pipeline over the 32-column chunks: chunk c+1 is in flight while chunk c is ranked
                    tmem_ld32(taddr, va);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int c0 = 0; c0 < COLS; c0 += 64) {
                        tmem_ld32(taddr + c0 + 32, vb);
                        consume(va, c0);
                        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                        if (c0 + 64 < COLS) tmem_ld32(taddr + c0 + 64, va); else release_stage();
                        consume(vb, c0 + 32);
                        if (c0 + 64 < COLS) asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    }
                }
            }
            // ---- two column parts per row: merge their top-2 through shared memory (one barrier per tile and super-tile) ----
            if (CP == 2) {
                const int mb = it & 1;
                if (cp == 1) { S.m_best[mb][m][rloc] = best; S.m_second[mb][m][rloc] = second; S.m_idx[mb][m][rloc] = bi; }
                asm volatile("bar.sync %0, 256;" ::"r"(1 + m) : "memory");   // the 8 epilogue warps of this tile
                if (cp == 1) continue;                                 // the exact part belongs to part 0
                const float ob = S.m_best[mb][m][rloc], os = S.m_second[mb][m][rloc]; const uint32_t oi = S.m_idx[mb][m][rloc];
                const bool take = ob > best || (ob == best && oi < bi);
                second = fmaxf(fmaxf(second, os), fminf(best, ob));
                bi = take ? oi : bi;
                best = fmaxf(best, ob);
            }
#ifdef SCKM_TC5H_TRACE
            const long long te0 = clock64();
#endif
            // ---- decide: near-tie mark; exact f64 distance to the winner and the update, cooperatively per row ----
            // columns of skipped chunks lie at or below max(second, prime): that is the runner-up the tie test must assume
            second = fmaxf(second, prime);
            const double rs = h_pow2d(-(s_row + s_c));                 // scaled score -> real units (exact: a power of two)
            const double gap = 2.0 * ((double)best - (double)second) * rs;
            const bool tie = !(gap > H_TIE_REL * (xn + cmax)) || bi >= k || force;
            const bool part_ok = valid && !tie;
            const uint32_t lab = part_ok ? bi : 0xffffffffu;
            const uint64_t wrow0 = st * rows_per_super + (uint64_t)m * H_BM + (uint64_t)q * 32;   // first row of this warp
            // (a) update: the warp walks its 32 rows in order; lane f handles features f, f+32, ... and adds the row's
            // value to the warp's private partial with a fire-and-forget RED (an address only ever receives adds from
            // one thread, in program order => fixed summation order).  No dependent loads: nothing waits.
#pragma unroll 4
            for (int r = 0; r < 32; r++) {
                const uint32_t lr = __shfl_sync(0xffffffffu, lab, r);
                if (lr == 0xffffffffu) continue;                       // warp-uniform
                if ((uint32_t)lane < d) {                              // d <= 32: lane f handles feature f
                    double xv;
                    if (sizeof(TXS) == 8) {
                        xv = (double)__ldg(xsrc + (wrow0 + r) * d + lane);
                    } else {
                        const int rt = q * 32 + r;
                        xv = (double)S.xraw[xs][m][rt * 32 + ((((lane >> 2) ^ (rt & 7)) << 2) | (lane & 3))];   // undo the 128-byte swizzle
                    }
                    atomicAdd(part + (size_t)lr * d + lane, xv);
                }
            }
            // (b) exact f64 distance of my row to its winner: all loads independent (one memory round trip)
            double mydist = 0.0;
            if (part_ok) {
                const double* cr = centroids + (size_t)bi * d;
                double a0 = 0.0, a1 = 0.0;
                for (uint32_t c = 0; c < d; c += 4) {                  // d % 4 == 0
                    double xv[4];
                    if (sizeof(TXS) == 8) {
                        const double2 p0 = __ldg(reinterpret_cast<const double2*>(xsrc + row * d + c));
                        const double2 p1 = __ldg(reinterpret_cast<const double2*>(xsrc + row * d + c + 2));
                        xv[0] = p0.x; xv[1] = p0.y; xv[2] = p1.x; xv[3] = p1.y;
                    } else {
                        const float4 p = reinterpret_cast<const float4*>(S.xraw[xs][m] + rloc * 32)[(c >> 2) ^ (rloc & 7)];
                        xv[0] = p.x; xv[1] = p.y; xv[2] = p.z; xv[3] = p.w;
                    }
                    const double2 c0 = __ldg(reinterpret_cast<const double2*>(cr + c));
                    const double2 c1 = __ldg(reinterpret_cast<const double2*>(cr + c + 2));
                    const double d0 = xv[0] - c0.x, d1 = xv[1] - c0.y, d2 = xv[2] - c1.x, d3 = xv[3] - c1.y;
                    a0 = fma(d0, d0, a0); a1 = fma(d1, d1, a1); a0 = fma(d2, d2, a0); a1 = fma(d3, d3, a1);
                }
                mydist = a0 + a1;
            }
            if (valid) { labels[row] = tie ? 0xffffffffu : bi; mind[row] = mydist; if (tie) atomicAdd(nmarked, 1ull); }
            // counts: one add per distinct label of the warp (the lowest lane of each group adds the group size)
            unsigned lanemask_lt;
            asm("mov.u32 %0, %%lanemask_lt;" : "=r"(lanemask_lt));
            const unsigned peers = __match_any_sync(0xffffffffu, part_ok ? bi : (0x80000000u | (uint32_t)lane));
            if (part_ok && (peers & lanemask_lt) == 0) atomicAdd(part + (size_t)k * d + bi, (double)__popc(peers));
            double vsum = part_ok ? mydist : 0.0;                     // fixed-order sum over the warp's 32 rows
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) vsum = __dadd_rn(vsum, __shfl_xor_sync(0xffffffffu, vsum, o));
            if (lane == 0) atomicAdd(part + pk - 1, vsum);
            __syncwarp();
            if (lane == 0) mbar_arrive(&S.raw_empty[xs]);            // the raw tile may be overwritten by super-tile it + 2
#ifdef SCKM_TC5H_TRACE
            if (threadIdx.x == 64) g_tc5h_trace[blockIdx.x * 8 + 6] += clock64() - te0;
#endif
        }
#ifdef SCKM_TC5H_TRACE
        if (threadIdx.x == 64) g_tc5h_trace[blockIdx.x * 8 + 7] += clock64() - tep0;
#endif
    }
    // ---- teardown ----
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS));
}

// ---- host side --------------------------------------------------------------------------------------------------
static int make_map_h(sckm_ctx* ctx, CUtensorMap* map, CUtensorMapDataType dt, uint32_t esize, const void* base, uint64_t rows,
                      uint64_t cols, uint32_t box_cols, uint32_t box_rows) {
    EncodeTiledFn enc = encode_fn();
    if (!enc) return fail(ctx, SCKM_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    cuuint64_t dims[2] = {cols, rows}, strides[1] = {cols * esize};
    cuuint32_t box[2] = {box_cols, box_rows}, estr[2] = {1, 1};
    CUresult r = enc(map, dt, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(ctx, SCKM_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) for %llu x %llu", (int)r,
                                       (unsigned long long)rows, (unsigned long long)cols);
    return SCKM_OK;
}

bool tc5h_supported(const sckm_dataset* ds, uint64_t k) {
    return ds->d >= 4 && ds->d <= 32 && ds->d % 4 == 0 && k >= 16 && k <= (1u << 20) && ds->n < 0x7FFFFFFFull &&
           encode_fn() != nullptr && !getenv("SCKM_TC5_TF32");
}

template <bool FOLD, int EW, int KS, typename TXS>
static int launch_tc5h_t(sckm_dataset* ds, uint64_t k, size_t pk, const float* x32) {
    sckm_ctx* ctx = ds->ctx;
    const unsigned grid = (unsigned)ctx->num_sms;
    const uint32_t nblocks = (uint32_t)((k + H_BN - 1) / H_BN), kpad = nblocks * H_BN;
    // per padded centroid: 64 halves (hi | lo) + 32 floats (f32 copy) + 16 bf16 (rank-one operand) + 1 float (scaled norm)
    const size_t need = (size_t)kpad * 80;                             // in floats: 32 + 32 + 8 + 1, rounded up
    if (need > ctx->cap_tc5) {
        if (ctx->d_tc5) { SCKM_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); cudaFree(ctx->d_tc5); ctx->d_tc5 = nullptr; ctx->cap_tc5 = 0; }
        SCKM_CUDA(ctx, cudaMalloc((void**)&ctx->d_tc5, need * sizeof(float)));
        ctx->cap_tc5 = need;
    }
    __half* chl = reinterpret_cast<__half*>(ctx->d_tc5);
    float* c32 = ctx->d_tc5 + (size_t)kpad * 32;
    __nv_bfloat16* ce_g = reinterpret_cast<__nv_bfloat16*>(ctx->d_tc5 + (size_t)kpad * 64);
    float* hcn_sc = ctx->d_tc5 + (size_t)kpad * 72;
    tc5h_prep_kernel<<<nblocks, 256, 0, ctx->stream>>>(ctx->d_centroids, ctx->d_cnorm, (uint32_t)k, (uint32_t)ds->d, chl, c32, hcn_sc, ce_g);
    LAUNCH_CHECK_H(ctx);
    CUtensorMap mapX, mapC;
    SCKM_TRY(make_map_h(ctx, &mapX, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, x32, ds->n, ds->d, 32, H_BM));
    SCKM_TRY(make_map_h(ctx, &mapC, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, chl, kpad, 64, 64, H_BN));
    const size_t smem = sizeof(HSmem) + 1024;
    auto kern = assign_tc5h_kernel<FOLD, EW, KS, TXS>;
    SCKM_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, (EW + 3) * 32, smem, ctx->stream>>>(mapX, mapC, (const TXS*)ds->x, ds->n, (uint32_t)ds->d, ctx->d_centroids, ctx->d_cnorm,
                                                 hcn_sc, c32, ce_g, (ds->have_labels && !getenv("SCKM_TC5_NOPRIME")) ? ds->labels : nullptr,
                                                 (uint32_t)k, nblocks, ds->labels, ds->mind, ctx->d_partials, pk, ctx->d_flags,
                                                 SCKM_LOOP_ARGS(ctx));
    LAUNCH_CHECK_H(ctx);
    return SCKM_OK;
}

// called by launch_assign_tc5 (sckm_tc5.cu) after the workspaces, norms and the f32 shadow are in place
int launch_tc5h(sckm_dataset* ds, uint64_t k, size_t pk, const float* x32) {
    const bool fold = !getenv("SCKM_TC5H_NOFOLD");
    // 16 epilogue warps (two threads per row) measured SLOWER in a sustained run (17.7 vs 16.9 ms per 10M-row step: more
    // instructions for the same work under the power cap), so one thread per row stays the default; SCKM_TC5H_EW16=1 for A/B
    const bool ew8 = getenv("SCKM_TC5H_EW16") == nullptr;
    const bool k1 = ds->d <= 16;                                      // one 16-column K-step per product
    if (ds->dtype == SCKM_F64) {
        if (!fold) return k1 ? launch_tc5h_t<false, 8, 1, double>(ds, k, pk, x32) : launch_tc5h_t<false, 8, 2, double>(ds, k, pk, x32);
        if (!ew8) return k1 ? launch_tc5h_t<true, 16, 1, double>(ds, k, pk, x32) : launch_tc5h_t<true, 16, 2, double>(ds, k, pk, x32);
        return k1 ? launch_tc5h_t<true, 8, 1, double>(ds, k, pk, x32) : launch_tc5h_t<true, 8, 2, double>(ds, k, pk, x32);
    }
    if (!fold) return k1 ? launch_tc5h_t<false, 8, 1, float>(ds, k, pk, x32) : launch_tc5h_t<false, 8, 2, float>(ds, k, pk, x32);
    if (!ew8) return k1 ? launch_tc5h_t<true, 16, 1, float>(ds, k, pk, x32) : launch_tc5h_t<true, 16, 2, float>(ds, k, pk, x32);
    return k1 ? launch_tc5h_t<true, 8, 1, float>(ds, k, pk, x32) : launch_tc5h_t<true, 8, 2, float>(ds, k, pk, x32);
}

}  // namespace sckm
