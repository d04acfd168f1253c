// sckm_tc5.cu -- K2: Lloyd assignment for f32 data on the 5th-generation tensor cores (tcgen05), 3xTF32.
//
// Replaces the per-iteration work of BBDTree::clustering (src/algorithm/neighbour/bbd_tree.rs:62-163) for
// TX = f32 with d <= 32: scores  x.c_j - ||c_j||^2/2  are computed as  Xh.Ch + Xh.Cl + Xl.Ch  (x = xh + xl, both
// TF32-representable, the split is exact) with FP32 accumulation in tensor memory; the decision is then made exact:
// rows whose best/second gap is within the error bound of that arithmetic are marked and re-decided by
// refine_rows_kernel in the reference's f64 arithmetic, and every row's distance to its centroid (inertia) is
// recomputed in f64.  So labels equal the exact direct-form argmin; only the *ranking* uses reduced precision.
//
// One persistent CTA per SM, warp-specialised:
//   warp 0   TMA producer for X: 256-row super-tiles (two 128 x 32 f32 boxes, SWIZZLE_128B, zero-filled out of bounds)
//   warp 10  TMA producer for the centroid blocks (128 centroids: hi and lo parts, prepared once per step)
//   warp 1   allocates the 512 TMEM columns and issues tcgen05.mma.kind::tf32 (one thread): per centroid block
//            2 tiles x 3 products x 4 k-steps of 128x128x8, accumulators double-buffered in TMEM (2 x 256 columns)
//   warps 2-9  one thread per row of the super-tile: split the landed X tile into hi/lo in place, then per centroid
//            block tcgen05.ld the row's 128 scores and keep a running top-2 + argmax in registers; at the end the
//            exact f64 distance, the near-tie mark, and the deterministic fused update (same scheme as sckm_dmma.cu:
//            fire-and-forget RED.ADD.F64 into the warp's private partial, same-label rows serialised by rank).
//            The top-2 tracking (five half-rate ALU instructions per score) is what bounds this kernel, so most of it
//            is skipped: every row is PRIMED with the score of the centroid it was assigned to in the previous step
//            (one 32-term dot product), and a 32-column chunk only goes through the tracking when, for some row of the
//            warp, its maximum (one FADD + half an FMNMX3 per score) exceeds max(second best so far, primed score -
//            2.5 tie margins) -- columns below that can neither win nor come within the tie margin of the winner.
//            After a few Lloyd steps few rows change cluster: a warp then tracks little more than the chunks that hold
//            its 32 rows' own centroids.
// mbarrier rings: x_full/x_ready/x_empty (2 stages), c_full/c_empty (2 stages), t_full/t_empty (2 TMEM stages).
#include "sckm_common.cuh"
#include "sckm_tile.cuh"
#include "sckm_umma.cuh"
#include <cuda.h>
#include <cfloat>
#include <algorithm>
#include <cstdlib>

namespace sckm {

#define LAUNCH_CHECK_T(ctx)                                                                        \
    do {                                                                                           \
        (ctx)->launches++;                                                                         \
        cudaError_t _e = cudaGetLastError();                                                       \
        if (_e != cudaSuccess)                                                                     \
            return fail((ctx), SCKM_ERR_CUDA, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), \
                        __FILE__, __LINE__);                                                       \
    } while (0)

int launch_refine_rows(sckm_dataset* ds, uint64_t k, size_t pk, unsigned grid_ctas);   // sckm_dmma.cu
int launch_cnorm(sckm_ctx* ctx, uint64_t k, uint64_t d, bool center);                  // sckm_dmma.cu
bool tc5h_supported(const sckm_dataset* ds, uint64_t k);                               // sckm_tc5h.cu
int launch_tc5h(sckm_dataset* ds, uint64_t k, size_t pk, const float* x32);            // sckm_tc5h.cu

constexpr int TC_BM = 128;               // rows per MMA tile (TMEM lanes)
constexpr int TC_EPI_WARPS = 8;
constexpr int TC_THREADS = 11 * 32;      // warps: 0 X-producer, 1 MMA, 2..9 epilogue, 10 C-producer
constexpr uint32_t TC_ATOM_FLOATS = TC_BM * 32;        // one 128-row x 128-byte swizzle atom of f32
constexpr double TC_TIE_REL = 2e-5;      // >= 10x the 3xTF32 + FP32-accumulate error bound (bench/tc5_probe.cu: 1e-6)


// NK: 128-byte swizzle atoms along K (d <= 32*NK); TILES: 128-row tiles per super-tile; BN: centroids per block.
// The 8 epilogue warps cover TILES tiles x 4 TMEM lane quadrants x CP column parts (CP = 2 / TILES).
template <int NK, int TILES, int BN>
struct alignas(16) TcSmemT {                  // dynamic shared memory image (base aligned to 1024 B)
    float xh[2][TILES][NK][TC_ATOM_FLOATS];   // raw rows on arrival, TF32 hi part after the split
    float xl[2][TILES][NK][TC_ATOM_FLOATS];   // TF32 lo part
    float ch[2][NK][BN * 32];                 // centroid block, hi
    float cl[2][NK][BN * 32];                 // centroid block, lo
    uint64_t x_full[2], x_ready[2], x_empty[2], c_full[2], c_empty[2], t_full[2], t_empty[2];
    double m_xn[2][TC_BM];                    // column-part merge scratch (CP == 2), double-buffered by super-tile
    float m_best[2][TC_BM], m_second[2][TC_BM];
    uint32_t m_idx[2][TC_BM];
    uint32_t tmem_base;
};

// centroids (f64 [k][d]) -> TF32 hi / lo parts in f32 [kpad][kdim] (zero padded) and -||c||^2/2 in f32
__global__ void tc5_prep_kernel(const double* __restrict__ centroids, const double* __restrict__ cnorm, uint32_t k, uint32_t d,
                                uint32_t kpad, uint32_t kdim, float* __restrict__ ch, float* __restrict__ cl,
                                float* __restrict__ hcn) {
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e < kpad * kdim) {
        const uint32_t r = e / kdim, c = e - r * kdim;
        float v = (r < k && c < d) ? (float)centroids[(size_t)r * d + c] : 0.f;
        const float h = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
        ch[e] = h; cl[e] = v - h;
    }
    if (e < kpad) hcn[e] = e < k ? (float)(-0.5 * cnorm[e]) : -INFINITY;
}

__global__ void tc5_shadow_kernel(const double* __restrict__ x, float* __restrict__ x32, uint64_t total) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x)
        x32[i] = (float)x[i];
}

// HCN_SMEM: -||c||^2/2 of all centroids resident in shared memory, else read through L1.
// TXS: type of the rows used for the exact part (distance to the winner, sums): float = the staged tile itself,
// double = the original f64 rows in HBM (the tensor cores then only see an f32 shadow copy, for the ranking).
template <int NK, int TILES, int BN, bool HCN_SMEM, typename TXS>
__global__ void __launch_bounds__(TC_THREADS, 1)
assign_tc5_kernel(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapCh,
                  const __grid_constant__ CUtensorMap mapCl, const TXS* __restrict__ xsrc, uint64_t n, uint32_t d,
                  const double* __restrict__ centroids, const double* __restrict__ cnorm, const float* __restrict__ hcn,
                  const float* __restrict__ ch_g, const float* __restrict__ cl_g, const uint32_t* prev_labels,
                  uint32_t k, uint32_t nblocks, uint32_t* labels, double* __restrict__ mind,
                  double* __restrict__ partials, size_t pk, unsigned long long* __restrict__ nmarked,
                  const LoopState* __restrict__ loop_st, uint32_t loop_it) {
    if (loop_done(loop_st, loop_it)) return;                          // the fit's stop rule already fired (kmeans.rs:305)
    using Smem = TcSmemT<NK, TILES, BN>;
    constexpr int CP = 2 / TILES;                    // column parts per row
    constexpr int COLS = BN / CP;                    // columns per epilogue thread per block
    constexpr int TSTAGE = TILES * BN;               // TMEM columns per accumulator stage
    constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
    static_assert(COLS % 32 == 0, "epilogue threads read TMEM in 32-column chunks");
    extern __shared__ unsigned char smem_raw[];
    Smem& S = *reinterpret_cast<Smem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint64_t rows_per_super = (uint64_t)TILES * TC_BM;
    const uint64_t nsuper = (n + rows_per_super - 1) / rows_per_super;
    const double cmax = cta_max(cnorm, k);                            // max_j ||c_j||^2 (all threads take part)
    float* s_hcn = reinterpret_cast<float*>(&S + 1);                 // [nblocks * BN] when HCN_SMEM
    if (HCN_SMEM) for (uint32_t i = threadIdx.x; i < nblocks * BN; i += blockDim.x) s_hcn[i] = hcn[i];

    if (threadIdx.x == 0) {
        for (int s = 0; s < 2; s++) {
            mbar_init(&S.x_full[s], 1); mbar_init(&S.x_ready[s], TC_EPI_WARPS); mbar_init(&S.x_empty[s], 1 + TC_EPI_WARPS);
            mbar_init(&S.c_full[s], 1); mbar_init(&S.c_empty[s], 1);
            mbar_init(&S.t_full[s], 1); mbar_init(&S.t_empty[s], TC_EPI_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    constexpr uint32_t TMEM_COLS = 2 * TSTAGE < 32 ? 32 : 2 * TSTAGE;   // 128 / 256 / 512: a power of two >= 32
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&S.tmem_base)), "r"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = S.tmem_base;

    if (warp == 0) {
        // ================= TMA producer: X super-tiles =================
        if (lane == 0) {
            uint32_t it = 0;
            for (uint64_t st = blockIdx.x; st < nsuper; st += gridDim.x, it++) {
                const int xs = it & 1; const uint32_t ph = (it >> 1) & 1;
                mbar_wait(&S.x_empty[xs], ph ^ 1);
                mbar_expect_tx(&S.x_full[xs], TILES * NK * TC_ATOM_FLOATS * 4);
                for (int m = 0; m < TILES; m++)
                    for (int a = 0; a < NK; a++)
                        tma_load_2d(S.xh[xs][m][a], &mapX, 32 * a, (int)(st * rows_per_super + (uint64_t)m * TC_BM), &S.x_full[xs]);
            }
        }
    } else if (warp == 10) {
        // ================= TMA producer: centroid blocks =================
        if (lane == 0) {
            uint32_t j = 0;
            for (uint64_t st = blockIdx.x; st < nsuper; st += gridDim.x)
                for (uint32_t b = 0; b < nblocks; b++, j++) {
                    const int cs = j & 1; const uint32_t ph = (j >> 1) & 1;
                    mbar_wait(&S.c_empty[cs], ph ^ 1);
                    mbar_expect_tx(&S.c_full[cs], 2 * NK * BN * 128);
                    for (int a = 0; a < NK; a++) {
                        tma_load_2d(S.ch[cs][a], &mapCh, 32 * a, (int)(b * BN), &S.c_full[cs]);
                        tma_load_2d(S.cl[cs][a], &mapCl, 32 * a, (int)(b * BN), &S.c_full[cs]);
                    }
                }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        // The whole warp walks the loops and waits on the barriers, one elected lane issues: the operands are then
        // warp-uniform (uniform registers, no per-MMA election loop), and the loops over products / atoms / K-steps are
        // unrolled with the descriptors as (low, high) words -- a stage or K-step change is one 32-bit add (see sckm_tc5h.cu)
        {
            uint32_t it = 0, j = 0;
            constexpr uint32_t SW128_HI = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);
            constexpr uint32_t ATOM16 = (uint32_t)(TC_ATOM_FLOATS * 4) >> 4, CATOM16 = (uint32_t)(BN * 32 * 4) >> 4;   // tile sizes in 16-byte units
            const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
            const uint32_t xh_lo0 = ((smem_u32(&S.xh[0][0][0][0]) >> 4) & 0x3FFF) | (1u << 16);
            const uint32_t xl_lo0 = ((smem_u32(&S.xl[0][0][0][0]) >> 4) & 0x3FFF) | (1u << 16);
            const uint32_t ch_lo0 = ((smem_u32(&S.ch[0][0][0]) >> 4) & 0x3FFF) | (1u << 16);
            const uint32_t cl_lo0 = ((smem_u32(&S.cl[0][0][0]) >> 4) & 0x3FFF) | (1u << 16);
            for (uint64_t st = blockIdx.x; st < nsuper; st += gridDim.x, it++) {
                const int xs = it & 1; const uint32_t xph = (it >> 1) & 1;
                mbar_wait(&S.x_ready[xs], xph);                      // hi/lo split done by the epilogue warps
                asm volatile("tcgen05.fence::after_thread_sync;");
                for (uint32_t b = 0; b < nblocks; b++, j++) {
                    const int cs = j & 1; const uint32_t ph = (j >> 1) & 1;   // centroid stage == TMEM stage index
                    mbar_wait(&S.c_full[cs], ph);
                    mbar_wait(&S.t_empty[cs], ph ^ 1);
                    asm volatile("tcgen05.fence::after_thread_sync;");
                    if (elect_one()) {
#pragma unroll
                        for (int m = 0; m < TILES; m++) {
                            const uint32_t tcol = tmem_u + (uint32_t)(cs * TSTAGE + m * BN);
#pragma unroll
                            for (int prod = 0; prod < 3; prod++)          // Xh.Ch + Xh.Cl + Xl.Ch
#pragma unroll
                                for (int a = 0; a < NK; a++) {
                                    const uint32_t xlo = (prod == 2 ? xl_lo0 : xh_lo0) + (uint32_t)((xs * TILES + m) * NK + a) * ATOM16;
                                    const uint32_t clo = (prod == 1 ? cl_lo0 : ch_lo0) + (uint32_t)(cs * NK + a) * CATOM16;
#pragma unroll
                                    for (int ks = 0; ks < 4; ks++)
                                        umma_tf32_parts(tcol, xlo + 2 * ks, SW128_HI, clo + 2 * ks, SW128_HI, IDESC, (prod | a | ks) != 0);
                                }
                        }
                        umma_commit(&S.c_empty[cs]);                 // centroid stage free once these MMAs have read it
                        umma_commit(&S.t_full[cs]);                  // accumulators ready for the epilogue
                        if (b + 1 == nblocks) umma_commit(&S.x_empty[xs]);   // X stage: MMA side done (epilogue warps arrive too)
                    }
                    __syncwarp();
                }
            }
        }
    } else {
        // ================= epilogue warps: one thread per (row, column part) =================
        const int ew = warp - 2;                                     // 0..7
        const int m = TILES == 2 ? (ew >> 2) : 0;                    // tile of the super-tile
        const int cp = TILES == 2 ? 0 : (ew >> 2);                   // column part
        const int q = warp & 3;                                      // TMEM lane quadrant this warp may access
        const int rloc = q * 32 + lane;                              // row within the tile
        double* part = partials + ((size_t)blockIdx.x * TC_EPI_WARPS + ew) * ((pk + 15) / 16 * 16);
        // split my share of X stage `xs` in place (a row is 128 bytes at rloc*128 inside each atom; the swizzle only
        // permutes 16-byte chunks inside it), return my part of ||x||^2, and tell the MMA warp the stage is ready
        auto split_stage = [&](int xs, uint32_t xph) -> double {
            mbar_wait(&S.x_full[xs], xph);
            double xn = 0.0;
#pragma unroll
            for (int a = cp; a < NK; a += CP) {
                float4* ph4 = reinterpret_cast<float4*>(S.xh[xs][m][a] + rloc * 32);
                float4* pl4 = reinterpret_cast<float4*>(S.xl[xs][m][a] + rloc * 32);
#pragma unroll
                for (int c0 = 0; c0 < 8; c0++) {
                    const int c = (c0 + lane) & 7;                     // rotate: 8 lanes hit 8 different 16-byte columns
                    float4 v = ph4[c], h, l;
                    h.x = __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u); l.x = v.x - h.x;
                    h.y = __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u); l.y = v.y - h.y;
                    h.z = __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u); l.z = v.z - h.z;
                    h.w = __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u); l.w = v.w - h.w;
                    ph4[c] = h; pl4[c] = l;
                    xn = fma((double)v.x, (double)v.x, xn); xn = fma((double)v.y, (double)v.y, xn);
                    xn = fma((double)v.z, (double)v.z, xn); xn = fma((double)v.w, (double)v.w, xn);
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic writes -> visible to the MMA (async proxy)
            __syncwarp();
            if (lane == 0) mbar_arrive(&S.x_ready[xs]);
            return xn;
        };
        uint32_t it = 0, j = 0;
        double xn_next = blockIdx.x < nsuper ? split_stage(0, 0) : 0.0;
        // CP == 2: a row's two atoms are split by two different warps, and the priming below reads both.  From the second
        // super-tile on the merge barrier at the end of the previous one orders that; the first needs its own (racecheck)
        if (CP == 2 && blockIdx.x < nsuper) asm volatile("bar.sync 1, 256;" ::: "memory");
        for (uint64_t st = blockIdx.x; st < nsuper; st += gridDim.x, it++) {
            const int xs = it & 1;
            const uint64_t row = st * rows_per_super + (uint64_t)m * TC_BM + rloc;
            const bool valid = row < n;
            double xn = xn_next;
            // ---- priming: a lower bound of this row's winning score from the centroid it had in the previous step ----
            float prime = -FLT_MAX;
            if (prev_labels != nullptr) {
                prime = FLT_MAX;                                       // rows past the end never ask for a scan
                if (valid) {
                    prime = -FLT_MAX;
                    const uint32_t pl = prev_labels[row];
                    if (pl < k) {
                        float dot = 0.f, xx = 0.f;
#pragma unroll
                        for (int a = 0; a < NK; a++) {
                            const float4* c_h = reinterpret_cast<const float4*>(ch_g + (size_t)pl * (32 * NK) + 32 * a);
                            const float4* c_l = reinterpret_cast<const float4*>(cl_g + (size_t)pl * (32 * NK) + 32 * a);
#pragma unroll
                            for (int c4 = 0; c4 < 8; c4++) {
                                const float4 h = reinterpret_cast<const float4*>(S.xh[xs][m][a] + rloc * 32)[c4 ^ (rloc & 7)];
                                const float4 l = reinterpret_cast<const float4*>(S.xl[xs][m][a] + rloc * 32)[c4 ^ (rloc & 7)];
                                const float4 ph = __ldg(c_h + c4), pw = __ldg(c_l + c4);
                                const float x0 = h.x + l.x, x1 = h.y + l.y, x2 = h.z + l.z, x3 = h.w + l.w;    // exact: x = hi + lo
                                dot = fmaf(x0, ph.x + pw.x, dot); dot = fmaf(x1, ph.y + pw.y, dot);
                                dot = fmaf(x2, ph.z + pw.z, dot); dot = fmaf(x3, ph.w + pw.w, dot);
                                xx = fmaf(x0, x0, xx); xx = fmaf(x1, x1, xx); xx = fmaf(x2, x2, xx); xx = fmaf(x3, x3, xx);
                            }
                        }
                        // 2.5 tie margins in score space (the tie test below works on 2 * (best - second)); the score the
                        // tensor cores produce for that centroid differs from `dot` by ~1e-6 (xx + cmax), a tenth of one margin
                        const float p = dot + __ldg(hcn + pl) - 2.5f * (float)(0.5 * TC_TIE_REL) * (xx + (float)cmax);
                        prime = p == p ? p : -FLT_MAX;                 // NaN centroid: no priming
                    }
                }
            }
            // ---- running top-2 over all centroid blocks (my columns only) ----
            float best = -FLT_MAX, second = -FLT_MAX; uint32_t bi = 0;
            for (uint32_t b = 0; b < nblocks; b++, j++) {
                const int ts = j & 1; const uint32_t ph = (j >> 1) & 1;
                // split the NEXT super-tile as soon as this one is under way, so the MMA warp never waits for it
                if (b == (nblocks > 1 ? 1u : 0u) && st + gridDim.x < nsuper) xn_next = split_stage((it + 1) & 1, ((it + 1) >> 1) & 1);
                mbar_wait(&S.t_full[ts], ph);
                asm volatile("tcgen05.fence::after_thread_sync;");
                const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(ts * TSTAGE + m * BN + cp * COLS);
                const uint32_t col0 = b * BN + cp * COLS;
                // -||c||^2/2 of this thread's columns, four at a time: an explicit ld.shared when the table is resident (the
                // compiler cannot tell the address space of `HCN_SMEM ? s_hcn : hcn` and emitted generic LD.E.128, whose
                // latency showed up as FADD stalls all over the chunk loops)
                const uint32_t h_sm = smem_u32(s_hcn + col0);
                const float4* h_gl = reinterpret_cast<const float4*>(hcn + col0);
                auto hcn4 = [&](int i) -> float4 {
                    if (HCN_SMEM) {
                        float4 r;
                        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(h_sm + 16u * (uint32_t)i));
                        return r;
                    }
                    return __ldg(h_gl + i);
                };
                // software pipeline over the 32-column chunks: chunk c+1 is in flight (tcgen05.ld is asynchronous
                // until tcgen05.wait::ld) while chunk c goes through the top-2 update
                uint32_t va[32], vb[32];
                auto consume = [&](const uint32_t (&v)[32], int c0) {
                    {   // can this chunk change any row's best or second, or come within the tie margin of a winner?
                        float mx = -FLT_MAX;
#pragma unroll
                        for (int u = 0; u < 8; u++) {
                            const float4 hv = hcn4((c0 >> 2) + u);
                            mx = fmaxf(fmaxf(mx, __uint_as_float(v[u * 4 + 0]) + hv.x), __uint_as_float(v[u * 4 + 1]) + hv.y);
                            mx = fmaxf(fmaxf(mx, __uint_as_float(v[u * 4 + 2]) + hv.z), __uint_as_float(v[u * 4 + 3]) + hv.w);
                        }
                        if (!__any_sync(0xffffffffu, mx > fmaxf(second, prime))) return;
                    }
                    // (four independent top-2 chains per chunk were tried: no gain -- the warp is not bound by this chain)
#pragma unroll
                    for (int u = 0; u < 8; u++) {
                        const float4 hv = hcn4((c0 >> 2) + u);
                        const float hh[4] = {hv.x, hv.y, hv.z, hv.w};
#pragma unroll
                        for (int e = 0; e < 4; e++) {
                            const float sc = __uint_as_float(v[u * 4 + e]) + hh[e];
                            const bool gt = sc > best;
                            second = fmaxf(second, gt ? best : sc);
                            bi = gt ? (col0 + c0 + u * 4 + e) : bi;
                            best = fmaxf(best, sc);
                        }
                    }
                };
                auto release_stage = [&]() {                           // all of this stage's columns are in registers
                    asm volatile("tcgen05.fence::before_thread_sync;");
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&S.t_empty[ts]);
                };
                tmem_ld32(taddr, va);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (COLS == 32) {
                    release_stage();
                    consume(va, 0);
                } else {
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
            // ---- CP == 2: the two column parts of a row live in different warps: merge through shared memory ----
            if (CP == 2) {
                const int mb = it & 1;
                if (cp == 1) { S.m_best[mb][rloc] = best; S.m_second[mb][rloc] = second; S.m_idx[mb][rloc] = bi; S.m_xn[mb][rloc] = xn; }
                asm volatile("bar.sync 1, 256;" ::: "memory");       // the 8 epilogue warps
                if (cp == 0) {
                    const float ob = S.m_best[mb][rloc], os = S.m_second[mb][rloc]; const uint32_t oi = S.m_idx[mb][rloc];   // (both parts used the same `prime`)
                    xn += S.m_xn[mb][rloc];
                    const bool take = ob > best || (ob == best && oi < bi);
                    second = fmaxf(fmaxf(second, os), fminf(best, ob));
                    bi = take ? oi : bi;
                    best = fmaxf(best, ob);
                }
            }
            if (cp == 0) {
                // ---- decide: near-tie mark; exact f64 distance to the winner and the update, cooperatively per row ----
                // columns of skipped chunks lie at or below max(second, prime): that is the runner-up the tie test must assume
                second = fmaxf(second, prime);
                const double gap = 2.0 * ((double)best - (double)second);
                const bool tie = !(gap > TC_TIE_REL * (xn + cmax)) || bi >= k;
                const bool part_ok = valid && !tie;
                const uint32_t lab = part_ok ? bi : 0xffffffffu;
                const uint64_t wrow0 = st * rows_per_super + (uint64_t)m * TC_BM + (uint64_t)q * 32;   // first row of this warp
                // (a) update: the warp walks its 32 rows in order; lane f handles features f, f+32, ... and adds the
                // row's value to the warp's private partial with a fire-and-forget RED (an address only ever receives
                // adds from one thread, in program order => fixed summation order).  No dependent loads: nothing waits.
#pragma unroll 4
                for (int r = 0; r < 32; r++) {
                    const uint32_t lr = __shfl_sync(0xffffffffu, lab, r);
                    if (lr == 0xffffffffu) continue;                   // warp-uniform
                    for (uint32_t f = lane; f < d; f += 32) {
                        double xv;
                        if (sizeof(TXS) == 8) {
                            xv = (double)__ldg(xsrc + (wrow0 + r) * d + f);
                        } else {
                            const int a = f >> 5, fi = f & 31, rt = q * 32 + r;
                            const int phys = rt * 32 + ((((fi >> 2) ^ (rt & 7)) << 2) | (fi & 3));   // undo the 128-byte swizzle
                            xv = (double)S.xh[xs][m][a][phys] + (double)S.xl[xs][m][a][phys];
                        }
                        atomicAdd(part + (size_t)lr * d + f, xv);
                    }
                }
                // (b) exact f64 distance of my row to its winner: all loads independent (one memory round trip)
                double mydist = 0.0;
                if (part_ok) {
                    const double* cr = centroids + (size_t)bi * d;
                    double a0 = 0.0, a1 = 0.0;
                    for (uint32_t c = 0; c < d; c += 4) {              // d % 4 == 0
                        double xv[4];
                        if (sizeof(TXS) == 8) {
                            const double2 p0 = __ldg(reinterpret_cast<const double2*>(xsrc + row * d + c));
                            const double2 p1 = __ldg(reinterpret_cast<const double2*>(xsrc + row * d + c + 2));
                            xv[0] = p0.x; xv[1] = p0.y; xv[2] = p1.x; xv[3] = p1.y;
                        } else {
                            const int a = c >> 5, ch4 = (c & 31) >> 2;
                            const float4 h = reinterpret_cast<const float4*>(S.xh[xs][m][a] + rloc * 32)[ch4 ^ (rloc & 7)];
                            const float4 l = reinterpret_cast<const float4*>(S.xl[xs][m][a] + rloc * 32)[ch4 ^ (rloc & 7)];
                            xv[0] = (double)h.x + (double)l.x; xv[1] = (double)h.y + (double)l.y;
                            xv[2] = (double)h.z + (double)l.z; xv[3] = (double)h.w + (double)l.w;
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
                double v = part_ok ? mydist : 0.0;                    // fixed-order sum over the warp's 32 rows
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v = __dadd_rn(v, __shfl_xor_sync(0xffffffffu, v, o));
                if (lane == 0) atomicAdd(part + pk - 1, v);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&S.x_empty[xs]);
        }
    }
    // ---- teardown ----
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS));
}


static int make_map(sckm_ctx* ctx, CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
    EncodeTiledFn enc = encode_fn();
    if (!enc) return fail(ctx, SCKM_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    cuuint64_t dims[2] = {cols, rows}, strides[1] = {cols * 4};
    cuuint32_t box[2] = {32, box_rows}, estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(ctx, SCKM_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) for %llu x %llu", (int)r,
                                       (unsigned long long)rows, (unsigned long long)cols);
    return SCKM_OK;
}

// f32 data: d <= 64.  f64 data only on explicit request (SCKM_ASSIGN_TC5): the tensor cores then rank an f32 shadow
// copy of X and everything that reaches the result (decision of near-ties, distances, sums) is still exact f64.
bool tc5_supported(const sckm_dataset* ds, uint64_t k) {
    return ds->d >= 4 && ds->d <= 64 && ds->d % 4 == 0 && k >= 16 && k <= (1u << 20) && ds->n < 0x7FFFFFFFull &&
           encode_fn() != nullptr;
}
bool tc5_auto(const sckm_dataset* ds, uint64_t k) { return ds->dtype == SCKM_F32 && tc5_supported(ds, k); }

template <int NK, int TILES, int BN, typename TXS>
static int launch_tc5_t(sckm_dataset* ds, uint64_t k, size_t pk, const float* x32) {
    sckm_ctx* ctx = ds->ctx;
    const unsigned grid = (unsigned)ctx->num_sms;
    const uint32_t kdim = 32 * NK;
    const uint32_t nblocks = (uint32_t)((k + BN - 1) / BN), kpad = nblocks * BN;
    const size_t need = (size_t)kpad * kdim * 2 + kpad;               // ch | cl | hcn   (floats)
    if (need > ctx->cap_tc5) {
        if (ctx->d_tc5) { SCKM_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); cudaFree(ctx->d_tc5); ctx->d_tc5 = nullptr; }
        SCKM_CUDA(ctx, cudaMalloc((void**)&ctx->d_tc5, need * sizeof(float)));
        ctx->cap_tc5 = need;
    }
    float* ch = ctx->d_tc5; float* cl = ch + (size_t)kpad * kdim; float* hcn = cl + (size_t)kpad * kdim;
    tc5_prep_kernel<<<(kpad * kdim + 255) / 256, 256, 0, ctx->stream>>>(ctx->d_centroids, ctx->d_cnorm, (uint32_t)k, (uint32_t)ds->d,
                                                                     kpad, kdim, ch, cl, hcn);
    LAUNCH_CHECK_T(ctx);
    CUtensorMap mapX, mapCh, mapCl;
    SCKM_TRY(make_map(ctx, &mapX, x32, ds->n, ds->d, TC_BM));
    SCKM_TRY(make_map(ctx, &mapCh, ch, kpad, kdim, BN));
    SCKM_TRY(make_map(ctx, &mapCl, cl, kpad, kdim, BN));
    using Smem = TcSmemT<NK, TILES, BN>;
    const size_t hcn_bytes = (size_t)kpad * sizeof(float);
    const bool hcn_smem = sizeof(Smem) + 1024 + hcn_bytes <= (size_t)ctx->smem_optin;
    const size_t smem = sizeof(Smem) + 1024 + (hcn_smem ? hcn_bytes : 0);
    auto kern = hcn_smem ? assign_tc5_kernel<NK, TILES, BN, true, TXS> : assign_tc5_kernel<NK, TILES, BN, false, TXS>;
    SCKM_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, TC_THREADS, smem, ctx->stream>>>(mapX, mapCh, mapCl, (const TXS*)ds->x, ds->n, (uint32_t)ds->d, ctx->d_centroids,
                                                  ctx->d_cnorm, hcn, ch, cl, (ds->have_labels && !getenv("SCKM_TC5_NOPRIME")) ? ds->labels : nullptr,
                                                  (uint32_t)k, nblocks, ds->labels, ds->mind,
                                                  ctx->d_partials, pk, ctx->d_flags, SCKM_LOOP_ARGS(ctx));
    LAUNCH_CHECK_T(ctx);
    return SCKM_OK;
}

int launch_assign_tc5(sckm_dataset* ds, uint64_t k) {
    sckm_ctx* ctx = ds->ctx;
    if (!tc5_supported(ds, k)) return fail(ctx, SCKM_ERR_INVALID, "shape not supported by the tcgen05 kernel");
    const size_t pk = (size_t)k * ds->d + k + 1;
    const unsigned grid = (unsigned)ctx->num_sms;
    SCKM_TRY(ensure_workspace(ctx, k, ds->d, (size_t)grid * TC_EPI_WARPS));
    ctx->partial_slots_used = grid * TC_EPI_WARPS;
    if (ds->n == 0) return SCKM_OK;
    SCKM_TRY(launch_cnorm(ctx, k, ds->d, false));                     // raw norms: this kernel ranks the raw f32 rows
    ctx->packed_centered = false;
    const float* x32 = (const float*)ds->x;
    if (ds->dtype == SCKM_F64) {                                      // f32 shadow copy of X for the ranking, built once
        if (!ds->x32) {
            SCKM_CUDA(ctx, dev_alloc(ctx, (void**)&ds->x32, ds->n * ds->d * sizeof(float)));
            tc5_shadow_kernel<<<ctx->num_sms * 8, 256, 0, ctx->stream>>>((const double*)ds->x, ds->x32, ds->n * ds->d);
            LAUNCH_CHECK_T(ctx);
        }
        x32 = ds->x32;
    }
    int rc;
    if (tc5h_supported(ds, k)) rc = launch_tc5h(ds, k, pk, x32);      // d <= 32: the 3xFP16 form (sckm_tc5h.cu); SCKM_TC5_TF32=1 keeps this file's
    else if (ds->d <= 32) rc = ds->dtype == SCKM_F64 ? launch_tc5_t<1, 2, 128, double>(ds, k, pk, x32) : launch_tc5_t<1, 2, 128, float>(ds, k, pk, x32);
    else             rc = ds->dtype == SCKM_F64 ? launch_tc5_t<2, 1, 64, double>(ds, k, pk, x32) : launch_tc5_t<2, 1, 64, float>(ds, k, pk, x32);
    SCKM_TRY(rc);
    if (getenv("SCKM_TRACE_MARKED")) {                              // diagnostic: rows this step hands to the exact re-decision
        unsigned long long m = 0;
        SCKM_CUDA(ctx, cudaMemcpyAsync(&m, ctx->d_flags, sizeof(m), cudaMemcpyDeviceToHost, ctx->stream));
        SCKM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        fprintf(stderr, "[sckm] tcgen05 step: %llu of %llu rows marked as near-ties (%.4f %%)\n", m, (unsigned long long)ds->n, 100.0 * (double)m / (double)ds->n);
    }
    return launch_refine_rows(ds, k, pk, grid);   // 8 warps per CTA: the same partial slots as the epilogue warps
}

}  // namespace sckm
