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
// mbarrier rings: x_full/x_ready/x_empty (2 stages), c_full/c_empty (2 stages), t_full/t_empty (2 TMEM stages).
#include "sckm_common.cuh"
#include <cuda.h>
#include <cfloat>
#include <algorithm>

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
int launch_cnorm(sckm_ctx* ctx, uint64_t k, uint64_t d);                               // sckm_dmma.cu

constexpr int TC_K = 32;                 // padded feature count = one 128-byte swizzle row of f32
constexpr int TC_BM = 128;               // rows per MMA tile (TMEM lanes)
constexpr int TC_TILES = 2;              // tiles per super-tile
constexpr int TC_BN = 128;               // centroids per block
constexpr int TC_EPI_WARPS = 8;
constexpr int TC_THREADS = 11 * 32;      // warps: 0 X-producer, 1 MMA, 2..9 epilogue, 10 C-producer
constexpr uint32_t TC_TILE_BYTES = TC_BM * TC_K * 4;   // 16 KB
constexpr double TC_TIE_REL = 2e-5;      // >= 10x the 3xTF32 + FP32-accumulate error bound (bench/tc5_probe.cu: 1e-6)
constexpr uint32_t TC_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC_BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);

// ---- PTX wrappers -----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile("{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n }"
                 ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
// K-major SWIZZLE_128B operand descriptor (cute::UMMA::SmemDescriptor): start>>4 | LBO=1 | SBO=1024 B | version 1 | layout 2
__device__ __forceinline__ uint64_t umma_desc_sw128(const void* smem) {
    return (uint64_t)((smem_u32(smem) >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_c, uint64_t da, uint64_t db, uint32_t accumulate) {
    asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n }"
                 ::"r"(tmem_c), "l"(da), "l"(db), "r"(TC_IDESC), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t addr, uint32_t (&v)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                   "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]),
                   "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]),
                   "=r"(v[30]), "=r"(v[31]) : "r"(addr));
}

struct alignas(16) TcSmem {           // dynamic shared memory image (base aligned to 1024 B)
    float xh[2][TC_TILES][TC_BM * TC_K];   // raw rows on arrival, TF32 hi part after the split
    float xl[2][TC_TILES][TC_BM * TC_K];   // TF32 lo part
    float ch[2][TC_BN * TC_K];             // centroid block, hi
    float cl[2][TC_BN * TC_K];             // centroid block, lo
    uint64_t x_full[2], x_ready[2], x_empty[2], c_full[2], c_empty[2], t_full[2], t_empty[2];
    uint32_t tmem_base;
};

// centroids (f64 [k][d]) -> TF32 hi / lo parts in f32 [kpad][32] (zero padded) and -||c||^2/2 in f32
__global__ void tc5_prep_kernel(const double* __restrict__ centroids, const double* __restrict__ cnorm, uint32_t k, uint32_t d,
                                uint32_t kpad, float* __restrict__ ch, float* __restrict__ cl, float* __restrict__ hcn) {
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e < kpad * TC_K) {
        const uint32_t r = e / TC_K, c = e - r * TC_K;
        float v = (r < k && c < d) ? (float)centroids[(size_t)r * d + c] : 0.f;
        const float h = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
        ch[e] = h; cl[e] = v - h;
    }
    if (e < kpad) hcn[e] = e < k ? (float)(-0.5 * cnorm[e]) : -INFINITY;
}

// HCN_SMEM: -||c||^2/2 of all centroids resident in shared memory (fits up to ~7000 centroids), else read through L1
template <bool HCN_SMEM>
__global__ void __launch_bounds__(TC_THREADS, 1)
assign_tc5_kernel(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapCh,
                  const __grid_constant__ CUtensorMap mapCl, uint64_t n, uint32_t d,
                  const double* __restrict__ centroids, const double* __restrict__ cnorm, const float* __restrict__ hcn,
                  uint32_t k, uint32_t nblocks, uint32_t* __restrict__ labels, double* __restrict__ mind,
                  double* __restrict__ partials, size_t pk) {
    extern __shared__ unsigned char smem_raw[];
    TcSmem& S = *reinterpret_cast<TcSmem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint64_t rows_per_super = TC_TILES * TC_BM;
    const uint64_t nsuper = (n + rows_per_super - 1) / rows_per_super;
    float* s_hcn = reinterpret_cast<float*>(&S + 1);                 // [nblocks * TC_BN] when HCN_SMEM
    if (HCN_SMEM) for (uint32_t i = threadIdx.x; i < nblocks * TC_BN; i += blockDim.x) s_hcn[i] = hcn[i];

    if (threadIdx.x == 0) {
        for (int s = 0; s < 2; s++) {
            mbar_init(&S.x_full[s], 1); mbar_init(&S.x_ready[s], TC_EPI_WARPS); mbar_init(&S.x_empty[s], 1 + TC_EPI_WARPS);
            mbar_init(&S.c_full[s], 1); mbar_init(&S.c_empty[s], 1);
            mbar_init(&S.t_full[s], 1); mbar_init(&S.t_empty[s], TC_EPI_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&S.tmem_base)), "r"(512));
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
                mbar_expect_tx(&S.x_full[xs], TC_TILES * TC_TILE_BYTES);
                for (int m = 0; m < TC_TILES; m++)
                    tma_load_2d(S.xh[xs][m], &mapX, 0, (int)(st * rows_per_super + (uint64_t)m * TC_BM), &S.x_full[xs]);
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
                    mbar_expect_tx(&S.c_full[cs], 2 * TC_TILE_BYTES);
                    tma_load_2d(S.ch[cs], &mapCh, 0, (int)(b * TC_BN), &S.c_full[cs]);
                    tma_load_2d(S.cl[cs], &mapCl, 0, (int)(b * TC_BN), &S.c_full[cs]);
                }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (lane == 0) {
            uint32_t it = 0, j = 0;
            for (uint64_t st = blockIdx.x; st < nsuper; st += gridDim.x, it++) {
                const int xs = it & 1; const uint32_t xph = (it >> 1) & 1;
                mbar_wait(&S.x_ready[xs], xph);                      // hi/lo split done by the epilogue warps
                asm volatile("tcgen05.fence::after_thread_sync;");
                for (uint32_t b = 0; b < nblocks; b++, j++) {
                    const int cs = j & 1; const uint32_t ph = (j >> 1) & 1;   // centroid stage == TMEM stage index
                    mbar_wait(&S.c_full[cs], ph);
                    mbar_wait(&S.t_empty[cs], ph ^ 1);
                    asm volatile("tcgen05.fence::after_thread_sync;");
                    const uint64_t dCh = umma_desc_sw128(S.ch[cs]), dCl = umma_desc_sw128(S.cl[cs]);
                    for (int m = 0; m < TC_TILES; m++) {
                        const uint64_t dXh = umma_desc_sw128(S.xh[xs][m]), dXl = umma_desc_sw128(S.xl[xs][m]);
                        const uint32_t tcol = tmem + (uint32_t)(cs * 256 + m * TC_BN);
                        for (int ks = 0; ks < TC_K / 8; ks++) umma_tf32(tcol, dXh + 2 * ks, dCh + 2 * ks, ks > 0);
                        for (int ks = 0; ks < TC_K / 8; ks++) umma_tf32(tcol, dXh + 2 * ks, dCl + 2 * ks, 1);
                        for (int ks = 0; ks < TC_K / 8; ks++) umma_tf32(tcol, dXl + 2 * ks, dCh + 2 * ks, 1);
                    }
                    umma_commit(&S.c_empty[cs]);                     // centroid stage free once these MMAs have read it
                    umma_commit(&S.t_full[cs]);                      // accumulators ready for the epilogue
                }
                umma_commit(&S.x_empty[xs]);                         // X stage: MMA side done (epilogue warps arrive too)
            }
        }
    } else {
        // ================= epilogue warps: thread = row =================
        const int ew = warp - 2;                                     // 0..7
        const int m = ew >> 2;                                       // tile of the super-tile
        const int q = warp & 3;                                      // TMEM lane quadrant this warp may access
        const int rloc = q * 32 + lane;                              // row within the tile
        const double cmax = cnorm[k];
        double* part = partials + ((size_t)blockIdx.x * TC_EPI_WARPS + ew) * ((pk + 15) / 16 * 16);
        unsigned lanemask_lt;
        asm("mov.u32 %0, %%lanemask_lt;" : "=r"(lanemask_lt));
        // split my row of X stage `xs` in place (128 bytes at rloc*128; the swizzle only permutes 16-byte chunks
        // inside it) and return ||x||^2; then tell the MMA warp that the stage is ready
        auto split_stage = [&](int xs, uint32_t xph) -> double {
            mbar_wait(&S.x_full[xs], xph);
            float4* ph4 = reinterpret_cast<float4*>(S.xh[xs][m] + rloc * TC_K);
            float4* pl4 = reinterpret_cast<float4*>(S.xl[xs][m] + rloc * TC_K);
            double xn = 0.0;
#pragma unroll
            for (int c0 = 0; c0 < 8; c0++) {
                const int c = (c0 + lane) & 7;                         // rotate: 8 lanes hit 8 different 16-byte columns
                float4 v = ph4[c], h, l;
                h.x = __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u); l.x = v.x - h.x;
                h.y = __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u); l.y = v.y - h.y;
                h.z = __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u); l.z = v.z - h.z;
                h.w = __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u); l.w = v.w - h.w;
                ph4[c] = h; pl4[c] = l;
                xn = fma((double)v.x, (double)v.x, xn); xn = fma((double)v.y, (double)v.y, xn);
                xn = fma((double)v.z, (double)v.z, xn); xn = fma((double)v.w, (double)v.w, xn);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic writes -> visible to the MMA (async proxy)
            __syncwarp();
            if (lane == 0) mbar_arrive(&S.x_ready[xs]);
            return xn;
        };
        uint32_t it = 0, j = 0;
        double xn_next = blockIdx.x < nsuper ? split_stage(0, 0) : 0.0;
        for (uint64_t st = blockIdx.x; st < nsuper; st += gridDim.x, it++) {
            const int xs = it & 1;
            const uint64_t row = st * rows_per_super + (uint64_t)m * TC_BM + rloc;
            const bool valid = row < n;
            const double xn = xn_next;
            float4* ph4 = reinterpret_cast<float4*>(S.xh[xs][m] + rloc * TC_K);
            float4* pl4 = reinterpret_cast<float4*>(S.xl[xs][m] + rloc * TC_K);
            // ---- running top-2 over all centroid blocks ----
            float best = -FLT_MAX, second = -FLT_MAX; uint32_t bi = 0;
            for (uint32_t b = 0; b < nblocks; b++, j++) {
                const int ts = j & 1; const uint32_t ph = (j >> 1) & 1;
                // split the NEXT super-tile as soon as this one is under way, so the MMA warp never waits for it
                if (b == (nblocks > 1 ? 1u : 0u) && st + gridDim.x < nsuper) xn_next = split_stage((it + 1) & 1, ((it + 1) >> 1) & 1);
                mbar_wait(&S.t_full[ts], ph);
                asm volatile("tcgen05.fence::after_thread_sync;");
                const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(ts * 256 + m * TC_BN);
                const float4* h4 = reinterpret_cast<const float4*>((HCN_SMEM ? s_hcn : hcn) + (size_t)b * TC_BN);
                // software pipeline over the four 32-column chunks: chunk c+1 is in flight (tcgen05.ld is asynchronous
                // until tcgen05.wait::ld) while chunk c goes through the top-2 update
                uint32_t va[32], vb[32];
                auto consume = [&](const uint32_t (&v)[32], int c0) {
#pragma unroll
                    for (int u = 0; u < 8; u++) {
                        const float4 hv = HCN_SMEM ? h4[(c0 >> 2) + u] : __ldg(h4 + (c0 >> 2) + u);
                        const float hh[4] = {hv.x, hv.y, hv.z, hv.w};
#pragma unroll
                        for (int e = 0; e < 4; e++) {
                            const float sc = __uint_as_float(v[u * 4 + e]) + hh[e];
                            const bool gt = sc > best;
                            second = fmaxf(second, gt ? best : sc);
                            bi = gt ? (b * TC_BN + c0 + u * 4 + e) : bi;
                            best = fmaxf(best, sc);
                        }
                    }
                };
                tmem_ld32(taddr, va);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                tmem_ld32(taddr + 32, vb);
                consume(va, 0);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                tmem_ld32(taddr + 64, va);
                consume(vb, 32);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                tmem_ld32(taddr + 96, vb);
                consume(va, 64);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                asm volatile("tcgen05.fence::before_thread_sync;");   // all of this stage's columns are in registers
                __syncwarp();
                if (lane == 0) mbar_arrive(&S.t_empty[ts]);
                consume(vb, 96);
            }
            // ---- decide: exact f64 distance to the winner, near-tie mark ----
            // my row again in logical order: chunk c of row r sits at physical chunk c ^ (r & 7)
            float xr[TC_K];
#pragma unroll
            for (int c = 0; c < 8; c++) {
                const float4 h = ph4[c ^ (rloc & 7)], l = pl4[c ^ (rloc & 7)];
                xr[4 * c + 0] = h.x + l.x; xr[4 * c + 1] = h.y + l.y; xr[4 * c + 2] = h.z + l.z; xr[4 * c + 3] = h.w + l.w;
            }
            const double gap = 2.0 * ((double)best - (double)second);
            const bool tie = !(gap > TC_TIE_REL * (xn + cmax)) || bi >= k;
            double dist = 0.0;
            if (valid && !tie) {
                const double* cr = centroids + (size_t)bi * d;
#pragma unroll
                for (int f = 0; f < TC_K; f++)
                    if ((uint32_t)f < d) { const double r = (double)xr[f] - cr[f]; dist = fma(r, r, dist); }
            }
            if (valid) { labels[row] = tie ? 0xffffffffu : bi; mind[row] = dist; }
            // ---- deterministic fused update: the warp walks its 32 rows in order; lane f adds feature f of the row
            // (read back from the staged tile) to the warp's private partial with a fire-and-forget RED.  Every
            // address only ever receives adds from one thread, in program order => fixed summation order. ----
            const bool part_ok = valid && !tie;
            const uint32_t lab = part_ok ? bi : 0xffffffffu;
            {
                const float* th = S.xh[xs][m] + (size_t)(q * 32) * TC_K;   // first row of this warp's quadrant
                const float* tl = S.xl[xs][m] + (size_t)(q * 32) * TC_K;
#pragma unroll 4
                for (int r = 0; r < 32; r++) {
                    const uint32_t lr = __shfl_sync(0xffffffffu, lab, r);
                    if (lr == 0xffffffffu) continue;                       // warp-uniform
                    const int phys = r * TC_K + ((((lane >> 2) ^ (r & 7)) << 2) | (lane & 3));   // undo the 128-byte swizzle
                    if ((uint32_t)lane < d) atomicAdd(part + (size_t)lr * d + lane, (double)th[phys] + (double)tl[phys]);
                }
                // counts: one add per distinct label of the warp (the lowest lane of each group adds the group size)
                const unsigned peers = __match_any_sync(0xffffffffu, part_ok ? bi : (0x80000000u | (uint32_t)lane));
                if (part_ok && (peers & lanemask_lt) == 0) atomicAdd(part + (size_t)k * d + bi, (double)__popc(peers));
            }
            double v = part_ok ? dist : 0.0;                          // fixed-order sum over the warp's 32 rows
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v = __dadd_rn(v, __shfl_xor_sync(0xffffffffu, v, o));
            if (lane == 0) { atomicAdd(part + pk - 1, v); mbar_arrive(&S.x_empty[xs]); }
            __syncwarp();
        }
    }
    // ---- teardown ----
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

// ---- host side --------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr; cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

static int make_map(sckm_ctx* ctx, CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols) {
    EncodeTiledFn enc = encode_fn();
    if (!enc) return fail(ctx, SCKM_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    cuuint64_t dims[2] = {cols, rows}, strides[1] = {cols * 4};
    cuuint32_t box[2] = {TC_K, TC_BM}, estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(ctx, SCKM_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) for %llu x %llu", (int)r,
                                       (unsigned long long)rows, (unsigned long long)cols);
    return SCKM_OK;
}

bool tc5_supported(const sckm_dataset* ds, uint64_t k) {
    return ds->dtype == SCKM_F32 && ds->d >= 4 && ds->d <= TC_K && ds->d % 4 == 0 && k >= 16 && k <= (1u << 20) &&
           ds->n < 0x7FFFFFFFull && encode_fn() != nullptr;
}

int launch_assign_tc5(sckm_dataset* ds, uint64_t k) {
    sckm_ctx* ctx = ds->ctx;
    if (!tc5_supported(ds, k)) return fail(ctx, SCKM_ERR_INVALID, "shape not supported by the tcgen05 kernel");
    const size_t pk = (size_t)k * ds->d + k + 1;
    const unsigned grid = (unsigned)ctx->num_sms;
    SCKM_TRY(ensure_workspace(ctx, k, ds->d, (size_t)grid * TC_EPI_WARPS));
    ctx->partial_slots_used = grid * TC_EPI_WARPS;
    if (ds->n == 0) return SCKM_OK;
    SCKM_TRY(launch_cnorm(ctx, k, ds->d));
    const uint32_t nblocks = (uint32_t)((k + TC_BN - 1) / TC_BN), kpad = nblocks * TC_BN;
    const size_t need = (size_t)kpad * TC_K * 2 + kpad;               // ch | cl | hcn   (floats)
    if (need > ctx->cap_tc5) {
        if (ctx->d_tc5) { SCKM_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); cudaFree(ctx->d_tc5); ctx->d_tc5 = nullptr; }
        SCKM_CUDA(ctx, cudaMalloc((void**)&ctx->d_tc5, need * sizeof(float)));
        ctx->cap_tc5 = need;
    }
    float* ch = ctx->d_tc5; float* cl = ch + (size_t)kpad * TC_K; float* hcn = cl + (size_t)kpad * TC_K;
    tc5_prep_kernel<<<(kpad * TC_K + 255) / 256, 256, 0, ctx->stream>>>(ctx->d_centroids, ctx->d_cnorm, (uint32_t)k, (uint32_t)ds->d,
                                                                     kpad, ch, cl, hcn);
    LAUNCH_CHECK_T(ctx);
    CUtensorMap mapX, mapCh, mapCl;
    SCKM_TRY(make_map(ctx, &mapX, ds->x, ds->n, ds->d));
    SCKM_TRY(make_map(ctx, &mapCh, ch, kpad, TC_K));
    SCKM_TRY(make_map(ctx, &mapCl, cl, kpad, TC_K));
    const size_t hcn_bytes = (size_t)kpad * sizeof(float);
    const bool hcn_smem = sizeof(TcSmem) + 1024 + hcn_bytes <= (size_t)ctx->smem_optin;
    const size_t smem = sizeof(TcSmem) + 1024 + (hcn_smem ? hcn_bytes : 0);
    auto kern = hcn_smem ? assign_tc5_kernel<true> : assign_tc5_kernel<false>;
    SCKM_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, TC_THREADS, smem, ctx->stream>>>(mapX, mapCh, mapCl, ds->n, (uint32_t)ds->d, ctx->d_centroids, ctx->d_cnorm,
                                                  hcn, (uint32_t)k, nblocks, ds->labels, ds->mind, ctx->d_partials, pk);
    LAUNCH_CHECK_T(ctx);
    return launch_refine_rows(ds, k, pk, grid);   // 8 warps per CTA: the same partial slots as the epilogue warps
}

}  // namespace sckm
