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
//   * Both operands are CENTRED on a per-fit shift mu (the mean of the fit's initial centroids, launch_cnorm):
//     rows become x - mu as they are loaded, the staged centroid block holds c - mu, the norms are ||c - mu||^2.
//     ||x - c||^2 does not change, but the cancellation error of the GEMM form now scales with the spread of the
//     data around mu instead of its distance from the origin, and so do the near-tie threshold and the per-point
//     distance.  The fused update accumulates the centred values; finalize_kernel adds mu back
//     (centroid = sum(x - mu) / count + mu).  The subtraction is not free on the FP64 datapath the DMMAs use (measured
//     on B200: +2.8 % on the resident kernel at C3, +12 % on the streamed-centroid kernel at C4), so it is a template
//     switch (CENTER) that launch_cnorm turns on only when the data sit further from the origin than they are wide
//     (||mu||^2 > max_j ||c_j - mu||^2): centring then buys more than a factor 4 in precision; otherwise mu = 0.
//   * Rows whose best/second gap is within 1e-10*(||x||^2 + max||c||^2) (>= 1e4 x the rounding
//     error bound of the GEMM form) are marked and re-decided by refine_rows_kernel with the
//     reference's exact arithmetic (euclidian.rs:56-63), so labels equal the exact direct-form argmin.
//   * The centroid update (per-label sums, counts, inertia) is fused: while a slab's rows are still
//     in registers they are added to the warp's PRIVATE partial [k*d | k | 1] in L2-resident global
//     memory, in a fixed order, so results are bit-reproducible run to run (needed by the stop rule
//     `distortion <= dist`, kmeans.rs:305) without any floating-point atomics.
#include "sckm_common.cuh"
#include "sckm_tile.cuh"
#include <cfloat>
#include <algorithm>
#include <cstdlib>

namespace sckm {

#define LAUNCH_CHECK_D(ctx)                                                                        \
    do {                                                                                           \
        (ctx)->launches++;                                                                         \
        cudaError_t _e = cudaGetLastError();                                                       \
        if (_e != cudaSuccess)                                                                     \
            return fail((ctx), SCKM_ERR_CUDA, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), \
                        __FILE__, __LINE__);                                                       \
    } while (0)

constexpr int DMMA_MAX_WARPS = 16;
constexpr double DMMA_TIE_REL = 1e-10;


// ---- building blocks of the tile kernel (all force-inlined; arrays stay in registers) ----
// accumulators start at -||c||^2/2 (plain copies out of shared memory, no FP64 instruction); the row's own
// -||x||^2/2 is constant over the centroids and is only added back once per row at the very end
template <int MT, int NT>
__device__ __forceinline__ void acc_init(double (&acc)[MT][NT][2], const double* __restrict__ cn_sub, int t) {
#pragma unroll
    for (int nt = 0; nt < NT; nt++) {
        const double2 c2 = *reinterpret_cast<const double2*>(cn_sub + nt * 8 + 2 * t);
#pragma unroll
        for (int mt = 0; mt < MT; mt++) { acc[mt][nt][0] = c2.x; acc[mt][nt][1] = c2.y; }
    }
}

// one epilogue item: element (nt, e) of m-tile mt -> running max / second max / argmax (ascending column order)
template <int MT, int NT>
__device__ __forceinline__ void epi_item(const double (&acc)[MT][NT][2], int item, uint32_t col_base, int t,
                                         key_t (&best)[MT], key_t (&second)[MT], uint32_t (&bidx)[MT]) {
    const int mt = item % MT, j = item / MT, nt = j >> 1, e = j & 1;
    const uint32_t col = col_base + nt * 8 + 2 * t + e;
    const key_t v = dkey(acc[mt][nt][e]);
    const bool gt = v > best[mt];
    const key_t lose = gt ? best[mt] : v;
    second[mt] = lose > second[mt] ? lose : second[mt];
    bidx[mt] = gt ? col : bidx[mt];
    best[mt] = gt ? v : best[mt];
}

// K loop of one sub-block: acc += A(rows x d) * B(d x 8*NT centroids), B fragments read from shared memory
template <int KSTEPS, int MT, int NT, int PITCH>
__device__ __forceinline__ void kloop(double (&acc)[MT][NT][2], const double (&a)[MT][KSTEPS],
                                      const double* __restrict__ bp) {
#pragma unroll
    for (int ks = 0; ks < KSTEPS; ks++) {
        double b[NT];
#pragma unroll
        for (int nt = 0; nt < NT; nt++) b[nt] = bp[(size_t)nt * 8 * PITCH + ks * 4];
#pragma unroll
        for (int mt = 0; mt < MT; mt++)
#pragma unroll
            for (int nt = 0; nt < NT; nt++) dmma884(acc[mt][nt][0], acc[mt][nt][1], a[mt][ks], b[nt]);
    }
}

template <int MT, int NT>
__device__ __forceinline__ void epilogue_all(const double (&acc)[MT][NT][2], uint32_t col_base, int t,
                                             key_t (&best)[MT], key_t (&second)[MT], uint32_t (&bidx)[MT]) {
#pragma unroll
    for (int item = 0; item < MT * NT * 2; item++) epi_item<MT, NT>(acc, item, col_base, t, best, second, bidx);
}

// The 4 lanes that share a row each hold the top-2 of their own columns: butterfly-merge them (lowest index wins ties)
template <int MT>
__device__ __forceinline__ void merge_row_lanes(key_t (&best)[MT], key_t (&second)[MT], uint32_t (&bidx)[MT]) {
#pragma unroll
    for (int mt = 0; mt < MT; mt++) {
#pragma unroll
        for (int o = 1; o <= 2; o <<= 1) {
            const key_t ob = __shfl_xor_sync(0xffffffffu, best[mt], o);
            const key_t os = __shfl_xor_sync(0xffffffffu, second[mt], o);
            const uint32_t oi = __shfl_xor_sync(0xffffffffu, bidx[mt], o);
            const bool take = ob > best[mt] || (ob == best[mt] && oi < bidx[mt]);
            const key_t lo_best = ob < best[mt] ? ob : best[mt];
            key_t s2 = os > second[mt] ? os : second[mt];
            second[mt] = lo_best > s2 ? lo_best : s2;
            bidx[mt] = take ? oi : bidx[mt];
            best[mt] = ob > best[mt] ? ob : best[mt];
        }
    }
}

// End of a slab (all centroids seen, lanes merged): labels out, near-ties marked, and the fused update while the rows
// are still in registers.  n_valid = number of rows that exist (0 for an inactive slab).
template <int KSTEPS, int MT, bool UPDATE>
__device__ __forceinline__ void finish_slab(const double (&a)[MT][KSTEPS], const double (&xn)[MT], const key_t (&best)[MT],
                                            const key_t (&second)[MT], const uint32_t (&bidx)[MT], uint64_t r0, uint64_t n_valid,
                                            uint32_t d, uint32_t k, double cmax, int g, int t, int lane, unsigned lanemask_lt,
                                            uint32_t* __restrict__ labels, double* __restrict__ mind, double* __restrict__ part,
                                            size_t pk, unsigned long long* __restrict__ nmarked) {
    double slab_inertia = 0.0;
#pragma unroll
    for (int mt = 0; mt < MT; mt++) {
        const uint64_t row = r0 + mt * 8 + g;
        const bool valid = row < n_valid;
        const double bestv = dunkey(best[mt]), secondv = dunkey(second[mt]);   // x.c - ||c||^2/2
        const double dist = fmax(0.0, fma(-2.0, bestv, xn[mt]));
        const double gap = 2.0 * (bestv - secondv);
        const bool tie = !(gap > DMMA_TIE_REL * (xn[mt] + cmax));   // also catches NaN
        if (valid && t == 0) {
            labels[row] = tie ? 0xffffffffu : bidx[mt];             // ties are re-decided by refine_rows_kernel
            if (UPDATE) mind[row] = dist;
            if (tie) atomicAdd(nmarked, 1ull);
        }
        // Deterministic per-label accumulation (the update of bbd_tree.rs:151-155) into this warp's
        // private partial: the 4 lanes of a row add their A-fragment elements.  Rows of this m-tile that
        // share a label are serialised in ascending row order (rank), so the order of every f64 addition is
        // fixed by (n, grid) alone.
        const bool part_ok = UPDATE && valid && !tie;        // UPDATE = false: labels only (predict)
        const uint32_t key = part_ok ? bidx[mt] : (0x80000000u | (uint32_t)g);
        const unsigned peers = __match_any_sync(0xffffffffu, key);
        const int rank = __popc(peers & lanemask_lt) >> 2;
        const int maxrank = __reduce_max_sync(0xffffffffu, part_ok ? rank : 0);
        for (int r = 0; r <= maxrank; r++) {
            if (r) { __threadfence(); __syncwarp(); }     // order the (rare) same-label rows of this m-tile
            if (part_ok && rank == r) {
                // fire-and-forget RED.ADD.F64 into the warp-private partial: a given address only ever
                // receives adds from this warp, same-thread adds stay in program order and cross-lane
                // same-label adds are separated by the fence above => the summation order is fixed.
                double* p = part + (size_t)bidx[mt] * d + t;
#pragma unroll
                for (int ks = 0; ks < KSTEPS; ks++)
                    if (ks * 4 + t < d) atomicAdd(p + ks * 4, a[mt][ks]);
                if (t == 0) atomicAdd(part + (size_t)k * d + bidx[mt], 1.0);
            }
        }
        double v = (part_ok && t == 0) ? dist : 0.0;                 // fixed-order sum over the 8 rows
        v = __dadd_rn(v, __shfl_xor_sync(0xffffffffu, v, 4));
        v = __dadd_rn(v, __shfl_xor_sync(0xffffffffu, v, 8));
        v = __dadd_rn(v, __shfl_xor_sync(0xffffffffu, v, 16));
        slab_inertia = __dadd_rn(slab_inertia, v);
    }
    if (UPDATE && lane == 0 && n_valid) atomicAdd(part + pk - 1, slab_inertia);
}

// Centroid block [c0, c0 + bn) -> shared memory (row pitch PITCH doubles, zero rows past k), -||c - mu||^2 / 2 -> cn.
// Whole rows (d == DP: no column padding) go as 16-byte cp.async copies, all of a thread's copies in flight at once and
// no register round trip: the streamed-centroid kernel switches blocks ~500 times per launch at config C4, and the
// first version of this loop (one dependent 8-byte load per iteration, ~50 L2 round trips in a row) cost ~35 us per
// switch -- 16 % of the step (measured: the step time grew linearly with the number of switches).
__device__ __forceinline__ void dmma_cp_async16(void* smem_dst, const void* gmem_src) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src));
}
template <int DP, int PITCH, bool CENTER>
__device__ __forceinline__ void stage_centroid_block(double* __restrict__ cbuf, double* __restrict__ cn,
                                                     const double* __restrict__ centroids, const double* __restrict__ cnorm,
                                                     const double* __restrict__ mu_s, uint32_t c0, uint32_t bn, uint32_t k, uint32_t d) {
    if (d == (uint32_t)DP) {
        constexpr uint32_t CPR = DP / 2;                           // 16-byte chunks per row
        const uint32_t total = bn * CPR;
        for (uint32_t e = threadIdx.x; e < total; e += blockDim.x) {
            const uint32_t r = e / CPR, q = e - r * CPR;
            double* dst = cbuf + (size_t)r * PITCH + 2 * q;
            if (c0 + r < k) dmma_cp_async16(dst, centroids + (size_t)(c0 + r) * DP + 2 * q);
            else *reinterpret_cast<double2*>(dst) = make_double2(0.0, 0.0);
        }
        asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
        if (CENTER) {                                              // (a thread sees its own copies after the wait)
            for (uint32_t e = threadIdx.x; e < total; e += blockDim.x) {
                const uint32_t r = e / CPR, q = e - r * CPR;
                if (c0 + r >= k) continue;
                double2* p = reinterpret_cast<double2*>(cbuf + (size_t)r * PITCH + 2 * q);
                double2 v = *p;
                v.x -= mu_s[2 * q]; v.y -= mu_s[2 * q + 1];
                *p = v;
            }
        }
    } else {                                                       // ragged d: eight independent loads in flight per thread
        const uint32_t total = bn * DP;
        for (uint32_t e0 = threadIdx.x; e0 < total; e0 += 8 * blockDim.x) {
            double v[8];
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const uint32_t e = e0 + i * blockDim.x, r = e / DP, c = e - r * DP;
                v[i] = (e < total && c0 + r < k && c < d) ? centroids[(size_t)(c0 + r) * d + c] : 0.0;
            }
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const uint32_t e = e0 + i * blockDim.x, r = e / DP, c = e - r * DP;
                if (e < total) cbuf[(size_t)r * PITCH + c] = (CENTER && c0 + r < k && c < d) ? v[i] - mu_s[c] : v[i];
            }
        }
    }
    for (uint32_t r = threadIdx.x; r < bn; r += blockDim.x)
        cn[r] = (c0 + r < k) ? -0.5 * cnorm[c0 + r] : -INFINITY;
}

// Resident-centroid specialisation (the whole centroid set fits in shared memory: config C3).  Kept as its own
// kernel: it is the headline path and its instruction schedule is tuned (87.8 % of the FP64 peak).
// KSTEPS: d padded to 4*KSTEPS; MT: m-tiles (8 rows each) per warp slab; DMMA_NT: n-tiles (8 centroids each) per
// accumulator sub-block; DMMA_WARPS: warps per CTA (one CTA per SM)
template <int KSTEPS, int MT, int DMMA_NT, int DMMA_WARPS, bool UPDATE, bool CENTER, typename TX>
__global__ void __launch_bounds__(DMMA_WARPS * 32, 1)
assign_dmma_resident_kernel(const TX* __restrict__ x, uint64_t n, uint32_t d, const double* __restrict__ centroids,
                   const double* __restrict__ cnorm, const double* __restrict__ mu, uint32_t k, uint32_t bn, uint32_t* __restrict__ labels,
                   double* __restrict__ mind, double* __restrict__ partials, size_t pk, unsigned long long* __restrict__ nmarked,
                   const LoopState* __restrict__ loop_st, uint32_t loop_it) {
    if (loop_done(loop_st, loop_it)) return;       // the fit's stop rule already fired (kmeans.rs:305)
    constexpr int DP = KSTEPS * 4;                 // padded feature count
    constexpr int PITCH = DP + 4;                  // doubles per staged centroid row (pitch = d*8+32 B)
    constexpr int ROWS = 8 * MT;
    extern __shared__ __align__(16) double smem_d[];
    double* cbuf = smem_d;                         // [bn][PITCH]  c - mu
    double* cn = smem_d + (size_t)bn * PITCH;      // [bn]  -||c - mu||^2/2, -inf for padding columns
    double* mu_s = cn + bn;                        // [DP]  centring shift (0 in the padding columns)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    for (uint32_t c = threadIdx.x; c < (uint32_t)DP; c += blockDim.x) mu_s[c] = c < d ? mu[c] : 0.0;
    const double cmax = cta_max(cnorm, k);         // max_j ||c_j - mu||^2 (its barriers also publish mu_s)

    const uint64_t nslabs = (n + ROWS - 1) / ROWS;
    const uint64_t stride = (uint64_t)gridDim.x * DMMA_WARPS;
    const uint64_t rounds = (nslabs + stride - 1) / stride;
    const uint32_t nchunks = (k + bn - 1) / bn;
    const double* bbase = cbuf + (size_t)g * PITCH + t;
    double* part = partials + ((size_t)blockIdx.x * DMMA_WARPS + warp) * ((pk + 15) / 16 * 16);   // this warp's private partial
    unsigned lanemask_lt;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(lanemask_lt));

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
                if (rok && col < d) v = CENTER ? (double)__ldg(xr + col) - mu_s[col] : (double)__ldg(xr + col);
                a[mt][ks] = v;
                s = fma(v, v, s);
            }
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            s += __shfl_xor_sync(0xffffffffu, s, 2);
            xn[mt] = s;
        }
        key_t best[MT], second[MT];
        uint32_t bidx[MT];
#pragma unroll
        for (int mt = 0; mt < MT; mt++) { best[mt] = KEY_MIN; second[mt] = KEY_MIN; bidx[mt] = 0; }

        for (uint32_t ch = 0; ch < nchunks; ch++) {
            const uint32_t c0 = ch * bn;
            if (nchunks > 1 || rd == 0) {
                // (re)load the centroid block; resident across rounds when it is the only one
                if (nchunks > 1) __syncthreads();
                stage_centroid_block<DP, PITCH, CENTER>(cbuf, cn, centroids, cnorm, mu_s, c0, bn, k, d);
                __syncthreads();
            }
            const uint32_t cols = min(bn, k - c0);
            constexpr int SUB = 8 * DMMA_NT;                       // centroids per accumulator sub-block
            const uint32_t nsub = (cols + SUB - 1) / SUB;
            for (uint32_t sb = 0; sb < nsub; sb++) {
                double acc[MT][DMMA_NT][2];
                acc_init<MT, DMMA_NT>(acc, cn + sb * SUB, t);
                kloop<KSTEPS, MT, DMMA_NT, PITCH>(acc, a, bbase + (size_t)sb * SUB * PITCH);
                epilogue_all<MT, DMMA_NT>(acc, c0 + sb * SUB, t, best, second, bidx);
            }
        }
        // ---- merge the 4 lanes that share a row, write out, mark near-ties, fused update ----
        merge_row_lanes<MT>(best, second, bidx);
        finish_slab<KSTEPS, MT, UPDATE>(a, xn, best, second, bidx, r0, active ? n : 0, d, k, cmax, g, t, lane, lanemask_lt,
                                        labels, mind, part, pk, nmarked);
    }
}

// KSTEPS: d padded to 4*KSTEPS; MT: m-tiles (8 rows each) per warp slab; DMMA_NT: n-tiles (8 centroids each) per
// accumulator sub-block; DMMA_WARPS: warps per CTA (one CTA per SM).
// When the centroids do not fit in shared memory they are streamed through it in blocks of `bn`; to amortise each
// block load (and its two CTA barriers) a warp then runs `sl` slabs against the resident block, keeping the running
// (best, second, argbest) of every row in a small per-warp shared-memory table between blocks.  Rows are re-read per
// block, but a round's working set (148 CTAs x warps x sl slabs) is L2-resident, so HBM still sees X once.
// (A fully resident centroid set takes assign_dmma_resident_kernel above.)
template <int KSTEPS, int MT, int DMMA_NT, int DMMA_WARPS, bool UPDATE, bool CENTER, typename TX>
__global__ void __launch_bounds__(DMMA_WARPS * 32, 1)
assign_dmma_kernel(const TX* __restrict__ x, uint64_t n, uint32_t d, const double* __restrict__ centroids,
                   const double* __restrict__ cnorm, const double* __restrict__ mu, uint32_t k, uint32_t bn, uint32_t sl_arg, uint32_t* __restrict__ labels,
                   double* __restrict__ mind, double* __restrict__ partials, size_t pk, unsigned long long* __restrict__ nmarked,
                   const LoopState* __restrict__ loop_st, uint32_t loop_it) {
    if (loop_done(loop_st, loop_it)) return;       // the fit's stop rule already fired (kmeans.rs:305)
    const uint32_t sl = sl_arg;
    constexpr int DP = KSTEPS * 4;                 // padded feature count
    constexpr int PITCH = DP + 4;                  // doubles per staged centroid row (pitch = d*8+32 B)
    constexpr int ROWS = 8 * MT;
    constexpr int SUB = 8 * DMMA_NT;               // centroids per accumulator sub-block
    extern __shared__ __align__(16) double smem_d[];
    double* cbuf = smem_d;                         // [bn][PITCH]  c - mu
    double* cn = smem_d + (size_t)bn * PITCH;      // [bn]  -||c - mu||^2/2, -inf for padding columns
    double* mu_s = cn + bn;                        // [DP]  centring shift (0 in the padding columns)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    // per-warp row state between centroid blocks: [sl][ROWS] x {best, second, idx}
    key_t* st_best = reinterpret_cast<key_t*>(mu_s + DP) + (size_t)warp * sl * ROWS * 3;
    key_t* st_second = st_best + (size_t)sl * ROWS;
    key_t* st_idx = st_second + (size_t)sl * ROWS;
    for (uint32_t c = threadIdx.x; c < (uint32_t)DP; c += blockDim.x) mu_s[c] = c < d ? mu[c] : 0.0;
    const double cmax = cta_max(cnorm, k);         // max_j ||c_j - mu||^2 (its barriers also publish mu_s)

    const uint64_t nslabs = (n + ROWS - 1) / ROWS;
    const uint64_t stride = (uint64_t)gridDim.x * DMMA_WARPS;
    const uint64_t rounds = (nslabs + stride * sl - 1) / (stride * sl);
    const uint32_t nchunks = (k + bn - 1) / bn;
    const double* bbase = cbuf + (size_t)g * PITCH + t;
    double* part = partials + ((size_t)blockIdx.x * DMMA_WARPS + warp) * ((pk + 15) / 16 * 16);   // this warp's private partial
    unsigned lanemask_lt;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(lanemask_lt));

    for (uint64_t rd = 0; rd < rounds; rd++) {
        for (uint32_t ch = 0; ch < nchunks; ch++) {
            const uint32_t c0 = ch * bn;
            if (nchunks > 1 || rd == 0) {
                // (re)load the centroid block; resident for the whole launch when it is the only one
                if (nchunks > 1) __syncthreads();
                stage_centroid_block<DP, PITCH, CENTER>(cbuf, cn, centroids, cnorm, mu_s, c0, bn, k, d);
                __syncthreads();
            }
            const uint32_t cols = min(bn, k - c0);
            const uint32_t nsub = (cols + SUB - 1) / SUB;
            const bool last = ch + 1 == nchunks;

            for (uint32_t si = 0; si < sl; si++) {
                const uint64_t slab = (rd * sl + si) * stride + (uint64_t)blockIdx.x * DMMA_WARPS + warp;
                const bool active = slab < nslabs;
                if (!active) continue;                                    // warp-uniform
                const uint64_t r0 = slab * ROWS;
                // ---- rows -> A fragments (registers), ||x||^2 ----
                double a[MT][KSTEPS];
                double xn[MT];
#pragma unroll
                for (int mt = 0; mt < MT; mt++) {
                    const uint64_t row = r0 + mt * 8 + g;
                    const bool rok = row < n;
                    const TX* xr = x + row * d;
                    double s = 0.0;
#pragma unroll
                    for (int ks = 0; ks < KSTEPS; ks++) {
                        const uint32_t col = ks * 4 + t;
                        double v = 0.0;
                        if (rok && col < d) v = CENTER ? (double)__ldg(xr + col) - mu_s[col] : (double)__ldg(xr + col);
                        a[mt][ks] = v;
                        s = fma(v, v, s);
                    }
                    s += __shfl_xor_sync(0xffffffffu, s, 1);
                    s += __shfl_xor_sync(0xffffffffu, s, 2);
                    xn[mt] = s;
                }
                // ---- running top-2 of every row: fresh on the first block, else resumed (lane t == 0 carries it) ----
                key_t best[MT], second[MT];
                uint32_t bidx[MT];
#pragma unroll
                for (int mt = 0; mt < MT; mt++) {
                    best[mt] = KEY_MIN; second[mt] = KEY_MIN; bidx[mt] = 0;
                    if (ch > 0 && t == 0) {
                        const uint32_t o = si * ROWS + mt * 8 + g;
                        best[mt] = st_best[o]; second[mt] = st_second[o]; bidx[mt] = (uint32_t)st_idx[o];
                    }
                }
                for (uint32_t sb = 0; sb < nsub; sb++) {
                    double acc[MT][DMMA_NT][2];
                    acc_init<MT, DMMA_NT>(acc, cn + sb * SUB, t);
                    kloop<KSTEPS, MT, DMMA_NT, PITCH>(acc, a, bbase + (size_t)sb * SUB * PITCH);
                    epilogue_all<MT, DMMA_NT>(acc, c0 + sb * SUB, t, best, second, bidx);
                }
                merge_row_lanes<MT>(best, second, bidx);
                if (!last) {                                              // park the row state until the next block
                    if (t == 0) {
#pragma unroll
                        for (int mt = 0; mt < MT; mt++) {
                            const uint32_t o = si * ROWS + mt * 8 + g;
                            st_best[o] = best[mt]; st_second[o] = second[mt]; st_idx[o] = (key_t)bidx[mt];
                        }
                    }
                    continue;
                }
                // ---- last block: write out, mark near-ties, fused update ----
                finish_slab<KSTEPS, MT, UPDATE>(a, xn, best, second, bidx, r0, n, d, k, cmax, g, t, lane, lanemask_lt,
                                                labels, mind, part, pk, nmarked);
            }
        }
    }
}

// Exact re-decision of the rows marked 0xffffffff by the tile kernel.  Warp w scans the fixed row range
// [w*R, (w+1)*R) in order; for a marked row, lane l scans centroids l, l+32, ... with the reference's
// arithmetic (widen to f64, diff, square, sequential sum, never fused; raw x and raw centroids, no centring), then
// a warp argmin with strict < and lowest index on ties (kmeans.rs:334-347 / bbd_tree.rs:101-111); the row is then
// added to this warp's private partial (lanes own columns), so the result does not depend on scheduling.
// The centroids reach the lanes through a per-warp shared-memory tile of 32 centroids x 32 features (pitch 33 doubles):
// the warp loads each centroid row as ONE coalesced request and every lane then reads ITS centroid conflict-free.
// (Until round 2b a lane read its centroid's features straight from global memory -- 32 different lines per request,
// k*d requests per row: 0.8 ms per marked row at k = 4096, d = 32; 6.7 ms of a 15 ms C5 step for 0.06 % of the rows.)
constexpr int REFINE_TILE = 32 * 33;    // doubles of dynamic shared memory per warp
template <typename TX, int DMMA_WARPS, bool UPDATE = true>
__global__ void __launch_bounds__(DMMA_WARPS * 32)
refine_rows_kernel(const TX* __restrict__ x, uint64_t n, uint32_t d, const double* __restrict__ centroids, uint32_t k,
                   uint32_t* __restrict__ labels, double* __restrict__ mind, double* __restrict__ partials, size_t pk,
                   const unsigned long long* __restrict__ nmarked, const double* __restrict__ mu,
                   const LoopState* __restrict__ loop_st, uint32_t loop_it) {
    pdl_wait();
    if (*nmarked == 0ull || loop_done(loop_st, loop_it)) return; // nothing was marked in this step (the common case)
    extern __shared__ __align__(16) double refine_smem[];
    const int lane = threadIdx.x & 31;
    double* tile = refine_smem + (size_t)(threadIdx.x >> 5) * REFINE_TILE;
    const uint64_t w = (uint64_t)blockIdx.x * DMMA_WARPS + (threadIdx.x >> 5);
    const uint64_t nw = (uint64_t)gridDim.x * DMMA_WARPS;
    const uint64_t per = ((n + nw - 1) / nw + 127) / 128 * 128;
    const uint64_t r_begin = min(n, w * per), r_end = min(n, r_begin + per);
    double* part = partials + w * ((pk + 15) / 16 * 16);
    double inertia = 0.0;
    bool any = false;
    for (uint64_t base4 = r_begin; base4 < r_end; base4 += 128) {
        unsigned marks[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {                            // 4 independent label loads in flight
            const uint64_t mine = base4 + u * 32 + lane;
            marks[u] = (mine < r_end) ? labels[mine] : 0u;
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
        const uint64_t base = base4 + u * 32;
        unsigned todo = __ballot_sync(0xffffffffu, marks[u] == 0xffffffffu);
        while (todo) {
            const int src = __ffs(todo) - 1;
            todo &= todo - 1;
            const uint64_t row = base + src;
            const TX* xr = x + row * d;
            double best = DBL_MAX; uint32_t bi = 0xffffffffu;
            double xv[32];                                           // the row's features of the current 32-feature chunk
            auto load_x = [&](uint32_t j0) {
#pragma unroll
                for (int jj = 0; jj < 32; jj++) xv[jj] = j0 + jj < d ? (double)xr[j0 + jj] : 0.0;
            };
            const bool one_chunk = d <= 32;
            if (one_chunk) load_x(0);
            const uint32_t tile_s = (uint32_t)__cvta_generic_to_shared(tile);
            for (uint32_t c0 = 0; c0 < k; c0 += 32) {                // 32 centroids at a time: lane l takes centroid c0 + l
                double dist = 0.0;
                for (uint32_t j0 = 0; j0 < d; j0 += 32) {            // ... 32 features at a time, in order
                    __syncwarp();
                    // centroid row cl: one coalesced request (lane = feature), straight into the tile; all 32 rows in
                    // flight at once -- a loop of dependent load/store pairs here cost one L2 round trip per ROW
                    if (j0 + lane < d) {
#pragma unroll
                        for (int cl = 0; cl < 32; cl++)
                            if (c0 + cl < k)
                                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(tile_s + (uint32_t)(cl * 33 + lane) * 8u),
                                             "l"(centroids + (size_t)(c0 + cl) * d + j0 + lane) : "memory");
                    }
                    asm volatile("cp.async.commit_group;" ::: "memory");
                    if (!one_chunk) load_x(j0);
                    asm volatile("cp.async.wait_group 0;" ::: "memory");
                    __syncwarp();
                    if (c0 + lane < k) {
                        const double* tr = tile + lane * 33;
#pragma unroll
                        for (int jj = 0; jj < 32; jj++)
                            if (j0 + jj < d) {
                                const double r = __dsub_rn(xv[jj], tr[jj]);
                                dist = __dadd_rn(dist, __dmul_rn(r, r));
                            }
                    }
                }
                if (c0 + lane < k && dist < best) { best = dist; bi = c0 + lane; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double ob = __shfl_xor_sync(0xffffffffu, best, o);
                const uint32_t oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (ob < best || (ob == best && oi < bi)) { best = ob; bi = oi; }
            }
            if (bi == 0xffffffffu) bi = 0;                      // all distances NaN: the reference keeps cluster 0
            if (UPDATE) {
                double* p = part + (size_t)bi * d;
                // (the slots of this step hold sums of x - mu when the tile kernel centred its rows: mu != nullptr)
                for (uint32_t j = lane; j < d; j += 32) __stcg(p + j, __dadd_rn(__ldcg(p + j), mu ? (double)xr[j] - mu[j] : (double)xr[j]));
            }
            if (lane == 0) {
                labels[row] = bi;
                if (UPDATE) {
                    mind[row] = best;
                    double* pc = part + (size_t)k * d + bi;
                    __stcg(pc, __ldcg(pc) + 1.0);
                    inertia = __dadd_rn(inertia, best);
                    any = true;
                }
            }
            __syncwarp();
        }
        }
    }
    if (UPDATE && lane == 0 && any) __stcg(part + pk - 1, __dadd_rn(__ldcg(part + pk - 1), inertia));
}

// mu = per-feature mean of the finite centroids (fixed order: one thread per feature walks the k centroids), the
// centring shift of the GEMM-form kernels.  NaN / inf centroids (an empty initial cluster leaves 0/0, kmeans.rs:288-292)
// are left out; any finite mu is correct, a central one is accurate.
__global__ void center_kernel(const double* __restrict__ centroids, uint32_t k, uint32_t d, double* __restrict__ mu) {
    for (uint32_t j = threadIdx.x; j < d; j += blockDim.x) {
        double s = 0.0; uint32_t cnt = 0;
        for (uint32_t c = 0; c < k; c++) {
            const double v = centroids[(size_t)c * d + j];
            if (isfinite(v)) { s += v; cnt++; }
        }
        const double m = cnt ? s / (double)cnt : 0.0;
        mu[j] = isfinite(m) ? m : 0.0;
    }
}

// ||c - mu||^2 of the centroids currently in ctx->d_centroids (the finalize kernel also writes them, but
// sckm_lloyd_step uploads centroids from the host)
__global__ void cnorm_kernel(const double* __restrict__ centroids, uint32_t k, uint32_t d, const double* __restrict__ mu,
                             double* __restrict__ cnorm) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= k) return;
    double s = 0.0;
    for (uint32_t j = 0; j < d; j++) { const double v = centroids[(size_t)c * d + j] - mu[j]; s = fma(v, v, s); }
    cnorm[c] = s;
}

// Any k >= 2: fewer than 32 centroids leave part of the one 32-column sub-block padded with -inf scores (wasted DMMAs,
// still several times the direct form's rate -- the streaming kernel takes k <= 15 when d <= 32, this kernel when d > 32).
bool dmma_supported(const sckm_dataset* ds, uint64_t k) {
    return ds->d >= 4 && ds->d <= 128 && k >= 2 && k <= (1u << 24) && ds->n < 0xFFFFFFFFull;
}

static unsigned dmma_grid(const sckm_ctx* ctx) { return (unsigned)ctx->num_sms; }

// number of per-warp partial slots the fused kernels may accumulate into (reduced by launch_reduce_partials)
uint32_t dmma_partial_slots(const sckm_ctx* ctx) { return dmma_grid(ctx) * DMMA_MAX_WARPS; }

template <int KSTEPS, int MT, int NT, int WARPS, bool UPDATE, typename TX>
static int launch_t(sckm_dataset* ds, uint64_t k, size_t pk) {
    sckm_ctx* ctx = ds->ctx;
    constexpr int DP = KSTEPS * 4, PITCH = DP + 4, ROWS = 8 * MT;
    const size_t row_bytes = (size_t)PITCH * 8 + 8;                   // staged row + its norm
    const uint32_t kpad = (uint32_t)((k + 8 * NT - 1) / (8 * NT) * (8 * NT));
    uint32_t sl = 1;
    bool multi = false;
    const size_t fixed = 2048 + (size_t)DP * 8;                       // static shared memory + slack, the centring shift
    uint32_t bn = (uint32_t)(((size_t)ctx->smem_optin - fixed) / row_bytes);
    bn = bn / (8 * NT) * (8 * NT);
    if (bn >= kpad) {
        bn = kpad;                                                     // whole centroid set resident, no row state
    } else {
        multi = true;
        // slabs per warp per resident centroid block.  Every block switch costs two CTA barriers, a staging pass and a
        // pipeline refill (fewer, longer phases win), but the rows of a round are re-read once per block and the round
        // already overflows L2 (148 x 12 warps x sl x 8 KB): measured on the C4 shard, fraction of the FP64 peak for
        // sl = 4 / 6 / 8 / 12 / 14 / 16: 0.775 / 0.79 / 0.80 / 0.86 / 0.81 / 0.81 -- a shallow optimum at 12
        sl = 12;
        if (const char* e = getenv("SCKM_DMMA_SL")) sl = (uint32_t)std::max(1, std::min(16, atoi(e)));   // tuning knob
        const size_t state = (size_t)WARPS * sl * ROWS * 3 * sizeof(long long);
        bn = (uint32_t)(((size_t)ctx->smem_optin - fixed - state) / row_bytes);
        bn = bn / (8 * NT) * (8 * NT);
        // balance the blocks: same count, even sizes
        const uint32_t nch = (kpad + bn - 1) / bn;
        bn = ((kpad + nch - 1) / nch + 8 * NT - 1) / (8 * NT) * (8 * NT);
    }
    if (bn < 8 * NT) return fail(ctx, SCKM_ERR_INVALID, "shared memory too small for the DMMA tile");
    const size_t smem = (size_t)bn * row_bytes + (size_t)DP * 8 + (multi ? (size_t)WARPS * sl * ROWS * 3 * sizeof(long long) : 0);
    ctx->partial_slots_used = dmma_grid(ctx) * WARPS;
    if (!multi) {
        auto kern = ctx->center_on ? assign_dmma_resident_kernel<KSTEPS, MT, NT, WARPS, UPDATE, true, TX>
                                   : assign_dmma_resident_kernel<KSTEPS, MT, NT, WARPS, UPDATE, false, TX>;
        SCKM_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<dmma_grid(ctx), WARPS * 32, smem, ctx->stream>>>((const TX*)ds->x, ds->n, (uint32_t)ds->d, ctx->d_centroids,
                                                              ctx->d_cnorm, ctx->d_mu, (uint32_t)k, bn, ds->labels, ds->mind,
                                                              ctx->d_partials, pk, ctx->d_flags, SCKM_LOOP_ARGS(ctx));
    } else {
        auto kern = ctx->center_on ? assign_dmma_kernel<KSTEPS, MT, NT, WARPS, UPDATE, true, TX>
                                   : assign_dmma_kernel<KSTEPS, MT, NT, WARPS, UPDATE, false, TX>;
        SCKM_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<dmma_grid(ctx), WARPS * 32, smem, ctx->stream>>>((const TX*)ds->x, ds->n, (uint32_t)ds->d, ctx->d_centroids,
                                                              ctx->d_cnorm, ctx->d_mu, (uint32_t)k, bn, sl, ds->labels, ds->mind,
                                                              ctx->d_partials, pk, ctx->d_flags, SCKM_LOOP_ARGS(ctx));
    }
    LAUNCH_CHECK_D(ctx);
    {
        auto rk = refine_rows_kernel<TX, WARPS, UPDATE>;
        SCKM_CUDA(ctx, cudaFuncSetAttribute(rk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(WARPS * REFINE_TILE * sizeof(double))));
    }
    refine_rows_kernel<TX, WARPS, UPDATE><<<dmma_grid(ctx), WARPS * 32, WARPS * REFINE_TILE * sizeof(double), ctx->stream>>>((const TX*)ds->x, ds->n, (uint32_t)ds->d,
        ctx->d_centroids, (uint32_t)k, ds->labels, ds->mind, ctx->d_partials, pk, ctx->d_flags, ctx->d_mu, SCKM_LOOP_ARGS(ctx));
    LAUNCH_CHECK_D(ctx);
    return SCKM_OK;
}

template <bool UPDATE, typename TX>
static int launch_by_d(sckm_dataset* ds, uint64_t k, size_t pk) {
    const uint64_t d = ds->d;
    if (d <= 16) return launch_t<4, 2, 4, 12, UPDATE, TX>(ds, k, pk);
    if (d <= 32) return launch_t<8, 2, 4, 12, UPDATE, TX>(ds, k, pk);
    if (d <= 64) return launch_t<16, 2, 4, 12, UPDATE, TX>(ds, k, pk);
    // (d > 64: 8-row slabs, 12 warps.  16-row slabs on 8 warps -- each B fragment feeding two DMMAs, 250 registers -- were
    // measured at 0.57 of the FP64 peak on the C4 shard against 0.78 for this shape: too few warps to cover the row loads.)
    return launch_t<32, 1, 4, 12, UPDATE, TX>(ds, k, pk);
}

// [0] = ||mu||^2, [1] = max_j ||c_j - mu||^2 (NaN norms ignored): the two numbers the centring decision needs
__global__ void center_stats_kernel(const double* __restrict__ mu, uint32_t d, const double* __restrict__ cnorm, uint32_t k,
                                    double* __restrict__ out2) {
    double s = 0.0;
    for (uint32_t j = threadIdx.x; j < d; j += blockDim.x) s = fma(mu[j], mu[j], s);
    const double total = cta_max(cnorm, k);                 // (called by every thread)
    __shared__ double s_sum[32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) s_sum[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (uint32_t w = 0; w < (blockDim.x + 31) / 32; w++) t += s_sum[w];
        out2[0] = t; out2[1] = total;
    }
}

// Centring shift and ||c - mu||^2 of ctx->d_centroids -- only needed when the centroids came from the host (start of
// a fit, sckm_lloyd_step, predict); inside the Lloyd loop mu stays what it was at the start of the fit and the
// finalize kernel keeps the norms current (an unchanged centroid keeps its norm).
// center = true (DMMA tile / streaming kernels): mu = mean of the finite centroids IF the data sit further from the
// origin than they are wide (||mu||^2 > max_j ||c_j - mu||^2; one 16-byte read-back per fit decides), else mu = 0 and
// the tile kernels run their subtraction-free instantiation.  Every rank of a multi-GPU fit holds the same centroids,
// hence takes the same decision.  SCKM_CENTER=0/1 forces it (tests).
// center = false (tcgen05 path, which ranks raw f32 rows): mu = 0, i.e. the raw norms.
int launch_cnorm(sckm_ctx* ctx, uint64_t k, uint64_t d, bool center) {
    if (ctx->cnorm_valid && ctx->mu_requested == center) return SCKM_OK;
    bool on = false;
    if (center) {
        center_kernel<<<1, 128, 0, ctx->stream>>>(ctx->d_centroids, (uint32_t)k, (uint32_t)d, ctx->d_mu);
        cnorm_kernel<<<(unsigned)((k + 127) / 128), 128, 0, ctx->stream>>>(ctx->d_centroids, (uint32_t)k, (uint32_t)d, ctx->d_mu, ctx->d_cnorm);
        center_stats_kernel<<<1, 256, 0, ctx->stream>>>(ctx->d_mu, (uint32_t)d, ctx->d_cnorm, (uint32_t)k, (double*)(ctx->d_flags + 2));
        ctx->launches += 3;
        SCKM_CUDA(ctx, cudaMemcpyAsync(ctx->h_pinned + 8, ctx->d_flags + 2, 2 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        SCKM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        on = ctx->h_pinned[8] > ctx->h_pinned[9];
        if (const char* e = getenv("SCKM_CENTER")) on = atoi(e) != 0;
    }
    if (!on) {
        SCKM_CUDA(ctx, cudaMemsetAsync(ctx->d_mu, 0, d * sizeof(double), ctx->stream));
        cnorm_kernel<<<(unsigned)((k + 127) / 128), 128, 0, ctx->stream>>>(ctx->d_centroids, (uint32_t)k, (uint32_t)d, ctx->d_mu, ctx->d_cnorm);
        ctx->launches++;
    }
    LAUNCH_CHECK_D(ctx);
    ctx->launches--;                                          // (LAUNCH_CHECK_D counted one that was already counted)
    ctx->mu_requested = center;
    ctx->center_on = on;
    ctx->mu_zero = !on;
    ctx->cnorm_valid = true;
    return SCKM_OK;
}

// refine pass for the streaming kernel's launch geometry (8 warps per CTA)
int launch_refine_rows(sckm_dataset* ds, uint64_t k, size_t pk, unsigned grid_ctas) {
    sckm_ctx* ctx = ds->ctx;
    constexpr size_t smem = 8 * REFINE_TILE * sizeof(double);
    SCKM_CUDA(ctx, cudaFuncSetAttribute(refine_rows_kernel<float, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SCKM_CUDA(ctx, cudaFuncSetAttribute(refine_rows_kernel<double, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (ds->dtype == SCKM_F32)
        SCKM_CUDA(ctx, launch_pdl(refine_rows_kernel<float, 8>, dim3(grid_ctas), dim3(256), smem, ctx->stream, (const float*)ds->x, ds->n, (uint32_t)ds->d,
            (const double*)ctx->d_centroids, (uint32_t)k, ds->labels, ds->mind, ctx->d_partials, pk, (const unsigned long long*)ctx->d_flags,
            (const double*)nullptr, SCKM_LOOP_ARGS(ctx)));
    else
        SCKM_CUDA(ctx, launch_pdl(refine_rows_kernel<double, 8>, dim3(grid_ctas), dim3(256), smem, ctx->stream, (const double*)ds->x, ds->n, (uint32_t)ds->d,
            (const double*)ctx->d_centroids, (uint32_t)k, ds->labels, ds->mind, ctx->d_partials, pk, (const unsigned long long*)ctx->d_flags,
            (const double*)nullptr, SCKM_LOOP_ARGS(ctx)));
    LAUNCH_CHECK_D(ctx);
    return SCKM_OK;
}

// labels + mind + per-warp partial [sums | counts | inertia] (fused update); the caller reduces the slots.
int launch_assign_dmma(sckm_dataset* ds, uint64_t k) {
    sckm_ctx* ctx = ds->ctx;
    if (!dmma_supported(ds, k)) return fail(ctx, SCKM_ERR_INVALID, "shape not supported by the DMMA kernel");
    const size_t pk = (size_t)k * ds->d + k + 1;
    SCKM_TRY(ensure_workspace(ctx, k, ds->d, dmma_partial_slots(ctx)));
    if (ds->n == 0) return SCKM_OK;
    SCKM_TRY(launch_cnorm(ctx, k, ds->d, true));
    ctx->packed_centered = ctx->center_on;                         // the slots receive sums of x - mu
    return ds->dtype == SCKM_F32 ? launch_by_d<true, float>(ds, k, pk) : launch_by_d<true, double>(ds, k, pk);
}

// ---- per-label sums and counts for GIVEN labels (initial centroids, kmeans.rs:275-292) at HBM rate ----
// Same row -> lane mapping and the same deterministic warp-private accumulation as the fused update of the tile
// kernel (rows of an m-tile that share a label are serialised in ascending row order), without any distance work.
template <int KSTEPS, int MT, int DMMA_WARPS, typename TX>
__global__ void __launch_bounds__(DMMA_WARPS * 32)
update_given_kernel(const TX* __restrict__ x, uint64_t n, uint32_t d, uint32_t k, const uint32_t* __restrict__ labels,
                    double* __restrict__ partials, size_t pk) {
    constexpr int ROWS = 8 * MT;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    const uint64_t nslabs = (n + ROWS - 1) / ROWS;
    const uint64_t stride = (uint64_t)gridDim.x * DMMA_WARPS;
    double* part = partials + ((size_t)blockIdx.x * DMMA_WARPS + warp) * ((pk + 15) / 16 * 16);
    unsigned lanemask_lt;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(lanemask_lt));
    for (uint64_t slab = (uint64_t)blockIdx.x * DMMA_WARPS + warp; slab < nslabs; slab += stride) {
        const uint64_t r0 = slab * ROWS;
        double a[MT][KSTEPS];
        uint32_t lab[MT];
#pragma unroll
        for (int mt = 0; mt < MT; mt++) {
            const uint64_t row = r0 + mt * 8 + g;
            const bool rok = row < n;
            const TX* xr = x + row * d;
            lab[mt] = rok ? labels[row] : 0xffffffffu;
#pragma unroll
            for (int ks = 0; ks < KSTEPS; ks++) {
                const uint32_t col = ks * 4 + t;
                a[mt][ks] = (rok && col < d) ? (double)__ldg(xr + col) : 0.0;
            }
        }
#pragma unroll
        for (int mt = 0; mt < MT; mt++) {
            const bool ok = lab[mt] < k;                        // rows past the end and labels outside [0,k) add nothing
            const uint32_t key = ok ? lab[mt] : (0x80000000u | (uint32_t)g);
            const unsigned peers = __match_any_sync(0xffffffffu, key);
            const int rank = __popc(peers & lanemask_lt) >> 2;
            const int maxrank = __reduce_max_sync(0xffffffffu, ok ? rank : 0);
            for (int r = 0; r <= maxrank; r++) {
                if (r) { __threadfence(); __syncwarp(); }
                if (ok && rank == r) {
                    double* p = part + (size_t)lab[mt] * d + t;
#pragma unroll
                    for (int ks = 0; ks < KSTEPS; ks++)
                        if (ks * 4 + t < d) atomicAdd(p + ks * 4, a[mt][ks]);
                    if (t == 0) atomicAdd(part + (size_t)k * d + lab[mt], 1.0);
                }
            }
        }
    }
}

bool update_given_supported(const sckm_dataset* ds, uint64_t k) { return ds->d >= 4 && ds->d <= 128 && k >= 1; }

template <int KSTEPS, typename TX>
static int update_given_t(sckm_dataset* ds, uint64_t k, size_t pk) {
    sckm_ctx* ctx = ds->ctx;
    constexpr int WARPS = 12, MT = 2;
    update_given_kernel<KSTEPS, MT, WARPS, TX><<<dmma_grid(ctx), WARPS * 32, 0, ctx->stream>>>(
        (const TX*)ds->x, ds->n, (uint32_t)ds->d, (uint32_t)k, ds->labels, ctx->d_partials, pk);
    LAUNCH_CHECK_D(ctx);
    return launch_reduce_partials(ctx, dmma_grid(ctx) * WARPS, pk);
}

// sums/counts of the dataset's current labels into ctx->d_packed (inertia slot = 0)
int launch_update_given(sckm_dataset* ds, uint64_t k) {
    sckm_ctx* ctx = ds->ctx;
    const size_t pk = (size_t)k * ds->d + k + 1;
    SCKM_TRY(ensure_workspace(ctx, k, ds->d, dmma_partial_slots(ctx)));
    const uint64_t d = ds->d;
    const bool f32 = ds->dtype == SCKM_F32;
    if (d <= 16) return f32 ? update_given_t<4, float>(ds, k, pk) : update_given_t<4, double>(ds, k, pk);
    if (d <= 32) return f32 ? update_given_t<8, float>(ds, k, pk) : update_given_t<8, double>(ds, k, pk);
    if (d <= 64) return f32 ? update_given_t<16, float>(ds, k, pk) : update_given_t<16, double>(ds, k, pk);
    return f32 ? update_given_t<32, float>(ds, k, pk) : update_given_t<32, double>(ds, k, pk);
}

// labels only (KMeans::predict at scale): same ranking + exact re-decision of near-ties, no update, no distances.
// For every row the result equals the exact direct-form argmin (kmeans.rs:334-347).
int launch_predict_dmma(sckm_dataset* ds, uint64_t k) {
    sckm_ctx* ctx = ds->ctx;
    if (!dmma_supported(ds, k)) return fail(ctx, SCKM_ERR_INVALID, "shape not supported by the DMMA kernel");
    const size_t pk = (size_t)k * ds->d + k + 1;
    SCKM_TRY(ensure_workspace(ctx, k, ds->d, 0));                 // no partial sums in this mode
    if (ds->n == 0) return SCKM_OK;
    SCKM_TRY(launch_cnorm(ctx, k, ds->d, true));
    SCKM_TRY(ds->dtype == SCKM_F32 ? (launch_by_d<false, float>(ds, k, pk)) : (launch_by_d<false, double>(ds, k, pk)));
    // the marked-row counter is normally cleared by the reduce of the step; there is none here
    SCKM_CUDA(ctx, cudaMemsetAsync(ctx->d_flags, 0, sizeof(unsigned long long), ctx->stream));
    return SCKM_OK;
}

}  // namespace sckm
