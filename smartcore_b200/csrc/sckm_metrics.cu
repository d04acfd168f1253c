// sckm_metrics.cu -- contingency table of (class, cluster) pairs from labels that are already resident on the
// device: the input of the reference's cluster-quality scores (src/metrics/cluster_helpers.rs:7-25
// contingency_matrix; entropy :27-48 and mutual_info_score :50-104 work on its row/column sums; HCVScore
// src/metrics/cluster_hcv.rs:36-55).  SURVEY.md section 8(f) rank 4: the step after the fit, without bringing n
// labels back to the host.  Integer counting: bit-exact by construction.
#include "sckm_common.cuh"
#include <algorithm>

namespace sckm {

constexpr int CONT_THREADS = 256;
constexpr uint32_t CONT_SMEM_CELLS = 8192;     // 32 KB of u32 counters per CTA

// a: class id of each row, b: cluster id of each row; out[na][nb] += 1.  Ids outside the table are counted in
// *bad (the caller turns that into an error) and otherwise ignored.
template <bool SMEM>
__global__ void __launch_bounds__(CONT_THREADS)
contingency_kernel(const uint32_t* __restrict__ a, const uint32_t* __restrict__ b, uint64_t n, uint32_t na, uint32_t nb,
                   unsigned long long* __restrict__ out, unsigned long long* __restrict__ bad) {
    __shared__ uint32_t cells[SMEM ? CONT_SMEM_CELLS : 1];
    const uint32_t ncell = na * nb;
    if (SMEM) {
        for (uint32_t i = threadIdx.x; i < ncell; i += blockDim.x) cells[i] = 0;
        __syncthreads();
    }
    uint32_t nbad = 0;
    // a CTA sees at most 2^32-1 rows (grid sized by the launcher), so the shared u32 counters cannot wrap
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t ai = a[i], bi = b[i];
        if (ai >= na || bi >= nb) { nbad++; continue; }
        if (SMEM) atomicAdd(&cells[ai * nb + bi], 1u);
        else atomicAdd(&out[(size_t)ai * nb + bi], 1ull);
    }
    if (nbad) atomicAdd(bad, (unsigned long long)nbad);
    if (SMEM) {
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < ncell; i += blockDim.x)
            if (cells[i]) atomicAdd(&out[i], (unsigned long long)cells[i]);
    }
}

// d_out: [na*nb + 1] u64 on the device, zeroed here; the last word counts out-of-range ids
int launch_contingency(sckm_ctx* ctx, const uint32_t* d_a, const uint32_t* d_b, uint64_t n, uint64_t na, uint64_t nb,
                       unsigned long long* d_out) {
    const uint64_t ncell = na * nb;
    SCKM_CUDA(ctx, cudaMemsetAsync(d_out, 0, (ncell + 1) * sizeof(unsigned long long), ctx->stream));
    if (n == 0) return SCKM_OK;
    const unsigned grid = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((n + CONT_THREADS - 1) / CONT_THREADS,
                                                                              (uint64_t)ctx->num_sms * 8));
    if (ncell <= CONT_SMEM_CELLS)
        contingency_kernel<true><<<grid, CONT_THREADS, 0, ctx->stream>>>(d_a, d_b, n, (uint32_t)na, (uint32_t)nb, d_out, d_out + ncell);
    else
        contingency_kernel<false><<<grid, CONT_THREADS, 0, ctx->stream>>>(d_a, d_b, n, (uint32_t)na, (uint32_t)nb, d_out, d_out + ncell);
    ctx->launches++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(ctx, SCKM_ERR_CUDA, "contingency kernel launch failed: %s", cudaGetErrorString(e));
    return SCKM_OK;
}

}  // namespace sckm
