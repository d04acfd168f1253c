// sckm_tile.cuh -- device helpers shared by the FP64 tensor-path kernels (sckm_dmma.cu, sckm_stream.cu).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace sckm {

// D(8x8) += A(8x4) * B(4x8) in f64: SASS DMMA.8x8x4.  Fragment layout (PTX ISA, m8n8k4 .f64), g = lane/4, t = lane%4:
//   A: row g, col t;   B: row t, col g;   C/D: row g, cols 2t and 2t+1.
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}



// Order-preserving map double -> int64 (an involution on the bit pattern): scalar FP64 instructions share the
// datapath with DMMA on B200 (bench/dmma_mix.cu: one DADD per DMMA costs 14 % of the DMMA rate, eight IMADs 2 %),
// so the whole top-2 tracking of the epilogue runs on integer keys.  NaNs sort to the extremes and are caught by
// the tie test at the end (gap is NaN -> exact re-decision).
typedef long long key_t;
__device__ __forceinline__ key_t dkey(double v) {
    const key_t b = __double_as_longlong(v);
    return b ^ ((b >> 63) & 0x7fffffffffffffffLL);
}
__device__ __forceinline__ double dunkey(key_t k) { return __longlong_as_double(k ^ ((k >> 63) & 0x7fffffffffffffffLL)); }
constexpr key_t KEY_MIN = (key_t)0x8000000000000000ULL;

// max of v[0..k) computed by the whole CTA (every thread must call it, every thread gets the result)
__device__ __forceinline__ double cta_max(const double* __restrict__ v, uint32_t k) {
    __shared__ double s_part[32];
    __shared__ double s_all;
    double m = 0.0;
    for (uint32_t j = threadIdx.x; j < k; j += blockDim.x) m = fmax(m, v[j]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (uint32_t w = 0; w < (blockDim.x + 31) / 32; w++) t = fmax(t, s_part[w]);
        s_all = t;
    }
    __syncthreads();
    return s_all;
}

}  // namespace sckm
