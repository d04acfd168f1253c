// bench/tc5_probe.cu -- validates the tcgen05 building blocks used by the f32 k-means tile kernel on B200:
// TMA (SWIZZLE_128B) -> shared memory -> tcgen05.mma kind::tf32 (K-major SW128 descriptors) -> TMEM -> tcgen05.ld.
// Computes S = X . C^T for one 128 x 128 tile with K = 32, once with plain TF32 and once as 3xTF32
// (hi*hi + hi*lo + lo*hi), and compares with a double-precision CPU result.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tc5_probe tc5_probe.cu   (no libcuda link: driver entry point)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile("{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n }"
                 ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
// K-major, SWIZZLE_128B operand descriptor (cute::UMMA::SmemDescriptor): start>>4 | LBO(1)<<16 | SBO(1024>>4)<<32 | version 1<<46 | layout 2<<61
__device__ __forceinline__ uint64_t umma_desc_sw128(const void* smem) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_u32(smem) >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_c, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n }"
                 ::"r"(tmem_c), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

constexpr int M = 128, N = 128, K = 32;
constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);

__global__ void __launch_bounds__(128) probe(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapC,
                                             float* out, int mode) {
    extern __shared__ __align__(1024) unsigned char smem[];
    float* sXh = reinterpret_cast<float*>(smem);                 // 16 KB each, 1024-byte aligned
    float* sXl = sXh + M * K;
    float* sCh = sXl + M * K;
    float* sCl = sCh + N * K;
    __shared__ uint64_t bar_tma, bar_mma;
    __shared__ uint32_t tmem_base;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) { mbar_init(&bar_tma, 1); mbar_init(&bar_mma, 1); asm volatile("fence.mbarrier_init.release.cluster;"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(128));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = tmem_base;

    if (threadIdx.x == 0) {
        mbar_expect_tx(&bar_tma, (M + N) * K * 4);
        tma_load_2d(sXh, &mapX, 0, 0, &bar_tma);
        tma_load_2d(sCh, &mapC, 0, 0, &bar_tma);
    }
    mbar_wait(&bar_tma, 0);
    // split in place: hi = x with the low 13 mantissa bits cleared (what kind::tf32 reads), lo = x - hi (exact)
    for (int i = threadIdx.x; i < M * K; i += blockDim.x) {
        float x = sXh[i], h = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
        if (mode == 1) sXh[i] = h;
        sXl[i] = x - h;
        float c = sCh[i], hc = __uint_as_float(__float_as_uint(c) & 0xFFFFE000u);
        if (mode == 1) sCh[i] = hc;
        sCl[i] = c - hc;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy writes -> visible to the MMA (async proxy)
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;");
        const uint64_t dXh = umma_desc_sw128(sXh), dXl = umma_desc_sw128(sXl), dCh = umma_desc_sw128(sCh), dCl = umma_desc_sw128(sCl);
        uint32_t acc = 0;
        for (int ks = 0; ks < K / 8; ks++) {                          // UMMA_K = 8 tf32 = 32 bytes -> +2 in 16-byte units
            umma_tf32(tmem, dXh + 2 * ks, dCh + 2 * ks, IDESC, acc); acc = 1;
        }
        if (mode == 1) {
            for (int ks = 0; ks < K / 8; ks++) umma_tf32(tmem, dXh + 2 * ks, dCl + 2 * ks, IDESC, 1);
            for (int ks = 0; ks < K / 8; ks++) umma_tf32(tmem, dXl + 2 * ks, dCh + 2 * ks, IDESC, 1);
        }
        umma_commit(&bar_mma);
    }
    mbar_wait(&bar_mma, 0);
    asm volatile("tcgen05.fence::after_thread_sync;");
    // TMEM -> registers: warp w reads lanes 32w..32w+31 (one row per thread), 32 columns per instruction
    for (int c0 = 0; c0 < N; c0 += 32) {
        uint32_t v[32];
        const uint32_t addr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                       "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]),
                       "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]),
                       "=r"(v[30]), "=r"(v[31]) : "r"(addr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 32; j++) out[(warp * 32 + lane) * N + c0 + j] = __uint_as_float(v[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128));
}

typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    void* fn = nullptr; cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    EncodeTiled encode = (EncodeTiled)fn;
    std::vector<float> hX(M * K), hC(N * K);
    srand(1);
    for (auto& v : hX) v = (float)rand() / RAND_MAX * 20.f - 10.f;
    for (auto& v : hC) v = (float)rand() / RAND_MAX * 20.f - 10.f;
    float *dX, *dC, *dO;
    CK(cudaMalloc(&dX, hX.size() * 4)); CK(cudaMalloc(&dC, hC.size() * 4)); CK(cudaMalloc(&dO, M * N * 4));
    CK(cudaMemcpy(dX, hX.data(), hX.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dC, hC.data(), hC.size() * 4, cudaMemcpyHostToDevice));
    CUtensorMap mapX, mapC;
    cuuint64_t dims[2] = {K, M}, strides[1] = {K * 4};
    cuuint32_t box[2] = {K, M}, estr[2] = {1, 1};
    CUresult r1 = encode(&mapX, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dX, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CUresult r2 = encode(&mapC, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dC, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r1 || r2) { printf("cuTensorMapEncodeTiled failed %d %d\n", r1, r2); return 1; }
    const size_t smem = 4 * M * K * 4 + 1024;
    CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    std::vector<float> hO(M * N);
    for (int mode = 0; mode < 2; mode++) {
        probe<<<1, 128, smem>>>(mapX, mapC, dO, mode);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(hO.data(), dO, hO.size() * 4, cudaMemcpyDeviceToHost));
        double maxerr = 0, maxref = 0;
        for (int i = 0; i < M; i++)
            for (int j = 0; j < N; j++) {
                double s = 0;
                for (int k = 0; k < K; k++) s += (double)hX[i * K + k] * (double)hC[j * K + k];
                maxerr = fmax(maxerr, fabs(s - hO[i * N + j])); maxref = fmax(maxref, fabs(s));
            }
        printf("%s: max |err| = %.3e (max |ref| = %.1f, relative %.2e)\n", mode ? "3xTF32" : "TF32  ", maxerr, maxref, maxerr / maxref);
    }
    return 0;
}
