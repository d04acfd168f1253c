// bench/dmma_micro.cu -- FP64 DMMA pipe characterisation on B200 (single-warp issue limits, ILP, latency).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_micro dmma_micro.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void k(double* out, int iters) {
    double a = threadIdx.x * 1e-3, b = threadIdx.x * 2e-3;
    double c[ILP][2];
#pragma unroll
    for (int i = 0; i < ILP; i++) { c[i][0] = i; c[i][1] = -i; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += c[i][0] + c[i][1];
    if (s == 12345.678) out[0] = s;
}
template <int ILP> void run(int warps_per_sm, double* d, int sms) {
    int iters = 8192 / ILP * 8;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9;
    for (int r = 0; r < 3; r++) {
        cudaEventRecord(e0);
        k<ILP><<<sms, warps_per_sm * 32>>>(d, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    double flops = (double)sms * warps_per_sm * iters * ILP * 512.0;
    printf("ILP=%2d warps/SM=%2d  %.2f TFLOP/s  (%.1f clk per DMMA per warp at 1.965 GHz)\n", ILP, warps_per_sm,
           flops / best / 1e9, best * 1e-3 * 1.965e9 / ((double)iters * ILP));
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    double* d; cudaMalloc(&d, 64);
    for (int w : {1, 4, 8, 12, 16, 32}) { run<1>(w, d, p.multiProcessorCount); }
    for (int w : {4, 8, 12, 16}) { run<2>(w, d, p.multiProcessorCount); run<4>(w, d, p.multiProcessorCount); run<8>(w, d, p.multiProcessorCount); run<16>(w, d, p.multiProcessorCount); }
    return 0;
}
