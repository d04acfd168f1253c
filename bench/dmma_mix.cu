// bench/dmma_mix.cu -- does scalar FP64 work (DADD/DSETP) steal DMMA throughput on B200?
// Per loop iteration: 8 independent DMMAs + NF scalar FP64 ops (or NI integer ops as a control).
#include <cstdio>
#include <cuda_runtime.h>
template <int NF, int NI, int KIND>
__global__ void k(double* out, int iters) {
    double a = threadIdx.x * 1e-3, b = threadIdx.x * 2e-3;
    double c[8][2];
#pragma unroll
    for (int i = 0; i < 8; i++) { c[i][0] = i; c[i][1] = -i; }
    double f[8]; unsigned u[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { f[i] = threadIdx.x + i; u[i] = threadIdx.x * 7 + i; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
#pragma unroll
        for (int j = 0; j < NF; j++) {
            if (KIND == 0) f[j % 8] = __dadd_rn(f[j % 8], 1.25);                       // DADD
            else { bool p = f[j % 8] > f[(j + 1) % 8]; u[j % 8] += p ? 3u : 5u; }      // DSETP + int
        }
#pragma unroll
        for (int j = 0; j < NI; j++) u[j % 8] = u[j % 8] * 1664525u + 1013904223u;     // IMAD
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += c[i][0] + c[i][1] + f[i] + u[i];
    if (s == 12345.678) out[0] = s;
}
template <int NF, int NI, int KIND> void run(const char* name, double* d, int sms, int warps) {
    int iters = 4096;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9;
    for (int r = 0; r < 3; r++) {
        cudaEventRecord(e0); k<NF, NI, KIND><<<sms, warps * 32>>>(d, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    double dm = (double)sms * warps * iters * 8 * 512.0 / best / 1e9;
    printf("%-28s warps/SM=%2d  DMMA %.2f TFLOP/s  (%.1f%% of 37.0)  extra ops/DMMA: fp64 %.2f int %.2f\n", name, warps, dm,
           100 * dm / 37.0, NF / 8.0, NI / 8.0);
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0); double* d; cudaMalloc(&d, 64); int s = p.multiProcessorCount;
    for (int w : {8, 12}) {
        run<0, 0, 0>("pure DMMA", d, s, w);
        run<2, 0, 0>("+DADD 0.25/DMMA", d, s, w);
        run<4, 0, 0>("+DADD 0.5/DMMA", d, s, w);
        run<8, 0, 0>("+DADD 1/DMMA", d, s, w);
        run<16, 0, 0>("+DADD 2/DMMA", d, s, w);
        run<4, 0, 1>("+DSETP 0.5/DMMA", d, s, w);
        run<8, 0, 1>("+DSETP 1/DMMA", d, s, w);
        run<0, 16, 0>("+IMAD 2/DMMA", d, s, w);
        run<0, 64, 0>("+IMAD 8/DMMA", d, s, w);
    }
    return 0;
}
