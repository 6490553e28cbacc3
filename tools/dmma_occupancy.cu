// DMMA m8n8k4 throughput vs warps per SM and independent accumulators per warp (register operands).
// Measurement tool (DESIGN.md K4 notes).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_occupancy dmma_occupancy.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int NACC>
__global__ void k(double* out, int iters, double seed) {
    double c[NACC][2];
#pragma unroll
    for (int j = 0; j < NACC; ++j) c[j][0] = c[j][1] = 0.0;
    double a = seed + threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-4;
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 16 / NACC; ++r)
#pragma unroll
            for (int j = 0; j < NACC; ++j) dmma(c[j][0], c[j][1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int j = 0; j < NACC; ++j) s += c[j][0] + c[j][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NACC>
float run(double* out, int warps, int sms) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int iters = 8192;
    k<NACC><<<sms, warps * 32>>>(out, iters, 1.0); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) { cudaEventRecord(e0); k<NACC><<<sms, warps * 32>>>(out, iters, 1.0); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); best = ms < best ? ms : best; }
    double flops = 2.0 * 256 * 16.0 * iters * warps * sms;
    return (float)(flops / best * 1e-9);
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    double* out; cudaMalloc(&out, sizeof(double) * sms * 1024);
    printf("{\"what\": \"DMMA TFLOP/s, one CTA per SM\"");
    for (int warps : {4, 8, 12, 16, 32}) {
        printf(", \"w%d_acc1\": %.2f", warps, run<1>(out, warps, sms));
        printf(", \"w%d_acc2\": %.2f", warps, run<2>(out, warps, sms));
        printf(", \"w%d_acc4\": %.2f", warps, run<4>(out, warps, sms));
        printf(", \"w%d_acc8\": %.2f", warps, run<8>(out, warps, sms));
    }
    printf("}\n");
    return 0;
}
