// FP64 ceiling probe for B200 (sm_100a): DFMA issue rate, DMMA m8n8k4 / m16n8k8 / m16n8k16 rate.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu
// Prints one JSON object. Measurement tool only (DESIGN.md "FP64 roofline denominators").
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__global__ void __launch_bounds__(256) k_dfma(double* out, int iters, double seed) {
    double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double b = 1.0000001, c = 1e-9;
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
            a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1688(double* c, const double* a, const double* b) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void dmma16816(double* c, const double* a, const double* b) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                   "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

__global__ void __launch_bounds__(256) k_dmma884(double* out, int iters, double seed) {
    double c[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) c[j] = 0.0;
    double a = seed + threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-4;
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) dmma884(c[2 * j], c[2 * j + 1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int j = 0; j < 16; ++j) s += c[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void __launch_bounds__(256) k_dmma1688(double* out, int iters, double seed) {
    double c[16], a[4], b[2];
#pragma unroll
    for (int j = 0; j < 16; ++j) c[j] = 0.0;
#pragma unroll
    for (int j = 0; j < 4; ++j) a[j] = seed + threadIdx.x * 1e-3 + j;
    b[0] = 1.0 + threadIdx.x * 1e-4; b[1] = b[0] * 0.5;
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma1688(c + 4 * j, a, b);
    }
    double s = 0;
#pragma unroll
    for (int j = 0; j < 16; ++j) s += c[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void __launch_bounds__(256) k_dmma16816(double* out, int iters, double seed) {
    double c[16], a[8], b[4];
#pragma unroll
    for (int j = 0; j < 16; ++j) c[j] = 0.0;
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = seed + threadIdx.x * 1e-3 + j;
#pragma unroll
    for (int j = 0; j < 4; ++j) b[j] = 1.0 + threadIdx.x * 1e-4 + j;
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma16816(c + 4 * j, a, b);
    }
    double s = 0;
#pragma unroll
    for (int j = 0; j < 16; ++j) s += c[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// Are the FP64 FMA pipe and the DMMA path separate issue resources?  Even warps run the DFMA loop, odd warps the DMMA loop
// (mode 0), or every warp interleaves both (mode 1).  If the aggregate exceeds either peak alone they overlap.
__global__ void __launch_bounds__(256) k_mixed(double* out, int iters, double seed, int mode) {
    double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double b = 1.0000001, cc = 1e-9;
    double c[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) c[j] = 0.0;
    double ma = seed + threadIdx.x * 1e-3, mb = 1.0 + threadIdx.x * 1e-4;
    const bool do_fma = mode == 1 || ((threadIdx.x >> 5) & 1) == 0, do_mma = mode == 1 || ((threadIdx.x >> 5) & 1) == 1;
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
        if (do_fma) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                a0 = fma(a0, b, cc); a1 = fma(a1, b, cc); a2 = fma(a2, b, cc); a3 = fma(a3, b, cc);
                a4 = fma(a4, b, cc); a5 = fma(a5, b, cc); a6 = fma(a6, b, cc); a7 = fma(a7, b, cc);
            }
        }
        if (do_mma) {
#pragma unroll
            for (int j = 0; j < 8; ++j) dmma884(c[2 * j], c[2 * j + 1], ma, mb);
        }
    }
    double s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
#pragma unroll
    for (int j = 0; j < 16; ++j) s += c[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static float time_ms(F launch, int reps) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch(); launch(); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int sms = p.multiProcessorCount;
    double* out; CK(cudaMalloc(&out, sizeof(double) * sms * 8 * 256));
    printf("{\"gpu\": \"%s\", \"sms\": %d", p.name, sms);
    for (int bps = 2; bps <= 8; bps *= 2) {
        int grid = sms * bps, iters = 4096;
        float ms = time_ms([&] { k_dfma<<<grid, 256>>>(out, iters, 1.0); }, 5);
        double flops = 2.0 * 64 * iters * 256.0 * grid;
        printf(", \"dfma_tflops_bps%d\": %.2f", bps, flops / ms * 1e-9);
        ms = time_ms([&] { k_dmma884<<<grid, 256>>>(out, iters, 1.0); }, 5);
        flops = 2.0 * 256 * 8 * iters * 8.0 * grid;   // 8 warps/CTA, 8 mma/iter, 256 FMA each
        printf(", \"dmma884_tflops_bps%d\": %.2f", bps, flops / ms * 1e-9);
        ms = time_ms([&] { k_dmma1688<<<grid, 256>>>(out, iters, 1.0); }, 5);
        flops = 2.0 * 1024 * 4 * iters * 8.0 * grid;
        printf(", \"dmma1688_tflops_bps%d\": %.2f", bps, flops / ms * 1e-9);
        ms = time_ms([&] { k_dmma16816<<<grid, 256>>>(out, iters, 1.0); }, 5);
        flops = 2.0 * 2048 * 4 * iters * 8.0 * grid;
        printf(", \"dmma16816_tflops_bps%d\": %.2f", bps, flops / ms * 1e-9);
    }
    for (int mode = 0; mode < 2; ++mode) {
        int grid = sms * 4, iters = 4096;
        float ms = time_ms([&] { k_mixed<<<grid, 256>>>(out, iters, 1.0, mode); }, 5);
        // per warp and iteration: 64 DFMA x 32 lanes (2048 FMA) and / or 8 DMMA x 256 FMA (2048 FMA)
        double fma_warps = mode == 1 ? 8.0 : 4.0, mma_warps = mode == 1 ? 8.0 : 4.0;
        double flops = 2.0 * 2048.0 * iters * (fma_warps + mma_warps) * grid;
        printf(", \"mixed_dfma_dmma_tflops_mode%d\": %.2f", mode, flops / ms * 1e-9);
    }
    CK(cudaGetLastError());
    printf("}\n");
    return 0;
}
