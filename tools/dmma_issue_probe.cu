// Does a stream of FP64 DMMAs (mma.sync.m8n8k4.f64) keep the other warps of its scheduler from issuing?
// The triangular sweeps interleave DMMA loops with ~80 bookkeeping instructions per item and warp; this probe measures how
// long a dependent integer chain / a shared-memory load chain takes on a warp whose scheduler neighbours issue DMMAs back
// to back, and what the neighbours lose.  One block of 16 warps per SM (warp w -> scheduler w % 4), like a sweep CTA.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/dmma_issue_probe tools/dmma_issue_probe.cu && tools/dmma_issue_probe
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// mode of the non-DMMA warps: 0 = dependent integer chain, 1 = dependent shared-memory load chain (pointer chase)
__global__ void __launch_bounds__(1024, 1) k_probe(int n_dmma, int n_other, int mode, int iters, long long* out, double* sink) {
    __shared__ int chase[1024];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, slot = warp >> 2;     // slot: 0..3 = which of the scheduler's four warps
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) chase[i] = (i * 37 + 11) & 1023;
    __syncthreads();
    long long t0 = 0, t1 = 0;
    if (slot < n_dmma) {
        double c[2][2] = {{0, 0}, {0, 0}};
        const double a = 1.0 + lane * 1e-3, b = 1.0 - lane * 1e-3;
        t0 = clock64();
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int k = 0; k < 16; ++k) { dmma(c[0][0], c[0][1], a, b); dmma(c[1][0], c[1][1], b, a); }     // two chains of 16, like a dense item
        }
        t1 = clock64();
        if (c[0][0] + c[1][1] == 123.456) sink[0] = c[0][1];
    } else if (slot < n_dmma + n_other) {
        int x = lane;
        t0 = clock64();
        if (mode == 0) {
            for (int it = 0; it < iters; ++it) {
#pragma unroll
                for (int k = 0; k < 32; ++k) x = x * 3 + (x >> 5) + k;                                       // 64 dependent integer instructions
            }
        } else {
            for (int it = 0; it < iters; ++it) {
#pragma unroll
                for (int k = 0; k < 16; ++k) x = chase[x & 1023];                                             // 16 dependent shared-memory loads
            }
        }
        t1 = clock64();
        if (x == -12345) sink[1] = x;
    }
    if (lane == 0 && blockIdx.x == 0) out[warp] = t1 - t0;
}

int main() {
    long long* d_out; double* d_sink;
    cudaMalloc(&d_out, 16 * sizeof(long long)); cudaMalloc(&d_sink, 16);
    int dev = 0, sms = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int iters = 2000;
    printf("16 warps per SM, warp w on scheduler w %% 4; per 'item': DMMA warp = 32 DMMAs (2 chains of 16), other warp = 64 dependent integer ops | 16 dependent LDS\n");
    printf("%-10s %-8s %-8s | %-22s | %-22s\n", "mode", "dmma/sch", "other/sch", "clk per DMMA item", "clk per other item");
    for (int mode = 0; mode < 2; ++mode)
        for (int nd = 0; nd <= 4; ++nd)
            for (int no = 0; no + nd <= 4; ++no) {
                if (nd == 0 && no == 0) continue;
                if (!(no == 0 || nd == 0 || nd + no == 4 || (nd == 1 && no == 1))) continue;
                cudaMemset(d_out, 0, 16 * sizeof(long long));
                k_probe<<<sms, 512>>>(nd, no, mode, iters, d_out, d_sink);
                if (cudaDeviceSynchronize() != cudaSuccess) { printf("launch failed\n"); return 1; }
                long long h[16]; cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
                double sd = 0, so = 0; int cd = 0, co = 0;
                for (int w = 0; w < 16; ++w) { const int slot = w >> 2; if (slot < nd) { sd += (double)h[w]; ++cd; } else if (slot < nd + no) { so += (double)h[w]; ++co; } }
                printf("%-10s %-8d %-8d | %-22.1f | %-22.1f\n", mode == 0 ? "int-chain" : "lds-chain", nd, no, cd ? sd / cd / iters : 0.0, co ? so / co / iters : 0.0);
            }
    // plain throughput by CUDA events: all warps issue DMMAs, 8 / 16 DMMA warps per block, 1 or 2 blocks per SM, 2 chains per warp
    printf("\nDMMA throughput by CUDA events (TFLOP/s at the clock the GPU actually ran):\n");
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int threads = 256; threads <= 1024; threads *= 2)
        for (int bps = 1; bps <= 2; ++bps) {
            if (threads * bps > 2048) continue;
            const int it2 = 20000;
            k_probe<<<sms * bps, threads>>>(4, 0, 0, 200, d_out, d_sink); cudaDeviceSynchronize();
            cudaEventRecord(e0);
            k_probe<<<sms * bps, threads>>>(4, 0, 0, it2, d_out, d_sink);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
            long long h[16]; cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
            const double dm = (double)sms * bps * (threads / 32 < 16 ? threads / 32 : 16) * 32.0 * it2;     // warps 16.. of a block idle (slot >= 4)
            printf("  %4d threads x %d blocks/SM: %.2f TFLOP/s  (%.3f ms; warp 0: %.1f clk per 32-DMMA item -> %.0f MHz effective)\n", threads, bps, dm * 512.0 / (ms * 1e-3) / 1e12, ms,
                   (double)h[0] / it2, (double)h[0] / (ms * 1e-3) / 1e6);
        }
    return 0;
}
