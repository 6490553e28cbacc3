// Micro-benchmark of the factorisation's critical chain on one CTA (B200): clocks of potrf8_warp, potrf64_smem, the
// panel TRSM and the SYRK of k_band_chol_cluster's CTA 0, with a correctness check against a host Cholesky.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/chol_probe tools/chol_probe.cu
#include <cstdio>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include "../small-fem-solver-based-on-a-lot-of-assumption_b200/csrc/jk_chol_cluster.cuh"
using namespace jk;

__global__ void __launch_bounds__(CHOL_THREADS, 1) k_probe(const double* __restrict__ Ain, const double* __restrict__ Pin, double* __restrict__ Lout,
                                                           double* __restrict__ Dout, double* __restrict__ Pout, double* __restrict__ Sout,
                                                           long long* __restrict__ clk, int* info, int variant) {
    extern __shared__ __align__(16) double smem[];
    double* As = smem; double* Bs = As + NB * LS_LD; double* Cs = Bs + NB * LS_LD; double* Di = Cs + NB * LS_LD;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, fr = lane >> 2, fk = lane & 3;
    for (int q = tid; q < 8 * DI_BLK; q += CHOL_THREADS) Di[q] = 0.0;
    for (int q = tid; q < NB * NB; q += CHOL_THREADS) { Cs[(q / NB) * LS_LD + q % NB] = Ain[q]; As[(q / NB) * LS_LD + q % NB] = Pin[q]; Bs[(q / NB) * LS_LD + q % NB] = Ain[q]; }
    __syncthreads();
    // pieces of the chain on scratch copies (As holds P: use Bs = copy of A)
    if (warp == 0) {
        long long a = clock64();
        potrf8_warp<true, true>(Bs, Di, 0, info, 0, lane);      // the <true,false> / <false,true> pair meets at a named barrier: never call one half alone
        __syncwarp();
        long long b = clock64();
        // rows-below solve of one 8-row tile + store
        double x0 = 0.0, x1 = 0.0;
#pragma unroll
        for (int k4 = 0; k4 < 2; ++k4) dmma(x0, x1, Bs[(8 + fr) * LS_LD + 4 * k4 + fk], Di[fr * DI_LD + 4 * k4 + fk]);
        __syncwarp();
        Bs[(8 + fr) * LS_LD + 2 * fk] = x0; Bs[(8 + fr) * LS_LD + 2 * fk + 1] = x1;
        __syncwarp();
        long long c = clock64();
        double* cp = Bs + (8 + fr) * LS_LD + 8 + 2 * fk;
        double c0v = cp[0], c1v = cp[1];
#pragma unroll
        for (int k4 = 0; k4 < 2; ++k4) dmma(c0v, c1v, -Bs[(8 + fr) * LS_LD + 4 * k4 + fk], Bs[(8 + fr) * LS_LD + 4 * k4 + fk]);
        cp[0] = c0v; cp[1] = c1v;
        __syncwarp();
        long long d = clock64();
        // latency of a dependent DFMA / DMMA chain (64 deep)
        double z = Bs[lane], y0 = 0.0, y1 = 0.0;
#pragma unroll
        for (int i = 0; i < 64; ++i) z = fma(z, 1.0000001, 1e-9);
        long long e = clock64();
#pragma unroll
        for (int i = 0; i < 64; ++i) dmma(y0, y1, z, z);
        long long f = clock64();
        double r = z;
#pragma unroll
        for (int i = 0; i < 16; ++i) r = fast_rcp(r + 1.5);
        long long g = clock64();
        double w[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) w[i] = 0.0;
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) dmma(w[2 * j], w[2 * j + 1], z, r);
        long long h = clock64();
        if (lane == 0) { clk[9] = h - g; double sacc = 0; for (int i = 0; i < 16; ++i) sacc += w[i]; Sout[1] = sacc; }
        if (lane == 0) { clk[3] = b - a; clk[4] = c - b; clk[5] = d - c; clk[6] = (e - d); clk[7] = (f - e); Sout[0] = y0 + y1 + r; }
        if (lane == 0) clk[8] = g - f;
    }
    __syncthreads();
    for (int q = tid; q < 8 * DI_BLK; q += CHOL_THREADS) Di[q] = 0.0;
    for (int q = tid; q < NB * NB; q += CHOL_THREADS) Bs[(q / NB) * LS_LD + q % NB] = Ain[q];
    __syncthreads();
    long long t0 = clock64();
    if (variant == 0) potrf64_smem(Cs, Di, info, 0); else potrf64_pipelined(Cs, Di, info, 0);
    __syncthreads();
    long long t1 = clock64();
    // panel TRSM: As <- As * L^{-T}
    trsm64_warp(As, Cs, Di);
    __syncthreads();
    long long t2 = clock64();
    // SYRK as in CTA 0: Bs -= As As^T (lower 8x8 tiles)
    syrk64_lower(As, Bs);
    __syncthreads();
    long long t3 = clock64();
    if (tid == 0) { clk[0] = t1 - t0; clk[1] = t2 - t1; clk[2] = t3 - t2; }
    // DMMA throughput of one SM with 8 warps (2 per scheduler): 256 DMMAs per warp, 8 / 2 / 1 independent accumulator chains
    for (int chains = 8; chains >= 1; chains /= 2 * (chains > 2 ? 2 : 1)) {
        double w[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) w[i] = 0.0;
        const double za = As[tid], zb = Bs[tid];
        __syncthreads();
        long long u0 = clock64();
        if (chains == 8) {
#pragma unroll 4
            for (int i = 0; i < 32; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) dmma(w[2 * j], w[2 * j + 1], za, zb);
        } else if (chains == 2) {
#pragma unroll 8
            for (int i = 0; i < 128; ++i) { dmma(w[0], w[1], za, zb); dmma(w[2], w[3], za, zb); }
        } else {
#pragma unroll 8
            for (int i = 0; i < 256; ++i) dmma(w[0], w[1], za, zb);
        }
        double sacc = 0;
#pragma unroll
        for (int i = 0; i < 16; ++i) sacc += w[i];
        if (sacc == 12345.678) Sout[2] = sacc;
        __syncthreads();
        long long u1 = clock64();
        if (tid == 0) clk[10 + (chains == 8 ? 0 : chains == 2 ? 1 : 2)] = u1 - u0;
    }
    for (int q = tid; q < NB * NB; q += CHOL_THREADS) { Lout[q] = Cs[(q / NB) * LS_LD + q % NB]; Pout[q] = As[(q / NB) * LS_LD + q % NB]; Sout[q] = Bs[(q / NB) * LS_LD + q % NB]; }
    for (int q = tid; q < 512; q += CHOL_THREADS) { int b = q >> 6, r = (q >> 3) & 7, c = q & 7; Dout[q] = Di[b * DI_BLK + r * DI_LD + c]; }
}

int main() {
    const int n = NB;
    std::vector<double> A(n * n), P(n * n), L(n * n, 0.0);
    srand(7);
    std::vector<double> G(n * n);
    for (auto& v : G) v = rand() / (double)RAND_MAX - 0.5;
    for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) { double s = 0; for (int k = 0; k < n; ++k) s += G[i * n + k] * G[j * n + k]; A[i * n + j] = s + (i == j ? 8.0 : 0.0); }
    for (auto& v : P) v = rand() / (double)RAND_MAX - 0.5;
    for (int j = 0; j < n; ++j) {          // host Cholesky
        double d = A[j * n + j]; for (int k = 0; k < j; ++k) d -= L[j * n + k] * L[j * n + k]; L[j * n + j] = sqrt(d);
        for (int i = j + 1; i < n; ++i) { double s = A[i * n + j]; for (int k = 0; k < j; ++k) s -= L[i * n + k] * L[j * n + k]; L[i * n + j] = s / L[j * n + j]; }
    }
    std::vector<double> X(n * n);          // host X = P L^{-T}
    for (int r = 0; r < n; ++r) for (int c = 0; c < n; ++c) { double s = P[r * n + c]; for (int k = 0; k < c; ++k) s -= X[r * n + k] * L[c * n + k]; X[r * n + c] = s / L[c * n + c]; }
    double *dA, *dP, *dL, *dD, *dPo, *dS; long long* dclk; int* dinfo;
    cudaMalloc(&dA, n * n * 8); cudaMalloc(&dP, n * n * 8); cudaMalloc(&dL, n * n * 8); cudaMalloc(&dD, 512 * 8); cudaMalloc(&dPo, n * n * 8); cudaMalloc(&dS, n * n * 8);
    cudaMalloc(&dclk, 256); cudaMalloc(&dinfo, 4); cudaMemset(dinfo, 0, 4);
    cudaMemcpy(dA, A.data(), n * n * 8, cudaMemcpyHostToDevice); cudaMemcpy(dP, P.data(), n * n * 8, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CHOL_CLUSTER_SMEM);
    for (int variant = 0; variant < 2; ++variant) {
        long long best[3] = {1LL << 60, 1LL << 60, 1LL << 60};
        for (int rep = 0; rep < 5; ++rep) {
            k_probe<<<1, CHOL_THREADS, CHOL_CLUSTER_SMEM>>>(dA, dP, dL, dD, dPo, dS, dclk, dinfo, variant);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("variant %d: CUDA error %s\n", variant, cudaGetErrorString(e)); return 1; }
            long long c[3]; cudaMemcpy(c, dclk, 24, cudaMemcpyDeviceToHost);
            for (int i = 0; i < 3; ++i) best[i] = c[i] < best[i] ? c[i] : best[i];
        }
        std::vector<double> Lg(n * n), Xg(n * n), Sg(n * n);
        cudaMemcpy(Lg.data(), dL, n * n * 8, cudaMemcpyDeviceToHost); cudaMemcpy(Xg.data(), dPo, n * n * 8, cudaMemcpyDeviceToHost); cudaMemcpy(Sg.data(), dS, n * n * 8, cudaMemcpyDeviceToHost);
        double eL = 0, eX = 0, eS = 0;
        for (int i = 0; i < n; ++i) for (int j = 0; j <= i; ++j) eL = fmax(eL, fabs(Lg[i * n + j] - L[i * n + j]));
        for (int i = 0; i < n * n; ++i) eX = fmax(eX, fabs(Xg[i] - X[i]));
        for (int i = 0; i < n; ++i) for (int j = 0; j <= i; ++j) { double s = A[i * n + j]; for (int k = 0; k < n; ++k) s -= X[i * n + k] * X[j * n + k]; eS = fmax(eS, fabs(Sg[i * n + j] - s)); }
        int info; cudaMemcpy(&info, dinfo, 4, cudaMemcpyDeviceToHost);
        { long long c[13]; cudaMemcpy(c, dclk, 104, cudaMemcpyDeviceToHost); printf("  8 warps x 256 DMMA on one SM: 8 chains %lld | 2 chains %lld | 1 chain %lld clk\n", c[10], c[11], c[12]);
          printf("  pieces (last run): potrf8 %lld | solve tile %lld | update tile %lld | 64 dep DFMA %lld | 64 dep DMMA %lld | 16 dep fast_rcp %lld\n", c[3], c[4], c[5], c[6], c[7], c[8]); }
        printf("variant %d: potrf64 %lld clk | trsm64 %lld | syrk64 %lld | err L %.2e X %.2e S %.2e info %d\n", variant, best[0], best[1], best[2], eL, eX, eS, info);
    }
    return 0;
}
