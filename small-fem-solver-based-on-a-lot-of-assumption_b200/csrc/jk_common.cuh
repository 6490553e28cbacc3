// jk_common.cuh -- shared constants, layouts and small device helpers (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace jk {

// ----------------------------------------------------------------------------------------------
// Layout constants
// ----------------------------------------------------------------------------------------------
constexpr int NB   = 64;    // solver tile edge (DOFs); K_ff and L live as NB x NB row-major tiles
constexpr int SLAB = 32;    // right-hand sides (phases) per solver slab
constexpr int PH_TPB = 128; // threads per block of the phase-parallel kernels (one thread = one phase)
#ifndef JK_MCHUNK
#define JK_MCHUNK 32
#endif
constexpr int MCHUNK = JK_MCHUNK;  // members per block of the Morison / member-post kernels (A/B switch)
constexpr int NCHUNK = 64;  // nodes per block of the node-post kernel

// member constant row (structure of the row is fixed; one row per member, AoS so a block can stage
// its chunk with one coalesced copy)
constexpr int MC_STRIDE = 48;
constexpr int MC_L = 0;       // length, m
constexpr int MC_E = 1;       // unit vector e[3]
constexpr int MC_R = 4;       // frame rows lx, ly, lz (9)
constexpr int MC_ALPHA = 13;  // E*Ax/L
constexpr int MC_BZ = 14;     // E*Iz/((1+Phi_y) L^3)
constexpr int MC_BY = 15;     // E*Iy/((1+Phi_z) L^3)
constexpr int MC_TORS = 16;   // G*Ix/L
constexpr int MC_PHIY = 17;
constexpr int MC_PHIZ = 18;
constexpr int MC_LMM = 19;    // length, mm
constexpr int MC_D = 20;      // outer diameter, m
constexpr int MC_ACROSS = 21; // pi D^2 / 4, m^2
constexpr int MC_AX = 22;
constexpr int MC_IY = 23;
constexpr int MC_IZ = 24;
constexpr int MC_IX = 25;
constexpr int MC_AY = 26;
constexpr int MC_AZ = 27;
constexpr int MC_RO = 28;     // outer radius, mm
// distinct entries of the local stiffness (GUI.py:406-421), products taken in the reference's order
constexpr int MC_K12Z = 29;   // 12 bz
constexpr int MC_K6ZL = 30;   // 6 bz L
constexpr int MC_K4Z = 31;    // (4 + Phi_y) bz L^2
constexpr int MC_K2Z = 32;    // (2 - Phi_y) bz L^2
constexpr int MC_K12Y = 33;   // 12 by
constexpr int MC_K6YL = 34;   // 6 by L
constexpr int MC_K4Y = 35;    // (4 + Phi_z) by L^2
constexpr int MC_K2Y = 36;    // (2 - Phi_z) by L^2
// reciprocals used by the stress evaluation
constexpr int MC_IAX = 37;
constexpr int MC_IIY = 38;
constexpr int MC_IIZ = 39;
constexpr int MC_IIX = 40;
constexpr int MC_IAY = 41;
constexpr int MC_IAZ = 42;

// Gauss-point table row (Airy): cos(k x_w), sin(k x_w), Cu, Cw, z, pad  (three 16-byte shared-memory loads)
constexpr int GP_STRIDE = 6;

// ----------------------------------------------------------------------------------------------
// Addressing helpers
// ----------------------------------------------------------------------------------------------
// Tile (I, J), J <= I, I - J <= bw of the lower block band.  Tiles of one tile-row are contiguous.
__host__ __device__ __forceinline__ size_t tile_off(int I, int J, int bw) {
    return ((size_t)I * (size_t)(bw + 1) + (size_t)(I - J)) * (size_t)(NB * NB);
}
// Right-hand sides / solutions: slab-packed [slab][row][SLAB]; phase p lives in slab p / SLAB.
__host__ __device__ __forceinline__ size_t rhs_off(int row, int p, int n_pad) {
    return ((size_t)(p / SLAB) * (size_t)n_pad + (size_t)row) * SLAB + (size_t)(p % SLAB);
}

struct WaveAiry {
    double a, k, omega, d, Uc, dt, inv_dt;
    double cos_w, sin_w;        // wave heading (math angle)
    double uc_cos_c, uc_sin_c;  // current velocity components
};

// mma.sync m8n8k4 f64: A row-major fragment (lane holds A[l/4][l%4]), B "col" fragment
// (lane holds B[l%4][l/4]), C/D lane holds C[l/4][2*(l%4) + {0,1}].  SASS: DMMA.8x8x4.
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" :: "r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" :: "n"(N)); }

}  // namespace jk
