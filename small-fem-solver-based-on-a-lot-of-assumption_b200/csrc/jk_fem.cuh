// jk_fem.cuh -- element construction, deterministic assembly and post-processing kernels.
//
// Reference behaviour restated here (file:line = /root/reference/JacketAnalysisGUI_v2.py):
//   k_member_setup      BeamElement3D.__init__/_compute_transformation_matrix/_compute_local_stiffness  361-422
//   k_assemble_blocks   FEMSolver._assemble_global_stiffness + apply_boundary_conditions                457-479
//   k_member_post       BeamElement3D.get_internal_forces + FEMSolver.get_member_internal_forces
//                       + TubularSection.calc_stress_at_point                                           424-432, 504-533, 147-160
//   k_reactions         FEMSolver.get_reactions                                                         492-502
//   k_node_post         max nodal translation of run_analysis                                           2035-2040
#pragma once
#include "jk_common.cuh"

#ifndef JK_POST_PACKED
#define JK_POST_PACKED 1      // per-member constants of k_member_post in one packed shared-memory row read with 16-byte loads
#endif
#ifndef JK_POST_TPB
#define JK_POST_TPB 128        // threads (= phases) per block of k_member_post (A/B switch: longer contiguous write runs)
#endif
#ifndef JK_POST_PREFETCH
#define JK_POST_PREFETCH 0   // 1: next member's displacements in flight during this member's arithmetic (96 registers; measured 4 % slower)
#endif

namespace jk {

// ----------------------------------------------------------------------------------------------
// K2a: one thread per member -> constant row, global element matrix Ke (symmetric by construction)
// ----------------------------------------------------------------------------------------------
__global__ void k_member_setup(int M, const double* __restrict__ xyz, const int* __restrict__ conn,
                               const int* __restrict__ sec, const double* __restrict__ secp, int nprop,
                               double E, double G, double* __restrict__ mc, double* __restrict__ Ke,
                               double* __restrict__ Kl_out) {
    int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    int a = conn[2 * m], b = conn[2 * m + 1];
    double dx = xyz[3 * b] - xyz[3 * a], dy = xyz[3 * b + 1] - xyz[3 * a + 1], dz = xyz[3 * b + 2] - xyz[3 * a + 2];
    double L = sqrt(dx * dx + dy * dy + dz * dz);
    double lx[3] = {dx / L, dy / L, dz / L}, ly[3], lz[3];
    if (fabs(lx[2]) > 0.999) {   // near-vertical member: ly = z x lx (GUI.py:374-378)
        ly[0] = -lx[1]; ly[1] = lx[0]; ly[2] = 0.0;
        double n = sqrt(ly[0] * ly[0] + ly[1] * ly[1] + ly[2] * ly[2]);
        if (n > 1e-10) { ly[0] /= n; ly[1] /= n; ly[2] /= n; } else { ly[0] = 0.0; ly[1] = 1.0; ly[2] = 0.0; }
        lz[0] = lx[1] * ly[2] - lx[2] * ly[1];
        lz[1] = lx[2] * ly[0] - lx[0] * ly[2];
        lz[2] = lx[0] * ly[1] - lx[1] * ly[0];
    } else {                      // lz = lx x z, ly = lz x lx (GUI.py:380-382)
        lz[0] = lx[1]; lz[1] = -lx[0]; lz[2] = 0.0;
        double n = sqrt(lz[0] * lz[0] + lz[1] * lz[1] + lz[2] * lz[2]);
        lz[0] /= n; lz[1] /= n; lz[2] /= n;
        ly[0] = lz[1] * lx[2] - lz[2] * lx[1];
        ly[1] = lz[2] * lx[0] - lz[0] * lx[2];
        ly[2] = lz[0] * lx[1] - lz[1] * lx[0];
    }
    const double* sp = secp + (size_t)sec[m] * nprop;
    double Do = sp[0], Ax = sp[1], Iy = sp[2], Iz = sp[3], Ix = sp[4], Ay = sp[5], Az = sp[6], Ro = sp[7];
    double Lmm = L * 1000.0;
    double L2 = Lmm * Lmm, L3 = L2 * Lmm;
    double phy = 0.0, phz = 0.0;
    if (Ay > 0.0 && Az > 0.0) {   // include_shear (GUI.py:394-396)
        phy = 12.0 * E * Iz / (G * Az * L2);
        phz = 12.0 * E * Iy / (G * Ay * L2);
    }
    double alpha = E * Ax / Lmm;
    double bz = E * Iz / ((1.0 + phy) * L3);
    double by = E * Iy / ((1.0 + phz) * L3);
    double tors = G * Ix / Lmm;
    double k12z = 12.0 * bz, k6zL = 6.0 * bz * Lmm, k4z = (4.0 + phy) * bz * L2, k2z = (2.0 - phy) * bz * L2;
    double k12y = 12.0 * by, k6yL = 6.0 * by * Lmm, k4y = (4.0 + phz) * by * L2, k2y = (2.0 - phz) * by * L2;

    double* c = mc + (size_t)m * MC_STRIDE;
    c[MC_L] = L;
    for (int i = 0; i < 3; ++i) { c[MC_E + i] = lx[i]; c[MC_R + i] = lx[i]; c[MC_R + 3 + i] = ly[i]; c[MC_R + 6 + i] = lz[i]; }
    c[MC_ALPHA] = alpha; c[MC_BZ] = bz; c[MC_BY] = by; c[MC_TORS] = tors; c[MC_PHIY] = phy; c[MC_PHIZ] = phz;
    c[MC_LMM] = Lmm;
    double Dm = Do / 1000.0;
    c[MC_D] = Dm; c[MC_ACROSS] = 3.141592653589793 * (Dm * Dm) / 4.0;
    c[MC_AX] = Ax; c[MC_IY] = Iy; c[MC_IZ] = Iz; c[MC_IX] = Ix; c[MC_AY] = Ay; c[MC_AZ] = Az; c[MC_RO] = Ro;
    c[MC_K12Z] = k12z; c[MC_K6ZL] = k6zL; c[MC_K4Z] = k4z; c[MC_K2Z] = k2z;
    c[MC_K12Y] = k12y; c[MC_K6YL] = k6yL; c[MC_K4Y] = k4y; c[MC_K2Y] = k2y;
    c[MC_IAX] = 1.0 / Ax; c[MC_IIY] = Iy > 0.0 ? 1.0 / Iy : 0.0; c[MC_IIZ] = Iz > 0.0 ? 1.0 / Iz : 0.0;
    c[MC_IIX] = Ix > 0.0 ? 1.0 / Ix : 0.0; c[MC_IAY] = Ay > 0.0 ? 1.0 / Ay : 0.0; c[MC_IAZ] = Az > 0.0 ? 1.0 / Az : 0.0;
    for (int i = 43; i < MC_STRIDE; ++i) c[i] = 0.0;

    // local stiffness, entries as listed at GUI.py:406-421
    double Kl[12][12];
    for (int i = 0; i < 12; ++i) for (int j = 0; j < 12; ++j) Kl[i][j] = 0.0;
    Kl[0][0] = Kl[6][6] = alpha;   Kl[0][6] = Kl[6][0] = -alpha;
    Kl[1][1] = Kl[7][7] = k12z;    Kl[1][7] = Kl[7][1] = -k12z;
    Kl[1][5] = Kl[5][1] = Kl[1][11] = Kl[11][1] = k6zL;
    Kl[7][5] = Kl[5][7] = Kl[7][11] = Kl[11][7] = -k6zL;
    Kl[5][5] = Kl[11][11] = k4z;   Kl[5][11] = Kl[11][5] = k2z;
    Kl[2][2] = Kl[8][8] = k12y;    Kl[2][8] = Kl[8][2] = -k12y;
    Kl[2][4] = Kl[4][2] = Kl[2][10] = Kl[10][2] = -k6yL;
    Kl[8][4] = Kl[4][8] = Kl[8][10] = Kl[10][8] = k6yL;
    Kl[4][4] = Kl[10][10] = k4y;   Kl[4][10] = Kl[10][4] = k2y;
    Kl[3][3] = Kl[9][9] = tors;    Kl[3][9] = Kl[9][3] = -tors;
    if (Kl_out) for (int i = 0; i < 12; ++i) for (int j = 0; j < 12; ++j) Kl_out[(size_t)m * 144 + i * 12 + j] = Kl[i][j];

    // Ke = T^T (Kl T), T = blkdiag(R,R,R,R), R rows = lx, ly, lz.  Lower triangle computed, mirrored.
    double Rm[3][3] = {{lx[0], lx[1], lx[2]}, {ly[0], ly[1], ly[2]}, {lz[0], lz[1], lz[2]}};
    double W[12][12];
    for (int i = 0; i < 12; ++i)
        for (int bb = 0; bb < 4; ++bb)
            for (int cc = 0; cc < 3; ++cc) {
                double s = 0.0;
                for (int j = 0; j < 3; ++j) s = fma(Kl[i][3 * bb + j], Rm[j][cc], s);
                W[i][3 * bb + cc] = s;
            }
    double* ke = Ke + (size_t)m * 144;
    for (int aa = 0; aa < 4; ++aa)
        for (int r = 0; r < 3; ++r)
            for (int col = 0; col <= 3 * aa + r; ++col) {
                double s = 0.0;
                for (int i = 0; i < 3; ++i) s = fma(Rm[i][r], W[3 * aa + i][col], s);
                ke[(3 * aa + r) * 12 + col] = s;
                ke[col * 12 + (3 * aa + r)] = s;
            }
}

// ----------------------------------------------------------------------------------------------
// K2b: deterministic segmented scatter-add.  One CUDA block (64 threads, 36 active) per 6x6 node
// block of K_ff; its contributions (member, quadrant) are summed in member order.
// ----------------------------------------------------------------------------------------------
struct KBlock { int row_slot, col_slot, start, count; };

__global__ void k_assemble_blocks(int nblocks, const KBlock* __restrict__ blocks, const int2* __restrict__ contrib,
                                  const double* __restrict__ Ke, double* __restrict__ tiles, int bw) {
    int bi = blockIdx.x;
    if (bi >= nblocks) return;
    int tid = threadIdx.x;
    if (tid >= 36) return;
    int r = tid / 6, c = tid % 6;
    KBlock kb = blocks[bi];
    double s = 0.0;
    for (int i = 0; i < kb.count; ++i) {
        int2 ct = contrib[kb.start + i];          // x = member, y = (row quadrant << 1) | col quadrant
        int qr = (ct.y >> 1) & 1, qc = ct.y & 1;
        s += Ke[(size_t)ct.x * 144 + (6 * qr + r) * 12 + (6 * qc + c)];
    }
    int Rg = 6 * kb.row_slot + r, Cg = 6 * kb.col_slot + c;
    if (Rg < Cg) return;                           // only the lower triangle is stored
    int I = Rg / NB, J = Cg / NB;
    tiles[tile_off(I, J, bw) + (size_t)(Rg % NB) * NB + (Cg % NB)] = s;
}

// identity on the padded tail of the last diagonal tile so the factor is defined there
__global__ void k_pad_identity(int n_free, int n_pad, double* __restrict__ tiles, int bw) {
    int r = n_free + blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_pad) return;
    int I = r / NB;
    tiles[tile_off(I, I, bw) + (size_t)(r % NB) * NB + (r % NB)] = 1.0;
}

// dense K in the reference's DOF order (FEMSolver.K_global accessor / parity tests only).
// One thread per (member, entry); summation order is not fixed here (atomics) -- the solver never reads this.
__global__ void k_dense_K(int M, const int* __restrict__ conn, const double* __restrict__ Ke, double* __restrict__ K, int n_dof) {
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)M * 144) return;
    int m = (int)(idx / 144), e = (int)(idx % 144), ii = e / 12, jj = e % 12;
    int di = 6 * conn[2 * m + (ii >= 6)] + (ii % 6), dj = 6 * conn[2 * m + (jj >= 6)] + (jj % 6);
    atomicAdd(&K[(size_t)di * n_dof + dj], Ke[idx]);
}

// ----------------------------------------------------------------------------------------------
// element end forces from element displacements (local axes)
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void element_local_forces(const double* __restrict__ c, const double* ue, double* Fl) {
    double ul[12];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int i = 0; i < 3; ++i)
            ul[3 * a + i] = fma(c[MC_R + 3 * i + 2], ue[3 * a + 2], fma(c[MC_R + 3 * i + 1], ue[3 * a + 1], c[MC_R + 3 * i] * ue[3 * a]));
    const double al = c[MC_ALPHA], to = c[MC_TORS];
    const double k12z = c[MC_K12Z], k6zL = c[MC_K6ZL], k4z = c[MC_K4Z], k2z = c[MC_K2Z];
    const double k12y = c[MC_K12Y], k6yL = c[MC_K6YL], k4y = c[MC_K4Y], k2y = c[MC_K2Y];
    Fl[0] = fma(-al, ul[6], al * ul[0]);
    Fl[1] = fma(k6zL, ul[11], fma(-k12z, ul[7], fma(k6zL, ul[5], k12z * ul[1])));
    Fl[2] = fma(-k6yL, ul[10], fma(-k12y, ul[8], fma(-k6yL, ul[4], k12y * ul[2])));
    Fl[3] = fma(-to, ul[9], to * ul[3]);
    Fl[4] = fma(k2y, ul[10], fma(k6yL, ul[8], fma(k4y, ul[4], -k6yL * ul[2])));
    Fl[5] = fma(k2z, ul[11], fma(-k6zL, ul[7], fma(k4z, ul[5], k6zL * ul[1])));
    Fl[6] = -Fl[0]; Fl[7] = -Fl[1]; Fl[8] = -Fl[2]; Fl[9] = -Fl[3];   // bitwise what the mirrored products of GUI.py:406-421 give
    Fl[10] = fma(k4y, ul[10], fma(k6yL, ul[8], fma(k2y, ul[4], -k6yL * ul[2])));
    Fl[11] = fma(k4z, ul[11], fma(-k6zL, ul[7], fma(k2z, ul[5], k6zL * ul[1])));
}

// cos/sin of the 8 stress points 0,45,...,315 deg (GUI.py:139-145); filled by the host with libm
struct StressPts { double cs[8], sn[8]; };

// per-member stress-point coefficients: pt[3*i + {0,1,2}] = y_i/Iz, z_i/Iy, R_i/Ix for the 8 points of
// TubularSection.get_stress_points (GUI.py:139-145)
__device__ __forceinline__ void stress_point_coeffs(const double* __restrict__ c, const StressPts& sp, int i, double* pt) {
    double y = c[MC_RO] * sp.cs[i], z = c[MC_RO] * sp.sn[i];
    pt[0] = y * c[MC_IIZ];
    pt[1] = z * c[MC_IIY];
    pt[2] = sqrt(fma(y, y, z * z)) * c[MC_IIX];
}

// 7 result fields of one member from its 12 local end forces (GUI.py:514-532).
// Two exact identities keep the FP64 count down: (1) axial force, shears and torque of node 2 are the bitwise negatives
// of node 1's (same products, opposite signs, GUI.py:406-421), so max(|n1|, |n2|) = |n1| for those rows; (2) the 8
// stress points (GUI.py:139-145) lie on one circle in antipodal pairs, so tau is the same at all of them and
// max_i sigma_i^2 = (|Fx/A| + max_{i<4} |Mz y_i/Iz + My z_i/Iy|)^2 -- 4 bending terms and one square root instead of 8
// full evaluations (differences to the literal loop: rounding of cos/sin(theta + 180 deg), 1e-16 relative).
__device__ __forceinline__ void member_row(const double* __restrict__ c, const double* __restrict__ pt, const double* Fl,
                                           double inv_fy, double* row) {
    // node-1 forces carry the sign flip of GUI.py:428-429 (irrelevant under the squares / absolute values below)
    const double sFx = fabs(Fl[0] * c[MC_IAX]);
    const double tFy = Fl[1] * c[MC_IAY], tFz = Fl[2] * c[MC_IAZ];
    const double tM = Fl[3] * pt[2];
    const double tau2 = fma(tM, tM, fma(tFy, tFy, tFz * tFz));
    double bmax = 0.0;
#pragma unroll
    for (int i = 0; i < 4; ++i) bmax = fmax(bmax, fabs(fma(Fl[5], pt[3 * i], Fl[4] * pt[3 * i + 1])));
    const double sig = sFx + bmax;
    const double vm = sqrt(fma(3.0, tau2, sig * sig));
    row[0] = fabs(Fl[0]) * 1e-3;
    row[1] = fabs(Fl[1]) * 1e-3;
    row[2] = fabs(Fl[2]) * 1e-3;
    row[3] = fmax(fabs(Fl[4]), fabs(Fl[10])) * 1e-6;
    row[4] = fmax(fabs(Fl[5]), fabs(Fl[11])) * 1e-6;
    row[5] = vm;
    row[6] = vm * inv_fy;
}

// Packed per-member constants of k_member_post (one 256-byte shared-memory row, read with sixteen 16-byte loads):
//   0-8 R | 9 alpha | 10 tors | 11 k12z 12 k6zL 13 k4z 14 k2z | 15 k12y 16 k6yL 17 k4y 18 k2y | 19 1/A 20 1/Ay 21 1/Az |
//   22 R_o/Ix | 23 pad | 24+2i, 25+2i: y_i/Iz, z_i/Iy of stress points i = 0..3
// The two functions below are element_local_forces / member_row with these indices: same operations in the same order,
// so the stored rows stay bit-identical to k_member_post_single's.
constexpr int PK_STRIDE = 32;
__device__ __forceinline__ void pack_member_consts(const double* __restrict__ c, const StressPts& sp, double* __restrict__ pk) {
#pragma unroll
    for (int i = 0; i < 9; ++i) pk[i] = c[MC_R + i];
    pk[9] = c[MC_ALPHA]; pk[10] = c[MC_TORS];
    pk[11] = c[MC_K12Z]; pk[12] = c[MC_K6ZL]; pk[13] = c[MC_K4Z]; pk[14] = c[MC_K2Z];
    pk[15] = c[MC_K12Y]; pk[16] = c[MC_K6YL]; pk[17] = c[MC_K4Y]; pk[18] = c[MC_K2Y];
    pk[19] = c[MC_IAX]; pk[20] = c[MC_IAY]; pk[21] = c[MC_IAZ];
    double pt[3];
    stress_point_coeffs(c, sp, 0, pt);
    pk[22] = pt[2]; pk[23] = 0.0; pk[24] = pt[0]; pk[25] = pt[1];
#pragma unroll
    for (int i = 1; i < 4; ++i) { stress_point_coeffs(c, sp, i, pt); pk[24 + 2 * i] = pt[0]; pk[25 + 2 * i] = pt[1]; }
}
__device__ __forceinline__ void element_local_forces_pk(const double* pk, const double* ue, double* Fl) {
    double ul[12];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int i = 0; i < 3; ++i)
            ul[3 * a + i] = fma(pk[3 * i + 2], ue[3 * a + 2], fma(pk[3 * i + 1], ue[3 * a + 1], pk[3 * i] * ue[3 * a]));
    const double al = pk[9], to = pk[10];
    const double k12z = pk[11], k6zL = pk[12], k4z = pk[13], k2z = pk[14];
    const double k12y = pk[15], k6yL = pk[16], k4y = pk[17], k2y = pk[18];
    Fl[0] = fma(-al, ul[6], al * ul[0]);
    Fl[1] = fma(k6zL, ul[11], fma(-k12z, ul[7], fma(k6zL, ul[5], k12z * ul[1])));
    Fl[2] = fma(-k6yL, ul[10], fma(-k12y, ul[8], fma(-k6yL, ul[4], k12y * ul[2])));
    Fl[3] = fma(-to, ul[9], to * ul[3]);
    Fl[4] = fma(k2y, ul[10], fma(k6yL, ul[8], fma(k4y, ul[4], -k6yL * ul[2])));
    Fl[5] = fma(k2z, ul[11], fma(-k6zL, ul[7], fma(k4z, ul[5], k6zL * ul[1])));
    Fl[6] = -Fl[0]; Fl[7] = -Fl[1]; Fl[8] = -Fl[2]; Fl[9] = -Fl[3];
    Fl[10] = fma(k4y, ul[10], fma(k6yL, ul[8], fma(k2y, ul[4], -k6yL * ul[2])));
    Fl[11] = fma(k4z, ul[11], fma(-k6zL, ul[7], fma(k2z, ul[5], k6zL * ul[1])));
}
__device__ __forceinline__ void member_row_pk(const double* pk, const double* Fl, double inv_fy, double* row) {
    const double sFx = fabs(Fl[0] * pk[19]);
    const double tFy = Fl[1] * pk[20], tFz = Fl[2] * pk[21];
    const double tM = Fl[3] * pk[22];
    const double tau2 = fma(tM, tM, fma(tFy, tFy, tFz * tFz));
    double bmax = 0.0;
#pragma unroll
    for (int i = 0; i < 4; ++i) bmax = fmax(bmax, fabs(fma(Fl[5], pk[24 + 2 * i], Fl[4] * pk[25 + 2 * i])));
    const double sig = sFx + bmax;
    const double vm = sqrt(fma(3.0, tau2, sig * sig));
    row[0] = fabs(Fl[0]) * 1e-3;
    row[1] = fabs(Fl[1]) * 1e-3;
    row[2] = fabs(Fl[2]) * 1e-3;
    row[3] = fmax(fabs(Fl[4]), fabs(Fl[10])) * 1e-6;
    row[4] = fmax(fabs(Fl[5]), fabs(Fl[11])) * 1e-6;
    row[5] = vm;
    row[6] = vm * inv_fy;
}

// packed per-member constants of the member post, once per assembly: pk[M][PK_STRIDE]
__global__ void k_pack_member_consts(int M, const double* __restrict__ mc, StressPts sp, double* __restrict__ pk) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m < M) pack_member_consts(mc + (size_t)m * MC_STRIDE, sp, pk + (size_t)m * PK_STRIDE);
}

__device__ __forceinline__ double load_u(const double* __restrict__ X, const int* __restrict__ node2slot,
                                         int node, int comp, int p, int n_pad) {
    int s = node2slot[node];
    return s >= 0 ? X[rhs_off(s + comp, p, n_pad)] : 0.0;   // node2slot holds the node's first ROW of the slab (or -1-f)
}

// ----------------------------------------------------------------------------------------------
// K5: block = 128 phases x MCHUNK members.  rows[m][7][ldP]; per-chunk running maxima.
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(JK_POST_TPB)
k_member_post(int M, int P, int ldP, int n_pad, const double* __restrict__ X, const int* __restrict__ node2slot,
              const int* __restrict__ conn, const double* __restrict__ mc, StressPts sp, double fy,
              double* __restrict__ rows, double* __restrict__ part_util, double* __restrict__ part_vm,
              int* __restrict__ part_mem, const int* __restrict__ chunk_list = nullptr /* nullable: member chunks of this launch */,
              const double* __restrict__ pk_all = nullptr /* [M][PK_STRIDE] from k_pack_member_consts */) {
    // 8.5 KB of shared memory per block (the packed rows come ready-made from k_pack_member_consts): two of these blocks fit
    // beside a sweep CTA, so the early member post runs on the SMs the backward sweep occupies, not only on the idle ones
    __shared__ int s_slot[MCHUNK * 2];
#if JK_POST_PACKED
    __shared__ __align__(16) double s_pk[MCHUNK * PK_STRIDE];
#else
    __shared__ double s_mc[MCHUNK * MC_STRIDE];
    __shared__ double s_pt[MCHUNK * 24];
#endif
    // chunk index is the FAST grid dimension: the blocks in flight share one or two 128-phase tiles, whose slice of
    // the solution (n x 128 doubles = 20 MB at c4) stays L2-resident while every member chunk re-reads its nodes
    int chunk = chunk_list ? chunk_list[blockIdx.x] : (int)blockIdx.x, m0 = chunk * MCHUNK;
    int nm = min(MCHUNK, M - m0);
    for (int i = threadIdx.x; i < nm * 2; i += blockDim.x) s_slot[i] = node2slot[conn[2 * m0 + i]];
#if JK_POST_PACKED
    for (int i = threadIdx.x; i < nm * PK_STRIDE; i += blockDim.x) s_pk[i] = pk_all[(size_t)m0 * PK_STRIDE + i];
#else
    for (int i = threadIdx.x; i < nm * MC_STRIDE; i += blockDim.x) s_mc[i] = mc[(size_t)m0 * MC_STRIDE + i];
    __syncthreads();
    for (int i = threadIdx.x; i < nm * 8; i += blockDim.x) stress_point_coeffs(s_mc + (i / 8) * MC_STRIDE, sp, i % 8, s_pt + 3 * i);
#endif
    __syncthreads();
    const double inv_fy = 1.0 / fy;
    int p = blockIdx.y * blockDim.x + threadIdx.x;
    if (p >= ldP) return;
    const size_t xbase = (size_t)(p / SLAB) * (size_t)n_pad * SLAB + (size_t)(p % SLAB);
    double best_u = -1.0, best_vm = 0.0; int best_m = 0;
    auto load_ue = [&](int mm, double* ue) {
        const int s1 = s_slot[2 * mm], s2 = s_slot[2 * mm + 1];
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            ue[k] = s1 >= 0 ? X[xbase + (size_t)(s1 + k) * SLAB] : 0.0;
            ue[6 + k] = s2 >= 0 ? X[xbase + (size_t)(s2 + k) * SLAB] : 0.0;
        }
    };
#if JK_POST_PREFETCH
    double un[12];
    load_ue(0, un);
#endif
    for (int mm = 0; mm < nm; ++mm) {
        double ue[12], Fl[12], row[7];
#if JK_POST_PREFETCH
        // the next member's displacements are requested before this member's arithmetic (L2 latency under ~150 FP64 ops)
#pragma unroll
        for (int k = 0; k < 12; ++k) ue[k] = un[k];
        if (mm + 1 < nm) load_ue(mm + 1, un);
#else
        load_ue(mm, ue);
#endif
#if JK_POST_PACKED
        double pk[PK_STRIDE];
#pragma unroll
        for (int q = 0; q < PK_STRIDE / 2; ++q) {
            const double2 v = *reinterpret_cast<const double2*>(s_pk + mm * PK_STRIDE + 2 * q);
            pk[2 * q] = v.x; pk[2 * q + 1] = v.y;
        }
        element_local_forces_pk(pk, ue, Fl);
        member_row_pk(pk, Fl, inv_fy, row);
#else
        const double* c = s_mc + mm * MC_STRIDE;
        element_local_forces(c, ue, Fl);
        member_row(c, s_pt + mm * 24, Fl, inv_fy, row);
#endif
        size_t o = ((size_t)(m0 + mm) * 7) * ldP + p;
#pragma unroll
        for (int k = 0; k < 7; ++k) __stcs(rows + o + (size_t)k * ldP, row[k]);      // write-once stream: keep the solution slab in L2
        if (row[6] > best_u) { best_u = row[6]; best_vm = row[5]; best_m = m0 + mm; }
    }
    size_t po = (size_t)chunk * ldP + p;
    part_util[po] = best_u; part_vm[po] = best_vm; part_mem[po] = best_m;
}

// single phase, compact outputs (jk_fetch_phase): rows[M*7], endf[M*12]
__global__ void k_member_post_single(int M, int p, int n_pad, const double* __restrict__ X, const int* __restrict__ node2slot,
                                     const int* __restrict__ conn, const double* __restrict__ mc, StressPts sp, double fy,
                                     double* __restrict__ rows, double* __restrict__ endf) {
    int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    const double* c = mc + (size_t)m * MC_STRIDE;
    double ue[12], Fl[12], row[7];
    for (int k = 0; k < 6; ++k) {
        ue[k] = load_u(X, node2slot, conn[2 * m], k, p, n_pad);
        ue[6 + k] = load_u(X, node2slot, conn[2 * m + 1], k, p, n_pad);
    }
    element_local_forces(c, ue, Fl);
    double pt[24];
    for (int i = 0; i < 8; ++i) stress_point_coeffs(c, sp, i, pt + 3 * i);
    member_row(c, pt, Fl, 1.0 / fy, row);
    if (rows) for (int k = 0; k < 7; ++k) rows[(size_t)m * 7 + k] = row[k];
    if (endf) for (int k = 0; k < 12; ++k) endf[(size_t)m * 12 + k] = k < 6 ? -Fl[k] : Fl[k];
}

// ----------------------------------------------------------------------------------------------
// max nodal translation per phase: block = 128 phases x NCHUNK nodes
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(PH_TPB)
k_node_post(int Nn, int ldP, int n_pad, const double* __restrict__ X, const int* __restrict__ node2slot,
            double* __restrict__ part_disp, int* __restrict__ part_node) {
    int chunk = blockIdx.x, n0 = chunk * NCHUNK;
    int nn = min(NCHUNK, Nn - n0);
    int p = blockIdx.y * blockDim.x + threadIdx.x;
    if (p >= ldP) return;
    const size_t xbase = (size_t)(p / SLAB) * (size_t)n_pad * SLAB + (size_t)(p % SLAB);
    double best = 0.0; int bnode = -1;
    for (int i = 0; i < nn; ++i) {
        int s = node2slot[n0 + i];
        double d = 0.0;
        if (s >= 0) {
            double ux = X[xbase + (size_t)s * SLAB], uy = X[xbase + (size_t)(s + 1) * SLAB], uz = X[xbase + (size_t)(s + 2) * SLAB];
            d = sqrt(ux * ux + uy * uy + uz * uz);
        }
        if (d > best) { best = d; bnode = n0 + i; }
    }
    part_disp[(size_t)chunk * ldP + p] = best;
    part_node[(size_t)chunk * ldP + p] = bnode;
}

// ----------------------------------------------------------------------------------------------
// K u - F at the 6 DOF of selected nodes, matrix-free over the incident members (member order).
//   mode 0: nodes = fixed nodes, F from Ffix[6*f+c][ldP]           -> reactions (GUI.py:493)
//   mode 1: nodes = free nodes (by slot), F from the saved RHS copy -> residual diagnostic
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(PH_TPB)
k_node_residual(int n_sel, const int* __restrict__ sel_nodes, int ldP, int n_pad,
                const double* __restrict__ X, const int* __restrict__ node2slot, const int* __restrict__ conn,
                const int* __restrict__ adj_ptr, const int* __restrict__ adj, const double* __restrict__ Ke,
                const double* __restrict__ Fsel /* [n_sel*6][ldP] */, double* __restrict__ out /* [n_sel*6][ldP] */) {
    int f = blockIdx.y;
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n_sel || p >= ldP) return;
    int node = sel_nodes[f];
    double r[6] = {0, 0, 0, 0, 0, 0};
    for (int q = adj_ptr[node]; q < adj_ptr[node + 1]; ++q) {
        int m = adj[q] >> 1, end = adj[q] & 1;
        double ue[12];
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            ue[k] = load_u(X, node2slot, conn[2 * m], k, p, n_pad);
            ue[6 + k] = load_u(X, node2slot, conn[2 * m + 1], k, p, n_pad);
        }
        const double* ke = Ke + (size_t)m * 144 + (size_t)(6 * end) * 12;
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            double s = 0.0;
#pragma unroll
            for (int j = 0; j < 12; ++j) s = fma(ke[i * 12 + j], ue[j], s);
            r[i] += s;
        }
    }
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        size_t o = (size_t)(6 * f + i) * ldP + p;
        out[o] = r[i] - Fsel[o];
    }
}

}  // namespace jk
