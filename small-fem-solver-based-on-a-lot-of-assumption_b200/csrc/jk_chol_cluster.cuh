// jk_chol_cluster.cuh -- the whole tile-banded Cholesky in ONE persistent kernel on a thread-block
// cluster (8 CTAs on 8 SMs of one GPC, hardware cluster barriers between phases).
//
// The factorisation of a narrow band is a latency chain (n pivots in sequence), not a throughput
// problem: the first version spent 60 us per tile column on three kernel launches.  Here a column
// costs two cluster barriers:
//
//   phase A(k)  panel:   tiles (k+1..k+w, k) <- A * L_kk^{-T}       one tile per CTA, warp-local blocked
//                                                                   TRSM on DMMA with the 16x16 diagonal-
//                                                                   block inverses produced by the potrf
//   phase B(k)  update:  tile (i,j) -= L_ik L_jk^T, k < j <= i <= k+w   DMMA, tiles dealt round-robin;
//               CTA 0 takes tile (k+1,k+1) first, keeps it in shared memory and factors it at once
//               (lookahead: the next potrf overlaps the rest of this column's update).
//
// potrf of a 64x64 tile: eight 8-column panels, see potrf64_smem.
#pragma once
#include <cooperative_groups.h>

#include "jk_chol.cuh"

namespace jk {

#ifndef JK_TRSM_UNROLL
#define JK_TRSM_UNROLL 1      // fully unrolled panel TRSM: the independent block updates of a step overlap (factor 2.92 -> 2.79 ms)
#endif
#ifndef JK_CHOL_CLUSTER
#define JK_CHOL_CLUSTER 8       // CTAs (SMs) per factorisation cluster
#endif
constexpr int CHOL_CLUSTER = JK_CHOL_CLUSTER;
constexpr int CHOL_THREADS = 256;
constexpr int DI_LD = 12;                                   // row stride of an 8x8 inverse block in smem (== 12 mod 16: conflict-free)
constexpr int DI_BLK = 8 * DI_LD;
constexpr size_t CHOL_CLUSTER_SMEM = (size_t)(3 * NB * LS_LD + 8 * DI_BLK) * sizeof(double);

// 1/d and 1/sqrt(d) to ~1 ulp from the MUFU seeds (rel. error ~2^-22) and ONE cubic correction each:
// a chain of 3-4 dependent FMAs instead of the 8-10 of the library routines.  d must be a normal positive number.
__device__ __forceinline__ double fast_rcp(double d) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
    double e = fma(-d, y, 1.0);
    double p = fma(e, e, e);
    return fma(y, p, y);
}
__device__ __forceinline__ double fast_rsqrt(double d) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
    double t = d * y;
    double e = fma(-t, y, 1.0);
    double p = fma(0.375, e, 0.5);
    double q = y * e;
    return fma(q, p, y);
}

// Cholesky of the 64x64 tile T (smem, row stride LS_LD, lower triangle) in place.  Di[8][8][DI_LD]
// receives the inverses of the eight 8x8 diagonal blocks of L.  All 256 threads of the CTA call this.
//
// Eight 8-column panels.  Per panel: warp 0 factors the 8x8 diagonal block entirely in registers (every
// lane the same values, no communication).  The pivot recurrence is square-root free (LDL^T form:
// d' = d11 - d10^2 / d00 needs only the reciprocal), the column scaling by 1/sqrt(d) happens off the
// chain.  Then one DMMA phase solves the rows below against the block inverse and one DMMA phase applies
// the rank-8 update to the trailing block.
// 8x8 diagonal block p of T (already carrying the updates of the panels before it): factor it in registers (every lane
// of the calling warp holds the same values, no communication), write L_pp back and its inverse to Dp.
__device__ __forceinline__ void potrf8_warp(double* __restrict__ T, double* __restrict__ Dp, int c0, int* __restrict__ info, int pivot_base, int lane) {
    double d[8][8], m[8][8], rs[8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j <= i; ++j) d[i][j] = T[(c0 + i) * LS_LD + c0 + j];
    bool bad = false;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        double dc = d[c][c];
        if (!(dc > 0.0)) { if (!bad && lane == 0) atomicCAS(info, 0, pivot_base + c0 + c + 1); bad = true; dc = 1.0; }
        const double inv = fast_rcp(dc);
        rs[c] = fast_rsqrt(dc);
        double t[8];
#pragma unroll
        for (int i = c + 1; i < 8; ++i) t[i] = d[i][c] * inv;
#pragma unroll
        for (int j = c + 1; j < 8; ++j)
#pragma unroll
            for (int i = j; i < 8; ++i) {
                if (i == j && j == c + 1) d[i][j] = fma(-(d[i][c] * d[i][c]), inv, d[i][j]);   // next pivot: shortest chain
                else d[i][j] = fma(-d[i][c], t[j], d[i][j]);
            }
        // the same row operations applied to the identity give Ltilde^-1 (product of the elimination matrices): the
        // inverse is complete one FMA after the last pivot instead of a 28-FMA substitution chain afterwards
#pragma unroll
        for (int i = c + 1; i < 8; ++i) {
#pragma unroll
            for (int j = 0; j < c; ++j) m[i][j] = fma(-t[i], m[c][j], m[i][j]);
            m[i][c] = -t[i];
        }
        d[c][c] = dc;
    }
    // L = Ltilde D^(1/2): scale the columns off the pivot chain
#pragma unroll
    for (int c = 0; c < 8; ++c) {
#pragma unroll
        for (int i = c + 1; i < 8; ++i) d[i][c] *= rs[c];
        d[c][c] *= rs[c];
    }
    // M = L^-1 = D^(-1/2) Ltilde^-1: scale the rows, 1/L_ii = rs_i
#pragma unroll
    for (int i = 0; i < 8; ++i) {
#pragma unroll
        for (int j = 0; j < i; ++j) m[i][j] *= rs[i];
        m[i][i] = rs[i];
    }
    if (lane == 0) {
        // 16-byte stores, lower triangle only (the entry just above the diagonal that a pair may cover gets 0;
        // the rest of Dp's upper triangle is zeroed once by the caller): 40 stores instead of 100 on the chain
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j <= i; j += 2) {
                *reinterpret_cast<double2*>(T + (c0 + i) * LS_LD + c0 + j) = make_double2(d[i][j], j + 1 <= i ? d[i][j + 1] : 0.0);
                *reinterpret_cast<double2*>(Dp + i * DI_LD + j) = make_double2(m[i][j], j + 1 <= i ? m[i][j + 1] : 0.0);
            }
    }
}

// Eight 8-column panels with LOOKAHEAD inside the tile: after the rows below panel p are solved, warp 0 alone applies
// the rank-8 update to the next diagonal block and factors it at once, while warps 1-7 apply the update to the rest of
// the trailing block.  The chain per panel is  8x8 factor -> barrier -> rows-below solve (DMMA) -> barrier;  the
// trailing update is off it.
__device__ void potrf64_smem(double* __restrict__ T, double* __restrict__ Di, int* __restrict__ info, int pivot_base) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int fr = lane >> 2, fk = lane & 3;
    if (warp == 0) potrf8_warp(T, Di, 0, info, pivot_base, lane);
    for (int p = 0; p < 8; ++p) {
        const int c0 = 8 * p;
        double* Dp = Di + p * DI_BLK;
        __syncthreads();                                    // L_pp, its inverse and every update of panel p-1 are in place
        const int nmt = 7 - p;                              // 8-row tiles below the diagonal block
        if (nmt == 0) break;
        // ---- rows below: X = A * M^T (in place), one 8-row tile per warp ----
        if (warp < nmt) {
            const int row0 = c0 + 8 + 8 * warp;
            double x0 = 0.0, x1 = 0.0;
#pragma unroll
            for (int k4 = 0; k4 < 2; ++k4)
                dmma(x0, x1, T[(row0 + fr) * LS_LD + c0 + 4 * k4 + fk], Dp[fr * DI_LD + 4 * k4 + fk]);
            __syncwarp();
            T[(row0 + fr) * LS_LD + c0 + 2 * fk] = x0;
            T[(row0 + fr) * LS_LD + c0 + 2 * fk + 1] = x1;
        }
        __syncthreads();
        // ---- trailing block: C(mi,nj) -= X_mi X_nj^T for the lower triangle of 8x8 tiles.  Tile 0 = the next diagonal
        //      block: warp 0 takes it and goes straight on to factor it; the other warps share the rest. ----
        const int ntile = nmt * (nmt + 1) / 2;
        auto update_tile = [&](int t) {
            int mi = 0;
            while ((mi + 1) * (mi + 2) / 2 <= t) ++mi;
            const int nj = t - mi * (mi + 1) / 2;
            const int ri = c0 + 8 + 8 * mi, rj = c0 + 8 + 8 * nj;
            double* cp = T + (ri + fr) * LS_LD + rj + 2 * fk;
            double c0v = cp[0], c1v = cp[1];
#pragma unroll
            for (int k4 = 0; k4 < 2; ++k4)
                dmma(c0v, c1v, -T[(ri + fr) * LS_LD + c0 + 4 * k4 + fk], T[(rj + fr) * LS_LD + c0 + 4 * k4 + fk]);
            cp[0] = c0v; cp[1] = c1v;
        };
        if (warp == 0) {
            update_tile(0);
            __syncwarp();
            potrf8_warp(T, Di + (p + 1) * DI_BLK, c0 + 8, info, pivot_base, lane);
        } else {
            for (int t = warp; t < ntile; t += CHOL_THREADS / 32 - 1) update_tile(t);
        }
    }
}

__device__ __forceinline__ void store_tile(double* __restrict__ g, const double* __restrict__ T, int tid, int nthreads) {
    for (int q = tid; q < NB * (NB / 2); q += nthreads) {
        int r = q / (NB / 2), cc = q % (NB / 2);
        *reinterpret_cast<double2*>(g + r * NB + 2 * cc) = make_double2(T[r * LS_LD + 2 * cc], T[r * LS_LD + 2 * cc + 1]);
    }
}

// tile (row-block of As) <- As * L^{-T}, warp-local: warp w owns rows 8w..8w+7 of As.  Blocked by 8 columns with
// the 8x8 diagonal-block inverses Di of L (Ls): X_b = A_b M_b^T ; A_b2 -= X_b L_{b2,b}^T (b2 > b).
__device__ __forceinline__ void trsm64_warp(double* __restrict__ As, const double* __restrict__ Ls, const double* __restrict__ Di) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int fr = lane >> 2, fk = lane & 3;
    double* arow = As + (8 * warp + fr) * LS_LD;
#if JK_TRSM_UNROLL
#pragma unroll
#else
#pragma unroll 1
#endif
    for (int b = 0; b < 8; ++b) {
        double x0 = 0.0, x1 = 0.0;
#pragma unroll
        for (int k4 = 0; k4 < 2; ++k4) dmma(x0, x1, arow[8 * b + 4 * k4 + fk], Di[b * DI_BLK + fr * DI_LD + 4 * k4 + fk]);
        __syncwarp();
        arow[8 * b + 2 * fk] = x0; arow[8 * b + 2 * fk + 1] = x1;
        __syncwarp();
        const double a0 = -arow[8 * b + fk], a1 = -arow[8 * b + 4 + fk];
#if JK_TRSM_UNROLL
#pragma unroll
#endif
        for (int b2 = b + 1; b2 < 8; ++b2) {
            double* cp = arow + 8 * b2 + 2 * fk;
            double c0v = cp[0], c1v = cp[1];
            dmma(c0v, c1v, a0, Ls[(8 * b2 + fr) * LS_LD + 8 * b + fk]);
            dmma(c0v, c1v, a1, Ls[(8 * b2 + fr) * LS_LD + 8 * b + 4 + fk]);
            cp[0] = c0v; cp[1] = c1v;
        }
        __syncwarp();
    }
}

// acc(8 rows x 64 cols per warp) = As * Bs^T over the full 64-deep contraction
__device__ __forceinline__ void gemm64_nt(const double* __restrict__ As, const double* __restrict__ Bs, double (&acc)[8][2]) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int fr = lane >> 2, fk = lane & 3;
    const double* ap = As + (8 * warp + fr) * LS_LD + fk;
    const double* bp = Bs + fr * LS_LD + fk;
#pragma unroll 4
    for (int k4 = 0; k4 < NB / 4; ++k4) {
        double av = ap[4 * k4];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) dmma(acc[nt][0], acc[nt][1], av, bp[nt * 8 * LS_LD + 4 * k4]);
    }
}

// One chain of the factorisation: columns [k_begin, k_end) of a tile-banded matrix are eliminated; the trailing
// rows (tiles >= k_end) only receive their Schur-complement updates.
struct CholChain { double* tiles; double* dinv; int NT, bw, k_begin, k_end; };

// Grid = n_chains clusters of CHOL_CLUSTER CTAs; cluster c works on chain[c].  With two chains the band is eliminated
// from both ends at once (the second chain is the same matrix in reversed order), halving the pivot chain.
__global__ void __cluster_dims__(CHOL_CLUSTER, 1, 1) __launch_bounds__(CHOL_THREADS, 1)
k_band_chol_cluster(CholChain chain0, CholChain chain1, int* __restrict__ info,
                    long long* __restrict__ prof /* optional [NT][8] clock stamps of chain 0, CTA 0 (debug) */,
                    unsigned* __restrict__ started = nullptr /* optional: every CTA adds 1 as soon as it is resident */) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    // Start gate: the stream that launches the (SM-filling) Morison kernel waits on this counter, so that the factor
    // clusters are resident first.  Otherwise the Morison grid can take every SM a moment earlier and the clusters -- which
    // need eight nearly empty SMs of one GPC at the same time -- starve until it has run out of blocks (+2.7 ms per step).
    if (started != nullptr && threadIdx.x == 0) { atomicAdd(started, 1u); __threadfence(); }
    const bool second = blockIdx.x >= CHOL_CLUSTER;
    const CholChain ch = second ? chain1 : chain0;
    double* __restrict__ tiles = ch.tiles;
    double* __restrict__ dinv = ch.dinv;
    const int NT = ch.NT, bw = ch.bw;
    if (second) prof = nullptr;
    extern __shared__ __align__(16) double smem[];
    double* As = smem;
    double* Bs = As + NB * LS_LD;
    double* Cs = Bs + NB * LS_LD;
    double* Di = Cs + NB * LS_LD;                            // [8][8][DI_LD]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int fr = lane >> 2, fk = lane & 3;
    const int cta = (int)cluster.block_rank();

    auto store_dinv = [&](int k) {
        double* g = dinv + (size_t)k * 512;
        for (int q = tid; q < 512; q += CHOL_THREADS) { int b = q >> 6, r = (q >> 3) & 7, c = q & 7; g[q] = Di[b * DI_BLK + r * DI_LD + c]; }
    };
    auto load_dinv = [&](int k) {
        const double* g = dinv + (size_t)k * 512;
        for (int q = tid; q < 512; q += CHOL_THREADS) { int b = q >> 6, r = (q >> 3) & 7, c = q & 7; Di[b * DI_BLK + r * DI_LD + c] = __ldcg(g + q); }
    };

    for (int q = tid; q < 8 * DI_BLK; q += CHOL_THREADS) Di[q] = 0.0;    // potrf64_smem only writes the lower triangles
    __syncthreads();
    if (cta == 0 && ch.k_begin < ch.k_end) {                 // prologue: potrf(k_begin)
        load_tile_async(Cs, tiles + tile_off(ch.k_begin, ch.k_begin, bw), tid, CHOL_THREADS);
        cp_async_commit(); cp_async_wait<0>();
        __syncthreads();
        potrf64_smem(Cs, Di, info, ch.k_begin * NB);
        store_tile(tiles + tile_off(ch.k_begin, ch.k_begin, bw), Cs, tid, CHOL_THREADS);
        store_dinv(ch.k_begin);
    }
    cluster.sync();

#define JK_STAMP(i) do { if (prof && cta == 0 && tid == 0) prof[(size_t)k * 8 + (i)] = clock64(); } while (0)
    // Split cluster barrier (arrive = release, wait = acquire): CTA 0 arrives on barrier A as soon as L_{k+1,k} is in
    // global memory and only waits on it after it has factored the next diagonal tile, so the critical chain
    //   trsm(k+1,k) -> syrk(k+1,k+1) -> potrf(k+1)
    // never stalls on the other CTAs, whose panel / update work for column k hides under it.
    auto cluster_arrive = [] { asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory"); };
    auto cluster_wait = [] { asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory"); };
    double* Lk = Cs;     // CTA 0: current diagonal factor L_kk (its 8x8 block inverses are in Di)
    double* Nx = Bs;     // CTA 0: next diagonal tile being built
    for (int k = ch.k_begin; k < ch.k_end; ++k) {
        const int w = min(bw, NT - 1 - k);
        const bool factor_next = (k + 1 < ch.k_end);         // lookahead potrf of the next diagonal tile
        const int ntile = w * (w + 1) / 2;
        JK_STAMP(0);
        if (cta == 0) {
            // ---------------- critical chain, all operands in this CTA's shared memory ----------------
            if (w >= 1) {
                load_tile_async(As, tiles + tile_off(k + 1, k, bw), tid, CHOL_THREADS);
                load_tile_async(Nx, tiles + tile_off(k + 1, k + 1, bw), tid, CHOL_THREADS);
                cp_async_commit(); cp_async_wait<0>();
                __syncthreads();
                JK_STAMP(1);
                trsm64_warp(As, Lk, Di);                      // L_{k+1,k} = A L_kk^{-T}
                __syncthreads();
                store_tile(tiles + tile_off(k + 1, k, bw), As, tid, CHOL_THREADS);
                JK_STAMP(2);
            } else if (factor_next) {                         // decoupled next tile (no panel): just fetch it
                load_tile_async(Nx, tiles + tile_off(k + 1, k + 1, bw), tid, CHOL_THREADS);
                cp_async_commit(); cp_async_wait<0>();
                __syncthreads();
            }
            cluster_arrive();                                 // barrier A: L_{k+1,k} published
            if (w >= 1) {
                // D_{k+1} -= L_{k+1,k} L_{k+1,k}^T, lower 8x8 tiles only; row blocks paired (w, 7-w) per SM sub-partition
                const int rb = (warp < 4) ? warp : 11 - warp;
                const double* ap = As + (8 * rb + fr) * LS_LD + fk;
                const double* bp = As + fr * LS_LD + fk;
                double acc[8][2];
#pragma unroll
                for (int nt = 0; nt < 8; ++nt) acc[nt][0] = acc[nt][1] = 0.0;
#pragma unroll 2
                for (int k4 = 0; k4 < NB / 4; ++k4) {
                    const double av = ap[4 * k4];
#pragma unroll
                    for (int nt = 0; nt < 8; ++nt)
                        if (nt <= rb) dmma(acc[nt][0], acc[nt][1], av, bp[nt * 8 * LS_LD + 4 * k4]);
                }
#pragma unroll
                for (int nt = 0; nt < 8; ++nt)
                    if (nt <= rb) {
                        double* cp = Nx + (8 * rb + fr) * LS_LD + 8 * nt + 2 * fk;
                        cp[0] -= acc[nt][0]; cp[1] -= acc[nt][1];
                    }
                __syncthreads();
            }
            if (w >= 1 || factor_next) {
                JK_STAMP(3);
                if (factor_next) potrf64_smem(Nx, Di, info, (k + 1) * NB);
                JK_STAMP(4);
                store_tile(tiles + tile_off(k + 1, k + 1, bw), Nx, tid, CHOL_THREADS);
                if (factor_next) store_dinv(k + 1);
                double* tmp = Lk; Lk = Nx; Nx = tmp;
            }
            cluster_wait();                                   // barrier A
        } else {
            // ---------------- phase A: the other panel tiles (k+2.., k) ----------------
            if (cta - 1 < w - 1) {
                load_tile_async(Bs, tiles + tile_off(k, k, bw), tid, CHOL_THREADS);
                cp_async_commit();
                load_dinv(k);
            }
            for (int q = 1 + (cta - 1); q < w; q += CHOL_CLUSTER - 1) {
                double* g = tiles + tile_off(k + 1 + q, k, bw);
                load_tile_async(As, g, tid, CHOL_THREADS);
                cp_async_commit(); cp_async_wait<0>();
                __syncthreads();
                trsm64_warp(As, Bs, Di);
                __syncthreads();
                store_tile(g, As, tid, CHOL_THREADS);
                __syncthreads();
            }
            cluster_arrive();
            cluster_wait();                                   // barrier A: every L_ik of column k is in global memory
            // ---------------- phase B: trailing update, tiles t = 1.. (tile 0 = (k+1,k+1) belongs to CTA 0) ----------------
            for (int t = 1 + (cta - 1); t < ntile; t += CHOL_CLUSTER - 1) {
                int bi = 0;
                while ((bi + 1) * (bi + 2) / 2 <= t) ++bi;
                const int bj = t - bi * (bi + 1) / 2;
                const int i = k + 1 + bi, j = k + 1 + bj;
                load_tile_async(As, tiles + tile_off(i, k, bw), tid, CHOL_THREADS);
                load_tile_async(Bs, tiles + tile_off(j, k, bw), tid, CHOL_THREADS);
                cp_async_commit(); cp_async_wait<0>();
                __syncthreads();
                double acc[8][2];
#pragma unroll
                for (int nt = 0; nt < 8; ++nt) acc[nt][0] = acc[nt][1] = 0.0;
                gemm64_nt(As, Bs, acc);
                double* gc = tiles + tile_off(i, j, bw);
#pragma unroll
                for (int nt = 0; nt < 8; ++nt) {
                    double2* p = reinterpret_cast<double2*>(gc + (8 * warp + fr) * NB + 8 * nt + 2 * fk);
                    double2 v = __ldcg(p);
                    v.x -= acc[nt][0]; v.y -= acc[nt][1];
                    *p = v;
                }
                __syncthreads();
            }
        }
        JK_STAMP(5);
        cluster.sync();                                       // barrier B: column k done
        JK_STAMP(6);
    }
#undef JK_STAMP
}

// Separator merge after the two chains stopped at their separator rows: the Schur complement of the separator is the
// sum of both chains' contributions.  Chain 1 holds K_SS - W_A W_A^T in its trailing block, chain 0's twin holds
// -W_B W_B^T with the separator NODES in reversed order (DOF order inside a node unchanged).  Lower triangle only.
__global__ void k_sep_merge_tiles(double* __restrict__ t0, int bw0, int kS0, const double* __restrict__ t1, int bw1, int kS1, int nS /* nodes */) {
    const int n = 6 * nS;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)n * n) return;
    const int r = (int)(idx / n), c = (int)(idx % n);
    if (c > r) return;
    const int r1 = 6 * (nS - 1 - r / 6) + r % 6, c1 = 6 * (nS - 1 - c / 6) + c % 6;
    const int hi = max(r1, c1), lo = min(r1, c1);
    const int R0 = kS0 * NB + r, C0 = kS0 * NB + c, R1 = kS1 * NB + hi, C1 = kS1 * NB + lo;
    if (R0 / NB - C0 / NB > bw0) return;                     // outside the band: both are structurally zero
    double v = (R1 / NB - C1 / NB > bw1) ? 0.0 : t1[tile_off(R1 / NB, C1 / NB, bw1) + (size_t)(R1 % NB) * NB + (C1 % NB)];
    t0[tile_off(R0 / NB, C0 / NB, bw0) + (size_t)(R0 % NB) * NB + (C0 % NB)] += v;
}

// right-hand sides / solutions of the separator rows between the two chains' row blocks of the slab-packed array:
//   mode 0:  X[row_a + r] += X[row_b + rev(r)]      (forward: add the second chain's partial sums)
//   mode 1:  X[row_b + rev(r)] = X[row_a + r]        (backward: hand the separator solution to the second chain)
__global__ void k_sep_exchange(double* __restrict__ X, int n_pad, int ldP, int row_a, int row_b, int nS, int mode) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    if (p >= ldP || r >= 6 * nS) return;
    const int r1 = 6 * (nS - 1 - r / 6) + r % 6;
    const size_t ia = rhs_off(row_a + r, p, n_pad), ib = rhs_off(row_b + r1, p, n_pad);
    if (mode == 0) X[ia] += X[ib]; else X[ib] = X[ia];
}

}  // namespace jk
