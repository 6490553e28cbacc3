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
// WANT_L / WANT_M: which of the two results this warp produces.  The pivot recurrence itself is cheap (about 150 of the ~400
// FP64 instructions); two warps on different schedulers running it redundantly -- one finishing L_pp, the other the inverse --
// halve the issue-bound time of the block (1,600 -> ~800 clocks) without any exchange between them.
template <bool WANT_L, bool WANT_M>
__device__ __forceinline__ void potrf8_warp(double* __restrict__ T, double* __restrict__ Dp, int c0, int* __restrict__ info, int pivot_base, int lane) {
    double d[8][8], m[8][8], rs[8];
    constexpr bool PAIR_L = WANT_L && !WANT_M, PAIR_M = WANT_M && !WANT_L;
    // A pair of warps works on the same block (one finishes L_pp, the other the inverse) and L_pp overwrites the block both
    // read.  The inverse's warp reads it with volatile loads (which the compiler may not sink below the barrier instruction)
    // and arrives on a named barrier at once; the L warp waits on that barrier just before its stores -- by then the other
    // warp's loads are a thousand clocks old, so nobody ever blocks.  (Without the hand-shake about one factorisation in
    // three met a "non-positive pivot": the compiler is free to read T late in a warp that never writes it.)
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j <= i; ++j) {
            if (PAIR_M) asm volatile("ld.shared.f64 %0, [%1];\n" : "=d"(d[i][j]) : "r"((unsigned)__cvta_generic_to_shared(T + (c0 + i) * LS_LD + c0 + j)) : "memory");
            else d[i][j] = T[(c0 + i) * LS_LD + c0 + j];
        }
    if (PAIR_M) asm volatile("bar.arrive 3, 64;\n" ::: "memory");
    bool bad = false;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        double dc = d[c][c];
        if (!(dc > 0.0)) { if (WANT_L && !bad && lane == 0) atomicCAS(info, 0, pivot_base + c0 + c + 1); bad = true; dc = 1.0; }
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
        if (WANT_M) {
#pragma unroll
            for (int i = c + 1; i < 8; ++i) {
#pragma unroll
                for (int j = 0; j < c; ++j) m[i][j] = fma(-t[i], m[c][j], m[i][j]);
                m[i][c] = -t[i];
            }
        }
        d[c][c] = dc;
    }
    if (WANT_L) {
        // L = Ltilde D^(1/2): scale the columns off the pivot chain
#pragma unroll
        for (int c = 0; c < 8; ++c) {
#pragma unroll
            for (int i = c + 1; i < 8; ++i) d[i][c] *= rs[c];
            d[c][c] *= rs[c];
        }
    }
    if (WANT_M) {
        // M = L^-1 = D^(-1/2) Ltilde^-1: scale the rows, 1/L_ii = rs_i
#pragma unroll
        for (int i = 0; i < 8; ++i) {
#pragma unroll
            for (int j = 0; j < i; ++j) m[i][j] *= rs[i];
            m[i][i] = rs[i];
        }
    }
    if (PAIR_L) asm volatile("bar.sync 3, 64;\n" ::: "memory");       // the partner has read the block (see above)
    if (lane == 0) {
        // 16-byte stores, lower triangle only (the entry just above the diagonal that a pair may cover gets 0;
        // the rest of Dp's upper triangle is zeroed once by the caller): 20 stores per result
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j <= i; j += 2) {
                if (WANT_L) *reinterpret_cast<double2*>(T + (c0 + i) * LS_LD + c0 + j) = make_double2(d[i][j], j + 1 <= i ? d[i][j + 1] : 0.0);
                if (WANT_M) *reinterpret_cast<double2*>(Dp + i * DI_LD + j) = make_double2(m[i][j], j + 1 <= i ? m[i][j + 1] : 0.0);
            }
    }
}

// The same 8x8 factorisation with the work DEALT OVER THE LANES instead of replicated in every lane: lane (r, q) = (lane / 4,
// lane % 4) owns the two entries (r, 2q) and (r, 2q + 1) of the (symmetric, both triangles carried) block and of the
// unit-lower inverse being accumulated -- the DMMA accumulator layout.  Per pivot a lane needs its row's multiplier
// d[r][c], the pivot row entries d[c][2q..2q+1] and the inverse's row c: six register shuffles.  The replicated form issues
// ~400 warp-wide FP64 instructions per block (two issue cycles each: 1,600 clocks measured, issue-bound); this one ~15 per
// pivot, so the block costs its dependency chain  fma -> shuffle -> reciprocal  per pivot.  Same recurrences
// (d' = d - (a b) / piv, LDL^T form, square roots off the chain); products are associated as (a b) * inv.
__device__ __forceinline__ void potrf8_lanes(double* __restrict__ T, double* __restrict__ Dp, int c0, int* __restrict__ info, int pivot_base, int lane) {
    constexpr unsigned FULL = 0xffffffffu;
    const int r = lane >> 2, q = lane & 3, j0 = 2 * q, j1 = 2 * q + 1;
    double d0 = T[(c0 + max(r, j0)) * LS_LD + c0 + min(r, j0)];
    double d1 = T[(c0 + max(r, j1)) * LS_LD + c0 + min(r, j1)];
    double m0 = 0.0, m1 = 0.0;                 // inverse of the unit-lower factor, strictly lower part
    double rs_r = 0.0, rs_0 = 0.0, rs_1 = 0.0;  // 1 / sqrt(pivot) of row r and of columns j0, j1
    int bad = -1;                               // first non-positive pivot (uniform over the lanes); no branches on the chain
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        double piv = __shfl_sync(FULL, (c & 1) ? d1 : d0, 4 * c + (c >> 1));
        // everything that does not need 1 / piv first: the shuffles and the products run under the reciprocal
        const double a = __shfl_sync(FULL, (c & 1) ? d1 : d0, (lane & ~3) | (c >> 1));      // d[r][c]
        const double b0 = __shfl_sync(FULL, d0, 4 * c + q), b1 = __shfl_sync(FULL, d1, 4 * c + q);   // d[c][j0], d[c][j1]
        const double mc0 = __shfl_sync(FULL, m0, 4 * c + q), mc1 = __shfl_sync(FULL, m1, 4 * c + q); // m[c][j0], m[c][j1]
        const bool neg = !(piv > 0.0);
        bad = (neg && bad < 0) ? c : bad;
        piv = neg ? 1.0 : piv;
        const double inv = fast_rcp(piv);
        const double ab0 = a * b0, ab1 = a * b1, am0 = a * mc0, am1 = a * mc1;
        const double nd0 = fma(-ab0, inv, d0), nd1 = fma(-ab1, inv, d1);
        const double nm0 = fma(-am0, inv, m0), nm1 = fma(-am1, inv, m1), ta = -(a * inv);
        const bool below = r > c;
        d0 = (below && j0 > c) ? nd0 : d0;
        d1 = (below && j1 > c) ? nd1 : d1;
        m0 = below ? ((j0 < c) ? nm0 : (j0 == c ? ta : m0)) : m0;
        m1 = below ? ((j1 < c) ? nm1 : (j1 == c ? ta : m1)) : m1;
        const double rs = fast_rsqrt(piv);
        rs_r = (c == r) ? rs : rs_r; rs_0 = (c == j0) ? rs : rs_0; rs_1 = (c == j1) ? rs : rs_1;
    }
    if (bad >= 0 && lane == 0) atomicCAS(info, 0, pivot_base + c0 + bad + 1);
    // L = Ltilde D^(1/2) (entry (r, j) holds d[r][j] as it stood at pivot j: scale by 1 / sqrt(piv_j)); M = D^(-1/2) Ltilde^-1
    const double l0 = (j0 <= r) ? d0 * rs_0 : 0.0, l1 = (j1 <= r) ? d1 * rs_1 : 0.0;
    const double i0 = (j0 < r) ? m0 * rs_r : (j0 == r ? rs_r : 0.0), i1 = (j1 < r) ? m1 * rs_r : (j1 == r ? rs_r : 0.0);
    *reinterpret_cast<double2*>(T + (c0 + r) * LS_LD + c0 + j0) = make_double2(l0, l1);
    *reinterpret_cast<double2*>(Dp + r * DI_LD + j0) = make_double2(i0, i1);
}

#ifndef JK_POTRF8_LANES
#define JK_POTRF8_LANES 0      // 1: lane-distributed 8x8 pivot blocks (potrf8_lanes: measured no faster, 1,590 clocks -- the shuffles sit on the chain), 0: replicated in every lane
#endif
__device__ __forceinline__ void potrf8(double* __restrict__ T, double* __restrict__ Dp, int c0, int* __restrict__ info, int pivot_base, int lane) {
#if JK_POTRF8_LANES
    potrf8_lanes(T, Dp, c0, info, pivot_base, lane);
#else
    potrf8_warp<true, true>(T, Dp, c0, info, pivot_base, lane);
#endif
}

// Eight 8-column panels with LOOKAHEAD inside the tile: after the rows below panel p are solved, warp 0 alone applies
// the rank-8 update to the next diagonal block and factors it at once, while warps 1-7 apply the update to the rest of
// the trailing block.  The chain per panel is  8x8 factor -> barrier -> rows-below solve (DMMA) -> barrier;  the
// trailing update is off it.
__device__ void potrf64_smem(double* __restrict__ T, double* __restrict__ Di, int* __restrict__ info, int pivot_base) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int fr = lane >> 2, fk = lane & 3;
    if (warp == 0) potrf8(T, Di, 0, info, pivot_base, lane);
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
            potrf8(T, Di + (p + 1) * DI_BLK, c0 + 8, info, pivot_base, lane);
        } else {
            for (int t = warp; t < ntile; t += CHOL_THREADS / 32 - 1) update_tile(t);
        }
    }
}

// potrf64_smem with the 8x8 pivot blocks factored by TWO warps at once (warp 0 -> L_pp, warp 1 -> its inverse; see
// potrf8_warp): warp 0 applies the rank-8 update to the next diagonal block, the pair meets at a 64-thread named barrier and
// both read the block; warps 2-7 share the rest of the trailing update.
__device__ void potrf64_pipelined(double* __restrict__ T, double* __restrict__ Di, int* __restrict__ info, int pivot_base) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int fr = lane >> 2, fk = lane & 3;
    if (warp == 0) potrf8_warp<true, false>(T, Di, 0, info, pivot_base, lane);
    else if (warp == 1) potrf8_warp<false, true>(T, Di, 0, info, pivot_base, lane);
    for (int p = 0; p < 8; ++p) {
        const int c0 = 8 * p;
        double* Dp = Di + p * DI_BLK;
        __syncthreads();                                    // L_pp, its inverse and every update of panel p-1 are in place
        const int nmt = 7 - p;                              // 8-row tiles below the diagonal block
        if (nmt == 0) break;
        if (warp < nmt) {                                   // rows below: X = A * M^T (in place), one 8-row tile per warp
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
            asm volatile("bar.sync 2, 64;\n" ::: "memory");          // the next diagonal block is final: hand it to warp 1 as well
            potrf8_warp<true, false>(T, Di + (p + 1) * DI_BLK, c0 + 8, info, pivot_base, lane);
        } else if (warp == 1) {
            asm volatile("bar.sync 2, 64;\n" ::: "memory");
            potrf8_warp<false, true>(T, Di + (p + 1) * DI_BLK, c0 + 8, info, pivot_base, lane);
        } else {
            for (int t = warp - 1; t < ntile; t += CHOL_THREADS / 32 - 2) update_tile(t);
        }
    }
}

#ifndef JK_POTRF_TWO_WARPS
#define JK_POTRF_TWO_WARPS 1
#endif
__device__ __forceinline__ void potrf64(double* __restrict__ T, double* __restrict__ Di, int* __restrict__ info, int pivot_base) {
#if JK_POTRF_TWO_WARPS
    potrf64_pipelined(T, Di, info, pivot_base);
#else
    potrf64_smem(T, Di, info, pivot_base);
#endif
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

// C (lower 8x8 tiles) -= A A^T for 64x64 tiles in shared memory, all 8 warps: row blocks paired (w, 7-w) per SM
// sub-partition so that every scheduler issues the same number of DMMAs.  Ends with a CTA barrier.
__device__ __forceinline__ void syrk64_lower(const double* __restrict__ As, double* __restrict__ Cs) {
    // The 36 lower 8x8 tiles are dealt round-robin over the 8 warps (4 or 5 each, 80 DMMAs at most) instead of one row block per
    // warp (16 ... 128 DMMAs): the warp with the longest list sets the time of this step of the critical chain
    // (tools/chol_probe.cu: 5,100 -> 4,050 clocks; 8 warps can saturate the SM's DMMA pipe, 2,300 clocks, so what is left is
    // operand-load latency).
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int fr = lane >> 2, fk = lane & 3;
#pragma unroll
    for (int q = 0; q < 5; ++q) {
        const int t = warp + 8 * q;                       // tile index in the row-major lower triangle
        if (t >= 36) break;
        int mi = 0;
        while ((mi + 1) * (mi + 2) / 2 <= t) ++mi;
        const int nj = t - mi * (mi + 1) / 2;
        const double* ap = As + (8 * mi + fr) * LS_LD + fk;
        const double* bp = As + (8 * nj + fr) * LS_LD + fk;
        double c0 = 0.0, c1 = 0.0, e0 = 0.0, e1 = 0.0;   // two partial sums: even / odd k-groups
#pragma unroll
        for (int k4 = 0; k4 < NB / 4; k4 += 2) {
            dmma(c0, c1, ap[4 * k4], bp[4 * k4]);
            dmma(e0, e1, ap[4 * k4 + 4], bp[4 * k4 + 4]);
        }
        double2* cp = reinterpret_cast<double2*>(Cs + (8 * mi + fr) * LS_LD + 8 * nj + 2 * fk);
        double2 v = *cp;
        v.x -= c0 + e0; v.y -= c1 + e1;
        *cp = v;
    }
    __syncthreads();
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
        potrf64(Cs, Di, info, ch.k_begin * NB);
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
                syrk64_lower(As, Nx);                         // D_{k+1} -= L_{k+1,k} L_{k+1,k}^T
            }
            if (w >= 1 || factor_next) {
                JK_STAMP(3);
                if (factor_next) potrf64(Nx, Di, info, (k + 1) * NB);
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
