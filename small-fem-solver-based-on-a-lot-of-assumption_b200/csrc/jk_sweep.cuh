// jk_sweep.cuh -- the multi-right-hand-side triangular sweeps of FEMSolver.solve (GUI.py:481-490) for narrow
// bands, as a warp-specialised TMA / mbarrier pipeline.
//
// The sweeps are rewritten so that one tile row needs ONE exchange between the warps of a CTA instead of three:
//
//   forward  (Z_k = L_kk Y_k):   Z_k = B_k               + sum_{j<k}  U_kj Z_j ,   U_kj = -L_kj  Linv_jj
//   backward (L^T X = Y):        X_k = G_k Z_k           + sum_{i>k}  V_ki X_i ,   V_ki = -Linv_kk^T L_ik^T ,
//                                                                                  G_k  =  Linv_kk^T Linv_kk
//
// U, V and G are built once per factorisation (k_sweep_build) in exactly the order the sweep consumes them (the
// "tile stream"), already in DMMA fragment order, together with a 16-bit mask per 8-row block that marks the
// 4-column groups holding non-zeros: the band's zero blocks (the envelope of L is much narrower than its tile band)
// are skipped.  A sweep CTA owns one slab of SLAB right-hand sides:
//
//   * one elected thread of warp 8 (producer) walks the item list and moves one 32 KB tile per item global -> shared
//     with a single cp.async.bulk (TMA), completion on a per-stage mbarrier; the backward diagonal item also brings Z_k;
//   * warps 0-7 (consumers) each own two 8-row blocks (p and 7-p, so triangular masks balance) x two 8-column
//     blocks of the 64 x 32 result tile and run mma.sync.m8n8k4.f64 (DMMA) from the fragment-ordered tiles.  They
//     never meet at a CTA barrier: a warp waits on the stage's "full" mbarrier, releases it on the "empty" one, and
//     only the last item of a row (the one that needs the row solved just before) waits on that row's mbarrier.
//     The last SW_RING solved tiles stay in shared memory (fragment order), so X is never re-read from L2.
//
// Layouts ("fragment order"):
//   A tile (64 x 64):  element (r, c) at ((r/8)*16 + c/4)*32 + (r%8)*4 + c%4   -> one A fragment = 32 consecutive doubles
//   X tile (64 x 32):  element (r, c) at ((r/4)*4  + c/8)*32 + (c%8)*4 + r%4   -> one B fragment = 32 consecutive doubles
// Between the forward and the backward sweep the slab holds Z in X-tile fragment order (it is an intermediate that
// only the backward sweep reads, with one 16 KB bulk copy per tile row); right-hand sides and solutions are row-major.
#pragma once
#include "jk_common.cuh"

namespace jk {

#ifndef JK_SW_STAGES
#define JK_SW_STAGES 3
#endif
constexpr int SW_STAGES = JK_SW_STAGES;
constexpr int SW_RING = 5;                  // solved tiles kept in shared memory: tile half-bandwidth <= SW_RING - 1
constexpr int SW_MAX_BW = SW_RING - 1;
constexpr int SW_CONSUMER_WARPS = 8;
constexpr int SW_CONSUMERS = 32 * SW_CONSUMER_WARPS;
constexpr int SW_THREADS = SW_CONSUMERS + 32;
constexpr int SW_TILE = NB * NB;            // doubles per A tile
constexpr int SW_XTILE = NB * SLAB;         // doubles per X tile

// item flags
constexpr int SW_ROW_BEGIN = 1;             // first item of a tile row: reset the accumulators
constexpr int SW_DIAG = 2;                  // backward diagonal item: operand = Z_k staged by the producer
constexpr int SW_ROW_END = 4;               // last item of a tile row: store the row
constexpr int SW_NO_RING = 8;               // ROW_END: row is not needed by later rows of this sweep (partial separator rows)
constexpr int SW_INIT_RHS = 16;             // ROW_BEGIN: accumulators start from the right-hand side rows (row-major)
constexpr int SW_OUT_FRAG = 32;             // ROW_END: store the row to the slab in fragment order (forward Z)
constexpr int SW_NO_OPERAND = 64;           // placeholder item of a row without any tile (accumulators pass through)

// One item = 3 x uint4 in the program array:
//   [0] = {row, src, flags, xinfo}   xinfo: bits 0-7 operand ring slot, bit 8 its mbarrier parity, bits 16-23 output slot
//   [1] = {next_row, next_init, 0, 0}   (valid on ROW_END items: right-hand side rows to prefetch before the row is published)
//   [2] = 8 x 16-bit k-group masks (written by k_sweep_build)
constexpr int SW_ITEM_U4 = 3;

constexpr size_t SW_SMEM = (size_t)(SW_STAGES * SW_TILE + SW_XTILE + SW_RING * SW_XTILE) * sizeof(double)
                         + (size_t)SW_STAGES * SW_ITEM_U4 * sizeof(uint4) + 16 * sizeof(unsigned long long) + 128;

__host__ __device__ __forceinline__ int sw_a_index(int r, int c) { return ((r >> 3) * 16 + (c >> 2)) * 32 + (r & 7) * 4 + (c & 3); }
__host__ __device__ __forceinline__ int sw_x_index(int r, int c) { return ((r >> 2) * 4 + (c >> 3)) * 32 + (c & 7) * 4 + (r & 3); }

// ----------------------------------------------------------------------------------------------
// mbarrier / bulk-copy primitives (PTX ISA 8.x, sm_90+)
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned bar, unsigned parity) {
    unsigned ok;
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
// Bounded wait: a broken pipeline traps (the launch fails with an error) instead of hanging the device.
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    unsigned spins = 0;
    while (!mbar_try_wait(bar, parity)) { if (++spins > (1u << 26)) __trap(); }
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                 :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void consumer_bar_sync() { asm volatile("bar.sync 1, %0;\n" :: "n"(SW_CONSUMERS) : "memory"); }

// ----------------------------------------------------------------------------------------------
// K3d: build the tile stream of one sweep program.  One CTA per item:
//   forward  item (row k, src j):  T = -L_kj Linv_jj
//   backward diagonal (row k):     T =  Linv_kk^T Linv_kk
//   backward item (row k, src i):  T = -Linv_kk^T L_ik^T
// T is written in A-fragment order; the item's masks mark the (8-row block, 4-column group) cells with a non-zero.
// Structural zeros of L are exact zeros (the envelope is never filled), so the test is exact.
// ----------------------------------------------------------------------------------------------
constexpr int SWB_LD = NB + 1;
constexpr size_t SWB_SMEM = (size_t)2 * NB * SWB_LD * sizeof(double);
__global__ void __launch_bounds__(256) k_sweep_build(uint4* __restrict__ prog, double* __restrict__ stream,
                                                     const double* __restrict__ tiles, const double* __restrict__ Linv,
                                                     int bw, int backward) {
    extern __shared__ __align__(16) double smem[];
    double* Ps = smem;                    // P[r][m]
    double* Qs = smem + NB * SWB_LD;      // Q[m][c]
    __shared__ unsigned msk[8];
    const int n = blockIdx.x, tid = threadIdx.x;
    const uint4 it = prog[(size_t)n * SW_ITEM_U4];
    const int row = (int)it.x, src = (int)it.y, flags = (int)it.z;
    double* out = stream + (size_t)n * SW_TILE;
    if (tid < 8) msk[tid] = 0u;
    if (flags & SW_NO_OPERAND) {
        for (int idx = tid; idx < SW_TILE; idx += 256) out[idx] = 0.0;
        if (tid == 0) prog[(size_t)n * SW_ITEM_U4 + 2] = make_uint4(0u, 0u, 0u, 0u);
        return;
    }
    const double *ps, *qs;
    bool tp, tq;
    double sgn;
    if (!backward) { ps = tiles + tile_off(row, src, bw); tp = false; qs = Linv + (size_t)src * SW_TILE; tq = false; sgn = -1.0; }
    else if (flags & SW_DIAG) { ps = Linv + (size_t)row * SW_TILE; tp = true; qs = ps; tq = false; sgn = 1.0; }
    else { ps = Linv + (size_t)row * SW_TILE; tp = true; qs = tiles + tile_off(src, row, bw); tq = true; sgn = -1.0; }
    for (int idx = tid; idx < SW_TILE; idx += 256) {
        const int a = idx / NB, b = idx % NB;
        const double pv = ps[idx], qv = qs[idx];
        if (tp) Ps[b * SWB_LD + a] = pv; else Ps[a * SWB_LD + b] = pv;
        if (tq) Qs[b * SWB_LD + a] = qv; else Qs[a * SWB_LD + b] = qv;
    }
    __syncthreads();
    const int ty = tid / 16, tx = tid % 16;
    double acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
    for (int m = 0; m < NB; ++m) {
        double p[4], q[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) p[i] = Ps[(4 * ty + i) * SWB_LD + m];
#pragma unroll
        for (int j = 0; j < 4; ++j) q[j] = Qs[m * SWB_LD + tx + 16 * j];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = fma(p[i], q[j], acc[i][j]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int r = 4 * ty + i, c = tx + 16 * j;
            const double v = sgn * acc[i][j];
            out[sw_a_index(r, c)] = v;
            if (v != 0.0) atomicOr(&msk[r >> 3], 1u << (c >> 2));
        }
    __syncthreads();
    if (tid == 0)
        prog[(size_t)n * SW_ITEM_U4 + 2] = make_uint4(msk[0] | (msk[1] << 16), msk[2] | (msk[3] << 16), msk[4] | (msk[5] << 16), msk[6] | (msk[7] << 16));
}

// Inner products of one item for one consumer warp: acc[a][b] += A(row block a) * X(column block b) over the k-groups.
// R0 / R1: which of the warp's two row blocks take part.  All offsets are compile-time constants.
template <bool R0, bool R1>
__device__ __forceinline__ void sweep_mma_full(double (&acc)[2][2][2], const double* __restrict__ a0p, const double* __restrict__ a1p,
                                               const double* __restrict__ bp) {
    double a0[2] = {0.0, 0.0}, a1[2] = {0.0, 0.0}, b0[2], b1[2];
    if (R0) a0[0] = a0p[0];
    if (R1) a1[0] = a1p[0];
    b0[0] = bp[0]; b1[0] = bp[32];
#pragma unroll
    for (int k4 = 0; k4 < 16; ++k4) {
        const int cur = k4 & 1, nxt = cur ^ 1;
        if (k4 + 1 < 16) {      // fragments of the next k-group before the DMMAs of this one
            if (R0) a0[nxt] = a0p[(k4 + 1) * 32];
            if (R1) a1[nxt] = a1p[(k4 + 1) * 32];
            b0[nxt] = bp[(k4 + 1) * 128]; b1[nxt] = bp[(k4 + 1) * 128 + 32];
        }
        if (R0) { dmma(acc[0][0][0], acc[0][0][1], a0[cur], b0[cur]); dmma(acc[0][1][0], acc[0][1][1], a0[cur], b1[cur]); }
        if (R1) { dmma(acc[1][0][0], acc[1][0][1], a1[cur], b0[cur]); dmma(acc[1][1][0], acc[1][1][1], a1[cur], b1[cur]); }
    }
}
template <bool R0, bool R1>
__device__ __forceinline__ void sweep_mma_masked(double (&acc)[2][2][2], const double* __restrict__ a0p, const double* __restrict__ a1p,
                                                 const double* __restrict__ bp, unsigned mask) {
#pragma unroll
    for (int k4 = 0; k4 < 16; ++k4) {
        if (mask & (1u << k4)) {      // warp-uniform
            const double b0 = bp[k4 * 128], b1 = bp[k4 * 128 + 32];
            if (R0) { const double a = a0p[k4 * 32]; dmma(acc[0][0][0], acc[0][0][1], a, b0); dmma(acc[0][1][0], acc[0][1][1], a, b1); }
            if (R1) { const double a = a1p[k4 * 32]; dmma(acc[1][0][0], acc[1][0][1], a, b0); dmma(acc[1][1][0], acc[1][1][1], a, b1); }
        }
    }
}

// ----------------------------------------------------------------------------------------------
// K4 (narrow bands): one sweep of one chain over one slab per CTA
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(SW_THREADS, 1)
k_sweep(const uint4* __restrict__ prog, const double* __restrict__ stream, double* __restrict__ X,
        int n_items, int n_pad /* rows of the whole slab */, int row0 /* first row of this chain in the slab */,
        int pre_row /* first known tile row (backward, second chain) */, int npre /* known tile rows to preload */, int ktop,
        long long* __restrict__ prof /* nullable: [8 warps][8] clock sums of CTA 0 (JK_SWEEP_PROFILE) */) {
    extern __shared__ __align__(128) unsigned char sw_smem[];
    double* As = reinterpret_cast<double*>(sw_smem);                 // [SW_STAGES][SW_TILE]   A tiles, fragment order
    double* Bs = As + SW_STAGES * SW_TILE;                           // [SW_XTILE]             Z_k of the backward diagonal item
    double* Xr = Bs + SW_XTILE;                                      // [SW_RING][SW_XTILE]    newest solved tiles, fragment order
    uint4* Ds = reinterpret_cast<uint4*>(Xr + SW_RING * SW_XTILE);   // [SW_STAGES][SW_ITEM_U4] item descriptors
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(Ds + SW_STAGES * SW_ITEM_U4);
    const unsigned bar_full = smem_u32(bars), bar_empty = smem_u32(bars + SW_STAGES), bar_x = smem_u32(bars + 2 * SW_STAGES),
                   bar_bfull = smem_u32(bars + 2 * SW_STAGES + SW_RING), bar_bempty = smem_u32(bars + 2 * SW_STAGES + SW_RING + 1);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double* Xslab = X + ((size_t)blockIdx.x * (size_t)n_pad + (size_t)row0) * SLAB;

    if (tid == 0) {
        for (int s = 0; s < SW_STAGES; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, SW_CONSUMER_WARPS); }
        for (int s = 0; s < SW_RING; ++s) mbar_init(bar_x + 8 * s, SW_CONSUMERS);
        mbar_init(bar_bfull, 1);
        mbar_init(bar_bempty, SW_CONSUMER_WARPS);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        asm volatile("fence.proxy.async;\n" ::: "memory");
    }
    __syncthreads();

    if (warp == SW_CONSUMER_WARPS) {
        // ------------------------------- producer (one elected thread) -------------------------------
        // tiles are contiguous in the stream AND in shared memory (fragment order needs no padding): one bulk copy each
        if (lane == 0) {
            int ndiag = 0;
            uint4 d0 = make_uint4(0, 0, 0, 0), d1 = d0, d2 = d0;
            if (n_items > 0) { d0 = prog[0]; d1 = prog[1]; d2 = prog[2]; }
            for (int n = 0; n < n_items; ++n) {
                const int s = n % SW_STAGES;
                const unsigned ph = (unsigned)(n / SW_STAGES) & 1u;
                mbar_wait(bar_empty + 8 * s, ph ^ 1u);                  // all consumer warps released the stage
                const int row = (int)d0.x, flags = (int)d0.z;
                Ds[s * SW_ITEM_U4 + 0] = d0; Ds[s * SW_ITEM_U4 + 1] = d1; Ds[s * SW_ITEM_U4 + 2] = d2;
                mbar_arrive_expect_tx(bar_full + 8 * s, SW_TILE * (unsigned)sizeof(double));
                bulk_g2s(smem_u32(As + s * SW_TILE), stream + (size_t)n * SW_TILE, SW_TILE * (unsigned)sizeof(double), bar_full + 8 * s);
                if (n + 1 < n_items) {                                  // next descriptor while the copy flies
                    const uint4* q = prog + (size_t)(n + 1) * SW_ITEM_U4;
                    d0 = q[0]; d1 = q[1]; d2 = q[2];
                }
                if (flags & SW_DIAG) {
                    mbar_wait(bar_bempty, ((unsigned)ndiag & 1u) ^ 1u);
                    mbar_arrive_expect_tx(bar_bfull, SW_XTILE * (unsigned)sizeof(double));
                    bulk_g2s(smem_u32(Bs), Xslab + (size_t)row * SW_XTILE, SW_XTILE * (unsigned)sizeof(double), bar_bfull);
                    ++ndiag;
                }
            }
        }
        return;
    }

    // ----------------------------------- consumer warps -----------------------------------
    const int fr = lane >> 2, fk = lane & 3;
    const int p = warp & 3, h = warp >> 2;
    const int mb0 = p, mb1 = 7 - p;               // row blocks of this warp (warps w and w+4 share a scheduler and the row pair)
    const int nt0 = 2 * h;                        // column blocks nt0, nt0 + 1
    // known rows of a backward sweep that starts below the top (second chain: separator solution): row-major -> ring
    for (int q = 0; q < npre; ++q) {
        const int i = pre_row + q, slot = (ktop - i) % SW_RING;
        const double* g = Xslab + (size_t)i * SW_XTILE;
        double* dst = Xr + slot * SW_XTILE;
        for (int e = tid; e < SW_XTILE; e += SW_CONSUMERS) dst[sw_x_index(e / SLAB, e % SLAB)] = g[e];
        mbar_arrive(bar_x + 8 * slot);
    }
    // this lane's elements of a 64 x 32 tile: rows 8*mb + fr, columns 8*nt + 2*fk + {0, 1}
    auto rm_off = [&](int mb, int nt) { return (8 * mb + fr) * SLAB + 8 * nt + 2 * fk; };                       // row-major, double2
    auto fx_off = [&](int mb, int nt, int e) { return ((2 * mb + (fr >> 2)) * 4 + nt) * 32 + (2 * fk + e) * 4 + (fr & 3); };   // fragment order

    double acc[2][2][2], rhs[2][2][2];
    {   // right-hand side rows of the first tile row (forward sweeps)
        const uint4 f0 = prog[0];
        if (n_items > 0 && ((int)f0.z & SW_INIT_RHS)) {
            const double* g = Xslab + (size_t)(int)f0.x * SW_XTILE;
#pragma unroll
            for (int a = 0; a < 2; ++a)
#pragma unroll
                for (int b = 0; b < 2; ++b) {
                    const double2 v = *reinterpret_cast<const double2*>(g + rm_off(a ? mb1 : mb0, nt0 + b));
                    rhs[a][b][0] = v.x; rhs[a][b][1] = v.y;
                }
        }
    }
    consumer_bar_sync();    // nobody stores a row before every consumer holds its first right-hand side
    int ndiag = 0;
    const bool profiling = prof != nullptr && blockIdx.x == 0;
    long long pc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, t0 = 0, t1 = 0;
    const long long t_begin = profiling ? clock64() : 0;
    for (int n = 0; n < n_items; ++n) {
        const int s = n % SW_STAGES;
        if (profiling) t0 = clock64();
        mbar_wait(bar_full + 8 * s, (unsigned)(n / SW_STAGES) & 1u);
        if (profiling) { t1 = clock64(); pc[0] += t1 - t0; }
        const uint4 d0 = Ds[s * SW_ITEM_U4], d1 = Ds[s * SW_ITEM_U4 + 1], mk = Ds[s * SW_ITEM_U4 + 2];
        const int row = (int)d0.x, flags = (int)d0.z, xinfo = (int)d0.w;
        if (flags & SW_ROW_BEGIN) {
#pragma unroll
            for (int a = 0; a < 2; ++a)
#pragma unroll
                for (int b = 0; b < 2; ++b) {
                    acc[a][b][0] = (flags & SW_INIT_RHS) ? rhs[a][b][0] : 0.0;
                    acc[a][b][1] = (flags & SW_INIT_RHS) ? rhs[a][b][1] : 0.0;
                }
        }
        const double* xb = Bs;
        if (flags & SW_DIAG) {
            mbar_wait(bar_bfull, (unsigned)ndiag & 1u);
        } else if (!(flags & SW_NO_OPERAND)) {
            const int slot = xinfo & 0xff;
            mbar_wait(bar_x + 8 * slot, (unsigned)(xinfo >> 8) & 1u);
            xb = Xr + slot * SW_XTILE;
        }
        if (profiling) { t0 = clock64(); pc[1] += t0 - t1; }
        // masks of this warp's two row blocks
        const unsigned w0 = mb0 < 2 ? mk.x : (mb0 < 4 ? mk.y : (mb0 < 6 ? mk.z : mk.w));
        const unsigned w1 = mb1 < 2 ? mk.x : (mb1 < 4 ? mk.y : (mb1 < 6 ? mk.z : mk.w));
        const unsigned m0 = (w0 >> (16 * (mb0 & 1))) & 0xffffu, m1 = (w1 >> (16 * (mb1 & 1))) & 0xffffu;
        const double* a0p = As + s * SW_TILE + (mb0 * 16) * 32 + lane;
        const double* a1p = As + s * SW_TILE + (mb1 * 16) * 32 + lane;
        const double* bp = xb + nt0 * 32 + lane;
        if (!(flags & SW_NO_OPERAND)) {
            // warp-uniform dispatch: dense row blocks run the unrolled, register-double-buffered loop; sparse ones the
            // unrolled loop with one uniform branch per k-group
            if (m0 == m1) {
                if (m0 == 0xffffu) sweep_mma_full<true, true>(acc, a0p, a1p, bp);
                else if (m0) sweep_mma_masked<true, true>(acc, a0p, a1p, bp, m0);
            } else {
                if (m0 == 0xffffu) sweep_mma_full<true, false>(acc, a0p, a1p, bp);
                else if (m0) sweep_mma_masked<true, false>(acc, a0p, a1p, bp, m0);
                if (m1 == 0xffffu) sweep_mma_full<false, true>(acc, a0p, a1p, bp);
                else if (m1) sweep_mma_masked<false, true>(acc, a0p, a1p, bp, m1);
            }
        }
        __syncwarp();
        if (profiling) { t1 = clock64(); pc[2] += t1 - t0; pc[4] += __popc(m0) + __popc(m1); pc[5] += 1; }
        if (lane == 0) {
            mbar_arrive(bar_empty + 8 * s);
            if (flags & SW_DIAG) mbar_arrive(bar_bempty);
        }
        if (flags & SW_DIAG) ++ndiag;
        if (flags & SW_ROW_END) {
            // right-hand side of the next row first: once this row is published another warp may overwrite that block
            if ((int)d1.y) {
                const double* g = Xslab + (size_t)(int)d1.x * SW_XTILE;
#pragma unroll
                for (int a = 0; a < 2; ++a)
#pragma unroll
                    for (int b = 0; b < 2; ++b) {
                        const double2 v = *reinterpret_cast<const double2*>(g + rm_off(a ? mb1 : mb0, nt0 + b));
                        rhs[a][b][0] = v.x; rhs[a][b][1] = v.y;
                    }
            }
            double* g = Xslab + (size_t)row * SW_XTILE;
            const int oslot = (xinfo >> 16) & 0xff;
            double* xr = Xr + oslot * SW_XTILE;
#pragma unroll
            for (int a = 0; a < 2; ++a)
#pragma unroll
                for (int b = 0; b < 2; ++b) {
                    const int mb = a ? mb1 : mb0, nt = nt0 + b;
                    if (flags & SW_OUT_FRAG) { g[fx_off(mb, nt, 0)] = acc[a][b][0]; g[fx_off(mb, nt, 1)] = acc[a][b][1]; }
                    else *reinterpret_cast<double2*>(g + rm_off(mb, nt)) = make_double2(acc[a][b][0], acc[a][b][1]);
                    if (!(flags & SW_NO_RING)) { xr[fx_off(mb, nt, 0)] = acc[a][b][0]; xr[fx_off(mb, nt, 1)] = acc[a][b][1]; }
                }
            if (!(flags & SW_NO_RING)) mbar_arrive(bar_x + 8 * oslot);
            if (profiling) { t0 = clock64(); pc[3] += t0 - t1; }
        }
    }
    if (profiling && lane == 0) {
        pc[6] = clock64() - t_begin;
        for (int i = 0; i < 8; ++i) prof[warp * 8 + i] = pc[i];
    }
}

}  // namespace jk
