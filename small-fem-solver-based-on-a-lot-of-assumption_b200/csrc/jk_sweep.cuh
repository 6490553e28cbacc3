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

// Measured and dropped (tools/ab_bench.sh, c4 workload): descriptors read from global by the consumers (-1 %), a test_wait
// probe of the next tile before the products (-3 %), a find-first-set walk of sparse masks (-2 %), 8 warps with 4 x 1 row
// blocks (-4 %), one row block x two column blocks per warp.  Kept: four consumer warps per 8-column block (16 for a whole
// slab), each owning the row-block pair (s, 7 - s) so that triangular masks balance, and 4 stages.
//
// A CTA owns NCB of the four 8-column blocks of a slab (template parameter): 4 = the whole slab (16 consumer warps), 2 or 1
// when there are fewer slabs than SMs (few phases: the sweeps are then a latency chain per CTA and more, narrower CTAs fill
// the GPU).  CTAs that share a slab never touch each other's data: right-hand sides / solutions are row-major in X (own
// columns), the forward sweep's intermediate Z lives in a SEPARATE array in fragment order (own column-block groups).
#ifndef JK_SW_STAGES
#define JK_SW_STAGES 3      // 32 KB tile stages of the TMA ring.  4 measure 1 % faster for the sweeps alone, but with 3 a sweep CTA leaves
                            // 35 KB of the SM's shared memory free: two member-post blocks (8.5 KB) run beside it (post 0.64 -> 0.56 ms)
#endif
constexpr int SW_STAGES = JK_SW_STAGES;
constexpr int SW_RING = 5;                  // solved tiles kept in shared memory: tile half-bandwidth <= SW_RING - 1
constexpr int SW_MAX_BW = SW_RING - 1;
constexpr int SW_MAX_CONSUMER_WARPS = 16;
#ifndef JK_SW_CTA_GROUPS
#define JK_SW_CTA_GROUPS 1      // CTA groups (blockIdx % groups) whose producers start JK_SW_CTA_STAGGER clocks apart.  MEASURED at c4: 4 / 8 / 16 groups give
                                // 5.199 / 5.212 / 5.254 ms per step against 5.198 -- the L2 slices holding the current tile are NOT what the sweeps wait for
#endif
#ifndef JK_SW_CTA_STAGGER
#define JK_SW_CTA_STAGGER 2000
#endif
#ifndef JK_SW_GROUP_BARS
#define JK_SW_GROUP_BARS 1  // row barriers per column group (4 warps) instead of per CTA.  A/B at c4: 5.209 vs 5.224 ms per step (noise level); kept: groups never wait for each other's rows
#endif
#ifndef JK_SW_STAGGER
#define JK_SW_STAGGER 0     // clocks between the starts of the column groups of a CTA.  MEASURED: 300 / 600 / 1200 change nothing at c4 (5.227 / 5.229 / 5.230 vs
                            // 5.209 ms per step) -- the groups are not held in lockstep by their bookkeeping, so there is nothing to hide
#endif
#ifndef JK_SW_ZPREFETCH
#define JK_SW_ZPREFETCH 0   // backward sweeps: L2 prefetch of the Z rows this many tile rows ahead of the diagonal item.  MEASURED with 3: no change at c4 (1.50 ms)
                            // nor at c5 (2 GB of Z: 2.91 vs 2.93 ms) -- the Z stage is not what the backward sweeps wait for
#endif
#ifndef JK_SW_CB_MINOR
#define JK_SW_CB_MINOR 0      // column group as the minor index of the warp id (each scheduler hosts all four roles of one column group).
                            // MEASURED SLOWER at c4: forward 1.17 vs 1.12 ms, backward equal -- see the comment at the mapping
#endif
#ifndef JK_SW_PAIR
#define JK_SW_PAIR 0          // paired fragment order (two k-groups per 16-byte shared-memory load: half the load instructions of the DMMA loops).
                            // MEASURED SLOWER at c4: forward 1.170 vs 1.118 ms, backward 1.48 vs 1.43 ms (parity suite green on both)
#endif
#ifndef JK_SW_ROLE_MASKS
#define JK_SW_ROLE_MASKS 1    // zero-block masks stored per row-block pair (s, 7 - s): one 32-bit load per warp and item (5.127 vs 5.144 ms per step at c4)
#endif
#ifndef JK_SW_MASKED
#define JK_SW_MASKED 0        // row blocks with different non-empty masks in ONE interleaved loop (B fragment loaded once) instead of one after the
                            // other: no measurable difference at c4 (5.133 vs 5.127 ms per step)
#endif
#ifndef JK_SW_PRED
#define JK_SW_PRED 0        // sparse items: 1 = unconditional double-buffered fragment loads + predicated DMMAs, 2 = only the split chain for a
                            // single active row block.  BOTH MEASURED SLOWER at c4 (forward sweeps 1.32 / 1.29 vs 1.20 ms): every extra
                            // shared-memory fragment load costs more than the latency chain it removes.
#endif
#ifndef JK_SW_CBN
#define JK_SW_CBN 1         // 8-column blocks per consumer warp.  2 (2 x 2 register blocking: half the warps, 2/3 of the shared-memory loads; whole-slab
                            // CTAs only) MEASURED SLOWER at c4: forward 1.25 vs 1.17 ms, backward 1.59 vs 1.47 ms -- the shared-memory pipe is not the bound
#endif
__host__ __device__ constexpr int sw_threads(int ncb) { return 32 * (4 * ncb / JK_SW_CBN) + 32; }   // consumer warps + the producer warp
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
constexpr int SW_WAIT_X = 128;              // first use of the operand row by this sweep: wait on its mbarrier (later uses need not)

// One item = 3 x uint4 in the program array:
//   [0] = {row, src, flags, xinfo}   xinfo: bits 0-7 operand ring slot, bit 8 its mbarrier parity, bits 16-23 output slot
//   [1] = {next_row, next_init, 0, 0}   (ROW_BEGIN items of forward sweeps: right-hand side rows to prefetch for the next row)
//   [2] = 8 x 16-bit k-group masks (written by k_sweep_build)
constexpr int SW_ITEM_U4 = 3;

constexpr size_t SW_SMEM = (size_t)(SW_STAGES * SW_TILE + SW_XTILE + SW_RING * SW_XTILE) * sizeof(double)
                         + (size_t)SW_STAGES * SW_ITEM_U4 * sizeof(uint4) + 40 * sizeof(unsigned long long) + 128;

#if JK_SW_PAIR
// Paired fragment order: the fragments of two consecutive k-groups are interleaved, so that one 16-byte shared-memory load
// brings a lane's elements of both (half the load instructions of the DMMA loops; same bytes, still conflict-free):
//   A tile:  ((r/8)*8 + c/8)*64 + ((r%8)*4 + c%4)*2 + (c/4)%2        X tile:  ((r/8)*4 + c/8)*64 + ((c%8)*4 + r%4)*2 + (r/4)%2
__host__ __device__ __forceinline__ int sw_a_index(int r, int c) { return ((((r >> 3) * 8 + (c >> 3)) * 32 + (r & 7) * 4 + (c & 3)) << 1) | ((c >> 2) & 1); }
__host__ __device__ __forceinline__ int sw_x_index(int r, int c) { return ((((r >> 3) * 4 + (c >> 3)) * 32 + (c & 7) * 4 + (r & 3)) << 1) | ((r >> 2) & 1); }
#define SW_A_OFF(k4) ((((k4) >> 1) * 64) + ((k4) & 1))      /* this lane's element of k-group k4, relative to its row block's pointer */
#define SW_B_OFF(k4) ((((k4) >> 1) * 256) + ((k4) & 1))
constexpr int SW_LANE_STRIDE = 2, SW_B_CB = 64;             // doubles per lane in a fragment pair / per 8-column block in a k-group pair
#else
__host__ __device__ __forceinline__ int sw_a_index(int r, int c) { return ((r >> 3) * 16 + (c >> 2)) * 32 + (r & 7) * 4 + (c & 3); }
__host__ __device__ __forceinline__ int sw_x_index(int r, int c) { return ((r >> 2) * 4 + (c >> 3)) * 32 + (c & 7) * 4 + (r & 3); }
#define SW_A_OFF(k4) ((k4) * 32)
#define SW_B_OFF(k4) ((k4) * 128)
constexpr int SW_LANE_STRIDE = 1, SW_B_CB = 32;
#endif

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
__device__ __forceinline__ bool mbar_test(unsigned bar, unsigned parity) {      // non-blocking probe
    unsigned ok;
    asm volatile("{\n .reg .pred p;\n mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
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
__device__ __forceinline__ void bulk_prefetch_l2(const void* src, unsigned bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;\n" :: "l"(src), "r"(bytes) : "memory");
}
template <int CONSUMERS>
__device__ __forceinline__ void consumer_bar_sync() { asm volatile("bar.sync 1, %0;\n" :: "n"(CONSUMERS) : "memory"); }

// ----------------------------------------------------------------------------------------------
// K3d: build the tile stream of one sweep program.  One CTA per item:
//   forward  item (row k, src j):  T = -L_kj Linv_jj
//   backward diagonal (row k):     T =  Linv_kk^T Linv_kk
//   backward item (row k, src i):  T = -Linv_kk^T L_ik^T
// T is written in A-fragment order; the item's masks mark the (8-row block, 4-column group) cells with a non-zero.
// Structural zeros of L are exact zeros (the envelope is never filled), so the test is exact.
// ----------------------------------------------------------------------------------------------
constexpr int SWB_KC = 16;                 // K chunk: small shared-memory footprint, so build CTAs fit beside a sweep CTA on an SM
constexpr int SWB_PLD = SWB_KC + 1, SWB_QLD = NB + 1;
constexpr size_t SWB_SMEM = (size_t)(NB * SWB_PLD + SWB_KC * SWB_QLD) * sizeof(double);
struct SweepBuildArgs { uint4* prog; double* stream; const double* tiles; const double* Linv; int bw, n_items; };
// blocks [0, a.n_items) build the stream of program a, the rest that of program b (both chains in one wave)
__global__ void __launch_bounds__(256) k_sweep_build(SweepBuildArgs a, SweepBuildArgs b, int backward) {
    extern __shared__ __align__(16) double smem[];
    double* Ps = smem;                    // P[r][m - m0]   (NB x SWB_KC)
    double* Qs = smem + NB * SWB_PLD;     // Q[m - m0][c]   (SWB_KC x NB)
    __shared__ unsigned msk[8];
    int n = blockIdx.x;
    const int tid = threadIdx.x;
    if (n >= a.n_items) { n -= a.n_items; a = b; }
    uint4* __restrict__ prog = a.prog;
    double* __restrict__ stream = a.stream;
    const double* __restrict__ tiles = a.tiles;
    const double* __restrict__ Linv = a.Linv;
    const int bw = a.bw;
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
    const int ty = tid / 16, tx = tid % 16;
    double acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
    for (int m0 = 0; m0 < NB; m0 += SWB_KC) {
        __syncthreads();
        // P[r][m] = tp ? ps[m][r] : ps[r][m];  Q[m][c] = tq ? qs[c][m] : qs[m][c]   for m in [m0, m0 + KC)
        for (int idx = tid; idx < NB * SWB_KC; idx += 256) {
            int r, m;
            if (tp) { m = idx / NB; r = idx % NB; Ps[r * SWB_PLD + m] = ps[(m0 + m) * NB + r]; }
            else { r = idx / SWB_KC; m = idx % SWB_KC; Ps[r * SWB_PLD + m] = ps[r * NB + m0 + m]; }
            int c;
            if (tq) { c = idx / SWB_KC; m = idx % SWB_KC; Qs[m * SWB_QLD + c] = qs[c * NB + m0 + m]; }
            else { m = idx / NB; c = idx % NB; Qs[m * SWB_QLD + c] = qs[(m0 + m) * NB + c]; }
        }
        __syncthreads();
#pragma unroll 4
        for (int m = 0; m < SWB_KC; ++m) {
            double p[4], q[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) p[i] = Ps[(4 * ty + i) * SWB_PLD + m];
#pragma unroll
            for (int j = 0; j < 4; ++j) q[j] = Qs[m * SWB_QLD + tx + 16 * j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fma(p[i], q[j], acc[i][j]);
        }
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
#if JK_SW_ROLE_MASKS
        // word s = masks of the row-block pair (s, 7 - s) a consumer warp owns: one 32-bit load, no selects
        prog[(size_t)n * SW_ITEM_U4 + 2] = make_uint4(msk[0] | (msk[7] << 16), msk[1] | (msk[6] << 16), msk[2] | (msk[5] << 16), msk[3] | (msk[4] << 16));
#else
        prog[(size_t)n * SW_ITEM_U4 + 2] = make_uint4(msk[0] | (msk[1] << 16), msk[2] | (msk[3] << 16), msk[4] | (msk[5] << 16), msk[6] | (msk[7] << 16));
#endif
}

// Inner products of one item for one consumer warp: acc[a][b] += A(row block a) * X(column block b) over the k-groups.
// ap[a] / bp point at this lane's element of the first fragment; all further offsets are compile-time constants.
constexpr int SW_RBN = 2, SW_CBN = JK_SW_CBN;     // per consumer warp: the row-block pair (s, 7 - s) x one 8-column block

// every row block of the warp is either dense or empty (act[a], warp-uniform): register-double-buffered, fully unrolled
template <bool ALL>
__device__ __forceinline__ void sweep_mma_dense(double (&acc)[SW_RBN][SW_CBN][2], const double* const (&ap)[SW_RBN], const bool (&act)[SW_RBN],
                                                const double* __restrict__ bp) {
    double af[2][SW_RBN], bf[2][SW_CBN];
#pragma unroll
    for (int a = 0; a < SW_RBN; ++a) { af[0][a] = 0.0; af[1][a] = 0.0; if (ALL || act[a]) af[0][a] = ap[a][0]; }
#pragma unroll
    for (int b = 0; b < SW_CBN; ++b) bf[0][b] = bp[b * SW_B_CB];
#pragma unroll
    for (int k4 = 0; k4 < 16; ++k4) {
        const int cur = k4 & 1, nxt = cur ^ 1;
        if (k4 + 1 < 16) {      // fragments of the next k-group before the DMMAs of this one
#pragma unroll
            for (int a = 0; a < SW_RBN; ++a) if (ALL || act[a]) af[nxt][a] = ap[a][SW_A_OFF(k4 + 1)];
#pragma unroll
            for (int b = 0; b < SW_CBN; ++b) bf[nxt][b] = bp[SW_B_OFF(k4 + 1) + b * SW_B_CB];
        }
#pragma unroll
        for (int a = 0; a < SW_RBN; ++a)
            if (ALL || act[a]) {
#pragma unroll
                for (int b = 0; b < SW_CBN; ++b) dmma(acc[a][b][0], acc[a][b][1], af[cur][a], bf[cur][b]);
            }
    }
}
#if JK_SW_PAIR
// paired layout, both row blocks dense: three 16-byte loads feed four DMMAs (k-groups 2q and 2q + 1 of both row blocks)
__device__ __forceinline__ void sweep_mma_dense_pair(double (&acc)[SW_RBN][SW_CBN][2], const double* const (&ap)[SW_RBN], const double* __restrict__ bp) {
    static_assert(SW_RBN == 2 && SW_CBN == 1, "paired dense loop");
    double2 a0[2], a1[2], b[2];
    a0[0] = *reinterpret_cast<const double2*>(ap[0]); a1[0] = *reinterpret_cast<const double2*>(ap[1]); b[0] = *reinterpret_cast<const double2*>(bp);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const int cur = q & 1, nxt = cur ^ 1;
        if (q + 1 < 8) {
            a0[nxt] = *reinterpret_cast<const double2*>(ap[0] + (q + 1) * 64);
            a1[nxt] = *reinterpret_cast<const double2*>(ap[1] + (q + 1) * 64);
            b[nxt] = *reinterpret_cast<const double2*>(bp + (q + 1) * 256);
        }
        dmma(acc[0][0][0], acc[0][0][1], a0[cur].x, b[cur].x);
        dmma(acc[1][0][0], acc[1][0][1], a1[cur].x, b[cur].x);
        dmma(acc[0][0][0], acc[0][0][1], a0[cur].y, b[cur].y);
        dmma(acc[1][0][0], acc[1][0][1], a1[cur].y, b[cur].y);
    }
}
// one dense row block on its own: two 16-byte loads feed two DMMAs
template <int A>
__device__ __forceinline__ void sweep_mma_single_dense_pair(double (&acc)[SW_RBN][SW_CBN][2], const double* __restrict__ ap, const double* __restrict__ bp) {
    double2 a[2], b[2];
    a[0] = *reinterpret_cast<const double2*>(ap); b[0] = *reinterpret_cast<const double2*>(bp);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const int cur = q & 1, nxt = cur ^ 1;
        if (q + 1 < 8) { a[nxt] = *reinterpret_cast<const double2*>(ap + (q + 1) * 64); b[nxt] = *reinterpret_cast<const double2*>(bp + (q + 1) * 256); }
        dmma(acc[A][0][0], acc[A][0][1], a[cur].x, b[cur].x);
        dmma(acc[A][0][0], acc[A][0][1], a[cur].y, b[cur].y);
    }
}
#endif
// general masks: unrolled, one warp-uniform branch per k-group and row block
__device__ __forceinline__ void sweep_mma_masked(double (&acc)[SW_RBN][SW_CBN][2], const double* const (&ap)[SW_RBN], const unsigned (&m)[SW_RBN],
                                                 const double* __restrict__ bp) {
    unsigned mu = 0u;
#pragma unroll
    for (int a = 0; a < SW_RBN; ++a) mu |= m[a];
#pragma unroll
    for (int k4 = 0; k4 < 16; ++k4) {
        if (mu & (1u << k4)) {
            double bf[SW_CBN];
#pragma unroll
            for (int b = 0; b < SW_CBN; ++b) bf[b] = bp[SW_B_OFF(k4) + b * SW_B_CB];
#pragma unroll
            for (int a = 0; a < SW_RBN; ++a)
                if (m[a] & (1u << k4)) {
                    const double af = ap[a][SW_A_OFF(k4)];
#pragma unroll
                    for (int b = 0; b < SW_CBN; ++b) dmma(acc[a][b][0], acc[a][b][1], af, bf[b]);
                }
        }
    }
}

// all row blocks of the warp share one sparse mask (backward tiles: the mask is a column pattern): one branch per k-group
__device__ __forceinline__ void sweep_mma_uniform(double (&acc)[SW_RBN][SW_CBN][2], const double* const (&ap)[SW_RBN], unsigned mask,
                                                  const double* __restrict__ bp) {
#pragma unroll
    for (int k4 = 0; k4 < 16; ++k4) {
        if (mask & (1u << k4)) {
            double af[SW_RBN], bf[SW_CBN];
#pragma unroll
            for (int a = 0; a < SW_RBN; ++a) af[a] = ap[a][SW_A_OFF(k4)];
#pragma unroll
            for (int b = 0; b < SW_CBN; ++b) bf[b] = bp[SW_B_OFF(k4) + b * SW_B_CB];
#pragma unroll
            for (int a = 0; a < SW_RBN; ++a)
#pragma unroll
                for (int b = 0; b < SW_CBN; ++b) dmma(acc[a][b][0], acc[a][b][1], af[a], bf[b]);
        }
    }
}
// one row block on its own (masks differ between the warp's row blocks): dense -> double-buffered, sparse -> branch per k-group
template <int A>
__device__ __forceinline__ void sweep_mma_single(double (&acc)[SW_RBN][SW_CBN][2], const double* __restrict__ ap, unsigned mask,
                                                 const double* __restrict__ bp) {
#if JK_SW_PAIR
    if (mask == 0xffffu) { sweep_mma_single_dense_pair<A>(acc, ap, bp); return; }
#endif
    if (mask == 0xffffu) {
        double af[2], bf[2][SW_CBN];
        af[0] = ap[0];
#pragma unroll
        for (int b = 0; b < SW_CBN; ++b) bf[0][b] = bp[b * SW_B_CB];
#pragma unroll
        for (int k4 = 0; k4 < 16; ++k4) {
            const int cur = k4 & 1, nxt = cur ^ 1;
            if (k4 + 1 < 16) {
                af[nxt] = ap[SW_A_OFF(k4 + 1)];
#pragma unroll
                for (int b = 0; b < SW_CBN; ++b) bf[nxt][b] = bp[SW_B_OFF(k4 + 1) + b * SW_B_CB];
            }
#pragma unroll
            for (int b = 0; b < SW_CBN; ++b) dmma(acc[A][b][0], acc[A][b][1], af[cur], bf[cur][b]);
        }
    } else if (mask) {
#pragma unroll
        for (int k4 = 0; k4 < 16; ++k4) {
            if (mask & (1u << k4)) {
                const double af = ap[SW_A_OFF(k4)];
#pragma unroll
                for (int b = 0; b < SW_CBN; ++b) dmma(acc[A][b][0], acc[A][b][1], af, bp[SW_B_OFF(k4) + b * SW_B_CB]);
            }
        }
    }
}

// Sparse items (JK_SW_PRED = 1, measured and dropped -- see the switch).  The paths above put a shared-memory load in front of every DMMA inside a warp-uniform branch
// and run the two row blocks of a warp one after the other when their masks differ: a sparse item then costs a chain of 16-32
// load + DMMA latencies (~850 clocks measured, option profile_sweep = 2) for a handful of DMMAs, as much as a dense item.
// Here the fragments of EVERY k-group are loaded (the tile is complete in shared memory, masked cells are zeros) in the
// register-double-buffered pattern of the dense loop and only the DMMAs are predicated, so the two row blocks' chains
// interleave and no load waits on a branch.
__device__ __forceinline__ void sweep_mma_pred(double (&acc)[SW_RBN][SW_CBN][2], const double* const (&ap)[SW_RBN], const unsigned (&m)[SW_RBN],
                                               const double* __restrict__ bp) {
    double af[2][SW_RBN], bf[2][SW_CBN];
#pragma unroll
    for (int a = 0; a < SW_RBN; ++a) af[0][a] = ap[a][0];
#pragma unroll
    for (int b = 0; b < SW_CBN; ++b) bf[0][b] = bp[b * SW_B_CB];
#pragma unroll
    for (int k4 = 0; k4 < 16; ++k4) {
        const int cur = k4 & 1, nxt = cur ^ 1;
        if (k4 + 1 < 16) {
#pragma unroll
            for (int a = 0; a < SW_RBN; ++a) af[nxt][a] = ap[a][SW_A_OFF(k4 + 1)];
#pragma unroll
            for (int b = 0; b < SW_CBN; ++b) bf[nxt][b] = bp[SW_B_OFF(k4 + 1) + b * SW_B_CB];
        }
#pragma unroll
        for (int a = 0; a < SW_RBN; ++a)
            if ((m[a] >> k4) & 1u) {
#pragma unroll
                for (int b = 0; b < SW_CBN; ++b) dmma(acc[a][b][0], acc[a][b][1], af[cur][a], bf[cur][b]);
            }
    }
}
// only row block A of the warp holds non-zeros: its k-groups alternate between the accumulator and a second partial sum, which
// halves the dependent DMMA chain (the only chain this warp has in the item)
template <int A>
__device__ __forceinline__ void sweep_mma_one(double (&acc)[SW_RBN][SW_CBN][2], const double* __restrict__ ap, unsigned mask,
                                              const double* __restrict__ bp) {
    double af[2], bf[2][SW_CBN], part[SW_CBN][2];
#pragma unroll
    for (int b = 0; b < SW_CBN; ++b) { part[b][0] = 0.0; part[b][1] = 0.0; bf[0][b] = bp[b * SW_B_CB]; }
    af[0] = ap[0];
#pragma unroll
    for (int k4 = 0; k4 < 16; ++k4) {
        const int cur = k4 & 1, nxt = cur ^ 1;
        if (k4 + 1 < 16) {
            af[nxt] = ap[SW_A_OFF(k4 + 1)];
#pragma unroll
            for (int b = 0; b < SW_CBN; ++b) bf[nxt][b] = bp[SW_B_OFF(k4 + 1) + b * SW_B_CB];
        }
        if ((mask >> k4) & 1u) {
#pragma unroll
            for (int b = 0; b < SW_CBN; ++b) {
                if (k4 & 1) dmma(part[b][0], part[b][1], af[cur], bf[cur][b]);
                else dmma(acc[A][b][0], acc[A][b][1], af[cur], bf[cur][b]);
            }
        }
    }
#pragma unroll
    for (int b = 0; b < SW_CBN; ++b) { acc[A][b][0] += part[b][0]; acc[A][b][1] += part[b][1]; }
}

// ----------------------------------------------------------------------------------------------
// K4 (narrow bands): one sweep of one chain over one slab per CTA
// ----------------------------------------------------------------------------------------------
template <int NCB /* 8-column blocks of the slab owned by this CTA: 4, 2 or 1 */,
          bool PROF = false /* clock counters / stamps of option profile_sweep compiled in (the production kernel carries none of them:
                               every bookkeeping instruction between two items' DMMAs is warp latency the FP64 pipe sits out) */>
__global__ void __launch_bounds__(sw_threads(NCB), 1)
k_sweep(const uint4* __restrict__ prog, const double* __restrict__ stream, double* __restrict__ X,
        double* __restrict__ Z /* same shape as X: the forward sweep's intermediate, fragment order */,
        int n_items, int n_pad /* rows of the whole slab */, int row0 /* first row of this chain in the slab */,
        int pre_row /* first known tile row */, int npre /* known tile rows to preload into the ring */, int ktop,
        int pre_mode /* 0: backward, rows hold X row-major (second chain's separator solution), ring slot (ktop - row) % R;
                        1: forward continuation, rows hold Z in fragment order, ring slot row % R */,
        int xphase_bits /* bit s: ring slot s starts one mbarrier phase ahead (continuation of a program whose slot parities
                           count from its first row) */,
        long long* __restrict__ prof /* nullable: [8 warps][8] clock sums of CTA 0 (option profile_sweep) */,
        unsigned* __restrict__ started = nullptr /* nullable: every CTA adds 1 as soon as it is resident (gate of the early member post) */,
        long long* __restrict__ trace = nullptr /* nullable: [items][2 warps][4] clock stamps of CTA 0 (option profile_sweep = 2) */) {
    constexpr int SW_CONSUMER_WARPS = 4 * NCB / SW_CBN, SW_CONSUMERS = 32 * SW_CONSUMER_WARPS, CTAS_PER_SLAB = 4 / NCB;
    extern __shared__ __align__(128) unsigned char sw_smem[];
    if (started != nullptr && threadIdx.x == 0) { atomicAdd(started, 1u); __threadfence(); }
    double* As = reinterpret_cast<double*>(sw_smem);                 // [SW_STAGES][SW_TILE]   A tiles, fragment order
    double* Bs = As + SW_STAGES * SW_TILE;                           // [SW_XTILE]             Z_k of the backward diagonal item
    double* Xr = Bs + SW_XTILE;                                      // [SW_RING][SW_XTILE]    newest solved tiles, fragment order
    uint4* Ds = reinterpret_cast<uint4*>(Xr + SW_RING * SW_XTILE);   // [SW_STAGES][SW_ITEM_U4] item descriptors, staged by the producer
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(Ds + SW_STAGES * SW_ITEM_U4);
    // row barriers: one per ring slot and COLUMN GROUP (JK_SW_GROUP_BARS): a warp only ever reads its own 8-column block of a solved
    // row, written by the warps of its group, so the groups need not wait for each other at a row end -- they stay coupled only
    // through the tile stages, and can run out of phase (see the stagger below)
    constexpr int XG = JK_SW_GROUP_BARS ? NCB : 1, X_ARRIVALS = JK_SW_GROUP_BARS ? SW_CONSUMER_WARPS / NCB : SW_CONSUMER_WARPS;
    static_assert(2 * SW_STAGES + SW_RING * XG + 2 <= 40, "barrier area");
    const unsigned bar_full = smem_u32(bars), bar_empty = smem_u32(bars + SW_STAGES), bar_x0 = smem_u32(bars + 2 * SW_STAGES),
                   bar_bfull = smem_u32(bars + 2 * SW_STAGES + SW_RING * XG), bar_bempty = smem_u32(bars + 2 * SW_STAGES + SW_RING * XG + 1);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int slab = blockIdx.x / CTAS_PER_SLAB, cbase = (blockIdx.x % CTAS_PER_SLAB) * NCB;
    double* Xslab = X + ((size_t)slab * (size_t)n_pad + (size_t)row0) * SLAB;
    double* Zslab = Z + ((size_t)slab * (size_t)n_pad + (size_t)row0) * SLAB;

    if (tid == 0) {
        for (int s = 0; s < SW_STAGES; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, SW_CONSUMER_WARPS); }
        for (int s = 0; s < SW_RING * XG; ++s) mbar_init(bar_x0 + 8 * s, X_ARRIVALS);
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
#if JK_SW_CTA_GROUPS > 1
            // All CTAs of a sweep stream the SAME tiles in the same order.  Started together they ask the L2 for the same 32 KB at
            // the same time: those few slices serve every SM while the others idle, and a tile takes as long to arrive as its
            // products take to compute, whatever the masks skip.  CTA groups that start a tile-time apart read different tiles
            // at any moment, which spreads the stream over all slices.
            { const long long t_go = clock64() + (long long)(blockIdx.x % JK_SW_CTA_GROUPS) * JK_SW_CTA_STAGGER; while (clock64() < t_go) { } }
#endif
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
#if JK_SW_ZPREFETCH
                    // Z_k is staged through ONE buffer, at most SW_STAGES items ahead of its use: with short tile rows and a slab set
                    // larger than L2 (sea-state ensembles: 2 GB of Z) that is less than an HBM round trip.  Pull the rows the sweep
                    // reaches next (backward: row - 1, row - 2, ...) into L2 now.
                    if (cbase == 0) {
                        const int first = ndiag == 0 ? 1 : JK_SW_ZPREFETCH;
                        for (int q = first; q <= JK_SW_ZPREFETCH; ++q)
                            if (row - q >= 0) bulk_prefetch_l2(Zslab + (size_t)(row - q) * SW_XTILE, SW_XTILE * (unsigned)sizeof(double));
                    }
#endif
                    mbar_wait(bar_bempty, ((unsigned)ndiag & 1u) ^ 1u);
                    mbar_arrive_expect_tx(bar_bfull, SW_XTILE * (unsigned)sizeof(double));
                    bulk_g2s(smem_u32(Bs), Zslab + (size_t)row * SW_XTILE, SW_XTILE * (unsigned)sizeof(double), bar_bfull);
                    ++ndiag;
                }
            }
        }
        return;
    }

    // ----------------------------------- consumer warps -----------------------------------
    // Warp w owns SW_RBN row blocks x SW_CBN column blocks of the 64 x 32 result tile.  With 4 x 1 the two warps of a
    // scheduler (w, w + 4) cover all eight row blocks of one column block, so every scheduler has the same DMMA count
    // in every tile row whatever the masks look like (rows end with an all-to-all exchange: imbalance is idle time).
    const int fr = lane >> 2, fk = lane & 3;
    // Which warp does what.  A warp's role is the row-block pair (role, 7 - role); its column group is one of the CTA's 8-column
    // blocks.  Consecutive warp ids sit on different schedulers (SM sub-partitions, each with its own share of the FP64 pipe), so
    // with the column group as the MINOR index (JK_SW_CB_MINOR) a scheduler hosts all four roles of one column group and does
    // a quarter of every item's DMMAs whatever the masks look like; with the role as the minor index a scheduler hosts ONE role
    // for all column groups, and a row-sparse item (the far tiles of the envelope staircase touch only the last row blocks)
    // lands on one scheduler while three idle.  That was the expectation; the measurement says otherwise (forward sweeps 1.17
    // instead of 1.12 ms), so the role stays the minor index: the four warps of a scheduler then share their A fragments'
    // addresses and run the same code path per item.
    constexpr int GROUPS = SW_CONSUMER_WARPS / 4;      // column groups of this CTA (4 warps each)
    const int cgroup = JK_SW_CB_MINOR ? (warp % GROUPS) : (warp >> 2);
    const int role = JK_SW_CB_MINOR ? (warp / GROUPS) : (warp & 3);
    const int cb0 = cbase + cgroup * SW_CBN;
    const int xgroup = JK_SW_GROUP_BARS ? cgroup : 0;
    const unsigned bar_x = bar_x0 + 8 * SW_RING * xgroup;      // this group's row barriers, [slot]
    const int rbs[SW_RBN] = {role, 7 - role};
    // a program that continues another launch: bring the slots' mbarrier phases in step with the item parities
    if (xphase_bits) {
        if (lane == 0)
            for (int sl = 0; sl < SW_RING; ++sl) if ((xphase_bits >> sl) & 1) mbar_arrive(bar_x + 8 * sl);
        consumer_bar_sync<SW_CONSUMERS>();
    }
    // rows solved before this launch that its first rows need: the second chain's separator solution (backward,
    // row-major) or the rows just before a forward continuation (Z, already in fragment order) -> ring
    for (int q = 0; q < npre; ++q) {
        const int i = pre_row + q, slot = (pre_mode ? i : ktop - i) % SW_RING;
        const double* g = (pre_mode ? Zslab : Xslab) + (size_t)i * SW_XTILE;
        double* dst = Xr + slot * SW_XTILE;
        if (pre_mode) { for (int e = tid; e < SW_XTILE; e += SW_CONSUMERS) dst[e] = g[e]; }
        else { for (int e = tid; e < SW_XTILE; e += SW_CONSUMERS) dst[sw_x_index(e / SLAB, e % SLAB)] = g[e]; }
    }
    if (npre > 0) {
        // the rows were written by all consumer threads, whatever their column group
        consumer_bar_sync<SW_CONSUMERS>();
        if (lane == 0)
            for (int q = 0; q < npre; ++q) mbar_arrive(bar_x + 8 * ((pre_mode ? pre_row + q : ktop - (pre_row + q)) % SW_RING));
    }
    // this lane's elements of a 64 x 32 tile: rows 8*mb + fr, columns 8*nt + 2*fk + {0, 1}
    auto rm_off = [&](int mb, int nt) { return (8 * mb + fr) * SLAB + 8 * nt + 2 * fk; };                       // row-major, double2
    auto fx_off = [&](int mb, int nt, int e) { return sw_x_index(8 * mb + fr, 8 * nt + 2 * fk + e); };   // fragment order
    double acc[SW_RBN][SW_CBN][2], rhs[SW_RBN][SW_CBN][2];
    auto load_rhs = [&](int r) {
        const double* g = Xslab + (size_t)r * SW_XTILE;
#pragma unroll
        for (int a = 0; a < SW_RBN; ++a)
#pragma unroll
            for (int b = 0; b < SW_CBN; ++b) {
                const double2 v = *reinterpret_cast<const double2*>(g + rm_off(rbs[a], cb0 + b));
                rhs[a][b][0] = v.x; rhs[a][b][1] = v.y;
            }
    };
    {   // right-hand side rows of the first tile row (forward sweeps)
        const uint4 f0 = prog[0];
        if (n_items > 0 && ((int)f0.z & SW_INIT_RHS)) load_rhs((int)f0.x);
    }
    consumer_bar_sync<SW_CONSUMERS>();    // nobody stores a row before every consumer holds its first right-hand side
    int ndiag = 0;
    const bool profiling = PROF && prof != nullptr && blockIdx.x == 0;
    const int trace_w = (warp == 0) ? 0 : (warp == SW_CONSUMER_WARPS - 3 ? 1 : -1);
    const bool tracing = PROF && trace != nullptr && blockIdx.x == 0 && lane == 0 && trace_w >= 0;
    long long pc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, t0 = 0, t1 = 0;
    const long long t_begin = profiling ? clock64() : 0;
    int s = 0; unsigned sphase = 0u;                 // stage of item n and its mbarrier parity, carried instead of n % / n /
    for (int n = 0; n < n_items; ++n, s = (s + 1 == SW_STAGES) ? 0 : s + 1, sphase ^= (s == 0) ? 1u : 0u) {
        if (profiling) t0 = clock64();
        mbar_wait(bar_full + 8 * s, sphase);
#if JK_SW_STAGGER
        // Identical column groups would run in lockstep: all sixteen warps in their DMMA loops (the FP64 pipe saturated, every loop
        // stretched) and then all in the per-item bookkeeping (barrier round trips, descriptor and mask decode: the pipe idle).
        // A one-off offset between the groups persists (the groups only meet at the tile stages), so one group's bookkeeping
        // hides under the other groups' DMMAs.
        if (n == 0 && xgroup > 0) { const long long t_go = clock64() + (long long)xgroup * JK_SW_STAGGER; while (clock64() < t_go) { } }
#endif
        if (profiling) { t1 = clock64(); pc[0] += t1 - t0; }
        if (tracing) trace[((size_t)n * 2 + trace_w) * 4 + 0] = clock64();
        const uint4 d0 = Ds[s * SW_ITEM_U4], d1 = Ds[s * SW_ITEM_U4 + 1];
#if !JK_SW_ROLE_MASKS
        const uint4 mk = Ds[s * SW_ITEM_U4 + 2];
#endif
        const int row = (int)d0.x, flags = (int)d0.z, xinfo = (int)d0.w;
        if (flags & SW_ROW_BEGIN) {
#pragma unroll
            for (int a = 0; a < SW_RBN; ++a)
#pragma unroll
                for (int b = 0; b < SW_CBN; ++b) {
                    acc[a][b][0] = (flags & SW_INIT_RHS) ? rhs[a][b][0] : 0.0;
                    acc[a][b][1] = (flags & SW_INIT_RHS) ? rhs[a][b][1] : 0.0;
                }
            // right-hand side of the NEXT row, a whole row ahead (HBM latency).  It must be issued before this row is
            // published: after that another warp may overwrite that block with Z in fragment order.
            if ((int)d1.y) load_rhs((int)d1.x);
        }
        if (profiling) { t0 = clock64(); pc[7] += t0 - t1; t1 = t0; }
        const double* xb = Bs;
        if (flags & SW_DIAG) {
            mbar_wait(bar_bfull, (unsigned)ndiag & 1u);
        } else if (!(flags & SW_NO_OPERAND)) {
            const int slot = xinfo & 0xff;
            if (flags & SW_WAIT_X) mbar_wait(bar_x + 8 * slot, (unsigned)(xinfo >> 8) & 1u);
            xb = Xr + slot * SW_XTILE;
        }
        if (profiling) { t0 = clock64(); pc[1] += t0 - t1; }
        if (tracing) trace[((size_t)n * 2 + trace_w) * 4 + 1] = clock64();
        // masks of this warp's row blocks
        unsigned m[SW_RBN];
        bool act[SW_RBN], all_dense = true, dense_or_empty = true, any = false, all_equal = true;
        const double* ap[SW_RBN];
#if JK_SW_ROLE_MASKS
        static_assert(SW_RBN == 2, "role-packed masks assume the row-block pair (s, 7 - s)");
        {
            const unsigned w = reinterpret_cast<const unsigned*>(Ds + s * SW_ITEM_U4 + 2)[role];
            m[0] = w & 0xffffu; m[1] = w >> 16;
            act[0] = m[0] == 0xffffu; act[1] = m[1] == 0xffffu;
            all_dense = w == 0xffffffffu;
            any = w != 0u;
            all_equal = m[0] == m[1];
            dense_or_empty = (act[0] || m[0] == 0u) && (act[1] || m[1] == 0u);
            const double* abase = As + s * SW_TILE + lane * SW_LANE_STRIDE;
            ap[0] = abase + rbs[0] * 512; ap[1] = abase + rbs[1] * 512;
        }
#else
        const unsigned mw[4] = {mk.x, mk.y, mk.z, mk.w};
#pragma unroll
        for (int a = 0; a < SW_RBN; ++a) {
            const int rb = rbs[a];
            const unsigned w = rb < 2 ? mw[0] : (rb < 4 ? mw[1] : (rb < 6 ? mw[2] : mw[3]));
            m[a] = (w >> (16 * (rb & 1))) & 0xffffu;
            act[a] = m[a] == 0xffffu;
            all_dense = all_dense && act[a];
            dense_or_empty = dense_or_empty && (act[a] || m[a] == 0u);
            any = any || m[a] != 0u;
            all_equal = all_equal && m[a] == m[0];
            ap[a] = As + s * SW_TILE + (rb * 16) * 32 + lane * SW_LANE_STRIDE;
        }
#endif
        const double* bp = xb + cb0 * SW_B_CB + lane * SW_LANE_STRIDE;
        if (any && !(flags & SW_NO_OPERAND)) {
#if JK_SW_PAIR
            if (all_dense) sweep_mma_dense_pair(acc, ap, bp);
#else
            if (all_dense) sweep_mma_dense<true>(acc, ap, act, bp);
#endif
#if JK_SW_PRED
            else if (SW_RBN == 2 && m[0] == 0u) sweep_mma_one<SW_RBN - 1>(acc, ap[SW_RBN - 1], m[SW_RBN - 1], bp);
            else if (SW_RBN == 2 && m[1] == 0u) sweep_mma_one<0>(acc, ap[0], m[0], bp);
            else if (SW_RBN == 2 && JK_SW_PRED == 1) sweep_mma_pred(acc, ap, m, bp);
#endif
            else if (all_equal) sweep_mma_uniform(acc, ap, m[0], bp);
#if JK_SW_MASKED
            else if (m[0] != 0u && m[SW_RBN - 1] != 0u) sweep_mma_masked(acc, ap, m, bp);     // both row blocks active, different masks: one interleaved loop, B loaded once
#endif
            else if (SW_RBN <= 2) { sweep_mma_single<0>(acc, ap[0], m[0], bp); sweep_mma_single<SW_RBN - 1>(acc, ap[SW_RBN - 1], m[SW_RBN - 1], bp); }
            else if (dense_or_empty) sweep_mma_dense<false>(acc, ap, act, bp);
            else sweep_mma_masked(acc, ap, m, bp);
        }
        __syncwarp();
        if (tracing) trace[((size_t)n * 2 + trace_w) * 4 + 2] = clock64();
        if (profiling) {
            t1 = clock64(); pc[2] += t1 - t0; pc[5] += 1;
            for (int a = 0; a < SW_RBN; ++a) pc[4] += __popc(m[a]) * SW_CBN;
        }
        if (lane == 0) {
            mbar_arrive(bar_empty + 8 * s);
            if (flags & SW_DIAG) mbar_arrive(bar_bempty);
        }
        if (flags & SW_DIAG) ++ndiag;
        if (flags & SW_ROW_END) {
            double* g = ((flags & SW_OUT_FRAG) ? Zslab : Xslab) + (size_t)row * SW_XTILE;
            const int oslot = (xinfo >> 16) & 0xff;
            double* xr = Xr + oslot * SW_XTILE;
#pragma unroll
            for (int a = 0; a < SW_RBN; ++a)
#pragma unroll
                for (int b = 0; b < SW_CBN; ++b) {
                    const int mb = rbs[a], nt = cb0 + b;
                    if (flags & SW_OUT_FRAG) { g[fx_off(mb, nt, 0)] = acc[a][b][0]; g[fx_off(mb, nt, 1)] = acc[a][b][1]; }
                    else *reinterpret_cast<double2*>(g + rm_off(mb, nt)) = make_double2(acc[a][b][0], acc[a][b][1]);
                    if (!(flags & SW_NO_RING)) { xr[fx_off(mb, nt, 0)] = acc[a][b][0]; xr[fx_off(mb, nt, 1)] = acc[a][b][1]; }
                }
            if (!(flags & SW_NO_RING)) { __syncwarp(); if (lane == 0) mbar_arrive(bar_x + 8 * oslot); }
            if (profiling) { t0 = clock64(); pc[3] += t0 - t1; }
        }
        if (tracing) trace[((size_t)n * 2 + trace_w) * 4 + 3] = clock64();
    }
    if (profiling && lane == 0 && warp < min(8, SW_CONSUMER_WARPS)) {
        pc[6] = clock64() - t_begin;
        for (int i = 0; i < 8; ++i) prof[warp * 8 + i] = pc[i];
    }
}

}  // namespace jk
