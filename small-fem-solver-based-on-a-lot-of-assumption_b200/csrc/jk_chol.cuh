// jk_chol.cuh -- blocked Cholesky of the tile-banded K_ff and the multi-right-hand-side triangular
// sweeps.  Replaces np.linalg.solve(K_ff, F_f) of FEMSolver.solve (GUI.py:481-490): the matrix is
// factored once (K_ff = L L^T) and every phase / load case is a pair of triangular sweeps.
//
// Storage: NB x NB row-major tiles, tile (I, J) (J <= I, I - J <= bw) at tile_off(I, J, bw).  With
// bw = n_tiles - 1 this is the reference's dense K_ff; with a reverse-Cuthill-McKee node order bw is a
// handful of tiles.  All dense tile products run on the FP64 tensor pipe (mma.sync m8n8k4, DMMA).
#pragma once
#include "jk_common.cuh"

namespace jk {

constexpr int LS_LD = NB + 4;     // smem row stride of an L tile  (== 4 mod 16 -> conflict-free DMMA fragment loads)
constexpr int XS_LD = SLAB + 4;   // smem row stride of an X tile
constexpr int SOLVE_STAGES = 3;
constexpr int SOLVE_NSPLIT = 1;       // column groups per tile row: warps = 8 * NSPLIT, each owns 8 rows x (SLAB / NSPLIT) columns
constexpr int SOLVE_NT = SLAB / 8 / SOLVE_NSPLIT;   // 8-column DMMA tiles per warp
constexpr int SOLVE_THREADS = 256 * SOLVE_NSPLIT;
constexpr size_t SOLVE_SMEM = (size_t)SOLVE_STAGES * (NB * LS_LD + NB * XS_LD) * sizeof(double) + (size_t)NB * XS_LD * sizeof(double);
constexpr size_t UPDATE_SMEM = (size_t)2 * NB * LS_LD * sizeof(double);

// ----------------------------------------------------------------------------------------------
// K3a: Cholesky of diagonal tile k (one CTA, right-looking in shared memory)
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_potrf_tile(double* __restrict__ tiles, int k, int bw, int* __restrict__ info) {
    __shared__ double A[NB][NB + 1];
    double* g = tiles + tile_off(k, k, bw);
    const int tid = threadIdx.x;
    for (int idx = tid; idx < NB * NB; idx += 256) {
        int r = idx / NB, c = idx % NB;
        A[r][c] = (c <= r) ? g[idx] : 0.0;
    }
    __syncthreads();
    const int ty = tid / 16, tx = tid % 16;
    for (int j = 0; j < NB; ++j) {
        double d = A[j][j];
        if (!(d > 0.0)) {
            if (tid == 0) atomicCAS(info, 0, k * NB + j + 1);
            d = 1.0;
        }
        double rinv = rsqrt(d);
        __syncthreads();
        if (tid == j) A[j][j] = d * rinv;
        else if (tid > j && tid < NB) A[tid][j] *= rinv;
        __syncthreads();
        for (int i = j + 1 + ty; i < NB; i += 16) {
            double lij = A[i][j];
            for (int c = j + 1 + tx; c <= i; c += 16) A[i][c] = fma(-lij, A[c][j], A[i][c]);
        }
        __syncthreads();
    }
    for (int idx = tid; idx < NB * NB; idx += 256) {
        int r = idx / NB, c = idx % NB;
        g[idx] = A[r][c];
    }
}

// ----------------------------------------------------------------------------------------------
// K3b: panel.  Tile (k+1+b, k) <- A * L_kk^{-T}; one CTA per tile, one thread per tile row, the row
// lives in registers and is solved by forward substitution (fully unrolled).
// ----------------------------------------------------------------------------------------------
constexpr size_t PANEL_SMEM = (size_t)(2 * NB * (NB + 1) + NB) * sizeof(double);
__global__ void __launch_bounds__(NB) k_panel_trsm(double* __restrict__ tiles, int k, int bw) {
    extern __shared__ __align__(16) double smem[];
    double (*Ls)[NB + 1] = reinterpret_cast<double (*)[NB + 1]>(smem);
    double (*As)[NB + 1] = reinterpret_cast<double (*)[NB + 1]>(smem + NB * (NB + 1));
    double* inv_d = smem + 2 * NB * (NB + 1);
    const int tid = threadIdx.x;
    const double* gl = tiles + tile_off(k, k, bw);
    double* ga = tiles + tile_off(k + 1 + blockIdx.x, k, bw);
    for (int idx = tid; idx < NB * NB; idx += NB) {
        int r = idx / NB, c = idx % NB;
        Ls[r][c] = gl[idx];
        As[r][c] = ga[idx];
    }
    __syncthreads();
    inv_d[tid] = 1.0 / Ls[tid][tid];
    double a[NB];
#pragma unroll
    for (int c = 0; c < NB; ++c) a[c] = As[tid][c];
    __syncthreads();
#pragma unroll
    for (int c = 0; c < NB; ++c) {
        double x = a[c] * inv_d[c];
        a[c] = x;
#pragma unroll
        for (int c2 = c + 1; c2 < NB; ++c2) a[c2] = fma(-x, Ls[c2][c], a[c2]);
    }
#pragma unroll
    for (int c = 0; c < NB; ++c) As[tid][c] = a[c];
    __syncthreads();
    for (int idx = tid; idx < NB * NB; idx += NB) ga[idx] = As[idx / NB][idx % NB];
}

// ----------------------------------------------------------------------------------------------
// K3c: trailing update.  Tile (i, j) -= L_ik * L_jk^T for k < j <= i <= k + w.  One CTA (4 warps)
// per tile, operands staged in shared memory by cp.async, product on DMMA.
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void load_tile_async(double* __restrict__ dst /* [NB][LS_LD] */, const double* __restrict__ src, int tid, int nthreads) {
    for (int q = tid; q < NB * (NB / 2); q += nthreads) {
        int r = q / (NB / 2), cc = q % (NB / 2);
        cp_async16(dst + r * LS_LD + 2 * cc, src + r * NB + 2 * cc);
    }
}

__global__ void __launch_bounds__(128) k_trailing_update(double* __restrict__ tiles, int k, int w, int bw) {
    extern __shared__ __align__(16) double smem[];
    double* As = smem;
    double* Bs = smem + NB * LS_LD;
    // decode (bi, bj), bi >= bj, from the linear block index
    int lin = blockIdx.x, bi = 0;
    while ((bi + 1) * (bi + 2) / 2 <= lin) ++bi;
    int bj = lin - bi * (bi + 1) / 2;
    const int i = k + 1 + bi, j = k + 1 + bj;
    (void)w;
    const int tid = threadIdx.x, lane = tid % 32, warp = tid / 32;
    load_tile_async(As, tiles + tile_off(i, k, bw), tid, 128);
    load_tile_async(Bs, tiles + tile_off(j, k, bw), tid, 128);
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    double acc[2][8][2];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;
    const int fr = lane / 4, fk = lane % 4;
    const double* ap = As + (16 * warp + fr) * LS_LD + fk;
    const double* bp = Bs + fr * LS_LD + fk;
#pragma unroll 4
    for (int k4 = 0; k4 < NB / 4; ++k4) {
        double a0 = ap[4 * k4], a1 = ap[8 * LS_LD + 4 * k4];
        double b[8];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) b[nt] = bp[nt * 8 * LS_LD + 4 * k4];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            dmma(acc[0][nt][0], acc[0][nt][1], a0, b[nt]);
            dmma(acc[1][nt][0], acc[1][nt][1], a1, b[nt]);
        }
    }
    double* gc = tiles + tile_off(i, j, bw);
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            int r = 16 * warp + 8 * mt + fr, c = 8 * nt + 2 * fk;
            double2* p = reinterpret_cast<double2*>(gc + r * NB + c);
            double2 v = *p;
            v.x -= acc[mt][nt][0]; v.y -= acc[mt][nt][1];
            *p = v;
        }
}

// ----------------------------------------------------------------------------------------------
// Inverses of all diagonal tiles (batched, off the critical path): Linv[k] = L_kk^{-1}, lower.
// The sweeps then apply diagonal blocks as DMMA products instead of scalar substitutions.
// ----------------------------------------------------------------------------------------------
constexpr size_t INVERSE_SMEM = (size_t)(2 * NB * (NB + 1)) * sizeof(double);
// Two chains in one launch: blocks [0, n0) invert the diagonal tiles of (tiles, Linv), blocks [n0, gridDim.x) those of
// (tiles1, Linv1) -- the kernel is a latency chain of 64 steps per tile, so one wave for both chains costs what one costs.
__global__ void __launch_bounds__(256) k_tile_inverse(const double* __restrict__ tiles, double* __restrict__ Linv, int bw, int n0,
                                                      const double* __restrict__ tiles1, double* __restrict__ Linv1, int bw1) {
    extern __shared__ __align__(16) double smem[];
    double (*Ls)[NB + 1] = reinterpret_cast<double (*)[NB + 1]>(smem);
    double (*X)[NB + 1] = reinterpret_cast<double (*)[NB + 1]>(smem + NB * (NB + 1));
    int k = blockIdx.x;
    const int tid = threadIdx.x;
    if (k >= n0) { k -= n0; tiles = tiles1; Linv = Linv1; bw = bw1; }
    const double* g = tiles + tile_off(k, k, bw);
    for (int idx = tid; idx < NB * NB; idx += 256) {
        int r = idx / NB, c = idx % NB;
        Ls[r][c] = (c <= r) ? g[idx] : 0.0;
        X[r][c] = (r == c) ? 1.0 : 0.0;
    }
    __syncthreads();
    const int ty = tid / 16, tx = tid % 16;
    for (int kk = 0; kk < NB; ++kk) {
        double dinv = 1.0 / Ls[kk][kk];
        if (tid <= kk) X[kk][tid] *= dinv;
        __syncthreads();
        for (int i = kk + 1 + ty; i < NB; i += 16) {
            double lik = Ls[i][kk];
            for (int c = tx; c <= kk; c += 16) X[i][c] = fma(-lik, X[kk][c], X[i][c]);
        }
        __syncthreads();
    }
    double* o = Linv + (size_t)k * NB * NB;
    for (int idx = tid; idx < NB * NB; idx += 256) o[idx] = X[idx / NB][idx % NB];
}

// Blocked variant for tiles factored by the cluster kernel, which leaves the inverses of the eight 8x8 diagonal blocks
// of every L_kk in dinv[k][8][8][8]: X = L^-1 by block forward substitution,
//     X_bc = Dinv_b ( delta_bc I - sum_{e=c}^{b-1} L_be X_ec ),   b = 0..7, c <= b,
// eight steps of two barriers each instead of 64 (5 us instead of 57 us on the tail of the factor stage).
constexpr size_t INVERSE_BLOCKED_SMEM = (size_t)(2 * NB * (NB + 1) + 8 * NB) * sizeof(double);
__global__ void __launch_bounds__(256) k_tile_inverse_blocked(const double* __restrict__ tiles, const double* __restrict__ dinv, double* __restrict__ Linv,
                                                              int bw, int k_first, int n0, const double* __restrict__ tiles1,
                                                              const double* __restrict__ dinv1, double* __restrict__ Linv1, int bw1, int k_first1) {
    extern __shared__ __align__(16) double smem[];
    double (*Ls)[NB + 1] = reinterpret_cast<double (*)[NB + 1]>(smem);
    double (*X)[NB + 1] = reinterpret_cast<double (*)[NB + 1]>(smem + NB * (NB + 1));
    double (*T)[NB] = reinterpret_cast<double (*)[NB]>(smem + 2 * NB * (NB + 1));      // [8][NB] rows of the current block step
    int k = blockIdx.x;
    const int tid = threadIdx.x;
    if (k >= n0) { k = k - n0 + k_first1; tiles = tiles1; dinv = dinv1; Linv = Linv1; bw = bw1; } else k += k_first;
    const double* g = tiles + tile_off(k, k, bw);
    const double* di = dinv + (size_t)k * 512;
    for (int idx = tid; idx < NB * NB; idx += 256) {
        int r = idx / NB, c = idx % NB;
        Ls[r][c] = (c <= r) ? g[idx] : 0.0;
        X[r][c] = 0.0;
    }
    __syncthreads();
    const int r8 = tid / 32, cq = tid % 32;                // row inside the block step, column (and column + 32)
    for (int b = 0; b < 8; ++b) {
        const int ncol = 8 * (b + 1);                      // columns 0 .. 8b+7 can be non-zero in block row b
        const int r = 8 * b + r8;
        for (int c = cq; c < ncol; c += 32) {
            double acc = (c == r) ? 1.0 : 0.0;
            for (int m = (c / 8) * 8; m < 8 * b; ++m) acc = fma(-Ls[r][m], X[m][c], acc);     // X[m][c] = 0 for m < 8*(c/8)
            T[r8][c] = acc;
        }
        __syncthreads();
        for (int c = cq; c < ncol; c += 32) {
            double acc = 0.0;
#pragma unroll
            for (int q = 0; q < 8; ++q) acc = fma(di[b * 64 + r8 * 8 + q], T[q][c], acc);
            X[r][c] = acc;
        }
        __syncthreads();
    }
    double* o = Linv + (size_t)k * NB * NB;
    for (int idx = tid; idx < NB * NB; idx += 256) o[idx] = X[idx / NB][idx % NB];
}

// ----------------------------------------------------------------------------------------------
// K4: triangular sweeps over one slab of SLAB right-hand sides per CTA (persistent over all tile rows).
//
//   forward :  X_k = Linv_kk ( B_k - sum_{j=k-bw}^{k-1} L_kj   X_j )      k = 0 .. NT-1
//   backward:  X_k = Linv_kk^T ( Y_k - sum_{i=k+1}^{k+bw} L_ik^T X_i )    k = NT-1 .. 0
//
// Every term is a (NB x NB) x (NB x SLAB) product on DMMA.  L tiles (and X tiles older than the
// previous step) stream from L2/HBM through a 3-stage cp.async ring; the newest X tile and the
// intermediate t stay in shared memory.  The slab's rows are updated in place.
// ----------------------------------------------------------------------------------------------
// Walks the tile products of a sweep.  kx = first "special" tile row:
//   forward : rows k >= kx are PARTIAL -- they only accumulate B_k - sum_{j < kx} L_kj X_j (no diagonal solve);
//             used for the separator rows of the second chain, whose sums are merged into the first chain.
//   backward: rows k >= kx are KNOWN   -- already solved (separator solution copied in); the sweep starts at kx-1.
// kx = NT gives the plain sweep.
template <bool BWD>
struct SweepIter {
    int k, j, NT, bw, kx;
    __device__ int first_j(int kk) const {
        if (!BWD) { int lo = max(0, kk - bw), hi = min(kk - 1, kx - 1); return lo <= hi ? lo : kk; }
        int hi = min(NT - 1, kk + bw); return hi >= kk + 1 ? hi : kk;
    }
    __device__ void init(int nt, int b, int kspecial) {
        NT = nt; bw = b; kx = kspecial;
        k = BWD ? kx - 1 : 0;
        j = BWD ? (k >= 0 ? first_j(k) : 0) : first_j(0);
    }
    __device__ bool done() const { return BWD ? (k < 0) : (k >= NT); }
    __device__ bool is_diag() const { return j == k; }
    __device__ bool partial_row() const { return !BWD && k >= kx; }
    __device__ bool x_in_smem() const { return BWD ? (j == k + 1 && k + 1 < kx) : (j == k - 1); }
    __device__ void next() {
        if (!BWD) {
            if (j == k) { ++k; j = first_j(k); }
            else { ++j; if (j > min(k - 1, kx - 1)) j = k; }
        } else {
            if (j == k) { --k; if (k >= 0) j = first_j(k); }
            else { --j; }
        }
    }
};

template <bool BWD>
__global__ void __launch_bounds__(SOLVE_THREADS, 1)
k_slab_sweep(const double* __restrict__ tiles, const double* __restrict__ Linv, double* __restrict__ X,
             int NT, int bw, int n_pad /* rows of the whole slab */, int row0 /* first row of this chain in the slab */,
             int kx /* first partial (forward) / known (backward) tile row, NT for a plain sweep */) {
    extern __shared__ __align__(16) double smem[];
    double* Ls = smem;                                         // [STAGES][NB][LS_LD]
    double* Xs = Ls + SOLVE_STAGES * NB * LS_LD;               // [STAGES][NB][XS_LD]
    double* Ts = Xs + SOLVE_STAGES * NB * XS_LD;               // [NB][XS_LD]  newest X tile / intermediate t
    const int tid = threadIdx.x, lane = tid % 32, warp = tid / 32;
    const int fr = lane / 4, fk = lane % 4;
    double* Xslab = X + ((size_t)blockIdx.x * (size_t)n_pad + (size_t)row0) * SLAB;

    auto issue = [&](const SweepIter<BWD>& it, int stage) {
        if (it.done() || (it.is_diag() && it.partial_row())) return;
        const double* lsrc = it.is_diag() ? (Linv + (size_t)it.k * NB * NB)
                                          : (tiles + (BWD ? tile_off(it.j, it.k, bw) : tile_off(it.k, it.j, bw)));
        load_tile_async(Ls + stage * NB * LS_LD, lsrc, tid, SOLVE_THREADS);
        if (!it.is_diag() && !it.x_in_smem()) {
            const double* xsrc = Xslab + (size_t)it.j * NB * SLAB;
            double* xdst = Xs + stage * NB * XS_LD;
            for (int q = tid; q < NB * (SLAB / 2); q += SOLVE_THREADS) {
                int r = q / (SLAB / 2), cc = q % (SLAB / 2);
                cp_async16(xdst + r * XS_LD + 2 * cc, xsrc + r * SLAB + 2 * cc);
            }
        }
    };

    SweepIter<BWD> it_load, it;
    it_load.init(NT, bw, kx); it.init(NT, bw, kx);
    // prologue: STAGES-1 items in flight
    for (int s = 0; s < SOLVE_STAGES - 1; ++s) { issue(it_load, s); cp_async_commit(); it_load.next(); }

    double acc[SOLVE_NT][2];
    double bk[SOLVE_NT][2];
    const int row = 8 * (warp % 8) + fr;                       // this lane's row inside the tile
    const int col0 = (warp / 8) * (SLAB / SOLVE_NSPLIT);        // first column of this warp's column group
    int n = 0;                                                 // item counter
    bool new_step = true;
    while (!it.done()) {
        const int stage = n % SOLVE_STAGES;
        if (new_step) {
            // right-hand side rows of this step (fragment layout), issued early
#pragma unroll
            for (int nt = 0; nt < SOLVE_NT; ++nt) {
                double2 v = *reinterpret_cast<const double2*>(Xslab + ((size_t)it.k * NB + row) * SLAB + col0 + 8 * nt + 2 * fk);
                bk[nt][0] = v.x; bk[nt][1] = v.y;
                acc[nt][0] = acc[nt][1] = 0.0;
            }
            new_step = false;
        }
        cp_async_wait<SOLVE_STAGES - 2>();
        __syncthreads();                                       // item n landed; everyone is done with item n-1
        // The copies of item n+STAGES-1 (into the stage item n-1 just released) are issued INSIDE the DMMA loop below,
        // one 16-byte chunk per thread per k-step, so the tensor pipe never waits for a copy-issue phase.
        const int lstage = (n + SOLVE_STAGES - 1) % SOLVE_STAGES;
        const bool ld_l = !it_load.done() && !(it_load.is_diag() && it_load.partial_row());
        const bool ld_x = ld_l && !it_load.is_diag() && !it_load.x_in_smem();
        const double* lsrc = nullptr; const double* xsrc = nullptr;
        if (ld_l) lsrc = (it_load.is_diag() ? (Linv + (size_t)it_load.k * NB * NB)
                                            : (tiles + (BWD ? tile_off(it_load.j, it_load.k, bw) : tile_off(it_load.k, it_load.j, bw))))
                         + (tid >> 5) * NB + 2 * (tid & 31);
        if (ld_x) xsrc = Xslab + (size_t)it_load.j * NB * SLAB + (tid >> 4) * SLAB + 2 * (tid & 15);
        double* ldst = Ls + lstage * NB * LS_LD + (tid >> 5) * LS_LD + 2 * (tid & 31);
        double* xdst = Xs + lstage * NB * XS_LD + (tid >> 4) * XS_LD + 2 * (tid & 15);
        constexpr int LROWS = SOLVE_THREADS / 32, XROWS = SOLVE_THREADS / 16;   // tile rows covered by one chunk round
        constexpr int LCH = NB / LROWS, XCH = NB / XROWS;                        // chunk rounds per thread
        it_load.next();

        const double* ls = Ls + stage * NB * LS_LD;
        const double* xs;
        if (it.is_diag() && it.partial_row()) {
            // partial row: store B_k - sum (no diagonal solve); nothing later in this sweep reads it
#pragma unroll
            for (int nt = 0; nt < SOLVE_NT; ++nt) {
                double2 v = make_double2(bk[nt][0] - acc[nt][0], bk[nt][1] - acc[nt][1]);
                *reinterpret_cast<double2*>(Xslab + ((size_t)it.k * NB + row) * SLAB + col0 + 8 * nt + 2 * fk) = v;
            }
            // no DMMA loop in this item: issue the copies of item n+2 here
#pragma unroll
            for (int q = 0; q < LCH; ++q) if (ld_l) cp_async16(ldst + q * LROWS * LS_LD, lsrc + q * LROWS * NB);
#pragma unroll
            for (int q = 0; q < XCH; ++q) if (ld_x) cp_async16(xdst + q * XROWS * XS_LD, xsrc + q * XROWS * SLAB);
            cp_async_commit();
            new_step = true;
            it.next();
            ++n;
            continue;
        }
        if (it.is_diag()) {
            // t = B_k - acc  -> Ts, then acc = Linv * t
#pragma unroll
            for (int nt = 0; nt < SOLVE_NT; ++nt) {
                Ts[row * XS_LD + col0 + 8 * nt + 2 * fk] = bk[nt][0] - acc[nt][0];
                Ts[row * XS_LD + col0 + 8 * nt + 2 * fk + 1] = bk[nt][1] - acc[nt][1];
                acc[nt][0] = acc[nt][1] = 0.0;
            }
            __syncthreads();
            xs = Ts;
        } else {
            xs = it.x_in_smem() ? Ts : (Xs + stage * NB * XS_LD);
        }
        // acc += op(L) * xs, op = identity (forward) or transpose (backward)
        const double* a_base = BWD ? (ls + fk * LS_LD + row) : (ls + row * LS_LD + fk);
        const double* b_base = xs + fk * XS_LD + col0 + fr;
        // fragments of step k4+1 are loaded before the DMMAs of step k4 are issued (register double buffer,
        // static indices after full unrolling) so the tensor pipe never waits on a shared-memory load
        double af[2], bf[2][SOLVE_NT];
        af[0] = a_base[0];
#pragma unroll
        for (int nt = 0; nt < SOLVE_NT; ++nt) bf[0][nt] = b_base[8 * nt];
#pragma unroll
        for (int k4 = 0; k4 < NB / 4; ++k4) {
            // copy chunks of item n+2: (LCH + XCH) x 16 B per thread = one 64x64 L tile + one 64x32 X tile per CTA
            if (k4 < LCH) { if (ld_l) cp_async16(ldst + k4 * LROWS * LS_LD, lsrc + k4 * LROWS * NB); }
            else if (k4 < LCH + XCH) { if (ld_x) cp_async16(xdst + (k4 - LCH) * XROWS * XS_LD, xsrc + (k4 - LCH) * XROWS * SLAB); }
            const int cur = k4 & 1, nxt = cur ^ 1;
            if (k4 + 1 < NB / 4) {
                af[nxt] = BWD ? a_base[4 * (k4 + 1) * LS_LD] : a_base[4 * (k4 + 1)];
#pragma unroll
                for (int nt = 0; nt < SOLVE_NT; ++nt) bf[nxt][nt] = b_base[4 * (k4 + 1) * XS_LD + 8 * nt];
            }
#pragma unroll
            for (int nt = 0; nt < SOLVE_NT; ++nt) dmma(acc[nt][0], acc[nt][1], af[cur], bf[cur][nt]);
        }
        cp_async_commit();
        if (it.is_diag()) {
            // acc is X_k: store to the slab (in place) and keep it in Ts as the newest tile
            __syncthreads();                                   // all warps finished reading Ts (= t)
#pragma unroll
            for (int nt = 0; nt < SOLVE_NT; ++nt) {
                double2 v = make_double2(acc[nt][0], acc[nt][1]);
                *reinterpret_cast<double2*>(Xslab + ((size_t)it.k * NB + row) * SLAB + col0 + 8 * nt + 2 * fk) = v;
                Ts[row * XS_LD + col0 + 8 * nt + 2 * fk] = v.x; Ts[row * XS_LD + col0 + 8 * nt + 2 * fk + 1] = v.y;
            }
            new_step = true;
        }
        it.next();
        ++n;
    }
    cp_async_wait<0>();
}

}  // namespace jk
