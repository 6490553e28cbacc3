"""Host logic of the TMA sweep (csrc/jk_sweep.cuh): the item list built by the library is interpreted here in NumPy --
same algebra (Z-form recurrences with U = -L Linv, G = Linv^T Linv, V = -Linv^T L^T), same ring-slot / mbarrier-parity
bookkeeping -- and checked against dense triangular solves.  No GPU needed: jk_sweep_program is host-only."""
import ctypes as C

import numpy as np
import pytest

import jacket_b200  # noqa: F401
from jacket_b200 import _lib

NB = 64
RING = 5
ROW_BEGIN, DIAG, ROW_END, NO_RING, INIT_RHS, OUT_FRAG, NO_OPERAND, WAIT_X = 1, 2, 4, 8, 16, 32, 64, 128


def program(NT, bw, kx, backward, first_tile=None):
    lib = _lib.lib()
    meta = np.zeros(3, dtype=np.int32)
    ft = None if first_tile is None else np.ascontiguousarray(first_tile, dtype=np.int32)
    n = lib.jk_sweep_program(NT, bw, kx, int(backward), _lib.iptr(ft), None, 0, _lib.iptr(meta))
    assert n > 0
    items = np.zeros(6 * n, dtype=np.int32)
    assert lib.jk_sweep_program(NT, bw, kx, int(backward), _lib.iptr(ft), _lib.iptr(items), n, _lib.iptr(meta)) == n
    return items.reshape(n, 6), meta


def banded_spd(NT, bw, rng):
    n = NT * NB
    A = np.zeros((n, n))
    hb = bw * NB - 7 if bw > 0 else NB - 1
    for i in range(n):
        lo = max(0, i - hb)
        A[i, lo:i] = rng.standard_normal(i - lo) * 0.1
    A = A + A.T + np.eye(n) * (2.0 + 0.2 * hb)
    return A


def tile(Lm, i, j):
    return Lm[i * NB:(i + 1) * NB, j * NB:(j + 1) * NB]


def run_program(items, meta, Lm, X, backward, known_rows=None, forward_continuation=None):
    """Interpret the item list on X ([NT*NB, nrhs], modified in place like the slab).  Returns tile products executed.
    forward_continuation = (first_row, npre): the items continue a forward program from first_row in a NEW launch: the
    npre rows before it are reloaded from the slab (they hold Z) and the slots' mbarrier phases are pre-advanced."""
    NT = Lm.shape[0] // NB
    Linv = [np.linalg.inv(tile(Lm, k, k)) for k in range(NT)]
    ring_row = [None] * RING
    ring_val = [None] * RING
    fills = [0] * RING
    pre_row, npre, ktop = (int(v) for v in meta)
    waited = set()      # operand rows some earlier item of the program already waited on
    if forward_continuation is not None:
        k1, npre = forward_continuation
        pre_row = k1 - npre
        for r in range(pre_row, pre_row + RING):          # xphase bits of jk_set_supports
            fills[r % RING] = (r // RING) & 1
        for i in range(pre_row, k1):
            slot = i % RING
            ring_row[slot], ring_val[slot] = i, X[i * NB:(i + 1) * NB].copy()
            fills[slot] += 1
            waited.add(i)                                  # made visible by the consumers' barrier after the preload
    else:
        for q in range(npre):
            i = pre_row + q
            slot = (ktop - i) % RING
            ring_row[slot], ring_val[slot] = i, X[i * NB:(i + 1) * NB].copy()
            fills[slot] += 1
    acc = None
    rhs = X[items[0][0] * NB:(items[0][0] + 1) * NB].copy() if items[0][2] & INIT_RHS else None
    nprod = 0
    for row, src, flags, xinfo, next_row, next_init in items:
        if flags & ROW_BEGIN:
            acc = rhs.copy() if flags & INIT_RHS else np.zeros((NB, X.shape[1]))
            if next_init:      # the next row's right-hand side is fetched a whole row ahead, before this row is stored
                rhs = X[next_row * NB:(next_row + 1) * NB].copy()
        else:
            assert not next_init
        if flags & DIAG:
            assert backward
            acc += (Linv[row].T @ Linv[row]) @ X[row * NB:(row + 1) * NB]      # operand = Z_k from the slab
            nprod += 1
        elif not flags & NO_OPERAND:
            slot, parity = xinfo & 0xFF, (xinfo >> 8) & 1
            assert ring_row[slot] == src, (row, src, slot, ring_row)
            if flags & WAIT_X:
                assert (fills[slot] - 1) & 1 == parity, "mbarrier parity of the operand slot"
                waited.add(src)
            else:
                assert src in waited, "operand row used without a wait must have been waited on earlier"
            if backward:
                A = -Linv[row].T @ tile(Lm, src, row).T
            else:
                A = -tile(Lm, row, src) @ Linv[src]
            acc += A @ ring_val[slot]
            nprod += 1
        if flags & ROW_END:
            X[row * NB:(row + 1) * NB] = acc
            if not flags & NO_RING:
                oslot = (xinfo >> 16) & 0xFF
                ring_row[oslot], ring_val[oslot] = row, acc.copy()
                fills[oslot] += 1
    return nprod


@pytest.mark.parametrize("NT,bw", [(1, 0), (2, 1), (7, 2), (13, 4), (12, 3)])
def test_plain_sweeps_solve_the_system(NT, bw):
    rng = np.random.default_rng(NT * 10 + bw)
    A = banded_spd(NT, bw, rng)
    Lm = np.linalg.cholesky(A)
    B = rng.standard_normal((NT * NB, 5))
    X = B.copy()
    f, mf = program(NT, bw, NT, False)
    b, mb = program(NT, bw, NT, True)
    run_program(f, mf, Lm, X, False)
    # after the forward sweep the slab holds Z_k = L_kk Y_k
    Y = np.linalg.solve(Lm, B)
    for k in range(NT):
        np.testing.assert_allclose(X[k * NB:(k + 1) * NB], tile(Lm, k, k) @ Y[k * NB:(k + 1) * NB], rtol=1e-9, atol=1e-11)
    run_program(b, mb, Lm, X, True)
    np.testing.assert_allclose(X, np.linalg.solve(A, B), rtol=1e-9, atol=1e-11)
    # every row has exactly one ROW_BEGIN and one ROW_END, in processing order
    for items, order in ((f, list(range(NT))), (b, list(range(NT - 1, -1, -1)))):
        assert [r for r, _, fl, *_ in items if fl & ROW_BEGIN] == order
        assert [r for r, _, fl, *_ in items if fl & ROW_END] == order


@pytest.mark.parametrize("NT,bw,kx", [(9, 3, 6), (12, 4, 9), (10, 4, 4), (5, 2, 3)])
def test_partial_forward_and_known_backward(NT, bw, kx):
    """Second-chain semantics: forward rows >= kx only accumulate B_k - sum_{j<kx} L_kj Y_j; the backward sweep starts
    below kx with rows >= kx known."""
    rng = np.random.default_rng(100 * NT + kx)
    A = banded_spd(NT, bw, rng)
    Lm = np.linalg.cholesky(A)
    B = rng.standard_normal((NT * NB, 3))
    n1 = kx * NB
    X = B.copy()
    f, mf = program(NT, bw, kx, False)
    run_program(f, mf, Lm, X, False)
    Y1 = np.linalg.solve(Lm[:n1, :n1], B[:n1])
    np.testing.assert_allclose(X[n1:], B[n1:] - Lm[n1:, :n1] @ Y1, rtol=1e-9, atol=1e-11)
    assert all(fl & NO_RING for r, _, fl, *_ in f if fl & ROW_END and r >= kx)
    # backward: take the exact solution of L^T x = y, give the rows >= kx, recover the rest from Z
    Yfull = rng.standard_normal((NT * NB, 3))
    Xtrue = np.linalg.solve(Lm.T, Yfull)
    S = np.zeros_like(Yfull)
    for k in range(kx):
        S[k * NB:(k + 1) * NB] = tile(Lm, k, k) @ Yfull[k * NB:(k + 1) * NB]     # Z rows
    S[n1:] = Xtrue[n1:]
    b, mb = program(NT, bw, kx, True)
    assert mb[0] == kx and mb[1] == min(bw, NT - kx)
    run_program(b, mb, Lm, S, True)
    np.testing.assert_allclose(S[:n1], Xtrue[:n1], rtol=1e-8, atol=1e-10)


def test_program_rejects_wide_bands():
    lib = _lib.lib()
    assert lib.jk_sweep_program(8, 5, 8, 0, None, None, 0, None) < 0
    assert lib.jk_sweep_program(0, 1, 0, 0, None, None, 0, None) < 0


def ragged_spd(NT, bw, rng):
    """SPD matrix whose row envelope varies from row to row (like a frame: some rows reach far back, most do not)."""
    n = NT * NB
    A = np.zeros((n, n))
    hbmax = bw * NB - 7
    for i in range(n):
        reach = hbmax if (i // 6) % 5 == 0 else int(rng.integers(3, max(4, hbmax // 3)))
        lo = max(0, i - reach)
        A[i, lo:i] = rng.standard_normal(i - lo) * 0.1
    A = A + A.T + np.eye(n) * (2.0 + 0.2 * hbmax)
    return A


def first_tiles(A):
    n = A.shape[0]
    first = np.array([np.flatnonzero(A[i, :i + 1])[0] for i in range(n)])
    return np.array([first[k * NB:(k + 1) * NB].min() // NB for k in range(n // NB)], dtype=np.int32)


@pytest.mark.parametrize("NT,bw,kx", [(11, 4, 11), (9, 3, 9), (12, 4, 8), (7, 2, 4)])
def test_envelope_drops_empty_tiles_and_still_solves(NT, bw, kx):
    """With the row envelope given, tiles left of it are not visited (fewer items) and the sweeps still solve the
    system: Cholesky never fills outside the envelope."""
    rng = np.random.default_rng(7 * NT + kx)
    A = ragged_spd(NT, bw, rng)
    # make some far tiles structurally empty: rows of every third tile row only reach one tile back
    for k in range(2, NT, 3):
        A[k * NB:(k + 1) * NB, :(k - 1) * NB] = 0.0
        A[:(k - 1) * NB, k * NB:(k + 1) * NB] = 0.0
    Lm = np.linalg.cholesky(A)
    ft = first_tiles(A)
    assert np.array_equal(first_tiles(np.where(np.abs(Lm) > 0, 1.0, 0.0)), ft)        # no fill left of the envelope
    B = rng.standard_normal((NT * NB, 4))
    f_all, _ = program(NT, bw, kx, False)
    f, mf = program(NT, bw, kx, False, ft)
    b_all, _ = program(NT, bw, kx, True)
    b, mb = program(NT, bw, kx, True, ft)
    assert len(f) < len(f_all) and len(b) < len(b_all)
    # every normal row still ends on the tile next to the diagonal (the item that orders ring and slab updates)
    for items, step in ((f, -1), (b, 1)):
        for row, src, fl, *_ in items:
            if fl & ROW_END and not fl & (NO_RING | NO_OPERAND) and not (fl & DIAG):
                assert src == row + step
    X = B.copy()
    run_program(f, mf, Lm, X, False)
    n1 = kx * NB
    Y1 = np.linalg.solve(Lm[:n1, :n1], B[:n1])
    for k in range(kx):
        np.testing.assert_allclose(X[k * NB:(k + 1) * NB], tile(Lm, k, k) @ Y1[k * NB:(k + 1) * NB], rtol=1e-9, atol=1e-11)
    if kx < NT:
        np.testing.assert_allclose(X[n1:], B[n1:] - Lm[n1:, :n1] @ Y1, rtol=1e-9, atol=1e-11)
        Xtrue = np.linalg.solve(Lm.T, rng.standard_normal((NT * NB, 4)))
        Yfull = Lm.T @ Xtrue
        S = np.zeros_like(Yfull)
        for k in range(kx):
            S[k * NB:(k + 1) * NB] = tile(Lm, k, k) @ Yfull[k * NB:(k + 1) * NB]
        S[n1:] = Xtrue[n1:]
        run_program(b, mb, Lm, S, True)
        np.testing.assert_allclose(S[:n1], Xtrue[:n1], rtol=1e-8, atol=1e-10)
    else:
        run_program(b, mb, Lm, X, True)
        np.testing.assert_allclose(X, np.linalg.solve(A, B), rtol=1e-9, atol=1e-11)


@pytest.mark.parametrize("NT,bw,kx,k1", [(20, 4, 20, 16), (23, 3, 18, 14), (17, 4, 12, 11)])
def test_forward_program_split_into_two_launches(NT, bw, kx, k1):
    """Split factor: the forward items of tile rows < k1 run in one launch, the rest in a second one that reloads the
    last band rows (Z) from the slab and pre-advances the ring slots' mbarrier phases.  Same result as one launch."""
    rng = np.random.default_rng(31 * NT + k1)
    A = banded_spd(NT, bw, rng)
    Lm = np.linalg.cholesky(A)
    B = rng.standard_normal((NT * NB, 3))
    f, mf = program(NT, bw, kx, False)
    X1 = B.copy()
    run_program(f, mf, Lm, X1, False)
    n1 = int(np.argmax(f[:, 0] >= k1))
    assert 0 < n1 < len(f) and f[n1, 2] & ROW_BEGIN
    X2 = B.copy()
    run_program(f[:n1], mf, Lm, X2, False)
    assert np.array_equal(X2[k1 * NB:], B[k1 * NB:])            # untouched rows still hold the right-hand side
    run_program(f[n1:], mf, Lm, X2, False, forward_continuation=(k1, min(bw, k1)))
    np.testing.assert_allclose(X2, X1, rtol=1e-12, atol=1e-13)
