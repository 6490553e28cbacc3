"""Host-side nonlinear wave fits (wavefit.py).  raschii is absent, so these are self-consistency checks: the
free-surface boundary conditions are satisfied to the order of the theory, the fits reduce to linear theory for
small waves, and the reference's model-selection thresholds (GUI.py:208-253) are reproduced."""
import numpy as np
import pytest

from jacket_b200 import wavefit as wf

T, D = 9.4, 50.0


def test_stokes_residual_order():
    """Residual of an order-N Stokes wave scales like eps^N relative (eps^(N+1) absolute)."""
    for N, expo in ((1, 1), (2, 2), (3, 3), (5, 5)):
        r1 = max(wf.bc_residuals(wf.stokes_fit(1.0, T, D, N)))
        r2 = max(wf.bc_residuals(wf.stokes_fit(2.0, T, D, N)))
        assert 0.6 * 2**expo < r2 / r1 < 1.6 * 2**expo, (N, r1, r2)
    assert max(wf.bc_residuals(wf.stokes_fit(1.0, T, D, 5))) < 1e-7


def test_fenton_satisfies_boundary_conditions():
    for H, N, tol in ((1.0, 5, 1e-7), (8.0, 10, 1e-5), (12.0, 12, 1e-4)):
        f = wf.fenton_fit(H, T, D, N)
        assert max(wf.bc_residuals(f, 256)) < tol
        assert abs(f.eta(0.0)[0] - f.eta(np.pi)[0] - H) < 1e-9 * H                  # wave height
        assert abs(np.mean(f.eta(np.linspace(0, 2 * np.pi, 400, endpoint=False)))) < 1e-9 * H   # mean level
        assert abs(f.k * f.c * T - 2 * np.pi) < 1e-10


def test_gui_default_wave_is_solvable():
    """H = 17.038 m, T = 9.4 s, d = 50 m (GUI.py:1810) is at ~90 % of the breaking limit."""
    f = wf.fenton_fit(17.038, T, D, 10)
    assert 150.0 < f.length < 155.0 and 10.0 < f.eta(0.0)[0] < 11.0
    assert max(wf.bc_residuals(wf.fenton_fit(17.038, T, D, 20), 256)) < 1e-4


def test_reduction_to_linear_theory_and_agreement():
    k0 = wf.airy_wavenumber(2 * np.pi / T, D)
    for fit in (wf.stokes_fit(0.01, T, D, 5), wf.fenton_fit(0.01, T, D, 6)):
        assert abs(fit.k - k0) < 1e-7 * k0
        assert abs(fit.E[0] - 0.005) < 1e-8 and np.all(np.abs(fit.E[1:]) < 1e-6)
        assert abs(fit.B[0] - 0.005 * (2 * np.pi / T) / np.tanh(k0 * D)) < 1e-8
    s5, fe = wf.stokes_fit(6.0, T, D, 5), wf.fenton_fit(6.0, T, D, 10)
    assert abs(s5.k - fe.k) < 2e-6 * fe.k and np.max(np.abs(s5.E[:3] - fe.E[:3])) < 2e-3


def test_model_selection_thresholds():
    L0 = 2 * np.pi / wf.airy_wavenumber(2 * np.pi / T, D)
    assert wf.select_model(0.005 * L0, T, D, "auto", 10) == ("Airy", 1)
    assert wf.select_model(0.02 * L0, T, D, "auto", 10) == ("Stokes", 3)
    assert wf.select_model(0.05 * L0, T, D, "auto", 10) == ("Stokes", 5)
    assert wf.select_model(0.07 * L0, T, D, "auto", 10) == ("Fenton", 14)
    assert wf.select_model(0.12 * L0, T, D, "auto", 10) == ("Fenton", 20)
    assert wf.select_model(5.0, T, D, "Stokes", 8) == ("Stokes", 5)
    assert wf.select_model(5.0, T, D, "Fenton", 12) == ("Fenton", 12)
    assert wf.select_model(5.0, T, D, "Airy", 12) == ("Airy", 1)


def test_raschiiwave_opt_in():
    import jacket_b200 as jb
    w = jb.RaschiiWave.with_own_fits(8.0, T, D, 1.0, "Stokes", 5)
    assert w.kind == "fourier" and w.actual_model == "Stokes" and w.actual_N == 5 and len(w.wave.E) == 5
    assert w.omega == 2 * np.pi / T and w.L == 2 * np.pi / w.k
    w0 = jb.RaschiiWave(8.0, T, D, 1.0, "Stokes", 5)
    assert w0.kind == "airy" and w0.actual_model == "Airy (fallback)"       # default = the pinned reference behaviour
