"""raschii is absent here, so the parity hook (tests/test_raschii_hook.py) skips.  This file runs the hook's own code
against a STAND-IN module with raschii's call surface (the attributes and methods the reference uses: GUI.py:185, 212-253,
261, 273) built on jacket_b200.wavefit -- it proves the hook executes, not parity."""
import sys
import types

import numpy as np
import pytest


class _StandIn:
    def __init__(self, fit, depth):
        self._fit, self._d = fit, depth
        self.omega, self.k, self.length, self.c = fit.omega, fit.k, fit.length, fit.c

    def surface_elevation(self, x, t=0.0):
        x = np.atleast_1d(np.asarray(x, dtype=float))
        return self._fit.eta(self.k * x - self.omega * t) + self._d            # measured from the bed, like raschii

    def velocity(self, x, z, t=0.0):
        u, w = self._fit.velocity(np.atleast_1d(self.k * np.asarray(x, dtype=float) - self.omega * t), z)
        return np.stack([u, w], axis=1)


@pytest.fixture
def fake_raschii(monkeypatch):
    from jacket_b200 import wavefit
    mod = types.ModuleType("raschii")

    def airy(height, depth, period):
        om = 2 * np.pi / period
        k = wavefit.airy_wavenumber(om, depth)
        return _StandIn(wavefit.FourierFit("Airy", 1, height, period, depth, float(k), om, om / k, np.array([height / 2]),
                                           np.array([height / 2 * om / np.tanh(k * depth)])), depth)
    mod.AiryWave = airy
    mod.StokesWave = lambda height, depth, period, N=5: _StandIn(wavefit.stokes_fit(height, period, depth, N), depth)
    mod.FentonWave = lambda height, depth, period, N=10: _StandIn(wavefit.fenton_fit(height, period, depth, N), depth)
    monkeypatch.setitem(sys.modules, "raschii", mod)
    return mod


def test_hook_runs_against_a_stand_in(fake_raschii):
    import jacket_b200 as jb
    from oracle import raschii_hook
    import test_raschii_hook as hook_tests
    assert raschii_hook.available()
    w = raschii_hook.WrappedRaschii(8.0, 9.4, 50.0, 0.7, "auto", 10)
    assert (w.actual_model, w.actual_N) == ("Stokes", 5)
    ours = jb.RaschiiWave.with_own_fits(8.0, 9.4, 50.0, 0.7, "auto", 10)
    top = w.eta(3.0, 1.0)
    assert abs(top - ours.eta(3.0, 1.0)) < 1e-12
    assert w.velocity(3.0, top + 0.1, 1.0) == (0.0, 0.0) and ours.velocity(3.0, top + 0.1, 1.0) == (0.0, 0.0)
    a, b = w.acceleration(3.0, -5.0, 1.0), ours.acceleration(3.0, -5.0, 1.0)
    assert abs(a[0] - b[0]) < 1e-9 * abs(a[0]) + 1e-12 and abs(a[1] - b[1]) < 1e-9 * abs(a[1]) + 1e-12
    # the parity test body itself (fields, model choice, wrapper semantics)
    for case in hook_tests.CASES[:3]:
        hook_tests.test_own_fit_reproduces_raschii_fields(*case)
    fn = raschii_hook.velocity_fn(w)
    u, ww, wet = fn(None, np.array([[0.0, 3.0]]), np.array([[-5.0, 50.0]]), np.array([[1.0], [2.0]]))
    assert u.shape == (2, 2) and wet[:, 0].all() and not wet[:, 1].any() and (u[:, 1] == 0).all()
