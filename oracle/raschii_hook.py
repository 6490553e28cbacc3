"""raschii parity hook (TEST INFRASTRUCTURE ONLY).

The reference delegates Stokes / Fenton kinematics to the third-party package ``raschii`` (requirements.txt:7,
``raschii>=1.0.0``, no lock file; call sites GUI.py:212-253, 261, 273).  It is not installable in the build container
or on the GPU box (no network), so the own fits of jacket_b200.wavefit and the Fourier-series Morison kernel are
"parity unpinned" there.  Wherever ``import raschii`` succeeds this module activates:

* ``reference_wave(...)``  the reference's own ``RaschiiWave`` on raschii when /root/reference is present too, else a
  wrapper around the raschii object that restates GUI.py:259-281 (elevation from the bed minus d, the z clamp, U_c
  added to a wet point, dry rule, forward-difference acceleration);
* ``velocity_fn(wave)``    adapter for oracle.jacket_oracle.morison_phases / phase_scan (vectorised over points);
* tests/test_raschii_hook.py compares wavefit's fields with raschii's and the CUDA Fourier kernel with the oracle
  driven by raschii; tests/golden/make_raschii_golden.py writes the vectors so that they travel to boxes without it.
"""
from __future__ import annotations

import numpy as np


def available() -> bool:
    try:
        import raschii  # noqa: F401
        return True
    except Exception:
        return False


class WrappedRaschii:
    """GUI.py:171-195 + 259-288 around a raschii wave object (used when the reference file itself is absent)."""

    def __init__(self, H, T, d, U_c=0.0, wave_model="Fenton", N=10, dt=0.001):
        import raschii
        self.H, self.T, self.d, self.U_c, self.dt = H, T, d, U_c, dt
        self.a = H / 2.0
        m = wave_model.lower()
        steep = H / raschii.AiryWave(height=H, depth=d, period=T).length          # GUI.py:212-213
        if m == "auto":                                                            # GUI.py:215-238
            if steep < 0.01:
                m, N = "airy", 1
            elif steep < 0.03:
                m, N = "stokes", 3
            elif steep < 0.06:
                m, N = "stokes", 5
            else:
                m, N = "fenton", min(max(int(steep * 200), 10), 20)
        if m == "fenton":
            self.wave, self.actual_model, self.actual_N = raschii.FentonWave(height=H, depth=d, period=T, N=N), "Fenton", N
        elif m == "stokes":
            n = min(N, 5)
            self.wave, self.actual_model, self.actual_N = raschii.StokesWave(height=H, depth=d, period=T, N=n), "Stokes", n
        else:
            self.wave, self.actual_model, self.actual_N = raschii.AiryWave(height=H, depth=d, period=T), "Airy", 1
        self.omega, self.k, self.L, self.c = self.wave.omega, self.wave.k, self.wave.length, self.wave.c
        self.steepness = H / self.L

    def eta(self, x, t=0.0):
        e = self.wave.surface_elevation(x, t=t)
        return (e[0] if hasattr(e, "__len__") else e) - self.d

    def velocity(self, x, z_mwl, t=0.0):
        surf = self.eta(x, t)
        if z_mwl > surf:
            return (0.0, 0.0)
        zb = max(0.01, min(z_mwl + self.d, self.d + surf - 0.01))
        v = np.asarray(self.wave.velocity(x, zb, t=t)).reshape(-1)
        return (float(v[0]) + self.U_c, float(v[1]))

    def acceleration(self, x, z_mwl, t=0.0):
        if z_mwl > self.eta(x, t):
            return (0.0, 0.0)
        a, b = self.velocity(x, z_mwl, t), self.velocity(x, z_mwl, t + self.dt)
        return ((b[0] - a[0]) / self.dt, (b[1] - a[1]) / self.dt)


def reference_wave(H, T, d, U_c=0.0, wave_model="Fenton", N=10, dt=0.001):
    """The reference's RaschiiWave running on raschii (preferred), else the restated wrapper."""
    from . import ref_loader
    if ref_loader.available():
        ref = ref_loader.load(allow_raschii=True)
        if getattr(ref, "RASCHII_AVAILABLE", False):
            return ref.RaschiiWave(H, T, d, U_c, wave_model, N, dt)
    return WrappedRaschii(H, T, d, U_c, wave_model, N, dt)


def velocity_fn(wave_like):
    """(wave, xw, z, t) -> (u, w, wet) over arrays, for oracle.morison_phases(velocity_fn=...): point-by-point calls of
    the wrapper's scalar velocity()."""
    def fn(_wave, xw, z, t):
        shape = np.broadcast(xw, z, t).shape
        xw_, z_, t_ = (np.broadcast_to(np.asarray(v, dtype=np.float64), shape).ravel() for v in (xw, z, t))
        u = np.zeros(xw_.shape); w = np.zeros(xw_.shape); wet = np.zeros(xw_.shape, dtype=bool)
        for i in range(xw_.size):
            wet[i] = not (z_[i] > wave_like.eta(float(xw_[i]), float(t_[i])))
            if wet[i]:
                u[i], w[i] = wave_like.velocity(float(xw_[i]), float(z_[i]), float(t_[i]))
        return u.reshape(shape), w.reshape(shape), wet.reshape(shape)
    return fn


def sample_fields(wave_like, n_x=48, n_z=12, t=0.0):
    """eta over one wave length and (u, w) without current on an (x, z) grid below the troughs -- what wavefit must reproduce."""
    L, d = wave_like.L, wave_like.d
    x = np.linspace(0.0, L, n_x, endpoint=False)
    eta = np.array([wave_like.eta(float(xx), t) for xx in x])
    z = np.linspace(-d + 0.05, eta.min() - 0.05, n_z)
    uw = np.array([[wave_like.velocity(float(xx), float(zz), t) for zz in z] for xx in x])
    uw[:, :, 0] -= wave_like.U_c
    return x, z, eta, uw
