"""RaschiiWave -- host side of the wave model (GUI.py:166-296).

Model selection and the dispersion solve stay on the host (scalar, once-per-sea-state work).  The bulk kinematics
-- every Gauss point of every member at every phase -- are evaluated on the GPU (csrc/jk_morison.cuh); the scalar
point queries of the reference's class surface (``eta``, ``velocity``, ``acceleration``, ``get_kinematics``,
GUI.py:259-296) are answered here with the same formulas, and tests/test_gpu_parity.py checks them against the
device evaluation of the same points (``MorisonCalculator.kinematics_points`` -> jk_kinematics_points).

``raschii`` is a third-party dependency of the reference (requirements.txt:7) that is absent here, so -- exactly
like the reference without it (GUI.py:187-195) -- every requested model runs the closed-form Airy branch and reports
``'Airy (fallback)'`` unless the own Stokes / Fenton fits (wavefit.py, Fourier-series kernel) are switched on with
``enable_nonlinear_waves()`` or ``RaschiiWave.with_own_fits(...)``; parity for those is unpinned here and is
checked against raschii by tests/test_raschii_hook.py wherever that package can be imported.
"""
from __future__ import annotations

import math

import numpy as np

g = 9.81                       # GUI.py:105

# Mirror of the reference's RASCHII_AVAILABLE switch (GUI.py:96-100).  False (default) reproduces the reference as it
# runs where raschii is not installed -- every model is the closed-form Airy fallback, which is the parity-pinned
# behaviour.  True routes 'Stokes' / 'Fenton' / 'auto' through this repo's own fits (wavefit.py) and the
# Fourier-series kernel; 'Airy' then also takes the wrapper's raschii branch semantics (z clamp, GUI.py:272).
NONLINEAR_MODELS_AVAILABLE = False


def enable_nonlinear_waves(flag=True):
    """Opt in to / out of the own Stokes / Fenton fits (parity unpinned: raschii is absent here)."""
    global NONLINEAR_MODELS_AVAILABLE
    NONLINEAR_MODELS_AVAILABLE = bool(flag)


def solve_dispersion(omega, d, gravity=g, tol=1e-10, max_iter=50):
    """Newton iteration for omega^2 = g k tanh(k d), deep-water start (GUI.py:197-206)."""
    w2 = omega**2
    k = w2 / gravity
    for _ in range(max_iter):
        th = np.tanh(k * d)
        resid = w2 - gravity * k * th
        slope = -gravity * (th + k * d / np.cosh(k * d)**2)
        k_next = k - resid / slope
        if abs(k_next - k) < tol:
            break
        k = k_next
    return k


class RaschiiWave:
    def __init__(self, H, T, d, U_c=0.0, wave_model='Fenton', N=10, dt=0.001):
        self._setup(H, T, d, U_c, wave_model, N, dt, NONLINEAR_MODELS_AVAILABLE)

    @classmethod
    def with_own_fits(cls, H, T, d, U_c=0.0, wave_model='Fenton', N=10, dt=0.001):
        """Same wave with this repo's Stokes / Fenton fits switched on for this object only."""
        self = cls.__new__(cls)
        self._setup(H, T, d, U_c, wave_model, N, dt, True)
        return self

    def _setup(self, H, T, d, U_c, wave_model, N, dt, use_fits):
        self.H, self.T, self.d, self.U_c = H, T, d, U_c
        self.requested_model, self.requested_N = wave_model, N
        self.dt = dt
        self.a = H / 2.0
        self.actual_model = self.actual_N = self.steepness = None
        self.wave = None
        self.kind = "airy"
        if use_fits:
            fit = self._create_wave(wave_model, N)
            self.wave, self.kind = fit, "fourier"
            self.omega, self.k, self.L, self.c = fit.omega, fit.k, fit.length, fit.c     # GUI.py:185
        else:
            self.omega = 2.0 * np.pi / T
            self.k = self._solve_dispersion(self.omega, d)
            self.L = 2.0 * np.pi / self.k
            self.c = self.L / T
            self.actual_model, self.actual_N = "Airy (fallback)", 1
        self.steepness = self.H / self.L

    def _solve_dispersion(self, omega, d):
        return solve_dispersion(omega, d)

    def _create_wave(self, model, N):
        """Model choice of GUI.py:208-253 (thresholds on the Airy steepness) -> a fitted periodic wave in series form
        (wavefit.FourierFit: .omega .k .length .c like the raschii objects the reference gets here)."""
        from . import wavefit
        name, order = wavefit.select_model(self.H, self.T, self.d, model, N)
        self.actual_model, self.actual_N = name, order
        if name == "Stokes":
            return wavefit.stokes_fit(self.H, self.T, self.d, order)
        if name == "Fenton":
            return wavefit.fenton_fit(self.H, self.T, self.d, order)
        om = 2.0 * np.pi / self.T                                           # Airy through the series form: one harmonic
        k = wavefit.airy_wavenumber(om, self.d)
        return wavefit.FourierFit("Airy", 1, self.H, self.T, self.d, float(k), om, om / k, np.array([self.H / 2.0]),
                                  np.array([self.H / 2.0 * om / np.tanh(k * self.d)]), ubar=om / k)

    def get_model_info(self):
        return f"{self.actual_model} (Order/N={self.actual_N}), Steepness H/L={self.steepness:.4f}"

    # -- scalar point queries (GUI.py:259-296) ----------------------------------------------------------------------
    def _theta(self, x, t):
        return self.k * x - self.omega * t

    def eta(self, x, t=0.0):
        """Free-surface elevation about the mean water level at wave coordinate x (GUI.py:259-265)."""
        th = self._theta(x, t)
        if self.wave is None:
            return self.a * np.cos(th)
        return float(self.wave.eta(th)[0])           # series elevation is already about MWL (the reference subtracts d from raschii's bed-based value)

    def velocity(self, x, z_mwl, t=0.0):
        """(u, w) at height z_mwl (z = 0 at MWL); a point above the instantaneous surface is dry -> (0, 0), and the
        current is only added to a wet point (GUI.py:267-281)."""
        surface = self.eta(x, t)
        if z_mwl > surface:
            return (0.0, 0.0)
        th = self._theta(x, t)
        if self.wave is None:
            amp = self.a * self.omega / np.sinh(self.k * self.d)
            arg = self.k * (z_mwl + self.d)
            u, w = amp * np.cosh(arg) * np.cos(th), amp * np.sinh(arg) * np.sin(th)
        else:
            zb = max(0.01, min(z_mwl + self.d, self.d + surface - 0.01))    # bed-based height, clamped (GUI.py:272)
            uu, ww = self.wave.velocity(th, zb)
            u, w = float(uu[0]), float(ww[0])
        return (u + self.U_c, w)

    def acceleration(self, x, z_mwl, t=0.0):
        """Forward difference of velocity() over dt; the dry test of velocity() applies at t + dt as well, so a point
        that leaves the water within dt gets -v(t)/dt (GUI.py:283-288)."""
        if z_mwl > self.eta(x, t):
            return (0.0, 0.0)
        now, later = self.velocity(x, z_mwl, t), self.velocity(x, z_mwl, t + self.dt)
        return ((later[0] - now[0]) / self.dt, (later[1] - now[1]) / self.dt)

    def get_kinematics(self, x, z_mwl, t=0.0):
        surface = self.eta(x, t)
        wet = not (z_mwl > surface)
        u = w = du = dw = 0
        if wet:
            u, w = self.velocity(x, z_mwl, t)
            du, dw = self.acceleration(x, z_mwl, t)
        return {"u": u, "w": w, "du_dt": du, "dw_dt": dw, "submerged": wet, "eta": surface}

    # -- device arguments -------------------------------------------------------------------------------------------
    def device_args(self):
        """Arguments of jk_set_wave_airy."""
        return (float(self.a), float(self.k), float(self.omega), float(self.d), float(self.U_c), float(self.dt))

    def fourier_args(self):
        """(k, omega, d, U_c, dt, E[], B[]) of jk_set_wave_fourier."""
        f = self.wave
        return (float(f.k), float(f.omega), float(self.d), float(self.U_c), float(self.dt),
                np.ascontiguousarray(f.E, dtype=np.float64), np.ascontiguousarray(f.B, dtype=np.float64))

    def signature(self):
        if self.kind == "fourier":
            a = self.fourier_args()
            return ("fourier",) + a[:5] + (a[5].tobytes(), a[6].tobytes())
        return ("airy",) + self.device_args()
