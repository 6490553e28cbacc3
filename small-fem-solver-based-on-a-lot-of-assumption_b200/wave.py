"""RaschiiWave -- host side of the wave model (GUI.py:166-296).

Model selection and the dispersion solve stay on the host (they are scalar,
once-per-sea-state work); the kinematics themselves are evaluated on the GPU
(csrc/jk_morison.cuh).  ``raschii`` is a third-party dependency of the
reference (requirements.txt:7) that is absent here, so -- exactly like the
reference without it (GUI.py:187-195) -- every requested model currently runs
the closed-form Airy branch and reports ``'Airy (fallback)'``.  Own Stokes /
Fenton coefficient fits (Fourier-series kinematics) are the next scope row
(SURVEY 8f-1); parity for them is unpinned.
"""
from __future__ import annotations

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
    def __init__(self, H, T, d, U_c=0.0, wave_model="Fenton", N=10, dt=0.001, nonlinear=None):
        self.H, self.T, self.d, self.U_c = H, T, d, U_c
        self.requested_model, self.requested_N = wave_model, N
        self.dt = dt
        self.a = H / 2.0
        self.wave = None
        self.kind = "airy"
        use_fits = NONLINEAR_MODELS_AVAILABLE if nonlinear is None else bool(nonlinear)
        if use_fits:
            from . import wavefit
            name, order = wavefit.select_model(H, T, d, wave_model, N)        # GUI.py:208-253
            self.actual_model, self.actual_N = name, order
            if name == "Stokes":
                fit = wavefit.stokes_fit(H, T, d, order)
            elif name == "Fenton":
                fit = wavefit.fenton_fit(H, T, d, order)
            else:                                                               # Airy through the series form (one harmonic)
                k = wavefit.airy_wavenumber(2.0 * np.pi / T, d)
                om = 2.0 * np.pi / T
                fit = wavefit.FourierFit("Airy", 1, H, T, d, float(k), om, om / k, np.array([H / 2.0]),
                                         np.array([H / 2.0 * om / np.tanh(k * d)]), ubar=om / k)
            self.wave, self.kind = fit, "fourier"
            self.omega, self.k, self.L, self.c = fit.omega, fit.k, fit.length, fit.c     # GUI.py:185
            self.steepness = self.H / self.L
        else:
            self.omega = 2.0 * np.pi / T
            self.k = solve_dispersion(self.omega, d)
            self.L = 2.0 * np.pi / self.k
            self.c = self.L / T
            self.steepness = self.H / self.L
            self.actual_model, self.actual_N = "Airy (fallback)", 1

    def get_model_info(self):
        return f"{self.actual_model} (Order/N={self.actual_N}), Steepness H/L={self.steepness:.4f}"

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
