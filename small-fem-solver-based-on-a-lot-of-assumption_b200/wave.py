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
NONLINEAR_MODELS_AVAILABLE = False


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
    def __init__(self, H, T, d, U_c=0.0, wave_model="Fenton", N=10, dt=0.001):
        self.H, self.T, self.d, self.U_c = H, T, d, U_c
        self.requested_model, self.requested_N = wave_model, N
        self.dt = dt
        self.a = H / 2.0
        self.wave = None
        self.kind = "airy"
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

    def signature(self):
        return ("airy",) + self.device_args()
