"""Host-side fits of periodic nonlinear waves -> Fourier coefficients for the GPU kinematics kernel.

The reference delegates Stokes / Fenton waves to the third-party ``raschii`` package
(requirements.txt:7; call sites GUI.py:212-253, 261, 273), which is absent here.  This module
restates the published algorithms and returns every model in ONE form (the form the reference's
wrapper consumes, GUI.py:259-281):

    eta(x,t) = sum_j E[j] cos(j phi)                                   phi = k x - omega t
    u(x,z,t) = sum_j B[j] cosh(j k zb) / cosh(j k d) cos(j phi)        zb measured from the sea bed
    w(x,z,t) = sum_j B[j] sinh(j k zb) / cosh(j k d) sin(j phi)        j = 1..N

with zero mean Eulerian current (the reference adds U_c afterwards, GUI.py:281).

* ``stokes_fit``  -- Fenton (1985) "A fifth-order Stokes theory for steady waves", orders 1..5.
* ``fenton_fit``  -- Fourier approximation / stream-function method (Rienecker & Fenton 1981,
                     Fenton 1988), Newton iteration with height stepping, any N.

PARITY UNPINNED: there is no raschii here to compare with.  ``bc_residuals`` measures how well a
fit satisfies the kinematic and dynamic free-surface conditions; tests assert the expected
convergence (Stokes residual ~ eps^(N+1), Fenton residual ~ 1e-10 at the collocation points) and
the reduction to linear theory as H -> 0.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

G = 9.81


@dataclass
class FourierFit:
    model: str
    N: int
    H: float
    T: float
    d: float
    k: float
    omega: float
    c: float
    E: np.ndarray          # eta harmonics, m
    B: np.ndarray          # velocity harmonics, m/s (already contain j k and the 1/cosh normalisation convention above)
    ubar: float = 0.0      # mean fluid speed in the wave frame (= c for zero Eulerian current)
    Q: float = 0.0
    R: float = 0.0

    @property
    def length(self):
        return 2.0 * np.pi / self.k

    def eta(self, phi):
        j = np.arange(1, self.N + 1)
        return np.sum(self.E[:, None] * np.cos(j[:, None] * np.atleast_1d(phi)[None, :]), axis=0)

    def velocity(self, phi, zb):
        """(u, w) in the fixed frame at bed-relative height zb (no clamping, no current)."""
        phi = np.atleast_1d(np.asarray(phi, dtype=np.float64)); zb = np.broadcast_to(np.asarray(zb, dtype=np.float64), phi.shape)
        u = np.zeros_like(phi); w = np.zeros_like(phi)
        for j in range(1, self.N + 1):
            den = np.cosh(j * self.k * self.d)
            u += self.B[j - 1] * np.cosh(j * self.k * zb) / den * np.cos(j * phi)
            w += self.B[j - 1] * np.sinh(j * self.k * zb) / den * np.sin(j * phi)
        return u, w


def airy_wavenumber(omega, d, g=G):
    k = omega**2 / g
    for _ in range(100):
        th = np.tanh(k * d)
        f = omega**2 - g * k * th
        df = -g * (th + k * d / np.cosh(k * d)**2)
        dk = f / df
        k -= dk
        if abs(dk) < 1e-14 * k:
            break
    return k


# ---------------------------------------------------------------------------------------------
# Stokes, orders 1..5 (Fenton 1985, table of coefficients in S = sech(2kd))
# ---------------------------------------------------------------------------------------------
def _stokes_tables(kd):
    S = 1.0 / np.cosh(2.0 * kd)
    sh, th = np.sinh(kd), np.tanh(kd)
    coth = 1.0 / th
    A = {}
    A[1, 1] = 1.0 / sh
    A[2, 2] = 3.0 * S**2 / (2.0 * (1 - S)**2)
    A[3, 1] = (-4 - 20 * S + 10 * S**2 - 13 * S**3) / (8 * sh * (1 - S)**3)
    A[3, 3] = (-2 * S**2 + 11 * S**3) / (8 * sh * (1 - S)**3)
    A[4, 2] = (12 * S - 14 * S**2 - 264 * S**3 - 45 * S**4 - 13 * S**5) / (24 * (1 - S)**5)
    A[4, 4] = (10 * S**3 - 174 * S**4 + 291 * S**5 + 278 * S**6) / (48 * (3 + 2 * S) * (1 - S)**5)
    A[5, 1] = (-1184 + 32 * S + 13232 * S**2 + 21712 * S**3 + 20940 * S**4 + 12554 * S**5 - 500 * S**6 - 3341 * S**7
               - 670 * S**8) / (64 * sh * (3 + 2 * S) * (4 + S) * (1 - S)**6)
    A[5, 3] = (4 * S + 105 * S**2 + 198 * S**3 - 1376 * S**4 - 1302 * S**5 - 117 * S**6 + 58 * S**7) / (
        32 * sh * (3 + 2 * S) * (1 - S)**6)
    A[5, 5] = (-6 * S**3 + 272 * S**4 - 1552 * S**5 + 852 * S**6 + 2029 * S**7 + 430 * S**8) / (
        64 * sh * (3 + 2 * S) * (4 + S) * (1 - S)**6)
    Bc = {}
    Bc[2, 2] = coth * (1 + 2 * S) / (2 * (1 - S))
    Bc[3, 1] = -3 * (1 + 3 * S + 3 * S**2 + 2 * S**3) / (8 * (1 - S)**3)
    Bc[4, 2] = coth * (6 - 26 * S - 182 * S**2 - 204 * S**3 - 25 * S**4 + 26 * S**5) / (6 * (3 + 2 * S) * (1 - S)**4)
    Bc[4, 4] = coth * (24 + 92 * S + 122 * S**2 + 66 * S**3 + 67 * S**4 + 34 * S**5) / (24 * (3 + 2 * S) * (1 - S)**4)
    Bc[5, 3] = 9 * (132 + 17 * S - 2216 * S**2 - 5897 * S**3 - 6292 * S**4 - 2687 * S**5 + 194 * S**6 + 467 * S**7
                    + 82 * S**8) / (128 * (3 + 2 * S) * (4 + S) * (1 - S)**6)
    Bc[5, 5] = 5 * (300 + 1579 * S + 3176 * S**2 + 2949 * S**3 + 1188 * S**4 + 675 * S**5 + 1326 * S**6 + 827 * S**7
                    + 130 * S**8) / (384 * (3 + 2 * S) * (4 + S) * (1 - S)**6)
    C0 = np.sqrt(th)
    C2 = np.sqrt(th) * (2 + 7 * S**2) / (4 * (1 - S)**2)
    C4 = np.sqrt(th) * (4 + 32 * S - 116 * S**2 - 400 * S**3 - 71 * S**4 + 146 * S**5) / (32 * (1 - S)**5)
    return A, Bc, (C0, C2, C4)


def stokes_fit(H, T, d, N=5, g=G):
    N = int(min(max(N, 1), 5))
    omega = 2.0 * np.pi / T

    def speed_residual(k):
        eps = k * H / 2.0
        _, _, (C0, C2, C4) = _stokes_tables(k * d)
        ub = C0 + (eps**2 * C2 if N >= 3 else 0.0) + (eps**4 * C4 if N >= 5 else 0.0)
        return np.sqrt(g / k) * ub - omega / k          # zero Eulerian current: c = ubar

    k = airy_wavenumber(omega, d, g)
    for _ in range(100):                                 # secant / Newton with numerical slope
        f = speed_residual(k)
        h = 1e-6 * k
        df = (speed_residual(k + h) - speed_residual(k - h)) / (2 * h)
        dk = f / df
        k -= dk
        if abs(dk) < 1e-14 * k:
            break
    eps = k * H / 2.0
    A, Bc, (C0, C2, C4) = _stokes_tables(k * d)
    # k*eta harmonics (Fenton 1985 eq. 14), truncated to order N
    ke = np.zeros(6)
    ke[1] = eps
    if N >= 2: ke[2] += eps**2 * Bc[2, 2]
    if N >= 3: ke[1] += eps**3 * Bc[3, 1]; ke[3] += -eps**3 * Bc[3, 1]
    if N >= 4: ke[2] += eps**4 * Bc[4, 2]; ke[4] += eps**4 * Bc[4, 4]
    if N >= 5: ke[1] += -eps**5 * (Bc[5, 3] + Bc[5, 5]); ke[3] += eps**5 * Bc[5, 3]; ke[5] += eps**5 * Bc[5, 5]
    E = ke[1:N + 1] / k
    # velocity harmonics: u = C0 sqrt(g/k) sum_i eps^i sum_j j A_ij cosh(j k z) cos(j phi)
    a = np.zeros(6)
    for (i, j), v in A.items():
        if i <= N:
            a[j] += eps**i * v
    jj = np.arange(1, N + 1)
    B = C0 * np.sqrt(g / k) * jj * a[1:N + 1] * np.cosh(jj * k * d)
    ub = np.sqrt(g / k) * (C0 + (eps**2 * C2 if N >= 3 else 0.0) + (eps**4 * C4 if N >= 5 else 0.0))
    return FourierFit("Stokes", N, H, T, d, float(k), omega, omega / k, E, B, ubar=float(ub))


# ---------------------------------------------------------------------------------------------
# Fourier approximation (stream function) method
# ---------------------------------------------------------------------------------------------
def _fenton_residual(x, H, T, d, N, g):
    """Unknowns x = [B_1..B_N, eta_0..eta_N, ubar, k, Q, R]; collocation on half a wave length."""
    B = x[:N]; eta = x[N:2 * N + 1]; ubar, k, Q, R = x[2 * N + 1:]
    m = np.arange(N + 1)
    j = np.arange(1, N + 1)
    ph = m * np.pi / N                                   # k X_m
    zb = d + eta
    jk = j * k
    den = np.cosh(jk * d)
    S = np.sinh(np.outer(zb, jk)) / den                  # [m, j]
    C = np.cosh(np.outer(zb, jk)) / den
    cosj = np.cos(np.outer(ph, j)); sinj = np.sin(np.outer(ph, j))
    psi = -ubar * zb + (S * cosj) @ B
    U = -ubar + (C * cosj) @ (B * jk)
    W = (S * sinj) @ (B * jk)
    f = np.empty(2 * N + 5)
    f[:N + 1] = psi + Q                                  # kinematic condition
    f[N + 1:2 * N + 2] = 0.5 * (U**2 + W**2) + g * eta - R   # dynamic condition
    f[2 * N + 2] = (eta[0] + eta[-1] + 2.0 * eta[1:-1].sum()) / (2.0 * N)   # mean level
    f[2 * N + 3] = eta[0] - eta[-1] - H                  # wave height
    f[2 * N + 4] = k * ubar * T - 2.0 * np.pi            # period, zero Eulerian current (c = ubar)
    return f


def fenton_fit(H, T, d, N=10, g=G, n_steps=None, tol=1e-11, max_iter=60):
    N = int(N)
    omega = 2.0 * np.pi / T
    k0 = airy_wavenumber(omega, d, g)
    steep = H * k0 / (2 * np.pi) / (0.142 * np.tanh(k0 * d))          # fraction of the breaking limit
    if n_steps is None:
        n_steps = int(np.clip(np.ceil(12 * steep), 3, 16))
    m = np.arange(N + 1)
    x = np.zeros(2 * N + 5)
    h0 = H / n_steps
    x[0] = (h0 / 2.0) * (omega / k0) / np.tanh(k0 * d)
    x[N:2 * N + 1] = (h0 / 2.0) * np.cos(m * np.pi / N)
    c0 = omega / k0
    x[2 * N + 1:] = [c0, k0, c0 * d, 0.5 * c0**2]
    scale = np.concatenate([np.full(N, c0 / k0), np.full(N + 1, 1.0), [c0, k0, c0 * d, c0**2]])
    x_prev = None
    for step in range(1, n_steps + 1):
        Hs = H * step / n_steps
        if x_prev is not None:                            # linear extrapolation in height
            x, x_prev = 2.0 * x - x_prev, x.copy()
        else:
            x_prev = x.copy()
        for it in range(max_iter):
            f = _fenton_residual(x, Hs, T, d, N, g)
            J = np.empty((len(x), len(x)))
            for q in range(len(x)):
                hq = 1e-7 * scale[q]
                xp = x.copy(); xp[q] += hq
                xm = x.copy(); xm[q] -= hq
                J[:, q] = (_fenton_residual(xp, Hs, T, d, N, g) - _fenton_residual(xm, Hs, T, d, N, g)) / (2 * hq)
            dx = np.linalg.solve(J, -f)
            x = x + dx
            if np.max(np.abs(dx) / scale) < tol:
                break
        else:
            raise RuntimeError(f"fenton_fit: Newton did not converge at height step {step}/{n_steps} (H={Hs:.3f})")
    Bs = x[:N]; eta = x[N:2 * N + 1]; ubar, k, Q, R = x[2 * N + 1:]
    # eta harmonics from the collocation values (discrete cosine transform on the half wave, trapezoid rule)
    j = np.arange(1, N + 1)
    wts = np.ones(N + 1); wts[0] = wts[-1] = 0.5
    E = (2.0 / N) * (np.cos(np.outer(j, m * np.pi / N)) * (wts * eta)).sum(axis=1)
    E[-1] *= 0.5                                          # Nyquist term
    Bv = Bs * j * k
    return FourierFit("Fenton", N, H, T, d, float(k), omega, float(omega / k), E, Bv, ubar=float(ubar), Q=float(Q), R=float(R))


# ---------------------------------------------------------------------------------------------
def bc_residuals(fit: FourierFit, n=64, g=G):
    """Max kinematic / dynamic free-surface residuals over a wave length, relative to c*H and g*H.

    In the frame moving with the wave the flow is steady: the surface must be a streamline
    ( (U, W) . normal = 0 ) and Bernoulli's constant must not vary along it."""
    phi = np.linspace(0.0, 2.0 * np.pi, n, endpoint=False)
    jv = np.arange(1, fit.N + 1)
    eta = fit.eta(phi)
    deta = -np.sum((fit.E * jv * fit.k)[:, None] * np.sin(jv[:, None] * phi[None, :]), axis=0)   # d eta / dX
    u, w = fit.velocity(phi, fit.d + eta)
    U = u - fit.c                                          # steady frame
    kin = (w - U * deta) / np.sqrt(1 + deta**2)
    bern = 0.5 * (U**2 + w**2) + g * eta
    return float(np.max(np.abs(kin)) / (fit.c * fit.H * fit.k)), float((bern.max() - bern.min()) / (g * fit.H))


def select_model(H, T, d, model, N, g=G):
    """The reference's model choice (GUI.py:208-253) -> (name, order).  Steepness uses the Airy length."""
    steep = H / (2.0 * np.pi / airy_wavenumber(2.0 * np.pi / T, d, g))
    m = model.lower()
    if m == "auto":
        if steep < 0.01:
            return "Airy", 1
        if steep < 0.03:
            return "Stokes", 3
        if steep < 0.06:
            return "Stokes", 5
        return "Fenton", min(max(int(steep * 200), 10), 20)
    if m == "fenton":
        return "Fenton", N
    if m == "stokes":
        return "Stokes", min(N, 5)
    return "Airy", 1
