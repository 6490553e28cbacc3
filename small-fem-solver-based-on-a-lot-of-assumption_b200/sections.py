"""TubularSection -- drop-in for the reference dataclass (GUI.py:115-160).

Attribute names, units (mm) and formulas follow the reference; the stress
evaluation of one load case is a host-side convenience (the per-phase von
Mises field is computed on the GPU by k_member_post).
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

STRESS_POINT_ANGLES = (0, 45, 90, 135, 180, 225, 270, 315)   # GUI.py:142


@dataclass
class TubularSection:
    D_outer: float
    t: float
    name: str = ""
    rho_steel: float = 7850

    def __post_init__(self):
        Do, wall = self.D_outer, self.t
        Di = Do - 2 * wall
        self.D_inner = Di
        self.R_outer, self.R_inner = Do / 2.0, Di / 2.0
        ring2 = Do**2 - Di**2          # area moment kernels of a circular ring
        ring4 = Do**4 - Di**4
        self.Ax_mm2 = np.pi / 4.0 * ring2
        self.Ax_m2 = self.Ax_mm2 / 1e6
        self.Iy_mm4 = self.Iz_mm4 = np.pi / 64.0 * ring4
        self.Ix_mm4 = np.pi / 32.0 * ring4
        self.Ay_mm2 = self.Az_mm2 = 0.5 * self.Ax_mm2        # thin-wall shear areas
        self.Wy_mm3 = self.Iy_mm4 / self.R_outer
        self.Wz_mm3 = self.Iz_mm4 / self.R_outer
        self.Wx_mm3 = self.Ix_mm4 / self.R_outer
        self.mass_per_m = self.Ax_m2 * self.rho_steel
        self.D_t_ratio = Do / wall

    # -- packing for the C ABI (JK_SEC_* order) -------------------------------------
    def prop_row(self):
        return [self.D_outer, self.Ax_mm2, self.Iy_mm4, self.Iz_mm4, self.Ix_mm4,
                self.Ay_mm2, self.Az_mm2, self.R_outer]

    # -- reference helpers ----------------------------------------------------------
    def get_stress_points(self):
        pts = {}
        for i, ang in enumerate(STRESS_POINT_ANGLES, start=1):
            a = np.radians(ang)
            pts[f"A{i}"] = (self.R_outer * np.cos(a), self.R_outer * np.sin(a))
        return pts

    def calc_stress_at_point(self, Fx, Fy, Fz, Mx, My, Mz, point_id):
        y, z = self.get_stress_points()[point_id]
        sigma = Fx / self.Ax_mm2
        if self.Iy_mm4 > 0:
            sigma += My * z / self.Iy_mm4
        if self.Iz_mm4 > 0:
            sigma += Mz * y / self.Iz_mm4
        r = math.hypot(y, z)
        parts = (Mx * r / self.Ix_mm4 if self.Ix_mm4 > 0 else 0.0,
                 Fy / self.Ay_mm2 if self.Ay_mm2 > 0 else 0.0,
                 Fz / self.Az_mm2 if self.Az_mm2 > 0 else 0.0)
        tau = math.sqrt(sum(p * p for p in parts))
        return {"sigma_total": sigma, "tau_total": tau, "von_mises": math.sqrt(sigma**2 + 3.0 * tau**2)}
