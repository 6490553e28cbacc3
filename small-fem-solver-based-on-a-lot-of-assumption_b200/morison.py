"""MorisonCalculator -- drop-in for GUI.py:539-724, evaluated on the GPU.

``compute_all_morison_forces`` and ``find_critical_phase`` return the same
dictionaries as the reference; the arithmetic runs in csrc/jk_morison.cuh.
"""
from __future__ import annotations

import numpy as np

from . import _lib as L
from .engine import get_engine


def phase_times(T, n_steps):
    """t_i = i*T/n_steps, evaluated like the reference does (GUI.py:696)."""
    return np.array([i * T / n_steps for i in range(n_steps)], dtype=np.float64)


def fill_phase_deg(table, omega):
    """Column 1 of the table, bit-identical to GUI.py:697-698.  Host mirror of what k_phase_reduce writes on the device
    (the library fills the column itself); kept for callers that build tables from their own times."""
    table[:, 1] = mod360(np.degrees(omega * table[:, 0]))
    return table


def mod360(x):
    """``x % 360`` for float arrays, bit-identical (C fmod, then +360 where the remainder is negative -- what Python's and
    NumPy's float ``%`` do) but twice as fast as NumPy's generic divmod loop."""
    r = np.fmod(x, 360.0)
    neg = r < 0
    if neg.any():
        r[neg] += 360.0
    return r


class MorisonCalculator:
    def __init__(self, structure, wave, wave_direction=0.0, current_direction=0.0,
                 Cd=0.7, Cm=2.0, rho_water=1025):
        self.structure, self.wave = structure, wave
        self.wave_dir_deg, self.current_dir_deg = wave_direction, current_direction
        self.theta_wave = np.deg2rad(90.0 - wave_direction)       # compass -> math angle (GUI.py:555)
        self.theta_current = np.deg2rad(90.0 - current_direction)
        self.Cd, self.Cm, self.rho = Cd, Cm, rho_water

    # ------------------------------------------------------------------------
    def _engine(self, n_gauss=15):
        eng = get_engine(self.structure)
        eng.set_wave(self.wave)
        eng.set_morison(self.theta_wave, self.theta_current, self.rho, self.Cd, self.Cm, n_gauss)
        return eng

    def get_kinematics_3d(self, x, y, z, t):
        """Wave and current velocity / wave acceleration at one point in global axes (GUI.py:559-589): the 2-D wave
        solution acts along the wave heading, the current (taken out of the wave's u again) along its own heading."""
        cw, sw = np.cos(self.theta_wave), np.sin(self.theta_wave)
        plane = self.wave.get_kinematics(x * cw + y * sw, z, t)
        if not plane["submerged"]:
            return dict(u_wave=0, v_wave=0, w_wave=0, u_current=0, v_current=0, du_dt=0, dv_dt=0, dw_dt=0,
                        submerged=False, eta=plane["eta"])
        Uc = self.wave.U_c
        along = plane["u"] - Uc
        return dict(u_wave=along * cw, v_wave=along * sw, w_wave=plane["w"],
                    u_current=Uc * np.cos(self.theta_current), v_current=Uc * np.sin(self.theta_current),
                    du_dt=plane["du_dt"] * cw, dv_dt=plane["du_dt"] * sw, dw_dt=plane["dw_dt"],
                    submerged=True, eta=plane["eta"])

    KINEMATICS_COLUMNS = ("u_wave", "v_wave", "w_wave", "u_current", "v_current", "du_dt", "dv_dt", "dw_dt", "submerged", "eta")

    def kinematics_points(self, points, t):
        """get_kinematics_3d for an [n, 3] array of points at time t, evaluated on the GPU with the device functions of
        the Morison kernels -> [n, 10] array (columns KINEMATICS_COLUMNS)."""
        return self._engine().kinematics_points(points, t)

    def compute_all_morison_forces(self, t=0.0, n_gauss=15):
        eng = self._engine(n_gauss)
        nodal, totals, details = eng.morison_single(t, want_details=True)
        st = self.structure
        nodal_forces = {}
        for i, name in enumerate(st.node_list):
            v = np.zeros(6)
            v[:3] = nodal[i]
            nodal_forces[name] = v
        member_details = [dict(member=m["name"], **{c: details[i, j] for j, c in enumerate(L.DETAIL_COLUMNS)})
                          for i, m in enumerate(st.members)]
        return {"nodal_forces": nodal_forces, "total_drag": totals[0:3].copy(), "total_inertia": totals[3:6].copy(),
                "total_morison": totals[6:9].copy(), "member_details": member_details}

    def scan_table(self, n_steps=36, t=None):
        """The Morison columns of the per-phase table as an array [P, 16] plus the critical index."""
        eng = self._engine()
        t = phase_times(self.wave.T, n_steps) if t is None else np.asarray(t, dtype=np.float64)
        table, crit = eng.morison_scan(t)
        return table, crit

    def find_critical_phase(self, n_steps=36):
        table, crit = self.scan_table(n_steps)
        keys = L.TABLE_COLUMNS[:8]
        rows = [dict(zip(keys, (float(v) for v in table[i, :8]))) for i in range(table.shape[0])]
        return {"all_phases": rows, "critical": rows[crit], "T": self.wave.T, "omega": self.wave.omega}
