"""Analysis drivers.

``run_analysis``   headless replay of JacketAnalysisGUI.run_analysis (GUI.py:1827-2082):
                   one load case at ``t_analysis`` (+ optional 36-step Morison scan),
                   returns the GUI's ``analysis_results`` dictionary.
``phase_scan``     the product of this repo: the same pipeline evaluated for EVERY phase
                   t_i = i T / P on the GPU (Morison -> loads -> multi-RHS solve ->
                   reactions / member forces / utilisation -> critical phase), returning
                   a per-phase result table; full rows of any phase are fetched on demand.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from . import _lib as L
from .engine import get_engine
from .fem import FEMSolver, member_rows_to_dicts
from .morison import MorisonCalculator, phase_times
from .sections import TubularSection
from .structure import CustomJacketStructure
from .wave import RaschiiWave, g


@dataclass
class AnalysisParams:
    """The GUI's entry fields with their defaults (GUI.py:1805-1817)."""
    E: float = 210000.0
    nu: float = 0.3
    fy: float = 355.0
    rho_steel: float = 7850.0
    rho_water: float = 1025.0
    D_leg: float = 2000.0
    t_leg: float = 75.0
    D_brace: float = 800.0
    t_brace: float = 30.0
    H: float = 17.038
    T: float = 9.4
    d: float = 50.0
    U_c: float = 1.7
    wave_dir: float = 38.0
    current_dir: float = 38.0
    wave_model: str = "Fenton"
    N_harm: int = 10
    Cd: float = 0.7
    Cm: float = 2.0
    F_axial: float = 25100.0      # kN
    F_shear: float = 2900.0       # kN
    M_moment: float = 0.0         # kNm
    M_torsion: float = 0.0        # kNm
    self_weight_mode: str = "calculated"   # 'calculated' | 'custom' | 'none'
    custom_sw: float = 1100.0     # tonnes
    t_analysis: float = 0.0
    do_phase_scan: bool = True
    n_phase_steps: int = 36       # GUI.py:1933


def interface_load_vector(wave_dir, n_legs, F_axial, F_shear, M_moment, M_torsion):
    """Per-top-node load of GUI.py:1962-1976 (kN, kNm in -> N, N mm out)."""
    theta = np.deg2rad(90.0 - wave_dir)
    Fa, Fs = F_axial * 1000.0, F_shear * 1000.0
    Mm, Mt = M_moment * 1e6, M_torsion * 1e6
    return np.array([Fs * np.cos(theta) / n_legs, Fs * np.sin(theta) / n_legs, -Fa / n_legs,
                     Mt / n_legs, Mm / n_legs, 0.0])


def apply_self_weight(F, structure, mode, custom_sw=0.0):
    """Self-weight exactly as run_analysis writes it into F_global (GUI.py:1994-2015). Returns total N."""
    if mode == "calculated":
        total = 0.0
        for m in structure.members:
            geom = structure.get_member_geometry(m)
            weight = m["section"].mass_per_m * g * geom["L"]
            total += weight
            half = weight / 2.0
            F[6 * structure.node_index[m["node1"]] + 2] -= half
            F[6 * structure.node_index[m["node2"]] + 2] -= half
        return total
    if mode == "custom":
        total = custom_sw * 1000 * g
        per_node = total / structure.n_nodes
        for i in range(structure.n_nodes):
            F[6 * i + 2] -= per_node
        return total
    return 0.0


def static_load(structure, p: AnalysisParams):
    """Phase-independent part of F_global: interface loads then self-weight.  Memoised per structure and load
    parameters (the self-weight loop walks every member in Python: 10 ms at 2k members, once per scan otherwise)."""
    key = (p.wave_dir, p.F_axial, p.F_shear, p.M_moment, p.M_torsion, p.self_weight_mode, p.custom_sw, p.rho_steel)
    cache = structure.__dict__.setdefault("_static_load_cache", {})
    if key not in cache:
        if len(cache) > 64:
            cache.clear()
        F = _static_load(structure, p)
        F.setflags(write=False)
        cache[key] = F
    return cache[key].copy()


def _static_load(structure, p: AnalysisParams):
    F = np.zeros(structure.n_dof)
    top = structure.get_top_nodes()
    vec = interface_load_vector(p.wave_dir, len(top), p.F_axial, p.F_shear, p.M_moment, p.M_torsion)
    for name in top:
        i = structure.node_index[name]
        F[6 * i:6 * i + 6] += vec
    apply_self_weight(F, structure, p.self_weight_mode, p.custom_sw)
    return F


def build_structure(nodes, members, fixed_nodes, top_nodes, p: AnalysisParams):
    leg = TubularSection(p.D_leg, p.t_leg, "Leg", p.rho_steel)
    brace = TubularSection(p.D_brace, p.t_brace, "Brace", p.rho_steel)
    nodes_np = {k: np.array(v, dtype=np.float64) for k, v in nodes.items()}
    return CustomJacketStructure(nodes_np, members, leg, brace, fixed_nodes, top_nodes, p.rho_steel)


def run_analysis(nodes, members, fixed_nodes, top_nodes, params: AnalysisParams | None = None, log=None):
    """Headless equivalent of the GUI's analysis step; returns ``analysis_results`` (GUI.py:2061-2067)."""
    p = params or AnalysisParams()
    say = log or (lambda *_: None)
    structure = build_structure(nodes, members, fixed_nodes, top_nodes, p)
    wave = RaschiiWave(p.H, p.T, p.d, p.U_c, p.wave_model, p.N_harm)
    say(f"[WAVE MODEL] requested {p.wave_model}, N={p.N_harm}; used {wave.get_model_info()}")
    morison = MorisonCalculator(structure, wave, p.wave_dir, p.current_dir, p.Cd, p.Cm, p.rho_water)
    morison_results = morison.compute_all_morison_forces(p.t_analysis)
    say(f"[MORISON t={p.t_analysis:.2f}s] |F_total| = {np.linalg.norm(morison_results['total_morison']) / 1000:.1f} kN")
    critical = None
    phase_results = None
    if p.do_phase_scan:
        phase_results = morison.find_critical_phase(n_steps=p.n_phase_steps)
        critical = phase_results["critical"]
        say(f"[PHASE SCAN] critical t = {critical['t']:.3f}s, total = {critical['total_kN']:.1f} kN")

    fem = FEMSolver(structure, p.E, p.nu)
    top = structure.get_top_nodes()
    vec = interface_load_vector(p.wave_dir, len(top), p.F_axial, p.F_shear, p.M_moment, p.M_torsion)
    for name in top:
        fem.apply_nodal_force(name, vec)
    for name, force in morison_results["nodal_forces"].items():
        fv = np.zeros(6)
        fv[:3] = force[:3]
        fem.apply_nodal_force(name, fv)
    apply_self_weight(fem.F_global, structure, p.self_weight_mode, p.custom_sw)
    fem.apply_boundary_conditions(structure.get_bottom_nodes())
    fem._fy = float(p.fy)          # yield stress of the device-side utilisation (the reference passes fy to get_member_internal_forces only)
    U = fem.solve()
    reactions = fem.get_reactions()
    internal_forces = fem.get_member_internal_forces(p.fy)
    max_util = max(m["utilization"] for m in internal_forces)
    say(f"[FEM] max utilisation {max_util:.2%}")
    return {"U": U, "reactions": reactions, "internal_forces": internal_forces, "structure": structure,
            "max_util": max_util, "morison_results": morison_results, "critical_phase": critical,
            "wave_info": wave.get_model_info(), "phase_scan_results": phase_results, "fem": fem}


# ---------------------------------------------------------------------------------
# per-phase scan (the hot path)
# ---------------------------------------------------------------------------------
@dataclass
class PhaseScanResult:
    structure: CustomJacketStructure
    wave: RaschiiWave
    table: np.ndarray              # [P, 16], columns = _lib.TABLE_COLUMNS
    critical_index: int            # first index of max total_kN (GUI.py:717)
    fy: float
    engine: object = field(repr=False, default=None)
    columns: tuple = L.TABLE_COLUMNS
    generation: int = -1           # engine generation the resident rows belong to (a later scan on the same structure overwrites them)

    def _resident(self):
        if self.generation >= 0:
            self.engine.check_generation(self.generation, "PhaseScanResult")

    @property
    def n_phases(self):
        return self.table.shape[0]

    def row(self, i):
        return dict(zip(self.columns, (float(v) for v in self.table[i])))

    @property
    def critical(self):
        return self.row(self.critical_index)

    @property
    def governing_index(self):
        """Phase with the largest member utilisation (first maximum)."""
        return int(np.argmax(self.table[:, L.TABLE_COLUMNS.index("max_util")]))

    def all_phases(self):
        """The reference's find_critical_phase()['all_phases'] rows (8 Morison columns)."""
        return [dict(zip(self.columns[:8], (float(v) for v in r[:8]))) for r in self.table]

    def phase(self, i, end_forces=False):
        """Full results of phase i in the reference's shapes: U, reactions{node: [6]}, internal_forces[list of dict]."""
        self._resident()
        got = self.engine.fetch_phase(i, U=True, reactions=True, rows=True, end_forces=end_forces, nodal=True)
        st = self.structure
        reactions = {st.node_list[int(n)]: got["reactions"][k].copy() for k, n in enumerate(self.engine.fixed_idx)}
        rows = member_rows_to_dicts(st, got["rows"], self.fy, self.engine)
        out = {"U": got["U"], "reactions": reactions, "internal_forces": rows,
               "max_util": max(r["utilization"] for r in rows), "nodal_forces": got["nodal_forces"]}
        if end_forces:
            out["end_forces"] = got["end_forces"]
        return out

    def member_series(self, member, column="utilization"):
        self._resident()
        mi = member if isinstance(member, int) else [m["name"] for m in self.structure.members].index(member)
        return self.engine.member_column(mi, L.MEMBER_COLUMNS.index(column), self.n_phases)

    def to_dataframe(self):
        import pandas as pd
        return pd.DataFrame(self.table, columns=list(self.columns))


def phase_scan(structure, wave, n_steps=360, *, wave_direction=0.0, current_direction=0.0, Cd=0.7, Cm=2.0,
               rho_water=1025.0, E=210000.0, nu=0.3, fy=355.0, static_F=None, params: AnalysisParams | None = None,
               t=None, n_gauss=15, engine=None):
    """Morison + FEM for every phase t_i = i*T/n_steps (or the given ``t``)."""
    eng = engine or get_engine(structure)
    G = E / (2 * (1 + nu))
    eng.ensure_factored(structure.indices(structure.get_bottom_nodes()), E, G)
    if static_F is None:
        static_F = static_load(structure, params) if params is not None else np.zeros(structure.n_dof)
    eng.set_static_load(static_F)
    eng.set_wave(wave)
    eng.set_morison(np.deg2rad(90.0 - wave_direction), np.deg2rad(90.0 - current_direction), rho_water, Cd, Cm, n_gauss)
    tt = phase_times(wave.T, n_steps) if t is None else np.asarray(t, dtype=np.float64)
    table, crit = eng.phase_scan(tt, fy)          # column 1 (phase_deg) is filled on the device, bit-identical to GUI.py:697-698
    return PhaseScanResult(structure, wave, table, crit, fy, eng, generation=eng.generation)


def phase_scan_from_params(nodes, members, fixed_nodes, top_nodes, params: AnalysisParams | None = None, n_steps=360):
    """Convenience: GUI-style inputs -> per-phase table."""
    p = params or AnalysisParams()
    structure = build_structure(nodes, members, fixed_nodes, top_nodes, p)
    wave = RaschiiWave(p.H, p.T, p.d, p.U_c, p.wave_model, p.N_harm)
    return phase_scan(structure, wave, n_steps, wave_direction=p.wave_dir, current_direction=p.current_dir, Cd=p.Cd,
                      Cm=p.Cm, rho_water=p.rho_water, E=p.E, nu=p.nu, fy=p.fy, params=p)


# ---------------------------------------------------------------------------------
# sea-state ensemble (BASELINE configs[4]): many (H, T, direction) states on one factor
# ---------------------------------------------------------------------------------
@dataclass
class EnsembleResult:
    structure: CustomJacketStructure
    H: np.ndarray
    T: np.ndarray
    wave_dir: np.ndarray
    k: np.ndarray
    table: np.ndarray              # [S, n_phase, 16]
    critical_phase: np.ndarray     # [S] first index of max total_kN inside each state (GUI.py:717 per state)
    fy: float
    engine: object = field(repr=False, default=None)
    columns: tuple = L.TABLE_COLUMNS
    generation: int = -1

    @property
    def n_states(self):
        return self.table.shape[0]

    @property
    def n_phase(self):
        return self.table.shape[1]

    def critical_rows(self):
        """[S, 16]: the critical-phase row of every sea state."""
        return self.table[np.arange(self.n_states), self.critical_phase]

    @property
    def governing(self):
        """(state, phase) of the largest member utilisation over the whole ensemble (first maximum)."""
        u = self.table[:, :, L.TABLE_COLUMNS.index("max_util")]
        s, p = np.unravel_index(int(np.argmax(u)), u.shape)
        return int(s), int(p)

    def case(self, state, phase, end_forces=False):
        if self.generation >= 0:
            self.engine.check_generation(self.generation, "EnsembleResult")
        got = self.engine.fetch_phase(int(state) * self.n_phase + int(phase), U=True, reactions=True, rows=True,
                                      end_forces=end_forces, nodal=True)
        st = self.structure
        got["reactions"] = {st.node_list[int(n)]: got["reactions"][i].copy() for i, n in enumerate(self.engine.fixed_idx)}
        got["internal_forces"] = member_rows_to_dicts(st, got.pop("rows"), self.fy, self.engine)
        return got


def dispersion_wavenumbers(T, d, gravity=g):
    """The reference's Newton iteration (GUI.py:197-206) for many periods at once.  Every element follows the scalar
    algorithm, including its own stopping test (the loop breaks BEFORE the last update is applied).  NumPy's array
    tanh/cosh may differ from the scalar path the reference takes by one ulp, so k can differ from
    ``solve_dispersion`` in the last bit (2.7e-16 relative over 5000 random periods); single waves (RaschiiWave)
    keep the scalar path and are bit-identical to the reference."""
    omega = 2.0 * np.pi / np.asarray(T, dtype=np.float64).reshape(-1)
    w2 = omega**2
    k = w2 / gravity
    active = np.ones(k.shape, dtype=bool)
    for _ in range(50):
        if not active.any():
            break
        ka = k[active]
        th = np.tanh(ka * d)
        resid = w2[active] - gravity * ka * th
        slope = -gravity * (th + ka * d / np.cosh(ka * d)**2)
        k_next = ka - resid / slope
        done = np.abs(k_next - ka) < 1e-10
        idx = np.flatnonzero(active)
        k[idx[~done]] = k_next[~done]
        active[idx[done]] = False
    return k


def ensemble_scan(structure, H, T, wave_dir, n_phase=16, *, d=50.0, U_c=0.0, current_direction=0.0, Cd=0.7, Cm=2.0,
                  rho_water=1025.0, E=210000.0, nu=0.3, fy=355.0, params: AnalysisParams | None = None, n_gauss=15,
                  dt=0.001, engine=None, host_dispersion=False):
    """Morison + FEM for n_phase phases of every sea state (H[i], T[i], wave_dir[i]) -- Airy kinematics (the pinned
    model), current and depth common to all states, one Cholesky factor for the whole ensemble.  With ``params`` the
    GUI's interface loads and self-weight are applied; the interface shear follows each state's wave direction as
    run_analysis does (GUI.py:1967-1971)."""
    H, T, wave_dir = (np.asarray(v, dtype=np.float64).reshape(-1) for v in (H, T, wave_dir))
    S = H.shape[0]
    eng = engine or get_engine(structure)
    G = E / (2 * (1 + nu))
    eng.ensure_factored(structure.indices(structure.get_bottom_nodes()), E, G)
    F_dir = None
    if params is not None:
        p0 = AnalysisParams(**{**params.__dict__, "F_shear": 0.0})
        eng.set_static_load(static_load(structure, p0))
        F_dir = np.zeros((2, structure.n_dof))
        top = structure.get_top_nodes()
        for name in top:
            i = structure.node_index[name]
            F_dir[0, 6 * i] = params.F_shear * 1000.0 / len(top)
            F_dir[1, 6 * i + 1] = params.F_shear * 1000.0 / len(top)
    else:
        eng.set_static_load(np.zeros(structure.n_dof))
    wave0 = RaschiiWave(float(H[0]), float(T[0]), d, U_c, "Airy", 1, dt)          # carries depth / current / dt
    eng.set_wave(wave0)
    eng.set_morison(0.0, np.deg2rad(90.0 - current_direction), rho_water, Cd, Cm, n_gauss)
    if host_dispersion:
        # host set-up (vectorised NumPy Newton): the arrays jk_ensemble_scan takes
        k = dispersion_wavenumbers(T, d)
        omega = 2.0 * np.pi / T
        t = np.arange(n_phase)[None, :] * T[:, None] / n_phase                       # (i*T)/n_steps, GUI.py:696, per state
        table, crit = eng.ensemble_scan(H / 2.0, k, omega, np.deg2rad(90.0 - wave_dir), t, fy, F_dir)
    else:
        # default: raw sea-state numbers in, dispersion / headings / case times on the device
        table, crit, k = eng.ensemble_scan_sea_states(H, T, wave_dir, n_phase, fy, F_dir, gravity=g)
    return EnsembleResult(structure, H, T, wave_dir, k, table, crit, fy, eng, generation=eng.generation)
