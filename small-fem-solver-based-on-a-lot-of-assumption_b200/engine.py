"""Engine: one GPU handle (include/jacket_b200.h) per structure.

Owns the call order of the C ABI (create -> supports -> assemble -> factor ->
loads / wave / Morison -> scan) and caches what has already been sent so the
reference-style facade classes can be used in any order the GUI uses them.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _lib as L


def default_device():
    for key in ("JK_DEVICE", "LOCAL_RANK"):
        if key in os.environ:
            return int(os.environ[key])
    return 0


class Engine:
    def __init__(self, structure, device=None, stream=None, ordering="rcm", solver="banded", options=None):
        self.structure = structure
        self.lib = L.lib()
        self.device = default_device() if device is None else int(device)
        self.ordering = {"rcm": L.ORDER_RCM, "natural": L.ORDER_NATURAL}[ordering]
        self.solver = {"banded": L.SOLVER_BANDED, "dense": L.SOLVER_DENSE}[solver]
        xyz, conn, sec_id, props, sections = structure.pack()
        self.xyz, self.conn, self.sec_id, self.sections = xyz, conn, sec_id, sections
        self.n_nodes, self.n_members = xyz.shape[0], conn.shape[0]
        self.h = C.c_void_p()
        rc = self.lib.jk_create(self.device, C.c_void_p(stream or 0), self.n_nodes, L.dptr(xyz), self.n_members,
                                L.iptr(conn), L.iptr(sec_id), props.shape[0], L.dptr(props), C.byref(self.h))
        if rc != 0:
            msg = self.lib.jk_last_error(self.h if self.h.value else None)
            if self.h.value:
                self.lib.jk_destroy(self.h)
                self.h = C.c_void_p()
            raise L.JacketError(rc, msg.decode() if msg else "")
        self.options = dict(options or {})
        for key, value in self.options.items():          # jk_set_option: documented switches of include/jacket_b200.h
            self._ck(self.lib.jk_set_option(self.h, str(key).encode(), int(value)))
        self._fixed = None
        self._moduli = None
        self._factored = False
        self._static = None
        self._wave_sig = None
        self._morison_sig = None
        self.fixed_idx = None
        self._lengths = None
        # Results live in HBM and are overwritten by the next scan / solve: every call that replaces or invalidates them
        # bumps the generation, result objects remember theirs and fetches through a stale object raise (check_generation)
        self.generation = 0

    def close(self):
        self.__dict__.pop("_dev_views", None)            # torch views of library-owned device buffers (distributed.device_views)
        self.__dict__.pop("_pinned_pair", None)
        if getattr(self, "h", None) is not None and self.h.value:
            self.lib.jk_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc == -4:                                  # JK_ENOTSPD: the factor is unusable
            self._factored = False
        L.check(rc, self.h)

    def _bump(self):
        self.generation += 1
        return self.generation

    def check_generation(self, generation, who="result"):
        if generation != self.generation:
            raise L.JacketError(-5, f"{who} is stale: the engine's device buffers were overwritten by a later scan / solve / "
                                    "re-assembly on the same structure (results are resident in HBM, one set per engine)")

    # -- options (documented run-time switches of the library, jk_set_option) ----
    def set_option(self, key, value):
        self._ck(self.lib.jk_set_option(self.h, str(key).encode(), int(value)))

    def get_option(self, key):
        v = C.c_int(0)
        self._ck(self.lib.jk_get_option(self.h, str(key).encode(), C.byref(v)))
        return int(v.value)

    # -- FEM side ---------------------------------------------------------------
    def set_supports(self, fixed_idx):
        fixed_idx = L.i32(fixed_idx)
        _, first = np.unique(fixed_idx, return_index=True)         # the library drops repeated nodes, keeping first occurrences
        fixed_idx = L.i32(fixed_idx[np.sort(first)])
        key = tuple(int(i) for i in fixed_idx)
        if key == self._fixed:
            return
        self._bump()
        self._ck(self.lib.jk_set_supports(self.h, len(fixed_idx), L.iptr(fixed_idx), self.ordering, self.solver))
        self._fixed, self.fixed_idx = key, fixed_idx
        self._assembled_for = None
        self._factored = False

    def assemble(self, E, G):
        self._bump()
        self._ck(self.lib.jk_assemble(self.h, float(E), float(G)))
        self._moduli = (float(E), float(G))
        self._factored = False
        self._lengths = None

    def factor(self, overlap=False):
        """Blocked Cholesky.  overlap=True queues it on the side stream so the next scan's Morison stage runs
        concurrently (a non-SPD matrix is then reported by that scan)."""
        self._ck(self.lib.jk_factor_begin(self.h) if overlap else self.lib.jk_factor(self.h))
        self._factored = True

    def ensure_factored(self, fixed_idx, E, G):
        self.set_supports(fixed_idx)
        if not self._factored or self._moduli != (float(E), float(G)):
            self.assemble(E, G)
            self.factor()

    def set_static_load(self, F):
        F = L.f64(F).reshape(-1)
        assert F.shape[0] == 6 * self.n_nodes
        self._ck(self.lib.jk_set_static_load(self.h, L.dptr(F)))

    # -- Morison side -----------------------------------------------------------
    def set_wave(self, wave):
        sig = wave.signature()
        if sig != self._wave_sig:
            if getattr(wave, "kind", "airy") == "fourier":
                k, om, d, Uc, dt, E, B = wave.fourier_args()
                self._ck(self.lib.jk_set_wave_fourier(self.h, k, om, d, Uc, dt, len(E), L.dptr(E), L.dptr(B)))
            else:
                self._ck(self.lib.jk_set_wave_airy(self.h, *wave.device_args()))
            self._wave_sig = sig

    def set_morison(self, theta_wave, theta_current, rho, Cd, Cm, n_gauss=15):
        sig = (float(theta_wave), float(theta_current), float(rho), float(Cd), float(Cm), int(n_gauss))
        if sig != self._morison_sig:
            xi, wts = np.polynomial.legendre.leggauss(n_gauss)   # same rule as GUI.py:615-617
            s = L.f64((xi + 1.0) / 2.0)
            w = L.f64(wts / 2.0)
            self._ck(self.lib.jk_set_morison(self.h, *sig[:5], n_gauss, L.dptr(s), L.dptr(w)))
            self._morison_sig = sig

    # -- scans ------------------------------------------------------------------
    def morison_scan(self, t):
        self._bump()
        t = L.f64(t).reshape(-1)
        table = np.zeros((t.shape[0], L.TABLE_NCOL))
        crit = C.c_int64(-1)
        self._ck(self.lib.jk_morison_scan(self.h, t.shape[0], L.dptr(t), L.dptr(table), C.byref(crit)))
        return table, int(crit.value)

    def morison_single(self, t, want_details=True):
        self._bump()
        nodal = np.zeros((self.n_nodes, 3))
        totals = np.zeros(9)
        details = np.zeros((self.n_members, L.DETAIL_NCOL)) if want_details else None
        self._ck(self.lib.jk_morison_single(self.h, float(t), L.dptr(nodal), L.dptr(totals), L.dptr(details)))
        return nodal, totals, details

    def phase_scan(self, t, fy):
        self._bump()
        t = L.f64(t).reshape(-1)
        table = np.empty((t.shape[0], L.TABLE_NCOL))     # the library overwrites every entry
        crit = C.c_int64(-1)
        self._ck(self.lib.jk_phase_scan(self.h, t.shape[0], L.dptr(t), float(fy), L.dptr(table), C.byref(crit)))
        return table, int(crit.value)

    def phase_scan_begin(self, t, fy):
        self._bump()
        """Queue a scan with host times and return without reading anything back (results: read_table / device views)."""
        t = L.f64(t).reshape(-1)
        self._ck(self.lib.jk_phase_scan(self.h, t.shape[0], L.dptr(t), float(fy), None, None))
        return t.shape[0]

    def read_critical(self, P):
        """Synchronise with the last scan and return its local critical index (also surfaces a failed factorisation)."""
        crit = C.c_int64(-1)
        self._ck(self.lib.jk_read_table(self.h, P, None, C.byref(crit)))
        return int(crit.value)

    def phase_scan_dev(self, P, t_dev_ptr, fy):
        self._bump()
        self._ck(self.lib.jk_phase_scan_dev(self.h, int(P), C.c_void_p(t_dev_ptr), float(fy)))

    # -- whole steps (assemble + asynchronous factor + scan as one call; replayed as a CUDA graph) ------------------
    def _after_step(self, E, G):
        self._moduli = (float(E), float(G))
        self._factored = True

    def step_dev(self, E, G, P, t_dev_ptr, fy):
        """jk_step_dev: resident step, nothing crosses PCIe, no host synchronisation."""
        self._bump()
        self._ck(self.lib.jk_step_dev(self.h, float(E), float(G), int(P), C.c_void_p(t_dev_ptr), float(fy)))
        self._after_step(E, G)

    def step(self, E, G, t, fy, read=True):
        """jk_step: host times in; with read=True the host table and the critical index come back (one synchronisation)."""
        self._bump()
        t = L.f64(t).reshape(-1)
        if not read:
            self._ck(self.lib.jk_step(self.h, float(E), float(G), t.shape[0], L.dptr(t), float(fy), None, None))
            self._after_step(E, G)
            return t.shape[0]
        table = np.empty((t.shape[0], L.TABLE_NCOL))
        crit = C.c_int64(-1)
        self._ck(self.lib.jk_step(self.h, float(E), float(G), t.shape[0], L.dptr(t), float(fy), L.dptr(table), C.byref(crit)))
        self._after_step(E, G)
        return table, int(crit.value)

    def read_table(self, P):
        table = np.zeros((P, L.TABLE_NCOL))
        crit = C.c_int64(-1)
        self._ck(self.lib.jk_read_table(self.h, P, L.dptr(table), C.byref(crit)))
        return table, int(crit.value)

    def ensemble_scan(self, a, k, omega, theta_wave, t, fy, F_dir=None):
        self._bump()
        a, k, omega, theta_wave = (L.f64(v).reshape(-1) for v in (a, k, omega, theta_wave))
        S = a.shape[0]
        t = L.f64(t).reshape(S, -1)
        n_phase = t.shape[1]
        table = np.zeros((S * n_phase, L.TABLE_NCOL))
        crit = np.zeros(S, dtype=np.int64)
        Fd = None if F_dir is None else L.f64(F_dir).reshape(2, 6 * self.n_nodes)
        self._ck(self.lib.jk_ensemble_scan(self.h, S, n_phase, L.dptr(a), L.dptr(k), L.dptr(omega), L.dptr(theta_wave), L.dptr(t),
                                           L.dptr(Fd), float(fy), L.dptr(table), crit.ctypes.data_as(C.POINTER(C.c_int64))))
        return table.reshape(S, n_phase, L.TABLE_NCOL), crit

    def ensemble_scan_sea_states(self, H, T, wave_dir_deg, n_phase, fy, F_dir=None, gravity=9.81):
        """Ensemble from the raw sea-state numbers: dispersion, headings and case times are computed on the device
        (jk_ensemble_scan_sea_states).  Returns (table[S][n_phase][16], critical[S], k[S])."""
        self._bump()
        H, T, wave_dir_deg = (L.f64(v).reshape(-1) for v in (H, T, wave_dir_deg))
        S = H.shape[0]
        if T.shape[0] != S or wave_dir_deg.shape[0] != S:
            raise ValueError("H, T and wave_dir must have one entry per sea state")
        table = np.zeros((S * int(n_phase), L.TABLE_NCOL))
        crit = np.zeros(S, dtype=np.int64)
        k = np.zeros(S)
        Fd = None if F_dir is None else L.f64(F_dir).reshape(2, 6 * self.n_nodes)
        self._ck(self.lib.jk_ensemble_scan_sea_states(self.h, S, int(n_phase), L.dptr(H), L.dptr(T), L.dptr(wave_dir_deg), float(gravity),
                                                      L.dptr(Fd), float(fy), L.dptr(table), crit.ctypes.data_as(C.POINTER(C.c_int64)), L.dptr(k)))
        return table.reshape(S, int(n_phase), L.TABLE_NCOL), crit, k

    def solve(self, F, fy=355.0):
        self._bump()
        F = L.f64(F).reshape(-1, 6 * self.n_nodes)
        self._ck(self.lib.jk_solve(self.h, F.shape[0], L.dptr(F), float(fy)))

    def fetch_phase(self, p, U=True, reactions=True, rows=True, end_forces=False, nodal=False):
        n_fixed = len(self._fixed) if self._fixed else 0
        out = {}
        aU = np.zeros(6 * self.n_nodes) if U else None
        aR = np.zeros((n_fixed, 6)) if reactions else None
        aM = np.zeros((self.n_members, L.MEMBER_NCOL)) if rows else None
        aE = np.zeros((self.n_members, 12)) if end_forces else None
        aN = np.zeros((self.n_nodes, 3)) if nodal else None
        self._ck(self.lib.jk_fetch_phase(self.h, int(p), L.dptr(aU), L.dptr(aR), L.dptr(aM), L.dptr(aE), L.dptr(aN)))
        out.update(U=aU, reactions=aR, rows=aM, end_forces=aE, nodal_forces=aN)
        return out

    def member_column(self, member, column, P):
        out = np.zeros(P)
        self._ck(self.lib.jk_fetch_member_column(self.h, int(member), int(column), int(P), L.dptr(out)))
        return out

    # -- introspection ----------------------------------------------------------
    def dims(self):
        d = np.zeros(12, dtype=np.int32)
        self._ck(self.lib.jk_get_dims(self.h, L.iptr(d)))
        return dict(zip(("n_nodes", "n_members", "n_fixed", "n_free_dof", "n_pad", "tile", "band_tiles", "n_tiles",
                         "dof_half_bandwidth", "n_chains", "separator_tile_row", "separator_nodes"), (int(v) for v in d)))

    def order(self):
        d = self.dims()
        o = np.zeros(d["n_nodes"] - d["n_fixed"], dtype=np.int32)
        self._ck(self.lib.jk_get_order(self.h, L.iptr(o)))
        return o

    def dense_K(self):
        n = 6 * self.n_nodes
        K = np.zeros((n, n))
        self._ck(self.lib.jk_get_K(self.h, L.dptr(K)))
        return K

    def elements(self):
        M = self.n_members
        Ke, Kl, R, Ln = np.zeros((M, 12, 12)), np.zeros((M, 12, 12)), np.zeros((M, 3, 3)), np.zeros(M)
        self._ck(self.lib.jk_get_elements(self.h, L.dptr(Ke), L.dptr(Kl), L.dptr(R), L.dptr(Ln)))
        return Ke, Kl, R, Ln

    def member_lengths(self):
        """Member lengths in m (cached: they only depend on the geometry)."""
        if self._lengths is None:
            Ln = np.zeros(self.n_members)
            self._ck(self.lib.jk_get_elements(self.h, None, None, None, L.dptr(Ln)))
            self._lengths = Ln
        return self._lengths

    def kinematics_points(self, xyz, t):
        """MorisonCalculator.get_kinematics_3d (GUI.py:559-589) for many points at once, on the device: [n, 10] =
        u_wave v_wave w_wave u_current v_current du_dt dv_dt dw_dt submerged eta."""
        xyz = L.f64(xyz).reshape(-1, 3)
        out = np.zeros((xyz.shape[0], 10))
        self._ck(self.lib.jk_kinematics_points(self.h, xyz.shape[0], L.dptr(xyz), float(t), L.dptr(out)))
        return out

    def timings(self):
        ms = np.zeros(L.NTIMERS)
        self._ck(self.lib.jk_get_timings(self.h, L.dptr(ms)))
        return dict(zip(L.TIMER_NAMES, (float(v) for v in ms)))

    def residual(self):
        r = C.c_double(0.0)
        self._ck(self.lib.jk_residual(self.h, C.byref(r)))
        return float(r.value)

    def solver_stats(self):
        """nnz(L), executed sweep flops per load case, sweep items per slab, TMA-pipeline flag."""
        out = np.zeros(8)
        self._ck(self.lib.jk_solver_stats(self.h, L.dptr(out)))
        return {"nnz_L": int(out[0]), "sweep_flops_executed_per_case": float(out[1]), "sweep_items": int(out[2]),
                "tma_sweep": bool(out[3]), "nnz_L_min_envelope": int(out[4]), "sweep_slab": int(out[5]),
                "step_graph": {1: "replayed", 0: "none", -1: "capture unsupported (eager)"}[int(out[6])], "graph_kernels": int(out[7])}

    def launch_count(self):
        return int(self.lib.jk_launch_count(self.h))

    def stream(self):
        return self.lib.jk_stream(self.h)


def get_engine(structure, **kw):
    """The structure's engine, created on first use (one handle per structure and GPU).  A different device, ordering or
    solver storage than the existing engine's replaces it (earlier result objects then report themselves stale)."""
    eng = getattr(structure, "_engine", None)
    if eng is not None and kw:
        want = dict(device=kw.get("device", eng.device),
                    ordering={"rcm": L.ORDER_RCM, "natural": L.ORDER_NATURAL}[kw["ordering"]] if "ordering" in kw else eng.ordering,
                    solver={"banded": L.SOLVER_BANDED, "dense": L.SOLVER_DENSE}[kw["solver"]] if "solver" in kw else eng.solver,
                    options=dict(kw["options"] or {}) if "options" in kw else eng.options)
        if any(getattr(eng, k) != v for k, v in want.items()):
            eng._bump()
            eng.close()
            eng = None
    if eng is None:
        eng = Engine(structure, **kw)
        structure._engine = eng
    return eng
