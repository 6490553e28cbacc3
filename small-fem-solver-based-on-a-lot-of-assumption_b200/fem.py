"""FEMSolver / BeamElement3D -- drop-ins for GUI.py:360-533, evaluated on the GPU.

Element stiffness, assembly, the Cholesky factorisation, the triangular sweeps,
reactions and member forces all run in the CUDA library; this module keeps the
reference's attributes (writable ``F_global``, ``U_global``, ``fixed_dofs`` ...)
and return types.
"""
from __future__ import annotations

import numpy as np

from . import _lib as L
from .engine import get_engine


class BeamElement3D:
    """Read-only view of one element as built on the device (K_local, T, K_global)."""

    def __init__(self, node1_coords, node2_coords, section, E, G, L_m, R, K_local, K_global):
        self.node1, self.node2 = node1_coords, node2_coords
        self.section, self.E, self.G = section, E, G
        self.dL = node2_coords - node1_coords
        self.L = L_m
        self.L_mm = L_m * 1000.0
        self.T = np.kron(np.eye(4), R)
        self.K_local = K_local
        self.K_global = K_global

    def get_internal_forces(self, u_global):
        F = self.K_local @ (self.T @ u_global)
        keys = ("Fx", "Fy", "Fz", "Mx", "My", "Mz")
        return {"node1": {k: -F[i] for i, k in enumerate(keys)},
                "node2": {k: F[6 + i] for i, k in enumerate(keys)}}


class FEMSolver:
    def __init__(self, structure, E=210000, nu=0.3, ordering=None, solver=None):
        self.structure = structure
        self.n_dof = structure.n_dof
        self.E = E
        self.G = E / (2 * (1 + nu))
        self.F_global = np.zeros(self.n_dof)
        self.U_global = np.zeros(self.n_dof)
        kw = {}
        if ordering is not None or solver is not None:
            structure._engine = None
            kw = dict(ordering=ordering or "rcm", solver=solver or "banded")
        self._eng = get_engine(structure, **kw)
        self._fy = 355.0
        self._solved = False
        self._elements = None
        self.fixed_dofs = np.array([], dtype=int)
        self.free_dofs = np.arange(self.n_dof)
        # the reference assembles in __init__; supports default to the structure's bottom nodes
        self._prepare(structure.get_bottom_nodes(), factor=False)

    # ------------------------------------------------------------------------
    def _prepare(self, fixed_nodes, factor):
        idx = self.structure.indices(fixed_nodes)
        self._eng.set_supports(idx)
        if self._eng._moduli != (float(self.E), float(self.G)) or not self._eng._factored:
            self._eng.assemble(self.E, self.G)
            if factor:
                self._eng.factor()
        self._fixed_names = list(fixed_nodes)

    @property
    def K_global(self):
        """Dense K in the reference's DOF order (fetched from the device on demand)."""
        if self._eng._factored:          # tile storage holds L now; rebuild element matrices only
            pass
        return self._eng.dense_K()

    @property
    def elements(self):
        if self._elements is None:
            Ke, Kl, R, Ln = self._eng.elements()
            st = self.structure
            self._elements = [BeamElement3D(st.nodes[m["node1"]], st.nodes[m["node2"]], m["section"], self.E, self.G,
                                            Ln[i], R[i], Kl[i], Ke[i]) for i, m in enumerate(st.members)]
        return self._elements

    def apply_nodal_force(self, node_name, force_vector):
        i = self.structure.node_index[node_name]
        self.F_global[6 * i:6 * i + 6] += force_vector

    def apply_boundary_conditions(self, fixed_nodes):
        dofs = []
        for name in fixed_nodes:
            i = self.structure.node_index[name]
            dofs.extend(range(6 * i, 6 * i + 6))
        self.fixed_dofs = np.array(dofs, dtype=int)
        self.free_dofs = np.setdiff1d(np.arange(self.n_dof), self.fixed_dofs)
        self._prepare(fixed_nodes, factor=True)

    def solve(self, fy=None):
        if fy is not None:
            self._fy = float(fy)
        if not self._eng._factored:
            self._prepare(self._fixed_names, factor=True)
        self._eng.solve(self.F_global, self._fy)
        self.U_global = self._eng.fetch_phase(0, U=True, reactions=False, rows=False)["U"]
        self._solved = True
        return self.U_global

    def _require_solution(self):
        if not self._solved:
            raise RuntimeError("FEMSolver: call solve() first")

    def get_reactions(self):
        self._require_solution()
        R = self._eng.fetch_phase(0, U=False, reactions=True, rows=False)["reactions"]
        st = self.structure
        return {st.node_list[int(n)]: R[i].copy() for i, n in enumerate(self._eng.fixed_idx)}

    def get_member_internal_forces(self, fy=355):
        self._require_solution()
        rows = self._eng.fetch_phase(0, U=False, reactions=False, rows=True)["rows"]
        return member_rows_to_dicts(self.structure, rows, fy, self._eng)


def member_rows_to_dicts(structure, rows, fy, engine):
    """Numeric member rows -> the reference's list of dicts (schema GUI.py:521-532)."""
    lengths = engine.elements()[3]
    out = []
    for i, m in enumerate(structure.members):
        d = {"member": m["name"], "type": m["type"], "node1": m["node1"], "node2": m["node2"],
             "length_m": float(lengths[i])}
        for j, c in enumerate(L.MEMBER_COLUMNS[:6]):
            d[c] = float(rows[i, j])
        d["utilization"] = d["von_mises_max_MPa"] / fy
        out.append(d)
    return out
