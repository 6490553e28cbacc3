"""FEMSolver / BeamElement3D -- drop-ins for GUI.py:360-533, evaluated on the GPU.

Element stiffness, assembly, the Cholesky factorisation, the triangular sweeps, reactions and member forces of a
structure all run in the CUDA library; this module keeps the reference's attributes (writable ``F_global``,
``U_global``, ``fixed_dofs`` ...), constructor signatures and return types.  ``BeamElement3D`` can also be built on
its own from two coordinates like the reference's (GUI.py:361): that is a 12 x 12 host computation, checked against
the device-built element matrices by tests/test_gpu_parity.py (``k_member_setup`` produces the same K_local, R and
K_global for every member of a structure; those are what ``FEMSolver.elements`` returns).
"""
from __future__ import annotations

import numpy as np

from . import _lib as L
from .engine import get_engine

_FORCE_KEYS = ("Fx", "Fy", "Fz", "Mx", "My", "Mz")
_UP = np.array([0.0, 0.0, 1.0])


def _put_sym(K, value, pairs):
    for i, j in pairs:
        K[i, j] = K[j, i] = value


class BeamElement3D:
    def __init__(self, node1_coords, node2_coords, section, E=210000, G=80769, include_shear=True):
        self.node1, self.node2 = node1_coords, node2_coords
        self.section, self.E, self.G = section, E, G
        self.dL = node2_coords - node1_coords
        self.L = np.linalg.norm(self.dL)
        self.L_mm = self.L * 1000.0
        self.T = self._compute_transformation_matrix()
        self.K_local = self._compute_local_stiffness(include_shear)
        self.K_global = self.T.T @ self.K_local @ self.T

    @classmethod
    def _from_device(cls, node1_coords, node2_coords, section, E, G, L_m, R, K_local, K_global):
        """View of one element as built on the device by k_member_setup (FEMSolver.elements)."""
        self = cls.__new__(cls)
        self.node1, self.node2 = node1_coords, node2_coords
        self.section, self.E, self.G = section, E, G
        self.dL = node2_coords - node1_coords
        self.L = L_m
        self.L_mm = L_m * 1000.0
        self.T = np.kron(np.eye(4), R)
        self.K_local, self.K_global = K_local, K_global
        return self

    def _compute_transformation_matrix(self):
        """Direction cosines (GUI.py:371-387): local x along the member; a (nearly) vertical member takes its local y
        from z^ x lx, any other member its local z from lx x z^; T = blockdiag(R, R, R, R)."""
        ex = self.dL / self.L
        if abs(ex @ _UP) > 0.999:
            ey = np.cross(_UP, ex)
            n = np.linalg.norm(ey)
            ey = ey / n if n > 1e-10 else np.array([0.0, 1.0, 0.0])
            ez = np.cross(ex, ey)
        else:
            ez = np.cross(ex, _UP)
            ez = ez / np.linalg.norm(ez)
            ey = np.cross(ez, ex)
        return np.kron(np.eye(4), np.vstack([ex, ey, ez]))

    def _compute_local_stiffness(self, include_shear):
        """Timoshenko beam in local axes, mm units (GUI.py:389-422): axial and torsional 2 x 2 blocks plus the two
        bending planes with shear parameters Phi = 12 E I / (G A_s L^2)."""
        Lm, sec, E, G = self.L_mm, self.section, self.E, self.G
        with_shear = include_shear and sec.Ay_mm2 > 0 and sec.Az_mm2 > 0
        phi_y = 12.0 * E * sec.Iz_mm4 / (G * sec.Az_mm2 * Lm**2) if with_shear else 0.0
        phi_z = 12.0 * E * sec.Iy_mm4 / (G * sec.Ay_mm2 * Lm**2) if with_shear else 0.0
        K = np.zeros((12, 12))
        for stiff, a, b in ((E * sec.Ax_mm2 / Lm, 0, 6), (G * sec.Ix_mm4 / Lm, 3, 9)):       # bar and shaft
            K[a, a] = K[b, b] = stiff
            K[a, b] = K[b, a] = -stiff
        # bending in the local x-y plane (deflection v = DOF 1 / 7, rotation about z = DOF 5 / 11) and in the x-z plane
        # (w = 2 / 8, rotation about y = 4 / 10; the slope there is -dw/dx, hence the opposite sign of the coupling)
        for EI, phi, v1, r1, v2, r2, sgn in ((E * sec.Iz_mm4, phi_y, 1, 5, 7, 11, 1.0), (E * sec.Iy_mm4, phi_z, 2, 4, 8, 10, -1.0)):
            b = EI / ((1.0 + phi) * Lm**3)
            K[v1, v1] = K[v2, v2] = 12.0 * b
            _put_sym(K, -12.0 * b, [(v1, v2)])
            _put_sym(K, sgn * (6.0 * b * Lm), [(v1, r1), (v1, r2)])
            _put_sym(K, -sgn * (6.0 * b * Lm), [(v2, r1), (v2, r2)])
            K[r1, r1] = K[r2, r2] = (4.0 + phi) * b * Lm**2
            _put_sym(K, (2.0 - phi) * b * Lm**2, [(r1, r2)])
        return K

    def get_internal_forces(self, u_global):
        """End forces in local axes; node 1 reports the negated stiffness forces (GUI.py:424-432)."""
        F = self.K_local @ (self.T @ u_global)
        return {"node1": {k: -F[i] for i, k in enumerate(_FORCE_KEYS)},
                "node2": {k: F[6 + i] for i, k in enumerate(_FORCE_KEYS)}}


class FEMSolver:
    def __init__(self, structure, E=210000, nu=0.3):
        self.structure = structure
        self.n_dof = structure.n_dof
        self.E = E
        self.G = E / (2 * (1 + nu))
        self.F_global = np.zeros(self.n_dof)
        self.U_global = np.zeros(self.n_dof)
        self._eng = get_engine(structure)        # ordering / storage: get_engine(structure, ordering=..., solver=...) beforehand
        self._fy = 355.0
        self._solved_gen = None
        self._elements = None
        self.fixed_dofs = np.array([], dtype=int)
        self.free_dofs = np.arange(self.n_dof)
        # the reference builds the elements and assembles in __init__; supports default to the structure's bottom nodes
        self._fixed_names = list(structure.get_bottom_nodes())
        self._build_elements()
        self._assemble_global_stiffness()

    # ------------------------------------------------------------------------
    def _build_elements(self):
        """GUI.py:451-455: the element matrices are built on the device together with the assembly; ``elements``
        fetches them on first use."""
        self._elements = None

    def _assemble_global_stiffness(self):
        """GUI.py:457-467 on the device: deterministic assembly of K_ff into the tile storage of the solver."""
        self._prepare(self._fixed_names, factor=False)

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
        """Dense K in the reference's DOF order (rebuilt from the element matrices on the device on demand)."""
        return self._eng.dense_K()

    @property
    def elements(self):
        if self._elements is None:
            Ke, Kl, R, Ln = self._eng.elements()
            st = self.structure
            self._elements = [BeamElement3D._from_device(st.nodes[m["node1"]], st.nodes[m["node2"]], m["section"], self.E, self.G,
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

    def solve(self):
        if not self._eng._factored or self._eng._moduli != (float(self.E), float(self.G)):
            self._prepare(self._fixed_names, factor=True)
        self._eng.solve(self.F_global, self._fy)
        self._solved_gen = self._eng.generation
        self.U_global = self._eng.fetch_phase(0, U=True, reactions=False, rows=False)["U"]
        return self.U_global

    def _require_solution(self):
        if self._solved_gen is None:
            raise RuntimeError("FEMSolver: call solve() first")
        self._eng.check_generation(self._solved_gen, "FEMSolver")

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
    lengths = engine.member_lengths()
    out = []
    for i, m in enumerate(structure.members):
        d = {"member": m["name"], "type": m["type"], "node1": m["node1"], "node2": m["node2"],
             "length_m": float(lengths[i])}
        for j, c in enumerate(L.MEMBER_COLUMNS[:6]):
            d[c] = float(rows[i, j])
        d["utilization"] = d["von_mises_max_MPa"] / fy
        out.append(d)
    return out
