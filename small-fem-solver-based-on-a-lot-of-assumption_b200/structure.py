"""Geometry carriers: CustomJacketStructure (GUI.py:302-354), the default 3-leg
jacket (GUI.py:730-803) and a synthetic L-leg x B-bay generator for the
benchmark configurations."""
from __future__ import annotations

import string

import numpy as np

from .sections import TubularSection


class CustomJacketStructure:
    """Name-keyed nodes / members with the reference's attribute surface."""

    def __init__(self, nodes_dict, members_list, section_leg, section_brace,
                 fixed_nodes, top_nodes, rho_steel=7850):
        self.nodes = nodes_dict
        self.node_list = list(nodes_dict)
        self.n_nodes = len(self.node_list)
        self.n_dof = 6 * self.n_nodes
        self.node_index = {n: i for i, n in enumerate(self.node_list)}
        self.section_leg, self.section_brace = section_leg, section_brace
        self.rho_steel = rho_steel
        self.members = [dict(name=m["name"], node1=m["node1"], node2=m["node2"],
                             section=section_leg if m.get("type", "brace") == "leg" else section_brace,
                             type=m.get("type", "brace"))
                        for m in members_list]
        self.n_members = len(self.members)
        self._fixed_nodes = fixed_nodes
        self._top_nodes = top_nodes
        self._engine = None          # lazily created GPU handle (engine.Engine)

    def get_member_geometry(self, member):
        c1, c2 = self.nodes[member["node1"]], self.nodes[member["node2"]]
        dL = c2 - c1
        L = np.linalg.norm(dL)
        return {"coord1": c1, "coord2": c2, "dL": dL, "L": L, "L_mm": L * 1000.0,
                "unit_vec": dL / L if L > 0 else np.array([1, 0, 0])}

    def get_top_nodes(self):
        return self._top_nodes

    def get_bottom_nodes(self):
        return self._fixed_nodes

    # -- structure-of-arrays view for the C ABI -----------------------------------
    def pack(self):
        """(xyz[Nn,3] f64, conn[M,2] i32, sec_id[M] i32, sec_props[S,8] f64, sections)"""
        xyz = np.array([np.asarray(self.nodes[n], dtype=np.float64) for n in self.node_list]).reshape(-1, 3)
        conn = np.array([[self.node_index[m["node1"]], self.node_index[m["node2"]]] for m in self.members],
                        dtype=np.int32).reshape(-1, 2)
        sections, ids = [], {}
        sec_id = np.empty(self.n_members, dtype=np.int32)
        for i, m in enumerate(self.members):
            s = m["section"]
            if id(s) not in ids:
                ids[id(s)] = len(sections)
                sections.append(s)
            sec_id[i] = ids[id(s)]
        props = np.array([s.prop_row() for s in sections], dtype=np.float64)
        return xyz, conn, sec_id, props, sections

    def indices(self, names):
        return np.array([self.node_index[n] for n in names], dtype=np.int32)


# ---------------------------------------------------------------------------------
# default geometry: plan positions of the three legs at the four levels, hinge nodes
# of the X-braced panels, all shifted by the water reference (GUI.py:730-765)
# ---------------------------------------------------------------------------------
_LEVEL_Z = (0.0, 28.41, 52.89, 74.0)
_LEG_XY = {
    "A": ((-9.2376, -16.0), (-7.9254, -13.7272), (-6.7947, -11.7688), (-5.8197, -10.08)),
    "B": ((18.4752, 0.0), (15.8508, 0.0), (13.5894, 0.0), (11.6394, 0.0)),
    "C": ((-9.2376, 16.0), (-7.9254, 13.7272), (-6.7947, 11.7688), (-5.8197, 10.08)),
}
_HINGE_Z = (15.291, 41.5902, 64.2608)
_HINGE_XY = {
    "AB": ((4.2657, -7.3884), (3.6583, -6.3364), (3.1348, -5.4296)),
    "BC": ((4.2657, 7.3884), (3.6583, 6.3364), (3.1348, 5.4296)),
    "CA": ((-8.5313, 0.0), (-7.3166, 0.0), (-6.2695, 0.0)),
}


def create_default_3leg_jacket(z_water_ref=47.0):
    """nodes, members, fixed_nodes, top_nodes of the reference's default model
    (21 nodes / 51 members), same names and ordering as GUI.py:730-803."""
    nodes = {}
    for leg, xy in _LEG_XY.items():
        for lvl, (x, y) in enumerate(xy):
            nodes[f"{leg}{lvl + 1}"] = np.array([x, y, _LEVEL_Z[lvl] - z_water_ref])
    for lvl in range(3):
        for face, xy in _HINGE_XY.items():
            nodes[f"H{face}{lvl + 1}"] = np.array([xy[lvl][0], xy[lvl][1], _HINGE_Z[lvl] - z_water_ref])

    members = []

    def add(prefix, a, b, kind):
        members.append({"name": f"{prefix}_{a}-{b}", "node1": a, "node2": b, "type": kind})

    for leg in "ABC":
        for i in (1, 2, 3):
            add("Leg", f"{leg}{i}", f"{leg}{i + 1}", "leg")
    faces = (("A", "B"), ("B", "C"), ("C", "A"))
    for lvl in (1, 2):                                   # ring braces exist at levels 1 and 2 only
        for p, q in faces:
            add("HBrace", f"{p}{lvl}", f"{q}{lvl}", "h_brace")
    for lvl in (1, 2, 3):                                # X panels through the hinge node
        for p, q in faces:
            hinge = f"H{p}{q}{lvl}"
            add("XBr", f"{p}{lvl}", hinge, "x_brace")
            add("XBr", hinge, f"{q}{lvl + 1}", "x_brace")
            add("XBr", f"{q}{lvl}", hinge, "x_brace")
            add("XBr", hinge, f"{p}{lvl + 1}", "x_brace")
    return nodes, members, ["A1", "B1", "C1"], ["A4", "B4", "C4"]


def _leg_names(n):
    letters = string.ascii_uppercase
    if n <= 26:
        return list(letters[:n])
    return [letters[i // 26 - 1] + letters[i % 26] if i >= 26 else letters[i] for i in range(n)]


def generate_jacket(n_legs, n_bays, r_bottom=30.0, r_top=10.0, z_bottom=-50.0, z_top=10.0):
    """Synthetic jacket in the reference's conventions: legs on a circle that
    tapers from r_bottom (mudline, z_bottom) to r_top (z_top); every bay of every
    face is X-braced through a hinge node at the mean of its four corners; a ring
    of horizontal braces at every level.  Members = L(6B+1), nodes = L(2B+1).
    Node order: legs (leg by leg, bottom to top), then hinges level by level --
    the same ordering style as the default model, so the natural K ordering is
    far from banded (that is what jk_set_supports' RCM option is for)."""
    if n_legs < 3 or n_bays < 1:
        raise ValueError("generate_jacket needs n_legs >= 3 and n_bays >= 1")
    legs = _leg_names(n_legs)
    nodes = {}
    for j, leg in enumerate(legs):
        ang = 2.0 * np.pi * j / n_legs
        for lvl in range(n_bays + 1):
            f = lvl / n_bays
            r = r_bottom + (r_top - r_bottom) * f
            nodes[f"{leg}{lvl + 1}"] = np.array([r * np.cos(ang), r * np.sin(ang), z_bottom + (z_top - z_bottom) * f])
    faces = [(legs[j], legs[(j + 1) % n_legs]) for j in range(n_legs)]
    for lvl in range(1, n_bays + 1):
        for p, q in faces:
            corners = [nodes[f"{p}{lvl}"], nodes[f"{q}{lvl}"], nodes[f"{p}{lvl + 1}"], nodes[f"{q}{lvl + 1}"]]
            nodes[f"H{p}{q}{lvl}"] = np.mean(corners, axis=0)
    members = []

    def add(prefix, a, b, kind):
        members.append({"name": f"{prefix}_{a}-{b}", "node1": a, "node2": b, "type": kind})

    for leg in legs:
        for i in range(1, n_bays + 1):
            add("Leg", f"{leg}{i}", f"{leg}{i + 1}", "leg")
    for lvl in range(1, n_bays + 2):
        for p, q in faces:
            add("HBrace", f"{p}{lvl}", f"{q}{lvl}", "h_brace")
    for lvl in range(1, n_bays + 1):
        for p, q in faces:
            hinge = f"H{p}{q}{lvl}"
            add("XBr", f"{p}{lvl}", hinge, "x_brace")
            add("XBr", hinge, f"{q}{lvl + 1}", "x_brace")
            add("XBr", f"{q}{lvl}", hinge, "x_brace")
            add("XBr", hinge, f"{p}{lvl + 1}", "x_brace")
    fixed = [f"{leg}1" for leg in legs]
    top = [f"{leg}{n_bays + 1}" for leg in legs]
    return nodes, members, fixed, top


def default_sections(rho_steel=7850):
    """GUI defaults (GUI.py:1808-1809): leg 2000x75, brace 800x30 mm."""
    return TubularSection(2000, 75, "Leg", rho_steel), TubularSection(800, 30, "Brace", rho_steel)
