import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
GOLDEN_CASES = ("default3_airy", "gen4x3_airy", "gen5x6_airy")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False))


def golden_params(g):
    p = {}
    for k, v in g.items():
        if k.startswith("param_"):
            p[k[6:]] = v.item() if v.ndim == 0 else v
    return p


def oracle_model(g):
    from oracle import jacket_oracle as orc
    p = golden_params(g)
    sections = [(p["D_leg"], p["t_leg"], p["rho_steel"]), (p["D_brace"], p["t_brace"], p["rho_steel"])]
    sec_id = np.where(g["is_leg"], 0, 1)
    return orc.Model(g["xyz"], g["conn"], sec_id, sections, g["fixed"], g["top"])


def product_structure(g):
    """The golden case rebuilt through the product's reference-style classes."""
    import jacket_b200 as jb
    p = golden_params(g)
    names = [str(n) for n in g["node_names"]]
    nodes = {n: np.array(g["xyz"][i]) for i, n in enumerate(names)}
    members = [{"name": str(g["member_names"][i]), "node1": names[g["conn"][i, 0]], "node2": names[g["conn"][i, 1]],
                "type": str(g["member_types"][i])} for i in range(g["conn"].shape[0])]
    fixed = [names[i] for i in g["fixed"]]
    top = [names[i] for i in g["top"]]
    ap = jb.AnalysisParams(**{k: p[k] for k in ("E", "nu", "fy", "rho_steel", "rho_water", "D_leg", "t_leg", "D_brace",
                                                 "t_brace", "H", "T", "d", "U_c", "wave_dir", "current_dir", "Cd", "Cm",
                                                 "F_axial", "F_shear", "M_moment", "M_torsion", "custom_sw")},
                           self_weight_mode=str(p["self_weight_mode"]), wave_model="Airy")
    return jb.build_structure(nodes, members, fixed, top, ap), ap


def relmax(a, b):
    """max|a-b| / max|b| -- the parity metric of SURVEY 7 (hard part 3)."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    den = np.max(np.abs(b))
    return float(np.max(np.abs(a - b)) / den) if den > 0 else float(np.max(np.abs(a - b)))


@pytest.fixture(scope="session", params=GOLDEN_CASES)
def golden(request):
    return request.param, load_golden(request.param)


def generated_case(g, legs, bays):
    """(jb, structure, params, wave kwargs) of a generator jacket whose inputs are stored in a large golden file; asserts
    that the regenerated geometry is the one the reference was run on."""
    import jacket_b200 as jb
    p = golden_params(g)
    ap = jb.AnalysisParams(**{k: p[k] for k in ("E", "nu", "fy", "rho_steel", "rho_water", "D_leg", "t_leg", "D_brace", "t_brace", "H", "T",
                                                 "d", "U_c", "wave_dir", "current_dir", "Cd", "Cm", "F_axial", "F_shear", "M_moment",
                                                 "M_torsion", "custom_sw")}, self_weight_mode=str(p["self_weight_mode"]), wave_model="Airy")
    nodes, members, fixed, top = jb.generate_jacket(legs, bays)
    st = jb.build_structure(nodes, members, fixed, top, ap)
    xyz, conn, sec_id, _, _ = st.pack()
    assert np.array_equal(xyz, g["xyz"]) and np.array_equal(conn, g["conn"]) and np.array_equal(sec_id == 0, g["is_leg"])
    assert np.array_equal(st.indices(fixed), g["fixed"]) and np.array_equal(st.indices(top), g["top"])
    return jb, st, ap, fixed


def fem_summary_columns(U, reactions, rows):
    """Columns 8..15 of the per-phase table from one phase's full results (run_analysis' log lines GUI.py:2027-2054):
    max |translation| and its node, max utilisation and its member, its von Mises stress, sum of the reactions."""
    tr = np.linalg.norm(np.asarray(U).reshape(-1, 6)[:, :3], axis=1)
    util = np.asarray(rows)[:, 6]
    m = int(np.argmax(util))
    R = np.asarray(reactions)[:, :3].sum(axis=0)
    return np.array([tr.max(), float(np.argmax(tr)), util[m], float(m), np.asarray(rows)[m, 5], R[0], R[1], R[2]])
