"""Generate the golden vectors under tests/golden/ by EXECUTING THE REFERENCE'S OWN CLASSES.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

Every case stores its inputs as plain arrays plus the reference outputs, so the
oracle (oracle/jacket_oracle.py) and the CUDA path can be checked on the GPU box
where the reference does not exist.  The per-phase FEM rows are the reference's
run_analysis sequence (GUI.py:1955-2024) replayed with t_analysis = t_i
 -- nothing is computed by code of this repo except the synthetic
input geometry of the `gen*` cases.
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

GUI_DEFAULTS = dict(E=210000.0, nu=0.3, fy=355.0, rho_steel=7850.0, rho_water=1025.0,
                    D_leg=2000.0, t_leg=75.0, D_brace=800.0, t_brace=30.0,
                    H=17.038, T=9.4, d=50.0, U_c=1.7, wave_dir=38.0, current_dir=38.0,
                    Cd=0.7, Cm=2.0, F_axial=25100.0, F_shear=2900.0, M_moment=0.0, M_torsion=0.0,
                    self_weight_mode="calculated", custom_sw=1100.0)


def reference_fem_case(ref, structure, morison_results, p):
    """GUI.py:1955-2024 verbatim in effect: build FEMSolver, loads, BCs, solve, reactions, member rows."""
    fem = ref.FEMSolver(structure, p["E"], p["nu"])
    top_nodes = structure.get_top_nodes()
    n_legs = len(top_nodes)
    theta = np.deg2rad(90.0 - p["wave_dir"])
    for node in top_nodes:
        force = np.array([p["F_shear"] * 1000.0 * np.cos(theta) / n_legs,
                          p["F_shear"] * 1000.0 * np.sin(theta) / n_legs,
                          -(p["F_axial"] * 1000.0) / n_legs,
                          p["M_torsion"] * 1e6 / n_legs,
                          p["M_moment"] * 1e6 / n_legs, 0.0])
        fem.apply_nodal_force(node, force)
    for node_name, force in morison_results["nodal_forces"].items():
        fv = np.zeros(6)
        fv[:3] = force[:3]
        fem.apply_nodal_force(node_name, fv)
    if p["self_weight_mode"] == "calculated":
        for member in structure.members:
            geom = structure.get_member_geometry(member)
            w = member["section"].mass_per_m * ref.g
            F_weight = w * geom["L"] / 2.0
            fem.F_global[6 * structure.node_index[member["node1"]] + 2] -= F_weight
            fem.F_global[6 * structure.node_index[member["node2"]] + 2] -= F_weight
    elif p["self_weight_mode"] == "custom":
        sw_per_node = p["custom_sw"] * 1000 * ref.g / structure.n_nodes
        for i in range(structure.n_nodes):
            fem.F_global[6 * i + 2] -= sw_per_node
    fem.apply_boundary_conditions(structure.get_bottom_nodes())
    U = fem.solve()
    reactions = fem.get_reactions()
    rows = fem.get_member_internal_forces(p["fy"])
    return fem, U.copy(), reactions, rows


ROW_KEYS = ("Fx_max_kN", "Fy_max_kN", "Fz_max_kN", "My_max_kNm", "Mz_max_kNm", "von_mises_max_MPa", "utilization")
DET_KEYS = ("drag_kN", "inertia_kN", "total_kN", "submerged_length")
PH_KEYS = ("t", "phase_deg", "total_kN", "drag_kN", "inertia_kN", "Fx_kN", "Fy_kN", "Fz_kN")


def make_case(ref, name, nodes, members, fixed, top, p, scans, fem_phases_of, n_fem_phases, store_elements=True):
    t0 = time.time()
    leg = ref.TubularSection(p["D_leg"], p["t_leg"], "Leg", p["rho_steel"])
    brace = ref.TubularSection(p["D_brace"], p["t_brace"], "Brace", p["rho_steel"])
    st = ref.CustomJacketStructure({k: np.array(v) for k, v in nodes.items()}, members, leg, brace, fixed, top, p["rho_steel"])
    wave = ref.RaschiiWave(p["H"], p["T"], p["d"], p["U_c"], "Airy", 10)
    mor = ref.MorisonCalculator(st, wave, p["wave_dir"], p["current_dir"], p["Cd"], p["Cm"], p["rho_water"])
    out = {}
    # inputs
    out["xyz"] = np.array([st.nodes[n] for n in st.node_list])
    out["conn"] = np.array([[st.node_index[m["node1"]], st.node_index[m["node2"]]] for m in st.members], dtype=np.int64)
    out["is_leg"] = np.array([m["type"] == "leg" for m in st.members])
    out["fixed"] = np.array([st.node_index[n] for n in fixed], dtype=np.int64)
    out["top"] = np.array([st.node_index[n] for n in top], dtype=np.int64)
    out["node_names"] = np.array(st.node_list)
    out["member_names"] = np.array([m["name"] for m in st.members])
    out["member_types"] = np.array([m["type"] for m in st.members])
    for k, v in p.items():
        out["param_" + k] = np.array(v)
    # wave
    out["wave_k"], out["wave_omega"], out["wave_L"] = np.array(wave.k), np.array(wave.omega), np.array(wave.L)
    # single-phase Morison at t = 0 and t = 2.35
    for tag, t in (("t0", 0.0), ("t1", 2.35)):
        r = mor.compute_all_morison_forces(t)
        out[f"mor_{tag}_t"] = np.array(t)
        out[f"mor_{tag}_nodal"] = np.array([r["nodal_forces"][n][:3] for n in st.node_list])
        out[f"mor_{tag}_totals"] = np.concatenate([r["total_drag"], r["total_inertia"], r["total_morison"]])
        out[f"mor_{tag}_details"] = np.array([[d[k] for k in DET_KEYS] for d in r["member_details"]])
        if tag == "t0":
            fem, U, reac, rows = reference_fem_case(ref, st, r, p)
            out["fem_t0_F"] = fem.F_global.copy()
            out["fem_t0_U"] = U
            out["fem_t0_reactions"] = np.array([reac[n] for n in fixed])
            out["fem_t0_rows"] = np.array([[row[k] for k in ROW_KEYS] for row in rows])
            out["fem_t0_length_m"] = np.array([row["length_m"] for row in rows])
            out["K_global"] = fem.K_global.copy() if fem.n_dof <= 200 else np.zeros(0)
            if store_elements:   # 2.3 MB each at 2k members: the small cases pin the element matrices
                out["Ke"] = np.array([e.K_global for e in fem.elements])
                out["Kl"] = np.array([e.K_local for e in fem.elements])
                out["T3"] = np.array([e.T[:3, :3] for e in fem.elements])
    # Morison scans
    for n_steps in scans:
        res = mor.find_critical_phase(n_steps=n_steps)
        tab = np.array([[row[k] for k in PH_KEYS] for row in res["all_phases"]])
        out[f"scan{n_steps}_table"] = tab
        out[f"scan{n_steps}_critical"] = np.array(res["all_phases"].index(res["critical"]))
    # per-phase FEM: n_fem_phases phases evenly taken from the `fem_phases_of`-step scan
    P = fem_phases_of
    idx = np.unique(np.linspace(0, P - 1, n_fem_phases).astype(int))
    Us, Rs, rows_all, Fs = [], [], [], []
    for i in idx:
        t = i * p["T"] / P
        r = mor.compute_all_morison_forces(t)
        fem, U, reac, rows = reference_fem_case(ref, st, r, p)
        Us.append(U); Fs.append(fem.F_global.copy())
        Rs.append(np.array([reac[n] for n in fixed]))
        rows_all.append(np.array([[row[k] for k in ROW_KEYS] for row in rows]))
    out["phasefem_P"] = np.array(P)
    out["phasefem_idx"] = idx
    out["phasefem_F"] = np.array(Fs)
    out["phasefem_U"] = np.array(Us)
    out["phasefem_reactions"] = np.array(Rs)
    out["phasefem_rows"] = np.array(rows_all)
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"{name}: {len(st.node_list)} nodes, {len(st.members)} members -> {path} "
          f"({os.path.getsize(path) / 1024:.0f} KiB, {time.time() - t0:.1f}s)")


def main():
    ref = ref_loader.load()
    sys.path.insert(0, ROOT)
    import jacket_b200 as jb   # only for the synthetic input geometry of the gen* cases (host code, no GPU)

    if "--only-large" in sys.argv:
        sys.argv.append("--large")
        nodes, members, fixed, top = jb.generate_jacket(8, 41)
        make_case(ref, "gen8x41_airy", nodes, members, fixed, top, dict(GUI_DEFAULTS), scans=(4, 16), fem_phases_of=16, n_fem_phases=3,
                  store_elements=False)
        return
    # case 1: the reference's own default model with the GUI defaults (BASELINE configs[0], [1] on the pinned Airy path)
    nodes, members, fixed, top = ref.create_default_3leg_jacket(47.0)
    make_case(ref, "default3_airy", nodes, members, fixed, top, dict(GUI_DEFAULTS), scans=(36, 360),
              fem_phases_of=36, n_fem_phases=12)

    # case 2: 4-leg x 3-bay generator, separate wave / current headings, custom self-weight, interface moments
    p = dict(GUI_DEFAULTS)
    p.update(H=9.5, T=11.0, U_c=0.9, wave_dir=20.0, current_dir=75.0, self_weight_mode="custom", custom_sw=800.0,
             M_moment=15000.0, M_torsion=2500.0, Cd=1.05, Cm=1.6)
    nodes, members, fixed, top = jb.generate_jacket(4, 3)
    make_case(ref, "gen4x3_airy", nodes, members, fixed, top, p, scans=(24,), fem_phases_of=24, n_fem_phases=24)

    # case 3: 5-leg x 6-bay generator, no current, no self-weight, deep draft wave (exercises dry members, many tiles)
    p = dict(GUI_DEFAULTS)
    p.update(H=6.0, T=8.0, U_c=0.0, wave_dir=300.0, current_dir=0.0, self_weight_mode="none",
             F_axial=9000.0, F_shear=700.0, D_leg=1500.0, t_leg=50.0, D_brace=600.0, t_brace=20.0)
    nodes, members, fixed, top = jb.generate_jacket(5, 6, r_bottom=22.0, r_top=9.0, z_bottom=-50.0, z_top=14.0)
    make_case(ref, "gen5x6_airy", nodes, members, fixed, top, p, scans=(16,), fem_phases_of=16, n_fem_phases=8)

    # case 4 (`--large`): BASELINE configs[2]/[4] geometry -- 8 legs x 41 bays, 1,976 members, 3,936 free DOF -- with the GUI
    # defaults; a 4- and a 16-phase Morison scan and 3 per-phase FEM solves by the reference (about 1.5 minutes of its Python loops)
    if "--large" in sys.argv:
        nodes, members, fixed, top = jb.generate_jacket(8, 41)
        make_case(ref, "gen8x41_airy", nodes, members, fixed, top, dict(GUI_DEFAULTS), scans=(4, 16), fem_phases_of=16, n_fem_phases=3,
                  store_elements=False)


if __name__ == "__main__":
    main()
