"""Golden vectors at BASELINE.json's full sizes, produced by EXECUTING THE REFERENCE'S OWN CLASSES in parallel processes.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden_large.py small   # gen4x3_endforces.npz  (76 members: the 12 member end forces, GUI.py:424-432)
    python tests/golden/make_golden_large.py c3      # gen8x41_scan1024.npz  (configs[2]: 1,976 members, 1,024-phase Morison scan)
    python tests/golden/make_golden_large.py c4      # gen16x104_c4.npz      (configs[3]: 10,000 members, phases of the 4,096-phase scan)

MorisonCalculator.find_critical_phase (GUI.py:684-724) is a loop over independent phases, so the phases are dealt to worker
processes; every worker builds the reference's objects itself and runs the reference's loop body (GUI.py:696-714) for its
phases -- the rows are exactly what the single-process loop would append.  The per-phase FEM cases are the reference's
run_analysis sequence (GUI.py:1955-2024, `make_golden.reference_fem_case`) with t_analysis = t_i, plus the 12 member end
forces of BeamElement3D.get_internal_forces (GUI.py:424-432) taken the way get_member_internal_forces does (GUI.py:506-511).
Nothing is computed by code of this repo except the synthetic input geometry (jacket_b200.generate_jacket, host code).
"""
from __future__ import annotations

import multiprocessing as mp
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from make_golden import GUI_DEFAULTS, PH_KEYS, ROW_KEYS, reference_fem_case  # noqa: E402

FORCE_KEYS = ("Fx", "Fy", "Fz", "Mx", "My", "Mz")


def _reference_objects(legs, bays, p):
    from oracle import ref_loader
    import jacket_b200 as jb
    ref = ref_loader.load()
    nodes, members, fixed, top = jb.generate_jacket(legs, bays)
    leg = ref.TubularSection(p["D_leg"], p["t_leg"], "Leg", p["rho_steel"])
    brace = ref.TubularSection(p["D_brace"], p["t_brace"], "Brace", p["rho_steel"])
    st = ref.CustomJacketStructure({k: np.array(v) for k, v in nodes.items()}, members, leg, brace, fixed, top, p["rho_steel"])
    wave = ref.RaschiiWave(p["H"], p["T"], p["d"], p["U_c"], "Airy", 10)
    mor = ref.MorisonCalculator(st, wave, p["wave_dir"], p["current_dir"], p["Cd"], p["Cm"], p["rho_water"])
    return ref, st, wave, mor, fixed, top


def _scan_rows(args):
    """Loop body of find_critical_phase (GUI.py:696-714) for the phases i in idx of an n_steps scan."""
    legs, bays, p, n_steps, idx = args
    ref, st, wave, mor, _, _ = _reference_objects(legs, bays, p)
    T, omega = wave.T, wave.omega
    rows = []
    for i in idx:
        t = i * T / n_steps
        phase = omega * t
        phase_deg = np.degrees(phase) % 360
        forces = mor.compute_all_morison_forces(t)
        rows.append([t, phase_deg, np.linalg.norm(forces["total_morison"]) / 1000, np.linalg.norm(forces["total_drag"]) / 1000,
                     np.linalg.norm(forces["total_inertia"]) / 1000, forces["total_morison"][0] / 1000,
                     forces["total_morison"][1] / 1000, forces["total_morison"][2] / 1000])
    return list(idx), rows


def _fem_case(args):
    """One per-phase FEM case by the reference + the 12 end forces of every member."""
    legs, bays, p, n_steps, i, want_end_forces = args
    os.environ.setdefault("OPENBLAS_NUM_THREADS", "2")
    ref, st, wave, mor, fixed, top = _reference_objects(legs, bays, p)
    t = i * p["T"] / n_steps
    r = mor.compute_all_morison_forces(t)
    fem, U, reac, rows = reference_fem_case(ref, st, r, p)
    out = {"i": i, "t": t,
           "nodal": np.array([r["nodal_forces"][n][:3] for n in st.node_list]),
           "totals": np.concatenate([r["total_drag"], r["total_inertia"], r["total_morison"]]),
           "U": U, "reactions": np.array([reac[n] for n in fixed]),
           "rows": np.array([[row[k] for k in ROW_KEYS] for row in rows])}
    if want_end_forces:
        ef = np.zeros((len(st.members), 12))
        for m, member in enumerate(st.members):                               # GUI.py:506-511
            i1, i2 = st.node_index[member["node1"]], st.node_index[member["node2"]]
            u_elem = np.concatenate([U[6 * i1:6 * i1 + 6], U[6 * i2:6 * i2 + 6]])
            f = fem.elements[m].get_internal_forces(u_elem)
            ef[m, :6] = [f["node1"][k] for k in FORCE_KEYS]
            ef[m, 6:] = [f["node2"][k] for k in FORCE_KEYS]
        out["end_forces"] = ef
    return out


def _inputs(st, fixed, top, p):
    out = {"xyz": np.array([st.nodes[n] for n in st.node_list]),
           "conn": np.array([[st.node_index[m["node1"]], st.node_index[m["node2"]]] for m in st.members], dtype=np.int64),
           "is_leg": np.array([m["type"] == "leg" for m in st.members]),
           "fixed": np.array([st.node_index[n] for n in fixed], dtype=np.int64),
           "top": np.array([st.node_index[n] for n in top], dtype=np.int64)}
    for k, v in p.items():
        out["param_" + k] = np.array(v)
    return out


def scan_parallel(legs, bays, p, n_steps, idx, workers):
    idx = list(idx)
    chunks = [idx[w::workers] for w in range(workers)]
    with mp.get_context("fork").Pool(workers) as pool:
        parts = pool.map(_scan_rows, [(legs, bays, p, n_steps, c) for c in chunks if c])
    rows = {}
    for ids, rs in parts:
        rows.update(dict(zip(ids, rs)))
    return np.array([rows[i] for i in idx])


def make_c3(workers):
    """configs[2]: the FULL 1,024-phase Morison scan of the 8 x 41 jacket (about an hour of the reference's loops, dealt over
    `workers` processes) + its critical index (first maximum, GUI.py:717)."""
    t0 = time.time()
    p = dict(GUI_DEFAULTS)
    legs, bays, P = 8, 41, 1024
    tab = scan_parallel(legs, bays, p, P, range(P), workers)
    crit = int(max(range(P), key=lambda i: tab[i, 2]))                        # max(results, key=...) -> first maximum
    _, st, wave, _, fixed, top = _reference_objects(legs, bays, p)
    out = _inputs(st, fixed, top, p)
    out.update(scan1024_table=tab, scan1024_critical=np.array(crit), wave_k=np.array(wave.k), wave_omega=np.array(wave.omega))
    path = os.path.join(HERE, "gen8x41_scan1024.npz")
    np.savez_compressed(path, **out)
    print(f"c3: 1024-phase scan, critical {crit} ({tab[crit, 2]:.3f} kN) -> {path} ({os.path.getsize(path) / 1024:.0f} KiB, {time.time() - t0:.0f}s)")


def make_c4(workers, fem_phases, scan_phases):
    """configs[3]: rows `scan_phases` of the 4,096-phase Morison scan and full per-phase FEM cases at `fem_phases`
    (the reference needs ~45 s and ~10 GB per FEM case at 19,968 free DOF: dense K, LU per case)."""
    t0 = time.time()
    p = dict(GUI_DEFAULTS)
    legs, bays, P = 16, 104, 4096
    tab = scan_parallel(legs, bays, p, P, scan_phases, workers)
    print(f"c4: {len(scan_phases)} scan rows in {time.time() - t0:.0f}s", flush=True)
    with mp.get_context("fork").Pool(min(workers, 4)) as pool:                # 4 x ~10 GB
        cases = pool.map(_fem_case, [(legs, bays, p, P, i, k < 2) for k, i in enumerate(fem_phases)], chunksize=1)
    _, st, wave, _, fixed, top = _reference_objects(legs, bays, p)
    out = _inputs(st, fixed, top, p)
    out.update(scan4096_idx=np.array(list(scan_phases)), scan4096_rows=tab, phasefem_P=np.array(P),
               phasefem_idx=np.array([c["i"] for c in cases]), phasefem_t=np.array([c["t"] for c in cases]),
               phasefem_nodal=np.array([c["nodal"] for c in cases]), phasefem_totals=np.array([c["totals"] for c in cases]),
               phasefem_U=np.array([c["U"] for c in cases]), phasefem_reactions=np.array([c["reactions"] for c in cases]),
               phasefem_rows=np.array([c["rows"] for c in cases]),
               endforce_idx=np.array([c["i"] for c in cases if "end_forces" in c]),
               end_forces=np.array([c["end_forces"] for c in cases if "end_forces" in c]),
               wave_k=np.array(wave.k), wave_omega=np.array(wave.omega))
    path = os.path.join(HERE, "gen16x104_c4.npz")
    np.savez_compressed(path, **out)
    print(f"c4: {len(cases)} FEM cases -> {path} ({os.path.getsize(path) / 1024:.0f} KiB, {time.time() - t0:.0f}s)")


def make_small():
    """4-leg x 3-bay generator jacket (76 members): every member's 12 end forces (a20, GUI.py:424-432) at 4 phases of a 24-phase scan."""
    p = dict(GUI_DEFAULTS)
    legs, bays, P = 4, 3, 24
    cases = [_fem_case((legs, bays, p, P, i, True)) for i in (0, 7, 13, 23)]
    _, st, wave, _, fixed, top = _reference_objects(legs, bays, p)
    out = _inputs(st, fixed, top, p)
    out.update(phasefem_P=np.array(P), phasefem_idx=np.array([c["i"] for c in cases]), phasefem_U=np.array([c["U"] for c in cases]),
               phasefem_rows=np.array([c["rows"] for c in cases]), phasefem_reactions=np.array([c["reactions"] for c in cases]),
               end_forces=np.array([c["end_forces"] for c in cases]))
    path = os.path.join(HERE, "gen4x3_endforces.npz")
    np.savez_compressed(path, **out)
    print(f"small: {path} ({os.path.getsize(path) / 1024:.0f} KiB)")


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "c3"
    workers = int(os.environ.get("JK_GOLDEN_WORKERS", max(1, (os.cpu_count() or 2) - 1)))
    if which == "small":
        make_small()
    elif which == "c3":
        make_c3(workers)
    elif which == "c4":
        # critical phase of the 4,096-phase scan (oracle estimate, passed on the command line) first: it gets the end forces
        crit = int(sys.argv[2])
        fem = [crit, 1234] + [i for i in (0, 585, 1755, 2340, 2925, 4095) if i not in (crit, 1234)]
        scan = sorted(set(list(range(0, 4096, 64)) + [crit - 2, crit - 1, crit, crit + 1, crit + 2, 4095]) & set(range(4096)))
        make_c4(workers, fem, scan)
