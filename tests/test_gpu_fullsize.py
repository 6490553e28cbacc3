"""Size-independent properties at BASELINE.json's full size (c4: 10,000 members, 19,968 free DOF, 4,096 phases) and
edge cases (single phase, ragged phase counts, tiny structures).  Needs a B200."""
import numpy as np
import pytest

from conftest import relmax

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def c4():
    import jacket_b200 as jb
    ap = jb.AnalysisParams(wave_model="Airy")
    nodes, members, fixed, top = jb.generate_jacket(16, 104)
    st = jb.build_structure(nodes, members, fixed, top, ap)
    wave = jb.RaschiiWave(ap.H, ap.T, ap.d, ap.U_c, "Airy")
    res = jb.phase_scan(st, wave, 4096, wave_direction=ap.wave_dir, current_direction=ap.current_dir, params=ap)
    return jb, ap, st, wave, res, fixed


def test_c4_dimensions_and_band(c4):
    jb, ap, st, wave, res, fixed = c4
    d = res.engine.dims()
    assert d["n_members"] == 10000 and d["n_nodes"] == 3344 and d["n_free_dof"] == 19968
    assert d["n_chains"] == 2 and d["band_tiles"] <= 5 and d["dof_half_bandwidth"] < 300      # RCM: a narrow band
    assert res.table.shape == (4096, 16) and np.all(np.isfinite(res.table))


def test_c4_residual_and_equilibrium(c4):
    jb, ap, st, wave, res, fixed = c4
    assert res.engine.residual() < 1e-9                       # max|K u - F| / max|F| over all 4096 phases
    F_static = jb.static_load(st, ap).reshape(-1, 6)
    for i in (0, 1234, res.critical_index, 4095):
        ph = res.phase(i)
        applied = F_static[:, :3].sum(axis=0) + ph["nodal_forces"].sum(axis=0)
        R = np.array([ph["reactions"][n] for n in fixed])[:, :3].sum(axis=0)
        assert np.max(np.abs(applied + R)) < 1e-9 * np.max(np.abs(applied))           # sum R + sum F = 0
        row = res.row(i)
        assert relmax([row["sum_Rx"], row["sum_Ry"], row["sum_Rz"]], R) < 1e-12
        # Morison totals = sum of nodal loads
        assert relmax(ph["nodal_forces"].sum(axis=0), np.array([row["Fx_kN"], row["Fy_kN"], row["Fz_kN"]]) * 1000) < 1e-10
        util = np.array([r["utilization"] for r in ph["internal_forces"]])
        assert abs(row["max_util"] - util.max()) <= 1e-12 * util.max() and int(row["max_util_member"]) == int(np.argmax(util))
        tr = np.linalg.norm(ph["U"].reshape(-1, 6)[:, :3], axis=1)
        assert abs(row["max_disp_mm"] - tr.max()) <= 1e-12 * tr.max() and int(row["max_disp_node"]) == int(np.argmax(tr))


def test_c4_critical_phase_is_first_maximum_and_deterministic(c4):
    jb, ap, st, wave, res, fixed = c4
    assert res.critical_index == int(np.argmax(res.table[:, 2]))
    again, crit = res.engine.phase_scan(res.table[:, 0].copy(), ap.fy)
    assert crit == res.critical_index and np.array_equal(again[:, 2:], res.table[:, 2:])     # bit-identical rerun


def test_c4_linearity_of_the_solve(c4):
    """K is factored once: solving a*F1 + b*F2 equals a*U1 + b*U2 (multi-RHS path, caller-built loads)."""
    jb, ap, st, wave, res, fixed = c4
    rng = np.random.default_rng(7)
    F = rng.normal(0, 1e5, (3, st.n_dof))
    F[2] = 2.5 * F[0] - 0.75 * F[1]
    eng = res.engine
    eng.solve(F, ap.fy)
    U = [eng.fetch_phase(i, U=True, reactions=False, rows=False)["U"] for i in range(3)]
    assert relmax(U[2], 2.5 * U[0] - 0.75 * U[1]) < 1e-9
    assert eng.residual() < 1e-9


@pytest.mark.parametrize("P", [1, 2, 31, 33, 97])
def test_ragged_phase_counts_vs_oracle(P):
    import jacket_b200 as jb
    from oracle import jacket_oracle as orc
    ap = jb.AnalysisParams(wave_model="Airy")
    nodes, members, fixed, top = jb.create_default_3leg_jacket()
    st = jb.build_structure(nodes, members, fixed, top, ap)
    res = jb.phase_scan(st, jb.RaschiiWave(ap.H, ap.T, ap.d, ap.U_c, "Airy"), P, wave_direction=ap.wave_dir,
                        current_direction=ap.current_dir, params=ap)
    xyz, conn, sec_id, _, sections = st.pack()
    model = orc.Model(xyz, conn, sec_id, [(s.D_outer, s.t, s.rho_steel) for s in sections], st.indices(fixed), st.indices(top))
    ref = orc.phase_scan(model, orc.AiryWave(ap.H, ap.T, ap.d, ap.U_c), orc.phase_times(ap.T, P), wave_direction=ap.wave_dir,
                         current_direction=ap.current_dir, Cd=ap.Cd, Cm=ap.Cm, rho_water=ap.rho_water, E=ap.E, nu=ap.nu,
                         fy=ap.fy, F_axial_kN=ap.F_axial, F_shear_kN=ap.F_shear, self_weight="calculated")
    assert res.table.shape == (P, 16) and res.critical_index == ref["critical"]
    for c in range(2, 8):
        assert relmax(res.table[:, c], ref["table"][:, c]) < 1e-9
    assert relmax(res.phase(P - 1)["U"], ref["U"][P - 1]) < 1e-9
    assert relmax(res.table[:, 10], ref["members"]["utilization"].max(axis=1)) < 1e-9


def test_single_member_cantilever_known_answer():
    """One element, clamped at one end, tip load P: delta = P L^3 / (3 E I) + P L / (G A_s)  (oracle-free check)."""
    import jacket_b200 as jb
    L_m, P_N = 12.0, 5.0e4
    sec = jb.TubularSection(800, 30)
    nodes = {"A": np.array([0.0, 0.0, -20.0]), "B": np.array([L_m, 0.0, -20.0])}
    st = jb.CustomJacketStructure(nodes, [{"name": "m", "node1": "A", "node2": "B", "type": "brace"}], sec, sec, ["A"], ["B"])
    fem = jb.FEMSolver(st, 210000, 0.3)
    fem.apply_nodal_force("B", np.array([0.0, 0.0, -P_N, 0, 0, 0]))
    fem.apply_boundary_conditions(["A"])
    U = fem.solve()
    Lmm = L_m * 1000.0
    G = 210000 / 2.6
    expect = P_N * Lmm**3 / (3 * 210000 * sec.Iy_mm4) + P_N * Lmm / (G * sec.Az_mm2)
    assert abs(-U[6 + 2] - expect) < 1e-9 * expect
    R = fem.get_reactions()["A"]
    assert abs(R[2] - P_N) < 1e-6 * P_N and abs(abs(R[4]) - P_N * Lmm) < 1e-6 * P_N * Lmm
    rows = fem.get_member_internal_forces(355)
    # a horizontal member's local y axis is global z (GUI.py:380-382), so the shear / moment appear as Fy / Mz
    shear = max(rows[0]["Fy_max_kN"], rows[0]["Fz_max_kN"]); moment = max(rows[0]["My_max_kNm"], rows[0]["Mz_max_kNm"])
    assert abs(shear - P_N / 1000) < 1e-9 * P_N / 1000 and abs(moment - P_N * Lmm / 1e6) < 1e-9 * P_N * Lmm / 1e6
    assert rows[0]["Fx_max_kN"] < 1e-9 * P_N / 1000
