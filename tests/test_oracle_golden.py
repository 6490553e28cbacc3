"""The oracle (oracle/jacket_oracle.py) pinned against vectors produced by the reference's own classes
(tests/golden/make_golden.py).  CPU only.  Tolerances are far below the 1e-9 parity bar because the oracle
follows the reference's operation order."""
import numpy as np

from conftest import golden_params, load_golden, oracle_model, relmax
from oracle import jacket_oracle as orc

TIGHT = 5e-13
TOL = 1e-9          # the north star's bar (LU per case in the reference vs one LU for all cases in the oracle: ~1e-11)
MEMBER_KEYS = ("Fx_max_kN", "Fy_max_kN", "Fz_max_kN", "My_max_kNm", "Mz_max_kNm", "von_mises_max_MPa", "utilization")


def _wave(p):
    return orc.AiryWave(p["H"], p["T"], p["d"], p["U_c"])


def _mor_kw(p):
    return dict(wave_direction=p["wave_dir"], current_direction=p["current_dir"], Cd=p["Cd"], Cm=p["Cm"],
                rho_water=p["rho_water"])


def test_dispersion_and_pinned_constants(golden):
    name, g = golden
    p = golden_params(g)
    w = _wave(p)
    assert w.k == g["wave_k"].item() and w.omega == g["wave_omega"].item() and w.L == g["wave_L"].item()
    if name == "default3_airy":   # values quoted in BASELINE.md section 3
        assert w.k == 0.046430003773529516 and w.omega == 0.6684239688488921 and w.L == 135.32597020295162


def test_morison_single_phase(golden):
    _, g = golden
    p = golden_params(g)
    m = oracle_model(g)
    for tag in ("t0", "t1"):
        out = orc.morison_phases(m, _wave(p), [g[f"mor_{tag}_t"].item()], want_details=True, **_mor_kw(p))
        tot = np.concatenate([out["total_drag"][0], out["total_inertia"][0], out["total_morison"][0]])
        assert relmax(out["nodal_forces"][0], g[f"mor_{tag}_nodal"]) < TIGHT
        assert relmax(tot, g[f"mor_{tag}_totals"]) < TIGHT
        assert relmax(out["member_details"][0], g[f"mor_{tag}_details"]) < TIGHT


def test_pinned_t0_total(golden):
    name, g = golden
    if name != "default3_airy":
        return
    np.testing.assert_allclose(g["mor_t0_totals"][6:9], [3392202.8797103227, 4344665.94447964, -433561.4819033571], rtol=1e-15)


def test_phase_scan_tables_and_critical_index(golden):
    name, g = golden
    p = golden_params(g)
    m = oracle_model(g)
    for key in [k for k in g if k.startswith("scan") and k.endswith("_table")]:
        n = int(key[4:-6])
        t = orc.phase_times(p["T"], n)
        out = orc.morison_phases(m, _wave(p), t, **_mor_kw(p))
        tab, crit = orc.phase_table(out, t, g["wave_omega"].item())
        assert crit == int(g[f"scan{n}_critical"])             # bit-exact index
        assert np.array_equal(tab[:, 0], g[key][:, 0])           # t_i
        assert np.array_equal(tab[:, 1], g[key][:, 1])           # phase_deg
        assert relmax(tab[:, 2:], g[key][:, 2:]) < TIGHT
    if name == "default3_airy":
        assert int(g["scan36_critical"]) == 35 and int(g["scan360_critical"]) == 353
        assert abs(g["scan36_table"][35, 2] - 5799.185270933482) < 1e-9


def test_elements_and_assembly(golden):
    _, g = golden
    p = golden_params(g)
    fem = orc.FEM(oracle_model(g), p["E"], p["nu"])
    assert relmax(fem.K_local, g["Kl"]) < 1e-15
    assert relmax(fem.R, g["T3"]) < 1e-15
    assert relmax(fem.K_elem, g["Ke"]) < 1e-14
    if g["K_global"].size:
        assert relmax(fem.K_global, g["K_global"]) < 1e-14


def _fem_inputs(g, p, fem, nodal):
    inter, sw = fem.static_loads(p["wave_dir"], p["F_axial"], p["F_shear"], p["M_moment"], p["M_torsion"],
                                 str(p["self_weight_mode"]), p["custom_sw"])
    return fem.load_matrix(nodal, inter, sw)


def test_fem_t0(golden):
    name, g = golden
    p = golden_params(g)
    m = oracle_model(g)
    fem = orc.FEM(m, p["E"], p["nu"])
    F = _fem_inputs(g, p, fem, g["mor_t0_nodal"][None])
    assert relmax(F[0], g["fem_t0_F"]) < 1e-14
    U = fem.solve(F)
    assert relmax(U[0], g["fem_t0_U"]) < 1e-10
    assert relmax(fem.reactions(U, F)[0], g["fem_t0_reactions"]) < 1e-10
    mf = fem.member_forces(U, p["fy"])
    rows = np.stack([mf[k][0] for k in ("Fx_max_kN", "Fy_max_kN", "Fz_max_kN", "My_max_kNm", "Mz_max_kNm",
                                        "von_mises_max_MPa", "utilization")], axis=1)
    for c in range(7):
        assert relmax(rows[:, c], g["fem_t0_rows"][:, c]) < 1e-10
    assert relmax(mf["length_m"], g["fem_t0_length_m"]) < 1e-15
    if name == "default3_airy":
        assert abs(np.max(g["fem_t0_rows"][:, 6]) - 0.2147147837812134) < 1e-12
        tr = np.linalg.norm(g["fem_t0_U"].reshape(-1, 6)[:, :3], axis=1)
        assert abs(np.max(tr) - 68.22044893416346) < 1e-9


def test_per_phase_fem(golden):
    _, g = golden
    p = golden_params(g)
    m = oracle_model(g)
    P = int(g["phasefem_P"])
    t = orc.phase_times(p["T"], P)[g["phasefem_idx"]]
    res = orc.phase_scan(m, _wave(p), t, E=p["E"], nu=p["nu"], fy=p["fy"], F_axial_kN=p["F_axial"],
                         F_shear_kN=p["F_shear"], M_moment_kNm=p["M_moment"], M_torsion_kNm=p["M_torsion"],
                         self_weight=str(p["self_weight_mode"]), custom_sw_tonnes=p["custom_sw"], **_mor_kw(p))
    assert relmax(res["F"], g["phasefem_F"]) < 1e-14
    for i in range(len(t)):
        assert relmax(res["U"][i], g["phasefem_U"][i]) < 1e-10
        assert relmax(res["reactions"][i], g["phasefem_reactions"][i]) < 1e-10
        for c, k in enumerate(("Fx_max_kN", "Fy_max_kN", "Fz_max_kN", "My_max_kNm", "Mz_max_kNm",
                               "von_mises_max_MPa", "utilization")):
            assert relmax(res["members"][k][i], g["phasefem_rows"][i][:, c]) < 1e-10


def test_equilibrium_invariant(golden):
    """sum of reactions + sum of applied forces = 0 (independent of any oracle)."""
    _, g = golden
    F = g["phasefem_F"].reshape(g["phasefem_F"].shape[0], -1, 6)
    R = g["phasefem_reactions"]
    applied = F[:, :, :3].sum(axis=1)
    # loads on the fixed nodes are part of F and also appear (negated) in R = K U - F
    total = applied + R[:, :, :3].sum(axis=1)
    assert np.max(np.abs(total)) / np.max(np.abs(applied)) < 1e-9


def test_c3_size_jacket_against_the_reference():
    """BASELINE configs[2] / [4] geometry (8 legs x 41 bays: 1,976 members, 3,936 free DOF): Morison at two times, a 4- and a 16-phase
    scan, the t = 0 FEM case and three per-phase FEM cases computed by the reference's own classes
    (tests/golden/make_golden.py --large) against the oracle.  Element matrices are not stored for this case."""
    from conftest import load_golden
    g = load_golden("gen8x41_airy")
    p = golden_params(g)
    m = oracle_model(g)
    assert g["conn"].shape == (1976, 2) and g["xyz"].shape == (664, 3)
    w = _wave(p)
    assert abs(w.k - g["wave_k"]) < 1e-15 and abs(w.omega - g["wave_omega"]) < 1e-15
    for tag in ("t0", "t1"):
        out = orc.morison_phases(m, w, np.array([float(g[f"mor_{tag}_t"])]), **_mor_kw(p))
        assert relmax(out["nodal_forces"][0], g[f"mor_{tag}_nodal"]) < TIGHT
        totals = np.concatenate([out["total_drag"][0], out["total_inertia"][0], out["total_morison"][0]])
        assert relmax(totals, g[f"mor_{tag}_totals"]) < TIGHT
    for n in (4, 16):
        tn = orc.phase_times(p["T"], n)
        tab, crit = orc.phase_table(orc.morison_phases(m, w, tn, **_mor_kw(p)), tn, w.omega)
        assert crit == int(g[f"scan{n}_critical"])
        assert np.array_equal(tab[:, :2], g[f"scan{n}_table"][:, :2]) and relmax(tab[:, 2:], g[f"scan{n}_table"][:, 2:]) < TIGHT
    fem = orc.FEM(m, p["E"], p["nu"])
    F = _fem_inputs(g, p, fem, g["mor_t0_nodal"][None])
    assert relmax(F[0], g["fem_t0_F"]) < 1e-14
    U = fem.solve(F)
    assert relmax(U[0], g["fem_t0_U"]) < 1e-9                 # two LU factorisations of a 3,936 x 3,936 system
    assert relmax(fem.reactions(U, F)[0], g["fem_t0_reactions"]) < 1e-9
    mf = fem.member_forces(U, p["fy"])
    for c, k in enumerate(("Fx_max_kN", "Fy_max_kN", "Fz_max_kN", "My_max_kNm", "Mz_max_kNm", "von_mises_max_MPa", "utilization")):
        assert relmax(mf[k][0], g["fem_t0_rows"][:, c]) < 1e-9
    P = int(g["phasefem_P"])
    t = orc.phase_times(p["T"], P)[g["phasefem_idx"]]
    res = orc.phase_scan(m, w, t, E=p["E"], nu=p["nu"], fy=p["fy"], F_axial_kN=p["F_axial"], F_shear_kN=p["F_shear"],
                         M_moment_kNm=p["M_moment"], M_torsion_kNm=p["M_torsion"], self_weight=str(p["self_weight_mode"]),
                         custom_sw_tonnes=p["custom_sw"], fem=fem, **_mor_kw(p))
    assert relmax(res["F"], g["phasefem_F"]) < 1e-14
    for i in range(len(t)):
        assert relmax(res["U"][i], g["phasefem_U"][i]) < 1e-9
        assert relmax(res["reactions"][i], g["phasefem_reactions"][i]) < 1e-9
        assert relmax(res["members"]["utilization"][i], g["phasefem_rows"][i][:, 6]) < 1e-9


# ---------------------------------------------------------------------------------------------------------------------
# BASELINE configs[2] / [3] at full size: vectors produced by the reference's own classes (tests/golden/make_golden_large.py)
# ---------------------------------------------------------------------------------------------------------------------
def _large_oracle_model(g):
    from oracle import jacket_oracle as orc
    p = golden_params(g)
    sections = [(p["D_leg"], p["t_leg"], p["rho_steel"]), (p["D_brace"], p["t_brace"], p["rho_steel"])]
    model = orc.Model(g["xyz"], g["conn"], np.where(g["is_leg"], 0, 1), sections, g["fixed"], g["top"])
    return orc, p, model, orc.AiryWave(p["H"], p["T"], p["d"], p["U_c"])


def test_oracle_full_1024_phase_scan_at_c3_size():
    """configs[2]: every row of the reference's 1,024-phase Morison scan of the 1,976-member jacket, and its critical index."""
    g = load_golden("gen8x41_scan1024")
    orc, p, model, wave = _large_oracle_model(g)
    t = orc.phase_times(p["T"], 1024)
    mor = orc.morison_phases(model, wave, t, p["wave_dir"], p["current_dir"], p["Cd"], p["Cm"], p["rho_water"])
    tab, crit = orc.phase_table(mor, t, wave.omega)
    assert crit == int(g["scan1024_critical"]) == 1009
    assert np.array_equal(tab[:, :2], g["scan1024_table"][:, :2])          # t and phase_deg: same expressions, same bits
    for c in range(2, 8):
        assert relmax(tab[:, c], g["scan1024_table"][:, c]) < 1e-13


LU_NOISE_C4 = 3e-8   # measured: how far the reference's own LU results at 19,968 DOF are from the converged solution of its equations


def test_oracle_at_c4_size_vs_reference():
    """configs[3]: 70 rows of the 4,096-phase Morison scan (incl. the critical phase 4044 and its neighbours) and 8 full FEM
    cases (U, reactions, member rows, 12 end forces) of the reference at 10,000 members / 19,968 free DOF; the committed
    oracle table of all 4,096 phases (c4_oracle_scan4096.npz) agrees with the reference on those rows.

    LU noise floor: at this size numpy.linalg.solve is only reproducible to ~2e-9 on U and ~8e-9 on the shear forces -- the
    reference's per-case dgesv and the oracle's multi-right-hand-side dgesv (same algorithm, different blocking) differ by
    that much, and BOTH are up to 7e-9 / 2e-8 away from the converged solution of the same equations (LU + iterative
    refinement with 80-bit residuals, FEM.solve_refined; the relative residual drops from 4e-8 to 2e-11).  So at c4 the
    1e-9 bar is asserted against the converged solution (the GPU path in tests/test_gpu_reference_sizes.py) and the raw
    reference output is matched to its own noise, LU_NOISE_C4."""
    g = load_golden("gen16x104_c4")
    fx = load_golden("c4_oracle_scan4096")
    idx = g["scan4096_idx"]
    assert int(fx["critical"]) == 4044 and int(idx[np.argmax(g["scan4096_rows"][:, 2])]) == 4044
    assert np.array_equal(fx["table"][idx, :2], g["scan4096_rows"][:, :2])
    for c in range(2, 8):
        assert relmax(fx["table"][idx, c], g["scan4096_rows"][:, c]) < 1e-13
    orc, p, model, wave = _large_oracle_model(g)
    fi = g["phasefem_idx"]
    t = orc.phase_times(p["T"], 4096)[fi]
    assert np.array_equal(t, g["phasefem_t"])
    mor = orc.morison_phases(model, wave, t, p["wave_dir"], p["current_dir"], p["Cd"], p["Cm"], p["rho_water"])
    assert relmax(mor["nodal_forces"], g["phasefem_nodal"]) < 1e-13
    tot = np.concatenate([mor["total_drag"], mor["total_inertia"], mor["total_morison"]], axis=1)
    assert relmax(tot, g["phasefem_totals"]) < 1e-13
    fem = orc.FEM(model, p["E"], p["nu"])
    inter, sw = fem.static_loads(p["wave_dir"], p["F_axial"], p["F_shear"], p["M_moment"], p["M_torsion"], str(p["self_weight_mode"]))
    F = fem.load_matrix(mor["nodal_forces"], inter, sw)
    U = fem.solve_refined(F, steps=2)
    assert fem.refine_history[0] > 1e-9 and fem.refine_history[-1] < 1e-10       # plain LU leaves a 4e-8 residual; refinement removes it
    R, mf = fem.reactions(U, F), fem.member_forces(U, p["fy"])
    worst = 0.0
    for k in range(len(fi)):
        errs = [relmax(g["phasefem_U"][k], U[k]), relmax(g["phasefem_reactions"][k], R[k])]
        errs += [relmax(g["phasefem_rows"][k][:, j], mf[key][k]) for j, key in enumerate(MEMBER_KEYS)]
        worst = max(worst, max(errs))
        assert max(errs) < LU_NOISE_C4, (int(fi[k]), errs)
    assert worst > TOL                                  # ... and it IS noise above the 1e-9 bar: the reference cannot define 1e-9 here
    for k, i in enumerate(g["endforce_idx"]):
        kk = int(np.flatnonzero(fi == i)[0])
        for blk in (slice(0, 3), slice(3, 6), slice(6, 9), slice(9, 12)):      # forces and moments of node 1 / node 2
            assert relmax(mf["end_forces"][kk][:, blk], g["end_forces"][k][:, blk]) < LU_NOISE_C4


def test_oracle_end_forces_small_case():
    """a20 (GUI.py:424-432): the 12 end forces of every member of the 4 x 3 generator jacket at 4 phases."""
    g = load_golden("gen4x3_endforces")
    orc, p, model, wave = _large_oracle_model(g)
    t = orc.phase_times(p["T"], int(g["phasefem_P"]))[g["phasefem_idx"]]
    res = orc.phase_scan(model, wave, t, wave_direction=p["wave_dir"], current_direction=p["current_dir"], Cd=p["Cd"], Cm=p["Cm"],
                         rho_water=p["rho_water"], E=p["E"], nu=p["nu"], fy=p["fy"], F_axial_kN=p["F_axial"], F_shear_kN=p["F_shear"],
                         self_weight="calculated")
    assert relmax(res["U"], g["phasefem_U"]) < TOL
    for blk in (slice(0, 3), slice(3, 6), slice(6, 9), slice(9, 12)):
        assert relmax(res["members"]["end_forces"][:, :, blk], g["end_forces"][:, :, blk]) < TOL
