"""CPU-side checks: the shared library loads and exports every symbol the header declares, fails loudly without a
GPU, and the host-side logic (sections, default geometry, generator, static loads, dispersion) matches the vectors
produced by the reference."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT, golden_params, load_golden, product_structure, relmax


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "jacket_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(jk_[a-z_0-9]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    from jacket_b200 import _lib
    lib = _lib.lib()
    names = _declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/jacket_b200.h but not exported"
    assert lib.jk_version() >= 100
    # constants shared between header and binding
    hdr = open(os.path.join(ROOT, "include", "jacket_b200.h")).read()
    for macro, val in (("JK_TABLE_NCOL", _lib.TABLE_NCOL), ("JK_MEMBER_NCOL", _lib.MEMBER_NCOL),
                       ("JK_DETAIL_NCOL", _lib.DETAIL_NCOL), ("JK_SEC_NPROP", _lib.SEC_NPROP), ("JK_NTIMERS", _lib.NTIMERS)):
        assert int(re.search(rf"#define\s+{macro}\s+(\d+)", hdr).group(1)) == val
    assert len(_lib.TABLE_COLUMNS) == _lib.TABLE_NCOL and len(_lib.TIMER_NAMES) == _lib.NTIMERS


def test_no_cpu_fallback_without_device():
    """On a box without a GPU the product must fail loudly (this test is skipped where a GPU exists)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import jacket_b200 as jb
    g = load_golden("default3_airy")
    st, ap = product_structure(g)
    with pytest.raises(jb.JacketError) as e:
        jb.MorisonCalculator(st, jb.RaschiiWave(ap.H, ap.T, ap.d, ap.U_c), ap.wave_dir, ap.current_dir).find_critical_phase(36)
    assert "JK_ENODEVICE" in str(e.value)
    with pytest.raises(jb.JacketError):
        jb.FEMSolver(st, ap.E, ap.nu)


def test_create_rejects_bad_geometry_before_touching_the_device():
    from jacket_b200 import _lib
    lib = _lib.lib()
    h = C.c_void_p()
    xyz = np.zeros((2, 3)); conn = np.array([[0, 5]], dtype=np.int32); sec = np.zeros(1, dtype=np.int32); props = np.ones((1, 8))
    rc = lib.jk_create(0, None, 2, _lib.dptr(xyz), 1, _lib.iptr(conn), _lib.iptr(sec), 1, _lib.dptr(props), C.byref(h))
    assert rc != 0 and not h.value
    assert lib.jk_last_error(None)


def test_sections_match_reference_values():
    import jacket_b200 as jb
    g = load_golden("default3_airy")
    p = golden_params(g)
    leg = jb.TubularSection(p["D_leg"], p["t_leg"], "Leg", p["rho_steel"])
    # closed forms (GUI.py:122-137)
    Do, Di = 2000.0, 1850.0
    assert leg.D_inner == Di and leg.R_outer == 1000.0
    assert leg.Ax_mm2 == np.pi / 4.0 * (Do**2 - Di**2)
    assert leg.Iy_mm4 == leg.Iz_mm4 == np.pi / 64.0 * (Do**4 - Di**4)
    assert leg.Ix_mm4 == np.pi / 32.0 * (Do**4 - Di**4) and leg.Ay_mm2 == 0.5 * leg.Ax_mm2
    assert leg.mass_per_m == leg.Ax_mm2 / 1e6 * 7850.0 and leg.D_t_ratio == Do / 75.0
    st = leg.calc_stress_at_point(1e6, 2e5, -3e5, 4e9, 5e9, -6e9, "A2")
    assert set(st) == {"sigma_total", "tau_total", "von_mises"} and st["von_mises"] > 0
    assert list(leg.get_stress_points()) == [f"A{i}" for i in range(1, 9)]


def test_default_geometry_and_generator():
    import jacket_b200 as jb
    g = load_golden("default3_airy")
    nodes, members, fixed, top = jb.create_default_3leg_jacket(47.0)
    assert list(nodes) == [str(n) for n in g["node_names"]]
    assert np.array_equal(np.array(list(nodes.values())), g["xyz"])
    assert [m["name"] for m in members] == [str(n) for n in g["member_names"]]
    assert [m["type"] for m in members] == [str(n) for n in g["member_types"]]
    assert fixed == ["A1", "B1", "C1"] and top == ["A4", "B4", "C4"]
    for L, B in ((4, 3), (8, 41), (16, 104)):
        n, m, f, t = jb.generate_jacket(L, B)
        assert len(m) == L * (6 * B + 1) and len(n) == L * (2 * B + 1) and len(f) == len(t) == L
    with pytest.raises(ValueError):
        jb.generate_jacket(2, 3)


def test_static_load_matches_reference_F(golden):
    """interface loads + self-weight (GUI.py:1962-2015): F_reference - Morison nodal loads == static_load()."""
    import jacket_b200 as jb
    _, g = golden
    st, ap = product_structure(g)
    F = jb.static_load(st, ap).reshape(-1, 6).copy()
    F[:, :3] += g["mor_t0_nodal"]
    assert relmax(F.ravel(), g["fem_t0_F"]) < 1e-14


def test_wave_host_side(golden):
    import jacket_b200 as jb
    _, g = golden
    _, ap = product_structure(g)
    for model in ("Airy", "Fenton", "Stokes", "auto"):
        w = jb.RaschiiWave(ap.H, ap.T, ap.d, ap.U_c, model, 10)
        assert w.k == g["wave_k"].item() and w.omega == g["wave_omega"].item() and w.L == g["wave_L"].item()
        assert w.actual_model == "Airy (fallback)" and "Steepness H/L=" in w.get_model_info()
    assert np.array_equal(jb.phase_times(9.4, 36), np.array([i * 9.4 / 36 for i in range(36)]))


def test_structure_pack_roundtrip():
    import jacket_b200 as jb
    g = load_golden("gen4x3_airy")
    st, _ = product_structure(g)
    xyz, conn, sec_id, props, sections = st.pack()
    assert np.array_equal(xyz, g["xyz"]) and np.array_equal(conn, g["conn"])
    assert np.array_equal(sec_id == sec_id[0], g["is_leg"] == g["is_leg"][0]) and props.shape == (2, 8)
    assert st.n_dof == 6 * st.n_nodes and st.get_member_geometry(st.members[0])["L_mm"] > 0


def test_occupancy_critical_kernels_keep_their_register_budget():
    """The Morison kernel is occupancy-bound (5 blocks of 128 threads per SM need <= 102 registers; with 224 it runs 15 %
    slower) and a sweep CTA of 544 threads needs <= 120: read the built library's resource usage so that a stray
    __launch_bounds__ or a code change that blows the budget fails here, on the CPU, before any GPU time is spent."""
    import re
    import shutil
    import subprocess
    from conftest import ROOT
    import os
    tool = shutil.which("cuobjdump")
    lib = os.path.join(ROOT, "small-fem-solver-based-on-a-lot-of-assumption_b200", "libjacket_b200.so")
    if tool is None or not os.path.isfile(lib):
        import pytest
        pytest.skip("cuobjdump or the built library is missing")
    out = subprocess.run([tool, "-res-usage", lib], capture_output=True, text=True, timeout=300).stdout
    regs = {}
    name = None
    for line in out.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            name = m.group(1)
        m = re.search(r"REG:(\d+)", line)
        if m and name:
            regs[name] = int(m.group(1))
    morison = [v for k, v in regs.items() if "k_morison_airyILb0ELi15ELb0" in k]
    sweep = [v for k, v in regs.items() if "k_sweepILi4ELb0" in k]
    post = [v for k, v in regs.items() if "k_member_post" in k and "single" not in k]
    assert morison and max(morison) <= 102, morison
    assert sweep and max(sweep) <= 120, sweep
    assert post and max(post) <= 64, post
