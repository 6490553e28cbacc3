"""The drop-in boundary: jacket_b200's analysis classes expose every function of the reference's classes
(GUI.py:115-803) with the same parameter names, order and defaults, and the host-side point queries / element
constructor return the reference's values.  CPU only: the surface is compared with tests/golden/reference_surface.json
(written by tests/golden/make_surface.py from the reference) and, where /root/reference exists, with the live module."""
import inspect
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN_DIR, load_golden, relmax

KIN_KEYS = ("u_wave", "v_wave", "w_wave", "u_current", "v_current", "du_dt", "dv_dt", "dw_dt", "submerged", "eta")


def _sig(fn):
    return "(" + ", ".join(p.name if p.default is inspect.Parameter.empty else f"{p.name}={p.default!r}"
                           for p in inspect.signature(fn).parameters.values()) + ")"


def _diffs(surface):
    import jacket_b200 as jb
    out = []
    for cname, methods in surface["classes"].items():
        cls = getattr(jb, cname, None)
        if cls is None:
            out.append(f"class {cname} missing")
            continue
        for mname, want in methods.items():
            fn = inspect.getattr_static(cls, mname, None)
            if fn is None or not callable(getattr(cls, mname)):
                out.append(f"{cname}.{mname} missing (reference {want})")
            elif _sig(getattr(cls, mname)) != want:
                out.append(f"{cname}.{mname}{_sig(getattr(cls, mname))} != reference {want}")
    for fname, want in surface["functions"].items():
        fn = getattr(jb, fname, None)
        if fn is None or _sig(fn) != want:
            out.append(f"{fname}: {None if fn is None else _sig(fn)} != reference {want}")
    for k, v in surface["constants"].items():
        if getattr(jb, k, None) != v:
            out.append(f"constant {k}: {getattr(jb, k, None)} != {v}")
    return out


def test_class_surface_matches_recorded_reference_surface():
    with open(os.path.join(GOLDEN_DIR, "reference_surface.json")) as f:
        surface = json.load(f)
    assert sum(len(m) for m in surface["classes"].values()) == 31
    assert _diffs(surface) == []


def test_class_surface_matches_live_reference():
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip("reference not present (GPU box)")
    import sys
    sys.path.insert(0, GOLDEN_DIR)
    import make_surface
    surface = make_surface.surface(ref_loader.load())
    with open(os.path.join(GOLDEN_DIR, "reference_surface.json")) as f:
        assert json.load(f) == json.loads(json.dumps(surface))          # the committed file is current
    assert _diffs(surface) == []


class _HostOnlyMorison:
    """MorisonCalculator.get_kinematics_3d is host code that needs no GPU handle: build the object without one."""

    def __new__(cls, wave, wave_dir, current_dir):
        import jacket_b200 as jb
        return jb.MorisonCalculator(None, wave, wave_dir, current_dir)


@pytest.mark.parametrize("tag", ["a", "b"])
def test_point_kinematics_vs_reference_vectors(tag):
    """RaschiiWave.eta / velocity / acceleration / get_kinematics and MorisonCalculator.get_kinematics_3d against the
    reference's own values, including the points that dry out within dt (acceleration spikes of thousands of m/s^2)."""
    import jacket_b200 as jb
    g = {k[2:]: v for k, v in load_golden("kinematics_airy").items() if k.startswith(tag + "_")}
    H, T, d, U_c, wave_dir, current_dir = g["params"]
    wave = jb.RaschiiWave(H, T, d, U_c, "Airy", 10)
    mor = _HostOnlyMorison(wave, wave_dir, current_dir)
    t = float(g["t"])
    assert int(g["spikes"]) > 10
    eta = np.array([wave.eta(x, t) for x in g["xw"]])
    assert np.array_equal(eta, g["eta"])                                     # same expression -> same bits
    vel = np.array([wave.velocity(x, z, t) for x, z in zip(g["xw"], g["points"][:, 2])])
    acc = np.array([wave.acceleration(x, z, t) for x, z in zip(g["xw"], g["points"][:, 2])])
    assert relmax(vel, g["velocity"]) < 1e-14 and relmax(acc, g["acceleration"]) < 1e-11
    assert np.array_equal(vel == 0.0, g["velocity"] == 0.0)                  # the same points are dry
    kin = [mor.get_kinematics_3d(*p, t) for p in g["points"]]
    assert all(tuple(k) == KIN_KEYS for k in kin)
    got = np.array([[float(k[c]) for c in KIN_KEYS] for k in kin])
    assert np.array_equal(got[:, 8], g["kin3"][:, 8])
    for c in range(10):
        assert relmax(got[:, c], g["kin3"][:, c]) < 1e-11, KIN_KEYS[c]
    k2 = wave.get_kinematics(float(g["xw"][1]), float(g["points"][1, 2]), t)
    assert tuple(k2) == ("u", "w", "du_dt", "dw_dt", "submerged", "eta")
    dry = wave.get_kinematics(0.0, 100.0, t)
    assert dry["submerged"] is False and dry["u"] == 0 and dry["dw_dt"] == 0


def test_beam_element_constructor_vs_reference_vectors():
    """BeamElement3D(node1, node2, section, E, G) built on its own (GUI.py:361-422) against the reference's element
    matrices of the default jacket (golden Ke / Kl / T3) and its end-force recovery (GUI.py:424-432)."""
    import jacket_b200 as jb
    g = load_golden("default3_airy")
    E, nu = float(g["param_E"]), float(g["param_nu"])
    G = E / (2 * (1 + nu))
    leg = jb.TubularSection(float(g["param_D_leg"]), float(g["param_t_leg"]), "Leg")
    brace = jb.TubularSection(float(g["param_D_brace"]), float(g["param_t_brace"]), "Brace")
    rng = np.random.default_rng(3)
    for m in range(g["conn"].shape[0]):
        a, b = g["conn"][m]
        el = jb.BeamElement3D(g["xyz"][a], g["xyz"][b], leg if g["is_leg"][m] else brace, E, G)
        assert np.array_equal(el.K_local, g["Kl"][m])                        # same products in the same order
        assert np.array_equal(el.T[:3, :3], g["T3"][m]) and np.array_equal(el.T[9:, 9:], g["T3"][m])
        assert relmax(el.K_global, g["Ke"][m]) < 1e-15
        assert el.L_mm == el.L * 1000.0
    u = rng.normal(0, 1, 12)
    f = el.get_internal_forces(u)
    F = g["Kl"][m] @ (np.kron(np.eye(4), g["T3"][m]) @ u)
    assert [f["node1"][k] for k in ("Fx", "Fy", "Fz", "Mx", "My", "Mz")] == list(-F[:6])
    assert [f["node2"][k] for k in ("Fx", "Fy", "Fz", "Mx", "My", "Mz")] == list(F[6:])
    # vertical member: local y from z^ x lx; without shear deformation Phi = 0
    v = jb.BeamElement3D(np.array([1.0, 2.0, -10.0]), np.array([1.0, 2.0, 5.0]), brace, include_shear=False)
    assert np.allclose(v.T[:3, :3], [[0, 0, 1], [0, 1, 0], [-1, 0, 0]]) and v.G == 80769
    EI = v.E * brace.Iz_mm4
    assert v.K_local[5, 5] == 4.0 * (EI / v.L_mm**3) * v.L_mm**2 and v.K_local[5, 11] == 2.0 * (EI / v.L_mm**3) * v.L_mm**2


def test_beam_element_constructor_vs_live_reference():
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip("reference not present (GPU box)")
    import jacket_b200 as jb
    ref = ref_loader.load()
    rng = np.random.default_rng(5)
    for i in range(40):
        a, b = rng.uniform(-30, 30, 3), rng.uniform(-30, 30, 3)
        if i % 5 == 0:
            b[:2] = a[:2] + rng.uniform(-1e-3, 1e-3, 2) * (i % 2)          # vertical and nearly vertical members
        D, t = rng.uniform(400, 2500), rng.uniform(10, 90)
        shear = bool(i % 3)
        ours = jb.BeamElement3D(a, b, jb.TubularSection(D, t), 205000, 79000, shear)
        theirs = ref.BeamElement3D(a, b, ref.TubularSection(D, t), 205000, 79000, shear)
        assert np.array_equal(ours.K_local, theirs.K_local)
        assert relmax(ours.T, theirs.T) < 1e-15 and relmax(ours.K_global, theirs.K_global) < 1e-14
        u = rng.normal(0, 1, 12)
        fo, ft = ours.get_internal_forces(u), theirs.get_internal_forces(u)
        for node in ("node1", "node2"):
            assert list(fo[node]) == list(ft[node])
            assert relmax([fo[node][k] for k in fo[node]], [ft[node][k] for k in ft[node]]) < 1e-13


def test_wave_private_helpers_keep_the_reference_contract():
    import jacket_b200 as jb
    w = jb.RaschiiWave(17.038, 9.4, 50.0, 1.7)
    assert w._solve_dispersion(w.omega, 50.0) == w.k == 0.046430003773529516
    fit = w._create_wave("auto", 10)                                              # own fits; sets actual_model / actual_N like GUI.py:208-253
    assert (w.actual_model, w.actual_N) == ("Fenton", 20) and fit.omega == 2 * np.pi / 9.4 and hasattr(fit, "length") and hasattr(fit, "c")
