"""Host-side helpers that need no GPU: the phase-degree column (GUI.py:697-698) and the memoised static load
(GUI.py:1962-2015)."""
import numpy as np

from conftest import load_golden, product_structure


def test_mod360_is_bit_identical_to_python_float_mod():
    from jacket_b200.morison import fill_phase_deg, mod360
    rng = np.random.default_rng(7)
    x = np.concatenate([rng.uniform(-5000, 5000, 20000), [0.0, -0.0, 360.0, -360.0, 720.0, 359.99999999999994, -1e-300]])
    got = mod360(x)
    assert np.array_equal(got, x % 360)
    assert all(float(g) == float(v) % 360 for g, v in zip(got[:3000], x[:3000]))       # the reference's scalar expression
    assert np.all((got >= 0) & (got <= 360))          # -1e-300 % 360 rounds to 360.0 in Python as well
    omega = 2 * np.pi / 9.4
    table = np.zeros((360, 16))
    table[:, 0] = [i * 9.4 / 360 for i in range(360)]
    fill_phase_deg(table, omega)
    assert all(table[i, 1] == np.degrees(omega * table[i, 0]) % 360 for i in range(360))


def test_static_load_memo_returns_private_copies():
    import jacket_b200 as jb
    from jacket_b200.analysis import _static_load
    st, ap = product_structure(load_golden("gen4x3_airy"))
    a = jb.static_load(st, ap)
    assert np.array_equal(a, _static_load(st, ap))
    a[:] = 0.0                                     # callers may scribble on what they get
    b = jb.static_load(st, ap)
    assert np.array_equal(b, _static_load(st, ap)) and b is not a
    ap2 = jb.AnalysisParams(**{**ap.__dict__, "F_shear": ap.F_shear + 100.0})
    c = jb.static_load(st, ap2)
    assert not np.array_equal(b, c) and np.array_equal(c, _static_load(st, ap2))
