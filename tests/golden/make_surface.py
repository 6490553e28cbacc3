"""Reference call surface and point-kinematics vectors, produced by the reference itself (build container only):

    python tests/golden/make_surface.py

* reference_surface.json -- inspect.signature of every function of the analysis classes (GUI.py:115-803): the
  conformance test compares jacket_b200's classes with it where /root/reference is absent (GPU box).
* kinematics_airy.npz    -- RaschiiWave.eta / velocity / acceleration and MorisonCalculator.get_kinematics_3d
  (GUI.py:259-296, 559-589) at random points and times, including points within dt of leaving the water (the
  finite-difference spike of SURVEY F2) -- for two wave / current heading pairs.
"""
from __future__ import annotations

import inspect
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402

CLASSES = ("TubularSection", "RaschiiWave", "CustomJacketStructure", "BeamElement3D", "FEMSolver", "MorisonCalculator")
FUNCTIONS = ("create_default_3leg_jacket",)
KIN_KEYS = ("u_wave", "v_wave", "w_wave", "u_current", "v_current", "du_dt", "dv_dt", "dw_dt", "submerged", "eta")


def signature_text(fn):
    """Parameter names, kinds and defaults (annotations dropped: the reference's dataclass spells them as objects, a
    module with `from __future__ import annotations` as strings)."""
    parts = []
    for p in inspect.signature(fn).parameters.values():
        parts.append(p.name if p.default is inspect.Parameter.empty else f"{p.name}={p.default!r}")
    return "(" + ", ".join(parts) + ")"


def surface(module):
    out = {"classes": {}, "functions": {}}
    for c in CLASSES:
        cls = getattr(module, c)
        out["classes"][c] = {name: signature_text(fn) for name, fn in inspect.getmembers(cls, inspect.isfunction)
                             if not (name.startswith("__") and name != "__init__")}
    for f in FUNCTIONS:
        out["functions"][f] = signature_text(getattr(module, f))
    out["constants"] = {k: getattr(module, k) for k in ("g", "DEFAULT_RHO_WATER", "DEFAULT_E", "DEFAULT_NU", "DEFAULT_FY", "DEFAULT_RHO_STEEL")}
    return out


def kinematics_case(ref, H, T, d, U_c, wave_dir, current_dir, seed, n=400):
    rng = np.random.default_rng(seed)
    wave = ref.RaschiiWave(H, T, d, U_c, "Airy", 10)
    nodes, members, fixed, top = ref.create_default_3leg_jacket(47.0)
    leg, brace = ref.TubularSection(2000, 75, "Leg"), ref.TubularSection(800, 30, "Brace")
    st = ref.CustomJacketStructure({k: np.array(v) for k, v in nodes.items()}, members, leg, brace, fixed, top)
    mor = ref.MorisonCalculator(st, wave, wave_dir, current_dir)
    pts = np.column_stack([rng.uniform(-40, 40, n), rng.uniform(-40, 40, n), rng.uniform(-d, 1.2 * H / 2, n)])
    t = float(rng.uniform(0, T))
    # a third of the points sit just below the surface at t: some of them are dry at t + dt (acceleration spike)
    cw, sw = np.cos(mor.theta_wave), np.sin(mor.theta_wave)
    for i in range(0, n, 3):
        xw = pts[i, 0] * cw + pts[i, 1] * sw
        pts[i, 2] = wave.eta(xw, t) - rng.uniform(0.0, 2e-3)
    kin3 = np.array([[float(mor.get_kinematics_3d(*p, t)[k]) for k in KIN_KEYS] for p in pts])
    xw = pts[:, 0] * cw + pts[:, 1] * sw
    eta = np.array([wave.eta(x, t) for x in xw])
    vel = np.array([wave.velocity(x, z, t) for x, z in zip(xw, pts[:, 2])])
    acc = np.array([wave.acceleration(x, z, t) for x, z in zip(xw, pts[:, 2])])
    spikes = int(np.sum(np.abs(kin3[:, 5:8]).max(axis=1) > 50.0))
    return dict(params=np.array([H, T, d, U_c, wave_dir, current_dir]), t=np.array(t), points=pts, xw=xw, kin3=kin3, eta=eta,
                velocity=vel, acceleration=acc, spikes=np.array(spikes))


def main():
    ref = ref_loader.load()
    with open(os.path.join(HERE, "reference_surface.json"), "w") as f:
        json.dump(surface(ref), f, indent=1, sort_keys=True)
    out = {}
    for tag, args in (("a", (17.038, 9.4, 50.0, 1.7, 38.0, 38.0, 11)), ("b", (9.5, 11.0, 50.0, 0.9, 20.0, 75.0, 12))):
        for k, v in kinematics_case(ref, *args).items():
            out[f"{tag}_{k}"] = v
    np.savez_compressed(os.path.join(HERE, "kinematics_airy.npz"), **out)
    print("spike points:", int(out["a_spikes"]), int(out["b_spikes"]))


if __name__ == "__main__":
    main()
