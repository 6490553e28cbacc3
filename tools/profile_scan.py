"""Minimal c4 driver for ncu captures: setup, then N full steps (assemble, factor, phase scan)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import jacket_b200 as jb
legs, bays, P, steps = 16, 104, 4096, 2
if len(sys.argv) > 4:
    legs, bays, P, steps = (int(a) for a in sys.argv[1:5])
p = jb.AnalysisParams(wave_model="Airy")
nodes, members, fixed, top = jb.generate_jacket(legs, bays)
st = jb.build_structure(nodes, members, fixed, top, p)
wave = jb.RaschiiWave(p.H, p.T, p.d, p.U_c, "Airy", p.N_harm)
eng = jb.get_engine(st)
eng.set_supports(st.indices(fixed))
eng.set_static_load(jb.static_load(st, p))
eng.set_wave(wave)
eng.set_morison(np.deg2rad(90 - p.wave_dir), np.deg2rad(90 - p.current_dir), p.rho_water, p.Cd, p.Cm, 15)
t = jb.phase_times(wave.T, P)
for i in range(steps):
    eng.assemble(p.E, p.E / 2.6)
    eng.factor(overlap=False)
    table, crit = eng.phase_scan(t, p.fy)
print("critical", crit, {k: round(v, 3) for k, v in eng.timings().items()})
