"""c4 sweep diagnostics: per-warp clock breakdown of the TMA sweep (option profile_sweep) and the factor chain (profile_chol)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import jacket_b200 as jb
legs, bays, P = 16, 104, 4096
opts = {"profile_sweep": 1}
for a in sys.argv[1:]:
    k, v = a.split("=")
    if k == "P": P = int(v)
    elif k == "legs": legs = int(v)
    elif k == "bays": bays = int(v)
    else: opts[k] = int(v)
p = jb.AnalysisParams(wave_model="Airy")
nodes, members, fixed, top = jb.generate_jacket(legs, bays)
st = jb.build_structure(nodes, members, fixed, top, p)
wave = jb.RaschiiWave(p.H, p.T, p.d, p.U_c, "Airy", p.N_harm)
eng = jb.get_engine(st, options=opts)
eng.set_supports(st.indices(fixed))
eng.set_static_load(jb.static_load(st, p))
eng.set_wave(wave)
eng.set_morison(np.deg2rad(90 - p.wave_dir), np.deg2rad(90 - p.current_dir), p.rho_water, p.Cd, p.Cm, 15)
t = jb.phase_times(wave.T, P)
for i in range(2):
    eng.assemble(p.E, p.E / 2.6)
    eng.factor(overlap=False)
    table, crit = eng.phase_scan(t, p.fy)
print("critical", crit, {k: round(v, 3) for k, v in eng.timings().items()}, eng.solver_stats())
