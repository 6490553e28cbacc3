"""Times assemble + factor of the c4 jacket a few times (debug aid; JK_CHOL_PROFILE=1 prints phase clocks)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import jacket_b200 as jb
legs, bays = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (16, 104)
p = jb.AnalysisParams(wave_model="Airy")
nodes, members, fixed, top = jb.generate_jacket(legs, bays)
st = jb.build_structure(nodes, members, fixed, top, p)
eng = jb.get_engine(st)
eng.set_supports(st.indices(fixed))
for i in range(3):
    eng.assemble(p.E, p.E / 2.6)
    eng.factor()
    print(eng.dims(), {k: round(v, 3) for k, v in eng.timings().items() if k in ("assemble", "factor")})
