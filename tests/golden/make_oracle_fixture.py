"""c4_oracle_scan4096.npz: the 8 Morison columns of the FULL 4,096-phase scan of the c4 jacket (16 legs x 104 bays, 10,000
members, GUI defaults) computed by the ORACLE (oracle/jacket_oracle.py, about 5 minutes on 8 cores) -- not by the reference,
which would need ~20 CPU-hours for it.  The reference itself pins 76 rows of this table (tests/golden/gen16x104_c4.npz,
rows 0, 64, ..., 4032, 4095 and the five around the critical phase; tests/test_oracle_golden.py compares them), and the GPU
tests use the fixture for the rows in between and for the critical index.

    python tests/golden/make_oracle_fixture.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)


def main():
    import jacket_b200 as jb                  # host code only: synthetic geometry and the GUI defaults
    from oracle import jacket_oracle as orc
    p = jb.AnalysisParams(wave_model="Airy")
    nodes, members, fixed, top = jb.generate_jacket(16, 104)
    st = jb.build_structure(nodes, members, fixed, top, p)
    xyz, conn, sec_id, _, sections = st.pack()
    model = orc.Model(xyz, conn, sec_id, [(s.D_outer, s.t, s.rho_steel) for s in sections], st.indices(fixed), st.indices(top))
    wave = orc.AiryWave(p.H, p.T, p.d, p.U_c)
    t = orc.phase_times(p.T, 4096)
    parts = []
    for lo in range(0, 4096, 128):
        out = orc.morison_phases(model, wave, t[lo:lo + 128], wave_direction=p.wave_dir, current_direction=p.current_dir,
                                 Cd=p.Cd, Cm=p.Cm, rho_water=p.rho_water)
        parts.append(orc.phase_table(out, t[lo:lo + 128], wave.omega)[0])
    table = np.concatenate(parts)
    crit = int(np.argmax(table[:, 2]))
    np.savez_compressed(os.path.join(HERE, "c4_oracle_scan4096.npz"), table=table, critical=np.array(crit))
    print("critical", crit, table[crit, 2])


if __name__ == "__main__":
    main()
