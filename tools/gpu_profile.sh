#!/bin/bash
# Final-build evidence in one GPU-box visit: full bench line, ncu launch list, one --set full capture of a whole step.
# usage: tools/gpu_profile.sh TAG
tag=$1
python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${tag}_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${tag}_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_sweep|k_morison_airy|k_member_post|k_rhs_gather|k_band_chol_cluster|k_tile_inverse" -s 57 -c 19 \
    -f -o gpurun_out/${tag}_prof python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${tag}_ncu2.log 2>&1
ncu -i gpurun_out/${tag}_prof.ncu-rep --page raw --csv > gpurun_out/${tag}_raw.csv 2>/dev/null
python - $tag <<'PY'
import json, sys
d = json.loads(open(f"gpurun_out/{sys.argv[1]}_bench.json").read().strip().splitlines()[-1])
print(sys.argv[1], round(d["value"]), round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"]), "cpu", d["cpu_baseline"]["value"], d["roofline"]["frac"], d["clocks"])
PY
