#!/bin/bash
# split factor vs one segment at different phase counts (is the early start of the forward sweeps ever harmful?)
for P in 512 1024 2048 8192; do for e in "" JK_NO_FACTOR_SPLIT=1; do
  env $e timeout 200 python bench.py --phases $P --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/abp.json 2> gpurun_out/abp.err
  python - "$P" "$e" <<'PY'
import json, sys
try:
    d = json.loads(open("gpurun_out/abp.json").read().strip().splitlines()[-1]); s = d["stage_ms"]
    print(f"P={sys.argv[1]:5s} {sys.argv[2]:22s} {d['value']:10.0f} cases/s  step {d['ms_per_step']:.3f} ms  scan {s['scan_total']:.3f} | factor {s['factor']:.2f} morison {s['morison']:.2f} rhs {s['rhs']:.2f} fwd {s['solve_fwd']:.3f} bwd {s['solve_bwd']:.3f} post {s['post']:.2f}")
except Exception as e:
    print(sys.argv[1], sys.argv[2], "ERR", e, open("gpurun_out/abp.err").read()[-300:])
PY
done; done
