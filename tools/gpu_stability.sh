#!/bin/bash
# Repeated short bench runs (fresh process each) to expose run-to-run bimodality of the step time.
# usage: tools/gpu_stability.sh TAG N [ENV=VAL ...]
tag=$1; n=$2; shift; shift
for i in $(seq 1 $n); do
  env "$@" timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${tag}_$i.json 2> gpurun_out/${tag}_$i.err
  python - gpurun_out/${tag}_$i.json "$tag $i" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); s = d["stage_ms"]
    print(f"{sys.argv[2]:16s} {d['value']:9.0f} cases/s step {d['ms_per_step']:.3f} ms scan {s['scan_total']:.3f} | factor {s['factor']:.2f} morison {s['morison']:.2f} rhs {s['rhs']:.2f} fwd {s['solve_fwd']:.3f} bwd {s['solve_bwd']:.3f} post {s['post']:.2f}")
except Exception as e:
    print(sys.argv[2], "ERR", e)
PY
done
