#!/bin/bash
# A/B runs of experimental builds of the library (JK_LIB) on the bench workload: prints one summary line per variant.
# usage: tools/ab_bench.sh tag variant1.so variant2.so ...   ("default" = the in-tree library)
tag=$1; shift
for v in "$@"; do
  name=$(basename $v .so)
  if [ "$v" = "default" ]; then unset JK_LIB; else export JK_LIB=$PWD/$v; fi
  timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/${tag}_${name}.json 2> gpurun_out/${tag}_${name}.err
  python - "$name" gpurun_out/${tag}_${name}.json <<'PY'
import json, sys
name, f = sys.argv[1], sys.argv[2]
try:
    d = json.loads(open(f).read().strip().splitlines()[-1]); s = d["stage_ms"]
    print(f"{name:14s} {d['value']:10.0f} cases/s  step {d['ms_per_step']:.3f} ms | factor {s['factor']:.2f} morison {s['morison']:.2f} rhs {s['rhs']:.2f} fwd {s['solve_fwd']:.3f} bwd {s['solve_bwd']:.3f} post {s['post']:.2f} | res {d['rel_residual']:.1e}")
except Exception as e:
    print(name, "ERR", e)
PY
done
