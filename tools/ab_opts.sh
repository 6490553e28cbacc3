#!/bin/bash
# A/B of run-time options on the bench workload: tools/ab_opts.sh tag "k=v,k=v" "k=v" ...   ("-" = defaults)
tag=$1; shift
i=0
for o in "$@"; do
  i=$((i+1)); args=""
  if [ "$o" != "-" ]; then for kv in ${o//,/ }; do args="$args --option $kv"; done; fi
  f=gpurun_out/${tag}_opt$i.json
  timeout 150 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e $args > $f 2> ${f%.json}.err
  python - "$o" $f <<'PY'
import json, sys
name, f = sys.argv[1], sys.argv[2]
try:
    d = json.loads(open(f).read().strip().splitlines()[-1]); s = d["stage_ms"]
    print(f"{name:34s} {d['value']:10.0f} cases/s  step {d['ms_per_step']:.3f} ms | factor {s['factor']:.2f} morison {s['morison']:.2f} rhs {s['rhs']:.2f} fwd {s['solve_fwd']:.3f} bwd {s['solve_bwd']:.3f} post {s['post']:.2f}")
except Exception as e:
    print(name, "ERR", e, open(f.replace('.json', '.err')).read()[-400:])
PY
done
