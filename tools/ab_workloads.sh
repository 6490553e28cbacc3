for w in c3_jacket2k c5_ensemble; do for e in "" JK_NO_FACTOR_SPLIT=1; do
  env $e timeout 200 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ab2.json 2> gpurun_out/ab2.err
  python - "$w" "$e" <<'PY'
import json, sys
try:
    d = json.loads(open("gpurun_out/ab2.json").read().strip().splitlines()[-1]); s = d["stage_ms"]
    print(f"{sys.argv[1]:14s} {sys.argv[2]:22s} {d['value']:10.0f} cases/s  step {d['ms_per_step']:.3f} ms | factor {s['factor']:.2f} morison {s['morison']:.2f} rhs {s['rhs']:.2f} fwd {s['solve_fwd']:.3f} bwd {s['solve_bwd']:.3f} post {s['post']:.2f}")
except Exception as e:
    print(sys.argv[1], sys.argv[2], "ERR", e, open("gpurun_out/ab2.err").read()[-300:])
PY
done; done
