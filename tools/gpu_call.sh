#!/bin/bash
# One GPU-box visit of round 2: GPU tests, bench lines of every named configuration, optional ncu evidence.
# usage: tools/gpu_call.sh TAG [tests] [benches] [ncu]
tag=$1; shift
mkdir -p gpurun_out
for what in "$@"; do
case $what in
tests)
  python -m pytest tests -m gpu -x -q -s 2>&1 | tail -25 > gpurun_out/${tag}_tests.log; tail -8 gpurun_out/${tag}_tests.log ;;
smoke)
  python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 ;;
benches)
  for w in c4_jacket10k c3_jacket2k c2_default3 c1_default3 c5_ensemble; do
    python bench.py --workload $w > gpurun_out/${tag}_bench_$w.json 2> gpurun_out/${tag}_bench_$w.err
    python - $tag $w <<'PY'
import json, sys
tag, w = sys.argv[1:3]
try:
    d = json.loads(open(f"gpurun_out/{tag}_bench_{w}.json").read().strip().splitlines()[-1])
    par = d.get("parity") or {}
    print(w, round(d["value"]), "cases/s", round(d["ms_per_step"], 3), "ms | e2e", round(d["e2e"]["value"]), "| cpu", d["cpu_baseline"] and d["cpu_baseline"]["value"],
          "| roofline", d["roofline"]["kernel"], round(d["roofline"]["frac"], 3), "| parity", par.get("max_rel"), par.get("max_rel_vs_plain_lu"), par.get("critical_index_match"),
          "|", {k: round(v, 3) for k, v in d["stage_ms"].items()})
except Exception as e:
    print(w, "ERR", e); print(open(f"gpurun_out/{tag}_bench_{w}.err").read()[-1500:])
PY
  done ;;
c4)
  python bench.py > gpurun_out/${tag}_bench_c4_jacket10k.json 2> gpurun_out/${tag}_bench_c4_jacket10k.err
  python - $tag <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/{tag}_bench_c4_jacket10k.json").read().strip().splitlines()[-1])
    par = d.get("parity") or {}
    print("c4", round(d["value"]), "cases/s", round(d["ms_per_step"], 3), "ms | e2e", round(d["e2e"]["value"]), "| roofline", d["roofline"]["kernel"], round(d["roofline"]["frac"], 3),
          "| parity", par.get("max_rel"), par.get("max_rel_vs_plain_lu"), par.get("critical_index_match"), "|", {k: round(v, 3) for k, v in d["stage_ms"].items()})
except Exception as e:
    print("c4 ERR", e); print(open(f"gpurun_out/{tag}_bench_c4_jacket10k.err").read()[-1500:])
PY
  ;;
ncu)
  cmd="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e"
  $cmd > gpurun_out/${tag}_plain.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/${tag}_launches.csv $cmd > gpurun_out/${tag}_ncu1.log 2>&1
  $cmd > gpurun_out/${tag}_plain2.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:"^k_sweep$|k_morison_airy|k_member_post|k_rhs_gather|k_band_chol_cluster|k_tile_inverse" -s ${NCU_SKIP:-60} -c ${NCU_COUNT:-15} \
      -f -o gpurun_out/${tag}_prof $cmd > gpurun_out/${tag}_ncu2.log 2>&1
  ncu -i gpurun_out/${tag}_prof.ncu-rep --page raw --csv > gpurun_out/${tag}_raw.csv 2>/dev/null
  tail -3 gpurun_out/${tag}_ncu2.log ;;
esac
done
