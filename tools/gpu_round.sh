#!/bin/bash
# One GPU-box visit: GPU tests, A/B of the listed library variants, the bench line (and the N-GPU line when N is given).
# usage: tools/gpu_round.sh TAG [NGPU] [variant.so ...]
tag=$1; n=${2:-1}; shift; shift
python -m pytest tests -m gpu -x -q 2>&1 | tail -6
[ $# -gt 0 ] && bash tools/ab_bench.sh $tag default "$@"
python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
files="${tag}_bench"
if [ "$n" -gt 1 ]; then
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/${tag}_scale_$n.json 2> gpurun_out/${tag}_scale_$n.err
  files="$files ${tag}_scale_$n"
fi
for f in $files; do python - $f <<'PY'
import json, sys
f = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
    print(f, d["n_gpus"], round(d["value"]), round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"]), {k: round(v, 3) for k, v in d["stage_ms"].items()})
except Exception as e:
    print(f, "ERR", e); print(open(f"gpurun_out/{f}.err").read()[-1500:])
PY
done
