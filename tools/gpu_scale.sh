#!/bin/bash
# Scaling lines on N GPUs of one box: weak c4 (the driver's line), strong c4 (configs[3]), strong c5 (configs[4]).
# usage: tools/gpu_scale.sh TAG N [tests]
tag=$1; n=$2
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $n "$@"; }
[ "$3" = tests ] && timeout 300 python -m pytest tests/test_gpu_multirank.py -m gpu -x -q 2>&1 | tail -3
timeout 300 bash -c "$(declare -f run); n=$n; run --no-cpu-baseline" > gpurun_out/${tag}_weak_c4_${n}gpu.json 2> gpurun_out/${tag}_weak_c4_${n}gpu.err
timeout 300 bash -c "$(declare -f run); n=$n; run --no-cpu-baseline --scaling strong" > gpurun_out/${tag}_strong_c4_${n}gpu.json 2> gpurun_out/${tag}_strong_c4_${n}gpu.err
timeout 300 bash -c "$(declare -f run); n=$n; run --no-cpu-baseline --scaling strong --workload c5_ensemble" > gpurun_out/${tag}_strong_c5_${n}gpu.json 2> gpurun_out/${tag}_strong_c5_${n}gpu.err
python tools/show_line.py gpurun_out/${tag}_weak_c4_${n}gpu.json gpurun_out/${tag}_strong_c4_${n}gpu.json gpurun_out/${tag}_strong_c5_${n}gpu.json
