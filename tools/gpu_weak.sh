#!/bin/bash
# weak-scaling line only: tools/gpu_weak.sh TAG N
tag=$1; n=$2
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $n --no-cpu-baseline > gpurun_out/${tag}_weak_c4_${n}gpu.json 2> gpurun_out/${tag}_weak_c4_${n}gpu.err
python tools/show_line.py gpurun_out/${tag}_weak_c4_${n}gpu.json
