n=4
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $n "$@"; }
timeout 200 bash -c "$(declare -f run); n=$n; run --no-cpu-baseline" > gpurun_out/r02g_weak_c4_4gpu.json 2> gpurun_out/r02g_weak_c4_4gpu.err
timeout 200 bash -c "$(declare -f run); n=$n; run --no-cpu-baseline --gather-table --no-e2e" > gpurun_out/r02g_weak_c4_4gpu_gt.json 2> gpurun_out/r02g_weak_c4_4gpu_gt.err
python tools/show_line.py gpurun_out/r02g_weak_c4_4gpu.json gpurun_out/r02g_weak_c4_4gpu_gt.json
