#!/bin/bash
# A/B of library builds on the sea-state ensemble workload (c5): tools/ab_c5.sh tag variant.so|default ...
tag=$1; shift
for v in "$@"; do
  name=$(basename $v .so)
  if [ "$v" = "default" ]; then unset JK_LIB; else export JK_LIB=$PWD/$v; fi
  timeout 200 python bench.py --workload c5_ensemble --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_c5_${name}.json 2> gpurun_out/${tag}_c5_${name}.err
  echo -n "$name: "; python tools/show_line.py gpurun_out/${tag}_c5_${name}.json | tail -1
done
