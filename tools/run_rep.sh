# repeat the scheduling-sensitive GPU tests (races show up as intermittent failures): tools/run_rep.sh N [lib.so|default ...]
n=$1; shift
for lib in "$@"; do
  for i in $(seq 1 $n); do
    if [ $lib = default ]; then unset JK_LIB; else export JK_LIB=$PWD/$lib; fi
    python -m pytest tests/test_gpu_parity.py -m gpu -q -k "split_factor or overlap_switches or two_chain or tma_sweep" 2>&1 | tail -1 | sed "s|^|$lib run $i: |"
  done
done
