#!/bin/bash
# Code-generation sweep of the thread-per-problem kernel.
#   scripts/sweep_tpp.sh build   -- here (no GPU): builds variants/<name>.so for every entry of VARIANTS
#   scripts/sweep_tpp.sh run     -- on the GPU box (via gpurun): times each variant on configs[1] and runs the tpp parity tests
# Add a line to VARIANTS = "<name>|<extra nvcc flags>".
VARIANTS=(
  "base|"
  "coop|-DMIRB200_TPP_COOP_REFILL=1"
  "mb2|-DMIRB200_TPP_MINBLOCKS=2"
)
cd "$(dirname "$0")/.."
case "$1" in
  build)
    for v in "${VARIANTS[@]}"; do name=${v%%|*}; flags=${v#*|}; echo "== building $name ($flags)"; bash scripts/build_variant.sh "$name" "$flags" | tail -2; done ;;
  run)
    mkdir -p gpurun_out; : > gpurun_out/sweep_tpp.txt
    for v in "${VARIANTS[@]}"; do name=${v%%|*}
      echo "== $name" | tee -a gpurun_out/sweep_tpp.txt
      MIR_B200_LIB=$PWD/variants/$name.so timeout 200 python scripts/profile_c2.py --batch 1048576 --launches 4 2>&1 | tail -4 | head -3 | tee -a gpurun_out/sweep_tpp.txt
      MIR_B200_LIB=$PWD/variants/$name.so timeout 600 python -m pytest tests/test_gpu_tpp_paths.py -m gpu -q -x 2>&1 | tail -1 | tee -a gpurun_out/sweep_tpp.txt
    done ;;
  *) echo "usage: $0 build|run"; exit 2 ;;
esac
