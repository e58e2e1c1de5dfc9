#!/bin/bash
# Round-1 GPU visit L: occupancy variants of the speculative tpp kernel.
mkdir -p gpurun_out; rm -f gpurun_out/l_variants.txt
run() { echo "== $1 yos=$2" | tee -a gpurun_out/l_variants.txt; MIR_B200_LIB=$3 MIRB200_TPP_YOS=$2 timeout 200 python scripts/profile_c2.py --batch 1048576 --launches 3 2>&1 | tail -3 | head -2 | tee -a gpurun_out/l_variants.txt; }
run default 1 ""
run default 0 ""
run mb3 0 $PWD/variants/mb3.so
run mb4 0 $PWD/variants/mb4.so
run mb3 1 $PWD/variants/mb3.so
