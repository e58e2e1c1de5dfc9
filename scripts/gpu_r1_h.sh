#!/bin/bash
# Round-1 GPU visit H: thread-per-problem kernel with unrolled rows; occupancy variants; parity in both modes.
set -x
mkdir -p gpurun_out; rm -f gpurun_out/kernel_choice.txt
run() { echo "== $1 $2" >> gpurun_out/kernel_choice.txt
  MIR_B200_LIB=$2 MIRB200_BATCH_KERNEL=$1 timeout 300 python scripts/profile_c2.py --batch 262144 2>&1 | tail -3 >> gpurun_out/kernel_choice.txt
  MIR_B200_LIB=$2 MIRB200_BATCH_KERNEL=$1 timeout 300 python scripts/profile_c2.py --batch 65536 --config c3 2>&1 | tail -2 >> gpurun_out/kernel_choice.txt
  MIR_B200_LIB=$2 MIRB200_BATCH_KERNEL=$1 timeout 300 python scripts/profile_c2.py --batch 65536 --config c3 --dtype f32 2>&1 | tail -2 >> gpurun_out/kernel_choice.txt; }
run group ""
run thread ""
for v in 2 3 4; do run thread $PWD/variants/tpp_mb$v.so; done
cat gpurun_out/kernel_choice.txt
MIRB200_BATCH_KERNEL=thread timeout 1200 python -m pytest tests/test_gpu_batched_parity.py -m gpu -q > gpurun_out/pytest_gpu_tpp.txt 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_tpp.txt
tail -30 gpurun_out/pytest_gpu_tpp.txt | cut -c1-400
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.txt
tail -12 gpurun_out/pytest_gpu.txt | cut -c1-300
MIRB200_BATCH_KERNEL=thread timeout 900 ncu --set full --clock-control none --import-source on -k regex:lm_tpp -s 1 -c 1 -f -o gpurun_out/lm_tpp_c2_r01b python scripts/profile_c2.py --batch 131072 --launches 2 > gpurun_out/ncu_full_tpp.log 2>&1
