#!/bin/bash
# Round-1 GPU visit G: tail fast-forward + thread-per-problem kernel: parity under both batched kernels, timings, ncu.
set -x
mkdir -p gpurun_out; rm -f gpurun_out/parity_report.jsonl gpurun_out/kernel_choice.txt
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.txt
tail -12 gpurun_out/pytest_gpu.txt | cut -c1-300
MIRB200_BATCH_KERNEL=thread timeout 1200 python -m pytest tests/test_gpu_batched_parity.py -m gpu -q > gpurun_out/pytest_gpu_tpp.txt 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_tpp.txt
tail -30 gpurun_out/pytest_gpu_tpp.txt | cut -c1-300
for k in group thread; do
  echo "== $k" >> gpurun_out/kernel_choice.txt
  MIRB200_BATCH_KERNEL=$k timeout 300 python scripts/profile_c2.py --batch 262144 >> gpurun_out/kernel_choice.txt 2>&1
  MIRB200_BATCH_KERNEL=$k timeout 300 python scripts/profile_c2.py --batch 65536 --config c3 >> gpurun_out/kernel_choice.txt 2>&1
  MIRB200_BATCH_KERNEL=$k timeout 300 python scripts/profile_c2.py --batch 65536 --config c3 --dtype f32 >> gpurun_out/kernel_choice.txt 2>&1
done
cat gpurun_out/kernel_choice.txt
MIRB200_BATCH_KERNEL=thread timeout 900 ncu --set full --clock-control none --import-source on -k regex:lm_tpp -s 1 -c 1 -f -o gpurun_out/lm_tpp_c2_r01 python scripts/profile_c2.py --batch 131072 --launches 2 > gpurun_out/ncu_full_tpp.log 2>&1
tail -2 gpurun_out/ncu_full_tpp.log
timeout 600 python scripts/profile_c4.py --reps 3 > gpurun_out/profile_c4_solve.txt 2>&1; cut -c1-330 gpurun_out/profile_c4_solve.txt
