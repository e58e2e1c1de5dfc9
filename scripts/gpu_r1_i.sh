#!/bin/bash
# Round-1 GPU visit I: tpp with observations in smem; batch-size crossover; parity in both modes; bench.
set -x
mkdir -p gpurun_out; rm -f gpurun_out/kernel_choice.txt
for k in group thread; do for b in 2048 8192 16384 65536 262144 1048576; do
  echo "== $k batch $b" >> gpurun_out/kernel_choice.txt
  MIRB200_BATCH_KERNEL=$k timeout 300 python scripts/profile_c2.py --batch $b --launches 3 2>&1 | tail -2 | head -1 >> gpurun_out/kernel_choice.txt
done; done
cat gpurun_out/kernel_choice.txt
MIRB200_BATCH_KERNEL=thread timeout 1200 python -m pytest tests/test_gpu_batched_parity.py -m gpu -q > gpurun_out/pytest_gpu_tpp.txt 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_tpp.txt
tail -8 gpurun_out/pytest_gpu_tpp.txt | cut -c1-400
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.txt
tail -8 gpurun_out/pytest_gpu.txt | cut -c1-300
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 2600 gpurun_out/bench_n1.json; tail -5 gpurun_out/bench_n1.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lm_tpp -s 1 -c 1 -f -o gpurun_out/lm_tpp_c2_r01c python scripts/profile_c2.py --batch 1048576 --launches 2 > gpurun_out/ncu_full_tpp.log 2>&1
