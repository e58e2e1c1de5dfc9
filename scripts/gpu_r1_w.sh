#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tpp_paths.py tests/test_gpu_batched_parity.py -m gpu -q -x 2>&1 | tail -2
MIRB200_NO_STAGING=1 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
echo "ncu exit $?"; wc -l gpurun_out/launches_bench.csv
