#!/bin/bash
# Round-1 GPU visit O: bench with secondary configs + pipelined e2e; launch list; full capture at the bench batch size.
mkdir -p gpurun_out
timeout 1200 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 6000 gpurun_out/bench_n1.json; tail -5 gpurun_out/bench_n1.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lm_tpp -s 1 -c 1 -f -o gpurun_out/lm_tpp_c2_1M_r01 python scripts/profile_c2.py --batch 1048576 --launches 2 > gpurun_out/ncu_full_tpp1M.log 2>&1
tail -2 gpurun_out/ncu_full_tpp1M.log
