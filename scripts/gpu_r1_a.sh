#!/bin/bash
# Round-1 GPU visit A: peaks, parity tests, smoke, per-config timings, bench, ncu launch list, full ncu captures.
set -x
mkdir -p gpurun_out; rm -f gpurun_out/parity_report.jsonl
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/host.txt; grep -m1 'model name' /proc/cpuinfo >> gpurun_out/host.txt
python - > gpurun_out/peaks.txt 2>&1 <<'PY'
import mir_optim_b200 as mo, json
L = mo.lib
out = {k: L.mir_b200_measure_peak_tflops(i, 5) for i, k in enumerate(("fp64_fma_tflops", "fp64_dmma_m8n8k4_tflops", "fp32_fma_tflops"))}
print(json.dumps(out))
PY
cat gpurun_out/peaks.txt
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.txt
tail -30 gpurun_out/pytest_gpu.txt | cut -c1-250
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.txt; tail -3 gpurun_out/smoke.txt
timeout 300 python scripts/profile_c2.py --batch 262144 > gpurun_out/profile_c2_plain.txt 2>&1; cat gpurun_out/profile_c2_plain.txt
timeout 300 python scripts/profile_c2.py --batch 65536 --config c3 > gpurun_out/profile_c3_plain.txt 2>&1; cat gpurun_out/profile_c3_plain.txt
timeout 300 python scripts/profile_c2.py --batch 65536 --config c3 --dtype f32 > gpurun_out/profile_c3_f32_plain.txt 2>&1; cat gpurun_out/profile_c3_f32_plain.txt
timeout 300 python scripts/profile_c5.py > gpurun_out/profile_c5.txt 2>&1; timeout 300 python scripts/profile_c5.py --dtype f32 >> gpurun_out/profile_c5.txt 2>&1; cat gpurun_out/profile_c5.txt
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 3000 gpurun_out/bench_n1.json; tail -5 gpurun_out/bench_n1.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 1500 gpurun_out/bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --batch 262144 > gpurun_out/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lm_small -s 1 -c 1 -f -o gpurun_out/lm_small_c2_r01 python scripts/profile_c2.py --batch 131072 --launches 2 > gpurun_out/ncu_full_c2.log 2>&1
tail -3 gpurun_out/ncu_full_c2.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:boxqp_cta -s 1 -c 1 -f -o gpurun_out/boxqp_c5_r01 python scripts/profile_c5.py --batch 20000 --launches 2 > gpurun_out/ncu_full_c5.log 2>&1
tail -3 gpurun_out/ncu_full_c5.log
ls -la gpurun_out
