#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/parity_report.jsonl
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.txt
tail -40 gpurun_out/pytest_gpu.txt | cut -c1-250
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 300 python scripts/profile_c5.py 2>&1 | tee gpurun_out/profile_c5.txt
timeout 300 python scripts/profile_c5.py --dtype f32 2>&1 | tee -a gpurun_out/profile_c5.txt
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 1500 gpurun_out/bench_n1.json; tail -5 gpurun_out/bench_n1.err
