#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_large.py tests/test_gpu_legacy_entry.py -m gpu -q 2>&1 | tail -2
timeout 300 python scripts/profile_c4.py --reps 4 2>&1 | cut -c1-150
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_c4_fused.csv python scripts/profile_c4.py --reps 1 > /dev/null 2>&1
