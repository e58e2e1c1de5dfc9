#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_boxqp.py tests/test_gpu_large.py tests/test_gpu_legacy_entry.py -m gpu -q -x 2>&1 | tail -3
timeout 300 python scripts/profile_c4.py --reps 3 2>&1 | cut -c1-200
