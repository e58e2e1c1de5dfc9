#!/bin/bash
# Round-1 GPU visit B: large-problem path (legacy entry, SYRK DMMA, C1/C4) tests + timings.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_large.py tests/test_gpu_legacy_entry.py -q -x > gpurun_out/pytest_large.txt 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_large.txt
tail -40 gpurun_out/pytest_large.txt | cut -c1-300
timeout 300 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_large.py -q -x -k "syrk and 32-8-8" > gpurun_out/sanitizer_syrk.txt 2>&1; tail -5 gpurun_out/sanitizer_syrk.txt
timeout 300 python scripts/profile_c4.py --syrk-only > gpurun_out/profile_c4_syrk.txt 2>&1; cat gpurun_out/profile_c4_syrk.txt
timeout 600 python scripts/profile_c4.py --reps 2 > gpurun_out/profile_c4_solve.txt 2>&1; cat gpurun_out/profile_c4_solve.txt | cut -c1-600
