#!/bin/bash
# Round-1 GPU visit C: whole GPU suite, C4 solve timings, C4 launch list, full ncu capture of the SYRK + Broyden kernels.
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.txt
tail -15 gpurun_out/pytest_gpu.txt | cut -c1-300
timeout 600 python scripts/profile_c4.py --reps 3 > gpurun_out/profile_c4_solve.txt 2>&1; cat gpurun_out/profile_c4_solve.txt | cut -c1-700
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c4_r01.csv python scripts/profile_c4.py --reps 1 --max-iterations 10 > gpurun_out/c4_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:syrk_dmma -s 2 -c 1 -f -o gpurun_out/syrk_dmma_r01 python scripts/profile_c4.py --syrk-only --reps 2 > gpurun_out/ncu_full_syrk.log 2>&1
tail -2 gpurun_out/ncu_full_syrk.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:large_broyden -s 1 -c 1 -f -o gpurun_out/broyden_r01 python scripts/profile_c4.py --reps 1 --max-iterations 4 > gpurun_out/ncu_full_broyden.log 2>&1
tail -2 gpurun_out/ncu_full_broyden.log
ls -la gpurun_out
