#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lm_small -s 1 -c 1 -f -o gpurun_out/lm_small_c3_r01 python scripts/profile_c2.py --batch 32768 --config c3 --launches 2 > gpurun_out/ncu_full_c3.log 2>&1
tail -2 gpurun_out/ncu_full_c3.log
