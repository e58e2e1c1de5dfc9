#!/bin/bash
# Round-1 GPU visit J: full ncu capture of the thread-per-problem kernel (C2) with source attribution.
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lm_tpp -s 1 -c 1 -f -o gpurun_out/lm_tpp_c2_r01d python scripts/profile_c2.py --batch 262144 --launches 2 > gpurun_out/ncu_full_tpp.log 2>&1
tail -3 gpurun_out/ncu_full_tpp.log
ls -la gpurun_out/
