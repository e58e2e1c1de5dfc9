#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lm_tpp -s 1 -c 1 -f -o gpurun_out/lm_tpp_c2_1M_final python scripts/profile_c2.py --batch 1048576 --launches 2 > gpurun_out/ncu_full_tpp1M.log 2>&1
tail -1 gpurun_out/ncu_full_tpp1M.log
