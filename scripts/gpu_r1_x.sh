#!/bin/bash
mkdir -p gpurun_out
timeout 200 python scripts/profile_c2.py --batch 1048576 --launches 3 2>&1 | tail -4 | head -3
MIRB200_TPP_L2_PERSIST=1 timeout 200 python scripts/profile_c2.py --batch 1048576 --launches 3 2>&1 | tail -5 | head -4
MIRB200_TPP_L2_PERSIST=1 timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:lm_tpp -s 1 -c 1 python scripts/profile_c2.py --batch 1048576 --launches 2 2>&1 | grep -E "dram__|duration"
