#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --warp-sampling-interval 0 --clock-control none --import-source on -k regex:large_ctl_mid -s 20 -c 1 -f -o gpurun_out/ctl_mid_r01 python scripts/profile_c4.py --m 262144 --reps 1 > gpurun_out/ncu_ctl.log 2>&1
tail -2 gpurun_out/ncu_ctl.log
