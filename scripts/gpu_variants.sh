#!/bin/bash
mkdir -p gpurun_out
python scripts/diag_parity.py > gpurun_out/diag_parity.txt 2>&1; grep -v "    prob" gpurun_out/diag_parity.txt | cut -c1-250
echo "== default" | tee gpurun_out/variants.txt
timeout 200 python scripts/profile_c2.py --batch 262144 2>&1 | tee -a gpurun_out/variants.txt
for v in r168_l32 r128_l32 r96_l32 r255_l16 r128_l16 r255_l8 r128_l8; do
  echo "== $v" | tee -a gpurun_out/variants.txt
  MIR_B200_LIB=$PWD/variants/$v.so timeout 200 python scripts/profile_c2.py --batch 262144 2>&1 | tee -a gpurun_out/variants.txt
done
timeout 300 python scripts/profile_c2.py --batch 65536 --config c3 2>&1 | tee -a gpurun_out/variants.txt
timeout 300 python scripts/profile_c2.py --batch 65536 --config c3 --dtype f32 2>&1 | tee -a gpurun_out/variants.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lm_small -s 1 -c 1 -f -o gpurun_out/lm_small_c2_r01b python scripts/profile_c2.py --batch 131072 --launches 2 > gpurun_out/ncu_full.log 2>&1
