#!/bin/bash
# Round-1 GPU visit K: speculative single-row-loop tpp kernel -- parity + timing (v-list vs stored J).
mkdir -p gpurun_out
timeout 200 python scripts/profile_c2.py --batch 1048576 --launches 3 2>&1 | tee gpurun_out/k_default.txt
MIRB200_TPP_STORED_J=1 timeout 200 python scripts/profile_c2.py --batch 1048576 --launches 3 2>&1 | tee gpurun_out/k_storedj.txt
timeout 1200 python -m pytest tests/test_gpu_batched_parity.py tests/test_gpu_legacy_entry.py -m gpu -q -x > gpurun_out/pytest_gpu_k.txt 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_k.txt
tail -15 gpurun_out/pytest_gpu_k.txt | cut -c1-300
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lm_tpp -s 1 -c 1 -f -o gpurun_out/lm_tpp_c2_r01e python scripts/profile_c2.py --batch 262144 --launches 2 > gpurun_out/ncu_full_tpp.log 2>&1
