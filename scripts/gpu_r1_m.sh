#!/bin/bash
# Round-1 GPU visit M: group kernel with batched exps -- C3 timing (double, float), C2 small batch, parity.
mkdir -p gpurun_out; rm -f gpurun_out/m_c3.txt
timeout 300 python scripts/profile_c2.py --batch 65536 --config c3 2>&1 | tee -a gpurun_out/m_c3.txt
timeout 300 python scripts/profile_c2.py --batch 65536 --config c3 --dtype f32 2>&1 | tee -a gpurun_out/m_c3.txt
MIRB200_BATCH_KERNEL=group timeout 300 python scripts/profile_c2.py --batch 262144 2>&1 | tee -a gpurun_out/m_c3.txt
timeout 1200 python -m pytest tests/test_gpu_batched_parity.py tests/test_gpu_legacy_entry.py -m gpu -q -x > gpurun_out/pytest_gpu_m.txt 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_m.txt
tail -5 gpurun_out/pytest_gpu_m.txt | cut -c1-300
MIRB200_BATCH_KERNEL=group timeout 1200 python -m pytest tests/test_gpu_batched_parity.py -m gpu -q -x > gpurun_out/pytest_gpu_m2.txt 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_m2.txt
tail -5 gpurun_out/pytest_gpu_m2.txt | cut -c1-300
