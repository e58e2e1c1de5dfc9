#!/bin/bash
timeout 200 python scripts/profile_c2.py --batch 1048576 --launches 4 2>&1 | tail -4 | head -3
timeout 200 python scripts/profile_c2.py --batch 65536 --config c3 2>&1 | tail -2 | head -1
timeout 200 python scripts/profile_c2.py --batch 65536 --config c3 --dtype f32 2>&1 | tail -2 | head -1
timeout 900 python -m pytest tests/test_gpu_tpp_paths.py tests/test_gpu_batched_parity.py -m gpu -q -x 2>&1 | tail -2
