#!/bin/bash
MIR_B200_LIB=$PWD/variants/coop.so timeout 100 python scripts/profile_c2.py --batch 1048576 --launches 3 2>&1 | tail -3 | head -2
MIR_B200_LIB=$PWD/variants/coop.so timeout 200 python -m pytest tests/test_gpu_tpp_paths.py -m gpu -q -x 2>&1 | tail -1
