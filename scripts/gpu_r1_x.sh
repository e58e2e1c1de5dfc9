#!/bin/bash
timeout 200 python scripts/profile_c2.py --batch 1048576 --launches 4 2>&1 | tail -4 | head -3
timeout 900 python -m pytest tests/test_gpu_tpp_paths.py -m gpu -q -x 2>&1 | tail -2
