#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_sharded.py -m gpu -q 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/c4_sharded.py --reps 3 2>/dev/null | grep "^{" | cut -c1-300
