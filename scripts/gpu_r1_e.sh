#!/bin/bash
set -x
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1; cat gpurun_out/topo.txt
NCCL_DEBUG=INFO timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 scripts/nccl_latency.py > gpurun_out/nccl_latency.txt 2>&1
grep -E "allreduce|can access|NVLS|P2P|SHM|via|Channel 00|transport|NET" gpurun_out/nccl_latency.txt | head -40 | cut -c1-220
timeout 900 python -m pytest tests/test_gpu_sharded.py -q -x > gpurun_out/pytest_sharded.txt 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_sharded.txt
tail -5 gpurun_out/pytest_sharded.txt | cut -c1-400
