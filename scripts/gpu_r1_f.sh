#!/bin/bash
set -x
mkdir -p gpurun_out
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $R --master-port 29621 scripts/c4_sharded.py --reps 4 --no-comm 2>/dev/null | cut -c1-330 > gpurun_out/c4_2gpu_nocomm.jsonl; cat gpurun_out/c4_2gpu_nocomm.jsonl
timeout 300 $R --master-port 29622 scripts/c4_sharded.py --reps 4 2>/dev/null | cut -c1-330 > gpurun_out/c4_2gpu.jsonl; cat gpurun_out/c4_2gpu.jsonl
timeout 300 $R --master-port 29623 scripts/c4_sharded.py --reps 4 --own-stream 2>/dev/null | cut -c1-330 > gpurun_out/c4_2gpu_ownstream.jsonl; cat gpurun_out/c4_2gpu_ownstream.jsonl
NCCL_PROTO=LL NCCL_ALGO=Ring timeout 300 $R --master-port 29624 scripts/c4_sharded.py --reps 4 --own-stream 2>/dev/null | cut -c1-330 > gpurun_out/c4_2gpu_ll.jsonl; cat gpurun_out/c4_2gpu_ll.jsonl
