#!/bin/bash
# Round-1 GPU visit D (2 GPUs): sharded C4 parity + timing, batched bench at N=2.
set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_sharded.py -q -x > gpurun_out/pytest_sharded.txt 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_sharded.txt
tail -15 gpurun_out/pytest_sharded.txt | cut -c1-400
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 scripts/c4_sharded.py --reps 3 2>gpurun_out/c4_2gpu.err | cut -c1-420 > gpurun_out/c4_2gpu.jsonl; cat gpurun_out/c4_2gpu.jsonl; tail -3 gpurun_out/c4_2gpu.err
timeout 600 python scripts/c4_sharded.py --reps 3 2>gpurun_out/c4_1gpu.err | cut -c1-420 > gpurun_out/c4_1gpu.jsonl; cat gpurun_out/c4_1gpu.jsonl
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; tail -c 2500 gpurun_out/bench_n2.json; tail -3 gpurun_out/bench_n2.err
