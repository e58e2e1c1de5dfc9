#!/bin/bash
# 2-GPU: sharded test + bench through torchrun, exactly as the driver launches it
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sharded.py -m gpu -q 2>&1 | tail -3
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
echo "exit $?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_n2.json'))
print("value",d["value"],"e2e",d["e2e"]["value"], d["e2e"]["ms_per_step"])
for k,v in d["secondary"].items(): print(k, v if not isinstance(v,dict) else {kk:vv for kk,vv in v.items() if kk in ("value","unit","ms","solve_s","iterations","passes","status")})
PY
tail -3 gpurun_out/bench_n2.err
