#!/bin/bash
N=${1:-4}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "exit $?"; python - <<PY
import json
d=json.load(open('gpurun_out/bench_n$N.json'))
print("n_gpus",d["n_gpus"],"value",d["value"],"e2e",d["e2e"]["value"], d["e2e"]["ms_per_step"])
for k,v in d["secondary"].items(): print(k, v if not isinstance(v,dict) else {kk:vv for kk,vv in v.items() if kk in ("value","unit","ms","solve_s","iterations","passes","status")})
PY
tail -2 gpurun_out/bench_n$N.err
