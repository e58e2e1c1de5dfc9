#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.txt
tail -4 gpurun_out/pytest_gpu.txt | cut -c1-300
timeout 1200 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_n1.json'))
print("value",d["value"],"e2e",d["e2e"],"frac",d["roofline"]["frac"])
for k,v in d["secondary"].items(): print(k, v if not isinstance(v,dict) else {kk:vv for kk,vv in v.items() if kk in ("value","unit","ms","solve_s","iterations","roofline")})
PY
tail -3 gpurun_out/bench_n1.err
