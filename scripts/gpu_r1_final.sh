#!/bin/bash
# Round-1 final evidence on 1 GPU: full GPU tests, smoke, bench (own arm + reference arm).
mkdir -p gpurun_out; rm -f gpurun_out/parity_report.jsonl
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.txt
tail -3 gpurun_out/pytest_gpu.txt | cut -c1-200
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_n1.json'))
print("value",d["value"],"e2e",d["e2e"]["value"],d["e2e"]["ms_per_step"],"frac",d["roofline"]["frac"],"cpu",d["cpu_baseline"]["value"], "launches", d["gpu_launches"], d["clocks"])
for k,v in d["secondary"].items(): print(k, v if not isinstance(v,dict) else {kk:vv for kk,vv in v.items() if kk in ("value","unit","ms","solve_s","iterations","passes")})
PY
