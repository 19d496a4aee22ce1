#!/bin/bash
# run AP: whole GPU suite + the default bench (both arms) with the split download / prepared-ahead chunks
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2ap_pytest.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/r2ap_pytest.log
timeout 600 python bench.py > gpurun_out/r2ap_bench_n1.json 2> gpurun_out/r2ap_bench_n1.err; echo bench rc=$?
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2ap_bench_ref_n1.json 2> gpurun_out/r2ap_bench_ref_n1.err; echo ref rc=$?
python - <<PY
import json
d=json.loads(open("gpurun_out/r2ap_bench_n1.json").read().strip().splitlines()[-1])
print("c4", d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"])
for k,v in d.get("other_configs",{}).items(): print(k, v.get("value"), v.get("ms_per_step"), v.get("roofline",{}).get("frac"), v.get("e2e",{}).get("ms_per_call"))
r=json.loads(open("gpurun_out/r2ap_bench_ref_n1.json").read().strip().splitlines()[-1])
print("ref", r["value"], r.get("cpu_baseline"))
PY
