#!/bin/bash
# run BD: whole GPU suite, smoke, the default bench (both arms) and its launch list with the final binary
mkdir -p gpurun_out
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2bd_smoke.log 2>&1; echo smoke rc=$?; tail -1 gpurun_out/r2bd_smoke.log
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2bd_pytest.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/r2bd_pytest.log
timeout 600 python bench.py --impl reference > gpurun_out/r2bd_bench_ref_n1.json 2> gpurun_out/r2bd_bench_ref_n1.err; echo ref rc=$?
timeout 600 python bench.py > gpurun_out/r2bd_bench_n1.json 2> gpurun_out/r2bd_bench_n1.err; echo bench rc=$?
python - <<PY
import json
d=json.loads(open("gpurun_out/r2bd_bench_n1.json").read().strip().splitlines()[-1])
print("c4", d["value"], d["ms_per_step"], d["roofline"]["frac"], {k:v for k,v in d["e2e"].items() if k!="call"})
for k,v in d.get("other_configs",{}).items(): print(k, v.get("value"), v.get("ms_per_step"), v.get("roofline",{}).get("frac"), {a:b for a,b in (v.get("e2e") or {}).items() if a!="call"})
print(d.get("pre_stages"))
r=json.loads(open("gpurun_out/r2bd_bench_ref_n1.json").read().strip().splitlines()[-1])
print("ref", r["value"], r.get("ms_per_step"))
PY
