#!/bin/bash
# run O: full GPU suite + default bench with the stroke / parse additions
mkdir -p gpurun_out
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2o_smoke.log 2>&1 || { echo "SMOKE FAILED"; tail -3 gpurun_out/r2o_smoke.log; exit 1; }
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2o_pytest.log 2>&1
echo "pytest rc=$?"; tail -8 gpurun_out/r2o_pytest.log | cut -c1-300
timeout 600 python bench.py > gpurun_out/r2o_bench.json 2> gpurun_out/r2o_bench.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2o_bench_ref.json 2> gpurun_out/r2o_bench_ref.err; echo "ref rc=$?"
python -c "
import json
d=json.load(open('gpurun_out/r2o_bench.json'))
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e'])
for k,v in d['other_configs'].items():
    print(k, {kk: v[kk] for kk in ('value','ms_per_step') if kk in v} or list(v.keys()))
r=json.load(open('gpurun_out/r2o_bench_ref.json')); print('ref', r['value'], r['unit'])
"
