#!/bin/bash
# run N: device parse parity + timing
mkdir -p gpurun_out
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2n_smoke.log 2>&1 || { echo "SMOKE FAILED"; tail -3 gpurun_out/r2n_smoke.log; exit 1; }
timeout 600 python -m pytest tests/test_gpu_parse.py tests/test_gpu_stroke.py -m gpu -x -q > gpurun_out/r2n_pytest.log 2>&1
echo "pytest rc=$?"; tail -25 gpurun_out/r2n_pytest.log | cut -c1-400
timeout 300 python tools/parse_time.py > gpurun_out/r2n_parse_time.txt 2>&1
timeout 300 python tools/stroke_time.py > gpurun_out/r2n_stroke_time.txt 2>&1; tail -6 gpurun_out/r2n_stroke_time.txt
tail -12 gpurun_out/r2n_parse_time.txt
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench.err; echo "bench rc=$?"
python -c "
import json
d=json.load(open('gpurun_out/r2n_bench.json'))
print(d['value'], d['ms_per_step'], d['roofline']['frac'])
print(json.dumps(d['other_configs'].get('pre_stages'), indent=1)[:3000])
"
