#!/bin/bash
# run Y: full suite, default bench, c2 at other cut depths, launch lists of the stroke / parse calls
mkdir -p gpurun_out
timeout 180 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2y_smoke.log 2>&1 || { echo "SMOKE FAILED"; tail -5 gpurun_out/r2y_smoke.log; exit 1; }
tail -1 gpurun_out/r2y_smoke.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2y_pytest.log 2>&1
echo "pytest rc=$?"; tail -6 gpurun_out/r2y_pytest.log | cut -c1-300
for d in 3 4 5; do
RGPU_CUT_DEPTH=$d timeout 200 python bench.py --workload c2 --no-others --steps 50 --warmup 5 > gpurun_out/r2y_c2_d$d.json 2> gpurun_out/r2y_c2_d$d.err
python -c "
import json
d=json.load(open('gpurun_out/r2y_c2_d$d.json'))
print('c2 depth $d', d['ms_per_step'], d['roofline']['stage_ms'])
"
done
timeout 600 python bench.py > gpurun_out/r2y_bench.json 2> gpurun_out/r2y_bench.err; echo "bench rc=$?"
python -c "
import json
d=json.load(open('gpurun_out/r2y_bench.json'))
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['ms_per_call'])
for k,v in d['other_configs'].items():
    print(k, {kk: v[kk] for kk in ('value','ms_per_step') if kk in v} or {a: (b.get('ms_per_call'), b.get('kernel_ms')) for a,b in v.items()})
"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2y_launches_stroke.csv python tools/stroke_time.py > gpurun_out/r2y_ncu_stroke.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r2y_launches_parse.csv python tools/parse_time.py > gpurun_out/r2y_ncu_parse.log 2>&1
echo "ncu done"
