#!/bin/bash
# round 2, GPU call G: winding guard + ADVICE regression tests, full suite, bench of every workload
mkdir -p gpurun_out
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2g_smoke.log 2>&1 || { echo "SMOKE FAILED rc=$?"; tail -5 gpurun_out/r2g_smoke.log; exit 1; }
timeout 600 python -m pytest tests/test_gpu_winding.py tests/test_gpu_parity.py tests/test_gpu_scene.py -m gpu -x -q > gpurun_out/r2g_pytest_new.log 2>&1
echo "pytest(new) rc=$?"; tail -12 gpurun_out/r2g_pytest_new.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2g_pytest_all.log 2>&1
echo "pytest(all) rc=$?"; tail -6 gpurun_out/r2g_pytest_all.log
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/r2g_bench_default.json 2> gpurun_out/r2g_bench_default.err
python - <<PY
import json
d=json.load(open('gpurun_out/r2g_bench_default.json'))
print('c4', d['n_gpus'], d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['e2e']['ms_per_call'], d['e2e'].get('rgba8_ms_per_call'))
for k,v in d.get('other_configs',{}).items():
    print(k, v.get('value'), v.get('ms_per_step'), v.get('roofline',{}).get('frac'), v.get('roofline',{}).get('stage_ms'), (v.get('e2e') or {}).get('value'), (v.get('e2e') or {}).get('ms_per_call'), v.get('error'))
PY
tail -3 gpurun_out/r2g_bench_default.err
