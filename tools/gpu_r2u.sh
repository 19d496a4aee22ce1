#!/bin/bash
# run U: device-resident stroke test, extended smoke, glyph kernel at 5 CTAs per SM
mkdir -p gpurun_out
timeout 180 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2u_smoke.log 2>&1 || { echo "SMOKE FAILED"; tail -5 gpurun_out/r2u_smoke.log; exit 1; }
tail -1 gpurun_out/r2u_smoke.log
timeout 600 python -m pytest tests/test_gpu_stroke.py tests/test_gpu_parse.py tests/test_gpu_cpp_host.py -m gpu -x -q > gpurun_out/r2u_pytest.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/r2u_pytest.log | cut -c1-300
run() {
timeout 200 python bench.py --workload c4 --no-others --steps 20 --warmup 3 > gpurun_out/r2u_c4_$1.json 2> gpurun_out/r2u_c4_$1.err
python -c "
import json
d=json.load(open('gpurun_out/r2u_c4_$1.json'))
print('$1', d['value'], d['ms_per_step'], d['roofline']['frac'], d.get('step_ms_min_med_max'))
"
}
run minb4
RGPU_SMALL_MINB=5 run minb5
