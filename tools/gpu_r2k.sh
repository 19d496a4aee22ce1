#!/bin/bash
mkdir -p gpurun_out
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2k_smoke.log 2>&1 || { echo "SMOKE FAILED"; tail -3 gpurun_out/r2k_smoke.log; exit 1; }
timeout 300 python -m pytest tests/test_gpu_fuzz.py tests/test_gpu_batch.py tests/test_gpu_batch_api.py tests/test_gpu_winding.py tests/test_gpu_scene_kernel.py -m gpu -x -q > gpurun_out/r2k_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r2k_pytest.log
export RB_GLYPHS=20000
timeout 200 python bench.py --workload c4 --no-others --steps 30 --warmup 5 > gpurun_out/r2k_c4.json 2> gpurun_out/r2k_c4.err
unset RB_GLYPHS
timeout 200 python bench.py --workload c4 --no-others --steps 20 --warmup 3 > gpurun_out/r2k_c4_100k.json 2> gpurun_out/r2k_c4_100k.err
for f in gpurun_out/r2k_c4*.json; do echo $f; python -c "
import json,sys
d=json.load(open('$f'))
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d.get('step_ms_min_med_max'))
"; done
export RB_GLYPHS=4000
timeout 300 ncu --set full --clock-control none --import-source on -k regex:small_canvas -s 2 -c 1 -o gpurun_out/r2k_c4 python tools/prof_step.py c4 3 > gpurun_out/r2k_ncu.log 2>&1
