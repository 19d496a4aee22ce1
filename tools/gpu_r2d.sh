#!/bin/bash
# round 2, GPU call D: balanced slots A/B, c5 merged bands, c2, ncu of c4
mkdir -p gpurun_out
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2d_smoke.log 2>&1 || { echo "SMOKE FAILED rc=$?"; tail -5 gpurun_out/r2d_smoke.log; exit 1; }
timeout 300 python -m pytest tests/test_gpu_fuzz.py tests/test_gpu_batch.py tests/test_gpu_batch_api.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2d_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r2d_pytest.log
export RB_GLYPHS=20000
for mb in 5 4; do
RGPU_SMALL_MINB=$mb timeout 200 python bench.py --workload c4 --no-others --steps 30 --warmup 5 > gpurun_out/r2d_c4_minb$mb.json 2> gpurun_out/r2d_c4_minb$mb.err
done
unset RB_GLYPHS
timeout 200 python bench.py --workload c5 --no-others --steps 20 --warmup 3 > gpurun_out/r2d_c5.json 2> gpurun_out/r2d_c5.err
timeout 200 python bench.py --workload c2 --no-others --steps 100 --warmup 10 > gpurun_out/r2d_c2.json 2> gpurun_out/r2d_c2.err
for f in gpurun_out/r2d_c*.json; do echo $f; python -c "
import json,sys
d=json.load(open('$f'))
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['stage_ms'], d.get('step_ms_min_med_max'), (d.get('e2e') or {}).get('ms_per_call'))
"; done
export RB_GLYPHS=4000
timeout 300 ncu --set full --clock-control none --import-source on -k regex:small_canvas -s 2 -c 1 -o gpurun_out/r2d_c4 python tools/prof_step.py c4 3 > gpurun_out/r2d_ncu.log 2>&1
tail -2 gpurun_out/r2d_ncu.log
