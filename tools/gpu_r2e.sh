#!/bin/bash
# round 2, GPU call E: pipelined glyph kernel — smoke (hang guard), tests, A/B, ncu
mkdir -p gpurun_out
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2e_smoke.log 2>&1 || { echo "SMOKE FAILED rc=$?"; tail -5 gpurun_out/r2e_smoke.log; exit 1; }
timeout 300 python -m pytest tests/test_gpu_fuzz.py tests/test_gpu_batch.py tests/test_gpu_batch_api.py tests/test_gpu_scene_kernel.py tests/test_gpu_fill.py -m gpu -x -q > gpurun_out/r2e_pytest.log 2>&1
echo "pytest rc=$?"; tail -8 gpurun_out/r2e_pytest.log
export RB_GLYPHS=20000
timeout 200 python bench.py --workload c4 --no-others --steps 30 --warmup 5 > gpurun_out/r2e_c4_pipe.json 2> gpurun_out/r2e_c4_pipe.err
RGPU_SMALL_NOPIPE=1 timeout 200 python bench.py --workload c4 --no-others --steps 30 --warmup 5 > gpurun_out/r2e_c4_nopipe.json 2> gpurun_out/r2e_c4_nopipe.err
RB_C4_MASK=1 timeout 200 python bench.py --workload c4 --no-others --steps 30 --warmup 5 > gpurun_out/r2e_c4mask_pipe.json 2> gpurun_out/r2e_c4mask_pipe.err
unset RB_GLYPHS
timeout 300 python bench.py --workload c4 --no-others --steps 20 --warmup 3 > gpurun_out/r2e_c4_100k.json 2> gpurun_out/r2e_c4_100k.err
for f in gpurun_out/r2e_c4*.json; do echo $f; python -c "
import json,sys
d=json.load(open('$f'))
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d.get('step_ms_min_med_max'), (d.get('e2e') or {}).get('ms_per_call'))
"; done
tail -3 gpurun_out/r2e_c4_pipe.err
export RB_GLYPHS=4000
timeout 300 ncu --set full --clock-control none --import-source on -k regex:glyph_pipeline -s 2 -c 1 -o gpurun_out/r2e_c4 python tools/prof_step.py c4 3 > gpurun_out/r2e_ncu.log 2>&1
tail -2 gpurun_out/r2e_ncu.log
