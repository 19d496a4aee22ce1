#!/bin/bash
# round 2, GPU call C: glyph kernel v3.5 — smoke first (hang guard), focused tests, full suite, A/B bench, ncu
mkdir -p gpurun_out
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2c_smoke.log 2>&1 || { echo "SMOKE FAILED rc=$?"; tail -5 gpurun_out/r2c_smoke.log; exit 1; }
tail -1 gpurun_out/r2c_smoke.log
timeout 300 python -m pytest tests/test_gpu_fuzz.py tests/test_gpu_batch.py tests/test_gpu_batch_api.py tests/test_gpu_scene_kernel.py -m gpu -x -q > gpurun_out/r2c_pytest_small.log 2>&1
echo "pytest(small) rc=$?" | tee -a gpurun_out/r2c_pytest_small.log
tail -15 gpurun_out/r2c_pytest_small.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2c_pytest_all.log 2>&1
echo "pytest(all) rc=$?" | tee -a gpurun_out/r2c_pytest_all.log
tail -15 gpurun_out/r2c_pytest_all.log
export RB_GLYPHS=20000
for mb in 5 4; do
RGPU_SMALL_MINB=$mb timeout 200 python bench.py --workload c4 --no-others --steps 30 --warmup 5 > gpurun_out/r2c_c4_minb$mb.json 2> gpurun_out/r2c_c4_minb$mb.err
done
RB_C4_MASK=1 timeout 200 python bench.py --workload c4 --no-others --steps 30 --warmup 5 > gpurun_out/r2c_c4mask.json 2> gpurun_out/r2c_c4mask.err
for f in gpurun_out/r2c_c4*.json; do echo $f; python -c "
import json,sys
d=json.load(open('$f'))
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_ms'], d.get('step_ms_min_med_max'), d.get('e2e'))
"; done
tail -3 gpurun_out/r2c_c4_minb5.err
export RB_GLYPHS=4000
timeout 300 ncu --set full --clock-control none --import-source on -k regex:small_canvas -s 2 -c 1 -o gpurun_out/r2c_c4 python tools/prof_step.py c4 3 > gpurun_out/r2c_ncu.log 2>&1
tail -2 gpurun_out/r2c_ncu.log
unset RB_GLYPHS
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r2c_bench_default.json 2> gpurun_out/r2c_bench_default.err
tail -c 3000 gpurun_out/r2c_bench_default.json; tail -5 gpurun_out/r2c_bench_default.err
