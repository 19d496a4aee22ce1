#!/bin/bash
# round 2, GPU call B: glyph kernel v3 correctness + A/B (occupancy variants, accumulate inline / noinline) + ncu
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_batch.py tests/test_gpu_fuzz.py tests/test_gpu_fill.py -m gpu -x -q > gpurun_out/r2b_pytest_small.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2b_pytest_small.log
tail -5 gpurun_out/r2b_pytest_small.log
export RB_GLYPHS=20000
for mb in 4 5 6; do
RGPU_SMALL_MINB=$mb timeout 300 python bench.py --workload c4 --steps 30 --warmup 5 > gpurun_out/r2b_c4_minb$mb.json 2> gpurun_out/r2b_c4_minb$mb.err
done
RB_C4_MASK=1 timeout 300 python bench.py --workload c4 --steps 30 --warmup 5 > gpurun_out/r2b_c4mask.json 2> gpurun_out/r2b_c4mask.err
export RB_GLYPHS=4000
timeout 600 ncu --set full --clock-control none --import-source on -k regex:small_canvas -s 2 -c 1 -o gpurun_out/r2b_c4 python tools/prof_step.py c4 3 > gpurun_out/r2b_ncu.log 2>&1
export RB_GLYPHS=20000
RGPU_NVCC_EXTRA=-DRGPU_ACC_NOINLINE python rasterize_b200/build.py > gpurun_out/r2b_rebuild.log 2>&1
for mb in 4 5; do
RGPU_SMALL_MINB=$mb timeout 300 python bench.py --workload c4 --steps 30 --warmup 5 > gpurun_out/r2b_c4_noinl_minb$mb.json 2> gpurun_out/r2b_c4_noinl_minb$mb.err
done
for f in gpurun_out/r2b_c4*.json; do echo $f; python -c "
import json,sys
d=json.load(open('$f'))
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_ms'], d.get('step_ms_min_med_max'))
"; done
