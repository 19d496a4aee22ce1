#!/bin/bash
# round 2, GPU call A: correctness of the new glyph kernel + A/B against the round-1 kernel + pipe rates + ncu
mkdir -p gpurun_out
tools/ubench/pipe_rate > gpurun_out/r2a_pipe_rate.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2a_pytest_gpu.log
export RB_GLYPHS=20000
timeout 300 python bench.py --workload c4 --steps 30 --warmup 5 > gpurun_out/r2a_c4_v2_minb8.json 2> gpurun_out/r2a_c4_v2_minb8.err
RGPU_SMALL_MINB=6 timeout 300 python bench.py --workload c4 --steps 30 --warmup 5 > gpurun_out/r2a_c4_v2_minb6.json 2> gpurun_out/r2a_c4_v2_minb6.err
RGPU_SMALL_V1=1 timeout 300 python bench.py --workload c4 --steps 30 --warmup 5 > gpurun_out/r2a_c4_v1.json 2> gpurun_out/r2a_c4_v1.err
RB_C4_MASK=1 timeout 300 python bench.py --workload c4 --steps 30 --warmup 5 > gpurun_out/r2a_c4mask_v2.json 2> gpurun_out/r2a_c4mask_v2.err
export RB_GLYPHS=4000
timeout 600 ncu --set full --clock-control none --import-source on -k regex:small_canvas -s 2 -c 1 -o gpurun_out/r2a_c4_v2 python tools/prof_step.py c4 3 > gpurun_out/r2a_ncu.log 2>&1
tail -3 gpurun_out/r2a_pytest_gpu.log
cat gpurun_out/r2a_pipe_rate.txt
for f in gpurun_out/r2a_c4*.json; do echo $f; python -c "
import json,sys
d=json.load(open('$f'))
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_ms'], d.get('step_ms_min_med_max'))
"; done
