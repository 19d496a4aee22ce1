#!/bin/bash
# round 2, GPU call J: accumulate variant A/B (lane = line vs span list) + ncu evidence for the committed numbers
mkdir -p gpurun_out
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2j_smoke.log 2>&1 || { echo "SMOKE FAILED"; tail -3 gpurun_out/r2j_smoke.log; exit 1; }
timeout 300 python -m pytest tests/test_gpu_fuzz.py tests/test_gpu_batch.py tests/test_gpu_batch_api.py tests/test_gpu_winding.py tests/test_gpu_scene_kernel.py -m gpu -x -q > gpurun_out/r2j_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r2j_pytest.log
export RB_GLYPHS=20000
timeout 200 python bench.py --workload c4 --no-others --steps 30 --warmup 5 > gpurun_out/r2j_c4_lane.json 2> gpurun_out/r2j_c4_lane.err
RGPU_SMALL_MINB=5 timeout 200 python bench.py --workload c4 --no-others --steps 30 --warmup 5 > gpurun_out/r2j_c4_lane_minb5.json 2> gpurun_out/r2j_c4_lane_minb5.err
export RB_GLYPHS=4000
timeout 300 ncu --set full --clock-control none --import-source on -k regex:small_canvas -s 2 -c 1 -o gpurun_out/r2j_c4 python tools/prof_step.py c4 3 > gpurun_out/r2j_ncu.log 2>&1
export RB_GLYPHS=20000
cp rasterize_b200/librasterize_b200.so /tmp/lib_lane.so
RGPU_NVCC_EXTRA=-DRGPU_ACC_SPANLIST python rasterize_b200/build.py > gpurun_out/r2j_rebuild.log 2>&1
timeout 200 python bench.py --workload c4 --no-others --steps 30 --warmup 5 > gpurun_out/r2j_c4_spanlist.json 2> gpurun_out/r2j_c4_spanlist.err
cp /tmp/lib_lane.so rasterize_b200/librasterize_b200.so
unset RB_GLYPHS
for f in gpurun_out/r2j_c4*.json; do echo $f; python -c "
import json,sys
d=json.load(open('$f'))
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d.get('step_ms_min_med_max'))
"; done
# ---- evidence for the committed numbers
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r2j_launches_default.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r2j_launches_bench.json 2> gpurun_out/r2j_launches_bench.err
echo "launch list rc=$?"; wc -l gpurun_out/r2j_launches_default.csv
timeout 600 ncu --set full --clock-control none --import-source on -k regex:small_canvas -s 2 -c 1 -o gpurun_out/r2j_c4_100k python tools/prof_step.py c4 3 > gpurun_out/r2j_ncu_c4_100k.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:raster_kernel -s 2 -c 1 -o gpurun_out/r2j_c5_raster python tools/prof_step.py c5 3 > gpurun_out/r2j_ncu_c5.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:flatten_bin -s 2 -c 1 -o gpurun_out/r2j_c5_flatten python tools/prof_step.py c5 3 > gpurun_out/r2j_ncu_c5f.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"raster_kernel|flatten_bin" -s 4 -c 2 -o gpurun_out/r2j_c2 python tools/prof_step.py c2 4 > gpurun_out/r2j_ncu_c2.log 2>&1
ls gpurun_out/r2j_*ncu-rep
