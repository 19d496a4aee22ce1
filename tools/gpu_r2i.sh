#!/bin/bash
# round 2, GPU call I: ncu evidence for the committed numbers (launch list of the default bench command, full captures)
mkdir -p gpurun_out
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2i_smoke.log 2>&1 || { echo "SMOKE FAILED"; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r2i_launches_default.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r2i_launches_bench.json 2> gpurun_out/r2i_launches_bench.err
echo "launch list rc=$?"; wc -l gpurun_out/r2i_launches_default.csv
timeout 600 ncu --set full --clock-control none --import-source on -k regex:small_canvas -s 2 -c 1 -o gpurun_out/r2i_c4_100k python tools/prof_step.py c4 3 > gpurun_out/r2i_ncu_c4.log 2>&1
tail -1 gpurun_out/r2i_ncu_c4.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:raster_kernel -s 2 -c 1 -o gpurun_out/r2i_c5_raster python tools/prof_step.py c5 3 > gpurun_out/r2i_ncu_c5.log 2>&1
tail -1 gpurun_out/r2i_ncu_c5.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:flatten_bin -s 2 -c 1 -o gpurun_out/r2i_c5_flatten python tools/prof_step.py c5 3 > gpurun_out/r2i_ncu_c5f.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"raster_kernel|flatten_bin" -s 4 -c 2 -o gpurun_out/r2i_c2 python tools/prof_step.py c2 4 > gpurun_out/r2i_ncu_c2.log 2>&1
tail -1 gpurun_out/r2i_ncu_c2.log
ls -la gpurun_out/r2i_*
