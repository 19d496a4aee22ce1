#!/bin/bash
# run AF: evidence for the committed state: launch list of the default bench command, full captures of the c4 kernel (100 000 and 4 000 glyphs)
mkdir -p gpurun_out
timeout 180 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2af_smoke.log 2>&1 || { echo "SMOKE FAILED"; tail -5 gpurun_out/r2af_smoke.log; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2af_launches_default.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r2af_launches_bench.json 2> gpurun_out/r2af_launches_bench.err
echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:small_canvas -s 2 -c 1 -o gpurun_out/r2af_c4_100k python tools/prof_step.py c4 3 > gpurun_out/r2af_ncu_100k.log 2>&1
echo "ncu 100k rc=$?"
