#!/bin/bash
# run Q: interleaved slot assignment in the glyph kernel
mkdir -p gpurun_out
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2q_smoke.log 2>&1 || { echo "SMOKE FAILED"; tail -3 gpurun_out/r2q_smoke.log; exit 1; }
timeout 400 python -m pytest tests/test_gpu_fuzz.py tests/test_gpu_batch.py tests/test_gpu_batch_api.py tests/test_gpu_winding.py tests/test_gpu_scene_kernel.py -m gpu -x -q > gpurun_out/r2q_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r2q_pytest.log
for v in a b; do
timeout 200 python bench.py --workload c4 --no-others --steps 20 --warmup 3 > gpurun_out/r2q_c4_100k_$v.json 2> gpurun_out/r2q_c4_100k_$v.err
python -c "
import json
d=json.load(open('gpurun_out/r2q_c4_100k_$v.json'))
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d.get('step_ms_min_med_max'))
"
done
export RB_GLYPHS=4000
timeout 300 ncu --set full --clock-control none --import-source on -k regex:small_canvas -s 2 -c 1 -o gpurun_out/r2q_c4 python tools/prof_step.py c4 3 > gpurun_out/r2q_ncu.log 2>&1
echo "ncu rc=$?"
