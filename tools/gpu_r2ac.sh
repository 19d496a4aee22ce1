#!/bin/bash
# run AC: glyph kernel, two classes of lines in the queue (A/B on one box)
mkdir -p gpurun_out
timeout 180 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2ac_smoke.log 2>&1 || { echo "SMOKE FAILED"; tail -5 gpurun_out/r2ac_smoke.log; exit 1; }
timeout 600 python -m pytest tests/test_gpu_fuzz.py tests/test_gpu_batch.py tests/test_gpu_batch_api.py tests/test_gpu_winding.py tests/test_gpu_scene_kernel.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2ac_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r2ac_pytest.log
run() {
timeout 200 python bench.py --workload c4 --no-others --steps 20 --warmup 3 > gpurun_out/r2ac_c4_$1.json 2> gpurun_out/r2ac_c4_$1.err
python -c "
import json
d=json.load(open('gpurun_out/r2ac_c4_$1.json'))
print('$1', d['value'], d['ms_per_step'], d['roofline']['frac'], d.get('step_ms_min_med_max'))
"
}
run two
export RB_GLYPHS=4000
timeout 300 ncu --set full --clock-control none --import-source on -k regex:small_canvas -s 2 -c 1 -o gpurun_out/r2ac_c4 python tools/prof_step.py c4 3 > gpurun_out/r2ac_ncu.log 2>&1
echo "ncu rc=$?"
unset RB_GLYPHS
RGPU_NVCC_EXTRA="-DRGPU_TWO_CLASS=0" timeout 600 python -c "from rasterize_b200 import build; build.build(force=True)" > gpurun_out/r2ac_build.log 2>&1 || { echo "build failed"; tail -5 gpurun_out/r2ac_build.log; }
run one
