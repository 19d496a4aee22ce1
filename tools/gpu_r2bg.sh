#!/bin/bash
# run BG: Paint::at — constant alpha polynomial, early-exit stop search: parity tests, then A/B on config 3 (one box)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fill.py tests/test_gpu_scene.py tests/test_gpu_scene_kernel.py tests/test_gpu_fuzz.py -x -q -m gpu > gpurun_out/r2bg_pytest.log 2>&1; echo pytest rc=$?; tail -2 gpurun_out/r2bg_pytest.log
run() {
timeout 200 python bench.py --workload c3 --no-others --steps 100 --warmup 10 > gpurun_out/r2bg_c3_$1.json 2> gpurun_out/r2bg_c3_$1.err
python -c "
import json
d=json.loads(open('gpurun_out/r2bg_c3_$1.json').read().strip().splitlines()[-1])
print('$1', d['value'], d['ms_per_step'], d['roofline']['frac'], d.get('step_ms_min_med_max'))
"
}
run early1
RGPU_NVCC_EXTRA="-DRGPU_STOPS_EARLY_EXIT=0" timeout 600 python -c "from rasterize_b200 import build; build.build(force=True)" > gpurun_out/r2bg_build_0.log 2>&1 || { echo "build failed"; tail -5 gpurun_out/r2bg_build_0.log; }
run early0
