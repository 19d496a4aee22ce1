#!/bin/bash
# run AG: plain render path, direct stores vs staging through shared memory (A/B on one box)
mkdir -p gpurun_out
timeout 180 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2ag_smoke.log 2>&1 || { echo "SMOKE FAILED"; tail -5 gpurun_out/r2ag_smoke.log; exit 1; }
timeout 600 python -m pytest tests/test_gpu_fuzz.py tests/test_gpu_batch.py tests/test_gpu_batch_api.py tests/test_gpu_fill.py -m gpu -x -q > gpurun_out/r2ag_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r2ag_pytest.log
run() {
timeout 200 python bench.py --workload c4 --no-others --steps 20 --warmup 3 > gpurun_out/r2ag_c4_$1.json 2> gpurun_out/r2ag_c4_$1.err
python -c "
import json
d=json.load(open('gpurun_out/r2ag_c4_$1.json'))
print('$1', d['value'], d['ms_per_step'], d['roofline']['frac'], d.get('step_ms_min_med_max'))
"
}
run direct
RGPU_NVCC_EXTRA="-DRGPU_PLAIN_DIRECT=0" timeout 600 python -c "from rasterize_b200 import build; build.build(force=True)" > gpurun_out/r2ag_build.log 2>&1 || { echo "build failed"; tail -5 gpurun_out/r2ag_build.log; }
run staged
