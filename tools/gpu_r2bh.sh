#!/bin/bash
# run BH: glyph kernel — row pairs without a cell skip the scan (plain solid fills), A/B on one box + parity tests
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_batch.py tests/test_gpu_batch_api.py -x -q -m gpu > gpurun_out/r2bh_pytest.log 2>&1; echo pytest rc=$?; tail -2 gpurun_out/r2bh_pytest.log
run() {
timeout 200 python bench.py --workload c4 --no-others --steps 30 --warmup 5 > gpurun_out/r2bh_c4_$1.json 2> gpurun_out/r2bh_c4_$1.err
python -c "
import json
d=json.loads(open('gpurun_out/r2bh_c4_$1.json').read().strip().splitlines()[-1])
print('$1', d['value'], d['ms_per_step'], d['roofline']['frac'], d.get('step_ms_min_med_max'))
"
}
run skip1
RGPU_NVCC_EXTRA="-DRGPU_SKIP_EMPTY_ROWS=0" timeout 600 python -c "from rasterize_b200 import build; build.build(force=True)" > gpurun_out/r2bh_build_0.log 2>&1 || { echo "build failed"; tail -5 gpurun_out/r2bh_build_0.log; }
run skip0
