#!/bin/bash
# run AL: prologue variants on one box
mkdir -p gpurun_out
timeout 180 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2al_smoke.log 2>&1 || { echo "SMOKE FAILED"; tail -5 gpurun_out/r2al_smoke.log; exit 1; }
run() {
timeout 200 python bench.py --workload c4 --no-others --steps 20 --warmup 3 > gpurun_out/r2al_c4_$1.json 2> gpurun_out/r2al_c4_$1.err
python -c "
import json
d=json.load(open('gpurun_out/r2al_c4_$1.json'))
print('$1', d['value'], d['ms_per_step'], d['roofline']['frac'], d.get('step_ms_min_med_max'))
"
}
run direct_hint
for v in "0 0" "0 1" "1 0"; do
set -- $v
RGPU_NVCC_EXTRA="-DRGPU_STAGE_DIRECT=$1 -DRGPU_STAGE_HINT=$2" timeout 600 python -c "from rasterize_b200 import build; build.build(force=True)" > gpurun_out/r2al_build_$1$2.log 2>&1 || { echo "build failed"; tail -5 gpurun_out/r2al_build_$1$2.log; continue; }
run direct$1_hint$2
done
