#!/bin/bash
# run W: one rank's block of c5 on one GPU (as at N = 8): dense middle block vs empty block, whole-warp wide spans on / off
mkdir -p gpurun_out
timeout 180 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2w_smoke.log 2>&1 || { echo "SMOKE FAILED"; tail -5 gpurun_out/r2w_smoke.log; exit 1; }
run() {
for blk in 24:32 0:8 16:24; do
RB_C5_BLOCK=$blk timeout 200 python bench.py --workload c5 --no-others --steps 30 --warmup 5 > gpurun_out/r2w_c5_$1_$blk.json 2> gpurun_out/r2w_c5_$1_$blk.err
python -c "
import json
d=json.load(open('gpurun_out/r2w_c5_$1_$blk.json'))
print('$1', '$blk', d['ms_per_step'], d['roofline']['stage_ms'])
"
done
}
run off
for c in 24 64; do
RGPU_NVCC_EXTRA="-DRGPU_COOP_SPAN=$c" timeout 600 python -c "from rasterize_b200 import build; build.build(force=True)" > gpurun_out/r2w_build_$c.log 2>&1 || { echo "build failed"; tail -5 gpurun_out/r2w_build_$c.log; }
run coop$c
done
