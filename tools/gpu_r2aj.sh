#!/bin/bash
# run AJ: per-warp line queue capacity (one drain per glyph wants the whole glyph to fit)
mkdir -p gpurun_out
timeout 180 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2aj_smoke.log 2>&1 || { echo "SMOKE FAILED"; tail -5 gpurun_out/r2aj_smoke.log; exit 1; }
run() {
timeout 200 python bench.py --workload c4 --no-others --steps 20 --warmup 3 > gpurun_out/r2aj_c4_$1.json 2> gpurun_out/r2aj_c4_$1.err
python -c "
import json
d=json.load(open('gpurun_out/r2aj_c4_$1.json'))
print('$1', d['value'], d['ms_per_step'], d['roofline']['frac'], d.get('step_ms_min_med_max'))
"
}
run q160
for q in 192 224 256; do
RGPU_NVCC_EXTRA="-DRGPU_GQUEUE=$q" timeout 600 python -c "from rasterize_b200 import build; build.build(force=True)" > gpurun_out/r2aj_build_$q.log 2>&1 || { echo "build failed"; tail -5 gpurun_out/r2aj_build_$q.log; continue; }
run q$q
done
