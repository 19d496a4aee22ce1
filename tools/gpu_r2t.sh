#!/bin/bash
# run T: tiled raster kernel, long wide spans by the whole warp (lane = column): parity + c5 / c2 / c3 A/B on one box
mkdir -p gpurun_out
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2t_smoke.log 2>&1 || { echo "SMOKE FAILED"; tail -3 gpurun_out/r2t_smoke.log; exit 1; }
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2t_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r2t_pytest.log
run() {
for wl in c5 c2 c3; do
timeout 200 python bench.py --workload $wl --no-others --steps 30 --warmup 5 > gpurun_out/r2t_${wl}_$1.json 2> gpurun_out/r2t_${wl}_$1.err
python -c "
import json
d=json.load(open('gpurun_out/r2t_${wl}_$1.json'))
print('$1', '$wl', d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['stage_ms'])
"
done
}
run coop24
sed -i 's/constexpr int kCoopSpan = 24;/constexpr int kCoopSpan = 1 << 20;/' rasterize_b200/csrc/raster_device.cuh
timeout 600 python -c "from rasterize_b200 import build; build.build(force=True)" > gpurun_out/r2t_build_off.log 2>&1 || { echo "build failed"; tail -5 gpurun_out/r2t_build_off.log; }
run off
sed -i 's/constexpr int kCoopSpan = 1 << 20;/constexpr int kCoopSpan = 64;/' rasterize_b200/csrc/raster_device.cuh
timeout 600 python -c "from rasterize_b200 import build; build.build(force=True)" > gpurun_out/r2t_build_64.log 2>&1 || { echo "build failed"; tail -5 gpurun_out/r2t_build_64.log; }
run coop64
