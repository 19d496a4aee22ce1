#!/bin/bash
# run AK: glyph kernel prologue: one barrier, first chunk staged from the descriptor in global memory, uniform-cubic hint
mkdir -p gpurun_out
timeout 180 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2ak_smoke.log 2>&1 || { echo "SMOKE FAILED"; tail -5 gpurun_out/r2ak_smoke.log; exit 1; }
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2ak_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r2ak_pytest.log | cut -c1-200
for v in a b; do
timeout 200 python bench.py --workload c4 --no-others --steps 20 --warmup 3 > gpurun_out/r2ak_c4_$v.json 2> gpurun_out/r2ak_c4_$v.err
python -c "
import json
d=json.load(open('gpurun_out/r2ak_c4_$v.json'))
print('$v', d['value'], d['ms_per_step'], d['roofline']['frac'], d.get('step_ms_min_med_max'), 'e2e', d['e2e']['ms_per_call'])
"
done
