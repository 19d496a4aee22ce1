#!/bin/bash
# run BE: cut depth of the flatten kernel on config 2 (material.path, 24 235 items), A/B on one box
mkdir -p gpurun_out
: > gpurun_out/r2be_c2_cut.txt
for rep in 1 2; do for d in 3 4 5; do
RGPU_CUT_DEPTH=$d timeout 200 python bench.py --workload c2 --no-others --steps 300 --warmup 20 > gpurun_out/r2be_c2_d$d.json 2> gpurun_out/r2be_c2_d$d.err
python -c "
import json
d=json.loads(open('gpurun_out/r2be_c2_d$d.json').read().strip().splitlines()[-1])
print('depth $d', d['ms_per_step'], d['roofline']['stage_ms'], d.get('step_ms_min_med_max'))
" >> gpurun_out/r2be_c2_cut.txt
done; done
cat gpurun_out/r2be_c2_cut.txt
