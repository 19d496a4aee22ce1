#!/bin/bash
# run AI: rgpu_fill_batch_host uploads the control points chunk by chunk
mkdir -p gpurun_out
timeout 180 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2ai_smoke.log 2>&1 || { echo "SMOKE FAILED"; tail -5 gpurun_out/r2ai_smoke.log; exit 1; }
timeout 600 python -m pytest tests/test_gpu_batch_api.py tests/test_gpu_batch.py tests/test_gpu_winding.py -m gpu -x -q > gpurun_out/r2ai_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r2ai_pytest.log
for v in a b; do
timeout 300 python bench.py --workload c4 --no-others --steps 10 --warmup 3 > gpurun_out/r2ai_c4_$v.json 2> gpurun_out/r2ai_c4_$v.err
python -c "
import json
d=json.load(open('gpurun_out/r2ai_c4_$v.json'))
print('$v', d['ms_per_step'], 'e2e', d['e2e']['ms_per_call'], d['e2e']['value'], d['e2e'].get('rgba8_ms_per_call'), d['e2e']['h2d_bytes_per_step'])
"
done
