#!/bin/bash
# run AS (N GPUs of one box): the driver's bench at N ranks with the split download; share trace of every rank
N=${1:-4}
mkdir -p gpurun_out
nproc > gpurun_out/r2as_host_n$N.txt; free -g | head -2 >> gpurun_out/r2as_host_n$N.txt
timeout 300 python -m pytest tests/test_gpu_batch_api.py -m gpu -x -q -k "multi or banded or run_coded or split" > gpurun_out/r2as_pytest_n$N.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2as_pytest_n$N.log
RGPU_E2E_TRACE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r2as_bench_n$N.json 2> gpurun_out/r2as_bench_n$N.err
echo "bench rc=$?"
grep "rgpu_fill_batch_host" gpurun_out/r2as_bench_n$N.err | sed 's/.*next share/share/' | tr '\n' ' ' | cut -c1-1500
python - <<PY
import json
d=json.loads(open('gpurun_out/r2as_bench_n$N.json').read().strip().splitlines()[-1])
print('c4', d['n_gpus'], d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['e2e']['ms_per_call'], d['e2e'].get('rgba8_ms_per_call'), d['e2e']['d2h_bytes_per_step'])
for k,v in d.get('other_configs',{}).items():
    print(k, v.get('value'), v.get('ms_per_step'), v.get('roofline',{}).get('frac'), (v.get('e2e') or {}).get('value'), (v.get('e2e') or {}).get('ms_per_call'), v.get('error'))
PY
