#!/bin/bash
# round 2, GPU call F (N GPUs of one box): multi-GPU library tests + the driver's bench at N ranks
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/r2f_gpus_n$N.txt 2>&1
df -h /dev/shm | tail -1 >> gpurun_out/r2f_gpus_n$N.txt; nproc >> gpurun_out/r2f_gpus_n$N.txt; free -g | head -2 >> gpurun_out/r2f_gpus_n$N.txt
timeout 300 python -m pytest tests/test_gpu_batch_api.py -m gpu -x -q -k "multi or banded" > gpurun_out/r2f_pytest_multi_n$N.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r2f_pytest_multi_n$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r2f_bench_n$N.json 2> gpurun_out/r2f_bench_n$N.err
echo "bench rc=$?"; tail -3 gpurun_out/r2f_bench_n$N.err
python - <<PY
import json
d=json.load(open('gpurun_out/r2f_bench_n$N.json'))
print('c4', d['n_gpus'], d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['e2e']['ms_per_call'], d['e2e'].get('rgba8_ms_per_call'))
for k,v in d.get('other_configs',{}).items():
    print(k, v.get('value'), v.get('ms_per_step'), v.get('roofline',{}).get('frac'), (v.get('e2e') or {}).get('value'), (v.get('e2e') or {}).get('ms_per_call'), v.get('error'))
PY
