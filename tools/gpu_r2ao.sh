#!/bin/bash
# run AO: split download with the chunk preparation one chunk ahead on a helper thread
mkdir -p gpurun_out
L=gpurun_out/r2ao_split.txt
: > $L
timeout 300 python -m pytest tests/test_gpu_batch_api.py tests/test_gpu_batch.py -x -q -m gpu > gpurun_out/r2ao_pytest.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/r2ao_pytest.log
RGPU_E2E_EXPAND=0 RGPU_E2E_TRACE=1 timeout 120 python tools/e2e_split.py 100000 4 >> $L 2>&1
for f in 0.5 0.7 0.8 0.9 0.95; do
RGPU_E2E_EXPAND_FRAC=$f RGPU_E2E_TRACE=1 timeout 120 python tools/e2e_split.py 100000 4 >> $L 2>&1
done
RGPU_E2E_TRACE=1 timeout 120 python tools/e2e_split.py 100000 16 >> $L 2>&1
grep -v "^rgpu_fill" $L
grep "^rgpu_fill" $L | awk 'NR%4==0 || NR>24'
