#!/bin/bash
# run AU: expansion tasks queued with their copies (pool threads sleep on the copy event), FIFO pool
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_batch_api.py tests/test_gpu_parity.py tests/test_gpu_batch.py -x -q -m gpu > gpurun_out/r2au_pytest.log 2>&1; echo pytest rc=$?; tail -4 gpurun_out/r2au_pytest.log
RGPU_E2E_TRACE=1 timeout 120 python tools/e2e_split.py 100000 14 > gpurun_out/r2au_split.txt 2>&1
grep -v "^rgpu_fill" gpurun_out/r2au_split.txt
grep "^rgpu_fill" gpurun_out/r2au_split.txt | tail -3
for f in 0.8 0.9 1.0; do RGPU_E2E_EXPAND_FRAC=$f timeout 120 python tools/e2e_split.py 100000 5; done
