#!/bin/bash
# run AR: host expansion / widening with AVX-512 streaming stores
mkdir -p gpurun_out
grep -o "avx512f\|avx2" /proc/cpuinfo | sort | uniq -c
timeout 600 python -m pytest tests/test_gpu_batch_api.py tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r2ar_pytest.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/r2ar_pytest.log
RGPU_E2E_TRACE=1 timeout 120 python tools/e2e_split.py 100000 16 > gpurun_out/r2ar_split.txt 2>&1
grep -v "^rgpu_fill" gpurun_out/r2ar_split.txt
grep "^rgpu_fill" gpurun_out/r2ar_split.txt | sed 's/.*next share/next share/' | tr '\n' ' '
for f in 0.9 1.0; do RGPU_E2E_EXPAND_FRAC=$f timeout 120 python tools/e2e_split.py 100000 5; done
RGPU_E2E_TRACE=1 python tools/e2e_trace.py 2>&1 | tail -4
