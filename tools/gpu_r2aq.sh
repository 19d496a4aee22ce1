#!/bin/bash
# run AQ: mask_iter compaction on the device, column-strided fill, share table of the split download
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fill.py tests/test_gpu_fuzz.py tests/test_gpu_batch_api.py -x -q -m gpu > gpurun_out/r2aq_pytest.log 2>&1; echo pytest rc=$?; tail -15 gpurun_out/r2aq_pytest.log
RGPU_E2E_TRACE=1 timeout 120 python tools/e2e_split.py 100000 16 > gpurun_out/r2aq_split.txt 2>&1
grep -v "^rgpu_fill" gpurun_out/r2aq_split.txt
grep "^rgpu_fill" gpurun_out/r2aq_split.txt | sed 's/.*next share/next share/' | tr '\n' ' '
