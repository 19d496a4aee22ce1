#!/bin/bash
# run BA: seg_classify with four segments in flight and 8-byte loads: tests + ncu
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_batch_api.py tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r2ba_pytest.log 2>&1; echo pytest rc=$?; tail -2 gpurun_out/r2ba_pytest.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"seg_" -c 4 -o gpurun_out/r2ba_compact python bench.py --workload c5 --no-others --steps 2 --warmup 1 > gpurun_out/r2ba_ncu.log 2>&1; echo ncu rc=$?
RGPU_E2E_TRACE=1 timeout 300 python bench.py --workload c5 --no-others --steps 5 --warmup 3 > gpurun_out/r2ba_c5.json 2> gpurun_out/r2ba_c5.err
grep download_runcoded gpurun_out/r2ba_c5.err | tail -2
