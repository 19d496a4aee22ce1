#!/bin/bash
# run Z: stroke / parse outputs from the stream-ordered pool: parity + call times
mkdir -p gpurun_out
timeout 180 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2z_smoke.log 2>&1 || { echo "SMOKE FAILED"; tail -5 gpurun_out/r2z_smoke.log; exit 1; }
timeout 900 python -m pytest tests/test_gpu_stroke.py tests/test_gpu_parse.py tests/test_gpu_batch_api.py tests/test_gpu_cpp_host.py -m gpu -x -q > gpurun_out/r2z_pytest.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/r2z_pytest.log | cut -c1-300
timeout 300 python tools/stroke_time.py > gpurun_out/r2z_stroke_time.txt 2>&1; tail -6 gpurun_out/r2z_stroke_time.txt
timeout 300 python tools/parse_time.py > gpurun_out/r2z_parse_time.txt 2>&1; tail -4 gpurun_out/r2z_parse_time.txt
