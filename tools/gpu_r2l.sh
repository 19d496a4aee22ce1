#!/bin/bash
# run L: device stroke parity + timing
mkdir -p gpurun_out
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2l_smoke.log 2>&1 || { echo "SMOKE FAILED"; tail -3 gpurun_out/r2l_smoke.log; exit 1; }
timeout 600 python -m pytest tests/test_gpu_stroke.py -m gpu -x -q > gpurun_out/r2l_pytest.log 2>&1
echo "pytest rc=$?"; tail -15 gpurun_out/r2l_pytest.log
timeout 300 python tools/stroke_time.py > gpurun_out/r2l_stroke_time.txt 2>&1
cat gpurun_out/r2l_stroke_time.txt | tail -12
