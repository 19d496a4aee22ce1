#!/bin/bash
# run P: parser with 16-byte windows: parity + timing
mkdir -p gpurun_out
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2p_smoke.log 2>&1 || { echo "SMOKE FAILED"; tail -3 gpurun_out/r2p_smoke.log; exit 1; }
timeout 600 python -m pytest tests/test_gpu_parse.py -m gpu -x -q > gpurun_out/r2p_pytest.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/r2p_pytest.log | cut -c1-300
timeout 300 python tools/parse_time.py > gpurun_out/r2p_parse_time.txt 2>&1
tail -5 gpurun_out/r2p_parse_time.txt
