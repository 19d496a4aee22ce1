#!/bin/bash
mkdir -p gpurun_out
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2h_smoke.log 2>&1 || { echo "SMOKE FAILED rc=$?"; tail -5 gpurun_out/r2h_smoke.log; exit 1; }
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2h_pytest_all.log 2>&1
echo "pytest(all) rc=$?"; tail -12 gpurun_out/r2h_pytest_all.log
