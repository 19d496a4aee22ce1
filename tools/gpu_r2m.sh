#!/bin/bash
mkdir -p gpurun_out
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2m_smoke.log 2>&1 || { echo "SMOKE FAILED"; tail -3 gpurun_out/r2m_smoke.log; exit 1; }
timeout 300 python tools/stroke_diff.py huyak material > gpurun_out/r2m_diff.txt 2>&1
tail -60 gpurun_out/r2m_diff.txt
