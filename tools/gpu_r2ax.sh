#!/bin/bash
# run AX: rgpu_mask (f64, C2) with and without the run-coded attempt
mkdir -p gpurun_out
RGPU_E2E_TRACE=1 python tools/e2e_trace.py > gpurun_out/r2ax_c2_runcode.txt 2>&1
RGPU_E2E_RUNCODE=0 RGPU_E2E_TRACE=1 python tools/e2e_trace.py > gpurun_out/r2ax_c2_dense.txt 2>&1
tail -5 gpurun_out/r2ax_c2_runcode.txt; echo; tail -4 gpurun_out/r2ax_c2_dense.txt
