#!/bin/bash
# run AZ: coverage share copied in pieces (does a piece stay in the last-level cache until it is expanded?)
mkdir -p gpurun_out
: > gpurun_out/r2az_pieces.txt
for mb in 0 8 2 1; do
for f in 0.8 0.9 1.0; do RGPU_E2E_PIECE_MB=$mb RGPU_E2E_EXPAND_FRAC=$f timeout 120 python tools/e2e_split.py 100000 5 2>&1 | sed "s/^/piece=${mb}MB /" >> gpurun_out/r2az_pieces.txt; done
done
cat gpurun_out/r2az_pieces.txt
timeout 300 python -m pytest tests/test_gpu_batch_api.py -x -q -m gpu 2>&1 | tail -2
