#!/bin/bash
# run AV: chunk size of rgpu_fill_batch_host against the call time (does a smaller coverage staging stay in the last-level cache?)
mkdir -p gpurun_out
: > gpurun_out/r2av_chunk.txt
for mb in 128 64 32 16; do
for f in 0.9 1.0; do RGPU_E2E_CHUNK_MB=$mb RGPU_E2E_EXPAND_FRAC=$f timeout 120 python tools/e2e_split.py 100000 5 2>&1 | sed "s/^/chunk=${mb}MB /" >> gpurun_out/r2av_chunk.txt; done
done
RGPU_E2E_CHUNK_MB=32 RGPU_E2E_EXPAND=0 timeout 120 python tools/e2e_split.py 100000 5 2>&1 | sed "s/^/chunk=32MB /" >> gpurun_out/r2av_chunk.txt
cat gpurun_out/r2av_chunk.txt
