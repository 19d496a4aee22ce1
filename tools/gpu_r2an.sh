#!/bin/bash
# run AN: split download — call time against the host-expanded share
mkdir -p gpurun_out
L=gpurun_out/r2an_split.txt
: > $L
nproc >> $L
RGPU_E2E_EXPAND=0 RGPU_E2E_TRACE=1 timeout 120 python tools/e2e_split.py 100000 4 >> $L 2>&1
for f in 0.3 0.5 0.6 0.7 0.8 0.9; do
RGPU_E2E_EXPAND_FRAC=$f RGPU_E2E_TRACE=1 timeout 120 python tools/e2e_split.py 100000 4 >> $L 2>&1
done
RGPU_HOST_THREADS=8 RGPU_E2E_EXPAND_FRAC=0.7 timeout 120 python tools/e2e_split.py 100000 4 >> $L 2>&1
grep -v "^rgpu_fill" $L
grep "^rgpu_fill" $L | awk 'NR%4==0'
