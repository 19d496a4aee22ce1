#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/d2h_streams.py > gpurun_out/r2ah_d2h_streams.txt 2>&1
cat gpurun_out/r2ah_d2h_streams.txt
