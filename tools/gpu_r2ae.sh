#!/bin/bash
# run AE: does running next to the GPU (NVML CPU affinity) change the e2e legs?  same box, both ways
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2ae_topo.txt 2>&1; lscpu | grep -i "numa\|socket\|model name\|^CPU(s)" >> gpurun_out/r2ae_topo.txt; nproc >> gpurun_out/r2ae_topo.txt
cat gpurun_out/r2ae_topo.txt | head -30
run() {
timeout 300 python bench.py --workload c4 --no-others --steps 10 --warmup 3 > gpurun_out/r2ae_c4_$1.json 2> gpurun_out/r2ae_c4_$1.err
python -c "
import json
d=json.load(open('gpurun_out/r2ae_c4_$1.json'))
print('$1', d['ms_per_step'], 'e2e', d['e2e']['ms_per_call'], d['e2e'].get('rgba8_ms_per_call'), 'cpus', d.get('host_cpus_of_this_rank'))
"
}
run aff
RB_NO_AFFINITY=1 run noaff
run aff2
RB_NO_AFFINITY=1 run noaff2
