#!/bin/bash
# run AT: run-coded download of large masks
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_batch_api.py tests/test_gpu_parity.py tests/test_gpu_fuzz.py -x -q -m gpu > gpurun_out/r2at_pytest.log 2>&1; echo pytest rc=$?; tail -8 gpurun_out/r2at_pytest.log
RGPU_E2E_TRACE=1 timeout 300 python bench.py --workload c5 --no-others --steps 5 --warmup 3 > gpurun_out/r2at_c5.json 2> gpurun_out/r2at_c5.err; echo rc=$?
grep download_runcoded gpurun_out/r2at_c5.err | tail -3
RGPU_E2E_RUNCODE=0 timeout 300 python bench.py --workload c5 --no-others --steps 5 --warmup 3 > gpurun_out/r2at_c5_dense.json 2> gpurun_out/r2at_c5_dense.err
RGPU_E2E_TRACE=1 timeout 300 python bench.py --workload c2 --no-others --steps 20 --warmup 3 > gpurun_out/r2at_c2.json 2> gpurun_out/r2at_c2.err
grep download_runcoded gpurun_out/r2at_c2.err | tail -2
python - <<PY
import json
for f in ("c5","c5_dense","c2"):
    d=json.loads(open(f"gpurun_out/r2at_{f}.json").read().strip().splitlines()[-1])
    print(f, d["value"], d["ms_per_step"], {k:v for k,v in d["e2e"].items() if k!="call"})
PY
