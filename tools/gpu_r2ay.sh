#!/bin/bash
# run AY: compute-sanitizer over the new kernels (pixel list, run-coded rows) and the split download; ncu --set full of the
# compaction kernels on config 5's canvas
mkdir -p gpurun_out
S="tests/test_gpu_parity.py::test_mask_iter_device_compaction tests/test_gpu_parity.py::test_mask_iter_capacity_and_empty tests/test_gpu_parity.py::test_mask_f64_run_coded_strided_rows tests/test_gpu_batch_api.py::test_run_coded_download_matches_dense_copies tests/test_gpu_batch_api.py::test_fill_batch_host_split_download_is_bit_identical tests/test_gpu_fill.py::test_fill_column_strided_view"
timeout 900 compute-sanitizer --tool memcheck --leak-check no python -m pytest $S -x -q > gpurun_out/r2ay_memcheck.txt 2>&1; echo memcheck rc=$?; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r2ay_memcheck.txt | tail -3
timeout 900 compute-sanitizer --tool racecheck python -m pytest $S -x -q > gpurun_out/r2ay_racecheck.txt 2>&1; echo racecheck rc=$?; grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/r2ay_racecheck.txt | tail -3
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"seg_|pixel_" -c 6 -o gpurun_out/r2ay_compact python bench.py --workload c5 --no-others --steps 2 --warmup 1 > gpurun_out/r2ay_ncu.log 2>&1; echo ncu rc=$?
ls -la gpurun_out/r2ay_compact.ncu-rep
