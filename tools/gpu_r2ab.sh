#!/bin/bash
# run AB: compute-sanitizer over the new code paths (stroke, parse, tall lines in the glyph kernel)
mkdir -p gpurun_out
timeout 180 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2ab_smoke.log 2>&1 || { echo "SMOKE FAILED"; tail -5 gpurun_out/r2ab_smoke.log; exit 1; }
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_stroke.py tests/test_gpu_parse.py -m gpu -x -q -k "not random_grammar and not garbage and not glyph_batch_from_text" > gpurun_out/r2ab_memcheck_stroke_parse.log 2>&1
echo "memcheck stroke/parse rc=$?"; tail -4 gpurun_out/r2ab_memcheck_stroke_parse.log | cut -c1-200
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_batch_api.py -m gpu -x -q -k "small_canvas or upload_batch or fill_batch_host" > gpurun_out/r2ab_memcheck_glyph.log 2>&1
echo "memcheck glyph rc=$?"; tail -4 gpurun_out/r2ab_memcheck_glyph.log | cut -c1-200
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_batch_api.py -m gpu -x -q -k "small_canvas_rare or upload_batch" > gpurun_out/r2ab_racecheck_glyph.log 2>&1
echo "racecheck glyph rc=$?"; tail -4 gpurun_out/r2ab_racecheck_glyph.log | cut -c1-200
