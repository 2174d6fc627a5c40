#!/bin/bash
# round-2 sanitizer pass over every kernel family incl. the round-2 kernels (k_tet_rows_w / cp.async stage, k_tet_winners, k_collapse_quality, k_prism_weights, v2t build)
mkdir -p gpurun_out
timeout 300 python scripts/sanitize_small.py > gpurun_out/r2san_plain.log 2>&1; echo "plain rc=$?"; tail -1 gpurun_out/r2san_plain.log
for tool in memcheck racecheck initcheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 7 python scripts/sanitize_small.py > gpurun_out/r2san_$tool.log 2>&1; echo "$tool rc=$?"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard" gpurun_out/r2san_$tool.log | tail -3
done
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/r2san_tests.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2san_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
