#!/bin/bash
mkdir -p gpurun_out
for tool in memcheck initcheck; do
  timeout 100 compute-sanitizer --tool $tool --error-exitcode 7 python scripts/sanitize_vqu.py > gpurun_out/r2end_san_$tool.log 2>&1; echo "$tool rc=$?"
  grep -E "ERROR SUMMARY|sanitize_vqu ok" gpurun_out/r2end_san_$tool.log | tail -2
done
