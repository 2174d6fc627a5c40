#!/bin/bash
# round-2: hybrid allocation policy (pool for blocks under 64 MB, cudaMalloc above): export time of the benchmark part, adapter timing, suite
mkdir -p gpurun_out
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --e2e-steps 0 --no-extras 2>/dev/null | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('n=203 export_s', d['config']['export_s'], 'field_upload_ms', d['config']['field_upload_ms'], 'ms', d['ms_per_step'])"
timeout 600 python scripts/adapter_run.py 48 1 2 > gpurun_out/r2fin2_adapter48.log 2>&1; cat gpurun_out/r2fin2_adapter48.log
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/r2fin2_tests.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2fin2_tests.log
