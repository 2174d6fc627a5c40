#!/bin/bash
# round-2 GPU pass p: export temporaries from the stream-ordered pool (was: 35 - 700 ms of cudaMalloc / cudaFree per export), full suite, adapter timing
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/r2p_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2p_tests.log
tail -6 gpurun_out/r2p_tests.log
{
echo "== direct, 1 thread"; MAG_TRACE=1 timeout 600 python scripts/adapter_run.py 48 1 1
echo "== direct, 1 thread"; timeout 600 python scripts/adapter_run.py 48 1
echo "== direct, 8 threads"; timeout 600 python scripts/adapter_run.py 48 8 2
} > gpurun_out/r2p_adapter48.log 2>&1
cat gpurun_out/r2p_adapter48.log
