#!/bin/bash
# round-2 GPU pass l: adapter host-path profile (first use / warm re-export)
mkdir -p gpurun_out
python -m pytest tests/test_adapter.py -m gpu -q -x 2>&1 | tail -3
{
echo "== direct, 1 thread"; timeout 600 python scripts/adapter_run.py 48 1
echo "== direct, 8 threads"; timeout 600 python scripts/adapter_run.py 48 8 2
} > gpurun_out/r2l_adapter48.log 2>&1
cat gpurun_out/r2l_adapter48.log
