#!/bin/bash
# round-2 GPU pass n: prism weights (mag_prism_weights, adapter on a mixed mesh), direct flag write-back, adapter timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_adapter.py tests/test_gpu_parity.py -m gpu -q -x -k "adapter or weights or golden" 2>&1 | tail -5
{
echo "== direct, 1 thread"; timeout 600 python scripts/adapter_run.py 48 1 2
echo "== direct, 8 threads"; timeout 600 python scripts/adapter_run.py 48 8 1
} > gpurun_out/r2n_adapter48.log 2>&1
cat gpurun_out/r2n_adapter48.log
