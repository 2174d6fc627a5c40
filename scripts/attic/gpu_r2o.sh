#!/bin/bash
# round-2 GPU pass o: (1) which adapter test crashes, (2) stage trace of mag_set_mesh inside the adapter timing
mkdir -p gpurun_out
timeout 900 python -X faulthandler -m pytest tests/test_adapter.py -m gpu -q -x -v > gpurun_out/r2o_adapter_tests.log 2>&1; echo "rc=$?"
grep -n "PASSED\|FAILED\|Fatal\|Segmentation\|adapter_check\|File \"/root" gpurun_out/r2o_adapter_tests.log | head -40
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "weights or golden" 2>&1 | tail -3
MAG_TRACE=1 timeout 600 python scripts/adapter_run.py 48 1 2 > gpurun_out/r2o_adapter48.log 2>&1
cat gpurun_out/r2o_adapter48.log
