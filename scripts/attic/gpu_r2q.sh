#!/bin/bash
# round-2 GPU pass q: collapse candidates (mag_collapse_quality + mag::collapseQualities against ma::Collapse), prism weights through the adapter
mkdir -p gpurun_out
timeout 900 python -X faulthandler -m pytest tests/test_adapter.py tests/test_gpu_parity.py -m gpu -q -x -k "collapse or weights or cavity or layer" > gpurun_out/r2q_tests.log 2>&1; echo "rc=$?"
tail -15 gpurun_out/r2q_tests.log
