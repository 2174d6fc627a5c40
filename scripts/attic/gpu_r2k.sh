#!/bin/bash
# round-2 GPU pass k: the adapter's direct MDS export (struct mds arrays instead of per-entity apf calls): parity suite + wall-clock of the five sweeps
mkdir -p gpurun_out
python -c "import core_b200._lib as l; l.lib(); print('libmag ok')" || exit 1
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2k_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2k_tests.log
tail -12 gpurun_out/r2k_tests.log
nproc; lscpu | grep "Model name"
{
echo "== direct, 1 thread"; timeout 600 python scripts/adapter_run.py 48 1
echo "== direct, 8 threads"; timeout 600 python scripts/adapter_run.py 48 8 2
echo "== public API route, 1 thread"; MAG_ADAPTER_PUBLIC_API_ONLY=1 timeout 600 python scripts/adapter_run.py 48 1 2
} > gpurun_out/r2k_adapter48.log 2>&1
cat gpurun_out/r2k_adapter48.log
