#!/bin/bash
# round-2 GPU pass s: MAG_FP_FAST_LISTED mode; full suite; bench lines
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/r2s_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2s_tests.log
tail -6 gpurun_out/r2s_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2s_smoke.log 2>&1; tail -2 gpurun_out/r2s_smoke.log
timeout 900 python bench.py > gpurun_out/r2s_default_bench.json 2> gpurun_out/r2s_default_bench.err; echo "default bench rc=$?"; tail -c 400 gpurun_out/r2s_default_bench.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2s_default_bench.json').read().strip().splitlines()[-1])
print({k: d.get(k) for k in ('value', 'ms_per_step', 'gpu_launches')})
print('roofline', {k: d['roofline'].get(k) for k in ('kernel', 'frac', 'step_frac', 'traffic', 'kernel_ms_all')})
print('config', {k: d['config'].get(k) for k in ('export_s', 'field_upload_ms', 'mesh_generation_s')})
print('e2e', d.get('e2e'))
print('cpu', d.get('cpu_baseline'))
print({k: (round(v.get('ms_per_step'), 3), v.get('kernel_ms_all')) for k, v in d.get('extra', {}).items() if 'ms_per_step' in v})
PY
