#!/bin/bash
# round-2 end pass A: full GPU suite on the final build (per-vertex uniform-edge cache), smoke, default bench line
T=r2end
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/${T}_tests.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/${T}_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2end_bench.json').read().strip().splitlines()[-1])
print({k: d.get(k) for k in ('value', 'ms_per_step', 'gpu_launches')})
print('roofline', {k: d['roofline'].get(k) for k in ('kernel', 'frac', 'step_frac', 'traffic', 'kernel_ms_all')})
print('config', {k: d['config'].get(k) for k in ('export_s', 'field_upload_ms')})
print('e2e', {k: d['e2e'].get(k) for k in ('value', 'ms_per_step')}, 'cpu', d.get('cpu_baseline'))
print({k: (round(v.get('ms_per_step'), 3)) for k, v in d.get('extra', {}).items() if 'ms_per_step' in v})
PY
