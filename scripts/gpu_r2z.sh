#!/bin/bash
# round-2 GPU pass z: the round's final profile (ncu --set full of the two sweep kernels, lattice and jittered; launch list) + the default bench lines
T=r2z
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/${T}_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_tests.log
tail -4 gpurun_out/${T}_tests.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_edge_rows_z|k_tet_rows_w' -c 2 -o gpurun_out/${T}_full -f \
    python bench.py --steps 1 --warmup 0 --no-cpu --e2e-steps 0 --no-extras > gpurun_out/${T}_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_edge_rows_z|k_tet_rows_w' -c 2 -o gpurun_out/${T}_full_jit -f \
    python bench.py --steps 1 --warmup 0 --no-cpu --e2e-steps 0 --no-extras --jitter 0.2 > gpurun_out/${T}_ncu_jit.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --e2e-steps 0 --no-extras > gpurun_out/${T}_l.log 2>&1
timeout 900 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_ref_bench.json 2> gpurun_out/${T}_ref_bench.err; echo "ref rc=$?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2z_bench.json').read().strip().splitlines()[-1])
print({k: d.get(k) for k in ('value', 'ms_per_step', 'gpu_launches')})
print('roofline', {k: d['roofline'].get(k) for k in ('kernel', 'frac', 'step_frac', 'traffic', 'kernel_ms_all')})
print('config', {k: d['config'].get(k) for k in ('export_s', 'field_upload_ms')})
print('e2e', {k: d['e2e'].get(k) for k in ('value', 'ms_per_step')})
print({k: (round(v.get('ms_per_step'), 3), v.get('kernel_ms_all')) for k, v in d.get('extra', {}).items() if 'ms_per_step' in v})
r = json.loads(open('gpurun_out/r2z_ref_bench.json').read().strip().splitlines()[-1])
print('ref', r.get('value'), r.get('cpu_baseline', {}).get('cores'))
PY
