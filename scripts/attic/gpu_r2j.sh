#!/bin/bash
# round-2 GPU pass j (8 GPUs): the driver's own command line at N = 8
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/r2j_gpus.txt
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu > gpurun_out/r2j_bench_g8.json 2> gpurun_out/r2j_bench_g8.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r2j_bench_g8.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2j_bench_g8.json').read().strip().splitlines()[-1])
print({k: d.get(k) for k in ('value', 'ms_per_step', 'parity_multi', 'parity_symmetry', 'per_part_owned_counts', 'multi_part_overhead')})
print(d.get('parity_multi_detail'))
print('e2e', d.get('e2e'))
print({k: (v.get('ms_per_step'), v.get('stats')) for k, v in d.get('extra', {}).items()})
PY
