#!/bin/bash
# round-2 multi-GPU pass: the driver's own command line at N GPUs (N = first argument)
N=$1
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/r2end_gpus_g$N.txt
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu --e2e-steps 3 > gpurun_out/r2end_bench_g$N.json 2> gpurun_out/r2end_bench_g$N.err; echo "bench rc=$?"
tail -c 300 gpurun_out/r2end_bench_g$N.err
python - $N <<'PY'
import json, sys
N = sys.argv[1]
d = json.loads(open('gpurun_out/r2end_bench_g%s.json' % N).read().strip().splitlines()[-1])
print({k: d.get(k) for k in ('value', 'ms_per_step', 'parity_multi', 'parity_symmetry', 'multi_part_overhead')})
print('e2e', {k: d['e2e'].get(k) for k in ('value', 'ms_per_step')})
print({k: (round(v.get('ms_per_step'), 3)) for k, v in d.get('extra', {}).items() if 'ms_per_step' in v})
PY
