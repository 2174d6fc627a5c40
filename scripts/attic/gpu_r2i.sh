#!/bin/bash
# round-2 GPU pass i (2 GPUs): the NCCL parity tests, then the driver's own command line at N = 2 (parity_multi, mixed extra, multi-part overhead)
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/r2i_gpus.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "nccl" > gpurun_out/r2i_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2i_tests.log
tail -6 gpurun_out/r2i_tests.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu > gpurun_out/r2i_bench_g2.json 2> gpurun_out/r2i_bench_g2.err; echo "bench rc=$?"
tail -c 800 gpurun_out/r2i_bench_g2.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2i_bench_g2.json').read().strip().splitlines()[-1])
print({k: d.get(k) for k in ('value', 'ms_per_step', 'parity_multi', 'parity_symmetry', 'per_part_owned_counts', 'multi_part_overhead')})
print(d.get('parity_multi_detail'))
print('e2e', d.get('e2e'))
print({k: (v.get('ms_per_step'), v.get('stats')) for k, v in d.get('extra', {}).items()})
PY
