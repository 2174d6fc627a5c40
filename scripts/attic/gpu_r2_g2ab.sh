for w in 1 0; do
MAG_TET_WINNER=$w timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu --no-extras --e2e-steps 0 2>/dev/null | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('winner=$w', round(d['ms_per_step'],3), d['roofline']['kernel_ms_all'], d['roofline']['kernel'])"
done
