#!/bin/bash
# on the GPU box (gpurun): the three captures a round's profile consists of.  usage: profile_run.sh <tag>
#   1. ncu --set full of one step of the default bench (top kernels; numbers printed under ncu are never bench values)
#   2. the launch list of a short bench (gpu__time_duration.sum per launch)
#   3. the bench itself, unprofiled (+ the jittered variant)
# scripts/profile_summary.py <tag> then turns gpurun_out/<tag>_* into profiles/<tag>_* and profiles/traffic.json.
T=$1
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:'k_edges|k_tets|k_vertex_pass' -c 3 -o gpurun_out/${T}_full -f \
    python bench.py --steps 1 --warmup 0 --no-cpu --e2e-steps 0 > gpurun_out/${T}_ncu.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --e2e-steps 0 > gpurun_out/${T}_l.log 2>&1
python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
python bench.py --jitter 0.2 --no-cpu --e2e-steps 0 > gpurun_out/${T}_bench_jit.json 2>> gpurun_out/${T}_bench.err
tail -c 400 gpurun_out/${T}_ncu.log; tail -1 gpurun_out/${T}_bench.json | cut -c1-400
