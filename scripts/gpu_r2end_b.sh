#!/bin/bash
# round-2 end pass B: ncu --set full of the two sweep kernels on the final build (lattice), launch list of a short bench
T=r2end
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_edge_rows_z|k_tet_rows_w' -c 2 -o gpurun_out/${T}_full -f \
    python bench.py --steps 1 --warmup 0 --no-cpu --e2e-steps 0 --no-extras > gpurun_out/${T}_ncu.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --e2e-steps 0 --no-extras > gpurun_out/${T}_l.log 2>&1
tail -c 300 gpurun_out/${T}_ncu.log; grep -c k_ gpurun_out/${T}_launches.csv
