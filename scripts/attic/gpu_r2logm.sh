#!/bin/bash
# round-2: ncu --set full of the log-Euclidean edge kernel (the default interpolation of ma::configure) and of the strict kernels, final build
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_edges|k_tet_rows_w' -c 2 -o gpurun_out/r2z_logm_full -f \
    python bench.py --steps 1 --warmup 0 --no-cpu --e2e-steps 0 --no-extras --field logm --jitter 0.2 > gpurun_out/r2z_logm_ncu.log 2>&1
tail -2 gpurun_out/r2z_logm_ncu.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_edges|k_tets' -c 2 -o gpurun_out/r2z_strict_full -f \
    python bench.py --steps 1 --warmup 0 --no-cpu --e2e-steps 0 --no-extras --fp strict > gpurun_out/r2z_strict_ncu.log 2>&1
tail -2 gpurun_out/r2z_strict_ncu.log
