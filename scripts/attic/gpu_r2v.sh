#!/bin/bash
# round-2 GPU pass v: tet rows with the max-Jacobian vertex in the slot word (k_tet_winners + k_tet_rows_w, transform staged by cp.async)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "lean or golden or random_box or hub or full_size or near_threshold or partition or baseline or listed or row_layout or coords" > gpurun_out/r2v_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2v_tests.log
tail -4 gpurun_out/r2v_tests.log
B="timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --e2e-steps 0 --no-extras"
S='import sys,json; d=json.loads(sys.stdin.read()); print(sys.argv[1], round(d["ms_per_step"],3), {k:round(x,3) for k,x in d["roofline"]["kernel_ms_all"].items()}, d["stats"]["n_split"], d["stats"]["n_collapse"], d["stats"]["n_bad"], d["stats"]["n_near_threshold"])'
run() { name=$1; shift; "$@" 2> gpurun_out/r2v_err_$name.log | tail -1 | python -c "$S" $name; }
{
run win $B
run win_jit $B --jitter 0.2
MAG_TET_WINNER=0 run nowin $B
MAG_TET_WINNER=0 run nowin_jit $B --jitter 0.2
run win_mixed $B --workload mixed --n 120
} > gpurun_out/r2v_bench.log 2>&1
cat gpurun_out/r2v_bench.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_tet_rows_w' -c 1 -o gpurun_out/r2v_full_jit -f \
    python bench.py --steps 1 --warmup 0 --no-cpu --e2e-steps 0 --no-extras --jitter 0.2 > gpurun_out/r2v_ncu_jit.log 2>&1
tail -2 gpurun_out/r2v_ncu_jit.log
