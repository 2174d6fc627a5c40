#!/bin/bash
# round-2 GPU pass y: next-slice headers loaded without the early R2UR
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "lean or golden or random_box or hub or full_size or near_threshold or partition or baseline or listed" > gpurun_out/r2y_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2y_tests.log
tail -4 gpurun_out/r2y_tests.log
B="timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --e2e-steps 0 --no-extras"
S='import sys,json; d=json.loads(sys.stdin.read()); print(sys.argv[1], round(d["ms_per_step"],3), {k:round(x,3) for k,x in d["roofline"]["kernel_ms_all"].items()}, d["stats"]["n_split"], d["stats"]["n_collapse"], d["stats"]["n_bad"], d["stats"]["n_near_threshold"])'
run() { name=$1; shift; "$@" 2> gpurun_out/r2y_err_$name.log | tail -1 | python -c "$S" $name; }
{
run deep $B
run deep_jit $B --jitter 0.2
cp core_b200/lib/libmag.so /tmp/libmag_base.so
cp core_b200/lib_var/t6/libmag.so core_b200/lib/libmag.so
run deep_t6 $B
run deep_t6_jit $B --jitter 0.2
cp /tmp/libmag_base.so core_b200/lib/libmag.so
} > gpurun_out/r2y_bench.log 2>&1
cat gpurun_out/r2y_bench.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_tet_rows_w|k_edge_rows_z' -c 2 -o gpurun_out/r2y_full_jit -f \
    python bench.py --steps 1 --warmup 0 --no-cpu --e2e-steps 0 --no-extras --jitter 0.2 > gpurun_out/r2y_ncu_jit.log 2>&1
tail -2 gpurun_out/r2y_ncu_jit.log
