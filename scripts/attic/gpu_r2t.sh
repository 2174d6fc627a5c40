#!/bin/bash
# round-2 GPU pass t: lean kernels with the ticket drawn by a PTX atomic (no compiler aggregation) and the slot-word stream prefetched into L2
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "lean or golden or random_box or hub or full_size or near_threshold or partition or baseline or listed" > gpurun_out/r2t_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2t_tests.log
tail -4 gpurun_out/r2t_tests.log
B="timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --e2e-steps 0 --no-extras"
S='import sys,json; d=json.loads(sys.stdin.read()); print(sys.argv[1], round(d["ms_per_step"],3), {k:round(x,3) for k,x in d["roofline"]["kernel_ms_all"].items()}, d["stats"]["n_split"], d["stats"]["n_collapse"], d["stats"]["n_bad"], d["stats"]["n_near_threshold"])'
run() { name=$1; shift; "$@" 2> gpurun_out/r2t_err_$name.log | tail -1 | python -c "$S" $name; }
{
run tk1_pf6 $B
run tk1_pf6_jit $B --jitter 0.2
cp core_b200/lib/libmag.so /tmp/libmag_base.so
for v in tk0_pf0 tk1_pf0 tk1_pf3 tk1_pf10; do
  cp core_b200/lib_var/$v/libmag.so core_b200/lib/libmag.so
  run ${v} $B
  run ${v}_jit $B --jitter 0.2
done
cp /tmp/libmag_base.so core_b200/lib/libmag.so
} > gpurun_out/r2t_bench.log 2>&1
cat gpurun_out/r2t_bench.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_edge_rows_z|k_tet_rows_z' -c 2 -o gpurun_out/r2t_full_jit -f \
    python bench.py --steps 1 --warmup 0 --no-cpu --e2e-steps 0 --no-extras --jitter 0.2 > gpurun_out/r2t_ncu_jit.log 2>&1
tail -2 gpurun_out/r2t_ncu_jit.log
