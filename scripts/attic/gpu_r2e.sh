#!/bin/bash
# round-2 GPU pass e: full parity suite incl. the adapter; fp64 issue-rate microbenchmark; tet kernel with pipelined dependent gathers; kernel family A/B; ncu
mkdir -p gpurun_out
python -c "import core_b200._lib as l; l.lib(); print('libmag ok')" || exit 1
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2e_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2e_tests.log
tail -12 gpurun_out/r2e_tests.log
scripts/microbench/fp64_rate > gpurun_out/r2e_fp64_rate.txt 2>&1; cat gpurun_out/r2e_fp64_rate.txt
B="python bench.py --steps 10 --warmup 3 --no-cpu --e2e-steps 0 --no-extras"
S='import sys,json; d=json.loads(sys.stdin.read()); print(sys.argv[1], round(d["ms_per_step"],3), {k:round(x,3) for k,x in d["roofline"]["kernel_ms_all"].items()}, d["stats"]["n_split"], d["stats"]["n_collapse"], d["stats"]["n_bad"], d["stats"]["n_near_threshold"])'
run() { name=$1; shift; "$@" 2> gpurun_out/r2e_err_$name.log | tail -1 | python -c "$S" $name; }
{
run lean $B
run lean_jit $B --jitter 0.2
run strict_tiles $B --fp strict
MAG_GENERAL_ROWS=1 run strict_rows $B --fp strict
run logm_tiles $B --field logm
run logm_tiles_jit $B --field logm --jitter 0.2
cp core_b200/lib/libmag.so /tmp/libmag_base.so
for v in ez_areg_3x128 ez_3x128 tz2_2x128 tz2_4x128; do
  cp core_b200/lib_var/$v/libmag.so core_b200/lib/libmag.so
  run ${v} $B
  run ${v}_jit $B --jitter 0.2
done
cp /tmp/libmag_base.so core_b200/lib/libmag.so
} > gpurun_out/r2e_bench.log 2>&1
cat gpurun_out/r2e_bench.log
ncu --set full --clock-control none --import-source on -k regex:'k_edge_rows|k_tet_rows' -c 2 -o gpurun_out/r2e_full_jit -f \
    python bench.py --steps 1 --warmup 0 --no-cpu --e2e-steps 0 --no-extras --jitter 0.2 > gpurun_out/r2e_ncu_jit.log 2>&1
tail -2 gpurun_out/r2e_ncu_jit.log
ncu --set full --clock-control none --import-source on -k regex:'k_edge_rows|k_tet_rows' -c 2 -o gpurun_out/r2e_full -f \
    python bench.py --steps 1 --warmup 0 --no-cpu --e2e-steps 0 --no-extras > gpurun_out/r2e_ncu.log 2>&1
tail -2 gpurun_out/r2e_ncu.log
