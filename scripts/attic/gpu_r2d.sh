#!/bin/bash
# round-2 GPU pass d: parity suite (incl. lean == general == tile kernels); lean kernels vs general rows; occupancy variants; ncu; full default bench line
mkdir -p gpurun_out
python -c "import core_b200._lib as l; l.lib(); print('libmag ok')" || exit 1
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2d_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2d_tests.log
tail -8 gpurun_out/r2d_tests.log
B="python bench.py --steps 10 --warmup 3 --no-cpu --e2e-steps 0 --no-extras"
S='import sys,json; d=json.loads(sys.stdin.read()); print(sys.argv[1], round(d["ms_per_step"],3), {k:round(x,3) for k,x in d["roofline"]["kernel_ms_all"].items()}, d["stats"]["n_split"], d["stats"]["n_collapse"], d["stats"]["n_bad"], d["stats"]["n_near_threshold"])'
run() { name=$1; shift; "$@" 2> gpurun_out/r2d_err_$name.log | tail -1 | python -c "$S" $name; }
{
run lean $B
run lean_jit $B --jitter 0.2
MAG_LEAN_SWEEP=0 run rows_jit $B --jitter 0.2
run lean_logm_jit $B --jitter 0.2 --field logm
run lean_strict $B --fp strict
cp core_b200/lib/libmag.so /tmp/libmag_base.so
for v in z_areg z_4x128 z_g2 z_5x128 z_3x128; do
  cp core_b200/lib_var/$v/libmag.so core_b200/lib/libmag.so
  run ${v} $B
  run ${v}_jit $B --jitter 0.2
done
cp /tmp/libmag_base.so core_b200/lib/libmag.so
} > gpurun_out/r2d_bench.log 2>&1
cat gpurun_out/r2d_bench.log
ncu --set full --clock-control none --import-source on -k regex:'k_edge_rows|k_tet_rows' -c 2 -o gpurun_out/r2d_full_jit -f \
    python bench.py --steps 1 --warmup 0 --no-cpu --e2e-steps 0 --no-extras --jitter 0.2 > gpurun_out/r2d_ncu_jit.log 2>&1
tail -2 gpurun_out/r2d_ncu_jit.log
ncu --set full --clock-control none --import-source on -k regex:'k_edge_rows|k_tet_rows' -c 2 -o gpurun_out/r2d_full -f \
    python bench.py --steps 1 --warmup 0 --no-cpu --e2e-steps 0 --no-extras > gpurun_out/r2d_ncu.log 2>&1
tail -2 gpurun_out/r2d_ncu.log
timeout 900 python bench.py > gpurun_out/r2d_default_bench.json 2> gpurun_out/r2d_default_bench.err; echo "default bench rc=$?"; tail -c 1500 gpurun_out/r2d_default_bench.err; cat gpurun_out/r2d_default_bench.json | cut -c1-3000
