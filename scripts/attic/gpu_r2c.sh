#!/bin/bash
# round-2 GPU pass c: parity suite; L2 bulk prefetch (UBLKPF) on/off for the tile kernels and the anchor-row kernels (ticket hand-out); occupancy variants
mkdir -p gpurun_out
python -c "import core_b200._lib as l; l.lib(); print('libmag ok')" || exit 1
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2c_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_tests.log
tail -5 gpurun_out/r2c_tests.log
B="python bench.py --steps 10 --warmup 3 --no-cpu --e2e-steps 0"
S='import sys,json; d=json.loads(sys.stdin.read()); print(sys.argv[1], round(d["ms_per_step"],3), {k:round(x,3) for k,x in d["roofline"]["kernel_ms_all"].items()}, d["stats"]["n_split"], d["stats"]["n_collapse"], d["stats"]["n_bad"], d["stats"]["n_near_threshold"])'
run() { name=$1; shift; "$@" 2> gpurun_out/r2c_err_$name.log | tail -1 | python -c "$S" $name; }
{
MAG_LEGACY_SWEEP=1 MAG_L2_PREFETCH=0 run legacy $B
MAG_LEGACY_SWEEP=1 run legacy_pf $B
MAG_LEGACY_SWEEP=1 MAG_L2_PREFETCH=0 run legacy_jit $B --jitter 0.2
MAG_LEGACY_SWEEP=1 run legacy_pf_jit $B --jitter 0.2
MAG_L2_PREFETCH=0 run rows $B
run rows_pf $B
MAG_L2_PREFETCH=0 run rows_jit $B --jitter 0.2
run rows_pf_jit $B --jitter 0.2
cp core_b200/lib/libmag.so /tmp/libmag_base.so
for v in r2 r2pf r5x128 r2pf_t3x192; do
  cp core_b200/lib_var/$v/libmag.so core_b200/lib/libmag.so
  run ${v}_pf $B
  run ${v}_pf_jit $B --jitter 0.2
done
cp /tmp/libmag_base.so core_b200/lib/libmag.so
} > gpurun_out/r2c_bench.log 2>&1
cat gpurun_out/r2c_bench.log
ncu --set full --clock-control none --import-source on -k regex:'k_edge_rows|k_tet_rows' -c 2 -o gpurun_out/r2c_full_jit -f \
    python bench.py --steps 1 --warmup 0 --no-cpu --e2e-steps 0 --jitter 0.2 > gpurun_out/r2c_ncu_jit.log 2>&1
tail -2 gpurun_out/r2c_ncu_jit.log
MAG_LEGACY_SWEEP=1 ncu --set full --clock-control none --import-source on -k regex:'k_edges|k_tets' -c 2 -o gpurun_out/r2c_full_legacy_jit -f \
    python bench.py --steps 1 --warmup 0 --no-cpu --e2e-steps 0 --jitter 0.2 > gpurun_out/r2c_ncu_legacy_jit.log 2>&1
tail -2 gpurun_out/r2c_ncu_legacy_jit.log
