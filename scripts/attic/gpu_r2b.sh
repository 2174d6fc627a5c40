#!/bin/bash
# round-2 GPU pass b: parity suite, legacy tile kernels vs anchor-row kernels, occupancy variants, ncu --set full of the row kernels
mkdir -p gpurun_out
python -c "import core_b200._lib as l; l.lib(); print('libmag ok')" || exit 1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2b_gpu.txt
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2b_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_tests.log
tail -15 gpurun_out/r2b_tests.log
B="python bench.py --steps 10 --warmup 3 --no-cpu --e2e-steps 0"
S='import sys,json; d=json.loads(sys.stdin.read()); print(sys.argv[1], round(d["ms_per_step"],3), {k:round(x,3) for k,x in d["roofline"]["kernel_ms_all"].items()}, d["stats"]["n_split"], d["stats"]["n_collapse"], d["stats"]["n_bad"], d["stats"]["n_near_threshold"])'
run() { name=$1; shift; "$@" 2> gpurun_out/r2b_err_$name.log | tail -1 | python -c "$S" $name; }
{
MAG_LEGACY_SWEEP=1 run legacy $B
MAG_LEGACY_SWEEP=1 run legacy_jit $B --jitter 0.2
run rows $B
run rows_jit $B --jitter 0.2
run rows_strict $B --fp strict
run rows_logm $B --field logm
run rows_logm_jit $B --field logm --jitter 0.2
} > gpurun_out/r2b_bench.log 2>&1
cat gpurun_out/r2b_bench.log
V="e2x256 e2x256pf e4x128_t4x128 e5x128_t5x128 ticket"
scripts/run_variants.sh $V > gpurun_out/r2b_variants.log 2>&1
scripts/run_variants.sh --jitter 0.2 $V > gpurun_out/r2b_variants_jit.log 2>&1
cat gpurun_out/r2b_variants.log gpurun_out/r2b_variants_jit.log
ncu --set full --clock-control none --import-source on -k regex:'k_edge_rows|k_tet_rows' -c 2 -o gpurun_out/r2b_full_jit -f \
    python bench.py --steps 1 --warmup 0 --no-cpu --e2e-steps 0 --jitter 0.2 > gpurun_out/r2b_ncu_jit.log 2>&1
tail -3 gpurun_out/r2b_ncu_jit.log
ncu --set full --clock-control none --import-source on -k regex:'k_edge_rows|k_tet_rows' -c 2 -o gpurun_out/r2b_full -f \
    python bench.py --steps 1 --warmup 0 --no-cpu --e2e-steps 0 > gpurun_out/r2b_ncu.log 2>&1
tail -3 gpurun_out/r2b_ncu.log
cp core_b200/lib/libmag.so /tmp/libmag_base.so; cp core_b200/lib_var/e2x256/libmag.so core_b200/lib/libmag.so
ncu --set full --clock-control none --import-source on -k regex:'k_edge_rows' -c 1 -o gpurun_out/r2b_full_e2x256_jit -f \
    python bench.py --steps 1 --warmup 0 --no-cpu --e2e-steps 0 --jitter 0.2 > gpurun_out/r2b_ncu_e2.log 2>&1
cp /tmp/libmag_base.so core_b200/lib/libmag.so
tail -3 gpurun_out/r2b_ncu_e2.log
