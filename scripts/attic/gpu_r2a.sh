#!/bin/bash
# round-2 first GPU pass: parity suite, then the anchor-row kernels against the round-1 tile kernels and the occupancy variants
mkdir -p gpurun_out
python -c "import core_b200._lib as l; l.lib(); print('libmag ok')" || exit 1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_gpu.txt
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2a_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_tests.log
tail -15 gpurun_out/r2a_tests.log
B="python bench.py --steps 10 --warmup 3 --no-cpu --e2e-steps 0"
S='import sys,json; d=json.loads(sys.stdin.read()); print(sys.argv[1], round(d["ms_per_step"],3), {k:round(x,3) for k,x in d["roofline"]["kernel_ms_all"].items()}, d["stats"]["n_split"], d["stats"]["n_collapse"], d["stats"]["n_bad"], d["stats"]["n_near_threshold"])'
{
MAG_LEGACY_SWEEP=1 $B | tail -1 | python -c "$S" legacy
MAG_LEGACY_SWEEP=1 $B --jitter 0.2 | tail -1 | python -c "$S" legacy_jit
$B | tail -1 | python -c "$S" rows
$B --jitter 0.2 | tail -1 | python -c "$S" rows_jit
$B --fp strict | tail -1 | python -c "$S" rows_strict
$B --field logm | tail -1 | python -c "$S" rows_logm
$B --field logm --jitter 0.2 | tail -1 | python -c "$S" rows_logm_jit
} > gpurun_out/r2a_bench.log 2>&1
cat gpurun_out/r2a_bench.log
V="ticket e2x256_t2x320 e2x256pf_t3x192 e5x128_t4x128 e4x128pf e3x256pf"
scripts/run_variants.sh $V > gpurun_out/r2a_variants.log 2>&1
scripts/run_variants.sh --jitter 0.2 $V > gpurun_out/r2a_variants_jit.log 2>&1
cat gpurun_out/r2a_variants.log gpurun_out/r2a_variants_jit.log
ncu --set full --clock-control none --import-source on -k regex:'k_edge_rows|k_tet_rows' -c 2 -o gpurun_out/r2a_full -f \
    python bench.py --steps 1 --warmup 0 --no-cpu --e2e-steps 0 > gpurun_out/r2a_ncu.log 2>&1
tail -3 gpurun_out/r2a_ncu.log
cp core_b200/lib/libmag.so /tmp/libmag_base.so; cp core_b200/lib_var/e2x256pf_t3x192/libmag.so core_b200/lib/libmag.so
ncu --set full --clock-control none --import-source on -k regex:'k_edge_rows|k_tet_rows' -c 2 -o gpurun_out/r2a_full_pf -f \
    python bench.py --steps 1 --warmup 0 --no-cpu --e2e-steps 0 --jitter 0.2 > gpurun_out/r2a_ncu_pf.log 2>&1
cp /tmp/libmag_base.so core_b200/lib/libmag.so
tail -3 gpurun_out/r2a_ncu_pf.log
