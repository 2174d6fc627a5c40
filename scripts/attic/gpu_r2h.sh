#!/bin/bash
# round-2 GPU pass h: stream kernels with cp.async-staged headers (edges and tets): parity, timing, ncu
mkdir -p gpurun_out
python -c "import core_b200._lib as l; l.lib(); print('libmag ok')" || exit 1
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "lean or golden or random_box or hub or full_size or near_threshold or partition or baseline" > gpurun_out/r2h_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2h_tests.log
tail -12 gpurun_out/r2h_tests.log
B="timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --e2e-steps 0 --no-extras"
S='import sys,json; d=json.loads(sys.stdin.read()); print(sys.argv[1], round(d["ms_per_step"],3), {k:round(x,3) for k,x in d["roofline"]["kernel_ms_all"].items()}, d["stats"]["n_split"], d["stats"]["n_collapse"], d["stats"]["n_bad"], d["stats"]["n_near_threshold"])'
run() { name=$1; shift; "$@" 2> gpurun_out/r2h_err_$name.log | tail -1 | python -c "$S" $name; }
{
run stream $B
run stream_jit $B --jitter 0.2
cp core_b200/lib/libmag.so /tmp/libmag_base.so
for v in t3x128 e3x128_t3x128 t2x192; do
  cp core_b200/lib_var/$v/libmag.so core_b200/lib/libmag.so
  run ${v} $B
  run ${v}_jit $B --jitter 0.2
done
} > gpurun_out/r2h_bench.log 2>&1
cat gpurun_out/r2h_bench.log
cp core_b200/lib_var/t3x128/libmag.so core_b200/lib/libmag.so
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_edge_rows_z|k_tet_rows_z' -c 2 -o gpurun_out/r2h_full_jit -f \
    python bench.py --steps 1 --warmup 0 --no-cpu --e2e-steps 0 --no-extras --jitter 0.2 > gpurun_out/r2h_ncu_jit.log 2>&1
cp /tmp/libmag_base.so core_b200/lib/libmag.so
tail -2 gpurun_out/r2h_ncu_jit.log
