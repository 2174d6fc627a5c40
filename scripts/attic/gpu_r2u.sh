#!/bin/bash
# round-2 GPU pass u: slot-word prefetch distance for the edge kernel, L2 prefetch of the other-end records
mkdir -p gpurun_out
B="timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --e2e-steps 0 --no-extras"
S='import sys,json; d=json.loads(sys.stdin.read()); print(sys.argv[1], round(d["ms_per_step"],3), {k:round(x,3) for k,x in d["roofline"]["kernel_ms_all"].items()}, d["stats"]["n_split"], d["stats"]["n_collapse"], d["stats"]["n_bad"], d["stats"]["n_near_threshold"])'
run() { name=$1; shift; "$@" 2> gpurun_out/r2u_err_$name.log | tail -1 | python -c "$S" $name; }
{
run e3 $B
run e3_jit $B --jitter 0.2
cp core_b200/lib/libmag.so /tmp/libmag_base.so
for v in e2 e4 e3rp e4rp e3t2; do
  cp core_b200/lib_var/$v/libmag.so core_b200/lib/libmag.so
  run ${v} $B
  run ${v}_jit $B --jitter 0.2
done
cp /tmp/libmag_base.so core_b200/lib/libmag.so
} > gpurun_out/r2u_bench.log 2>&1
cat gpurun_out/r2u_bench.log
