#!/bin/bash
# round-2: ticket group sizes of the final lean kernels
mkdir -p gpurun_out
B="timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --e2e-steps 0 --no-extras"
S='import sys,json; d=json.loads(sys.stdin.read()); print(sys.argv[1], round(d["ms_per_step"],3), {k:round(x,3) for k,x in d["roofline"]["kernel_ms_all"].items()})'
run() { name=$1; shift; "$@" 2>/dev/null | tail -1 | python -c "$S" $name; }
{
cp core_b200/lib/libmag.so /tmp/libmag_base.so
for v in ez1 ez3 ez4 tz2 tz6 tz8; do
  cp core_b200/lib_var/$v/libmag.so core_b200/lib/libmag.so
  run ${v} $B
  run ${v}_jit $B --jitter 0.2
done
cp /tmp/libmag_base.so core_b200/lib/libmag.so
} > gpurun_out/r2grp_bench.log 2>&1
cat gpurun_out/r2grp_bench.log
