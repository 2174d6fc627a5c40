#!/bin/bash
# builds kernel variants ON the GPU box and benches each (kernel times only).  usage: variants.sh [--jitter X] "flags" ...
BARGS=""
if [ "$1" == "--jitter" ]; then BARGS="--jitter $2"; shift 2; fi
for v in "$@"; do
  rm -f core_b200/lib/mag_*.o core_b200/lib/libmag.so; make -s -C core_b200/csrc -j8 EXTRA="$v" >/dev/null 2>&1 || { echo "BUILD FAILED $v"; continue; }
  python bench.py --steps 10 --warmup 3 --no-cpu --e2e-steps 0 $BARGS 2>&1 | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('$v', round(d['ms_per_step'],3), {k:round(x,3) for k,x in d['roofline']['kernel_ms_all'].items()}, d['stats']['n_split'], d['stats']['n_bad'])"
done
rm -f core_b200/lib/mag_*.o core_b200/lib/libmag.so; make -s -C core_b200/csrc -j8 >/dev/null 2>&1
