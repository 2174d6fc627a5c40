#!/bin/bash
# on the GPU box: bench every prebuilt variant under core_b200/lib_var/ (kernel times only).  usage: run_variants.sh [--jitter X] name ...
BARGS=""
if [ "$1" == "--jitter" ]; then BARGS="--jitter $2"; shift 2; fi
cp core_b200/lib/libmag.so /tmp/libmag_base.so
for v in "$@"; do
  if [ "$v" == "base" ]; then cp /tmp/libmag_base.so core_b200/lib/libmag.so; else cp core_b200/lib_var/$v/libmag.so core_b200/lib/libmag.so; fi
  python bench.py --steps 10 --warmup 3 --no-cpu --e2e-steps 0 $BARGS 2>&1 | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('$v', round(d['ms_per_step'],3), {k:round(x,3) for k,x in d['roofline']['kernel_ms_all'].items()}, d['stats']['n_split'], d['stats']['n_collapse'], d['stats']['n_bad'])"
done
cp /tmp/libmag_base.so core_b200/lib/libmag.so
