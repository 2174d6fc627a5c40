#!/bin/bash
# builds one kernel variant HERE (no GPU needed) into core_b200/lib_var/<name>/libmag.so.  usage: build_variant.sh <name> "<-D flags>"
# run on the GPU box with scripts/run_variants.sh (copies each variant over core_b200/lib/libmag.so, benches, restores)
set -e
cd "$(dirname "$0")/.."
make -s -C core_b200/csrc -j8 OUT=../lib_var/$1 EXTRA="$2" >/dev/null
python scripts/regs.py core_b200/lib_var/$1/mag_kernels.ptxas.log 2>/dev/null | grep "rows_z<2>" || true
