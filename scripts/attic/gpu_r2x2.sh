#!/bin/bash
# round-2: both Gauss points of a LogAniso edge solved in lockstep (quad_expm_qr3_x2): bit-for-bit A/B against the sequential form, timing, parity tests
mkdir -p gpurun_out
python scripts/logm_ab.py /tmp/x2on.npy
cp core_b200/lib/libmag.so /tmp/libmag_base.so
cp core_b200/lib_var/x2off/libmag.so core_b200/lib/libmag.so
python scripts/logm_ab.py /tmp/x2off.npy
python -c "
import numpy as np
a, b = np.load('/tmp/x2on.npy'), np.load('/tmp/x2off.npy')
print('lockstep vs sequential: entries', len(a), 'differing', int(np.count_nonzero(a != b)), 'max rel', float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300))))"
B="timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --e2e-steps 0 --no-extras --field logm"
S='import sys,json; d=json.loads(sys.stdin.read()); print(sys.argv[1], round(d["ms_per_step"],3), {k:round(x,3) for k,x in d["roofline"]["kernel_ms_all"].items()}, d["stats"]["n_split"], d["stats"]["n_collapse"])'
run() { name=$1; shift; "$@" 2>/dev/null | tail -1 | python -c "$S" $name; }
run x2off $B
run x2off_jit $B --jitter 0.2
cp /tmp/libmag_base.so core_b200/lib/libmag.so
run x2on $B
run x2on_jit $B --jitter 0.2
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_adapter.py -m gpu -q -x -k "logm or log or golden or loganiso or adapter_reproduces or ma_adapt" 2>&1 | tail -3
