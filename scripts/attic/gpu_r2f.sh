#!/bin/bash
# round-2 GPU pass f: full parity suite on the final kernel selection (lean rows for the full fast sweep, tiles elsewhere), bench lines, launch list
mkdir -p gpurun_out
python -c "import core_b200._lib as l; l.lib(); print('libmag ok')" || exit 1
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2f_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f_tests.log
tail -12 gpurun_out/r2f_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2f_smoke.log 2>&1; tail -2 gpurun_out/r2f_smoke.log
B="python bench.py --steps 10 --warmup 3 --no-cpu --e2e-steps 0 --no-extras"
S='import sys,json; d=json.loads(sys.stdin.read()); print(sys.argv[1], round(d["ms_per_step"],3), {k:round(x,3) for k,x in d["roofline"]["kernel_ms_all"].items()}, d["stats"]["n_split"], d["stats"]["n_collapse"], d["stats"]["n_bad"], d["stats"]["n_near_threshold"])'
run() { name=$1; shift; "$@" 2> gpurun_out/r2f_err_$name.log | tail -1 | python -c "$S" $name; }
{
run lean $B
run lean_jit $B --jitter 0.2
MAG_LEGACY_SWEEP=1 run tiles $B
run mixed $B --workload mixed --n 120
} > gpurun_out/r2f_bench.log 2>&1
cat gpurun_out/r2f_bench.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2f_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --e2e-steps 0 --no-extras > gpurun_out/r2f_l.log 2>&1
tail -2 gpurun_out/r2f_l.log
ncu --set full --clock-control none --import-source on -k regex:'k_edge_rows_z|k_tet_rows_z' -c 2 -o gpurun_out/r2f_full -f \
    python bench.py --steps 1 --warmup 0 --no-cpu --e2e-steps 0 --no-extras > gpurun_out/r2f_ncu.log 2>&1
tail -2 gpurun_out/r2f_ncu.log
ncu --set full --clock-control none --import-source on -k regex:'k_edge_rows_z|k_tet_rows_z' -c 2 -o gpurun_out/r2f_full_jit -f \
    python bench.py --steps 1 --warmup 0 --no-cpu --e2e-steps 0 --no-extras --jitter 0.2 > gpurun_out/r2f_ncu_jit.log 2>&1
tail -2 gpurun_out/r2f_ncu_jit.log
timeout 900 python bench.py > gpurun_out/r2f_default_bench.json 2> gpurun_out/r2f_default_bench.err; echo "default bench rc=$?"; tail -c 600 gpurun_out/r2f_default_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2f_ref_bench.json 2> gpurun_out/r2f_ref_bench.err; echo "ref rc=$?"; cut -c1-600 gpurun_out/r2f_ref_bench.json
timeout 600 python scripts/adapter_run.py 48 > gpurun_out/r2f_adapter48.log 2>&1; tail -5 gpurun_out/r2f_adapter48.log
