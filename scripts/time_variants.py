"""One process, one n = 203 mesh: step time and per-kernel times (CUDA events inside mag_sweep) of the bench's variants --
lattice / jittered x AnisoSizeField fast / strict / LogAniso fast.  Usage: time_variants.py [n] [steps]"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import core_b200 as cb

n = int(sys.argv[1]) if len(sys.argv) > 1 else 203
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
tag = sys.argv[3] if len(sys.argv) > 3 else "variants"
ops = cb.OP_LENGTHS | cb.OP_MARK_SPLIT | cb.OP_MARK_COLLAPSE | cb.OP_QUALITIES | cb.OP_MARK_BAD
xyz0, ev, tv = cb.boxmesh.kuhn_box(n, n, n)
hbar = 1.0 / n
stream = torch.cuda.Stream()
out = {}
p = cb.Part(0)
p.set_stream(stream.cuda_stream)
for jit in (0.0, 0.2):
    xyz = cb.fields.jitter(xyz0, jit * hbar) if jit > 0 else xyz0
    h, R = cb.fields.shock_rotating(xyz, hbar)
    p.set_mesh(xyz, ev, tv)
    for field, mode in (("aniso", cb.FP_FAST), ("aniso", cb.FP_STRICT), ("logm", cb.FP_FAST)):
        if field == "aniso":
            p.set_size_field_aniso(h, R)
        else:
            p.set_size_field_logm_from_frames(h, R, 0)
        p.synchronize()
        def step():
            p.clear_flags()
            p.sweep(ops, fp_mode=mode)
            return p.stats()
        for _ in range(3):
            st = step()
        p.timing_begin(steps)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        with torch.cuda.stream(stream):
            e0.record(stream)
            for _ in range(steps):
                st = step()
            e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        kt = p.timing_read()
        key = "jit%.1f_%s_%s" % (jit, field, "fast" if mode == cb.FP_FAST else "strict")
        out[key] = {"ms_per_step": ms, "edges_ms": float(kt[:, 1].mean()), "elements_ms": float(kt[:, 2].mean()),
                    "counts": [st[k] for k in ("n_split", "n_collapse", "n_bad", "n_near_threshold")],
                    "min_quality": st["min_quality"], "max_length": st["max_length"]}
        print(key, json.dumps(out[key]), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/%s.json" % tag, "w"), indent=1)
