"""r2co experiment: MAG_COSCHED = 0 / 1 / 2 on the n = 203 box, lattice and jittered; results of the three must be identical.
One process, one mesh build; every variant gets its own context (MAG_COSCHED is read at mag_create)."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import core_b200 as cb

n = int(sys.argv[1]) if len(sys.argv) > 1 else 203
steps = 10
ops = cb.OP_LENGTHS | cb.OP_MARK_SPLIT | cb.OP_MARK_COLLAPSE | cb.OP_QUALITIES | cb.OP_MARK_BAD
xyz0, ev, tv = cb.boxmesh.kuhn_box(n, n, n)
hbar = 1.0 / n
stream = torch.cuda.Stream()
out = {}
for jit in (0.0, 0.2):
    xyz = cb.fields.jitter(xyz0, jit * hbar) if jit > 0 else xyz0
    h, R = cb.fields.shock_rotating(xyz, hbar)
    ref = None
    for co in (0, 1, 2):
        os.environ["MAG_COSCHED"] = str(co)
        p = cb.Part(0)
        p.set_stream(stream.cuda_stream)
        p.set_mesh(xyz, ev, tv)
        p.set_size_field_aniso(h, R)
        p.synchronize()
        def step():
            p.clear_flags()
            p.sweep(ops, fp_mode=cb.FP_FAST)
            return p.stats()
        for _ in range(3):
            st = step()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        with torch.cuda.stream(stream):
            e0.record(stream)
            for _ in range(steps):
                st = step()
            e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        L, q = p.edge_lengths(), p.qualities()
        ef, lf = p.flags()
        key = (L.tobytes(), q.tobytes(), ef.tobytes(), lf.tobytes())
        cnt = {k: st[k] for k in ("n_split", "n_collapse", "n_bad", "n_near_threshold", "min_quality", "max_length")}
        if ref is None:
            ref = (key, cnt)
            same = True
        else:
            same = all(a == b for a, b in zip(key, ref[0])) and cnt == ref[1]
        out["jit%.1f_co%d" % (jit, co)] = {"ms_per_step": ms, "identical_to_co0": same}
        print("jitter %.1f cosched %d: %.4f ms per step, identical %s, %s" % (jit, co, ms, same, cnt), flush=True)
        del p
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/r2co_cosched.json", "w"), indent=1)
