"""Device time of the SURVEY 8f sweeps (element weights, split-vertex transfer) on the n=203 benchmark part."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import core_b200 as cb
n = int(sys.argv[1]) if len(sys.argv) > 1 else 203
xyz, ev, tv = cb.boxmesh.kuhn_box(n, n, n)
h, R = cb.fields.shock_rotating(xyz, 1.0 / n)
p = cb.Part(0)
p.set_mesh(xyz, ev, tv)
p.set_size_field_aniso(h, R)
p.sweep(cb.OP_MARK_SPLIT | cb.OP_MARK_COLLAPSE, fp_mode=cb.FP_FAST)
print("n_split", p.stats()["n_split"])
L = p._L
import ctypes as C
for mode, name in ((cb.FP_STRICT, "strict"), (cb.FP_FAST, "fast")):
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        p._ck(L.mag_element_weights(p._h, float("inf"), float("-inf"), mode, None))
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print("element_weights %s: %.3f ms for %d tets (%.3g tets/s)" % (name, dt * 1e3, len(tv), len(tv) / dt))
for rep in range(2):
    t0 = time.perf_counter()
    idx, sx, sa, sb = p.split_vertices(cb.FP_STRICT)
    dt = time.perf_counter() - t0
print("split_vertices (incl. D2H of %d vertices, %.0f MB): %.1f ms" % (len(idx), (idx.nbytes + sx.nbytes + sa.nbytes + sb.nbytes) / 1e6, dt * 1e3))
