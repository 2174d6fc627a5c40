"""BASELINE configs[4] shape on one GPU: mixed prism / tet boundary-layer box (n^3 cells, bottom k layers prisms), rotating
shock-layer AnisoSizeField; tets -> mean-ratio kernel, prisms -> isPrismOk (k_layer) + LAYER closure flags.  Prints device
times per sweep (CUDA events inside mag_sweep)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import core_b200 as cb
n = int(sys.argv[1]) if len(sys.argv) > 1 else 120
k = int(sys.argv[2]) if len(sys.argv) > 2 else 12
t0 = time.perf_counter()
xyz, ev, tv, pv = cb.boxmesh.mixed_box(n, k)
ef, lf = cb.boxmesh.layer_closure_flags(ev, pv, None, len(tv))
h, R = cb.fields.shock_rotating(xyz, 1.0 / n)
print("mesh: %d prisms + %d tets, %d edges, %d vertices (%.1f s on the host)" % (len(pv), len(tv), len(ev), len(xyz), time.perf_counter() - t0))
p = cb.Part(0)
p.set_mesh(xyz, ev, tv, prism_v=pv)
p.set_size_field_aniso(h, R)
steps = 20
for mode, name in ((cb.FP_STRICT, "strict"), (cb.FP_FAST, "fast")):
    for _ in range(3):
        p.set_flags(ef, lf); p.sweep(cb.OP_ALL, fp_mode=mode)
    p.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p.set_flags(ef, lf)
    dflags = None
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    p.timing_begin(steps)
    for _ in range(steps):
        p.set_flags(ef, lf)          # includes the H2D of the flag words (0.1 GB); kernel times come from the events below
        p.sweep(cb.OP_ALL, fp_mode=mode)
    kt = p.timing_read()
    st = p.stats()
    ents = len(ev) + len(tv) + len(pv)
    ms = kt.sum(axis=1).mean()
    print("%s: vertex %.3f + edges %.3f + elements(tets + prisms) %.3f = %.3f ms / sweep, %.3g entities/s; split %d collapse %d bad %d unsafe prisms %d"
          % (name, kt[:, 0].mean(), kt[:, 1].mean(), kt[:, 2].mean(), ms, ents / (ms * 1e-3), st["n_split"], st["n_collapse"], st["n_bad"], st["n_layer_unsafe"]))
