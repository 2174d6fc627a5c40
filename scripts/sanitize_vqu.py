"""compute-sanitizer exercise of the per-vertex uniform-edge transform cache (k_vertex_uniform, the Q_u branch of the strict
re-evaluation in near_edges / drain_edges and its L2 prefetch): a lattice whose field does not vary along z, so that the z edges
sit on the collapse threshold with identical end values; aniso (lean rows) and LogAniso (tile kernel), owned and not, then the
same after new coordinates, a new field and a new mesh of another size."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import core_b200 as cb
rng = np.random.default_rng(5)
ops = cb.OP_ALL & ~cb.OP_LAYER_CHECK
p = cb.Part(0)
total = 0
for n in (7, 5):
    xyz, ev, tv = cb.boxmesh.kuhn_box(n, n + 1, n)
    h, R = cb.fields.shock_rotating(xyz, 1.0 / n)
    eo = (rng.random(len(ev)) < 0.9).astype(np.uint8)
    p.set_mesh(xyz, ev, tv, edge_owned=eo)
    for kind in ("aniso", "logm"):
        if kind == "aniso": p.set_size_field_aniso(h, R)
        else: p.set_size_field_logm_from_frames(h, R, 0)
        for x in (xyz, cb.fields.jitter(xyz, 0.02 / n), xyz):
            p.set_coords(x)
            for mode in (cb.FP_FAST, cb.FP_STRICT, cb.FP_FAST_LISTED):
                p.clear_flags(); p.sweep(ops, fp_mode=mode); st = p.stats(); p.flags(); p.edge_lengths()
                idx, cnt = p.near_threshold(0)
                total += cnt
            ef = np.zeros(len(ev), np.int32); ef[rng.random(len(ev)) < 0.2] |= cb.DONT_SPLIT
            p.set_flags(ef, None); p.sweep(ops, fp_mode=cb.FP_FAST); p.stats()     # incoming words: the tile kernel's queue (drain_edges)
p.close()
assert total > 0
print("sanitize_vqu ok, near-threshold entries seen:", total)
