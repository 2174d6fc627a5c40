"""Small end-to-end exercise of every kernel family for compute-sanitizer (memcheck): all kinds, both fp modes, ownership,
incoming flags, mixed mesh with layer check, 2-D part, streamed host call, weights (3-D and 2-D), split vertices, cavity batches, sliver codes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import core_b200 as cb
rng = np.random.default_rng(3)
n = 9
xyz, ev, tv = cb.boxmesh.kuhn_box(n, n + 1, n - 1)
xyz = cb.fields.jitter(xyz, 0.3 / n)
h, R = cb.fields.shock_rotating(xyz, 1.0 / n)
eo = (rng.random(len(ev)) < 0.9).astype(np.uint8)
lo = (rng.random(len(tv)) < 0.9).astype(np.uint8)
ef = np.zeros(len(ev), np.int32); ef[rng.random(len(ev)) < 0.2] |= cb.DONT_SPLIT
lf = np.zeros(len(tv), np.int32); lf[rng.random(len(tv)) < 0.3] |= cb.OK_QUALITY
p = cb.Part(0)
p.set_mesh(xyz, ev, tv, edge_owned=eo, elem_owned=lo)
for kind in ("identity", "iso", "aniso", "logm"):
    if kind == "identity": p.set_size_field_identity()
    elif kind == "iso": p.set_size_field_iso(h[:, 1].copy())
    elif kind == "aniso": p.set_size_field_aniso(h, R)
    else: p.set_size_field_logm_from_frames(h, R, 0)
    for mode in (cb.FP_STRICT, cb.FP_FAST):
        p.set_flags(ef, lf); p.sweep(cb.OP_ALL & ~cb.OP_LAYER_CHECK | cb.OP_LENGTH_SUM, fp_mode=mode); p.stats(); p.flags()
        p.clear_flags(); p.sweep(cb.OP_ALL & ~cb.OP_LAYER_CHECK, use_max=False, fp_mode=mode); p.stats()
        p.element_weights(fp_mode=mode); p.element_weights(0, 1, fp_mode=mode)
        p.split_vertices(fp_mode=mode); p.near_threshold(0); p.near_threshold(1)
        off = np.arange(0, len(tv) + 1, 3, dtype=np.int64); off[-1] = len(tv)
        p.cavity_quality(off, tv, fp_mode=mode)
    p.sliver_codes(None, 0.3); p.sliver_codes(np.ascontiguousarray(tv[:, [1, 2, 0]]), 0.027, only_bad=True)
key = lambda a, b: np.minimum(a, b).astype(np.int64) * (1 << 32) + np.maximum(a, b)
ek = key(ev[:, 0], ev[:, 1]); order = np.argsort(ek)
te = np.stack([order[np.searchsorted(ek[order], key(tv[:, a], tv[:, b]))] for a, b in ((0, 1), (1, 2), (2, 0), (0, 3), (1, 3), (2, 3))], axis=1).astype(np.int32)
p.short_edge_test(te, 2.0)
p.clearFlagFromDimension(cb.SPLIT | cb.COLLAPSE, 1); p.unMarkBadQuality()
p.set_size_field_uniform_refiner(); p.clear_flags(); p.sweep(cb.OP_MARK_SPLIT | cb.OP_MARK_COLLAPSE); p.stats()
p.set_size_field_identity(); p.getAverageEdgeLength()
bad = tv.copy(); bad[3, 1] = len(xyz) + 5
try:
    p.set_mesh(xyz, ev, bad)          # the export-time range check must stop this before anything gathers through it
    raise SystemExit("bad connectivity accepted")
except cb.MagError:
    pass
p.set_mesh(xyz, ev, tv, edge_owned=eo, elem_owned=lo); p.set_size_field_aniso(h, R)
oL, oq = np.empty(len(ev)), np.empty(len(tv))
oe, ol = np.empty(len(ev), np.int32), np.empty(len(tv), np.int32)
p.sweep_host(xyz, ev, tv, 2, h, R, edge_flags=ef, elem_flags=lf, edge_owned=eo, elem_owned=lo, out_lengths=oL, out_qualities=oq,
             out_edge_flags=oe, out_elem_flags=ol, fp_mode=cb.FP_FAST, slice_entities=61440)
p.clear_flags(); p.sweep(cb.OP_ALL & ~cb.OP_LAYER_CHECK, fp_mode=cb.FP_FAST); p.stats()
# round 2: the winner-in-slot tet kernel (cp.async stage) after a field change and after new coordinates, the listed-only fast
# mode, edge-collapse candidates over the device-built vertex -> tet incidence (both ends of every 7th edge)
p.set_size_field_aniso(h * 1.5, R); p.clear_flags(); p.sweep(cb.OP_ALL & ~cb.OP_LAYER_CHECK, fp_mode=cb.FP_FAST); p.stats()
p.set_coords(xyz * 1.01); p.clear_flags(); p.sweep(cb.OP_ALL & ~cb.OP_LAYER_CHECK, fp_mode=cb.FP_FAST_LISTED); p.stats(); p.row_layout(1)
ce = np.arange(0, len(ev), 7, dtype=np.int32)
for mode in (cb.FP_STRICT, cb.FP_FAST):
    p.collapse_quality(np.concatenate([ce, ce]), np.concatenate([np.zeros(len(ce), np.uint8), np.ones(len(ce), np.uint8)]), fp_mode=mode)
x2, e2, t2, pr2 = cb.boxmesh.mixed_box(5, 2)
ef2, lf2 = cb.boxmesh.layer_closure_flags(e2, pr2, None, len(t2))
h2, R2 = cb.fields.shock_rotating(x2, 0.2)
p.set_mesh(x2, e2, t2, prism_v=pr2); p.set_size_field_aniso(h2, R2); p.set_flags(ef2, lf2); p.sweep(cb.OP_ALL, fp_mode=cb.FP_FAST); p.stats(); p.layer_ok()
# round 2: layer prisms weighed by their base triangle, the device LAYER closure, the listed-only fast mode on a mixed part
p.prism_weights(np.ascontiguousarray(pr2[:, :3]), fp_mode=cb.FP_STRICT); p.prism_weights(np.ascontiguousarray(pr2[:, :3]), 0, 1, False, False, True, fp_mode=cb.FP_FAST)
p.clear_flags(); p.reset_layer(); p.sweep(cb.OP_ALL, fp_mode=cb.FP_FAST_LISTED); p.stats()
x3, e3, tr3 = cb.boxmesh.tri_box(9, 7)
h3, R3 = cb.fields.shock_rotating(x3, 1.0 / 8)
p.set_mesh_2d(x3, e3, tr3); p.set_size_field_aniso(h3, R3)
for mode in (cb.FP_STRICT, cb.FP_FAST):
    p.clear_flags(); p.sweep(cb.OP_ALL & ~cb.OP_LAYER_CHECK, good_quality=0.2, fp_mode=mode); p.stats()
    p.element_weights(fp_mode=mode); p.element_weights(0, 1, fp_mode=mode, dim=2); p.split_vertices(fp_mode=mode)
for setter in (p.set_size_field_identity, lambda: p.set_size_field_iso(h3[:, 1].copy()), lambda: p.set_size_field_logm_from_frames(h3, R3, 0)):
    setter(); p.element_weights(fp_mode=cb.FP_STRICT); p.element_weights(fp_mode=cb.FP_FAST)
p.close()
print("sanitize_small: done")
