"""Parity of the CUDA path (through the C ABI, core_b200.Part) with the restated oracle and with the golden vectors
the compiled reference produced.  Bar (BASELINE.json north_star): flags and counts bit-exact; lengths and qualities
bit-exact in MAG_FP_STRICT (LogAniso: CUDA exp() vs glibc exp() -> 1e-12 relative) and within 1e-12 relative in
MAG_FP_FAST; every entity whose value lies within 1e-12 relative of a threshold is listed."""
import numpy as np
import pytest

import util

pytestmark = pytest.mark.gpu
TOL = 1e-12


@pytest.fixture(scope="module")
def cb(built):
    import torch
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    import core_b200
    return core_b200


def check_against(cb, p, want, kind, mode, nns=0, exact_values=None):
    from oracle import mao
    st = p.stats()
    L, q = p.edge_lengths(), p.qualities()
    ef, lf = p.flags()
    if exact_values is None:
        exact_values = (mode == cb.FP_STRICT and kind != mao.LOGM)
    if exact_values:
        assert np.array_equal(L, want["lengths"]), "lengths not bit-exact (max rel %.3g)" % util.rel_err(L, want["lengths"])
        assert np.array_equal(q, want["qualities"]), "qualities not bit-exact"
        assert st["min_quality"] == want["min_quality"] and st["max_length"] == want["max_length"]
    else:
        assert util.rel_err(L, want["lengths"]) < TOL
        assert util.rel_err(q[nns:], want["qualities"][nns:]) < TOL
        assert abs(st["min_quality"] - want["min_quality"]) <= TOL * abs(want["min_quality"])
        assert abs(st["max_length"] - want["max_length"]) <= TOL * want["max_length"]
    near_e, n_e = p.near_threshold(0)
    near_l, n_l = p.near_threshold(1)
    assert st["n_near_threshold"] == n_e + n_l
    # flags: identical everywhere except (LogAniso only) on listed near-threshold entities
    de = np.nonzero(ef != want["edge_flags"])[0]
    dl = np.nonzero(lf != want["elem_flags"])[0]
    if kind == mao.LOGM:
        assert set(de.tolist()) <= set(near_e.tolist()) and set(dl.tolist()) <= set(near_l.tolist())
    else:
        assert len(de) == 0 and len(dl) == 0, "flags differ on %d edges / %d elements" % (len(de), len(dl))
        assert (st["n_split"], st["n_collapse"], st["n_bad"]) == (want["n_split"], want["n_collapse"], want["n_bad"])
    return st


@pytest.mark.parametrize("mode", ["strict", "fast"])
@pytest.mark.parametrize("name", util.golden_cases())
def test_golden_reference_vectors(cb, name, mode):
    """CUDA path vs the compiled reference's own outputs (tests/golden)."""
    from oracle import mao
    g = util.load(name)
    mode = cb.FP_STRICT if mode == "strict" else cb.FP_FAST
    kind, ma, mb = util.metric_arrays(g)
    two_d = util.is_2d(g)
    prism_v, pyr_v, tet_v = (np.zeros((0, 6), np.int32), np.zeros((0, 5), np.int32), None) if two_d else util.split_elements(g)
    nns = len(prism_v) + len(pyr_v)
    ef_in, lf_in = g["edge_flags_in"].copy(), g["elem_flags_in"].copy()
    if nns:
        ef_in |= g["edge_flags_ctor"]
        lf_in |= g["elem_flags_ctor"]
    gq = float(g["good_quality"])
    gq = 0.027 if gq < 0 else gq
    p = cb.Part(0)
    if two_d:
        p.set_mesh_2d(g["xyz"], g["edge_v"], np.ascontiguousarray(g["elem_v"][:, :3]))
    else:
        p.set_mesh(g["xyz"], g["edge_v"], tet_v, prism_v if len(prism_v) else None, pyr_v if len(pyr_v) else None)
    if kind == mao.LOGM:
        lm = p.set_size_field_logm_from_frames(g["h"], g["R"], util.logm_variant(g), want_logm=True)
        assert np.array_equal(lm, g["logM"]), "host-built logM field differs from the reference's ma_logM"
    else:
        util.set_part_metric(p, kind, ma, mb)
    p.set_flags(ef_in, lf_in)
    p.sweep(cb.OP_ALL, good_quality=gq, fp_mode=mode)
    q_want = np.concatenate([np.zeros(nns), g["qualities"]]) if "qualities" in g else None
    want = dict(lengths=g["lengths"], qualities=q_want, edge_flags=g["edge_flags_out"], elem_flags=g["elem_flags_out"],
                n_split=int(g["counts"][0]), n_collapse=int(g["counts"][1]), n_bad=int(g["counts"][2]),
                min_quality=float(g["min_q"]), max_length=float(g["max_len"]))
    if q_want is None:   # mixed mesh: the reference cannot run getLinearQualities on it; use the pinned oracle for tets
        q = mao.tet_qualities(kind, g["xyz"], ma, mb, tet_v)
        want["qualities"] = np.concatenate([np.zeros(nns), q])
        want["min_quality"] = mao.min_quality(q)
    check_against(cb, p, want, kind, mode, nns)
    if nns:
        ok, codes = p.layer_ok()
        assert np.array_equal(ok, g["layer_ok"][:nns]) and np.array_equal(codes, g["layer_codes"][:nns])
    # centroid metric (useMax = false, maQuality.cc:148-153)
    if "qualities_centroid" in g:
        p.set_flags(None, None)
        p.sweep(cb.OP_QUALITIES, use_max=False, fp_mode=mode)
        q = p.qualities()
        if mode == cb.FP_STRICT and kind != mao.LOGM:
            assert np.array_equal(q, g["qualities_centroid"])
        else:
            assert util.rel_err(q, g["qualities_centroid"]) < TOL
    p.close()


@pytest.mark.parametrize("name", [n for n in util.golden_cases() if "weights_raw" in util.load(n)])
def test_weights_and_split_vertices_golden(cb, name):
    """SURVEY 8f rows on the device against the compiled reference's outputs: ma::getElementWeight (raw and clamped)
    and ma::makeSplitVert's transfer (position + size-field values of the vertex splitting every SPLIT edge)."""
    from oracle import mao
    g = util.load(name)
    kind, ma, mb = util.metric_arrays(g)
    p = cb.Part(0)
    two_d = util.is_2d(g)
    if two_d:   # triangles are the elements: measure(triangle) / (1/2), clampForIterations with dimension 2
        p.set_mesh_2d(g["xyz"], g["edge_v"], np.ascontiguousarray(g["elem_v"][:, :3]))
    else:
        p.set_mesh(g["xyz"], g["edge_v"], util.split_elements(g)[2])
    util.set_part_metric(p, kind, ma, mb)
    exact = kind != mao.LOGM          # LogAniso: CUDA exp() vs glibc exp() (see MAG_FP_STRICT in mag.h)
    for mode in (cb.FP_STRICT, cb.FP_FAST):
        w = p.element_weights(fp_mode=mode)
        wc = p.element_weights(0, 1, fp_mode=mode, dim=2 if two_d else 3)
        if exact and mode == cb.FP_STRICT:
            assert np.array_equal(w, g["weights_raw"]) and np.array_equal(wc, g["weights_r0_c1"])
        else:
            assert util.rel_err(w, g["weights_raw"]) < TOL and util.rel_err(wc, g["weights_r0_c1"]) < TOL
    if "split_edges" in g:
        gq = float(g["good_quality"])
        p.set_flags(g["edge_flags_in"], g["elem_flags_in"])
        p.sweep(cb.OP_MARK_SPLIT | cb.OP_MARK_COLLAPSE, fp_mode=cb.FP_STRICT)
        ef, _ = p.flags()
        for mode in (cb.FP_STRICT, cb.FP_FAST):
            idx, sx, sa, sb = p.split_vertices(fp_mode=mode)
            assert np.array_equal(idx, np.nonzero(ef & cb.SPLIT)[0])        # edge order, every SPLIT edge once
            if kind != mao.LOGM:
                assert np.array_equal(idx, g["split_edges"])
            sel = np.searchsorted(g["split_edges"], idx)
            ok = (sel < len(g["split_edges"])) & (g["split_edges"][np.minimum(sel, len(g["split_edges"]) - 1)] == idx)
            assert ok.mean() > 0.99                                            # LogAniso: flags may differ in the listed band
            sel, keep = sel[ok], np.nonzero(ok)[0]
            if mode == cb.FP_STRICT:
                assert np.array_equal(sx[keep], g["split_xyz"][sel]) and np.array_equal(sb[keep], g["split_b"][sel])
                if kind == mao.ANISO:
                    assert np.array_equal(sa[keep], g["split_a"][sel])
            else:
                assert util.rel_err(sx[keep], g["split_xyz"][sel]) < TOL
                assert np.max(np.abs(sb[keep] - g["split_b"][sel])) < TOL * max(1.0, np.max(np.abs(g["split_b"])))
    p.close()


@pytest.mark.parametrize("name", [n for n in util.golden_cases() if "prism_base_v" in util.load(n)])
def test_prism_weights_golden(cb, name):
    """ma::getElementWeight of layer prisms (maBalance.cc:21-81: the base triangle's getWeight, clamps, layer permissions) on the
    device against the compiled reference: raw weights bit for bit in strict arithmetic, ma::getElementWeight with the Input's
    defaults (no layer refinement / coarsening: exactly 1), and the tets of the same mixed part through mag_element_weights."""
    g = util.load(name)
    kind, ma, mb = util.metric_arrays(g)
    pr, py, te = util.split_elements(g)
    p = cb.Part(0)
    p.set_mesh(g["xyz"], g["edge_v"], te, prism_v=pr, pyr_v=py)
    util.set_part_metric(p, kind, ma, mb)
    base = g["prism_base_v"][: len(pr)]
    assert np.all(base >= 0)
    raw, cl = g["layer_weights_raw"], g["layer_weights_r0_c1"]
    n0 = len(pr) + len(py)
    for mode in (cb.FP_STRICT, cb.FP_FAST):
        w = p.prism_weights(base, fp_mode=mode)
        if mode == cb.FP_STRICT:
            assert np.array_equal(w, raw[: len(pr)])
        else:
            assert util.rel_err(w, raw[: len(pr)]) < TOL
        wc = p.prism_weights(base, 0, 1, refine_layer=False, coarsen_layer=False, fp_mode=mode)
        assert np.array_equal(wc, cl[: len(pr)])
        w3 = p.prism_weights(base, 0, 1, refine_layer=True, coarsen_layer=True, to_tets=True, fp_mode=cb.FP_STRICT)
        assert np.array_equal(w3, 3.0 * np.clip(raw[: len(pr)], 0.25, 1.0))
        wt = p.element_weights(0, 1, fp_mode=mode)
        if mode == cb.FP_STRICT:
            assert np.array_equal(wt[n0:], cl[n0:])
        else:
            assert util.rel_err(wt[n0:], cl[n0:]) < TOL
    with pytest.raises(cb.MagError):
        bad = base.copy(); bad[0, 0] = len(g["xyz"])
        p.prism_weights(bad)
    p.close()


@pytest.mark.parametrize("name", [n for n in util.golden_cases() if "sliver_codes" in util.load(n)])
def test_sliver_codes_golden(cb, name):
    """SURVEY 8f row 1, second classification sweep: ma::getSliverCode / matchSliver on the device against the compiled
    reference's own codes (every tet; then only the BAD_QUALITY ones after a marking sweep)."""
    from oracle import mao
    g = util.load(name)
    kind, ma, mb = util.metric_arrays(g)
    tet_v = util.split_elements(g)[2]
    gq = float(g["good_quality"])
    gq = 0.027 if gq < 0 else gq
    p = cb.Part(0)
    p.set_mesh(g["xyz"], g["edge_v"], tet_v)
    util.set_part_metric(p, kind, ma, mb)
    codes, match = p.sliver_codes(g["face0_v"], gq)
    if kind != mao.LOGM:
        assert np.array_equal(codes, g["sliver_codes"]) and np.array_equal(match, g["sliver_match"])
    else:   # CUDA exp() vs glibc exp(): a tet whose area coordinate sits within rounding of a cut may flip a bit
        assert (codes == g["sliver_codes"]).mean() > 0.995
        same = codes == g["sliver_codes"]
        assert np.array_equal(match[same], g["sliver_match"][same])
    # the tet's own (v0, v1, v2) instead of the face's order: same face, same code unless the face quality ties the cut
    c2, _ = p.sliver_codes(None, gq)
    assert (c2 == codes).mean() > 0.995
    # only the BAD_QUALITY tets of the resident flags
    p.set_flags(g["edge_flags_in"], g["elem_flags_in"])
    p.sweep(cb.OP_MARK_BAD, good_quality=gq, fp_mode=cb.FP_STRICT)
    _, lf = p.flags()
    bad = (lf & cb.BAD_QUALITY) != 0
    c3, m3 = p.sliver_codes(g["face0_v"], gq, only_bad=True)
    assert np.array_equal(c3[bad], codes[bad]) and np.array_equal(m3[bad], match[bad])
    assert np.all(c3[~bad] == 0) and np.all(m3[~bad] == -1)
    p.close()


def edge_cavities(edge_v, tet_v):
    """CSR of the tets around every edge (the cavity of an edge collapse / swap), built from the connectivity alone."""
    pairs = [(0, 1), (1, 2), (2, 0), (0, 3), (1, 3), (2, 3)]
    key = lambda a, b: np.minimum(a, b).astype(np.int64) * (1 << 32) + np.maximum(a, b)
    ek = key(edge_v[:, 0], edge_v[:, 1])
    order = np.argsort(ek)
    tk = np.concatenate([key(tet_v[:, a], tet_v[:, b]) for a, b in pairs])
    tid = np.tile(np.arange(len(tet_v)), 6)
    pos = order[np.searchsorted(ek[order], tk)]          # edge index of every (tet, local edge)
    o = np.argsort(pos, kind="stable")
    counts = np.bincount(pos, minlength=len(edge_v))
    offsets = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    return offsets, tid[o]


@pytest.mark.parametrize("name", [n for n in util.golden_cases() if "sliver_codes" in util.load(n)])
def test_cavity_quality_golden(cb, name):
    """batch ma::getWorstQuality.  Cavities = the tets around every edge of the golden mesh; the worst
    quality of each must equal the minimum of the compiled reference's own per-tet qualities."""
    from oracle import mao
    g = util.load(name)
    kind, ma, mb = util.metric_arrays(g)
    _, _, tet_v = util.split_elements(g)
    offsets, members = edge_cavities(g["edge_v"], tet_v)
    assert offsets[-1] == 6 * len(tet_v) and np.all(np.diff(offsets) > 0)
    want = np.minimum.reduceat(g["qualities"][members], offsets[:-1])
    want_c = np.minimum.reduceat(g["qualities_centroid"][members], offsets[:-1])
    p = cb.Part(0)
    p.set_mesh(g["xyz"], g["edge_v"], tet_v)
    util.set_part_metric(p, kind, ma, mb)
    for mode in (cb.FP_STRICT, cb.FP_FAST):
        worst, q = p.cavity_quality(offsets, tet_v[members], fp_mode=mode, want_qualities=True)
        worst_c = p.cavity_quality(offsets, tet_v[members], use_max=False, fp_mode=mode)
        if mode == cb.FP_STRICT and kind != mao.LOGM:
            assert np.array_equal(worst, want) and np.array_equal(q, g["qualities"][members])
            assert np.array_equal(worst_c, want_c)
        else:
            assert util.rel_err(worst, want) < TOL and util.rel_err(worst_c, want_c) < TOL
    p.close()


def tet_edge_table(edge_v, tet_v):
    """[nt,6] edge indices of every tet in getDownward(tet, 1) order = tet_edge_verts {01,12,20,03,13,23} (apfMesh.cc:53-60)."""
    key = lambda a, b: np.minimum(a, b).astype(np.int64) * (1 << 32) + np.maximum(a, b)
    ek = key(edge_v[:, 0], edge_v[:, 1])
    order = np.argsort(ek)
    cols = [order[np.searchsorted(ek[order], key(tet_v[:, a], tet_v[:, b]))] for a, b in ((0, 1), (1, 2), (2, 0), (0, 3), (1, 3), (2, 3))]
    return np.ascontiguousarray(np.stack(cols, axis=1), dtype=np.int32)


@pytest.mark.parametrize("name", ["jbox7_shock_rot_aniso", "jbox7_random_aniso_flags", "box6_shock_planar_aniso"])
def test_short_edge_classification(cb, name):
    """ShortEdgeFixer::shouldApply (maShape.cc:188-219) over the BAD_QUALITY tets against the COMPILED REFERENCE: the class is
    local to maShape.cc, so oracle/ref/ref_shape_shim.cc compiles that file into the test driver and the golden vectors hold
    what the reference's own object answered for both defaults of maximumEdgeRatio (maInput.cc:35,43) -- the edge handed to
    the ShortEdgeRemover (first shortest edge in getDownward order) and the flag words afterwards (BAD_QUALITY cleared below
    the ratio).  Ratio 1.0 (clears nothing) is checked against the same few lines restated with numpy."""
    g = util.load(name)
    kind, ma, mb = util.metric_arrays(g)
    _, _, tet_v = util.split_elements(g)
    te = tet_edge_table(g["edge_v"], tet_v)
    gq = float(g["good_quality"])
    gq = 0.027 if gq < 0 else gq
    p = cb.Part(0)
    p.set_mesh(g["xyz"], g["edge_v"], tet_v)
    util.set_part_metric(p, kind, ma, mb)
    bad = (g["elem_flags_out"] & cb.BAD_QUALITY) != 0
    for ratio in (2.0, 100.0, 1.0):
        p.set_flags(g["edge_flags_in"], g["elem_flags_in"])
        p.sweep(cb.OP_ALL & ~cb.OP_LAYER_CHECK, good_quality=gq, fp_mode=cb.FP_STRICT)
        short, n_cleared, n_short = p.short_edge_test(te, ratio)
        lf = p.flags()[1]
        if ratio != 1.0:
            want, want_lf = g["short_edge_%g" % ratio], g["short_flags_%g" % ratio]
        else:
            l = g["lengths"][te]
            cleared = bad & (l.max(axis=1) / l.min(axis=1) < ratio)
            want = np.where(bad & ~cleared, te[np.arange(len(te)), l.argmin(axis=1)], -1)
            want_lf = np.where(cleared, g["elem_flags_out"] & ~cb.BAD_QUALITY, g["elem_flags_out"])
        assert np.array_equal(short, want)
        assert np.array_equal(lf, want_lf)
        assert n_short == int((want >= 0).sum()) and n_cleared == int(bad.sum()) - n_short
    assert bad.sum() > 0
    p.close()


@pytest.mark.parametrize("kindname", ["aniso", "iso"])
def test_collapse_candidates_vs_oracle(cb, kindname):
    """mag_collapse_quality (ma::Collapse's quality test, maCollapse.cc:88-113,353-383,425-433) against the restated oracle: for
    both ends of sampled edges of a jittered box, the tets around the collapsing vertex that do not hold the edge are rebuilt
    with the vertex replaced (numpy), their worst quality and the worst quality of all tets around the vertex must come out
    bit for bit in strict arithmetic, within 1e-12 in fast arithmetic; boundary vertices give small one-sided cavities, and a
    collapse across the box inverts tets (negative qualities)."""
    from oracle import mao
    n = 9
    rng = np.random.default_rng(5)
    xyz, ev, tv = cb.boxmesh.kuhn_box(n, n, n)
    xyz = cb.fields.jitter(xyz, 0.3 / n)
    h, R = cb.fields.shock_rotating(xyz, 1.0 / n)
    if kindname == "aniso":
        kind, ma, mb = mao.ANISO, h, R
    else:
        kind, ma, mb = mao.ISO, np.ascontiguousarray(h[:, 1] * (1 + xyz[:, 0])), None
    eo = (rng.random(len(ev)) < 0.9).astype(np.uint8)        # ownership bits ride in the connectivity: must not leak into ids
    lo = (rng.random(len(tv)) < 0.9).astype(np.uint8)
    p = cb.Part(0)
    p.set_mesh(xyz, ev, tv, edge_owned=eo, elem_owned=lo)
    util.set_part_metric(p, kind, ma, mb)
    q_all = mao.tet_qualities(kind, xyz, ma, mb, tv)
    edges = rng.choice(len(ev), 1500, replace=False).astype(np.int32)
    edges = np.concatenate([edges, edges])
    ends = np.concatenate([np.zeros(1500, np.uint8), np.ones(1500, np.uint8)])
    # vertex -> tets
    order = np.argsort(tv.ravel(), kind="stable")
    owner = (order // 4).astype(np.int64)
    start = np.searchsorted(tv.ravel()[order], np.arange(len(xyz) + 1))
    want_new, want_old, want_keep = np.empty(len(edges)), np.empty(len(edges)), np.empty(len(edges), np.int32)
    cand, offs = [], [0]
    for k, (e, w) in enumerate(zip(edges, ends)):
        vc, vk = (ev[e, 1], ev[e, 0]) if w else (ev[e, 0], ev[e, 1])
        around = owner[start[vc]:start[vc + 1]]
        want_old[k] = q_all[around].min()
        keep = around[~np.any(tv[around] == vk, axis=1)]
        t = tv[keep].copy()
        t[t == vc] = vk
        cand.append(t)
        offs.append(offs[-1] + len(t))
        want_keep[k] = len(t)
    cand = np.ascontiguousarray(np.concatenate(cand), np.int32)
    qn = mao.tet_qualities(kind, xyz, ma, mb, cand)
    offs = np.array(offs)
    has = want_keep > 0
    want_new[has] = np.minimum.reduceat(qn, offs[:-1][has])
    want_new[~has] = np.inf
    assert (want_new < 0).any() and has.mean() > 0.95      # a corner vertex may keep nothing: +inf
    new_w, old_w, keep_n = p.collapse_quality(edges, ends, fp_mode=cb.FP_STRICT)
    assert np.array_equal(keep_n, want_keep)
    assert np.array_equal(old_w, want_old) and np.array_equal(new_w, want_new)
    new_f, old_f, _ = p.collapse_quality(edges, ends, fp_mode=cb.FP_FAST)
    assert np.all(np.abs(new_f[has] - want_new[has]) <= TOL * np.abs(want_new[has]) + 1e-15) and np.all(np.isinf(new_f[~has]))
    assert util.rel_err(old_f, want_old) < TOL
    # the batch through mag_cavity_quality gives the same numbers (the incidence lists are the only new ingredient)
    offs_ne = np.concatenate([[0], np.cumsum(want_keep[has])]).astype(np.int64)        # mag_cavity_quality takes no empty cavity
    worst = p.cavity_quality(offs_ne, cand, fp_mode=cb.FP_STRICT)
    assert np.array_equal(worst, want_new[has])
    # a new mesh in the same context: the incidence is rebuilt
    xyz2, ev2, tv2 = cb.boxmesh.kuhn_box(4, 5, 3)
    h2, R2 = cb.fields.shock_rotating(xyz2, 0.25)
    p.set_mesh(xyz2, ev2, tv2)
    p.set_size_field_aniso(h2, R2)
    q2 = mao.tet_qualities(mao.ANISO, xyz2, h2, R2, tv2)
    e2 = np.arange(len(ev2), dtype=np.int32)
    _, old2, _ = p.collapse_quality(e2, np.zeros(len(ev2), np.uint8))
    want2 = np.array([q2[np.any(tv2 == ev2[e, 0], axis=1)].min() for e in e2])
    assert np.array_equal(old2, want2)
    with pytest.raises(cb.sweep.MagError):
        p.collapse_quality(np.array([len(ev2)], np.int32), np.zeros(1, np.uint8))
    p.close()


def test_cavity_quality_candidate_tets_vs_oracle(cb):
    """Would-be elements: random vertex quadruples (many inverted -> negative qualities) that are not mesh entities,
    ragged cavities of 1..40 tets, against the restated oracle; error paths of the batch call."""
    from oracle import mao
    n = 10
    rng = np.random.default_rng(11)
    xyz, ev, tv = cb.boxmesh.kuhn_box(n, n, n)
    xyz = cb.fields.jitter(xyz, 0.3 / n)
    h, R = cb.fields.shock_rotating(xyz, 1.0 / n)
    p = cb.Part(0)
    p.set_mesh(xyz, ev, tv)
    p.set_size_field_aniso(h, R)
    sizes = rng.integers(1, 41, 3000)
    offsets = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    base = rng.integers(0, len(xyz) - 150, offsets[-1])
    cand = (base[:, None] + rng.integers(0, 150, (offsets[-1], 4))).astype(np.int32)     # nearby, arbitrary orientation
    q0 = mao.tet_qualities(mao.ANISO, xyz, h, R, cand)
    assert (q0 < 0).mean() > 0.2
    want = np.minimum.reduceat(q0, offsets[:-1])
    worst, q = p.cavity_quality(offsets, cand, fp_mode=cb.FP_STRICT, want_qualities=True)
    assert np.array_equal(q, q0) and np.array_equal(worst, want)
    worst_f = p.cavity_quality(offsets, cand, fp_mode=cb.FP_FAST)
    # degenerate candidates (repeated / coplanar vertices) have |V| ~ 0 by cancellation: relative error is meaningless
    # there, so the bound is 1e-12 relative plus an absolute floor far below any quality threshold
    assert np.all(np.abs(worst_f - want) <= TOL * np.abs(want) + 1e-15)
    # a sweep in between (new vertex pass) changes nothing; a new size field does
    p.sweep(cb.OP_QUALITIES, fp_mode=cb.FP_FAST)
    assert np.array_equal(p.cavity_quality(offsets, cand), want)
    p.set_size_field_iso(h[:, 1].copy())
    want_iso = np.minimum.reduceat(mao.tet_qualities(mao.ISO, xyz, h[:, 1].copy(), None, cand), offsets[:-1])
    assert np.array_equal(p.cavity_quality(offsets, cand), want_iso)
    with pytest.raises(cb.sweep.MagError):
        p.cavity_quality(np.array([0, 2, 2], np.int64), cand[:2])          # empty cavity: the reference asserts n > 0
    bad = cand[:4].copy()
    bad[1, 2] = len(xyz)
    with pytest.raises(cb.sweep.MagError):
        p.cavity_quality(np.array([0, 4], np.int64), bad)                   # vertex id out of range
    p.close()


@pytest.mark.parametrize("name", ["mixed5_shock_rot_aniso", "pyrslab_shock_rot_aniso"])
def test_reset_layer_golden(cb, name):
    """mag_reset_layer = ma::resetLayer (maLayer.cc:94-103): the edge / element flag words the reference's Adapt constructor
    leaves on a mesh with prisms / pyramids (golden edge_flags_ctor / elem_flags_ctor), bit for bit; then the full sweep on
    top of them equals the reference's."""
    g = util.load(name)
    pr, py, te = util.split_elements(g)
    kind, ma, mb = util.metric_arrays(g)
    p = cb.Part(0)
    p.set_mesh(g["xyz"], g["edge_v"], te, prism_v=pr, pyr_v=py)
    util.set_part_metric(p, kind, ma, mb)
    p.clear_flags()
    n = p.reset_layer()
    assert n == len(pr) + len(py)
    ef, lf = p.flags()
    assert np.array_equal(ef, g["edge_flags_ctor"]) and np.array_equal(lf, g["elem_flags_ctor"])
    assert p.reset_layer() == n                                  # idempotent
    ef2, lf2 = p.flags()
    assert np.array_equal(ef2, ef) and np.array_equal(lf2, lf)
    p.sweep(cb.OP_ALL, good_quality=float(g["good_quality"]) if float(g["good_quality"]) > 0 else cb.GOOD_QUALITY_3D, fp_mode=cb.FP_STRICT)
    st = p.stats()
    ef3, lf3 = p.flags()
    if np.array_equal(g["edge_flags_in"], g["edge_flags_ctor"]) or not g["edge_flags_in"].any():
        assert np.array_equal(ef3, g["edge_flags_out"]) and np.array_equal(lf3, g["elem_flags_out"])
        assert (st["n_split"], st["n_collapse"], st["n_bad"]) == tuple(int(x) for x in g["counts"])
    p.close()


def test_reset_layer_user_tag_and_errors(cb):
    """Input::userDefinedLayerTagName (maLayer.cc:24-39): tagged elements -- tets included -- join the layer; a layer element
    whose edges the part does not hold is refused; a 2-D part is refused."""
    rng = np.random.default_rng(11)
    xyz, ev, tv, pv = cb.boxmesh.mixed_box(6, 2)
    nel = len(pv) + len(tv)
    tag = np.zeros(nel, np.int32)
    tagged = rng.choice(np.arange(len(pv), nel), 40, replace=False)
    tag[tagged] = 3
    p = cb.Part(0)
    p.set_mesh(xyz, ev, tv, prism_v=pv)
    p.clear_flags()
    assert p.reset_layer(tag) == len(pv) + 40
    ef, lf = p.flags()
    # numpy model: closure edges of the prisms and of the tagged tets
    want_e, want_l = cb.boxmesh.layer_closure_flags(ev, pv, None, len(tv))
    tt = tv[tagged - len(pv)].astype(np.int64)
    pairs = np.sort(np.concatenate([tt[:, [a, b]] for a, b in ((0, 1), (1, 2), (2, 0), (0, 3), (1, 3), (2, 3))]), axis=1)
    nvmax = len(xyz)
    hit = np.isin(np.sort(ev.astype(np.int64), axis=1) @ np.array([nvmax, 1]), pairs @ np.array([nvmax, 1]))
    want_e[hit] = cb.LAYER | cb.DONT_COLLAPSE | cb.DONT_SPLIT | cb.DONT_SWAP
    want_l[tagged] = cb.LAYER | cb.OK_QUALITY
    assert np.array_equal(ef, want_e) and np.array_equal(lf, want_l)
    # incoming bits survive (the call ORs)
    ef_in = np.zeros(len(ev), np.int32)
    ef_in[::7] = cb.CHECKED
    p.set_flags(ef_in, np.zeros(nel, np.int32))
    p.reset_layer()
    ef2, _ = p.flags()
    assert np.array_equal(ef2, cb.boxmesh.layer_closure_flags(ev, pv, None, len(tv))[0] | ef_in)
    # a part that misses an edge of a prism
    p.set_mesh(xyz, ev[1:], tv, prism_v=pv)
    p.clear_flags()
    with pytest.raises(cb.MagError) as e:
        p.reset_layer()
    assert e.value.code == 2
    x2, e2, t2 = cb.boxmesh.tri_box(4, 3)
    p.set_mesh_2d(x2, e2, t2)
    with pytest.raises(cb.MagError):
        p.reset_layer()
    p.close()


def test_unsafe_prisms(cb):
    g = util.load("mixed5_unsafe_layer")
    prism_v, _, tet_v = util.split_elements(g)
    p = cb.Part(0)
    p.set_mesh(g["xyz"], g["edge_v"], tet_v, prism_v)
    p.set_size_field_identity()
    p.sweep(cb.OP_LAYER_CHECK)
    ok, codes = p.layer_ok()
    assert np.array_equal(ok, g["layer_ok"][:len(prism_v)]) and np.array_equal(codes, g["layer_codes"][:len(prism_v)])
    assert p.stats()["n_layer_unsafe"] == 8
    p.close()


def test_unsafe_pyramids(cb):
    """ma::isPyramidOk on the device against the compiled reference (good rotations -1 / 0 / 1), prisms of the same mesh too."""
    g = util.load("pyrslab_unsafe_layer")
    prism_v, pyr_v, tet_v = util.split_elements(g)
    p = cb.Part(0)
    p.set_mesh(g["xyz"], g["edge_v"], None, prism_v, pyr_v)
    p.set_size_field_identity()
    p.sweep(cb.OP_LAYER_CHECK)
    ok, codes = p.layer_ok()
    assert np.array_equal(ok, g["layer_ok"]) and np.array_equal(codes, g["layer_codes"])
    assert p.stats()["n_layer_unsafe"] == int((g["layer_ok"] == 0).sum())
    p.close()


@pytest.mark.parametrize("kindname", ["identity", "iso", "aniso", "logm"])
def test_random_box_vs_oracle(cb, kindname):
    """Seeded jittered box, random frames / sizes, random incoming flag words and ownership, both fp modes."""
    from oracle import mao
    n = 14
    rng = np.random.default_rng(42)
    xyz, ev, tv = cb.boxmesh.kuhn_box(n, n + 1, n - 1)
    xyz = cb.fields.jitter(xyz, 0.3 / n)
    nv = len(xyz)
    R = util.random_frames(nv, rng)
    H = (1.0 / n) * np.exp(rng.uniform(-1.5, 1.5, (nv, 3)))
    s = (1.0 / n) * np.exp(rng.uniform(-1, 1, nv))
    kind, ma, mb = {"identity": (mao.IDENTITY, None, None), "iso": (mao.ISO, s, None), "aniso": (mao.ANISO, H, R),
                    "logm": (mao.LOGM, None, mao.logm_from_frames(H, R, 0))}[kindname]
    ef = np.zeros(len(ev), np.int32)
    lf = np.zeros(len(tv), np.int32)
    ef[rng.random(len(ev)) < 0.2] |= cb.DONT_SPLIT
    ef[rng.random(len(ev)) < 0.2] |= cb.DONT_COLLAPSE
    ef[rng.random(len(ev)) < 0.1] |= cb.NEED_NOT_SPLIT
    ef[rng.random(len(ev)) < 0.1] |= cb.NEED_NOT_COLLAPSE
    ef[rng.random(len(ev)) < 0.1] |= cb.DONT_SWAP
    lf[rng.random(len(tv)) < 0.3] |= cb.OK_QUALITY
    eo = (rng.random(len(ev)) < 0.9).astype(np.uint8)
    lo = (rng.random(len(tv)) < 0.9).astype(np.uint8)
    want = util.oracle_sweep(kind, xyz, ma, mb, ev, tv, ef, lf, eo, lo, good_quality=0.1)
    p = cb.Part(0)
    p.set_mesh(xyz, ev, tv, edge_owned=eo, elem_owned=lo)
    util.set_part_metric(p, kind, ma, mb)
    for mode in (cb.FP_STRICT, cb.FP_FAST):
        p.set_flags(ef, lf)
        p.sweep(cb.OP_ALL & ~cb.OP_LAYER_CHECK, good_quality=0.1, fp_mode=mode)
        st = check_against(cb, p, want, kind, mode)
        assert st["n_edges_evaluated"] == int(np.count_nonzero(
            ~((ef & (cb.DONT_SPLIT | cb.NEED_NOT_SPLIT)) != 0) | ~((ef & (cb.DONT_COLLAPSE | cb.NEED_NOT_COLLAPSE)) != 0)))
    p.close()


def _shuffled_hub_mesh(cb, rng, n=9):
    """A jittered box whose entities arrive in an order and orientation unrelated to the vertex numbering (what an
    unstructured mesh looks like to the row layout), plus a hub: vertex 0 is made the FIRST vertex of 100 extra edges and 70
    extra tets, so its rows must be cut (kRowMax = 32 in mag_rows.cuh)."""
    xyz, ev, tv = cb.boxmesh.kuhn_box(n, n, n + 1)
    xyz = cb.fields.jitter(xyz, 0.3 / n)
    xyz += (0.05 / n) * (rng.random(xyz.shape) - 0.5)      # boundary vertices too: no hub tet is exactly flat (volume = rounding noise)
    nv = len(xyz)
    hub_e = np.stack([np.zeros(100, np.int64), rng.choice(np.arange(1, nv), 100, replace=False)], axis=1)
    hub_t = np.concatenate([np.zeros((70, 1), np.int64), np.stack([rng.choice(np.arange(1, nv), 3, replace=False) for _ in range(70)])], axis=1)
    flip = rng.random(len(ev)) < 0.5                       # box entities: random first vertex; hub entities keep vertex 0 first
    ev[flip] = ev[flip][:, ::-1]
    rot = rng.integers(0, 3, len(tv))                      # even permutations keep the orientation
    even = np.array([[0, 1, 2, 3], [1, 2, 0, 3], [3, 0, 2, 1]])
    tv = np.take_along_axis(tv, even[rot], axis=1)
    ev = np.concatenate([ev, hub_e]).astype(np.int32)
    tv = np.concatenate([tv, hub_t]).astype(np.int32)
    ev = np.ascontiguousarray(ev[rng.permutation(len(ev))])
    tv = np.ascontiguousarray(tv[rng.permutation(len(tv))])
    return xyz, ev, tv


def test_row_layout_invariants(cb):
    """mag_get_row_layout: every entity sits in exactly one slot, under its first vertex, with its other vertices in the
    caller's order and its ownership bit; slices are as wide as their longest row; no row is longer than 32."""
    rng = np.random.default_rng(5)
    xyz, ev, tv = _shuffled_hub_mesh(cb, rng)
    eo = (rng.random(len(ev)) < 0.8).astype(np.uint8)
    lo = (rng.random(len(tv)) < 0.8).astype(np.uint8)
    p = cb.Part(0)
    p.set_mesh(xyz, ev, tv, edge_owned=eo, elem_owned=lo)
    for which, conn, owned in ((0, ev, eo), (1, tv, lo)):
        lay = p.row_layout(which)
        sl, off, anchor = lay["slot"], lay["slice_off"], lay["anchor"]
        assert lay["slots"] == off[-1] and lay["slices"] == (lay["rows"] + 31) // 32
        ids = sl[:, -1]
        used = np.nonzero(ids >= 0)[0]
        assert np.array_equal(np.sort(ids[used]), np.arange(len(conn))), "every entity exactly once"
        slice_of = np.searchsorted(off, used, side="right") - 1
        lane = (used - off[slice_of]) % 32
        a = anchor[32 * slice_of + lane]
        e = ids[used]
        assert np.array_equal(a, conn[e, 0]), "filed under the first vertex"
        assert np.array_equal(sl[used, 0] & 0x7fffffff, conn[e, 1])
        assert np.array_equal((sl[used, 0] >= 0).astype(np.uint8), owned[e])
        if which:
            assert np.array_equal(sl[used, 1], conn[e, 2]) and np.array_equal(sl[used, 2], conn[e, 3])
        width = np.diff(off) // 32
        assert width.min() >= 1 and width.max() <= 32
        k = (used - off[slice_of]) // 32
        assert np.array_equal(np.bincount(slice_of, minlength=lay["slices"]) > 0, np.ones(lay["slices"], bool))
        assert np.array_equal(np.maximum.reduceat(k[np.argsort(slice_of, kind="stable")],
                                                  np.searchsorted(np.sort(slice_of), np.arange(lay["slices"]))) + 1, width)
        # an anchor's entities keep the caller's order along its row(s)
        hub = np.nonzero(a == 0)[0]
        assert len(hub) >= (100 if which == 0 else 70)
        rows_hub = len(set(zip(slice_of[hub].tolist(), lane[hub].tolist())))
        assert rows_hub >= (4 if which == 0 else 3)
    p.close()


@pytest.mark.parametrize("kindname", ["iso", "aniso", "logm"])
def test_shuffled_hub_mesh_vs_oracle(cb, kindname):
    """The whole sweep on a mesh whose entity order and orientation are unrelated to the vertex numbering."""
    from oracle import mao
    rng = np.random.default_rng(6)
    xyz, ev, tv = _shuffled_hub_mesh(cb, rng)
    nv = len(xyz)
    R = util.random_frames(nv, rng)
    H = (1.0 / 9) * np.exp(rng.uniform(-1.0, 1.0, (nv, 3)))
    s = (1.0 / 9) * np.exp(rng.uniform(-1, 1, nv))
    kind, ma, mb = {"iso": (mao.ISO, s, None), "aniso": (mao.ANISO, H, R), "logm": (mao.LOGM, None, mao.logm_from_frames(H, R, 0))}[kindname]
    ef = np.zeros(len(ev), np.int32)
    lf = np.zeros(len(tv), np.int32)
    ef[rng.random(len(ev)) < 0.2] |= cb.DONT_SPLIT
    ef[rng.random(len(ev)) < 0.2] |= cb.NEED_NOT_COLLAPSE
    lf[rng.random(len(tv)) < 0.3] |= cb.OK_QUALITY
    eo = (rng.random(len(ev)) < 0.9).astype(np.uint8)
    lo = (rng.random(len(tv)) < 0.9).astype(np.uint8)
    want = util.oracle_sweep(kind, xyz, ma, mb, ev, tv, ef, lf, eo, lo, good_quality=0.1)
    p = cb.Part(0)
    p.set_mesh(xyz, ev, tv, edge_owned=eo, elem_owned=lo)
    util.set_part_metric(p, kind, ma, mb)
    for mode in (cb.FP_STRICT, cb.FP_FAST):
        p.set_flags(ef, lf)
        p.sweep(cb.OP_ALL & ~cb.OP_LAYER_CHECK, good_quality=0.1, fp_mode=mode)
        check_against(cb, p, want, kind, mode)
        p.clear_flags()                                   # and with the incoming words all zero (not materialised)
        p.sweep(cb.OP_ALL & ~cb.OP_LAYER_CHECK, good_quality=0.1, fp_mode=mode)
        want0 = util.oracle_sweep(kind, xyz, ma, mb, ev, tv, None, None, eo, lo, good_quality=0.1)
        check_against(cb, p, want0, kind, mode)
    p.close()


def test_reference_named_entry_points(cb):
    """The host mirror's reference-named calls (markEdgesToSplit ... getMaximumEdgeLength) one at a time, and the
    second-sweep behaviour: NEED_NOT_* / OK_QUALITY set by the first sweep make the second one skip (maAdapt.cc:311)."""
    from oracle import mao
    n = 8
    xyz, ev, tv = cb.boxmesh.kuhn_box(n, n, n)
    h, R = cb.fields.shock_rotating(xyz, 1.0 / n)
    want = util.oracle_sweep(mao.ANISO, xyz, h, R, ev, tv)
    p = cb.Part(0)
    p.set_mesh(xyz, ev, tv)
    p.set_size_field_aniso(h, R)
    assert p.markEdgesToSplit() == want["n_split"]
    ef, lf = p.flags()       # a single mark touches only its own two bits
    assert np.array_equal(ef, want["edge_flags"] & (cb.SPLIT | cb.NEED_NOT_SPLIT)) and not lf.any()
    assert p.markEdgesToCollapse() == want["n_collapse"]
    assert p.markBadQuality() == want["n_bad"]
    ef, lf = p.flags()
    assert np.array_equal(ef, want["edge_flags"]) and np.array_equal(lf, want["elem_flags"])
    assert p.getMinQuality() == want["min_quality"]
    assert p.getMaximumEdgeLength() == want["max_length"]
    assert np.array_equal(p.getEdgeLengthsInMetricSpace(), want["lengths"])
    import ctypes
    libm = ctypes.CDLL("libm.so.6")
    libm.cbrt.restype, libm.cbrt.argtypes = ctypes.c_double, [ctypes.c_double]
    assert np.array_equal(p.getLinearQualitiesInMetricSpace(), np.array([libm.cbrt(float(q)) for q in want["qualities"]]))
    # ma::getAverageEdgeLength (maSize.cc:654-671): identity measure, every edge of the part whether owned or not
    with pytest.raises(cb.MagError):
        p.getAverageEdgeLength()
    eo = (np.arange(len(ev)) % 3 != 0).astype(np.uint8)
    p2 = cb.Part(0)
    p2.set_mesh(xyz, ev, tv, edge_owned=eo)
    p2.set_size_field_identity()
    L_id = mao.edge_lengths(mao.IDENTITY, xyz, None, None, ev)
    serial = 0.0
    for v in L_id:
        serial += float(v)
    assert abs(p2.getAverageEdgeLength() - serial / len(ev)) <= 1e-13 * serial / len(ev)
    p2.close()
    # a second markEdgesToSplit on the marked mesh: the reference asserts because SPLIT is still set
    with pytest.raises(cb.MagError) as ei:
        p.markEdgesToSplit()
    assert ei.value.code == 3
    # clear only the true flags (what ma::refine / unMarkBadQuality do) and re-mark: only un-decided entities are evaluated
    p.set_flags(ef, lf)
    p.clearFlagFromDimension(cb.SPLIT | cb.COLLAPSE, 1)      # ma::clearFlagFromDimension (maAdapt.cc:139-147)
    p.unMarkBadQuality()                                      # maShape.cc:138-150
    ef2, lf2 = p.flags()
    assert np.array_equal(ef2, ef & ~(cb.SPLIT | cb.COLLAPSE)) and np.array_equal(lf2, lf & ~cb.BAD_QUALITY)
    with pytest.raises(cb.MagError):
        p.clearFlagFromDimension(cb.SPLIT, 2)                 # a 3-D part holds words on edges and elements only
    p.sweep(cb.OP_MARK_SPLIT | cb.OP_MARK_COLLAPSE | cb.OP_MARK_BAD)
    st = p.stats()
    assert (st["n_split"], st["n_collapse"], st["n_bad"]) == (want["n_split"], want["n_collapse"], want["n_bad"])
    assert st["n_edges_evaluated"] == int(np.count_nonzero(
        ((want["edge_flags"] & cb.SPLIT) != 0) | ((want["edge_flags"] & cb.COLLAPSE) != 0)))
    assert st["n_elems_evaluated"] == want["n_bad"]
    p.close()


def test_set_coords_stream_and_reshape(cb):
    """Coordinates replaced under a resident mesh (vertex motion: the packed records, the per-vertex pass and cached
    cavity transforms must follow), a caller-owned CUDA stream, and a context re-used for a part of another shape."""
    import torch
    from oracle import mao
    n = 7
    xyz, ev, tv = cb.boxmesh.kuhn_box(n, n, n)
    x0 = cb.fields.jitter(xyz, 0.2 / n, seed=3)
    x1 = cb.fields.jitter(xyz, 0.3 / n, seed=4)
    h, R = cb.fields.shock_rotating(xyz, 1.0 / n)
    p = cb.Part(0)
    p.set_mesh(x0, ev, tv)
    p.set_size_field_aniso(h, R)
    p.sweep(cb.OP_ALL & ~cb.OP_LAYER_CHECK)
    check_against(cb, p, util.oracle_sweep(mao.ANISO, x0, h, R, ev, tv), mao.ANISO, cb.FP_STRICT)
    off = np.arange(0, len(tv) + 1, 2, dtype=np.int64)
    w0 = p.cavity_quality(off, tv)                       # caches the per-vertex transforms of x0
    p.set_coords(x1)
    p.clear_flags()
    p.sweep(cb.OP_ALL & ~cb.OP_LAYER_CHECK)
    want1 = util.oracle_sweep(mao.ANISO, x1, h, R, ev, tv)
    check_against(cb, p, want1, mao.ANISO, cb.FP_STRICT)
    w1 = p.cavity_quality(off, tv)
    assert np.array_equal(w1, np.minimum.reduceat(want1["qualities"], off[:-1])) and not np.array_equal(w0, w1)
    # the same sweep on a stream the caller owns
    s = torch.cuda.Stream()
    p.set_stream(s.cuda_stream)
    for mode in (cb.FP_STRICT, cb.FP_FAST):
        p.clear_flags()
        p.sweep(cb.OP_ALL & ~cb.OP_LAYER_CHECK, fp_mode=mode)
        check_against(cb, p, want1, mao.ANISO, mode)
    p.set_stream(None)
    # another shape on the same context: smaller mesh, other size-field kind
    xs, es, ts = cb.boxmesh.kuhn_box(3, 4, 2)
    siz = cb.fields.iso_linear(xs, 0.3)
    p.set_mesh(xs, es, ts)
    with pytest.raises(cb.MagError):
        p.sweep()                                        # the size field went with the old mesh
    p.set_size_field_iso(siz)
    p.sweep(cb.OP_ALL & ~cb.OP_LAYER_CHECK)
    check_against(cb, p, util.oracle_sweep(mao.ISO, xs, siz, None, es, ts), mao.ISO, cb.FP_STRICT)
    p.close()


def test_near_threshold_listing(cb):
    """An edge whose metric length is exactly the threshold: listed, and marked as the reference's strict
    comparison decides (1.5 > 1.5 is false) in both modes."""
    xyz = np.array([[0, 0, 0], [1.5, 0, 0], [0, 0.5, 0], [0, 0, 1.0]], dtype=np.float64)
    ev = np.array([[0, 1], [0, 2], [0, 3], [1, 2]], dtype=np.int32)
    tv = np.array([[0, 1, 2, 3]], dtype=np.int32)
    p = cb.Part(0)
    p.set_mesh(xyz, ev, tv)
    p.set_size_field_iso(np.ones(4))
    for mode in (cb.FP_STRICT, cb.FP_FAST):
        p.set_flags(None, None)
        p.sweep(cb.OP_LENGTHS | cb.OP_MARK_SPLIT | cb.OP_MARK_COLLAPSE, fp_mode=mode)
        idx, n = p.near_threshold(0)
        assert n == 2 and sorted(idx.tolist()) == [0, 1]
        ef, _ = p.flags()
        assert ef[0] == cb.NEED_NOT_SPLIT | cb.NEED_NOT_COLLAPSE      # 1.5: neither > 1.5 nor < 0.5
        assert ef[1] == cb.NEED_NOT_SPLIT | cb.NEED_NOT_COLLAPSE      # 0.5: not < 0.5
        assert ef[2] == cb.NEED_NOT_SPLIT | cb.NEED_NOT_COLLAPSE
        assert ef[3] == cb.SPLIT | cb.NEED_NOT_COLLAPSE               # sqrt(2.5) = 1.58
        assert np.array_equal(p.edge_lengths()[:3], [1.5, 0.5, 1.0])
    p.close()


def test_edge_cases_and_errors(cb):
    p = cb.Part(0)
    # sweep before any size field
    p.set_mesh(np.zeros((2, 3)), np.array([[0, 1]], np.int32))
    with pytest.raises(cb.MagError) as ei:
        p.sweep()
    assert ei.value.code == 2
    # a vertex id outside [0, nv): rejected at export (device-side check), the context stays usable
    with pytest.raises(cb.MagError) as ei:
        p.set_mesh(np.zeros((3, 3)), np.array([[0, 1], [1, 3]], np.int32))
    assert ei.value.code == 2 and "vertex ids" in str(ei.value)
    with pytest.raises(cb.MagError):
        p.set_mesh(np.zeros((4, 3)), np.array([[0, 1]], np.int32), np.array([[0, 1, 2, -1]], np.int32))
    with pytest.raises(cb.MagError):
        p.set_mesh(np.zeros((0, 3)), np.array([[0, 1]], np.int32))
    # ... and in the streamed one-call path (the ids are made harmless on the device, the call reports MAG_ERR_ARG)
    xb, eb, tb = cb.boxmesh.kuhn_box(2, 2, 2)
    tb_bad = tb.copy()
    tb_bad[5, 2] = len(xb)
    with pytest.raises(cb.MagError) as ei:
        p.sweep_host(xb, eb, tb_bad, 1, cb.fields.iso_linear(xb, 0.5), None, out_lengths=np.empty(len(eb)), out_qualities=np.empty(len(tb)))
    assert ei.value.code == 2
    # maximum sizes: entity ids are int32 (MDS_ID_TYPE = int); a count the kernels could not index is refused before
    # anything is allocated or read (the arrays behind these pointers are tiny)
    import ctypes as C
    tiny = np.zeros(8, np.int32)
    for ne_big, nt_big in ((2 ** 31 - 1, 0), (0, 2 ** 31 - 1), (2 ** 33, 0)):
        rc = p._L.mag_set_mesh(p._h, 4, np.zeros(12).ctypes.data, ne_big, tiny.ctypes.data, nt_big, tiny.ctypes.data, 0, None, 0, None, None, None)
        assert rc == 2 and b"int32" in p._L.mag_last_error(p._h)
    assert p._L.mag_set_mesh(p._h, -1, None, 0, None, 0, None, 0, None, 0, None, None, None) == 2
    # empty mesh: every count zero, statistics at their initial values (getMinQuality 1, getMaximumEdgeLength 0)
    p.set_mesh(np.zeros((0, 3)), np.zeros((0, 2), np.int32), np.zeros((0, 4), np.int32))
    p.set_size_field_identity()
    p.sweep()
    st = p.stats()
    assert (st["n_split"], st["n_collapse"], st["n_bad"]) == (0, 0, 0)
    assert st["min_quality"] == 1.0 and st["max_length"] == 0.0
    assert len(p.edge_lengths()) == 0 and len(p.qualities()) == 0
    # edges only (a part with no tets), zero-length edge: measure = 0 -> collapse
    xyz = np.array([[0, 0, 0], [1, 0, 0], [1, 0, 0]], dtype=np.float64)
    p.set_mesh(xyz, np.array([[0, 1], [1, 2]], np.int32))
    p.set_size_field_iso(np.full(3, 0.25))
    p.sweep()
    st = p.stats()
    assert (st["n_split"], st["n_collapse"]) == (1, 1)
    assert np.array_equal(p.edge_lengths(), [4.0, 0.0])
    # inverted tet: negative quality, BAD_QUALITY, min quality negative
    xyz = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, -1]], dtype=np.float64)
    p.set_mesh(xyz, np.array([[0, 1]], np.int32), np.array([[0, 1, 2, 3]], np.int32))
    p.set_size_field_identity()
    p.sweep()
    st = p.stats()
    assert st["min_quality"] < 0 and st["n_bad"] == 1
    # a prism that reaches markBadQuality without OK_QUALITY: the reference would call a null table entry
    xyz = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1], [1, 0, 1], [0, 1, 1]], dtype=np.float64)
    p.set_mesh(xyz, np.array([[0, 1]], np.int32), None, np.array([[0, 1, 2, 3, 4, 5]], np.int32))
    p.set_size_field_identity()
    p.sweep(cb.OP_MARK_BAD)
    with pytest.raises(cb.MagError) as ei:
        p.stats()
    assert ei.value.code == 5
    p.set_flags(None, np.array([cb.OK_QUALITY | cb.LAYER], np.int32))
    p.sweep(cb.OP_MARK_BAD | cb.OP_LAYER_CHECK)
    assert p.stats()["n_bad"] == 0 and p.layer_ok()[0][0] == 1
    # bad arguments
    with pytest.raises(cb.MagError):
        p.sweep(ops=1 << 9)
    with pytest.raises(cb.MagError):
        p.sweep(fp_mode=7)
    p.close()
    with pytest.raises(cb.MagError):
        cb.Part(9999)


def test_two_parts_one_gpu_partition_invariance(cb):
    """Two slab parts (two contexts on one device, no NCCL): shared-edge flags agree copy by copy and the owned counts
    add up to the serial sweep of the glued box -- the partition invariance the multi-GPU path relies on."""
    from oracle import mao
    gnx, ny, nz = 12, 7, 6

    def field(xyz):
        f = xyz.copy()
        f[:, 0] /= 2.0
        return cb.fields.shock_rotating(f, 1.0 / ny)
    xyz, ev, tv = cb.boxmesh.kuhn_box(gnx, ny, nz, wx=2.0)
    h, R = field(xyz)
    want = util.oracle_sweep(mao.ANISO, xyz, h, R, ev, tv)
    tot = np.zeros(3, np.int64)
    flags = []
    parts = [cb.boxmesh.slab_part(gnx, ny, nz, 2, r, wx=2.0) for r in range(2)]
    mx, mn = 0.0, 1.0
    for part in parts:
        p = cb.Part(0)
        p.set_mesh(part["xyz"], part["edge_v"], part["tet_v"], edge_owned=part["edge_owned"], elem_owned=part["elem_owned"])
        hh, RR = field(part["xyz"])
        p.set_size_field_aniso(hh, RR)
        p.sweep(fp_mode=cb.FP_FAST)
        st = p.stats()
        tot += [st["n_split"], st["n_collapse"], st["n_bad"]]
        mx, mn = max(mx, st["max_length"]), min(mn, st["min_quality"])
        flags.append(p.flags()[0])
        p.close()
    assert tot.tolist() == [want["n_split"], want["n_collapse"], want["n_bad"]]
    (peer, idx0, _), = parts[0]["links"]
    (_, idx1, _), = parts[1]["links"]
    assert np.array_equal(flags[0][idx0], flags[1][idx1])
    assert abs(mx - want["max_length"]) <= TOL * mx and abs(mn - want["min_quality"]) <= TOL * abs(mn)


@pytest.mark.parametrize("tagged", [False, True])
def test_mixed_part_lean_tets_after_reset_layer(cb, tagged):
    """BASELINE configs[4] in small: on a mixed prism / tet part the device LAYER closure (mag_reset_layer) materialises the
    element flag words -- LAYER | OK_QUALITY on the prisms -- but touches no tet, so the element sweep that follows may still run
    the lean tet kernel (it writes every tet word from zero).  Lengths, qualities, flag words, counts, layer verdicts and lists
    equal bit for bit what the tile kernels return (MAG_LEAN_SWEEP=0); with a user layer tag on some tets the lean kernel must
    not be chosen, and the results are the same again."""
    rng = np.random.default_rng(17)
    xyz, ev, tv, pr = cb.boxmesh.mixed_box(9, 3)
    xyz = cb.fields.jitter(xyz, 0.15 / 9)
    h, R = cb.fields.shock_rotating(xyz, 1.0 / 9)
    tag = None
    if tagged:
        tag = np.zeros(len(pr) + len(tv), np.int32)
        tag[len(pr) + rng.choice(len(tv), 40, replace=False)] = 1
    res = []
    for env in ({}, {"MAG_LEAN_SWEEP": "0"}):
        p = _part_with_env(cb, env)
        p.set_mesh(xyz, ev, tv, prism_v=pr)
        p.set_size_field_aniso(h, R)
        for rep in range(2):                       # the second sweep starts from the state the first one left
            p.clear_flags()
            nl = p.reset_layer(tag)
            p.sweep(cb.OP_ALL, fp_mode=cb.FP_FAST)
        st = p.stats()
        res.append((nl, p.edge_lengths(), p.qualities(), *p.flags(), st, np.sort(p.near_threshold(1)[0]), p.layer_ok()))
        p.close()
    (n0, L0, q0, ef0, lf0, s0, ne0, ok0), (n1, L1, q1, ef1, lf1, s1, ne1, ok1) = res
    assert n0 == n1 == len(pr) + (40 if tagged else 0)
    assert np.array_equal(L0, L1) and np.array_equal(q0, q1) and np.array_equal(ef0, ef1) and np.array_equal(lf0, lf1)
    assert np.array_equal(ne0, ne1) and np.array_equal(ok0[0] if isinstance(ok0, tuple) else ok0, ok1[0] if isinstance(ok1, tuple) else ok1)
    for k in ("n_split", "n_collapse", "n_bad", "n_elems_evaluated", "n_near_threshold", "min_quality", "max_length", "n_layer_unsafe"):
        assert s0[k] == s1[k], k
    assert np.all(lf0[: len(pr)] & cb.LAYER) and s0["n_bad"] > 0


def test_eight_parts_counts_equal_the_one_part_sweep(cb):
    """BASELINE configs[3] in small: a 208 x 40 x 40 box (2.0 M tets) cut into EIGHT x-slabs, each part swept in its own context
    with the lean kernels, against the same box swept as ONE part on the device: the owned counts add up to the one-part counts
    exactly, min quality / max length are the one-part values bit for bit, and every shared edge carries the same word on both
    of its copies -- what the 8-GPU line's global statistics rest on (there the parts are 50 M tets each and the sum cannot be
    checked against anything but itself)."""
    gnx, ny, nz, nparts = 208, 40, 40, 8

    def field(xyz):
        f = xyz.copy()
        f[:, 0] /= float(nparts)
        return cb.fields.shock_rotating(f, 1.0 / ny)
    xyz, ev, tv = cb.boxmesh.kuhn_box(gnx, ny, nz, wx=float(nparts))
    h, R = field(xyz)
    one = cb.Part(0)
    one.set_mesh(xyz, ev, tv)
    one.set_size_field_aniso(h, R)
    one.clear_flags()
    one.sweep(fp_mode=cb.FP_FAST)
    want = one.stats()
    one.close()
    assert want["n_split"] > 10000 and want["n_collapse"] > 10000 and want["n_bad"] > 10000
    tot = np.zeros(3, np.int64)
    mx, mn = 0.0, 1.0
    words, links = [], []
    for r in range(nparts):
        part = cb.boxmesh.slab_part(gnx, ny, nz, nparts, r, wx=float(nparts))
        p = cb.Part(0)
        p.set_mesh(part["xyz"], part["edge_v"], part["tet_v"], edge_owned=part["edge_owned"], elem_owned=part["elem_owned"])
        hh, RR = field(part["xyz"])
        p.set_size_field_aniso(hh, RR)
        p.clear_flags()
        p.sweep(fp_mode=cb.FP_FAST)
        st = p.stats()
        tot += [st["n_split"], st["n_collapse"], st["n_bad"]]
        mx, mn = max(mx, st["max_length"]), min(mn, st["min_quality"])
        words.append(p.flags()[0])
        links.append({peer: idx for peer, idx, _ in part["links"]})
        p.close()
    assert tot.tolist() == [want["n_split"], want["n_collapse"], want["n_bad"]]
    assert mx == want["max_length"] and mn == want["min_quality"]
    for r in range(nparts - 1):
        assert np.array_equal(words[r][links[r][r + 1]], words[r + 1][links[r + 1][r]])


@pytest.mark.parametrize("kindname", ["iso", "aniso", "logm"])
def test_sweep_host_streamed_equals_resident(cb, kindname):
    """mag_sweep_host (export + sweep + results in one streamed call, several slices) returns bit for bit what the
    resident calls return, with incoming flag words and ownership, in both fp modes; the part stays resident."""
    from oracle import mao
    n = 24
    rng = np.random.default_rng(5)
    xyz, ev, tv = cb.boxmesh.kuhn_box(n, n, n)
    xyz = cb.fields.jitter(xyz, 0.3 / n)
    nv = len(xyz)
    R = util.random_frames(nv, rng)
    H = (1.0 / n) * np.exp(rng.uniform(-1.0, 1.0, (nv, 3)))
    s = (1.0 / n) * np.exp(rng.uniform(-1, 1, nv))
    kind, ma, mb = {"iso": (mao.ISO, s, None), "aniso": (mao.ANISO, H, R),
                    "logm": (mao.LOGM, None, mao.logm_from_frames(H, R, 0))}[kindname]
    ef = np.zeros(len(ev), np.int32)
    lf = np.zeros(len(tv), np.int32)
    ef[rng.random(len(ev)) < 0.2] |= cb.DONT_SPLIT
    ef[rng.random(len(ev)) < 0.1] |= cb.NEED_NOT_COLLAPSE
    lf[rng.random(len(tv)) < 0.3] |= cb.OK_QUALITY
    eo = (rng.random(len(ev)) < 0.9).astype(np.uint8)
    lo = (rng.random(len(tv)) < 0.9).astype(np.uint8)
    ops = cb.OP_ALL & ~cb.OP_LAYER_CHECK
    p = cb.Part(0)
    q = cb.Part(0)
    assert len(ev) > 61440 and len(tv) > 61440          # more than one slice each
    for mode in (cb.FP_STRICT, cb.FP_FAST):
        p.set_mesh(xyz, ev, tv, edge_owned=eo, elem_owned=lo)
        util.set_part_metric(p, kind, ma, mb)
        p.set_flags(ef, lf)
        p.sweep(ops, good_quality=0.1, fp_mode=mode)
        st0, L0, q0 = p.stats(), p.edge_lengths(), p.qualities()
        ef0, lf0 = p.flags()
        oL, oq = np.empty(len(ev)), np.empty(len(tv))
        oef, olf = np.empty(len(ev), np.int32), np.empty(len(tv), np.int32)
        st1 = q.sweep_host(xyz, ev, tv, kind, ma, mb, edge_flags=ef, elem_flags=lf, edge_owned=eo, elem_owned=lo,
                           out_lengths=oL, out_qualities=oq, out_edge_flags=oef, out_elem_flags=olf, ops=ops,
                           good_quality=0.1, fp_mode=mode, slice_entities=61440)
        assert np.array_equal(oL, L0), "mode %d: %d lengths differ, max rel %.3g" % (mode, np.count_nonzero(oL != L0), util.rel_err(oL, L0))
        assert np.array_equal(oq, q0), "mode %d: %d qualities differ" % (mode, np.count_nonzero(oq != q0))
        assert np.array_equal(oef, ef0) and np.array_equal(olf, lf0)
        for k in ("n_split", "n_collapse", "n_bad", "n_edges_evaluated", "n_elems_evaluated", "n_near_threshold",
                  "min_quality", "max_length"):
            assert st0[k] == st1[k], k
        assert sorted(q.near_threshold(0)[0].tolist()) == sorted(p.near_threshold(0)[0].tolist())
        # the part is resident: the getters and a plain sweep work on it
        assert np.array_equal(q.edge_lengths(), L0)
        q.set_flags(ef, lf)
        q.sweep(ops, good_quality=0.1, fp_mode=mode)
        assert np.array_equal(q.flags()[0], ef0) and q.stats()["n_bad"] == st0["n_bad"]
    # NULL flag words = zeros, no ownership arrays, outputs optional
    st = q.sweep_host(xyz, ev, tv, kind, ma, mb, ops=cb.OP_MARK_SPLIT | cb.OP_MARK_BAD, good_quality=0.1, fp_mode=cb.FP_FAST)
    p.set_mesh(xyz, ev, tv)
    util.set_part_metric(p, kind, ma, mb)
    p.sweep(cb.OP_MARK_SPLIT | cb.OP_MARK_BAD, good_quality=0.1, fp_mode=cb.FP_FAST)
    assert (st["n_split"], st["n_bad"]) == (p.stats()["n_split"], p.stats()["n_bad"])
    assert np.array_equal(q.flags()[0], p.flags()[0]) and np.array_equal(q.flags()[1], p.flags()[1])
    p.close()
    q.close()


@pytest.mark.parametrize("kindname", ["iso", "aniso", "logm"])
def test_resweep_host_equals_resident(cb, kindname):
    """mag_resweep_host (connectivity resident; coordinates + size field + mark bytes up, mark bytes down) returns what
    mag_set_coords + mag_set_metric_* + mag_set_flags + mag_sweep + mag_get_flags return, restricted to the eight mark
    bits; moved vertices and a changed field are picked up (the cached per-vertex pass is invalidated)."""
    from oracle import mao
    n = 14
    rng = np.random.default_rng(11)
    xyz, ev, tv = cb.boxmesh.kuhn_box(n, n, n)
    xyz = cb.fields.jitter(xyz, 0.3 / n)
    nv = len(xyz)
    ops = cb.OP_ALL & ~cb.OP_LAYER_CHECK
    p, q = cb.Part(0), cb.Part(0)
    p.set_mesh(xyz, ev, tv)
    q.set_mesh(xyz, ev, tv)
    mask = np.int32(cb.Part.MARK_WORD_MASK)
    w = rng.integers(0, 2**31 - 1, 1000).astype(np.int32)
    assert np.array_equal(cb.Part.mark_to_word(cb.Part.word_to_mark(w)), w & mask)
    for it, mode in enumerate((cb.FP_STRICT, cb.FP_FAST, cb.FP_FAST)):
        xyz_i = cb.fields.jitter(xyz, 0.05 / n, seed=100 + it) if it else xyz
        R = util.random_frames(nv, rng)
        H = (1.0 / n) * np.exp(rng.uniform(-1.0, 1.0, (nv, 3)))
        s = (1.0 / n) * np.exp(rng.uniform(-1, 1, nv))
        kind, ma, mb = {"iso": (mao.ISO, s, None), "aniso": (mao.ANISO, H, R),
                        "logm": (mao.LOGM, None, mao.logm_from_frames(H, R, 0))}[kindname]
        ef = np.zeros(len(ev), np.int32)
        lf = np.zeros(len(tv), np.int32)
        ef[rng.random(len(ev)) < 0.2] |= cb.DONT_SPLIT
        ef[rng.random(len(ev)) < 0.1] |= cb.NEED_NOT_COLLAPSE
        lf[rng.random(len(tv)) < 0.3] |= cb.OK_QUALITY
        p.set_coords(xyz_i)
        util.set_part_metric(p, kind, ma, mb)
        p.set_flags(ef, lf)
        p.sweep(ops, good_quality=0.1, fp_mode=mode)
        st0, L0, q0 = p.stats(), p.edge_lengths(), p.qualities()
        ef0, lf0 = p.flags()
        oem, olm = np.empty(len(ev), np.uint8), np.empty(len(tv), np.uint8)
        oL, oq = np.empty(len(ev)), np.empty(len(tv))
        st1 = q.resweep_host(xyz=xyz_i, kind=kind, field_a=ma, field_b=mb, edge_marks=cb.Part.word_to_mark(ef),
                             elem_marks=cb.Part.word_to_mark(lf), out_edge_marks=oem, out_elem_marks=olm,
                             out_lengths=oL, out_qualities=oq, ops=ops, good_quality=0.1, fp_mode=mode)
        assert np.array_equal(oem, cb.Part.word_to_mark(ef0)) and np.array_equal(olm, cb.Part.word_to_mark(lf0))
        assert np.array_equal(oL, L0) and np.array_equal(oq, q0)
        for k in ("n_split", "n_collapse", "n_bad", "n_edges_evaluated", "n_elems_evaluated", "n_near_threshold",
                  "min_quality", "max_length"):
            assert st0[k] == st1[k], k
        # getters on the resident part agree as well
        em, lm = q.mark_bytes()
        assert np.array_equal(em, oem) and np.array_equal(lm, olm)
    # nothing new uploaded: same field, all-zero marks, marks only
    st2 = q.resweep_host(out_edge_marks=oem, out_elem_marks=olm, ops=cb.OP_MARK_SPLIT | cb.OP_MARK_COLLAPSE | cb.OP_MARK_BAD,
                         good_quality=0.1, fp_mode=cb.FP_FAST)
    p.set_flags(None, None)
    p.sweep(cb.OP_MARK_SPLIT | cb.OP_MARK_COLLAPSE | cb.OP_MARK_BAD, good_quality=0.1, fp_mode=cb.FP_FAST)
    assert np.array_equal(oem, cb.Part.word_to_mark(p.flags()[0])) and np.array_equal(olm, cb.Part.word_to_mark(p.flags()[1]))
    assert (st2["n_split"], st2["n_collapse"], st2["n_bad"]) == tuple(p.stats()[k] for k in ("n_split", "n_collapse", "n_bad"))
    # set / get of mark bytes alone
    b = rng.integers(0, 256, len(ev)).astype(np.uint8)
    q.set_mark_bytes(b, None)
    assert np.array_equal(q.mark_bytes()[0], b) and not q.mark_bytes()[1].any()
    assert np.array_equal(q.flags()[0], cb.Part.mark_to_word(b))
    p.close()
    q.close()


def _part_with_env(cb, env):
    """A Part created under the given environment (the library reads its A/B switches in mag_create)."""
    import os
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        return cb.Part(0)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


@pytest.mark.parametrize("kindname", ["iso", "aniso", "logm"])
@pytest.mark.parametrize("mesh", ["lattice", "jittered", "hub"])
def test_lean_kernels_equal_general(cb, kindname, mesh):
    """The lean whole-part kernels (mag_lean.cuh: MAG_FP_FAST, zero incoming flag words, full marking sweep) return bit for
    bit what the general row kernels and the tile kernels of the sub-range path return: lengths, qualities, flag words,
    counts, min / max and the near-threshold lists (the lattice puts a whole edge family exactly on the collapse threshold)."""
    from oracle import mao
    rng = np.random.default_rng(21)
    n = 13
    if mesh == "hub":
        xyz, ev, tv = _shuffled_hub_mesh(cb, rng)
    else:
        xyz, ev, tv = cb.boxmesh.kuhn_box(n, n, n)
        if mesh == "jittered":
            xyz = cb.fields.jitter(xyz, 0.3 / n)
    nv = len(xyz)
    if mesh == "lattice":
        H, R = cb.fields.shock_rotating(xyz, 1.0 / n)
        s = cb.fields.iso_linear(xyz, 1.0 / n)
    else:
        R = util.random_frames(nv, rng)
        H = (1.0 / n) * np.exp(rng.uniform(-1.0, 1.0, (nv, 3)))
        s = (1.0 / n) * np.exp(rng.uniform(-1, 1, nv))
    kind, ma, mb = {"iso": (mao.ISO, s, None), "aniso": (mao.ANISO, H, R),
                    "logm": (mao.LOGM, None, mao.logm_from_frames(H, R, 0))}[kindname]
    eo = (rng.random(len(ev)) < 0.9).astype(np.uint8)
    lo = (rng.random(len(tv)) < 0.9).astype(np.uint8)
    ops = cb.OP_ALL & ~cb.OP_LAYER_CHECK
    res = []
    # lean rows (tets: winner-in-slot kernel); lean rows with the dependent gather of the max-Jacobian transform; general rows; tiles
    for env in ({}, {"MAG_TET_WINNER": "0"}, {"MAG_LEAN_SWEEP": "0"}, {"MAG_LEGACY_SWEEP": "1"}):
        p = _part_with_env(cb, env)
        p.set_mesh(xyz, ev, tv, edge_owned=eo, elem_owned=lo)
        util.set_part_metric(p, kind, ma, mb)
        p.clear_flags()
        p.sweep(ops, good_quality=0.1, fp_mode=cb.FP_FAST)
        st = p.stats()
        res.append((p.edge_lengths(), p.qualities(), p.flags(), st, sorted(p.near_threshold(0)[0].tolist()),
                    sorted(p.near_threshold(1)[0].tolist())))
        p.close()
    L0, q0, (ef0, lf0), st0, ne0, nl0 = res[0]
    if mesh == "lattice" and kindname == "aniso":
        assert st0["n_near_threshold"] > 1000          # the z edges measure exactly 0.5
    for L, q, (ef, lf), st, ne_, nl in res[1:]:
        assert np.array_equal(L, L0), "%d lengths differ, max rel %.3g" % (np.count_nonzero(L != L0), util.rel_err(L, L0))
        assert np.array_equal(q, q0), "%d qualities differ" % np.count_nonzero(q != q0)
        assert np.array_equal(ef, ef0) and np.array_equal(lf, lf0)
        for k in ("n_split", "n_collapse", "n_bad", "n_edges_evaluated", "n_elems_evaluated", "n_near_threshold", "min_quality", "max_length"):
            assert st[k] == st0[k], k
        assert ne_ == ne0 and nl == nl0


def test_baseline_configs_1_and_2_reference_values(cb):
    """BASELINE configs[0] (n = 20, isotropic h = hbar (1 + 2x)) and configs[1] (n = 55, 998,250 tets, planar shock
    layer AnisoSizeField) on the device against the values the compiled reference printed at survey time
    (SURVEY.md 8c): counts exact in both fp modes, min quality / max length bit-exact in strict mode."""
    cases = [(20, "iso", (800, 15710, 0), 0.43199999999999933, 1.6508199830081591),
             (20, "aniso", (26732, 0, 9600), 0.00064551397077518649, 7.0712523453038738),
             (55, "aniso", (528471, None, 151250), 0.0012373498034575676, 8.5789062902035695)]
    p = cb.Part(0)
    for n, field, counts, minq, maxlen in cases:
        xyz, ev, tv = cb.boxmesh.kuhn_box(n, n, n)
        p.set_mesh(xyz, ev, tv)
        if field == "iso":
            p.set_size_field_iso(cb.fields.iso_linear(xyz, 1.0 / n))
        else:
            p.set_size_field_aniso(*cb.fields.shock_planar(xyz, 1.0 / n))
        for mode in (cb.FP_STRICT, cb.FP_FAST):
            p.clear_flags()
            p.sweep(cb.OP_ALL & ~cb.OP_LAYER_CHECK, fp_mode=mode)
            st = p.stats()
            got = (st["n_split"], st["n_collapse"], st["n_bad"])
            assert all(w is None or g == w for g, w in zip(got, counts)), (n, field, got)
            if mode == cb.FP_STRICT:
                assert st["min_quality"] == minq and st["max_length"] == maxlen
            else:
                assert abs(st["min_quality"] - minq) <= TOL * minq and abs(st["max_length"] - maxlen) <= TOL * maxlen
    p.close()


def test_full_size_properties(cb):
    """BASELINE config 3 size (n=203: 50.2 M tets, 58.9 M edges).  The oracle cannot sweep this in seconds, so:
    (1) a seeded random sample of 300k edges and 300k tets is checked against the oracle; (2) fast and strict flags
    and counts are identical, values within 1e-12; (3) counts equal the number of flagged entities; (4) every entity got
    exactly one of the true / false flags; (5) reversing the entity order leaves every per-entity result unchanged."""
    from oracle import mao
    n = 203
    xyz, ev, tv = cb.boxmesh.kuhn_box(n, n, n)
    h, R = cb.fields.shock_rotating(xyz, 1.0 / n)
    p = cb.Part(0)
    p.set_mesh(xyz, ev, tv)
    p.set_size_field_aniso(h, R)
    res = {}
    for mode in (cb.FP_STRICT, cb.FP_FAST):
        p.clear_flags()
        p.sweep(cb.OP_ALL & ~cb.OP_LAYER_CHECK, fp_mode=mode)
        st = p.stats()
        res[mode] = (st, p.edge_lengths(), p.qualities(), *p.flags())
    (s0, L0, q0, ef0, lf0), (s1, L1, q1, ef1, lf1) = res[cb.FP_STRICT], res[cb.FP_FAST]
    assert np.array_equal(ef0, ef1) and np.array_equal(lf0, lf1)
    for k in ("n_split", "n_collapse", "n_bad"):
        assert s0[k] == s1[k]
    assert util.rel_err(L1, L0) < TOL and util.rel_err(q1, q0) < TOL
    assert s0["n_split"] == int(np.count_nonzero(ef0 & cb.SPLIT)) and s0["n_collapse"] == int(np.count_nonzero(ef0 & cb.COLLAPSE))
    assert s0["n_bad"] == int(np.count_nonzero(lf0 & cb.BAD_QUALITY))
    assert np.all(((ef0 & cb.SPLIT) != 0) ^ ((ef0 & cb.NEED_NOT_SPLIT) != 0))
    assert np.all(((ef0 & cb.COLLAPSE) != 0) ^ ((ef0 & cb.NEED_NOT_COLLAPSE) != 0))
    assert np.all(((lf0 & cb.BAD_QUALITY) != 0) ^ ((lf0 & cb.OK_QUALITY) != 0))
    assert s0["max_length"] == L0.max() and s0["min_quality"] == min(1.0, q0.min())
    rng = np.random.default_rng(7)
    se = rng.choice(len(ev), 300000, replace=False)
    stt = rng.choice(len(tv), 300000, replace=False)
    assert np.array_equal(mao.edge_lengths(mao.ANISO, xyz, h, R, ev[se]), L0[se])
    assert np.array_equal(mao.tet_qualities(mao.ANISO, xyz, h, R, tv[stt]), q0[stt])
    # the rows either side of the path at full size: fast weights agree with the reference-order ones, every SPLIT edge
    # gets exactly one split vertex in edge order, and the streamed one-call path returns what the resident calls returned
    w_strict, w_fast = p.element_weights(fp_mode=cb.FP_STRICT), p.element_weights(fp_mode=cb.FP_FAST)
    assert util.rel_err(w_fast, w_strict) < TOL
    assert np.array_equal(mao.tet_weights(mao.ANISO, xyz, h, R, tv[stt[:50000]]), w_strict[stt[:50000]])
    idx, sx, sa, sb = p.split_vertices()
    assert len(idx) == s1["n_split"] and np.array_equal(idx, np.nonzero(ef1 & cb.SPLIT)[0])
    assert np.array_equal(sx[:1000], 0.5 * xyz[ev[idx[:1000], 0]] + 0.5 * xyz[ev[idx[:1000], 1]])
    del sx, sa, sb
    oL, oq = np.empty(len(ev)), np.empty(len(tv))
    oef, olf = np.empty(len(ev), np.int32), np.empty(len(tv), np.int32)
    sth = p.sweep_host(xyz, ev, tv, 2, h, R, out_lengths=oL, out_qualities=oq, out_edge_flags=oef, out_elem_flags=olf,
                       ops=cb.OP_ALL & ~cb.OP_LAYER_CHECK, fp_mode=cb.FP_FAST)
    assert np.array_equal(oL, L1) and np.array_equal(oq, q1) and np.array_equal(oef, ef1) and np.array_equal(olf, lf1)
    assert all(sth[k] == s1[k] for k in ("n_split", "n_collapse", "n_bad", "min_quality", "max_length", "n_near_threshold"))
    del oL, oq, oef, olf
    # entity-order invariance
    p.set_mesh(xyz, ev[::-1].copy(), tv[::-1].copy())   # a new mesh starts with zero flag words
    p.set_size_field_aniso(h, R)
    p.sweep(cb.OP_ALL & ~cb.OP_LAYER_CHECK, fp_mode=cb.FP_FAST)
    st = p.stats()
    assert np.array_equal(p.edge_lengths()[::-1], L1) and np.array_equal(p.qualities()[::-1], q1)
    efr, lfr = p.flags()
    assert np.array_equal(efr[::-1], ef0) and np.array_equal(lfr[::-1], lf0)
    assert (st["n_split"], st["n_bad"]) == (s0["n_split"], s0["n_bad"])
    p.close()


@pytest.mark.parametrize("kindname", ["aniso", "logm"])
def test_uniform_edge_transform_cache(cb, kindname):
    """Near-threshold edges whose two ends carry bit-identical size-field values are re-evaluated with the per-vertex transform
    of k_vertex_uniform (both Gauss points see the same interpolated values: maSize.cc:395-413 / 511-521 at apfShape.cc:123-124's
    swapped shape values).  On a lattice whose field does not vary along z the z edges are such edges AND sit exactly on the
    collapse threshold: their fast-mode lengths and flag words must be the strict sweep's bit for bit (the strict sweep
    evaluates both points in place, without the cache), through the lean rows (aniso) and the tile kernel's queue (LogAniso);
    aniso also against the oracle.  Then the same after the coordinates moved (the cache follows the per-vertex pass)."""
    from oracle import mao
    n = 10
    xyz, ev, tv = cb.boxmesh.kuhn_box(n, n, n)
    H, R = cb.fields.shock_rotating(xyz, 1.0 / n)
    lm = mao.logm_from_frames(H, R, 0)
    p = cb.Part(0)
    p.set_mesh(xyz, ev, tv)
    ops = cb.OP_LENGTHS | cb.OP_MARK_SPLIT | cb.OP_MARK_COLLAPSE
    same_ends = np.all(H[ev[:, 0]] == H[ev[:, 1]], axis=1) & np.all(R[ev[:, 0]] == R[ev[:, 1]], axis=1)
    assert same_ends.sum() >= n * (n + 1) ** 2
    if kindname == "aniso":
        p.set_size_field_aniso(H, R)
    else:
        p.set_size_field_logm(lm)
    for x in (xyz, cb.fields.jitter(xyz, 0.05 / n)):
        if x is not xyz:
            p.set_coords(x)
        out = {}
        for mode in (cb.FP_STRICT, cb.FP_FAST):
            p.clear_flags()
            p.sweep(ops, fp_mode=mode)
            idx, cnt = p.near_threshold(0)
            out[mode] = (p.edge_lengths(), p.flags()[0], np.sort(idx[:cnt]), p.stats())
        (L0, f0, near0, s0), (L1, f1, near1, s1) = out[cb.FP_STRICT], out[cb.FP_FAST]
        assert np.array_equal(f0, f1) and (s0["n_split"], s0["n_collapse"]) == (s1["n_split"], s1["n_collapse"])
        assert np.array_equal(near0, near1)
        if x is xyz:
            assert len(near1) >= n * (n + 1) ** 2 and same_ends[near1].sum() >= n * (n + 1) ** 2   # the z family is listed
        assert np.array_equal(L1[near1], L0[near1])                   # re-evaluated = the reference's own operation order
        assert util.rel_err(L1, L0) < TOL
        if kindname == "aniso":
            assert np.array_equal(L0, mao.edge_lengths(mao.ANISO, x, H, R, ev))
    p.close()


@pytest.mark.parametrize("kindname", ["aniso", "logm", "iso"])
def test_fast_listed_mode(cb, kindname):
    """MAG_FP_FAST_LISTED (the parity rule's own exception: an edge within 1e-12 of a threshold is decided by the fast value and
    LISTED instead of being re-evaluated in strict arithmetic) on a lattice whose z edges sit exactly on the collapse threshold,
    through the lean rows (aniso, iso) and the tile kernel (LogAniso), whole-part and with incoming flag words: the same lengths
    and qualities as MAG_FP_FAST bit for bit, the same list, flag words identical to MAG_FP_FAST outside the list, counts off by
    at most the flips inside it; element flags untouched by the mode."""
    from oracle import mao
    n = 12
    rng = np.random.default_rng(3)
    xyz, ev, tv = cb.boxmesh.kuhn_box(n, n, n)
    H, R = cb.fields.shock_rotating(xyz, 1.0 / n)
    kind, ma, mb = {"iso": (mao.ISO, np.full(len(xyz), 2.0 / n), None), "aniso": (mao.ANISO, H, R),
                    "logm": (mao.LOGM, None, mao.logm_from_frames(H, R, 0))}[kindname]
    p = cb.Part(0)
    p.set_mesh(xyz, ev, tv)
    util.set_part_metric(p, kind, ma, mb)
    ops = cb.OP_ALL & ~cb.OP_LAYER_CHECK
    ef_in = np.where(rng.random(len(ev)) < 0.2, cb.DONT_COLLAPSE, 0).astype(np.int32)
    for incoming in (False, True):
        res = {}
        for mode in (cb.FP_FAST, cb.FP_FAST_LISTED):
            if incoming:
                p.set_flags(ef_in, np.zeros(len(tv), np.int32))
            else:
                p.clear_flags()
            p.sweep(ops, fp_mode=mode)
            res[mode] = (p.stats(), p.edge_lengths(), p.qualities(), *p.flags(), np.sort(p.near_threshold(0)[0]))
        (s0, L0, q0, ef0, lf0, near0), (s1, L1, q1, ef1, lf1, near1) = res[cb.FP_FAST], res[cb.FP_FAST_LISTED]
        assert len(near0) >= n * (n + 1) ** 2 * (0.7 if incoming else 1.0)      # the z edges
        assert np.array_equal(near0, near1) and np.array_equal(q0, q1) and np.array_equal(lf0, lf1) and s0["n_bad"] == s1["n_bad"]
        # the strict re-evaluation of MAG_FP_FAST also replaces the stored length of a listed edge by the strict value
        off = np.ones(len(ev), bool); off[near0] = False
        assert np.array_equal(L0[off], L1[off]) and util.rel_err(L1, L0) < TOL
        de = np.nonzero(ef0 != ef1)[0]
        assert set(de.tolist()) <= set(near0.tolist())
        assert abs(s0["n_collapse"] - s1["n_collapse"]) <= len(de) and abs(s0["n_split"] - s1["n_split"]) <= len(de)
        assert s0["n_edges_evaluated"] == s1["n_edges_evaluated"] and s0["n_near_threshold"] == s1["n_near_threshold"]
    with pytest.raises(cb.sweep.MagError):        # the streamed host calls take the two base modes only
        p.sweep_host(xyz, ev, tv, int(kind), ma, mb, fp_mode=cb.FP_FAST_LISTED)
    p.close()


def test_full_size_loganiso(cb):
    """BASELINE config 3, log-Euclidean variant (makeSizeField(m, sizes, frames, true)) at n=203: the fast path (the
    reference's QR iteration in its second form) against the strict path on every entity -- flags and counts identical,
    lengths and qualities within 1e-12 -- and against the oracle on a seeded sample."""
    from oracle import mao
    n = 203
    xyz, ev, tv = cb.boxmesh.kuhn_box(n, n, n)
    xyz = cb.fields.jitter(xyz, 0.2 / n)
    h, R = cb.fields.shock_rotating(xyz, 1.0 / n)
    p = cb.Part(0)
    p.set_mesh(xyz, ev, tv)
    lm = p.set_size_field_logm_from_frames(h, R, 0, want_logm=True)
    res = {}
    for mode in (cb.FP_STRICT, cb.FP_FAST):
        p.clear_flags()
        p.sweep(cb.OP_ALL & ~cb.OP_LAYER_CHECK, fp_mode=mode)
        res[mode] = (p.stats(), p.edge_lengths(), p.qualities(), *p.flags())
    (s0, L0, q0, ef0, lf0), (s1, L1, q1, ef1, lf1) = res[cb.FP_STRICT], res[cb.FP_FAST]
    assert np.array_equal(ef0, ef1) and np.array_equal(lf0, lf1)
    assert all(s0[k] == s1[k] for k in ("n_split", "n_collapse", "n_bad"))
    assert util.rel_err(L1, L0) < TOL and util.rel_err(q1, q0) < TOL
    rng = np.random.default_rng(11)
    se = rng.choice(len(ev), 100000, replace=False)
    stt = rng.choice(len(tv), 100000, replace=False)
    assert util.rel_err(L1[se], mao.edge_lengths(mao.LOGM, xyz, None, lm, ev[se])) < TOL
    assert util.rel_err(q1[stt], mao.tet_qualities(mao.LOGM, xyz, None, lm, tv[stt])) < TOL
    p.close()


def test_nccl_two_gpus(cb):
    """mag_comm_* over NCCL needs two devices; exercised by bench.py --gpus 2 as well."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import subprocess, sys, os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29611", os.path.join(root, "bench.py"),
                          "--gpus", "2", "--steps", "2", "--warmup", "1", "--cells", "40", "--no-cpu", "--e2e-steps", "1"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    import json
    line = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert line["n_gpus"] == 2 and line["stats"]["n_flag_mismatch"] == 0


def test_nccl_two_part_device(cb):
    """Two slab parts on two GPUs through mag_comm_* (tests/nccl_two_part.py): consistent copies and global statistics
    equal to the serial oracle, injected disagreements counted and resolved by the owner, syncFlag's OR."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import subprocess, sys, os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29613", os.path.join(root, "tests", "nccl_two_part.py")],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "NCCL_TWO_PART_OK" in out.stdout, (out.stdout[-1500:], out.stderr[-3000:])
