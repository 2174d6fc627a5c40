"""The restated oracle (oracle/ma_oracle.c) against golden vectors produced by the compiled, unmodified
reference (tests/golden/make_golden.py), bit for bit.  This is what pins the oracle."""
import numpy as np
import pytest

from oracle import mao
import util


@pytest.mark.parametrize("name", util.golden_cases())
def test_oracle_matches_reference_golden(name, built):
    g = util.load(name)
    kind, ma, mb = util.metric_arrays(g)
    if util.is_2d(g):
        return check_2d(g, kind, ma, mb)
    prism_v, pyr_v, tet_v = util.split_elements(g)
    nns = len(prism_v) + len(pyr_v)
    if kind == mao.LOGM:
        # the logM vertex field itself is part of the reference's output: rebuild it from sizes + frames
        lm = mao.logm_from_frames(g["h"], g["R"], util.logm_variant(g))
        assert np.array_equal(lm, g["logM"]), "logM field differs from the reference's ma_logM"
    L = mao.edge_lengths(kind, g["xyz"], ma, mb, g["edge_v"])
    assert np.array_equal(L, g["lengths"]), "edge lengths are not bit-identical to ma::SizeField::measure"
    q = mao.tet_qualities(kind, g["xyz"], ma, mb, tet_v, True)
    if "qualities" in g:
        assert np.array_equal(q, g["qualities"])
        assert np.array_equal(mao.tet_qualities(kind, g["xyz"], ma, mb, tet_v, False), g["qualities_centroid"])
        assert np.array_equal(mao.vertex_transforms(kind, ma, mb, len(g["xyz"])), g["vertex_Q"])
    # marks: start from the reference's incoming words (plus, on mixed meshes, what ma::Adapt's constructor set)
    ef = g["edge_flags_in"].copy()
    lf = g["elem_flags_in"].copy()
    if nns:
        ef |= g["edge_flags_ctor"]
        lf |= g["elem_flags_ctor"]
    gq = float(g["good_quality"])
    gq = 0.027 if gq < 0 else gq
    qall = np.concatenate([np.zeros(nns), q])
    counts = [mao.mark_edges_to_split(L, ef, None, kind), mao.mark_edges_to_collapse(L, ef, None, kind),
              mao.mark_bad_quality(qall, lf, gq)]
    assert counts == g["counts"].tolist()
    assert np.array_equal(ef, g["edge_flags_out"])
    assert np.array_equal(lf, g["elem_flags_out"])
    if "qualities" in g:
        assert mao.min_quality(q) == float(g["min_q"])
    assert mao.max_length(L) == float(g["max_len"])
    # SURVEY 8f rows, pinned the same way: ma::getElementWeight and ma::makeSplitVert's size-field transfer
    if "weights_raw" in g:
        assert np.array_equal(mao.tet_weights(kind, g["xyz"], ma, mb, tet_v), g["weights_raw"])
        assert np.array_equal(mao.tet_weights(kind, g["xyz"], ma, mb, tet_v, 0, 1), g["weights_r0_c1"])
        if kind != mao.IDENTITY:
            assert g["weights_r0_c1"].min() == 0.25 or g["weights_r0_c1"].max() == 1.0   # the clamp is exercised
    if "split_edges" in g:
        se = g["split_edges"]
        assert np.array_equal(se, np.nonzero(g["edge_flags_out"] & mao.SPLIT)[0])
        sx, sa, sb = mao.split_vertices(kind, g["xyz"], ma, mb, g["edge_v"][se])
        assert np.array_equal(sx, g["split_xyz"]) and np.array_equal(sb, g["split_b"])
        if kind == mao.ANISO:
            assert np.array_equal(sa, g["split_a"])
    if "sliver_codes" in g:   # ma::getSliverCode / matchSliver, both projections exercised across the fixtures
        codes, match = mao.sliver_codes(kind, g["xyz"], ma, mb, tet_v, g["face0_v"], gq)
        assert np.array_equal(codes, g["sliver_codes"]) and np.array_equal(match, g["sliver_match"])
        assert np.all(codes != 0)                                   # the reference asserts a non-zero code
    if nns:
        ok, codes = mao.prism_ok(g["xyz"], prism_v)
        assert np.array_equal(ok, g["layer_ok"][:len(prism_v)])
        assert np.array_equal(codes, g["layer_codes"][:len(prism_v)])
        ok, codes = mao.pyramid_ok(g["xyz"], pyr_v)        # ma::isPyramidOk: the good rotation (maQuality.cc:533-560)
        assert np.array_equal(ok, g["layer_ok"][len(prism_v):nns])
        assert np.array_equal(codes, g["layer_codes"][len(prism_v):nns])


def check_2d(g, kind, ma, mb):
    """2-D meshes: edges as in 3-D, elements are triangles (measureTriQuality, maQuality.cc:110-136)."""
    tri_v = np.ascontiguousarray(g["elem_v"][:, :3])
    L = mao.edge_lengths(kind, g["xyz"], ma, mb, g["edge_v"])
    assert np.array_equal(L, g["lengths"])
    q = mao.tri_qualities(kind, g["xyz"], ma, mb, tri_v, True)
    assert np.array_equal(q, g["qualities"])
    assert np.array_equal(mao.tri_qualities(kind, g["xyz"], ma, mb, tri_v, False), g["qualities_centroid"])
    ef, lf = g["edge_flags_in"].copy(), g["elem_flags_in"].copy()
    counts = [mao.mark_edges_to_split(L, ef, None, kind), mao.mark_edges_to_collapse(L, ef, None, kind),
              mao.mark_bad_quality(q, lf, float(g["good_quality"]))]
    assert counts == g["counts"].tolist()
    assert np.array_equal(ef, g["edge_flags_out"]) and np.array_equal(lf, g["elem_flags_out"])
    assert mao.min_quality(q) == float(g["min_q"]) and mao.max_length(L) == float(g["max_len"])
    # ma::getElementWeight of a triangle (3-point rule, parent measure 1/2) and the split-vertex transfer
    assert np.array_equal(mao.tri_weights(kind, g["xyz"], ma, mb, tri_v), g["weights_raw"])
    assert np.array_equal(mao.tri_weights(kind, g["xyz"], ma, mb, tri_v, 0, 1), g["weights_r0_c1"])
    if "split_edges" in g:
        se = g["split_edges"]
        assert np.array_equal(se, np.nonzero(g["edge_flags_out"] & mao.SPLIT)[0])
        sx, sa, sb = mao.split_vertices(kind, g["xyz"], ma, mb, g["edge_v"][se])
        assert np.array_equal(sx, g["split_xyz"]) and np.array_equal(sb, g["split_b"])
        if kind == mao.ANISO:
            assert np.array_equal(sa, g["split_a"])


def test_layer_closure_flags_golden():
    """LAYER closure + freezeLayer flags of ma::Adapt's constructor (maLayer.cc:11-71), restated with numpy
    from the connectivity alone: every edge of a prism gets LAYER|DONT_COLLAPSE|DONT_SPLIT|DONT_SWAP, every prism
    LAYER|OK_QUALITY."""
    g = util.load("mixed5_shock_rot_aniso")
    prism_v, _, tet_v = util.split_elements(g)
    from core_b200 import boxmesh
    ef, lf = boxmesh.layer_closure_flags(g["edge_v"], prism_v, None, len(tet_v))
    assert np.array_equal(ef, g["edge_flags_ctor"])
    assert np.array_equal(lf, g["elem_flags_ctor"])
    g = util.load("pyrslab_shock_rot_aniso")             # pyramids too: their eight edges join the closure
    prism_v, pyr_v, tet_v = util.split_elements(g)
    assert len(pyr_v) == 36 and len(tet_v) == 0
    ef, lf = boxmesh.layer_closure_flags(g["edge_v"], prism_v, pyr_v, 0)
    assert np.array_equal(ef, g["edge_flags_ctor"])
    assert np.array_equal(lf, g["elem_flags_ctor"])


def test_unsafe_pyramids_golden(built):
    """ma::isPyramidOk on a prism + pyramid slab thrown about until half of the pyramids are unsafe: all three good-rotation
    values (-1, 0, 1) occur; also the unsafe prisms of the same mesh."""
    g = util.load("pyrslab_unsafe_layer")
    prism_v, pyr_v, _ = util.split_elements(g)
    ok, codes = mao.pyramid_ok(g["xyz"], pyr_v)
    assert set(codes.tolist()) == {-1, 0, 1} and 0 < (ok == 0).sum() < len(ok)
    assert np.array_equal(ok, g["layer_ok"][len(prism_v):]) and np.array_equal(codes, g["layer_codes"][len(prism_v):])
    ok, codes = mao.prism_ok(g["xyz"], prism_v)
    assert np.array_equal(ok, g["layer_ok"][:len(prism_v)]) and np.array_equal(codes, g["layer_codes"][:len(prism_v)])


def test_unsafe_prisms_golden(built):
    g = util.load("mixed5_unsafe_layer")
    prism_v, _, _ = util.split_elements(g)
    ok, codes = mao.prism_ok(g["xyz"], prism_v)
    assert (ok == 0).sum() == 8
    assert np.array_equal(ok, g["layer_ok"][:len(prism_v)])
    assert np.array_equal(codes, g["layer_codes"][:len(prism_v)])


def test_eigen_golden_and_reference_kat(built):
    """mth::eigenQR restated: bit-identical to the compiled reference on 206 matrices, and within the
    reference test's own tolerance of the Octave eigenpairs listed in test/eigen_test.cc:13-56."""
    g = dict(np.load(util.GOLDEN + "/eigen.npz"))
    for A, vals, vecs in zip(g["A"], g["vals"], g["vecs"]):
        v, E, rc = mao.eigen(A)
        assert rc == 3 or rc == 1
        assert np.array_equal(v, vals) and np.array_equal(E, vecs)
    octave_l = np.array([[2.214902e-02, 9.112746e-01, 8.145053e+00], [4.661900e-01, 1.298152e+00, 2.055468e+00],
                         [2.321140e-02, 5.742601e-01, 3.021677e+00], [7.303538e-02, 2.154649e+00, 7.177734e+00],
                         [2.627564e-02, 2.587634e+00, 7.214654e+00], [9.493002e-01, 2.404829e+00, 1.008340e+01]])
    for A, l in zip(g["A"][:6], octave_l):
        v, E, _ = mao.eigen(A)
        o = np.argsort(v)
        assert np.sum((v[o] - l) ** 2) < 1e-10          # eigen_test.cc:144
        for j in o:                                      # A e = lambda e
            assert np.allclose(A @ E[j], v[j] * E[j], atol=1e-9)


def test_qr_kat(built):
    """test/qr.cc:84-109 (testEigenQR) and :57-82 (the all-ones matrix of testHessenberg): the restated iteration
    converges, its eigenvector matrix is orthogonal and Q L Q^T reproduces A, both to the reference test's 1e-10."""
    for A in (np.array([[1, 5, 4], [5, 6, 3], [4, 3, 2]], dtype=np.float64), np.ones((3, 3))):
        vals, vecs, rc = mao.eigen(A)
        assert rc in (1, 3)
        Q = vecs.T                                   # eigenvectors are returned as rows (apfMatrix.cc:68-83)
        assert np.max(np.abs(Q @ Q.T - np.eye(3))) < 1e-10
        assert np.max(np.abs(Q @ np.diag(vals) @ Q.T - A)) < 1e-10
    assert np.allclose(sorted(mao.eigen(np.ones((3, 3)))[0]), [0, 0, 3], atol=1e-10)


def test_det_kat(built):
    """test/ma_insphere.cc:13-29: exact integer determinants of the 3x3 minors of the test's 4x4 matrix."""
    M = np.array([[2, 5, 3, 5], [14, 9, 6, 7], [4, 9, 3, 2], [3, 7, 8, 6]], dtype=np.float64)
    for row, want in ((0, 135.0), (1, 145.0), (2, 35.0), (3, -45.0)):
        minor = np.ascontiguousarray(np.delete(np.delete(M, row, axis=0), 0, axis=1))
        assert mao.det3(minor) == want


def test_survey_smoke_values(built):
    """n=20 unit box, iso field h = hbar (1 + 2x): nSplit 800, nCollapse 15710, nBad 0,
    minQ 0.43199999999999933, max length 1.6508199830081591 (measured on the compiled reference)."""
    import core_b200 as cb
    n = 20
    xyz, ev, tv = cb.boxmesh.kuhn_box(n, n, n)
    s = cb.fields.iso_linear(xyz, 1.0 / n)
    r = util.oracle_sweep(mao.ISO, xyz, s, None, ev, tv)
    assert (r["n_split"], r["n_collapse"], r["n_bad"]) == (800, 15710, 0)
    assert r["min_quality"] == 0.43199999999999933
    assert r["max_length"] == 1.6508199830081591
    assert np.cumsum(r["lengths"])[-1] / len(ev) == 0.70227441816624803   # serial sum, as the survey probe did
    h, R = cb.fields.shock_planar(xyz, 1.0 / n)
    r = util.oracle_sweep(mao.ANISO, xyz, h, R, ev, tv)
    assert (r["n_split"], r["n_collapse"], r["n_bad"]) == (26732, 0, 9600)
    assert r["min_quality"] == 0.00064551397077518649
    assert r["max_length"] == 7.0712523453038738


@pytest.mark.parametrize("name", util.golden_cases())
def test_stats_tables_match_ma_stats(name, built, tmp_path):
    """the vectors of ma::stats (maStats.cc:12-45,115-134) rebuilt from per-entity lengths and qualities by
    core_b200.stats, bit for bit, and the table files in measureAnisoStats' format (one `ostream << double` per line)."""
    from core_b200 import stats
    g = util.load(name)
    kind, ma, mb = util.metric_arrays(g)
    two_d = util.is_2d(g)
    simplex = (g["elem_type"] == util.TRIANGLE) | (g["elem_type"] == util.TET)
    if "qualities" in g:
        q = g["qualities"]
    else:
        q = np.zeros(len(simplex))
        q[simplex] = mao.tet_qualities(kind, g["xyz"], ma, mb, util.split_elements(g)[2])
    el = stats.edge_lengths(g["lengths"])
    lq = stats.linear_qualities(q, simplex=simplex, dim=2 if two_d else 3)
    assert np.array_equal(el, g["stats_el"]) and np.array_equal(lq, g["stats_lq"])
    pe, pq = stats.write_linear_tables(str(tmp_path), 0, g["lengths"], q, simplex=simplex, dim=2 if two_d else 3)
    assert pe.endswith("linear_tables/linearETable_0.dat") and pq.endswith("linear_tables/linearQTable_0.dat")
    rows = open(pe).read().split("\n")
    assert len(rows) == len(el) + 1 and rows[-1] == ""
    assert rows[0] == "%g" % el[0] and np.allclose(np.array(rows[:-1], float), el, rtol=1e-5)
    got = np.array([float(x) for x in open(pq).read().split()])   # empty on a mesh without simplex elements
    assert len(got) == len(lq) and np.allclose(got, lq, rtol=1e-5, atol=1e-300)


@pytest.mark.parametrize("name", [n for n in util.golden_cases() if "short_edge_2" in util.load(n)])
def test_short_edge_fixer_restatement_matches_reference(name):
    """ShortEdgeFixer::shouldApply (maShape.cc:188-219), reached in the compiled reference through oracle/ref/ref_shape_shim.cc:
    its answers (golden vectors) equal the few lines restated with numpy on the reference's own lengths and flag words --
    ratio test max/min < maximumEdgeRatio clears BAD_QUALITY, otherwise the FIRST shortest edge in getDownward order."""
    g = util.load(name)
    _, _, tet_v = util.split_elements(g)
    if len(tet_v) == 0:
        pytest.skip("no tets")
    from test_gpu_parity import tet_edge_table
    te = tet_edge_table(g["edge_v"], tet_v)
    BAD = 1 << 5
    bad = (g["elem_flags_out"] & BAD) != 0
    l = g["lengths"][te]
    for ratio in (2.0, 100.0):
        cleared = bad & (l.max(axis=1) / l.min(axis=1) < ratio)
        want = np.where(bad & ~cleared, te[np.arange(len(te)), l.argmin(axis=1)], -1)
        assert np.array_equal(want, g["short_edge_%g" % ratio])
        assert np.array_equal(np.where(cleared, g["elem_flags_out"] & ~BAD, g["elem_flags_out"]), g["short_flags_%g" % ratio])


@pytest.mark.parametrize("name", [n for n in util.golden_cases() if "prism_base_v" in util.load(n)])
def test_prism_weights_restatement_matches_reference(name):
    """ma::getElementWeight weighs a layer prism by its base triangle (maBalance.cc:31-37): the restated triangle measure on the
    reference's own face vertex order reproduces the reference's raw weights bit for bit; with the Input's default layer
    permissions every prism then comes out at exactly 1 (clampForLayerPermissions)."""
    from oracle import mao
    g = util.load(name)
    kind, ma, mb = util.metric_arrays(g)
    base = g["prism_base_v"]
    pr = np.nonzero(base[:, 0] >= 0)[0]
    assert len(pr) > 0
    w = mao.tri_weights(kind, g["xyz"], ma, mb, np.ascontiguousarray(base[pr]))
    assert np.array_equal(w, g["layer_weights_raw"][pr])
    assert np.all(g["layer_weights_r0_c1"][pr] == 1.0)


@pytest.mark.parametrize("name", [n for n in util.golden_cases() if not n.startswith(("tri9x7_iso", "box6_identity", "box6_iso", "box6_uniform"))])
def test_uniform_edge_identity_on_reference_lengths(name, built):
    """The claim behind k_vertex_uniform (core_b200/csrc/mag_kernels.cu), pinned on the compiled reference's own lengths: an edge
    whose two ends carry bit-identical size-field values sees the same interpolated values at both Gauss points (the second point
    uses the first one's shape values exchanged, apfShape.cc:123-124, and a N0 + a N1 commutes), so its measure
    (maSize.cc:158-216) is twice |row0(J) Q_u| with Q_u = getTransform of fl(fl(a N0) + fl(a N1)) -- a function of the vertex
    alone.  Rebuilt here with numpy in the reference's operation order and compared bit for bit with the golden lengths of every
    such edge (the planar shock-layer fields do not vary along y and z; random-frame fixtures have no such edge and only check
    that the selection is empty)."""
    g = util.load(name)
    kind, ma, mb = util.metric_arrays(g)
    if kind not in (mao.ANISO, mao.LOGM):
        pytest.skip("no frames")
    xyz, ev, L = g["xyz"], g["edge_v"], g["lengths"]
    vals = np.concatenate([ma, mb], axis=1) if kind == mao.ANISO else mb
    same = np.all(vals[ev[:, 0]].view(np.int64) == vals[ev[:, 1]].view(np.int64), axis=1)
    if name.startswith("box6_shock_planar"):
        assert same.sum() >= 2 * 6 * 7 * 7          # the y and z families of the 6^3 lattice at least
    if not same.any():
        return
    XI = 0.577350269189626
    N0, N1 = (1.0 - XI) / 2.0, (1.0 + XI) / 2.0
    assert (1.0 - (-XI)) / 2.0 == N1 and (1.0 + (-XI)) / 2.0 == N0
    c = vals * N0 + vals * N1                        # c = 0; c += a N0; c += a N1 (apfElement.cc:109-113), per component
    nv = len(xyz)
    Qu = (mao.vertex_transforms(kind, np.ascontiguousarray(c[:, :3]), np.ascontiguousarray(c[:, 3:]), nv) if kind == mao.ANISO
          else mao.vertex_transforms(kind, None, np.ascontiguousarray(c), nv)).reshape(nv, 3, 3)
    e = ev[same]
    x0, x1 = xyz[e[:, 0]], xyz[e[:, 1]]
    j = x0 * (-0.5) + x1 * 0.5                       # row 0 of the edge Jacobian (apfVectorElement.cc:44-52)
    Q = Qu[e[:, 0]]
    r = [(j[:, 0] * Q[:, 0, k] + j[:, 1] * Q[:, 1, k]) + j[:, 2] * Q[:, 2, k] for k in range(3)]
    ln = np.sqrt((r[0] * r[0] + r[1] * r[1]) + r[2] * r[2])
    assert np.array_equal(ln + ln, L[same])
    assert np.array_equal(Qu[e[:, 0]], Qu[e[:, 1]])
