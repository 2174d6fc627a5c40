"""Seeded fuzz of the restated oracle against the LIVE compiled reference (oracle/_ref, built here from /root/reference):
random box shapes, heavy vertex jitter, random skewed frames with sizes over three decades, every size-field kind --
lengths, qualities (both metric choices), vertex transforms, element weights, sliver codes and split-vertex transfer,
bit for bit.  Complements the committed golden vectors (which travel to the GPU box); skipped where the reference is not built."""
import numpy as np
import pytest

from oracle import mao
import core_b200.fields as fields
import util


def _ref():
    from oracle import refo
    if not refo.available():
        pytest.skip("compiled reference not built here")
    return refo


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_oracle_equals_live_reference_3d(seed, built):
    refo = _ref()
    rng = np.random.default_rng(1000 + seed)
    dims = tuple(int(x) for x in rng.integers(2, 6, 3))
    m = refo.RefMesh.box(*dims, *(float(x) for x in rng.uniform(0.5, 2.0, 3)))
    xyz = fields.jitter(m.export()[0], 0.35 / max(dims), seed=seed)
    m.set_coords(xyz)
    _, ev, et, elv = m.export()
    tv = np.ascontiguousarray(elv[:, :4])
    nv = len(xyz)
    R = util.random_frames(nv, rng, skew=3e-2)
    h = (1.0 / max(dims)) * np.exp(rng.uniform(-3.5, 3.5, (nv, 3)))
    s = (1.0 / max(dims)) * np.exp(rng.uniform(-2.0, 2.0, nv))
    gq = float(rng.uniform(0.02, 0.6))
    for rkind, kind, a, b in ((refo.KIND_IDENTITY, mao.IDENTITY, None, None), (refo.KIND_ISO_FIELD, mao.ISO, s, None),
                              (refo.KIND_ISO_FN, mao.ISO, s, None), (refo.KIND_ANISO_FIELD, mao.ANISO, h, R),
                              (refo.KIND_ANISO_FN, mao.ANISO, h, R), (refo.KIND_LOG_FIELD, mao.LOGM, h, R),
                              (refo.KIND_LOG_FN, mao.LOGM, h, R)):
        m.set_sizefield(rkind, a, b)
        oa, ob = a, b
        if kind == mao.LOGM:
            lm = m.logm()
            assert np.array_equal(mao.logm_from_frames(h, R, 1 if rkind == refo.KIND_LOG_FN else 0), lm)
            oa, ob = None, lm
        assert np.array_equal(mao.edge_lengths(kind, xyz, oa, ob, ev), m.lengths())
        assert np.array_equal(mao.tet_qualities(kind, xyz, oa, ob, tv, True), m.qualities(True))
        assert np.array_equal(mao.tet_qualities(kind, xyz, oa, ob, tv, False), m.qualities(False))
        assert np.array_equal(mao.vertex_transforms(kind, oa, ob, nv), m.vertex_transforms())
        assert np.array_equal(mao.tet_weights(kind, xyz, oa, ob, tv), m.weights())
        assert np.array_equal(mao.tet_weights(kind, xyz, oa, ob, tv, 1, 1), m.weights(1, 1))
        codes, match, f0 = m.sliver_codes(gq)
        oc, om = mao.sliver_codes(kind, xyz, oa, ob, tv, f0, gq)
        assert np.array_equal(oc, codes) and np.array_equal(om, match)
        if rkind in (refo.KIND_ANISO_FIELD, refo.KIND_LOG_FIELD):
            se = rng.choice(len(ev), min(200, len(ev)), replace=False)
            sx, sa, sb = m.split_vertices(np.sort(se))
            ox, oa2, ob2 = mao.split_vertices(kind, xyz, oa, ob, ev[np.sort(se)])
            assert np.array_equal(ox, sx) and np.array_equal(ob2, sb)
            if kind == mao.ANISO:
                assert np.array_equal(oa2, sa)
        r = m.mark(which=15, good_quality=gq)
        L, q = m.lengths(), m.qualities(True)
        ef, lf = np.zeros(len(ev), np.int32), np.zeros(len(tv), np.int32)
        counts = [mao.mark_edges_to_split(L, ef, None, kind), mao.mark_edges_to_collapse(L, ef, None, kind),
                  mao.mark_bad_quality(q, lf, gq)]
        assert counts == [r["n_split"], r["n_collapse"], r["n_bad"]]
        assert np.array_equal(ef, r["edge_flags"]) and np.array_equal(lf, r["elem_flags"])
        assert mao.min_quality(q) == r["min_q"] and mao.max_length(L) == m.max_edge_length()
    m.close()


@pytest.mark.parametrize("seed", [1, 2])
def test_oracle_equals_live_reference_2d(seed, built):
    refo = _ref()
    rng = np.random.default_rng(2000 + seed)
    nx, ny = (int(x) for x in rng.integers(3, 9, 2))
    m = refo.RefMesh.box(nx, ny, 0)
    xyz = m.export()[0]
    inner = (xyz[:, 0] > 1e-9) & (xyz[:, 0] < 1 - 1e-9) & (xyz[:, 1] > 1e-9) & (xyz[:, 1] < 1 - 1e-9)
    xyz[inner, :2] += (0.35 / max(nx, ny)) * (rng.random((int(inner.sum()), 2)) - 0.5)
    m.set_coords(xyz)
    _, ev, et, elv = m.export()
    tri = np.ascontiguousarray(elv[:, :3])
    nv = len(xyz)
    R = util.random_frames(nv, rng, skew=3e-2)
    h = (1.0 / max(nx, ny)) * np.exp(rng.uniform(-2.5, 2.5, (nv, 3)))
    for rkind, kind in ((refo.KIND_ANISO_FIELD, mao.ANISO), (refo.KIND_LOG_FIELD, mao.LOGM), (refo.KIND_IDENTITY, mao.IDENTITY)):
        m.set_sizefield(rkind, h if kind != mao.IDENTITY else None, R if kind != mao.IDENTITY else None)
        oa, ob = (h, R) if kind == mao.ANISO else (None, m.logm()) if kind == mao.LOGM else (None, None)
        assert np.array_equal(mao.edge_lengths(kind, xyz, oa, ob, ev), m.lengths())
        assert np.array_equal(mao.tri_qualities(kind, xyz, oa, ob, tri, True), m.qualities(True))
        assert np.array_equal(mao.tri_qualities(kind, xyz, oa, ob, tri, False), m.qualities(False))
        assert np.array_equal(mao.tri_weights(kind, xyz, oa, ob, tri), m.weights())
        assert np.array_equal(mao.tri_weights(kind, xyz, oa, ob, tri, 1, 2), m.weights(1, 2))
    m.close()
