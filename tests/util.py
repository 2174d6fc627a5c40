"""Shared helpers of the parity tests: golden fixtures, kind mapping, synthetic inputs."""
import glob
import os

import numpy as np

from oracle import mao

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# reference driver kinds (oracle/refo.py) -> restated-oracle kinds
REF_KIND_TO_MAO = {0: mao.IDENTITY, 1: mao.ISO, 2: mao.ANISO, 3: mao.LOGM, 4: mao.ANISO, 5: mao.LOGM, 6: mao.ISO, 7: mao.UNIFORM}
# apf::Mesh::Type
TRIANGLE, TET, PRISM, PYRAMID = 2, 4, 6, 7


def golden_cases():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "*.npz"))
                  if not os.path.basename(p).startswith(("eigen", "mixed5_unsafe", "pyrslab_unsafe", "smb_")))


def load(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


def is_2d(g):
    return bool(np.all(g["elem_type"] == TRIANGLE))


def split_elements(g):
    """(prism_v [np,6], pyr_v [npy,5], tet_v [nt,4]) from the padded golden element table; iteration
    order of dimension 3 is prisms, pyramids, tets, which the fixture must obey."""
    et, ev = g["elem_type"], g["elem_v"]
    pr, py, te = ev[et == PRISM][:, :6], ev[et == PYRAMID][:, :5], ev[et == TET][:, :4]
    order = np.concatenate([np.nonzero(et == PRISM)[0], np.nonzero(et == PYRAMID)[0], np.nonzero(et == TET)[0]])
    assert np.array_equal(order, np.arange(len(et))), "fixture not in prism|pyramid|tet order"
    return (np.ascontiguousarray(pr, np.int32), np.ascontiguousarray(py, np.int32), np.ascontiguousarray(te, np.int32))


def metric_arrays(g):
    """(mao kind, ma, mb) as the restated oracle / mag_set_metric_* take them."""
    kind = REF_KIND_TO_MAO[int(g["kind"])]
    if kind in (mao.IDENTITY, mao.UNIFORM):
        return kind, None, None
    if kind == mao.ISO:
        return kind, g["h"], None
    if kind == mao.ANISO:
        return kind, g["h"], g["R"]
    return kind, None, g["logM"]


def logm_variant(g):
    """0 = LogAnisoSizeField::init from fields (refo kind 3); 1 = LogMEval user function (kind 5)."""
    return 1 if int(g["kind"]) == 5 else 0


def set_part_metric(p, kind, ma, mb):
    if kind == mao.IDENTITY:
        p.set_size_field_identity()
    elif kind == mao.UNIFORM:
        p.set_size_field_uniform_refiner()
    elif kind == mao.ISO:
        p.set_size_field_iso(ma)
    elif kind == mao.ANISO:
        p.set_size_field_aniso(ma, mb)
    else:
        p.set_size_field_logm(mb)


def random_frames(nv, rng, skew=1e-3):
    A = rng.standard_normal((nv, 3, 3))
    Q, _ = np.linalg.qr(A)
    Q[:, :, 2] *= np.sign(np.linalg.det(Q))[:, None]
    return (Q + skew * rng.standard_normal((nv, 3, 3))).reshape(nv, 9)


def rel_err(a, b):
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300))) if len(a) else 0.0


def oracle_sweep(kind, xyz, ma, mb, edge_v, tet_v, edge_flags=None, elem_flags=None, edge_owned=None,
                 elem_owned=None, good_quality=0.027, n_nonsimplex=0, use_max=True):
    """Full sweep through the restated oracle: lengths, qualities, three marks, statistics.
    elem arrays cover [nonsimplex | tets]; non-simplex elements must carry OK_QUALITY."""
    L = mao.edge_lengths(kind, xyz, ma, mb, edge_v)
    q = mao.tet_qualities(kind, xyz, ma, mb, tet_v, use_max)
    ef = np.zeros(len(edge_v), np.int32) if edge_flags is None else np.array(edge_flags, np.int32)
    nel = n_nonsimplex + len(tet_v)
    lf = np.zeros(nel, np.int32) if elem_flags is None else np.array(elem_flags, np.int32)
    qall = np.concatenate([np.zeros(n_nonsimplex), q])
    ns = mao.mark_edges_to_split(L, ef, edge_owned, kind)
    nc = mao.mark_edges_to_collapse(L, ef, edge_owned, kind)
    nb = mao.mark_bad_quality(qall, lf, good_quality, elem_owned)
    return dict(lengths=L, qualities=qall, edge_flags=ef, elem_flags=lf, n_split=ns, n_collapse=nc, n_bad=nb,
                min_quality=mao.min_quality(q), max_length=mao.max_length(L, edge_owned))
