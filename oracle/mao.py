"""ctypes binding of the restated C oracle (oracle/ma_oracle.c) -- TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import
this; the product (core_b200/) never does.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_build", "libma_oracle.so")
IDENTITY, ISO, ANISO, LOGM = range(4)
UNIFORM = 100   # ma::UniformRefiner: IdentitySizeField arithmetic, shouldSplit constant true (maSize.h:75-85); marks only

# ma flag bits, ma/maAdapt.h:17-37
SPLIT, DONT_SPLIT, COLLAPSE, DONT_COLLAPSE, CHECKED, BAD_QUALITY, OK_QUALITY = (1 << i for i in range(7))
DONT_SWAP, LAYER = 1 << 9, 1 << 10
NEED_NOT_SPLIT, NEED_NOT_COLLAPSE = 1 << 17, 1 << 18

_lib = None


def _arith(kind):
    """UniformRefiner measures, transforms and weighs like IdentitySizeField (maSize.h:75-85)."""
    return IDENTITY if kind == UNIFORM else kind


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _lib
    if _lib is None:
        src = [os.path.join(_HERE, f) for f in ("ma_oracle.c", "ma_oracle.h")]
        if not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < max(os.path.getmtime(f) for f in src):
            build()
        L = C.CDLL(LIB_PATH)
        vp, i64, i32, f64 = C.c_void_p, C.c_int64, C.c_int32, C.c_double
        L.mao_det3.restype = f64
        L.mao_det3.argtypes = [vp]
        L.mao_eigen_qr.argtypes = [vp, vp, vp, vp]
        L.mao_eigen.argtypes = [vp, vp, vp]
        L.mao_transform_aniso.argtypes = [vp, vp, vp]
        L.mao_transform_logm.argtypes = [vp, vp]
        L.mao_logm_from_frame.argtypes = [C.c_int, vp, vp, vp]
        L.mao_edge_lengths.argtypes = [C.c_int, vp, vp, vp, i64, vp, vp]
        L.mao_tet_qualities.argtypes = [C.c_int, vp, vp, vp, i64, vp, C.c_int, vp]
        L.mao_tri_qualities.argtypes = [C.c_int, vp, vp, vp, i64, vp, C.c_int, vp]
        L.mao_vertex_transforms.argtypes = [C.c_int, vp, vp, i64, vp]
        L.mao_prism_ok.argtypes = [vp, vp, vp]
        L.mao_pyramid_ok.argtypes = [vp, vp, vp]
        L.mao_mark_entities.restype = i64
        L.mao_mark_entities.argtypes = [i64, vp, C.c_int, f64, vp, vp, i32, i32, i32]
        L.mao_tet_weights.argtypes = [C.c_int, vp, vp, vp, i64, vp, f64, f64, vp]
        L.mao_tri_weights.argtypes = [C.c_int, vp, vp, vp, i64, vp, f64, f64, vp]
        L.mao_sliver_codes.argtypes = [C.c_int, vp, vp, vp, i64, vp, vp, f64, vp, vp]
        L.mao_split_vertices.argtypes = [C.c_int, vp, vp, vp, i64, vp, vp, vp, vp]
        L.mao_min_quality.restype = f64
        L.mao_min_quality.argtypes = [i64, vp]
        L.mao_max_length.restype = f64
        L.mao_max_length.argtypes = [i64, vp, vp]
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.int32)


def det3(A):
    A = _f64(A)
    return lib().mao_det3(_p(A))


def eigen(A):
    """apf::eigen restated: returns (vals[3], vecs[3][3] eigenvector j in row j, rc)."""
    A = _f64(A)
    vecs = np.zeros((3, 3))
    vals = np.zeros(3)
    rc = lib().mao_eigen(_p(A), _p(vecs), _p(vals))
    return vals, vecs, rc


def logm_from_frames(h, R, variant=0):
    """Per-vertex logM.  variant 0: LogAnisoSizeField::init (fields); 1: LogMEval (user function)."""
    h = _f64(h).reshape(-1, 3)
    R = _f64(R).reshape(-1, 9)
    out = np.zeros_like(R)
    L = lib()
    for i in range(len(h)):
        L.mao_logm_from_frame(variant, _p(h[i]), _p(R[i]), _p(out[i]))
    return out


def edge_lengths(kind, xyz, ma, mb, edge_v):
    kind = _arith(kind)
    xyz, ma, mb, edge_v = _f64(xyz), _f64(ma), _f64(mb), _i32(edge_v)
    ne = edge_v.size // 2
    out = np.zeros(ne)
    rc = lib().mao_edge_lengths(kind, _p(xyz), _p(ma), _p(mb), ne, _p(edge_v), _p(out))
    assert rc == 1, "eigenQR did not converge / assert in reference"
    return out


def tet_qualities(kind, xyz, ma, mb, tet_v, use_max=True):
    kind = _arith(kind)
    xyz, ma, mb, tet_v = _f64(xyz), _f64(ma), _f64(mb), _i32(tet_v)
    nt = tet_v.size // 4
    out = np.zeros(nt)
    rc = lib().mao_tet_qualities(kind, _p(xyz), _p(ma), _p(mb), nt, _p(tet_v), int(use_max), _p(out))
    assert rc == 1
    return out


def tet_weights(kind, xyz, ma, mb, tet_v, refines_left=None, coarsens_left=0, dim=3):
    """ma::getElementWeight restated (maBalance.cc:21-52,74-81); refines_left None = raw SizeField::getWeight."""
    kind = _arith(kind)
    xyz, ma, mb, tet_v = _f64(xyz), _f64(ma), _f64(mb), _i32(tet_v)
    nt = tet_v.size // 4
    out = np.zeros(nt)
    w_max, w_min = weight_clamps(refines_left, coarsens_left, dim)
    rc = lib().mao_tet_weights(kind, _p(xyz), _p(ma), _p(mb), nt, _p(tet_v), w_max, w_min, _p(out))
    assert rc == 1
    return out


def sliver_codes(kind, xyz, ma, mb, tet_v, face0_v, good_quality=0.027):
    """ma::getSliverCode / matchSliver (maShape.cc:35-120) of every tet; face0_v = vertices of each tet's first face in
    the face's own order.  Returns (codes [nt], match [nt][2] = {rotation, code_index})."""
    kind = _arith(kind)
    xyz, ma, mb, tet_v, face0_v = _f64(xyz), _f64(ma), _f64(mb), _i32(tet_v), _i32(face0_v)
    nt = tet_v.size // 4
    codes, match = np.zeros(nt, np.int32), np.zeros((nt, 2), np.int32)
    rc = lib().mao_sliver_codes(kind, _p(xyz), _p(ma), _p(mb), nt, _p(tet_v), _p(face0_v), float(good_quality), _p(codes), _p(match))
    assert rc == 1
    return codes, match


def tri_weights(kind, xyz, ma, mb, tri_v, refines_left=None, coarsens_left=0):
    """ma::getElementWeight on a 2-D mesh (triangle measure / (1/2), clampForIterations with dimension 2)."""
    kind = _arith(kind)
    xyz, ma, mb, tri_v = _f64(xyz), _f64(ma), _f64(mb), _i32(tri_v)
    nt = tri_v.size // 3
    out = np.zeros(nt)
    w_max, w_min = weight_clamps(refines_left, coarsens_left, 2)
    rc = lib().mao_tri_weights(kind, _p(xyz), _p(ma), _p(mb), nt, _p(tri_v), w_max, w_min, _p(out))
    assert rc == 1
    return out


def weight_clamps(refines_left, coarsens_left, dim=3):
    """clampForIterations bounds (maBalance.cc:41-52)."""
    if refines_left is None:
        return float("inf"), float("-inf")
    return 2.0 ** (dim * refines_left), 4.0 ** (-coarsens_left)


def split_vertices(kind, xyz, ma, mb, edge_v):
    """ma::makeSplitVert restated: (xyz, a, b) of the vertex splitting each edge, (a, b) in the kind's (ma, mb) layout."""
    kind = _arith(kind)
    xyz, ma, mb, edge_v = _f64(xyz), _f64(ma), _f64(mb), _i32(edge_v)
    n = edge_v.size // 2
    oxyz = np.zeros((n, 3))
    oa = np.zeros(n) if kind == ISO else (np.zeros((n, 3)) if kind == ANISO else None)
    ob = np.zeros((n, 9)) if kind in (ANISO, LOGM) else None
    lib().mao_split_vertices(kind, _p(xyz), _p(ma), _p(mb), n, _p(edge_v), _p(oxyz), _p(oa), _p(ob))
    return oxyz, oa, ob


def tri_qualities(kind, xyz, ma, mb, tri_v, use_max=True):
    """measureTriQuality on a 2-D mesh (maQuality.cc:110-136)."""
    kind = _arith(kind)
    xyz, ma, mb, tri_v = _f64(xyz), _f64(ma), _f64(mb), _i32(tri_v)
    nt = tri_v.size // 3
    out = np.zeros(nt)
    rc = lib().mao_tri_qualities(kind, _p(xyz), _p(ma), _p(mb), nt, _p(tri_v), int(use_max), _p(out))
    assert rc == 1
    return out


def vertex_transforms(kind, ma, mb, nv):
    kind = _arith(kind)
    ma, mb = _f64(ma), _f64(mb)
    out = np.zeros((nv, 9))
    lib().mao_vertex_transforms(kind, _p(ma), _p(mb), nv, _p(out))
    return out


def prism_ok(xyz, prism_v):
    xyz, prism_v = _f64(xyz), _i32(prism_v).reshape(-1, 6)
    ok = np.zeros(len(prism_v), dtype=np.int32)
    codes = np.zeros(len(prism_v), dtype=np.int32)
    c = C.c_int(0)
    for i in range(len(prism_v)):
        ok[i] = lib().mao_prism_ok(_p(xyz), _p(prism_v[i]), C.addressof(c))
        codes[i] = c.value
    return ok, codes


def pyramid_ok(xyz, pyr_v):
    xyz, pyr_v = _f64(xyz), _i32(pyr_v).reshape(-1, 5)
    ok = np.zeros(len(pyr_v), dtype=np.int32)
    codes = np.zeros(len(pyr_v), dtype=np.int32)
    c = C.c_int(0)
    for i in range(len(pyr_v)):
        ok[i] = lib().mao_pyramid_ok(_p(xyz), _p(pyr_v[i]), C.addressof(c))
        codes[i] = c.value
    return ok, codes


def mark_entities(value, cmp, thr, flags, owned, true_flag, set_false_flag, all_false_flags=0):
    """ma::markEntities restated; flags (int32) updated in place; returns owned-true count."""
    value = _f64(value)
    assert flags.dtype == np.int32 and flags.flags.c_contiguous
    ow = None if owned is None else np.ascontiguousarray(owned, dtype=np.uint8)
    return lib().mao_mark_entities(value.size, _p(value), cmp, thr, _p(flags), _p(ow),
                                   true_flag, set_false_flag, all_false_flags)


def mark_edges_to_split(lengths, flags, owned=None, kind=None):      # maRefine.cc:395-400
    # IdentitySizeField::shouldSplit / shouldCollapse are constant false (maSize.cc:64-72)
    thr = float("inf") if kind == IDENTITY else (float("-inf") if kind == UNIFORM else 1.5)
    return mark_entities(lengths, 0, thr, flags, owned, SPLIT, NEED_NOT_SPLIT, DONT_SPLIT | NEED_NOT_SPLIT)


def mark_edges_to_collapse(lengths, flags, owned=None, kind=None):   # maCoarsen.cc:287-292
    thr = float("-inf") if kind in (IDENTITY, UNIFORM) else 0.5
    return mark_entities(lengths, 1, thr, flags, owned, COLLAPSE, NEED_NOT_COLLAPSE,
                         DONT_COLLAPSE | NEED_NOT_COLLAPSE)


def mark_bad_quality(q, flags, good_quality, owned=None):  # maShape.cc:132-136
    return mark_entities(q, 1, good_quality, flags, owned, BAD_QUALITY, OK_QUALITY, 0)


def min_quality(q):
    q = _f64(q)
    return lib().mao_min_quality(q.size, _p(q))


def max_length(lengths, owned=None):
    lengths = _f64(lengths)
    ow = None if owned is None else np.ascontiguousarray(owned, dtype=np.uint8)
    return lib().mao_max_length(lengths.size, _p(lengths), _p(ow))
