"""ctypes binding of oracle/_ref/libref_oracle.so -- TEST INFRASTRUCTURE.

The shared object is the UNMODIFIED SCOREC/core reference (compiled in place
from /root/reference by oracle/ref/Makefile) behind the small C driver
oracle/ref/ref_driver.cc.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module; the product
(core_b200/) never does.
"""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libref_oracle.so")

KIND_IDENTITY, KIND_ISO_FIELD, KIND_ANISO_FIELD, KIND_LOG_FIELD, \
    KIND_ANISO_FN, KIND_LOG_FN, KIND_ISO_FN, KIND_UNIFORM = range(8)

# apf::Mesh::Type (apf/apfMesh.h:149-167)
VERTEX, EDGE, TRIANGLE, QUAD, TET, HEX, PRISM, PYRAMID = range(8)

_lib = None


def available():
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB_PATH)
        L.refo_box.restype = C.c_void_p
        L.refo_box.argtypes = [C.c_int] * 3 + [C.c_double] * 3
        L.refo_build.restype = C.c_void_p
        L.refo_build.argtypes = [C.c_int64, C.c_void_p] + [C.c_int64, C.c_void_p] * 3
        L.refo_load.restype = C.c_void_p
        L.refo_load.argtypes = [C.c_char_p, C.c_char_p]
        L.refo_free.argtypes = [C.c_void_p]
        L.refo_counts.argtypes = [C.c_void_p, C.c_void_p]
        L.refo_export.argtypes = [C.c_void_p] * 5
        L.refo_set_coords.argtypes = [C.c_void_p, C.c_void_p]
        L.refo_set_sizefield.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.refo_get_logm.argtypes = [C.c_void_p, C.c_void_p]
        L.refo_lengths.restype = C.c_double
        L.refo_lengths.argtypes = [C.c_void_p, C.c_void_p]
        L.refo_qualities.restype = C.c_double
        L.refo_qualities.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.refo_vertex_transforms.argtypes = [C.c_void_p, C.c_void_p]
        L.refo_edge_transform.argtypes = [C.c_void_p, C.c_int64, C.c_double, C.c_void_p]
        L.refo_mark.argtypes = [C.c_void_p, C.c_int, C.c_double] + [C.c_void_p] * 7
        L.refo_layer_ok.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.refo_max_edge_length.restype = C.c_double
        L.refo_max_edge_length.argtypes = [C.c_void_p]
        L.refo_avg_edge_length.restype = C.c_double
        L.refo_avg_edge_length.argtypes = [C.c_void_p]
        L.refo_eigen.argtypes = [C.c_void_p] * 3
        L.refo_write_smb.argtypes = [C.c_void_p, C.c_char_p]
        L.refo_store_fields.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.refo_weights.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.refo_stats.argtypes = [C.c_void_p] * 4
        L.refo_sliver_codes.argtypes = [C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]
        L.refo_split_vertices.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.int32)


class RefMesh:
    """One reference apf::Mesh2 (MDS) plus an ma::SizeField on it."""

    def __init__(self, handle):
        self.h = handle
        c = np.zeros(8, dtype=np.int64)
        lib().refo_counts(self.h, _p(c))
        self.counts = c
        self.nv, self.ne = int(c[VERTEX]), int(c[EDGE])
        self.nelem = int(c[TET] + c[HEX] + c[PRISM] + c[PYRAMID])
        if self.nelem == 0:          # 2-D mesh: the elements are the faces
            self.nelem = int(c[TRIANGLE] + c[QUAD])

    @classmethod
    def box(cls, nx, ny, nz, wx=1.0, wy=1.0, wz=1.0):
        return cls(lib().refo_box(nx, ny, nz, wx, wy, wz))

    @classmethod
    def build(cls, xyz, tets=None, prisms=None, pyramids=None):
        xyz = _f64(xyz)
        t, p, y = _i32(tets), _i32(prisms), _i32(pyramids)
        n = lambda a, k: 0 if a is None else a.size // k
        return cls(lib().refo_build(xyz.size // 3, _p(xyz), n(t, 4), _p(t),
                                    n(p, 6), _p(p), n(y, 5), _p(y)))

    @classmethod
    def load(cls, model, smb):
        return cls(lib().refo_load(model.encode(), smb.encode()))

    def close(self):
        if self.h:
            lib().refo_free(self.h)
            self.h = None

    def export(self):
        xyz = np.zeros((self.nv, 3))
        ev = np.zeros((self.ne, 2), dtype=np.int32)
        et = np.zeros(self.nelem, dtype=np.int32)
        elv = np.zeros((self.nelem, 8), dtype=np.int32)
        lib().refo_export(self.h, _p(xyz), _p(ev), _p(et), _p(elv))
        return xyz, ev, et, elv

    def set_coords(self, xyz):
        xyz = _f64(xyz)
        assert xyz.size == 3 * self.nv
        lib().refo_set_coords(self.h, _p(xyz))

    def set_sizefield(self, kind, h=None, R=None):
        h, R = _f64(h), _f64(R)
        rc = lib().refo_set_sizefield(self.h, kind, _p(h), _p(R))
        assert rc == 0

    def logm(self):
        out = np.zeros((self.nv, 9))
        assert lib().refo_get_logm(self.h, _p(out)) == 0
        return out

    def lengths(self):
        out = np.zeros(self.ne)
        self.t_lengths = lib().refo_lengths(self.h, _p(out))
        return out

    def qualities(self, use_max=True):
        out = np.zeros(self.nelem)
        self.t_qualities = lib().refo_qualities(self.h, _p(out), int(use_max))
        return out

    def vertex_transforms(self):
        out = np.zeros((self.nv, 9))
        lib().refo_vertex_transforms(self.h, _p(out))
        return out

    def edge_transform(self, edge, xi):
        out = np.zeros(9)
        lib().refo_edge_transform(self.h, edge, xi, _p(out))
        return out

    def mark(self, which=15, good_quality=-1.0, edge_flags=None, elem_flags=None):
        """Runs the reference marks through a fresh ma::Adapt.

        Returns dict(edge_flags, elem_flags, n_split, n_collapse, n_bad, min_q, times)."""
        efi, lfi = _i32(edge_flags), _i32(elem_flags)
        efo = np.zeros(self.ne, dtype=np.int32)
        lfo = np.zeros(self.nelem, dtype=np.int32)
        counts = np.zeros(3, dtype=np.int64)
        minq = np.zeros(1)
        times = np.zeros(4)
        rc = lib().refo_mark(self.h, which, good_quality, _p(efi), _p(lfi),
                             _p(efo), _p(lfo), _p(counts), _p(minq), _p(times))
        assert rc == 0
        return dict(edge_flags=efo, elem_flags=lfo, n_split=int(counts[0]),
                    n_collapse=int(counts[1]), n_bad=int(counts[2]),
                    min_q=float(minq[0]), times=times)

    def layer_ok(self):
        ok = np.zeros(self.nelem, dtype=np.int32)
        codes = np.zeros(self.nelem, dtype=np.int32)
        lib().refo_layer_ok(self.h, _p(ok), _p(codes))
        return ok, codes

    def write_smb(self, path):
        """mds_write_smb; a serial mesh lands in <path minus .smb>0.smb."""
        lib().refo_write_smb(self.h, path.encode())

    def store_fields(self, h, R):
        h, R = _f64(h), _f64(R)
        lib().refo_store_fields(self.h, _p(h), _p(R))

    def weights(self, refines_left=None, coarsens_left=0):
        """SizeField::getWeight per element (refines_left None) or ma::getElementWeight with the given iteration
        counts (maBalance.cc:41-52,74-81)."""
        out = np.zeros(self.nelem)
        if refines_left is None:
            assert lib().refo_weights(self.h, 0, 0, _p(out), None) == 0
        else:
            assert lib().refo_weights(self.h, int(refines_left), int(coarsens_left), None, _p(out)) == 0
        return out

    def layer_weights(self, refines_left=0, coarsens_left=0):
        """Every element as ma::getElementWeights weighs it, layer elements included (maBalance.cc:21-81): (raw getWeight -- of
        the base triangle for a prism --, ma::getElementWeight, first face of every prism in the face's own vertex order)."""
        raw, cl = np.zeros(self.nelem), np.zeros(self.nelem)
        base = np.zeros((self.nelem, 3), np.int32)
        f = lib().refo_layer_weights
        f.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        assert f(self.h, int(refines_left), int(coarsens_left), _p(raw), _p(cl), _p(base)) == 0
        return raw, cl, base

    def stats(self):
        """ma::stats in metric space: (edge lengths of owned edges, cbrt / signed sqrt of owned simplex qualities)."""
        el, lq = np.zeros(self.ne), np.zeros(self.nelem)
        n = np.zeros(2, np.int64)
        assert lib().refo_stats(self.h, _p(el), _p(lq), _p(n)) == 0
        return el[:n[0]].copy(), lq[:n[1]].copy()

    def sliver_codes(self, good_quality=-1.0):
        """ma::getSliverCode / matchSliver of every tet, and the first face's vertices in the face's own order."""
        codes = np.zeros(self.nelem, np.int32)
        match = np.zeros((self.nelem, 2), np.int32)
        f0 = np.zeros((self.nelem, 3), np.int32)
        assert lib().refo_sliver_codes(self.h, float(good_quality), _p(codes), _p(match), _p(f0)) == 0
        return codes, match, f0

    def short_edge(self, max_edge_ratio):
        """ShortEdgeFixer::shouldApply (maShape.cc:188-219) on every element, on the flags the last mark() left on the Adapt:
        (short edge index or -1 per element, element flag words afterwards)."""
        short = np.zeros(self.nelem, np.int32)
        lf = np.zeros(self.nelem, np.int32)
        f = lib().refo_short_edge
        f.restype = C.c_int64
        n = f(self.h, C.c_double(max_edge_ratio), _p(short), _p(lf))
        assert n == self.nelem, n
        return short, lf

    def split_vertices(self, edges):
        """Position and size-field values ma::makeSplitVert gives the vertex splitting each listed edge."""
        edges = np.ascontiguousarray(edges, dtype=np.int64)
        n = len(edges)
        xyz, a, b = np.zeros((n, 3)), np.zeros((n, 3)), np.zeros((n, 9))
        rc = lib().refo_split_vertices(self.h, n, _p(edges), _p(xyz), _p(a), _p(b))
        assert rc == 0, rc
        return xyz, a, b

    def max_edge_length(self):
        return lib().refo_max_edge_length(self.h)

    def avg_edge_length(self):
        return lib().refo_avg_edge_length(self.h)


def eigen(A):
    A = _f64(A)
    vals = np.zeros(3)
    vecs = np.zeros((3, 3))
    lib().refo_eigen(_p(A), _p(vals), _p(vecs))
    return vals, vecs
