"""Host-side mirror of the reference interface for the MeshAdapt marking / quality sweep.

`Part` is one mesh part resident on one B200 (one `mag_ctx`).  Method names follow the
reference's free functions and `ma::SizeField` so tests read like the reference's own
drivers (test/measureAnisoStats.cc, ma/maStats.cc):

    markEdgesToSplit / markEdgesToCollapse / markBadQuality / getMinQuality
    getMaximumEdgeLength / getEdgeLengthsInMetricSpace / getLinearQualitiesInMetricSpace

Everything runs through the C ABI (include/mag.h); there is no CPU fallback.
"""
import ctypes as C
import numpy as np

from ._lib import lib, MagStats, MagHostPart, MagHostResult, MagHostUpdate, MagHostMarks

# ma/maSize.h:26-27, ma/maInput.cc:32-46
MAXLENGTH = 1.5
MINLENGTH = 0.5
GOOD_QUALITY_3D = 0.027
GOOD_QUALITY_2D = 0.2

# ma/maAdapt.h:17-37
SPLIT, DONT_SPLIT, COLLAPSE, DONT_COLLAPSE, CHECKED, BAD_QUALITY, OK_QUALITY = (1 << i for i in range(7))
SNAP, DONT_SNAP, DONT_SWAP, LAYER, LAYER_BASE, LAYER_TOP = (1 << i for i in range(7, 13))
NEED_NOT_SPLIT, NEED_NOT_COLLAPSE = 1 << 17, 1 << 18

OP_LENGTHS, OP_MARK_SPLIT, OP_MARK_COLLAPSE, OP_QUALITIES, OP_MARK_BAD, OP_LAYER_CHECK = (1 << i for i in range(6))
OP_ALL = 63
OP_LENGTH_SUM = 1 << 6
FP_STRICT, FP_FAST, FP_FAST_LISTED = 0, 1, 2

ERR_NAMES = {1: "MAG_ERR_CUDA", 2: "MAG_ERR_ARG", 3: "MAG_ERR_FLAG_STATE", 4: "MAG_ERR_EIGEN",
             5: "MAG_ERR_NONSIMPLEX", 6: "MAG_ERR_NCCL", 7: "MAG_ERR_INCONSISTENT"}


class MagError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("%s: %s" % (ERR_NAMES.get(code, code), msg))
        self.code = code


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(C.c_void_p)
    # torch tensor (pinned host memory) or anything with data_ptr()
    return C.c_void_p(a.data_ptr())


def _arr(a, dtype, shape=None):
    """Contiguous view of the right dtype without copying when already so (numpy or torch)."""
    if a is None:
        return None
    if not isinstance(a, np.ndarray) and hasattr(a, "data_ptr"):
        return a
    a = np.ascontiguousarray(a, dtype=dtype)
    return a


class Part:
    def __init__(self, device=0):
        self._L = lib()
        h = C.c_void_p()
        rc = self._L.mag_create(C.byref(h), int(device))
        if rc:
            raise MagError(rc, self._L.mag_last_error(None).decode())
        self._h = h
        self.device = device
        self.nv = self.ne = self.nt = self.np_ = self.npy = self.ntri = 0
        self._kind = -1
        self._keep = []

    # ---- plumbing
    def _ck(self, rc):
        if rc:
            raise MagError(rc, self._L.mag_last_error(self._h).decode())

    def close(self):
        if self._h:
            self._L.mag_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream_ptr):
        self._ck(self._L.mag_set_stream(self._h, C.c_void_p(cuda_stream_ptr) if cuda_stream_ptr else None))

    def synchronize(self):
        self._ck(self._L.mag_synchronize(self._h))

    @property
    def nelem(self):
        return self.np_ + self.npy + self.nt + self.ntri

    # ---- export
    def set_mesh(self, xyz, edge_v, tet_v=None, prism_v=None, pyr_v=None, edge_owned=None, elem_owned=None):
        xyz = _arr(xyz, np.float64)
        edge_v, tet_v = _arr(edge_v, np.int32), _arr(tet_v, np.int32)
        prism_v, pyr_v = _arr(prism_v, np.int32), _arr(pyr_v, np.int32)
        edge_owned, elem_owned = _arr(edge_owned, np.uint8), _arr(elem_owned, np.uint8)
        n = lambda a, k: 0 if a is None else int(a.numel() if hasattr(a, "numel") else a.size) // k
        self.nv, self.ne, self.nt = n(xyz, 3), n(edge_v, 2), n(tet_v, 4)
        self.np_, self.npy = n(prism_v, 6), n(pyr_v, 5)
        self.ntri = 0
        self._ck(self._L.mag_set_mesh(self._h, self.nv, _ptr(xyz), self.ne, _ptr(edge_v), self.nt, _ptr(tet_v),
                                      self.np_, _ptr(prism_v), self.npy, _ptr(pyr_v),
                                      _ptr(edge_owned), _ptr(elem_owned)))
        self.synchronize()  # host buffers may be temporaries

    def set_mesh_2d(self, xyz, edge_v, tri_v, edge_owned=None, elem_owned=None):
        """A 2-D part (triangles are the elements; ma::measureTriQuality)."""
        xyz = _arr(xyz, np.float64)
        edge_v, tri_v = _arr(edge_v, np.int32), _arr(tri_v, np.int32)
        edge_owned, elem_owned = _arr(edge_owned, np.uint8), _arr(elem_owned, np.uint8)
        n = lambda a, k: 0 if a is None else int(a.numel() if hasattr(a, "numel") else a.size) // k
        self.nv, self.ne, self.ntri = n(xyz, 3), n(edge_v, 2), n(tri_v, 3)
        self.nt = self.np_ = self.npy = 0
        self._ck(self._L.mag_set_mesh_2d(self._h, self.nv, _ptr(xyz), self.ne, _ptr(edge_v), self.ntri, _ptr(tri_v),
                                         _ptr(edge_owned), _ptr(elem_owned)))
        self.synchronize()

    def set_coords(self, xyz):
        xyz = _arr(xyz, np.float64)
        self._ck(self._L.mag_set_coords(self._h, _ptr(xyz)))
        self.synchronize()

    def set_size_field_identity(self):
        self._kind = 0
        self._ck(self._L.mag_set_metric_identity(self._h))

    def set_size_field_uniform_refiner(self):
        """ma::UniformRefiner (maSize.h:75-85), what ma::runUniformRefinement configures: every edge that may be split is."""
        self._kind = 0
        self._ck(self._L.mag_set_metric_uniform_refiner(self._h))

    def set_size_field_iso(self, size):
        size = _arr(size, np.float64)
        self._kind = 1
        self._ck(self._L.mag_set_metric_iso(self._h, _ptr(size)))
        self.synchronize()

    def set_size_field_aniso(self, h, R):
        h, R = _arr(h, np.float64), _arr(R, np.float64)
        self._kind = 2
        self._ck(self._L.mag_set_metric_aniso(self._h, _ptr(h), _ptr(R)))
        self.synchronize()

    def set_size_field_logm(self, logM):
        logM = _arr(logM, np.float64)
        self._kind = 3
        self._ck(self._L.mag_set_metric_logm(self._h, _ptr(logM)))
        self.synchronize()

    def set_size_field_logm_from_frames(self, h, R, variant=0, want_logm=False):
        """LogAnisoSizeField built from sizes/frames (variant 0, maSize.cc:491-499) or from a user
        function (variant 1, maSize.cc:343-346); the log is the host libm's, as in the reference."""
        h, R = _arr(h, np.float64), _arr(R, np.float64)
        out = np.empty((self.nv, 9)) if want_logm else None
        self._kind = 3
        self._ck(self._L.mag_set_metric_logm_from_frames(self._h, _ptr(h), _ptr(R), int(variant), _ptr(out)))
        return out

    def set_flags(self, edge_flags=None, elem_flags=None):
        ef, lf = _arr(edge_flags, np.int32), _arr(elem_flags, np.int32)
        self._ck(self._L.mag_set_flags(self._h, _ptr(ef), _ptr(lf)))
        self.synchronize()

    def clear_flags(self):
        """Incoming flag words = 0 on every entity (asynchronous device memset)."""
        self._ck(self._L.mag_set_flags(self._h, None, None))

    def reset_layer(self, user_layer_tag=None):
        """ma::resetLayer (ma/maLayer.cc:94-103) on the resident flag words: LAYER on the closure of every prism / pyramid (and
        of every element with a non-zero user tag), synchronised across parts, then the freeze (DONT_* on LAYER edges,
        OK_QUALITY on LAYER elements).  Returns this part's number of layer elements."""
        tag = _arr(user_layer_tag, np.int32)
        n = C.c_int64(0)
        self._ck(self._L.mag_reset_layer(self._h, _ptr(tag), C.byref(n)))
        return int(n.value)

    # ---- sweep + results
    def sweep(self, ops=OP_ALL, max_len=MAXLENGTH, min_len=MINLENGTH, good_quality=GOOD_QUALITY_3D,
              use_max=True, fp_mode=FP_STRICT, reconcile_mask=None):
        """mag_sweep; with reconcile_mask: mag_sweep_reconciled (the part-boundary exchange of those edge bits runs under the
        element sweep)."""
        if reconcile_mask is None:
            self._ck(self._L.mag_sweep(self._h, int(ops), float(max_len), float(min_len), float(good_quality),
                                       int(bool(use_max)), int(fp_mode)))
        else:
            self._ck(self._L.mag_sweep_reconciled(self._h, int(ops), float(max_len), float(min_len), float(good_quality),
                                                  int(bool(use_max)), int(fp_mode), int(reconcile_mask)))

    def sweep_host(self, xyz, edge_v, tet_v, kind, field_a=None, field_b=None, edge_flags=None, elem_flags=None,
                   edge_owned=None, elem_owned=None, out_lengths=None, out_qualities=None, out_edge_flags=None,
                   out_elem_flags=None, ops=OP_ALL & ~OP_LAYER_CHECK, max_len=MAXLENGTH, min_len=MINLENGTH,
                   good_quality=GOOD_QUALITY_3D, use_max=True, fp_mode=FP_STRICT, slice_entities=0):
        """mag_sweep_host: export + sweep + results of a tet part in one streamed call (host buffers, ideally pinned).
        kind: 0 identity, 1 iso (field_a = size), 2 aniso (field_a = h, field_b = R), 3 logm (field_b = logM).
        Returns the statistics dict; outputs land in the out_* buffers that were given."""
        xyz = _arr(xyz, np.float64)
        edge_v, tet_v = _arr(edge_v, np.int32), _arr(tet_v, np.int32)
        fa, fb = _arr(field_a, np.float64), _arr(field_b, np.float64)
        ef, lf = _arr(edge_flags, np.int32), _arr(elem_flags, np.int32)
        eo, lo = _arr(edge_owned, np.uint8), _arr(elem_owned, np.uint8)
        n = lambda a, k: 0 if a is None else int(a.numel() if hasattr(a, "numel") else a.size) // k
        self.nv, self.ne, self.nt = n(xyz, 3), n(edge_v, 2), n(tet_v, 4)
        self.np_ = self.npy = self.ntri = 0
        self._kind = int(kind)
        pv = lambda a: None if a is None else _ptr(a).value
        part = MagHostPart(self.nv, pv(xyz), self.ne, pv(edge_v), self.nt, pv(tet_v), pv(eo), pv(lo), int(kind),
                           pv(fa), pv(fb), pv(ef), pv(lf), int(slice_entities))
        res = MagHostResult(pv(out_lengths), pv(out_qualities), pv(out_edge_flags), pv(out_elem_flags))
        s = MagStats()
        self._ck(self._L.mag_sweep_host(self._h, C.byref(part), C.byref(res), int(ops), float(max_len), float(min_len),
                                        float(good_quality), int(bool(use_max)), int(fp_mode), C.byref(s)))
        return s.as_dict()

    # ---- mark bytes: the 8 bits of a flag word the sweep reads or writes (include/mag.h MAG_MARK_*)
    MARK_WORD_MASK = SPLIT | DONT_SPLIT | COLLAPSE | DONT_COLLAPSE | NEED_NOT_SPLIT | NEED_NOT_COLLAPSE | BAD_QUALITY | OK_QUALITY

    @staticmethod
    def word_to_mark(w):
        u = np.asarray(w).astype(np.uint32)
        return ((u & 0xF) | ((u >> 13) & 0x30) | ((u << 1) & 0xC0)).astype(np.uint8)

    @staticmethod
    def mark_to_word(b):
        u = np.asarray(b).astype(np.uint32)
        return ((u & 0xF) | ((u & 0x30) << 13) | ((u & 0xC0) >> 1)).astype(np.int32)

    def set_mark_bytes(self, edge_marks=None, elem_marks=None):
        e, l = _arr(edge_marks, np.uint8), _arr(elem_marks, np.uint8)
        self._ck(self._L.mag_set_mark_bytes(self._h, _ptr(e), _ptr(l)))
        self.synchronize()

    def mark_bytes(self, edge_out=None, elem_out=None):
        ef = edge_out if edge_out is not None else np.empty(self.ne, dtype=np.uint8)
        lf = elem_out if elem_out is not None else np.empty(self.nelem, dtype=np.uint8)
        self._ck(self._L.mag_get_mark_bytes(self._h, _ptr(ef), _ptr(lf)))
        return ef, lf

    def resweep_host(self, xyz=None, kind=-1, field_a=None, field_b=None, edge_marks=None, elem_marks=None,
                     out_edge_marks=None, out_elem_marks=None, out_lengths=None, out_qualities=None,
                     ops=OP_ALL & ~OP_LAYER_CHECK, max_len=MAXLENGTH, min_len=MINLENGTH,
                     good_quality=GOOD_QUALITY_3D, use_max=True, fp_mode=FP_STRICT):
        """mag_resweep_host: the connectivity stays resident; coordinates (xyz), the size field (kind >= 0 with field_a /
        field_b as in sweep_host) and the incoming mark bytes go up, the outgoing mark bytes (and, opt-in, lengths /
        qualities) come back, in one streamed call.  Returns the statistics dict."""
        xyz = _arr(xyz, np.float64)
        fa, fb = _arr(field_a, np.float64), _arr(field_b, np.float64)
        em, lm = _arr(edge_marks, np.uint8), _arr(elem_marks, np.uint8)
        pv = lambda a: None if a is None else _ptr(a).value
        if int(kind) >= 0:
            self._kind = int(kind)
        upd = MagHostUpdate(pv(xyz), int(kind), pv(fa), pv(fb), pv(em), pv(lm))
        res = MagHostMarks(pv(out_edge_marks), pv(out_elem_marks), pv(out_lengths), pv(out_qualities))
        s = MagStats()
        self._ck(self._L.mag_resweep_host(self._h, C.byref(upd), C.byref(res), int(ops), float(max_len), float(min_len),
                                          float(good_quality), int(bool(use_max)), int(fp_mode), C.byref(s)))
        return s.as_dict()

    def element_weights(self, refines_left=None, coarsens_left=0, fp_mode=FP_STRICT, dim=3):
        """ma::getElementWeights (maBalance.cc:83-97): refines_left None = raw SizeField::getWeight, otherwise clamped
        like clampForIterations with the Adapt's refinesLeft / coarsensLeft."""
        if refines_left is None:
            w_max, w_min = float("inf"), float("-inf")
        else:
            w_max, w_min = 2.0 ** (dim * refines_left), 4.0 ** (-coarsens_left)
        out = np.empty(self.nelem, dtype=np.float64)
        self._ck(self._L.mag_element_weights(self._h, w_max, w_min, int(fp_mode), _ptr(out)))
        return out

    def collapse_quality(self, edges, which_end, use_max=True, fp_mode=FP_STRICT):
        """ma::Collapse's quality test for many candidates (maCollapse.cc:88-113): candidate k collapses end which_end[k] of
        edge edges[k] onto the other end.  Returns (new_worst, old_worst, n_keep): worst quality of the rebuilt tets, of all tets
        around the collapsing vertex, number of rebuilt tets."""
        edges = np.ascontiguousarray(edges, dtype=np.int32)
        which_end = np.ascontiguousarray(which_end, dtype=np.uint8)
        if edges.shape != which_end.shape or edges.ndim != 1:
            raise ValueError("edges and which_end must be 1-D arrays of the same length")
        n = len(edges)
        new_w, old_w, keep = np.empty(n), np.empty(n), np.empty(n, np.int32)
        self._ck(self._L.mag_collapse_quality(self._h, n, _ptr(edges), _ptr(which_end), int(bool(use_max)), int(fp_mode),
                                              _ptr(new_w), _ptr(old_w), _ptr(keep)))
        return new_w, old_w, keep

    def prism_weights(self, base_v, refines_left=None, coarsens_left=0, refine_layer=True, coarsen_layer=True, to_tets=False,
                      fp_mode=FP_STRICT):
        """ma::getElementWeight of every prism (maBalance.cc:21-81): its base triangle's getWeight (base_v [np][3]: the first
        face in the face's own vertex order), clamped like clampForIterations (refines_left None: raw), then the layer
        permissions and the tet conversion factor of ma::Input."""
        base_v = np.ascontiguousarray(base_v, dtype=np.int32)
        if base_v.shape != (self.np_, 3):
            raise ValueError("base_v must be [np][3]")
        if refines_left is None:
            w_max, w_min = float("inf"), float("-inf")
        else:
            w_max, w_min = 2.0 ** (3 * refines_left), 4.0 ** (-coarsens_left)
        out = np.empty(self.np_, dtype=np.float64)
        self._ck(self._L.mag_prism_weights(self._h, _ptr(base_v), w_max, w_min, int(bool(refine_layer)), int(bool(coarsen_layer)),
                                           int(bool(to_tets)), int(fp_mode), _ptr(out)))
        return out

    def cavity_quality(self, offsets, tet_v, use_max=True, fp_mode=FP_STRICT, want_qualities=False):
        """Batch ma::getWorstQuality (maQuality.cc:184-210): worst[k] = min quality over the candidate tets
        tet_v[offsets[k]:offsets[k+1]] (vertex quadruples; the tets need not exist in the mesh)."""
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        tet_v = np.ascontiguousarray(tet_v, dtype=np.int32)
        ncav = len(offsets) - 1
        worst = np.empty(ncav)
        q = np.empty(len(tet_v)) if want_qualities else None
        self._ck(self._L.mag_cavity_quality(self._h, ncav, _ptr(offsets), _ptr(tet_v), int(bool(use_max)), int(fp_mode),
                                            _ptr(worst), _ptr(q)))
        return (worst, q) if want_qualities else worst

    def short_edge_test(self, tet_edges, max_edge_ratio):
        """ShortEdgeFixer::shouldApply over every BAD_QUALITY tet (maShape.cc:188-219): returns (short_edge [nelem] with
        -1 where nothing is to be removed, n_cleared, n_short); BAD_QUALITY is cleared on the resident flags where the
        edge ratio is below max_edge_ratio."""
        te = np.ascontiguousarray(tet_edges, dtype=np.int32)
        out = np.empty(self.nelem, dtype=np.int32)
        nc, ns = C.c_int64(0), C.c_int64(0)
        self._ck(self._L.mag_short_edge_test(self._h, _ptr(te), float(max_edge_ratio), _ptr(out), C.byref(nc), C.byref(ns)))
        return out, nc.value, ns.value

    def sliver_codes(self, face0_v=None, good_quality=0.027, only_bad=False):
        """ma::getSliverCode / matchSliver (maShape.cc:35-120) of every tet (or every BAD_QUALITY tet): (codes [nelem],
        match [nelem][2] = {rotation, code_index}, -1 = no match).  face0_v [nt][3] = the first face's vertices in the
        face's own order (None: the tet's v0, v1, v2)."""
        f0 = None if face0_v is None else np.ascontiguousarray(face0_v, dtype=np.int32)
        codes = np.empty(self.nelem, dtype=np.int32)
        match = np.empty((self.nelem, 2), dtype=np.int32)
        self._ck(self._L.mag_sliver_codes(self._h, _ptr(f0), float(good_quality), int(bool(only_bad)), _ptr(codes), _ptr(match)))
        return codes, match

    def split_vertices(self, fp_mode=FP_STRICT):
        """ma::makeSplitVert for every SPLIT-marked edge, in edge order: (edge_idx, xyz, field_a, field_b) with
        (field_a, field_b) in the layout of the resident size field (iso: size[n]; aniso: h[n][3], R[n][9];
        logm: None, logM[n][9]; identity: None, None)."""
        n = C.c_int64(0)
        self._ck(self._L.mag_split_vertices(self._h, int(fp_mode), 0, C.byref(n), None, None, None, None))
        k = n.value
        idx = np.empty(k, dtype=np.int32)
        xyz = np.empty((k, 3))
        fa = {1: np.empty(k), 2: np.empty((k, 3))}.get(self._kind)
        fb = np.empty((k, 9)) if self._kind in (2, 3) else None
        if k:
            self._ck(self._L.mag_split_vertices(self._h, int(fp_mode), k, C.byref(n), _ptr(idx), _ptr(xyz), _ptr(fa), _ptr(fb)))
        return idx, xyz, fa, fb

    def stats(self):
        s = MagStats()
        self._ck(self._L.mag_get_stats(self._h, C.byref(s)))
        return s.as_dict()

    def edge_lengths(self, out=None):
        out = np.empty(self.ne, dtype=np.float64) if out is None else out
        self._ck(self._L.mag_get_edge_lengths(self._h, _ptr(out)))
        return out

    def qualities(self, out=None):
        out = np.empty(self.nelem, dtype=np.float64) if out is None else out
        self._ck(self._L.mag_get_qualities(self._h, _ptr(out)))
        return out

    def flags(self, edge_out=None, elem_out=None):
        ef = np.empty(self.ne, dtype=np.int32) if edge_out is None else edge_out
        lf = np.empty(self.nelem, dtype=np.int32) if elem_out is None else elem_out
        self._ck(self._L.mag_get_flags(self._h, _ptr(ef), _ptr(lf)))
        return ef, lf

    def layer_ok(self):
        n = self.np_ + self.npy
        ok, codes = np.empty(n, dtype=np.int32), np.empty(n, dtype=np.int32)
        self._ck(self._L.mag_get_layer_ok(self._h, _ptr(ok), _ptr(codes)))
        return ok, codes

    def near_threshold(self, which=0, cap=1 << 20):
        idx = np.empty(cap, dtype=np.int64)
        n = C.c_int64(0)
        self._ck(self._L.mag_get_near_threshold(self._h, which, _ptr(idx), cap, C.byref(n)))
        return idx[:min(n.value, cap)].copy(), n.value

    def timing_begin(self, max_sweeps):
        self._ck(self._L.mag_timing_begin(self._h, int(max_sweeps)))
        self._tslots = int(max_sweeps)

    def timing_read(self):
        """rows of (vertex_ms, edge_ms, elem_ms), one per sweep since timing_begin."""
        ms = np.zeros((max(self._tslots, 1), 3), dtype=np.float32)
        n = C.c_int(0)
        self._ck(self._L.mag_timing_read(self._h, _ptr(ms), C.byref(n)))
        return ms[:n.value].astype(np.float64)

    def launch_count(self):
        return int(self._L.mag_launch_count(self._h))

    def row_layout(self, which):
        """The anchor-row device layout of the edges (which = 0) or tets (1): dict(rows, slices, slots, anchor [32*slices],
        slice_off [slices+1], slot [slots, 2 or 4]) -- see mag_get_row_layout / core_b200/csrc/mag_rows.cuh."""
        counts = np.zeros(3, np.int64)
        self._ck(self._L.mag_get_row_layout(self._h, int(which), _ptr(counts), None, None, None))
        rows, slices, slots = (int(x) for x in counts)
        anchor = np.empty(32 * slices, np.int32)
        off = np.empty(slices + 1 if slices else 0, np.int32)
        sl = np.empty((slots, 4 if which else 2), np.int32)
        self._ck(self._L.mag_get_row_layout(self._h, int(which), _ptr(counts), _ptr(anchor), _ptr(off), _ptr(sl)))
        return dict(rows=rows, slices=slices, slots=slots, anchor=anchor, slice_off=off, slot=sl)

    # ---- the reference's entry points, same names and meaning
    def markEdgesToSplit(self, fp_mode=FP_STRICT):          # ma/maRefine.cc:395-400
        self.sweep(OP_MARK_SPLIT, fp_mode=fp_mode)
        return self.stats()["n_split"]

    def markEdgesToCollapse(self, fp_mode=FP_STRICT):       # ma/maCoarsen.cc:287-292
        self.sweep(OP_MARK_COLLAPSE, fp_mode=fp_mode)
        return self.stats()["n_collapse"]

    def markBadQuality(self, good_quality=GOOD_QUALITY_3D, fp_mode=FP_STRICT):   # ma/maShape.cc:132-136
        self.sweep(OP_MARK_BAD, good_quality=good_quality, fp_mode=fp_mode)
        return self.stats()["n_bad"]

    def getMinQuality(self, fp_mode=FP_STRICT):             # ma/maShape.cc:152-169
        self.sweep(OP_QUALITIES, fp_mode=fp_mode)
        return self.stats()["min_quality"]

    def getMaximumEdgeLength(self, fp_mode=FP_STRICT):      # ma/maSize.cc:673-691
        self.sweep(OP_LENGTHS, fp_mode=fp_mode)
        return self.stats()["max_length"]

    def getAverageEdgeLength(self, fp_mode=FP_STRICT):      # ma/maSize.cc:654-671
        """Mean PHYSICAL edge length over every edge of the part (the reference measures with an IdentitySizeField whatever
        the adapt's size field is): call with the identity size field set."""
        if self._kind != 0:
            raise MagError(2, "getAverageEdgeLength measures with ma::IdentitySizeField: set_size_field_identity() first")
        self.sweep(OP_LENGTHS | OP_LENGTH_SUM, fp_mode=fp_mode)
        return self.stats()["sum_length"] / self.ne

    def getEdgeLengthsInMetricSpace(self, fp_mode=FP_STRICT):   # ma/maStats.cc:33-45
        self.sweep(OP_LENGTHS, fp_mode=fp_mode)
        return self.edge_lengths()

    def getLinearQualitiesInMetricSpace(self, fp_mode=FP_STRICT):   # ma/maStats.cc:12-31 (cbrt in 3D)
        self.sweep(OP_QUALITIES, fp_mode=fp_mode)
        from . import stats   # the host's libm cbrt, as the reference (numpy's own differs in the last bit)
        return stats.linear_qualities(self.qualities()[self.np_ + self.npy:], dim=2 if self.ntri else 3)

    def clearFlagFromDimension(self, flag, dimension):          # ma/maAdapt.cc:139-147
        self._ck(self._L.mag_clear_flag(self._h, int(dimension), int(flag)))

    def unMarkBadQuality(self):                                 # ma/maShape.cc:138-150
        self.clearFlagFromDimension(BAD_QUALITY, 2 if self.ntri else 3)

    # ---- multi-GPU (NCCL)
    @staticmethod
    def comm_unique_id():
        buf = (C.c_char * 128)()
        rc = lib().mag_comm_unique_id(buf)
        if rc:
            raise MagError(rc, lib().mag_last_error(None).decode())
        return bytes(buf)

    def comm_init(self, nranks, rank, unique_id):
        buf = (C.c_char * 128).from_buffer_copy(unique_id)
        self._ck(self._L.mag_comm_init(self._h, nranks, rank, buf))

    def set_edge_links(self, links):
        """links: list of (peer, idx int32 array, peer_owns uint8 array or None)."""
        k = len(links)
        peers = (C.c_int32 * k)(*[int(l[0]) for l in links])
        ns = (C.c_int64 * k)(*[int(len(l[1])) for l in links])
        idx = [np.ascontiguousarray(l[1], dtype=np.int32) for l in links]
        own = [None if l[2] is None else np.ascontiguousarray(l[2], dtype=np.uint8) for l in links]
        idxp = (C.c_void_p * k)(*[a.ctypes.data for a in idx])
        ownp = (C.c_void_p * k)(*[None if a is None else a.ctypes.data for a in own])
        self._ck(self._L.mag_set_edge_links(self._h, k, peers, ns, idxp, ownp))

    def reconcile_edge_flags(self, mask):
        self._ck(self._L.mag_reconcile_edge_flags(self._h, int(mask)))

    def check_edge_flag_consistency(self, mask):
        """ma::checkFlagConsistency: raises MagError(MAG_ERR_INCONSISTENT) when copies of a shared edge disagree; returns the
        number of disagreeing local copies otherwise (0)."""
        n = C.c_int64(0)
        self._ck(self._L.mag_check_edge_flag_consistency(self._h, int(mask), C.byref(n)))
        return int(n.value)

    def sync_edge_flags(self, mask):
        self._ck(self._L.mag_sync_edge_flags(self._h, int(mask)))

    def allreduce_stats(self):
        s = MagStats()
        self._ck(self._L.mag_allreduce_stats(self._h, C.byref(s)))
        return s.as_dict()
