"""The C++ drop-in adapter (core_b200/adapter) against the unmodified reference, in one process: libmag_ma.so holds the
adapter, the compiled reference it plugs into, and the self-check of tests/adapter/adapter_check.cc.  The library is
built where /root/reference exists (the build container) and travels to the GPU box prebuilt."""
import ctypes as C
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "core_b200", "lib", "libmag_ma.so")


def test_adapter_sources_cite_reference():
    src = open(os.path.join(ROOT, "core_b200", "adapter", "magAdapt.h")).read()
    for cite in ("ma/maSize.h", "ma/maRefine.cc", "ma/maCoarsen.cc", "ma/maShape.cc", "ma/maStats.cc"):
        assert cite in src


@pytest.mark.gpu
@pytest.mark.parametrize("log_interp,fp_mode", [(0, 0), (0, 1), (1, 0), (1, 1)])
def test_adapter_reproduces_reference(built, log_interp, fp_mode):
    if not os.path.exists(LIB):
        pytest.skip("libmag_ma.so not built (needs the reference headers)")
    L = C.CDLL(LIB)
    L.mag_adapter_check.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double, C.c_void_p]
    rep = np.zeros(20)
    rc = L.mag_adapter_check(16, log_interp, fp_mode, 0.25, rep.ctypes.data_as(C.c_void_p))
    ref, bulk, loops = rep[0:5], rep[5:10], rep[10:15]
    assert rc == 0, "adapter differs from the reference: ref=%s bulk=%s loops=%s flagdiffs=%s" % (ref, bulk, loops, rep[15:19])
    assert ref[0] > 0 and ref[2] > 0                       # the case marks something
    assert np.array_equal(ref[:3], bulk[:3]) and np.array_equal(ref[:3], loops[:3])
    assert np.all(rep[15:19] == 0)
    # the unmodified reference loops (5 whole-mesh sweeps: split, collapse, bad, minq, maxlen) were served by a handful
    # of device sweeps, not by per-entity evaluation
    # the reference's own loops were served from device sweeps, not per entity: either the snapshot the last bulk sweep left
    # still matched the mesh (vertex hash: 0 launches) or a handful of launches re-swept it; per-entity evaluation would be ~1e5
    assert 0 <= rep[19] <= 5 * 40


@pytest.mark.gpu
@pytest.mark.parametrize("log_interp,fp_mode", [(0, 0), (0, 1), (1, 0)])
def test_adapter_reproduces_reference_2d(built, log_interp, fp_mode):
    """The same on a 2-D box: triangles are the elements (ma::measureTriQuality, goodQuality 0.2 as ma::Input picks in 2-D)."""
    if not os.path.exists(LIB):
        pytest.skip("libmag_ma.so not built (needs the reference headers)")
    L = C.CDLL(LIB)
    L.mag_adapter_check_2d.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double, C.c_void_p]
    rep = np.zeros(20)
    rc = L.mag_adapter_check_2d(60, log_interp, fp_mode, 0.25, rep.ctypes.data_as(C.c_void_p))
    ref, bulk, loops = rep[0:5], rep[5:10], rep[10:15]
    assert rc == 0, "adapter differs from the reference: ref=%s bulk=%s loops=%s flagdiffs=%s" % (ref, bulk, loops, rep[15:19])
    assert ref[0] > 0 and ref[2] > 0
    assert np.array_equal(ref[:3], bulk[:3]) and np.array_equal(ref[:3], loops[:3])
    assert np.all(rep[15:19] == 0)


@pytest.mark.gpu
def test_adapter_threaded_export(built):
    """The opt-in multi-threaded MDS walk of the export (setExportThreads) gives the same export: the whole self-check
    passes with 8 host threads on a box large enough for the threads to start (n = 24: 83 k tets)."""
    if not os.path.exists(LIB):
        pytest.skip("libmag_ma.so not built (needs the reference headers)")
    L = C.CDLL(LIB)
    L.mag_adapter_check.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double, C.c_void_p]
    L.mag_adapter_set_threads.argtypes = [C.c_int]
    L.mag_adapter_set_threads(8)
    try:
        rep = np.zeros(20)
        rc = L.mag_adapter_check(24, 0, 0, 0.25, rep.ctypes.data_as(C.c_void_p))
        assert rc == 0 and np.all(rep[15:19] == 0) and np.array_equal(rep[0:3], rep[5:8])
    finally:
        L.mag_adapter_set_threads(1)


@pytest.mark.gpu
@pytest.mark.parametrize("dim,log_interp,fp_mode,jitter", [(3, 0, 0, 0.0), (3, 0, 1, 0.25), (3, 1, 0, 0.25), (3, 1, 1, 0.0),
                                                           (2, 0, 0, 0.0), (2, 1, 1, 0.25), (4, 0, 0, 0.0), (4, 1, 0, 0.0)])
def test_ma_adapt_through_the_adapter_gives_the_same_mesh(built, dim, log_interp, fp_mode, jitter):
    """The drop-in claim end to end: the UNMODIFIED ma::adapt driver (refine / coarsen / shape correction, two iterations) on
    a 10^3 box of tets or a 60^2 box of triangles (lattice or jittered) with the rotating shock-layer fields, once with the reference's own size field
    (AnisoSizeField or LogAnisoSizeField) and once with mag::GpuSizeField + mag::shapeHandler plugged into ma::Input.
    Same adapted mesh: counts, coordinates and connectivity in iteration order, same longest metric edge (the quantity
    test/aniso_adapt.h:65-74 checks); and the device did serve whole-mesh sweeps on the way.  (The log-Euclidean case is
    what exposed that the export has to read the size field's own ma_logM field: the sizes / frames it was built from go
    stale for vertices created by refinement.)
    dim = 4: a unit ball (8 tets around the centre of an octahedron) on an ANALYTIC sphere model with snapping on, three
    iterations -- the situation of test/ma_test_analytic_model.cc: snap() (ma/ma.cc:37) moves the vertices refinement created on
    the boundary with no size-field callback, so the adapter has to notice that its device copy of the coordinates is stale
    (vertex hash, magAdapt.cc ensureExported) to keep producing the reference's mesh."""
    if not os.path.exists(LIB):
        pytest.skip("libmag_ma.so not built (needs the reference headers)")
    L = C.CDLL(LIB)
    L.mag_adapter_adapt_check2.argtypes = [C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int, C.c_void_p]
    L.mag_adapter_set_adapt_jitter.argtypes = [C.c_double]
    L.mag_adapter_set_adapt_dim.argtypes = [C.c_int]
    L.mag_adapter_set_adapt_jitter(jitter)
    L.mag_adapter_set_adapt_dim(dim)
    try:
        out = np.zeros(11)
        rc = L.mag_adapter_adapt_check2({3: 10, 2: 60, 4: 4}[dim], 3, 1.0, 2 if dim != 4 else 3, log_interp, fp_mode, out.ctypes.data_as(C.c_void_p))
    finally:
        L.mag_adapter_set_adapt_jitter(0.0)
        L.mag_adapter_set_adapt_dim(3)
    assert rc == 0, out.tolist()
    assert np.array_equal(out[0:3], out[3:6]) and out[6] == 0
    assert out[2] > 3 * {3: 6000, 2: 7200, 4: 8}[dim]   # the mesh was really adapted (6000 tets / 7200 triangles / 8 tets before)
    assert out[8] > 0                             # device sweeps happened


def test_edge_links_from_stub_sharing(built):
    """mag::buildEdgeLinks (the adapter's part-boundary lists, built from apf::Sharing with no communication) on a stub
    Sharing that glues two boxes into the two slab parts of one box: the lists are boxmesh.slab_part's -- same edges, same
    order on both sides (sorted by the lower part's entity), same owner bits (apfPM.cc:109-126).  No device needed."""
    if not os.path.exists(LIB):
        pytest.skip("libmag_ma.so not built (needs the reference headers)")
    import core_b200.boxmesh as boxmesh
    L = C.CDLL(LIB)
    L.mag_adapter_links_check.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
    for gnx, ny, nz in ((5, 3, 2), (4, 2, 3)):
        (lo0, hi0), (lo1, hi1) = boxmesh.slab_bounds(gnx, 2)
        for part in (0, 1):
            idx = np.zeros(4096, np.int32)
            own = np.zeros(4096, np.uint8)
            n = L.mag_adapter_links_check(hi0 - lo0, hi1 - lo1, ny, nz, part, idx.ctypes.data_as(C.c_void_p), own.ctypes.data_as(C.c_void_p), 4096)
            assert n > 0
            (peer, want_idx, want_own), = boxmesh.slab_part(gnx, ny, nz, 2, part)["links"]
            assert peer == 1 - part
            assert np.array_equal(idx[:n], want_idx), (gnx, part)
            assert np.array_equal(own[:n], want_own), (gnx, part)


@pytest.mark.parametrize("n,kind,threads,iters", [(8, 0, 1, 0), (7, 0, 3, 0), (9, 2, 1, 0), (6, 3, 1, 0), (6, 1, 1, 2), (8, 1, 4, 3)])
def test_direct_mds_export_equals_public_walk(built, n, kind, threads, iters):
    """The adapter's direct export (MDS's own arrays: struct mds one-level-down tables, point[][3], tag arrays) against its
    walk through apf::Mesh2's public interface, host only: every exported array, the entity lists, the slot tables, the
    change detection (a moved vertex, an edited field value) and the direct read of the ma_flags tag.  kind 0: jittered box of
    tets; 1: the same after `iters` iterations of the reference's own ma::adapt (free-list holes, new entities between old
    ones); 2: box of triangles; 3: prisms below Kuhn tets plus a detached pyramid.  No device needed."""
    if not os.path.exists(LIB):
        pytest.skip("libmag_ma.so not built (needs the reference headers)")
    L = C.CDLL(LIB)
    L.mag_adapter_export_check.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
    t = np.zeros(8)
    mask = L.mag_adapter_export_check(n, kind, threads, iters, t.ctypes.data_as(C.c_void_p))
    assert mask == 0, "routes differ, mask %#x (1: direct route not taken, 2 vertices, 4 edges, 8 connectivity, 16 elements, 32/1024/2048 slot tables, 64-256 change detection, 512 flag tag)" % mask
    assert t[2] > 0 and t[3] > 0 and t[4] > 0
    if kind == 1:
        assert t[4] > 6 * n ** 3        # the mesh was adapted


@pytest.mark.gpu
def test_adapter_public_api_route(built):
    """The export route of round 1 (every entity through apf::Mesh2's public interface) still reproduces the reference: what a
    non-MDS apf::Mesh2 would get."""
    if not os.path.exists(LIB):
        pytest.skip("libmag_ma.so not built (needs the reference headers)")
    L = C.CDLL(LIB)
    L.mag_adapter_check.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double, C.c_void_p]
    L.mag_adapter_set_direct.argtypes = [C.c_int]
    L.mag_adapter_set_direct(0)
    try:
        rep = np.zeros(20)
        rc = L.mag_adapter_check(12, 0, 0, 0.25, rep.ctypes.data_as(C.c_void_p))
        assert rc == 0 and np.all(rep[15:19] == 0) and np.array_equal(rep[0:3], rep[5:8])
    finally:
        L.mag_adapter_set_direct(1)


@pytest.mark.gpu
@pytest.mark.parametrize("fp_mode,refine_layer,coarsen_layer,to_tets", [(0, 0, 0, 0), (0, 1, 1, 0), (0, 1, 1, 1), (1, 1, 1, 0)])
def test_adapter_layer_element_weights(built, fp_mode, refine_layer, coarsen_layer, to_tets):
    """mag::getElementWeights on a mixed mesh (prisms under Kuhn tets) against ma::getElementWeight of the unmodified reference
    for every element: prisms go through mag_prism_weights on their base triangles (face order from MDS), tets through
    mag_element_weights -- for the layer permissions ma::Input offers.  (No pyramid: the reference's own getWeight crashes on
    one, apf has no order-2 pyramid rule.)"""
    if not os.path.exists(LIB):
        pytest.skip("libmag_ma.so not built (needs the reference headers)")
    L = C.CDLL(LIB)
    L.mag_adapter_layer_weights_check.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
    L.mag_adapter_layer_weights_check.restype = C.c_long
    out = np.zeros(3)
    diffs = L.mag_adapter_layer_weights_check(6, fp_mode, refine_layer, coarsen_layer, to_tets, out.ctypes.data_as(C.c_void_p))
    assert out[0] == 2 * 6 * 6 * 2 and out[1] == 0 and out[2] == 6 * 6 * 6 * 4
    assert diffs == 0


@pytest.mark.gpu
@pytest.mark.parametrize("fp_mode", [0, 1])
def test_adapter_collapse_candidates_against_ma_collapse(built, fp_mode):
    """mag::collapseQualities (one mag_collapse_quality call) against ma::Collapse ITSELF: on a jittered 10^3 box whose size
    field asks for coarsening, the unmodified reference really rebuilds the cavity of every collapse candidate that passes its
    classification / topology checks, in each permitted direction (computeElementSets + rebuildElements), measures old and new
    elements through its shape handler and destroys the new ones again; the device answers the same ~6800 candidates from its
    export of the restored mesh (MDS now has free-list holes).  Old and new worst qualities agree bit for bit in strict
    arithmetic (1e-12 fast), including the ~1600 candidates whose rebuilt cavity holds an inverted tet."""
    if not os.path.exists(LIB):
        pytest.skip("libmag_ma.so not built (needs the reference headers)")
    L = C.CDLL(LIB)
    L.mag_adapter_collapse_check.argtypes = [C.c_int, C.c_int, C.c_double, C.c_void_p]
    L.mag_adapter_collapse_check.restype = C.c_long
    out = np.zeros(3)
    diffs = L.mag_adapter_collapse_check(10, fp_mode, 2.5, out.ctypes.data_as(C.c_void_p))
    assert out[0] > 5000 and out[1] > 1000 and out[2] > 2000
    assert diffs == 0
