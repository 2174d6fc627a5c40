"""Closed-form Kuhn box meshes in the entity order of apf::makeMdsBox.

`kuhn_box(nx, ny, nz, ...)` produces, directly as flat arrays, the same vertices, edges
and tets -- in the same order and with the same per-entity vertex order -- that
`apf::makeMdsBox(nx, ny, nz, wx, wy, wz, true)` followed by iteration over
`m->begin(d)` / `getDownward(e, 0, ..)` would export (reference: mds/apfBox.cc:179-263;
edge orientation of the diagonals follows apf::buildElement's tri_edge_verts /
tet_tri_verts walk, apf/apfMesh.cc:28-33,80-91).  tests/test_boxmesh.py checks the
equivalence against the compiled reference.  It exists because MDS itself cannot hold the
50 M-tet benchmark part in reasonable time/memory (hard part 4).

`slab_part(...)` cuts a global box into x-slabs, one PUMI-style part per GPU: every part
keeps full copies of its boundary entities, and the per-peer shared-edge lists are ordered
identically on both sides (the contract of struct mds_links, mds/mds_net.h:33-38).
"""
import numpy as np

# tets of one cell, vertices numbered as in BoxBuilder::buildCellRegion (apfBox.cc:250-256)
_TET_VERTS = np.array([[0, 1, 2, 6], [0, 2, 3, 6], [0, 3, 7, 6],
                       [0, 7, 4, 6], [0, 4, 5, 6], [0, 5, 1, 6]], dtype=np.int64)
# cell-corner offsets (dx, dy, dz) of rv[0..7]
_CORNER = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0],
                    [0, 0, 1], [1, 0, 1], [1, 1, 1], [0, 1, 1]], dtype=np.int64)


def box_counts(nx, ny, nz):
    nv = (nx + 1) * (ny + 1) * (nz + 1)
    ne = (nx * (ny + 1) * (nz + 1) + (nx + 1) * ny * (nz + 1) + (nx + 1) * (ny + 1) * nz
          + nx * ny * (nz + 1) + (nx + 1) * ny * nz + nx * (ny + 1) * nz + nx * ny * nz)
    nt = 6 * nx * ny * nz
    return nv, ne, nt


def kuhn_box(nx, ny, nz, wx=1.0, wy=1.0, wz=1.0, x0=0, gnx=None, index_dtype=np.int32):
    """Returns (xyz [nv,3] f64, edge_v [ne,2], tet_v [nt,4]).

    x0 / gnx: this box is the x-slab [x0, x0+nx] of a global box with gnx cells along x and
    total width wx, so that shared vertices of neighbouring slabs get bit-identical
    coordinates (w * global_index, exactly as apfBox.cc:181-184 computes them)."""
    if gnx is None:
        gnx = nx
    sx, sy, sz = nx + 1, ny + 1, nz + 1
    nv = sx * sy * sz
    hx, hy, hz = wx / gnx, wy / ny, wz / nz
    ix = np.arange(sx, dtype=np.int64)
    iy = np.arange(sy, dtype=np.int64)
    iz = np.arange(sz, dtype=np.int64)
    xyz = np.empty((sz, sy, sx, 3), dtype=np.float64)
    xyz[..., 0] = (hx * (ix + x0))[None, None, :]
    xyz[..., 1] = (hy * iy)[None, :, None]
    xyz[..., 2] = (hz * iz)[:, None, None]
    xyz = xyz.reshape(nv, 3)

    # vertex grid index and coordinates, grid order (x fastest)
    vid = np.arange(nv, dtype=np.int64)
    vx = vid % sx
    vy = (vid // sx) % sy
    vz = vid // (sx * sy)
    stride = np.array([1, sx, sx * sy], dtype=np.int64)
    notmax = [vx < nx, vy < ny, vz < nz]

    # --- dimension-1 pass: per vertex, edges to +x, +y, +z that exist (apfBox.cc:192-207)
    cand = np.stack([vid, vid, vid], axis=1)                       # [nv,3] first vertex
    other = cand + stride[None, :]
    mask = np.stack(notmax, axis=1)
    axis_edges = np.stack([cand[mask], other[mask]], axis=1)       # row-major flatten keeps (vertex, j) order

    # --- dimension-2 pass: per vertex, faces (jx, jy=(jx+1)%3); the new edge is the diagonal
    # (fv2, fv0) = (v + e_jx + e_jy, v) created by triangle (fv0, fv1, fv2) (apfBox.cc:209-246)
    diag_other = np.stack([vid + stride[j] + stride[(j + 1) % 3] for j in range(3)], axis=1)
    fmask = np.stack([notmax[j] & notmax[(j + 1) % 3] for j in range(3)], axis=1)
    face_edges = np.stack([diag_other[fmask], cand[fmask]], axis=1)

    # --- dimension-3 pass: per cell the body diagonal (rv6, rv0) from tet (0,1,2,6), then 6 tets
    cmask = notmax[0] & notmax[1] & notmax[2]
    c0 = vid[cmask]
    corner = c0[:, None] + (_CORNER @ stride)[None, :]              # [ncell, 8]
    body_edges = np.stack([corner[:, 6], corner[:, 0]], axis=1)
    tet_v = corner[:, _TET_VERTS.reshape(-1)].reshape(-1, 4)

    edge_v = np.concatenate([axis_edges, face_edges, body_edges], axis=0)
    return xyz, np.ascontiguousarray(edge_v.astype(index_dtype)), np.ascontiguousarray(tet_v.astype(index_dtype))


def tri_box(nx, ny, wx=1.0, wy=1.0, index_dtype=np.int32):
    """The 2-D box apf::makeMdsBox(nx, ny, 0, wx, wy, 0, true) builds: (xyz [nv,3] with z = 0, edge_v, tri_v [2 nx ny, 3]).
    Per cell the triangles (v, v+x, v+x+y) and (v+x+y, v+y, v) in that vertex order (BoxBuilder::buildTriangles,
    apfBox.cc:209-217); edges: axis edges in vertex order, then the cell diagonals (v+x+y, v)."""
    sx, sy = nx + 1, ny + 1
    nv = sx * sy
    xyz = np.zeros((sy, sx, 3), dtype=np.float64)
    xyz[..., 0] = ((wx / nx) * np.arange(sx))[None, :]
    xyz[..., 1] = ((wy / ny) * np.arange(sy))[:, None]
    xyz = xyz.reshape(nv, 3)
    vid = np.arange(nv, dtype=np.int64)
    vx, vy = vid % sx, vid // sx
    notmax = [vx < nx, vy < ny]
    stride = np.array([1, sx], dtype=np.int64)
    cand = np.stack([vid, vid], axis=1)
    mask = np.stack(notmax, axis=1)
    axis_edges = np.stack([cand[mask], (cand + stride[None, :])[mask]], axis=1)
    cell = notmax[0] & notmax[1]
    c0 = vid[cell]
    diag = np.stack([c0 + 1 + sx, c0], axis=1)
    tri = np.stack([np.stack([c0, c0 + 1, c0 + 1 + sx], axis=1), np.stack([c0 + 1 + sx, c0 + sx, c0], axis=1)], axis=1).reshape(-1, 3)
    edge_v = np.concatenate([axis_edges, diag], axis=0)
    return xyz, np.ascontiguousarray(edge_v.astype(index_dtype)), np.ascontiguousarray(tri.astype(index_dtype))


def slab_bounds(gnx, nparts):
    """x-cell ranges [lo, hi) of each slab; remainders go to the first parts."""
    base, rem = divmod(gnx, nparts)
    lo, out = 0, []
    for p in range(nparts):
        n = base + (1 if p < rem else 0)
        out.append((lo, lo + n))
        lo += n
    return out


def slab_part(gnx, ny, nz, nparts, part, wx=1.0, wy=1.0, wz=1.0):
    """One x-slab part of the global gnx x ny x nz Kuhn box.

    Returns dict(xyz, edge_v, tet_v, edge_owned [ne] u8, elem_owned [nt] u8,
                 links = [(peer, idx int32[], peer_owns u8[])...], x0, nx).
    Shared entities are the vertices/edges lying in the cut planes x = lo and x = hi.  Owner
    rule (apfPM.cc:109-126): the resident part with the fewest elements, ties -> lowest part
    id.  Shared-edge lists are sorted by the LOWER part's local edge index on both sides."""
    bounds = slab_bounds(gnx, nparts)
    lo, hi = bounds[part]
    nx = hi - lo
    xyz, edge_v, tet_v = kuhn_box(nx, ny, nz, wx, wy, wz, x0=lo, gnx=gnx)
    sx = nx + 1
    ne, nt = len(edge_v), len(tet_v)
    nelem = [6 * (b[1] - b[0]) * ny * nz for b in bounds]

    def plane_edges(nx_local, plane_ix):
        """local indices (box order) of edges with both ends in the plane x = plane_ix, plus a
        part-independent key (the (y,z) ids of both end vertices)."""
        _, ev, _ = kuhn_box(nx_local, ny, nz) if nx_local != nx else (None, edge_v, None)
        s = nx_local + 1
        a, b = ev[:, 0].astype(np.int64), ev[:, 1].astype(np.int64)
        inplane = ((a % s) == plane_ix) & ((b % s) == plane_ix)
        idx = np.nonzero(inplane)[0]
        key = np.stack([a[idx] // s, b[idx] // s], axis=1)         # (y + sy*z) of each end
        return idx, key

    edge_owned = np.ones(ne, dtype=np.uint8)
    links = []
    for peer, my_plane in ((part - 1, 0), (part + 1, nx)):
        if peer < 0 or peer >= nparts:
            continue
        my_idx, my_key = plane_edges(nx, my_plane)
        pnx = bounds[peer][1] - bounds[peer][0]
        peer_plane = pnx if peer < part else 0
        p_idx, p_key = plane_edges(pnx, peer_plane)
        # match by key; order by the lower part's local index
        def order(k):
            return np.lexsort((k[:, 1], k[:, 0]))
        mo, po = order(my_key), order(p_key)
        assert np.array_equal(my_key[mo], p_key[po])
        mine, theirs = my_idx[mo], p_idx[po]
        lower_local = mine if part < peer else theirs
        perm = np.argsort(lower_local, kind="stable")
        mine = mine[perm]
        owner = part if (nelem[part], part) < (nelem[peer], peer) else peer
        peer_owns = np.full(len(mine), 1 if owner == peer else 0, dtype=np.uint8)
        if owner == peer:
            edge_owned[mine] = 0
        links.append((peer, np.ascontiguousarray(mine.astype(np.int32)), peer_owns))
    return dict(xyz=xyz, edge_v=edge_v, tet_v=tet_v, edge_owned=edge_owned,
                elem_owned=np.ones(nt, dtype=np.uint8), links=links, x0=lo, nx=nx, sx=sx)


# ---------------------------------------------------------------------------- mixed prism / tet boxes (config 5)
def mixed_box(n, k, index_dtype=np.int32, ny=None, nz=None, x0=0, gnx=None, wx=1.0):
    """n x ny x nz cells (default a cube): the bottom k cell layers are 2 prisms per cell, (0,1,2|4,5,6) and (0,2,3|4,6,7)
    -- conforming to the 0-2 bottom diagonal of the Kuhn tets (apfBox.cc:250-256) -- and Kuhn tets above.  Returns
    (xyz, edge_v, tet_v, prism_v); edges are the unique vertex pairs of all elements sorted by (min, max) vertex
    (a mixed mesh cannot come from makeMdsBox, so there is no reference creation order to follow; the parity
    fixtures of tests/golden use the order apf::buildElement produced).
    x0 / gnx / wx: this box is the x-slab [x0, x0 + n] of a global box with gnx cells of total width wx along x (shared
    vertices of neighbouring slabs get bit-identical coordinates, as in kuhn_box)."""
    nx = n
    ny = n if ny is None else ny
    nz = n if nz is None else nz
    gnx = nx if gnx is None else gnx
    sx, sy, sz = nx + 1, ny + 1, nz + 1
    xyz = np.empty((sz, sy, sx, 3), dtype=np.float64)
    xyz[..., 0] = ((wx / gnx) * (np.arange(sx) + x0))[None, None, :]
    xyz[..., 1] = (np.arange(sy) / ny)[None, :, None]
    xyz[..., 2] = (np.arange(sz) / nz)[:, None, None]
    xyz = xyz.reshape(-1, 3)
    cz, cy, cx = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    c0 = (cx + sx * (cy + sy * cz)).reshape(-1)
    zc = cz.reshape(-1)
    stride = np.array([1, sx, sx * sy], dtype=np.int64)
    corner = c0[:, None] + (_CORNER @ stride)[None, :]
    lay = zc < k
    pc = corner[lay]
    prism_v = np.stack([pc[:, [0, 1, 2, 4, 5, 6]], pc[:, [0, 2, 3, 4, 6, 7]]], axis=1).reshape(-1, 6)
    tet_v = corner[~lay][:, _TET_VERTS.reshape(-1)].reshape(-1, 4)
    pairs = [tet_v[:, [a, b]] for a, b in ((0, 1), (1, 2), (2, 0), (0, 3), (1, 3), (2, 3))]
    pairs += [prism_v[:, [a, b]] for a, b in _PRISM_EDGES]
    e = np.sort(np.concatenate(pairs, axis=0), axis=1)
    nv = sx * sy * sz
    key = np.unique(e[:, 0].astype(np.int64) * nv + e[:, 1])     # sorted by (min, max)
    edge_v = np.stack([key // nv, key % nv], axis=1)
    return (xyz, np.ascontiguousarray(edge_v.astype(index_dtype)), np.ascontiguousarray(tet_v.astype(index_dtype)),
            np.ascontiguousarray(prism_v.astype(index_dtype)))


def mixed_slab_part(gnx, ny, nz, k, nparts, part, wx=1.0):
    """One x-slab part of the global gnx x ny x nz mixed prism / tet box (BASELINE configs[4] on several GPUs).  Same contract
    as slab_part: full copies of the entities in the cut planes, per-peer shared-edge lists in the same order on both sides
    (sorted by the (y, z) grid ids of the two end vertices, which both parts see identically), owner = the part with the
    fewest elements, ties -> lowest part id (apfPM.cc:109-126).
    Returns dict(xyz, edge_v, tet_v, prism_v, edge_owned, links, x0, nx)."""
    bounds = slab_bounds(gnx, nparts)
    lo, hi = bounds[part]
    nx = hi - lo
    xyz, edge_v, tet_v, prism_v = mixed_box(nx, k, ny=ny, nz=nz, x0=lo, gnx=gnx, wx=wx)
    sx = nx + 1
    nelem = [(b[1] - b[0]) * ny * (2 * min(k, nz) + 6 * max(nz - k, 0)) for b in bounds]
    a, b = edge_v[:, 0].astype(np.int64), edge_v[:, 1].astype(np.int64)
    edge_owned = np.ones(len(edge_v), dtype=np.uint8)
    links = []
    for peer, plane in ((part - 1, 0), (part + 1, nx)):
        if peer < 0 or peer >= nparts:
            continue
        idx = np.nonzero(((a % sx) == plane) & ((b % sx) == plane))[0]
        ka, kb = a[idx] // sx, b[idx] // sx                    # y + (ny+1) z of each end: the same numbers on the peer
        kmin, kmax = np.minimum(ka, kb), np.maximum(ka, kb)
        idx = idx[np.lexsort((kmax, kmin))]
        owner = part if (nelem[part], part) < (nelem[peer], peer) else peer
        peer_owns = np.full(len(idx), 1 if owner == peer else 0, dtype=np.uint8)
        if owner == peer:
            edge_owned[idx] = 0
        links.append((peer, np.ascontiguousarray(idx.astype(np.int32)), peer_owns))
    return dict(xyz=xyz, edge_v=edge_v, tet_v=tet_v, prism_v=prism_v, edge_owned=edge_owned, links=links, x0=lo, nx=nx)


# apf prism_edge_verts / pyramid_edge_verts (apf/apfMesh.cc:62-78)
_PRISM_EDGES = ((0, 1), (1, 2), (2, 0), (0, 3), (1, 4), (2, 5), (3, 4), (4, 5), (5, 3))
_PYRAMID_EDGES = ((0, 1), (1, 2), (2, 3), (3, 0), (0, 4), (1, 4), (2, 4), (3, 4))
_LAYER, _DONT_SPLIT, _DONT_COLLAPSE, _DONT_SWAP, _OK_QUALITY = 1 << 10, 1 << 1, 1 << 3, 1 << 9, 1 << 6


def layer_closure_flags(edge_v, prism_v, pyr_v, nt):
    """The "ma_flags" words ma::Adapt's constructor leaves on a mesh with layer elements, before any mark
    (markLayerElements + freezeLayer, ma/maLayer.cc:11-71): every edge in the closure of a prism / pyramid gets
    LAYER | DONT_COLLAPSE | DONT_SPLIT | DONT_SWAP, every prism / pyramid LAYER | OK_QUALITY, so all three marks skip
    them.  This is one-time set-up done by the caller (in the reference: the Adapt constructor), not part of a
    sweep; the result is what mag_set_flags receives.  Returns (edge_flags [ne], elem_flags [np+npy+nt])."""
    edge_v = np.asarray(edge_v, dtype=np.int64)
    ne = len(edge_v)
    np_ = 0 if prism_v is None else len(prism_v)
    npy = 0 if pyr_v is None else len(pyr_v)
    ef = np.zeros(ne, dtype=np.int32)
    lf = np.zeros(np_ + npy + nt, dtype=np.int32)
    lf[:np_ + npy] = _LAYER | _OK_QUALITY
    pairs = []
    if np_:
        pairs += [np.asarray(prism_v, dtype=np.int64)[:, [a, b]] for a, b in _PRISM_EDGES]
    if npy:
        pairs += [np.asarray(pyr_v, dtype=np.int64)[:, [a, b]] for a, b in _PYRAMID_EDGES]
    if pairs:
        nvmax = int(edge_v.max()) + 1
        lay = np.sort(np.concatenate(pairs, axis=0), axis=1)
        lay_key = np.unique(lay[:, 0] * nvmax + lay[:, 1])
        es = np.sort(edge_v, axis=1)
        hit = np.isin(es[:, 0] * nvmax + es[:, 1], lay_key)
        ef[hit] = _LAYER | _DONT_COLLAPSE | _DONT_SPLIT | _DONT_SWAP
    return ef, lf
