"""Reader of PUMI's native mesh format (.smb) straight into the flat arrays `Part.set_mesh` takes.

The reference loads an .smb by rebuilding the whole MDS database entity by entity (mds/mds_smb.c:562-600 read_smb,
mds_create_entity per entity) and the adapter would then walk it again to export.  This reader skips the database: it
parses the file with numpy and derives element -> vertex connectivity with the same rules MDS uses, so the arrays come
out in the reference's own order (entity index order = `m->begin(d)` order of a freshly loaded mesh, downward vertices as
`getDownward(e, 0, .)` returns them).

File layout (mds/mds_smb.c; all integers unsigned 32-bit, everything big-endian, pcu/pcu_io.c:238-241):
  header   magic, version (<= 6), dim, number of parts                                   (:120-133)
  counts   entities per type in SMB order VERT EDGE TRI QUAD HEX PRIS PYR TET           (:25-35, :577)
  conn     for every type but VERT: the ONE-LEVEL-DOWN adjacency (edge: 2 vertices, triangle: 3 edges, quad: 4 edges,
           tet: 4 triangles, prism: tri + 3 quads + tri, pyramid: quad + 4 tris), indices within the down type (:158-183)
  points   3 doubles per vertex, then (version >= 2) 2 parametric doubles per vertex     (:586-594)
  remotes  part-boundary vertex links (struct mds_links, :95-113)
  class    (model id, model dim) per entity of every type                                (:230-250)
  tags     n headers {type int|double, components, name\\0}; then per entity type, per tag: ids + values (:257-300, :448-473)
  matches, meta (not needed here)

Deriving lower adjacencies (mds/mds.c:634-670 step_down / convert_down with the `convs` tables :62-180): entity i of
dimension d-2 of an element is the entity the (d-1)-dimensional faces `conv[2i]` and `conv[2i+1]` have in common.
"""
import numpy as np

SMB_VERT, SMB_EDGE, SMB_TRI, SMB_QUAD, SMB_HEX, SMB_PRIS, SMB_PYR, SMB_TET = range(8)
_TYPE_NAMES = ("vertex", "edge", "triangle", "quad", "hex", "prism", "pyramid", "tet")
_DOWN_DEGREE = {SMB_EDGE: 2, SMB_TRI: 3, SMB_QUAD: 4, SMB_HEX: 6, SMB_PRIS: 5, SMB_PYR: 5, SMB_TET: 4}

# mds/mds.c:62-132: pairs of (d-1)-faces whose common (d-2)-entity is entity i
_T10 = ((2, 0), (0, 1), (1, 2))                                      # triangle: vertex i from edges
_Q10 = ((3, 0), (0, 1), (1, 2), (2, 3))                              # quad
_TET21 = ((0, 1), (0, 2), (0, 3), (1, 3), (1, 2), (2, 3))            # tet: edge i from triangles
_TET10 = ((2, 0), (0, 1), (1, 2), (3, 4))                            # tet: vertex i from edges
_W21 = ((0, 1), (0, 2), (0, 3), (1, 3), (1, 2), (2, 3), (1, 4), (2, 4), (3, 4))   # prism: edge i from faces
_W10 = ((0, 2), (0, 1), (1, 2), (6, 8), (6, 7), (7, 8))              # prism: vertex i from edges
_P21 = ((0, 1), (0, 2), (0, 3), (0, 4), (1, 4), (1, 2), (2, 3), (3, 4))   # pyramid
_P10 = ((0, 3), (0, 1), (1, 2), (2, 3), (4, 5))


class _Reader:
    def __init__(self, data):
        self.b = data
        self.o = 0

    def u4(self, n):
        a = np.frombuffer(self.b, dtype=">u4", count=n, offset=self.o)
        self.o += 4 * n
        return a.astype(np.int64)

    def f8(self, n):
        a = np.frombuffer(self.b, dtype=">f8", count=n, offset=self.o)
        self.o += 8 * n
        return a.astype(np.float64)

    def string(self):
        end = self.b.index(b"\0", self.o)
        s = self.b[self.o:end].decode()
        self.o = end + 1
        return s


def _common(a, b):
    """Row-wise: the one entry rows of a [n,ka] and b [n,kb] have in common (mds.c:496-508 common_down)."""
    eq = a[:, :, None] == b[:, None, :]
    hit = eq.any(axis=2)
    assert np.all(hit.sum(axis=1) >= 1), "faces without a common lower entity: not a valid element"
    return a[np.arange(len(a)), hit.argmax(axis=1)]


def _step(faces_down, pairs):
    """faces_down[j]: [n, k_j] lower entities of face j of every element -> [n, len(pairs)] derived entities."""
    return np.stack([_common(faces_down[a], faces_down[b]) for a, b in pairs], axis=1)


def read_smb(path):
    """Returns a dict: dim, version, nparts, counts, xyz [nv,3], edge_v [ne,2], tri_v, quad_v, tet_v [nt,4], prism_v [np,6],
    pyr_v [npy,5] (int32, vertex ids = vertex indices of the file), remotes (peers, lists), classification, and
    tags: {name: {"components": c, "dtype": "int"|"double", entity type name: (ids, values [n,c])}}."""
    r = _Reader(open(path, "rb").read())
    magic, version, dim, nparts = r.u4(4)
    assert version <= 6, "unknown .smb version %d" % version
    n = r.u4(8)
    conn = {t: r.u4(_DOWN_DEGREE[t] * int(n[t])).reshape(-1, _DOWN_DEGREE[t]) for t in range(1, 8)}
    xyz = r.f8(3 * int(n[SMB_VERT])).reshape(-1, 3)
    param = r.f8(2 * int(n[SMB_VERT])).reshape(-1, 2) if version >= 2 else np.zeros((int(n[SMB_VERT]), 2))
    npeers = int(r.u4(1)[0])
    remotes = {}
    if npeers:
        peers, cnt = r.u4(npeers), r.u4(npeers)
        remotes = {int(p): r.u4(int(c)) for p, c in zip(peers, cnt)}
    classification = {_TYPE_NAMES[t]: r.u4(2 * int(n[t])).reshape(-1, 2) for t in range(8)}
    ntags = int(r.u4(1)[0])
    heads = []
    for _ in range(ntags):
        ttype, comps = r.u4(2)
        heads.append((int(ttype), int(comps), r.string()))
    tags = {name: {"components": comps, "dtype": "int" if ttype == 0 else "double"} for ttype, comps, name in heads}
    for t in range(8):
        sizes = r.u4(ntags)
        for (ttype, comps, name), cnt in zip(heads, sizes):
            cnt = int(cnt)
            ids = r.u4(cnt)
            vals = (r.u4(cnt * comps).astype(np.int32) if ttype == 0 else r.f8(cnt * comps)).reshape(cnt, comps)
            if cnt:
                tags[name][_TYPE_NAMES[t]] = (ids, vals)

    edge_v = conn[SMB_EDGE]
    tri_e, quad_e = conn[SMB_TRI], conn[SMB_QUAD]
    tri_v = _step([edge_v[tri_e[:, j]] for j in range(3)], _T10) if len(tri_e) else np.zeros((0, 3), np.int64)
    quad_v = _step([edge_v[quad_e[:, j]] for j in range(4)], _Q10) if len(quad_e) else np.zeros((0, 4), np.int64)

    def element(faces, face_edges, p21, p10):
        if not len(faces):
            return np.zeros((0, len(p10)), np.int64)
        el_e = _step(face_edges, p21)                                  # element -> edges
        return _step([edge_v[el_e[:, j]] for j in range(el_e.shape[1])], p10)

    tf = conn[SMB_TET]
    tet_v = element(tf, [tri_e[tf[:, j]] for j in range(4)], _TET21, _TET10)
    pf = conn[SMB_PRIS]        # faces: tri, quad, quad, quad, tri (mds.c W2)
    prism_v = element(pf, [tri_e[pf[:, 0]], quad_e[pf[:, 1]], quad_e[pf[:, 2]], quad_e[pf[:, 3]], tri_e[pf[:, 4]]]
                      if len(pf) else [], _W21, _W10)
    yf = conn[SMB_PYR]         # faces: quad, tri, tri, tri, tri (mds.c P2)
    pyr_v = element(yf, [quad_e[yf[:, 0]]] + [tri_e[yf[:, j]] for j in range(1, 5)] if len(yf) else [], _P21, _P10)
    i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
    return dict(dim=int(dim), version=int(version), nparts=int(nparts), counts={_TYPE_NAMES[t]: int(n[t]) for t in range(8)},
                xyz=np.ascontiguousarray(xyz), param=param, edge_v=i32(edge_v), tri_v=i32(tri_v), quad_v=i32(quad_v),
                tet_v=i32(tet_v), prism_v=i32(prism_v), pyr_v=i32(pyr_v), remotes=remotes,
                classification=classification, tags=tags)


def vertex_field(mesh, name):
    """Values of a vertex tag / field as a dense [nv, components] array (every vertex must carry it).  apf stores the
    vertex nodes of field <name> in the tag <name>_ver (apf/apfTagData.cc)."""
    tag = mesh["tags"].get(name) or mesh["tags"][name + "_ver"]
    ids, vals = tag["vertex"]
    out = np.empty((len(mesh["xyz"]), vals.shape[1]), dtype=vals.dtype)
    out[ids] = vals
    assert len(ids) == len(out), "tag %s is not set on every vertex" % name
    return out


def read_smb_native(path):
    """The same file through the C reader of the library (mag_smb_read, core_b200/csrc/mag_smb.cu: what the C ABI and the C++
    adapter use; no device needed).  Returns (dict with the same array keys as read_smb, field getter name -> [nv, c])."""
    import ctypes as C
    from ._lib import lib, MagSmbArrays
    L = lib()
    h = C.c_void_p()
    rc = L.mag_smb_read(str(path).encode(), C.byref(h))
    try:
        if rc:
            raise ValueError("mag_smb_read: " + L.mag_smb_last_error(h).decode())
        a = MagSmbArrays()
        assert L.mag_smb_get(h, C.byref(a)) == 0

        def arr(ptr, n, k, ctype, dtype):
            if n == 0:
                return np.zeros((0, k), dtype)
            return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ctype)), shape=(n * k,)).reshape(n, k).astype(dtype).copy()
        out = dict(dim=a.dim, version=a.version, nparts=a.nparts,
                   counts=dict(vertex=a.nv, edge=a.ne, triangle=a.ntri, quad=a.nquad, hex=a.nhex, prism=a.np, pyramid=a.npy, tet=a.nt),
                   xyz=arr(a.xyz, a.nv, 3, C.c_double, np.float64), edge_v=arr(a.edge_v, a.ne, 2, C.c_int32, np.int32),
                   tri_v=arr(a.tri_v, a.ntri, 3, C.c_int32, np.int32), tet_v=arr(a.tet_v, a.nt, 4, C.c_int32, np.int32),
                   prism_v=arr(a.prism_v, a.np, 6, C.c_int32, np.int32), pyr_v=arr(a.pyr_v, a.npy, 5, C.c_int32, np.int32))
        fields = {}

        def field(name):
            if name not in fields:
                comps, ptr = C.c_int(0), C.c_void_p()
                if L.mag_smb_vertex_field(h, name.encode(), C.byref(comps), C.byref(ptr)):
                    raise KeyError(L.mag_smb_last_error(h).decode())
                fields[name] = arr(ptr, a.nv, comps.value, C.c_double, np.float64)
            return fields[name]
        out["fields"] = {n: field(n) for n in ("sizes", "frames") if _has_field(L, h, n)}
        return out
    finally:
        L.mag_smb_free(h)


def _has_field(L, h, name):
    import ctypes as C
    comps, ptr = C.c_int(0), C.c_void_p()
    return L.mag_smb_vertex_field(h, name.encode(), C.byref(comps), C.byref(ptr)) == 0
