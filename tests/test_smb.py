"""core_b200/smb.py (the .smb reader) against files the reference itself wrote (mds_write_smb) and the
arrays its own API exports from the same meshes (tests/golden/make_golden.py): entity order, downward vertex order,
coordinates and vertex fields must be identical, so the arrays can go straight into mag_set_mesh / the oracle."""
import os

import numpy as np

from core_b200 import smb
from oracle import mao
import util


def test_kuhn_box_with_fields(built):
    s = smb.read_smb(os.path.join(util.GOLDEN, "kbox322_0.smb"))
    g = util.load("smb_kbox322")
    assert s["dim"] == 3 and s["nparts"] == 1 and s["version"] == 6
    assert s["counts"]["tet"] == 72 and s["counts"]["vertex"] == 36
    assert np.array_equal(s["xyz"], g["xyz"])
    assert np.array_equal(s["edge_v"], g["edge_v"])
    assert np.array_equal(s["tet_v"], g["elem_v"][:, :4])
    h, R = smb.vertex_field(s, "sizes"), smb.vertex_field(s, "frames")
    assert np.array_equal(h, g["h"]) and np.array_equal(R, g["R"])
    # every triangle of the file is the face of a tet, with the vertices MDS would report
    tets = {tuple(sorted(t)) for t in s["tet_v"].tolist()}
    assert all(any(set(tri) <= set(t) for t in tets) for tri in s["tri_v"][:20].tolist())
    # straight from the file into the sweep arithmetic: the reference's own lengths and qualities come out
    assert np.array_equal(mao.edge_lengths(mao.ANISO, s["xyz"], h, R, s["edge_v"]), g["lengths"])
    assert np.array_equal(mao.tet_qualities(mao.ANISO, s["xyz"], h, R, s["tet_v"]), g["qualities"])


def test_mixed_prism_tet_mesh(built):
    s = smb.read_smb(os.path.join(util.GOLDEN, "mixed3_0.smb"))
    g = util.load("smb_mixed3")
    prism_v, pyr_v, tet_v = util.split_elements(g)
    assert s["counts"]["prism"] == len(prism_v) == 18 and s["counts"]["quad"] == 33
    assert np.array_equal(s["xyz"], g["xyz"]) and np.array_equal(s["edge_v"], g["edge_v"])
    assert np.array_equal(s["prism_v"], prism_v) and np.array_equal(s["tet_v"], tet_v)
    assert len(s["pyr_v"]) == 0 and s["quad_v"].shape == (33, 4)


def test_native_reader_equals_numpy_reader(built, tmp_path):
    """mag_smb_read (the C reader behind the C ABI, core_b200/csrc/mag_smb.cu) returns the arrays of the numpy reader -- and
    thereby of the reference's own export -- on the files the reference wrote; errors come back as codes with a text."""
    import ctypes as C
    from core_b200._lib import lib
    for name in ("kbox322_0.smb", "mixed3_0.smb"):
        path = os.path.join(util.GOLDEN, name)
        a, b = smb.read_smb(path), smb.read_smb_native(path)
        assert (a["dim"], a["version"], a["nparts"]) == (b["dim"], b["version"], b["nparts"])
        assert a["counts"] == b["counts"]
        for k in ("xyz", "edge_v", "tri_v", "tet_v", "prism_v", "pyr_v"):
            assert np.array_equal(a[k], b[k]), (name, k)
    b = smb.read_smb_native(os.path.join(util.GOLDEN, "kbox322_0.smb"))
    a = smb.read_smb(os.path.join(util.GOLDEN, "kbox322_0.smb"))
    assert np.array_equal(b["fields"]["sizes"], smb.vertex_field(a, "sizes"))
    assert np.array_equal(b["fields"]["frames"], smb.vertex_field(a, "frames"))
    # a truncated file and a missing file are refused with a message
    data = open(os.path.join(util.GOLDEN, "mixed3_0.smb"), "rb").read()
    bad = tmp_path / "cut.smb"
    bad.write_bytes(data[:len(data) // 2])
    L = lib()
    for p in (str(bad), str(tmp_path / "none.smb")):
        h = C.c_void_p()
        assert L.mag_smb_read(p.encode(), C.byref(h)) != 0
        assert len(L.mag_smb_last_error(h)) > 0
        L.mag_smb_free(h)
