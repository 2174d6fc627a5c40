"""The closed-form Kuhn box generator against apf::makeMdsBox (golden export of the compiled reference, and the
live reference when it is built here), and the slab partition's link / ownership rules."""
import numpy as np
import pytest

import core_b200.boxmesh as boxmesh
import util


def test_kuhn_box_matches_makeMdsBox_golden():
    g = util.load("box6_identity")
    xyz, ev, tv = boxmesh.kuhn_box(6, 6, 6)
    assert np.array_equal(xyz, g["xyz"])
    assert np.array_equal(ev, g["edge_v"])
    assert np.array_equal(tv, g["elem_v"][:, :4])


@pytest.mark.parametrize("dims", [(1, 1, 1), (2, 3, 4), (5, 2, 3)])
def test_kuhn_box_matches_live_reference(dims):
    from oracle import refo
    if not refo.available():
        pytest.skip("compiled reference not built here")
    m = refo.RefMesh.box(*dims, 1.0, 2.0, 0.5)
    rx, rev, _, relv = m.export()
    m.close()
    xyz, ev, tv = boxmesh.kuhn_box(*dims, 1.0, 2.0, 0.5)
    assert np.array_equal(xyz, rx) and np.array_equal(ev, rev) and np.array_equal(tv, relv[:, :4])


def test_tri_box_matches_makeMdsBox_golden():
    g = util.load("tri9x7_iso")
    xyz, ev, tv = boxmesh.tri_box(9, 7)
    assert np.array_equal(ev, g["edge_v"]) and np.array_equal(tv, g["elem_v"][:, :3])
    assert len(xyz) == len(g["xyz"])


def test_box_counts():
    for n in (1, 3, 20):
        nv, ne, nt = boxmesh.box_counts(n, n, n)
        assert nv == (n + 1) ** 3 and nt == 6 * n ** 3
        assert ne == 3 * n * (n + 1) ** 2 + 3 * n * n * (n + 1) + n ** 3
    xyz, ev, tv = boxmesh.kuhn_box(3, 4, 5)
    assert (len(xyz), len(ev), len(tv)) == boxmesh.box_counts(3, 4, 5)
    # every tet edge is in the edge list exactly once (Euler: unique pairs == ne)
    pairs = np.concatenate([tv[:, [a, b]] for a, b in ((0, 1), (1, 2), (2, 0), (0, 3), (1, 3), (2, 3))])
    assert len(np.unique(np.sort(pairs, axis=1), axis=0)) == len(ev)
    assert len(np.unique(np.sort(ev, axis=1), axis=0)) == len(ev)


def test_slab_parts_cover_global_box():
    """Slab parts glue back into the global box: same geometry, shared edges listed identically on both sides,
    exactly one owner per shared edge (apfPM.cc:109-126: fewest elements, ties -> lowest part id)."""
    gnx, ny, nz, P = 7, 3, 2, 3
    parts = [boxmesh.slab_part(gnx, ny, nz, P, r) for r in range(P)]
    gx, gev, gtv = boxmesh.kuhn_box(gnx, ny, nz)
    assert sum(len(p["tet_v"]) for p in parts) == len(gtv)
    # global edge set == union of part edge sets (by end-point coordinates)
    def keyset(xyz, ev):
        a, b = xyz[ev[:, 0]], xyz[ev[:, 1]]
        k = np.concatenate([np.minimum(a, b), np.maximum(a, b)], axis=1)
        return set(map(tuple, np.round(k * 840).astype(np.int64)))
    allk = set()
    for p in parts:
        allk |= keyset(p["xyz"], p["edge_v"])
    assert allk == keyset(gx, gev)
    # owned edges partition the global edge set
    assert sum(int(p["edge_owned"].sum()) for p in parts) == len(gev)
    # links: same length both sides, same geometric edges in the same order, complementary ownership
    for r, p in enumerate(parts):
        for peer, idx, peer_owns in p["links"]:
            q = parts[peer]
            back = [l for l in q["links"] if l[0] == r]
            assert len(back) == 1
            _, qidx, qowns = back[0]
            assert len(idx) == len(qidx) > 0
            mine = p["xyz"][p["edge_v"][idx]].reshape(len(idx), 6)
            theirs = q["xyz"][q["edge_v"][qidx]].reshape(len(idx), 6)
            assert np.array_equal(mine, theirs), "shared edges not listed in the same order on both sides"
            assert np.all(peer_owns + qowns == 1)
            assert np.array_equal(p["edge_owned"][idx] == 0, peer_owns == 1)
    # owner rule: parts 0 (3 cells) vs 1 (2 cells): fewest elements wins
    assert boxmesh.slab_bounds(7, 3) == [(0, 3), (3, 5), (5, 7)]
    l01 = [l for l in parts[0]["links"] if l[0] == 1][0]
    assert np.all(l01[2] == 1)          # part 1 has fewer elements -> owns the 0|1 interface
    l12 = [l for l in parts[1]["links"] if l[0] == 2][0]
    assert np.all(l12[2] == 0)          # tie 2 vs 2 cells -> lower id (part 1) owns


def test_mixed_box_conforms():
    xyz, ev, tv, pv = boxmesh.mixed_box(4, 1)
    assert len(pv) == 2 * 16 and len(tv) == 6 * 48
    ef, lf = boxmesh.layer_closure_flags(ev, pv, None, len(tv))
    assert np.all(lf[:len(pv)] == (1 << 10 | 1 << 6)) and np.all(lf[len(pv):] == 0)
    # layer edges are exactly those with both ends at z <= 1 cell
    z = xyz[:, 2]
    inlayer = (z[ev[:, 0]] <= 0.25 + 1e-12) & (z[ev[:, 1]] <= 0.25 + 1e-12)
    # top-face diagonals 1-3 of a cell are not prism edges; all prism top-face edges are
    assert np.all(ef[~inlayer] == 0)
    assert ef[inlayer].astype(bool).sum() == len(np.unique(np.sort(
        np.concatenate([pv[:, [a, b]] for a, b in boxmesh._PRISM_EDGES]), axis=1), axis=0))
