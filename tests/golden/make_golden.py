"""Generates tests/golden/*.npz from the COMPILED, UNMODIFIED reference (oracle/_ref/libref_oracle.so,
built from /root/reference by oracle/ref/Makefile).  Run in the build container only:

    python tests/golden/make_golden.py

Every array below except the inputs (coordinates jitter, size fields) is an OUTPUT OF THE REFERENCE:
apf::makeMdsBox / apf::buildElement entity order, ma::SizeField::measure, ma::measureElementQuality,
ma::markEdgesToSplit / markEdgesToCollapse / markBadQuality / getMinQuality / getMaximumEdgeLength,
ma::isPrismOk / isPyramidOk, the ma_logM field, apf::eigen, ma::getElementWeight, ma::makeSplitVert's
size-field transfer, ma::getSliverCode / matchSliver.  The fixtures are small (a few 100 kB) and
committed, because /root/reference does not exist on the GPU box.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import refo  # noqa: E402
import core_b200.fields as fields  # noqa: E402


def random_frames(nv, rng, skew=1e-3):
    A = rng.standard_normal((nv, 3, 3))
    Q, _ = np.linalg.qr(A)
    Q[:, :, 2] *= np.sign(np.linalg.det(Q))[:, None]
    return (Q + skew * rng.standard_normal((nv, 3, 3))).reshape(nv, 9)


def case(name, m, kind, h=None, R=None, good_quality=-1.0, edge_flags=None, elem_flags=None, use_max=True):
    xyz, ev, et, elv = m.export()
    m.set_sizefield(kind, h, R)
    out = dict(xyz=xyz, edge_v=ev, elem_type=et, elem_v=elv, kind=np.int64(kind), good_quality=np.float64(good_quality))
    if h is not None:
        out["h"] = np.asarray(h, dtype=np.float64)
    if R is not None:
        out["R"] = np.asarray(R, dtype=np.float64)
    if kind in (refo.KIND_LOG_FIELD, refo.KIND_LOG_FN):
        out["logM"] = m.logm()
    out["lengths"] = m.lengths()
    simplex_only = not np.any((et == refo.PRISM) | (et == refo.PYRAMID))
    if simplex_only:
        out["qualities"] = m.qualities(use_max)
        out["qualities_centroid"] = m.qualities(False)
        out["vertex_Q"] = m.vertex_transforms()
    out["stats_el"], out["stats_lq"] = m.stats()   # ma::stats (maStats.cc:115-134): the tables of measureAnisoStats
    which = 15 if simplex_only else 3 + 4  # markBadQuality works on mixed meshes because layer elements carry OK_QUALITY
    r = m.mark(which=which, good_quality=good_quality, edge_flags=edge_flags, elem_flags=elem_flags)
    out["edge_flags_in"] = np.zeros(m.ne, np.int32) if edge_flags is None else np.asarray(edge_flags, np.int32)
    out["elem_flags_in"] = np.zeros(m.nelem, np.int32) if elem_flags is None else np.asarray(elem_flags, np.int32)
    out["edge_flags_out"] = r["edge_flags"]
    out["elem_flags_out"] = r["elem_flags"]
    out["counts"] = np.array([r["n_split"], r["n_collapse"], r["n_bad"]], dtype=np.int64)
    out["min_q"] = np.float64(r["min_q"])
    out["max_len"] = np.float64(m.max_edge_length())
    if simplex_only and (np.all(et == refo.TET) or np.all(et == refo.TRIANGLE)):   # 2-D: measure(triangle) / (1/2)
        # ma::getElementWeight (raw SizeField::getWeight, and clamped with refinesLeft = 0,
        # coarsensLeft = 1 -> [0.25, 1]) and what ma::makeSplitVert gives the vertex splitting every SPLIT-marked edge
        out["weights_raw"] = m.weights()
        out["weights_r0_c1"] = m.weights(0, 1)
        if kind in (refo.KIND_ANISO_FIELD, refo.KIND_LOG_FIELD):   # stored vertex fields (user-function fields hold no data)
            se = np.nonzero(r["edge_flags"] & 1)[0]
            sx, sa, sb = m.split_vertices(se)
            out["split_edges"], out["split_xyz"], out["split_b"] = se.astype(np.int32), sx, sb
            if kind == refo.KIND_ANISO_FIELD:
                out["split_a"] = sa
    if not simplex_only:
        # layer elements as ma::getElementWeights weighs them (maBalance.cc:21-81): prisms by their base triangle
        out["layer_weights_raw"], out["layer_weights_r0_c1"], out["prism_base_v"] = m.layer_weights(0, 1)
    if simplex_only and np.all(et == refo.TET):
        # ShortEdgeFixer::shouldApply (maShape.cc:188-219; the class is local to that file: oracle/ref/ref_shape_shim.cc) on
        # the BAD_QUALITY marks: the edge it hands to its ShortEdgeRemover (-1: not applied) and the flag words afterwards
        for ratio in (2.0, 100.0):            # maInput.cc:35,43: the two defaults of maximumEdgeRatio
            m.mark(which=which, good_quality=good_quality, edge_flags=edge_flags, elem_flags=elem_flags)
            out["short_edge_%g" % ratio], out["short_flags_%g" % ratio] = m.short_edge(ratio)
        # ma::getSliverCode / matchSliver of every tet (maShape.cc:35-120) + the first face's own vertex order
        out["sliver_codes"], out["sliver_match"], out["face0_v"] = m.sliver_codes(good_quality)
    if not simplex_only:
        ok, codes = m.layer_ok()
        out["layer_ok"], out["layer_codes"] = ok, codes
        # flags the Adapt constructor put on the mesh before any mark (LAYER closure, maLayer.cc:11-103)
        r0 = m.mark(which=0)
        out["edge_flags_ctor"], out["elem_flags_ctor"] = r0["edge_flags"], r0["elem_flags"]
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print("%-28s nv=%d ne=%d nel=%d counts=%s minq=%.17g maxlen=%.17g" % (
        name, m.nv, m.ne, m.nelem, out["counts"].tolist(), float(out["min_q"]), float(out["max_len"])))


def mixed_box(n, k):
    """n^3 cells; bottom k cell layers are 2 prisms per cell, Kuhn tets above."""
    s = n + 1
    vid = lambda x, y, z: x + s * (y + s * z)
    g = np.arange(s) / n
    xyz = np.stack(np.meshgrid(g, g, g, indexing="ij"), axis=-1)  # [x][y][z]
    xyz = xyz.transpose(2, 1, 0, 3).reshape(-1, 3)
    prisms, tets = [], []
    T = [[0, 1, 2, 6], [0, 2, 3, 6], [0, 3, 7, 6], [0, 7, 4, 6], [0, 4, 5, 6], [0, 5, 1, 6]]
    for z in range(n):
        for y in range(n):
            for x in range(n):
                c = [vid(x, y, z), vid(x + 1, y, z), vid(x + 1, y + 1, z), vid(x, y + 1, z),
                     vid(x, y, z + 1), vid(x + 1, y, z + 1), vid(x + 1, y + 1, z + 1), vid(x, y + 1, z + 1)]
                if z < k:
                    prisms.append([c[0], c[1], c[2], c[4], c[5], c[6]])
                    prisms.append([c[0], c[2], c[3], c[4], c[6], c[7]])
                else:
                    tets += [[c[i] for i in t] for t in T]
    return xyz, np.array(tets, np.int32), np.array(prisms, np.int32)


def prism_pyramid_slab(nx, ny):
    """One layer of nx x ny cells: cell (x, y) with x + y even = 2 prisms, odd = 6 pyramids around the cell centre (every
    face between neighbouring cells is a quad, so the mesh conforms).  Returns xyz, prisms [.,6], pyramids [.,5]."""
    s = nx + 1
    vid = lambda x, y, z: x + s * (y + (ny + 1) * z)
    xyz = [[x / nx, y / ny, z / max(nx, ny)] for z in (0, 1) for y in range(ny + 1) for x in range(s)]
    prisms, pyrs = [], []
    for y in range(ny):
        for x in range(nx):
            c = [vid(x, y, 0), vid(x + 1, y, 0), vid(x + 1, y + 1, 0), vid(x, y + 1, 0),
                 vid(x, y, 1), vid(x + 1, y, 1), vid(x + 1, y + 1, 1), vid(x, y + 1, 1)]
            if (x + y) % 2 == 0:
                prisms += [[c[0], c[1], c[2], c[4], c[5], c[6]], [c[0], c[2], c[3], c[4], c[6], c[7]]]
            else:
                a = len(xyz)
                xyz.append(np.mean([xyz[i] for i in c], axis=0).tolist())
                for q in ([0, 1, 2, 3], [4, 7, 6, 5], [0, 4, 5, 1], [1, 5, 6, 2], [2, 6, 7, 3], [3, 7, 4, 0]):   # bases seen from the apex
                    pyrs.append([c[q[0]], c[q[1]], c[q[2]], c[q[3]], a])
    return np.array(xyz), np.array(prisms, np.int32), np.array(pyrs, np.int32)


def main():
    assert refo.available(), "build oracle/_ref first (make -C oracle/ref)"
    rng = np.random.default_rng(20261017)

    # config 1 shape, small: plain Kuhn box, isotropic user function h = hbar (1 + 2x)
    n = 6
    m = refo.RefMesh.box(n, n, n)
    xyz = m.export()[0]
    case("box6_iso_fn", m, refo.KIND_ISO_FN, fields.iso_linear(xyz, 1.0 / n))
    case("box6_iso_field", m, refo.KIND_ISO_FIELD, fields.iso_linear(xyz, 1.0 / n))
    case("box6_identity", m, refo.KIND_IDENTITY)
    # ma::UniformRefiner (what ma::runUniformRefinement configures): every edge not carrying DONT_SPLIT is marked
    ef = np.zeros(m.ne, np.int32)
    ef[::7] |= 1 << 1
    case("box6_uniform_refiner", m, refo.KIND_UNIFORM, edge_flags=ef)
    h, R = fields.shock_planar(xyz, 1.0 / n)
    case("box6_shock_planar_aniso", m, refo.KIND_ANISO_FN, h, R)
    case("box6_shock_planar_logfn", m, refo.KIND_LOG_FN, h, R)
    m.close()

    # jittered box, rotating shock layer (config 3 field), vertex fields, both interpolations
    n = 7
    m = refo.RefMesh.box(n, n, n)
    xyz = fields.jitter(m.export()[0], 0.3 / n)
    m.set_coords(xyz)
    h, R = fields.shock_rotating(xyz, 1.0 / n)
    case("jbox7_shock_rot_aniso", m, refo.KIND_ANISO_FIELD, h, R)
    case("jbox7_shock_rot_log", m, refo.KIND_LOG_FIELD, h, R)
    # random, slightly non-orthogonal frames and sizes over 1.5 decades: exercises Gram-Schmidt and eigenQR
    R = random_frames(m.nv, rng)
    h = (1.0 / n) * np.exp(rng.uniform(-1.5, 1.5, (m.nv, 3)))
    # incoming flag words: DONT_SPLIT / DONT_COLLAPSE / NEED_NOT_* / OK_QUALITY on random entities
    ef = np.zeros(m.ne, np.int32)
    lf = np.zeros(m.nelem, np.int32)
    ef[rng.random(m.ne) < 0.2] |= 1 << 1      # DONT_SPLIT
    ef[rng.random(m.ne) < 0.2] |= 1 << 3      # DONT_COLLAPSE
    ef[rng.random(m.ne) < 0.1] |= 1 << 17     # NEED_NOT_SPLIT
    ef[rng.random(m.ne) < 0.1] |= 1 << 18     # NEED_NOT_COLLAPSE
    ef[rng.random(m.ne) < 0.1] |= 1 << 9      # DONT_SWAP (unrelated bit must survive)
    lf[rng.random(m.nelem) < 0.3] |= 1 << 6   # OK_QUALITY
    case("jbox7_random_aniso_flags", m, refo.KIND_ANISO_FIELD, h, R, good_quality=0.2, edge_flags=ef, elem_flags=lf)
    case("jbox7_random_log_flags", m, refo.KIND_LOG_FIELD, h, R, good_quality=0.2, edge_flags=ef, elem_flags=lf)
    case("jbox7_random_logfn", m, refo.KIND_LOG_FN, h, R, good_quality=0.2)
    m.close()

    # 2-D box (triangles are the elements: ma::measureTriQuality), jittered in the plane, rotating shock field
    m = refo.RefMesh.box(9, 7, 0)
    xyz = m.export()[0]
    inner = (xyz[:, 0] > 1e-9) & (xyz[:, 0] < 1 - 1e-9) & (xyz[:, 1] > 1e-9) & (xyz[:, 1] < 1 - 1e-9)
    xyz[inner, :2] += (0.3 / 9) * (rng.random((int(inner.sum()), 2)) - 0.5)
    m.set_coords(xyz)
    h, R = fields.shock_rotating(xyz, 1.0 / 8)
    case("tri9x7_shock_rot_aniso", m, refo.KIND_ANISO_FIELD, h, R, good_quality=0.2)
    case("tri9x7_shock_rot_log", m, refo.KIND_LOG_FIELD, h, R, good_quality=0.2)
    case("tri9x7_iso", m, refo.KIND_ISO_FIELD, fields.iso_linear(xyz, 1.0 / 8), good_quality=0.2)
    m.close()

    # mixed prism / tet boundary-layer box (config 5 shape)
    xyz, tets, prisms = mixed_box(5, 2)
    xyz = fields.jitter(xyz, 0.3 / 5, seed=7)
    m = refo.RefMesh.build(xyz, tets=tets, prisms=prisms)
    h, R = fields.shock_rotating(xyz, 1.0 / 5)
    case("mixed5_shock_rot_aniso", m, refo.KIND_ANISO_FIELD, h, R)
    m.close()
    # the same mesh jittered until some prisms become unsafe: ma::isPrismOk only (constructing an ma::Adapt on a
    # mesh with unsafe layer elements crashes inside the compiled reference's reporting path, so no marks here)
    xyz, tets, prisms = mixed_box(5, 2)
    xyz = fields.jitter(xyz, 0.9 / 5, seed=7)
    m = refo.RefMesh.build(xyz, tets=tets, prisms=prisms)
    _, ev, et, elv = m.export()
    ok, codes = m.layer_ok()
    np.savez_compressed(os.path.join(HERE, "mixed5_unsafe_layer.npz"), xyz=xyz, edge_v=ev, elem_type=et, elem_v=elv,
                        layer_ok=ok, layer_codes=codes)
    print("mixed5_unsafe_layer          unsafe prisms: %d of %d" % (int((ok == 0).sum()), len(prisms)))
    m.close()

    # prisms + pyramids (ma::isPyramidOk, LAYER closure over pyramid edges): a mildly jittered slab through a real ma::Adapt,
    # and the same slab with the apexes and corners thrown far enough that pyramids become unsafe (isPyramidOk only)
    rng2 = np.random.default_rng(20261018)          # its own stream: the fixtures generated after it keep their values
    xyz, prisms, pyrs = prism_pyramid_slab(4, 3)
    nc = 2 * 5 * 4                                   # corner vertices; the rest are pyramid apexes
    x1 = xyz.copy()
    x1[nc:] += 0.05 * (rng2.random((len(xyz) - nc, 3)) - 0.5)
    m = refo.RefMesh.build(x1, prisms=prisms, pyramids=pyrs)
    h, R = fields.shock_rotating(x1, 1.0 / 4)
    case("pyrslab_shock_rot_aniso", m, refo.KIND_ANISO_FIELD, h, R)
    m.close()
    xyz, prisms, pyrs = prism_pyramid_slab(6, 5)
    nc = 2 * 7 * 6
    x2 = xyz.copy()
    x2 += 0.2 * (rng2.random(xyz.shape) - 0.5) * np.array([1.0, 1.0, 0.5])
    x2[nc:] += 0.3 * (rng2.random((len(xyz) - nc, 3)) - 0.5)
    m = refo.RefMesh.build(x2, prisms=prisms, pyramids=pyrs)
    _, ev, et, elv = m.export()
    ok, codes = m.layer_ok()
    np.savez_compressed(os.path.join(HERE, "pyrslab_unsafe_layer.npz"), xyz=x2, edge_v=ev, elem_type=et, elem_v=elv,
                        layer_ok=ok, layer_codes=codes)
    py = et == refo.PYRAMID
    print("pyrslab_unsafe_layer         unsafe pyramids: %d of %d, good rotations %s; unsafe prisms %d" % (
        int((ok[py] == 0).sum()), int(py.sum()), np.bincount(codes[py] + 1, minlength=3).tolist(), int((ok[~py] == 0).sum())))
    m.close()

    # native-format fixtures written by the reference itself (mds_write_smb) with the arrays its own API exports from the
    # same meshes: pins core_b200/smb.py
    import shutil
    import tempfile
    tmp = tempfile.mkdtemp()
    m = refo.RefMesh.box(3, 2, 2)
    xyz = fields.jitter(m.export()[0], 0.2 / 3)
    m.set_coords(xyz)
    h, R = fields.shock_rotating(xyz, 1.0 / 3)
    m.store_fields(h, R)                       # vertex fields "sizes" / "frames" travel as tags
    m.write_smb(os.path.join(tmp, "kbox322.smb"))
    shutil.copy(os.path.join(tmp, "kbox3220.smb"), os.path.join(HERE, "kbox322_0.smb"))
    _, ev, et, elv = m.export()
    m.set_sizefield(refo.KIND_ANISO_FIELD, h, R)
    np.savez_compressed(os.path.join(HERE, "smb_kbox322.npz"), xyz=xyz, edge_v=ev, elem_type=et, elem_v=elv, h=h, R=R,
                        lengths=m.lengths(), qualities=m.qualities(True))
    m.close()
    xyz, tets, prisms = mixed_box(3, 1)
    m = refo.RefMesh.build(xyz, tets=tets, prisms=prisms)
    m.write_smb(os.path.join(tmp, "mixed3.smb"))
    shutil.copy(os.path.join(tmp, "mixed30.smb"), os.path.join(HERE, "mixed3_0.smb"))
    _, ev, et, elv = m.export()
    np.savez_compressed(os.path.join(HERE, "smb_mixed3.npz"), xyz=xyz, edge_v=ev, elem_type=et, elem_v=elv)
    m.close()
    shutil.rmtree(tmp)
    print("smb fixtures: kbox322_0.smb, mixed3_0.smb")

    # apf::eigen on the six matrices of test/eigen_test.cc:13-56 plus random symmetric ones
    A6 = np.array([
        [[1.001575e+00, -3.138397e-01, 8.107355e-01], [-3.138397e-01, 4.946182e-01, -1.860431e+00], [8.107355e-01, -1.860431e+00, 7.582283e+00]],
        [[1.668786e+00, -7.051850e-02, 4.561114e-01], [-7.051850e-02, 1.022233e+00, -5.465726e-01], [4.561114e-01, -5.465726e-01, 1.128792e+00]],
        [[7.139854e-01, 1.220499e+00, -8.923666e-02], [1.220499e+00, 2.289174e+00, -4.099688e-01], [-8.923666e-02, -4.099688e-01, 6.159893e-01]],
        [[1.662534e+00, 1.834307e+00, 1.113124e-01], [1.834307e+00, 4.465579e+00, 2.825547e+00], [1.113124e-01, 2.825547e+00, 3.277306e+00]],
        [[2.869781e+00, -1.879048e+00, 1.982487e+00], [-1.879048e+00, 1.296580e+00, -9.613331e-01], [1.982487e+00, -9.613331e-01, 5.662202e+00]],
        [[2.763219e+00, 1.444043e+00, 1.102805e+00], [1.444043e+00, 4.883458e+00, 4.285960e+00], [1.102805e+00, 4.285960e+00, 5.790849e+00]]])
    B = rng.standard_normal((200, 3, 3))
    B = B + B.transpose(0, 2, 1)
    B[:20] *= 1e-3
    B[20:40] *= 1e3
    for i in range(40, 60):   # already (block-)diagonal inputs exercise the deflation branches
        B[i, 0, 1] = B[i, 1, 0] = B[i, 0, 2] = B[i, 2, 0] = 0.0
    for i in range(60, 70):
        B[i] = np.diag(np.diag(B[i]))
    A = np.concatenate([A6, B])
    vals = np.zeros((len(A), 3))
    vecs = np.zeros((len(A), 3, 3))
    for i in range(len(A)):
        vals[i], vecs[i] = refo.eigen(A[i])
    np.savez_compressed(os.path.join(HERE, "eigen.npz"), A=A, vals=vals, vecs=vecs)
    print("eigen: %d matrices" % len(A))


if __name__ == "__main__":
    main()
