"""Quick GPU sanity: CUDA path vs restated oracle on a jittered box, every kind, both fp modes."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import core_b200 as cb
from oracle import mao

n = int(sys.argv[1]) if len(sys.argv) > 1 else 12
xyz, ev, tv = cb.boxmesh.kuhn_box(n, n, n)
hbar = 1.0 / n
xyz = cb.fields.jitter(xyz, 0.3 * hbar)
nv = len(xyz)
rng = np.random.default_rng(1)
A = rng.standard_normal((nv, 3, 3)); Qm, _ = np.linalg.qr(A); Qm[:, :, 2] *= np.sign(np.linalg.det(Qm))[:, None]
R = (Qm + 1e-3 * rng.standard_normal((nv, 3, 3))).reshape(nv, 9)
H = hbar * np.exp(rng.uniform(-1.5, 1.5, (nv, 3)))
s = hbar * np.exp(rng.uniform(-1, 1, nv))
lm = mao.logm_from_frames(H, R, 0)
p = cb.Part(0)
p.set_mesh(xyz, ev, tv)
def rel(a, b): return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))
ok = True
for name, kind, ma, mb, setter in (("identity", mao.IDENTITY, None, None, lambda: p.set_size_field_identity()),
                                   ("iso", mao.ISO, s, None, lambda: p.set_size_field_iso(s)),
                                   ("aniso", mao.ANISO, H, R, lambda: p.set_size_field_aniso(H, R)),
                                   ("logm", mao.LOGM, None, lm, lambda: p.set_size_field_logm(lm))):
    setter()
    L0 = mao.edge_lengths(kind, xyz, ma, mb, ev)
    q0 = mao.tet_qualities(kind, xyz, ma, mb, tv)
    ef0 = np.zeros(len(ev), np.int32); lf0 = np.zeros(len(tv), np.int32)
    ns = mao.mark_edges_to_split(L0, ef0, None, kind); nc = mao.mark_edges_to_collapse(L0, ef0, None, kind); nb = mao.mark_bad_quality(q0, lf0, 0.027)
    for mode in (cb.FP_STRICT, cb.FP_FAST):
        p.set_flags(None, None)
        p.sweep(cb.OP_ALL & ~cb.OP_LAYER_CHECK, fp_mode=mode)
        st = p.stats()
        L, q = p.edge_lengths(), p.qualities()
        ef, lf = p.flags()
        line = "%-8s mode=%d len_exact=%s rel=%.2e q_exact=%s rel=%.2e flags=%s/%s counts=%s near=%d minq=%s maxl=%s" % (
            name, mode, np.array_equal(L, L0), rel(L, L0), np.array_equal(q, q0), rel(q, q0),
            np.array_equal(ef, ef0), np.array_equal(lf, lf0),
            (st["n_split"], st["n_collapse"], st["n_bad"]) == (ns, nc, nb), st["n_near_threshold"],
            st["min_quality"] == mao.min_quality(q0) if mode == 0 else "-", st["max_length"] == mao.max_length(L0) if mode == 0 else "-")
        print(line)
        if not (np.array_equal(ef, ef0) and np.array_equal(lf, lf0)): ok = False
print("OK" if ok else "FLAG MISMATCH")
