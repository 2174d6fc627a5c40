"""World-size-2 host-side test of the multi-part path on CPU (gloo): each rank builds its slab part, evaluates its
edges with the restated oracle, exchanges the shared-edge flag words in link order (the host-side twin of
mag_reconcile_edge_flags: pack by per-peer index list -> send/recv -> compare, owner wins) and reduces the
statistics (sum of owned counts, max length) -- the result must equal the serial sweep of the glued box.  This
covers the link ordering / ownership / reduction logic the NCCL path relies on."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _field(xyz, hbar):
    # a function of the GLOBAL position only, so shared vertices get bit-identical values on both parts
    import core_b200.fields as fields
    f = xyz.copy()
    f[:, 0] = f[:, 0] / 2.0
    return fields.shock_rotating(f, hbar)


def _worker(rank, world, port, gnx, ny, nz, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import core_b200.boxmesh as boxmesh
    from oracle import mao
    import util
    part = boxmesh.slab_part(gnx, ny, nz, world, rank, wx=2.0)
    h, R = _field(part["xyz"], 1.0 / ny)
    r = util.oracle_sweep(mao.ANISO, part["xyz"], h, R, part["edge_v"], part["tet_v"], edge_owned=part["edge_owned"])
    ef = r["edge_flags"]
    mismatches = 0
    for peer, idx, peer_owns in part["links"]:
        send = torch.from_numpy(ef[idx].copy())
        recv = torch.empty_like(send)
        ops = [dist.P2POp(dist.isend, send, peer), dist.P2POp(dist.irecv, recv, peer)]
        for w in dist.batch_isend_irecv(ops):
            w.wait()
        got = recv.numpy()
        mismatches += int(np.count_nonzero(got != ef[idx]))
        take = peer_owns.astype(bool)
        ef[idx[take]] = got[take]
    counts = torch.tensor([r["n_split"], r["n_collapse"], r["n_bad"], mismatches], dtype=torch.int64)
    dist.all_reduce(counts, op=dist.ReduceOp.SUM)
    mx = torch.tensor([r["max_length"], -r["min_quality"]], dtype=torch.float64)
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    if rank == 0:
        q.put((counts.tolist(), mx.tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_part_sweep_equals_serial(built):
    from oracle import mao
    import core_b200.boxmesh as boxmesh
    import util
    gnx, ny, nz, world = 6, 4, 3, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, gnx, ny, nz, q)) for r in range(world)]
    for p in procs:
        p.start()
    counts, mx = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    xyz, ev, tv = boxmesh.kuhn_box(gnx, ny, nz, wx=2.0)
    h, R = _field(xyz, 1.0 / ny)
    s = util.oracle_sweep(mao.ANISO, xyz, h, R, ev, tv)
    assert counts[:3] == [s["n_split"], s["n_collapse"], s["n_bad"]]
    assert counts[3] == 0, "part-boundary copies disagreed"
    assert mx[0] == s["max_length"] and -mx[1] == s["min_quality"]
    assert s["n_split"] > 0 and s["n_bad"] > 0
