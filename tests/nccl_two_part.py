"""Two parts on two GPUs (launched by torchrun from test_gpu_parity.py::test_nccl_two_part_device): the device twin of
tests/test_multipart_gloo.py.  Each rank sweeps its slab part on its own GPU through the C ABI, then
  1. mag_reconcile_edge_flags: 0 disagreeing copies; the NCCL-reduced statistics equal the serial oracle sweep of the
     glued box (sum of owned counts, min quality, max length) -- ma::checkFlagConsistency + PCU Add/Min/Max;
  2. a flipped bit on the NON-owner's copies is counted on both sides and the owner's word wins on both;
  3. mag_sync_edge_flags ORs a bit set on one copy into every copy (ma::syncFlag, maAdapt.cc:498-520).
Prints one line "NCCL_TWO_PART_OK" from rank 0."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")
    import core_b200 as cb
    from oracle import mao
    import util
    from test_multipart_gloo import _field
    gnx, ny, nz = 6, 4, 3
    part = cb.boxmesh.slab_part(gnx, ny, nz, world, rank, wx=2.0)
    h, R = _field(part["xyz"], 1.0 / ny)
    p = cb.Part(local)
    p.set_mesh(part["xyz"], part["edge_v"], part["tet_v"], edge_owned=part["edge_owned"])
    p.set_size_field_aniso(h, R)
    uid = [cb.Part.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    p.comm_init(world, rank, uid[0])
    try:   # an index past the part's edges is refused on the host
        p.set_edge_links([(1 - rank, np.array([len(part["edge_v"])], np.int32), None)])
        raise AssertionError("out-of-range link index accepted")
    except cb.MagError as e:
        assert e.code == 2
    p.set_edge_links(part["links"])
    mask = cb.SPLIT | cb.COLLAPSE | cb.NEED_NOT_SPLIT | cb.NEED_NOT_COLLAPSE
    # 1. consistent copies, global statistics
    p.clear_flags()
    p.sweep(cb.OP_ALL & ~cb.OP_LAYER_CHECK, fp_mode=cb.FP_STRICT)
    p.reconcile_edge_flags(mask)
    g = p.allreduce_stats()
    xyz, ev, tv = cb.boxmesh.kuhn_box(gnx, ny, nz, wx=2.0)
    hs, Rs = _field(xyz, 1.0 / ny)
    s = util.oracle_sweep(mao.ANISO, xyz, hs, Rs, ev, tv)
    assert g["n_flag_mismatch"] == 0
    assert (g["n_split"], g["n_collapse"], g["n_bad"]) == (s["n_split"], s["n_collapse"], s["n_bad"]), (g, s["n_split"])
    assert g["min_quality"] == s["min_quality"] and g["max_length"] == s["max_length"]
    ef_ok, lf = p.flags()
    # 1b. the same through mag_sweep_reconciled (exchange overlapped with the element sweep), both arithmetic modes
    for mode in (cb.FP_STRICT, cb.FP_FAST):
        p.clear_flags()
        p.sweep(cb.OP_ALL & ~cb.OP_LAYER_CHECK, fp_mode=mode, reconcile_mask=mask)
        g2 = p.allreduce_stats()
        ef_b, lf_b = p.flags()
        assert np.array_equal(ef_b, ef_ok) and np.array_equal(lf_b, lf)
        for key in ("n_split", "n_collapse", "n_bad", "n_flag_mismatch"):
            assert g2[key] == g[key], key
    # 2. the non-owner's copies flip SPLIT: counted on both sides, owner wins
    ef = ef_ok.copy()
    flipped = 0
    for peer, idx, peer_owns in part["links"]:
        sel = idx[peer_owns.astype(bool)]
        ef[sel] ^= cb.SPLIT
        flipped += len(sel)
    n_flipped = torch.tensor([flipped], dtype=torch.int64)
    dist.all_reduce(n_flipped)
    assert int(n_flipped) > 0
    p.set_flags(ef, lf)
    p.sweep(cb.OP_LENGTHS, fp_mode=cb.FP_STRICT)       # a sweep that marks nothing: resets the statistics, keeps the words
    p.reconcile_edge_flags(mask)
    g = p.allreduce_stats()
    assert g["n_flag_mismatch"] == 2 * int(n_flipped), (g["n_flag_mismatch"], int(n_flipped))   # each copy of a pair reports it
    ef2, _ = p.flags()
    assert np.array_equal(ef2, ef_ok), "the owner's word did not win"
    # 3. syncFlag: rank 0 alone sets DONT_SWAP on its copies of the shared edges
    ef = ef_ok.copy()
    shared = np.unique(np.concatenate([idx for _, idx, _ in part["links"]])) if part["links"] else np.zeros(0, np.int64)
    if rank == 0:
        ef[shared] |= cb.DONT_SWAP
    p.set_flags(ef, lf)
    p.sync_edge_flags(cb.DONT_SWAP)
    ef3, _ = p.flags()
    want = ef_ok.copy()
    want[shared] |= cb.DONT_SWAP
    assert np.array_equal(ef3, want), "syncFlag did not OR the bit into every copy"
    p.close()
    dist.barrier()
    if rank == 0:
        print("NCCL_TWO_PART_OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
