#!/usr/bin/env python
"""bench.py -- MeshAdapt marking / quality sweep throughput on B200 (and the reference's CPU path).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--n CELLS] [--fp strict|fast] [--field aniso|logm|iso]
    python bench.py --impl reference [--steps K] [--warmup W]          # the reference's own CPU implementation
    torchrun --nproc-per-node N bench.py --gpus N ...                  # one rank (= one PUMI part) per GPU

Workload (BASELINE.json configs[2], the configuration the north-star target is quoted on): a
Kuhn box of n^3 cells (n = 203: 50,192,562 tets, 58,929,479 edges, 8,489,664 vertices) per GPU with
the vertex-stored rotating-shock-layer anisotropic size field (sizes VECTOR + frames MATRIX,
AnisoSizeField).  One step = one pass of the hot path over the part: incoming flag words cleared
(device memset), per-vertex transforms, every edge's metric length + SPLIT/COLLAPSE marking, every
tet's mean-ratio quality + BAD_QUALITY marking, the reduced statistics read back by the host
(and, for N > 1, the NCCL part-boundary edge-flag reconciliation + global statistics).
Metric: (edges + elements) evaluated per second, whole job.  All device inputs are far larger than
L2 (>= 4 GB per step vs 126 MB), so no explicit L2 flush is needed between timed steps.

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every key.
"""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

METRIC = "edges+elements evaluated/sec"
UNIT = "entities/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    # "--cells" is the spelling to use under torchrun, whose own parser rejects "--n" as an ambiguous abbreviation
    ap.add_argument("--n", "--cells", dest="n", type=int, default=203, help="cells per side of each part's box")
    ap.add_argument("--fp", default=os.environ.get("MAG_BENCH_FP", "fast"), choices=["strict", "fast"])
    ap.add_argument("--field", default="aniso", choices=["aniso", "logm", "iso"])
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--cpu-n", type=int, default=0, help="box size of the CPU-baseline sample (0 = auto)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--jitter", type=float, default=0.0,
                    help="move interior vertices by this fraction of the cell size (SURVEY 8d optional perturbed variant)")
    ap.add_argument("--workload", default="box", choices=["box", "mixed"],
                    help="box = BASELINE configs[2] (the headline); mixed = configs[4] as the main timed workload (use --n 120)")
    ap.add_argument("--layers", type=int, default=12, help="prism layers of the mixed workload")
    ap.add_argument("--mixed-n", type=int, default=120, help="cells per side of the mixed extra (configs[4]: ~10 M elements)")
    ap.add_argument("--no-extras", action="store_true",
                    help="only the headline workload (skips strict / jittered / LogAniso / mixed extras and the N>1 parity check)")
    return ap.parse_args()


# --------------------------------------------------------------------------- workload
def workload_name(n, field):
    return "kuhn-box n=%d per part (%d tets, %d edges), rotating shock-layer %s size field, fp64" % (
        n, 6 * n ** 3, box_ne(n), {"aniso": "AnisoSizeField(sizes+frames)", "logm": "LogAnisoSizeField",
                                    "iso": "IsoSizeField"}[field])


def box_ne(n):
    return 3 * n * (n + 1) ** 2 + 3 * n * n * (n + 1) + n ** 3


def algorithmic_bytes(nv, ne, nt, field):
    """every array touched once.  F = per-vertex metric bytes."""
    F = {"aniso": 96, "logm": 72, "iso": 8}[field]
    vert = nv * (24 + F)
    return dict(total=vert + 24 * ne + 32 * nt, edge_kernel=vert + 24 * ne, elem_kernel=vert + 32 * nt)


# --------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "20"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.p = None

    def stop(self):
        if not self.p:
            return None
        time.sleep(0.15)
        self.p.terminate()
        try:
            out, _ = self.p.communicate(timeout=5)
        except Exception:
            self.p.kill()
            return None
        sm, mx, reasons = [], [], set()
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return None
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "samples": len(sm),
                "reasons": sorted(reasons)}


# --------------------------------------------------------------------------- CPU baseline (reference / port)
def cpu_sweep_reference(n):
    """The UNMODIFIED reference (oracle/_ref) on an n^3 box of the same field: lengths, qualities, 3 marks.
    Returns (entities, seconds).  The only place outside tests where oracle/ is executed: it is the
    thing being timed as the CPU baseline, never the product path."""
    from oracle import refo
    import core_b200.fields as fields
    m = refo.RefMesh.box(n, n, n)
    xyz, _, _, _ = m.export()
    h, R = fields.shock_rotating(xyz, 1.0 / n)
    m.set_sizefield(refo.KIND_ANISO_FIELD, h, R)
    t0 = time.perf_counter()
    m.lengths()
    m.qualities()
    m.mark(which=7)
    dt = time.perf_counter() - t0
    ents = m.ne + m.nelem
    m.close()
    return ents, dt


def cpu_sweep_port(n):
    from oracle import mao
    import core_b200.boxmesh as boxmesh
    import core_b200.fields as fields
    xyz, ev, tv = boxmesh.kuhn_box(n, n, n)
    h, R = fields.shock_rotating(xyz, 1.0 / n)
    t0 = time.perf_counter()
    L = mao.edge_lengths(mao.ANISO, xyz, h, R, ev)
    q = mao.tet_qualities(mao.ANISO, xyz, h, R, tv)
    ef = np.zeros(len(ev), np.int32)
    lf = np.zeros(len(tv), np.int32)
    mao.mark_edges_to_split(L, ef)
    mao.mark_edges_to_collapse(L, ef)
    mao.mark_bad_quality(q, lf, 0.027)
    return len(ev) + len(tv), time.perf_counter() - t0


def cpu_kind():
    from oracle import refo
    return "reference" if refo.available() else "port"


def cpu_baseline(n_sample):
    kind = cpu_kind()
    ents, dt = cpu_sweep_reference(n_sample) if kind == "reference" else cpu_sweep_port(n_sample)
    return {"value": ents / dt, "unit": UNIT, "cores": 1, "kind": kind,
            "sample": "one full sweep (lengths + qualities + split/collapse/bad marks) of the same size field on an "
                      "n=%d box (%d entities) in %.1f s, 1 process = 1 core (the reference is single-threaded per rank)"
                      % (n_sample, ents, dt)}


def _ref_worker(kind, n, conn):
    """One reference 'rank': builds its own n^3 part once, then sweeps it each time it is told to."""
    import core_b200.fields as fields
    if kind == "reference":
        from oracle import refo
        m = refo.RefMesh.box(n, n, n)
        xyz, _, _, _ = m.export()
        h, R = fields.shock_rotating(xyz, 1.0 / n)
        m.set_sizefield(refo.KIND_ANISO_FIELD, h, R)
        ents = m.ne + m.nelem

        def sweep():
            m.lengths()
            m.qualities()
            m.mark(which=7)
    else:
        from oracle import mao
        import core_b200.boxmesh as boxmesh
        xyz, ev, tv = boxmesh.kuhn_box(n, n, n)
        h, R = fields.shock_rotating(xyz, 1.0 / n)
        ents = len(ev) + len(tv)

        def sweep():
            L = mao.edge_lengths(mao.ANISO, xyz, h, R, ev)
            q = mao.tet_qualities(mao.ANISO, xyz, h, R, tv)
            ef = np.zeros(len(ev), np.int32)
            lf = np.zeros(len(tv), np.int32)
            mao.mark_edges_to_split(L, ef)
            mao.mark_edges_to_collapse(L, ef)
            mao.mark_bad_quality(q, lf, 0.027)
    conn.send(ents)
    while conn.recv():
        t0 = time.perf_counter()
        sweep()
        conn.send(time.perf_counter() - t0)


def run_reference_arm(a):
    """--impl reference: the reference's own CPU implementation on all host cores: P independent
    single-threaded processes (the reference has no threads; its parallel unit is one part per
    process), each sweeping its own resident n^3 sample part, no communication."""
    import multiprocessing as mp
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    kind = cpu_kind()
    P = max(1, min(os.cpu_count() or 1, 64))
    # size the sample so (steps + warmup) sweeps stay within ~2 minutes at ~9 us / entity / core
    budget = 110.0 / max(1, a.steps + a.warmup)
    n = a.cpu_n or int(max(8, min(40, round((budget / 9e-6 / 13.0) ** (1.0 / 3.0)))))
    ctx = mp.get_context("spawn")
    workers = []
    for _ in range(P):
        parent, child = ctx.Pipe()
        pr = ctx.Process(target=_ref_worker, args=(kind, n, child), daemon=True)
        pr.start()
        workers.append((pr, parent))
    ents_each = [c.recv() for _, c in workers]

    def step():
        for _, c in workers:
            c.send(True)
        return [c.recv() for _, c in workers]

    for _ in range(a.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        step()
    dt = time.perf_counter() - t0
    for pr, c in workers:
        c.send(False)
        pr.join(timeout=10)
    ents = sum(ents_each) * a.steps
    value = ents / dt
    sample = ("%d independent single-threaded processes, each sweeping (lengths + qualities + split/collapse/bad "
              "marks) its own resident n=%d box (%d entities) once per step" % (P, n, ents_each[0]))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": 1e3 * dt / max(1, a.steps), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "bounded sample of [%s]: the same size field on an n=%d box per process (%d entities each, %d "
                                   "processes), swept once per step" % (workload_name(a.n, a.field), n, ents_each[0], P),
                       "sample_n": n, "processes": P, "host_cpus": os.cpu_count(), "cpu_sample": sample},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": P, "kind": kind, "sample": sample,
                             "value_per_core": value / P},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def bind_to_gpu_numa_node(gpu_index):
    """Runs this rank on the CPUs nearest to its GPU (NVML's ideal CPU affinity), so that the pinned host buffers of the
    end-to-end leg are allocated on that NUMA node and their copies do not cross the socket interconnect."""
    if os.environ.get("MAG_BENCH_NO_BIND"):
        return
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
    except Exception:
        pass


# --------------------------------------------------------------------------- B200 arm
def _pin(torch, x):
    return None if x is None else torch.from_numpy(np.ascontiguousarray(x)).pin_memory()


class Timed:
    """K steps of `step` on `stream`, bracketed by barrier + synchronize, CUDA events on the launching stream,
    max over ranks; per-kernel times from the events mag_sweep records itself."""

    def __init__(self, torch, dist, world, stream):
        self.torch, self.dist, self.world, self.stream = torch, dist, world, stream

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        if self.world == 1:
            return x
        t = self.torch.tensor([x], device="cuda", dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def run(self, p, step, steps, warmup):
        torch = self.torch
        st = None
        for _ in range(warmup):
            st = step()
        l0 = p.launch_count()
        p.timing_begin(steps)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.barrier()
        with torch.cuda.stream(self.stream):
            ev0.record(self.stream)
            for _ in range(steps):
                st = step()
            ev1.record(self.stream)
        self.barrier()
        ms = self.max_over_ranks(ev0.elapsed_time(ev1)) / steps
        kt = p.timing_read()
        return {"ms_per_step": ms, "vertex_ms": float(kt[:, 0].mean()), "edge_ms": float(kt[:, 1].mean()),
                "elem_ms": float(kt[:, 2].mean()), "stats": st, "launches": p.launch_count() - l0}


def _peak():
    try:
        peaks = json.load(open(os.path.join(HERE, "MEASURED_PEAKS.json")))
        return float(peaks["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback 6650 GB/s (B200_PROFILING.md)"


def _kernel_names(cb, field, fp, legacy):
    """Names as ncu prints them (profiles/traffic.json keys).  The full fast sweep over zero incoming flag words runs the
    lean row kernels (mag_lean.cuh; the log-Euclidean edge kernel stays with the tiles), everything else the tile kernels."""
    kind = {"iso": 1, "aniso": 2, "logm": 3}[field]
    fast = 1 if fp == "fast" else 0
    tiles = "k_edges<%d, %d, 0>" % (kind, fast), "k_tets<%d, %d, 1>" % (kind, fast)
    if legacy or not fast:
        return tiles
    tet = "k_tet_rows_z<%d>" % kind if os.environ.get("MAG_TET_WINNER") == "0" else "k_tet_rows_w<%d>" % kind
    return (tiles[0] if field == "logm" else "k_edge_rows_z<%d>" % kind), tet


def _roofline(r, nv, ne, nt, field, fp, n, world, jitter):
    peak, peak_src = _peak()
    ab = algorithmic_bytes(nv, ne, nt, field)
    legacy = os.environ.get("MAG_LEGACY_SWEEP") == "1"
    k_edge, k_elem = _kernel_names(None, field, fp, legacy)
    dom_is_edge = r["edge_ms"] >= r["elem_ms"]
    dom = k_edge if dom_is_edge else k_elem
    dom_ms = max(r["edge_ms"], r["elem_ms"])
    dom_bytes = ab["edge_kernel"] if dom_is_edge else ab["elem_kernel"]
    achieved = dom_bytes / (dom_ms * 1e-3) / 1e9
    # DRAM bytes per launch of that kernel: from the committed ncu --set full capture of THIS kernel on THIS workload
    # (profiles/traffic.json, written by scripts/profile_summary.py); null when no capture of that exact kernel exists
    traffic, units = None, None
    try:
        if n == 203 and world == 1 and jitter == 0:
            tj = json.load(open(os.path.join(HERE, "profiles", "traffic.json")))
            if dom in tj and tj[dom].get("workload", "aniso") == field:
                traffic = float(tj[dom]["dram_bytes_per_launch"])
                units = {k: tj[dom].get(k) for k in ("fp64_pipe_pct", "l1_data_pipe_pct", "dram_pct", "warps_active_pct", "source")}
    except Exception:
        traffic, units = None, None
    return {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
            "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
            "algorithmic_bytes_per_launch": int(dom_bytes), "kernel_ms": dom_ms,
            "kernel_ms_all": {"vertex_pass": r["vertex_ms"], "edges": r["edge_ms"], "elements": r["elem_ms"]},
            "step_algorithmic_bytes": int(ab["total"]),
            "step_frac": ab["total"] / (r["ms_per_step"] * 1e-3) / 1e9 / peak,
            "unit_utilisation_ncu": units}


def parity_multi(cb, torch, dist, world, rank, local):
    """Device-side parity of the multi-part path, run before the timed steps when N > 1 (so the driver's scaling run
    carries it): a small global box cut into `world` slab parts, one per GPU, against the SAME box swept as one part on
    rank 0's GPU (the single-part sweep is pinned on the reference in tests/; nothing under oracle/ is touched here).
      1. reconciled flags: 0 disagreements; NCCL-reduced owned counts / min quality / max length == the one-part sweep
         (ma::checkFlagConsistency + PCU Add / Min / Max, maAdapt.cc:226-256,323; maShape.cc:152-169);
      2. both copies of every shared edge carry the same word as the one-part sweep gives that edge;
      3. a bit flipped on the NON-owner's copies is counted (once per copy) and the owner's word wins everywhere;
      4. mag_sync_edge_flags ORs a bit set on one copy into all copies (ma::syncFlag, maAdapt.cc:498-520)."""
    gnx, ny, nz = 6 * world, 5, 4
    part = cb.boxmesh.slab_part(gnx, ny, nz, world, rank, wx=float(world))
    hbar = 1.0 / ny

    def field(xyz):
        f = xyz.copy()
        f[:, 0] = 1.0 - np.abs(1.0 - np.mod(xyz[:, 0], 2.0))
        return cb.fields.shock_rotating(f, hbar)

    h, R = field(part["xyz"])
    p = cb.Part(local)
    p.set_mesh(part["xyz"], part["edge_v"], part["tet_v"], edge_owned=part["edge_owned"])
    p.set_size_field_aniso(h, R)
    uid = [cb.Part.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    p.comm_init(world, rank, uid[0])
    p.set_edge_links(part["links"])
    mask = cb.SPLIT | cb.COLLAPSE | cb.NEED_NOT_SPLIT | cb.NEED_NOT_COLLAPSE
    ops = cb.OP_ALL & ~cb.OP_LAYER_CHECK
    out = {}
    for mode, name in ((cb.FP_STRICT, "strict"), (cb.FP_FAST, "fast")):
        p.clear_flags()
        p.sweep(ops, fp_mode=mode)
        p.reconcile_edge_flags(mask)
        g = p.allreduce_stats()
        ef_ok, lf = p.flags()
        # the glued box as ONE part on rank 0
        serial = [None]
        if rank == 0:
            xyz, ev, tv = cb.boxmesh.kuhn_box(gnx, ny, nz, wx=float(world))
            hs, Rs = field(xyz)
            s = cb.Part(local)
            s.set_mesh(xyz, ev, tv)
            s.set_size_field_aniso(hs, Rs)
            s.clear_flags()
            s.sweep(ops, fp_mode=mode)
            ss = s.stats()
            sf = s.flags()[0]
            s.close()
            # word of a global edge, keyed by its end-vertex coordinates' grid ids
            serial[0] = (ss, xyz, ev, sf)
        dist.broadcast_object_list(serial, src=0)
        ss, gxyz, gev, gsf = serial[0]
        ok = g["n_flag_mismatch"] == 0
        ok &= (g["n_split"], g["n_collapse"], g["n_bad"]) == (ss["n_split"], ss["n_collapse"], ss["n_bad"])
        ok &= g["min_quality"] == ss["min_quality"] and g["max_length"] == ss["max_length"]
        # 2. every local edge's word == the word of the same edge of the glued box (match by global vertex grid ids)
        sxg = gnx + 1
        sxl = part["nx"] + 1
        def gid(v):   # local vertex id -> global vertex id
            v = v.astype(np.int64)
            return (v % sxl + part["x0"]) + sxg * (v // sxl)
        a, b = gid(part["edge_v"][:, 0]), gid(part["edge_v"][:, 1])
        nvg = len(gxyz)
        key_l = a * nvg + b
        key_g = gev[:, 0].astype(np.int64) * nvg + gev[:, 1].astype(np.int64)
        order = np.argsort(key_g)
        pos = order[np.searchsorted(key_g[order], key_l)]
        ok &= bool(np.array_equal(key_g[pos], key_l)) and bool(np.array_equal(ef_ok & mask, gsf[pos] & mask))
        out[name] = bool(ok)
        if mode == cb.FP_STRICT:
            # 3. injected disagreement on the non-owner's copies
            ef = ef_ok.copy()
            flipped = 0
            for peer, idx, peer_owns in part["links"]:
                sel = idx[peer_owns.astype(bool)]
                ef[sel] ^= cb.SPLIT
                flipped += len(sel)
            t = torch.tensor([flipped], device="cuda", dtype=torch.int64)
            dist.all_reduce(t)
            p.set_flags(ef, lf)
            p.sweep(cb.OP_LENGTHS, fp_mode=cb.FP_STRICT)
            p.reconcile_edge_flags(mask)
            g2 = p.allreduce_stats()
            ef2, _ = p.flags()
            out["owner_wins"] = bool(int(t.item()) > 0 and g2["n_flag_mismatch"] == 2 * int(t.item()) and np.array_equal(ef2, ef_ok))
            # 4. syncFlag
            ef = ef_ok.copy()
            shared = np.unique(np.concatenate([idx for _, idx, _ in part["links"]])) if part["links"] else np.zeros(0, np.int64)
            if rank == 0:
                ef[shared] |= cb.DONT_SWAP
            p.set_flags(ef, lf)
            p.sync_edge_flags(cb.DONT_SWAP)
            ef3, _ = p.flags()
            want = ef_ok.copy()
            if rank <= 1:
                want[shared if rank == 0 else part["links"][0][1]] |= cb.DONT_SWAP   # rank 1 shares a plane with rank 0
            out["sync_flag"] = bool(np.array_equal(ef3, want))
    p.close()
    t = torch.tensor([int(all(out.values()))], device="cuda", dtype=torch.int64)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    out["status"] = "ok" if int(t.item()) == 1 else "FAILED"
    out["what"] = ("%d x %d x %d box in %d slab parts vs the same box as one part on one GPU: counts, min/max, every edge word, "
                   "injected disagreement (owner wins), syncFlag" % (gnx, ny, nz, world))
    return out


def run_b200(a):
    import torch
    import torch.distributed as dist
    import core_b200 as cb

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 arm has no CPU fallback")
    torch.cuda.set_device(local)
    bind_to_gpu_numa_node(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    n = a.n
    fp_mode = cb.FP_FAST if a.fp == "fast" else cb.FP_STRICT
    ops = cb.OP_LENGTHS | cb.OP_MARK_SPLIT | cb.OP_MARK_COLLAPSE | cb.OP_QUALITIES | cb.OP_MARK_BAD
    mark_mask = cb.SPLIT | cb.COLLAPSE | cb.NEED_NOT_SPLIT | cb.NEED_NOT_COLLAPSE
    stream = torch.cuda.Stream()
    T = Timed(torch, dist, world, stream)
    extras = {} if not a.no_extras else None

    pm = None
    if world > 1 and not a.no_extras:
        pm = parity_multi(cb, torch, dist, world, rank, local)

    def field_coords(xyz):
        # every slab is a unit cube; the field sees the triangle-wave coordinate u(x) = 1 - |1 - (x mod 2)|, a
        # continuous function of the GLOBAL position, so both copies of a shared vertex get bit-identical values
        if world == 1:
            return xyz
        f = xyz.copy()
        f[:, 0] = 1.0 - np.abs(1.0 - np.mod(xyz[:, 0], 2.0))
        return f

    def connect(p, links):
        uid = [cb.Part.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        p.comm_init(world, rank, uid[0])
        p.set_edge_links(links)

    # ---- this rank's part: x-slab `rank` of a (world*n) x n x n global box (weak scaling)
    t_gen = time.perf_counter()
    if a.workload == "mixed":
        part = cb.boxmesh.mixed_slab_part(world * n, n, n, a.layers, world, rank, wx=float(world))
        prism_v = part["prism_v"]
    else:
        part = (cb.boxmesh.slab_part(world * n, n, n, world, rank, wx=float(world)) if world > 1 else
                dict(zip(("xyz", "edge_v", "tet_v"), cb.boxmesh.kuhn_box(n, n, n)), edge_owned=None, links=[]))
        prism_v = None
    xyz0, edge_v, tet_v = part["xyz"], part["edge_v"], part["tet_v"]
    edge_owned, links = part["edge_owned"], part["links"]
    hbar = 1.0 / n
    xyz = cb.fields.jitter(xyz0, a.jitter * hbar) if a.jitter > 0 else xyz0
    if a.field == "iso":
        size = cb.fields.iso_linear(field_coords(xyz), hbar)
    else:
        h, R = cb.fields.shock_rotating(field_coords(xyz), hbar)
    nv, ne, nt = len(xyz), len(edge_v), len(tet_v)
    npr = 0 if prism_v is None else len(prism_v)
    t_gen = time.perf_counter() - t_gen

    p = cb.Part(local)
    p.set_stream(stream.cuda_stream)
    t_exp = time.perf_counter()
    p.set_mesh(xyz, edge_v, tet_v, prism_v=prism_v, edge_owned=edge_owned)
    p.synchronize()
    t_exp = time.perf_counter() - t_exp
    t_field = time.perf_counter()
    if a.field == "iso":
        p.set_size_field_iso(size)
    elif a.field == "aniso":
        p.set_size_field_aniso(h, R)
    else:
        p.set_size_field_logm_from_frames(h, R, 0)
    p.synchronize()
    t_field = time.perf_counter() - t_field     # upload + gather records + per-vertex transforms (cached across sweeps)
    if world > 1:
        connect(p, links)

    def make_step(part_obj, mode, has_layer):
        def step():
            part_obj.clear_flags()               # incoming flag words = 0 (async device memset)
            if has_layer:
                part_obj.reset_layer()           # ma::resetLayer on the device: LAYER closure (+ syncFlag per dimension) + freeze
            if world > 1:
                # the part-boundary exchange of the edge marks runs under the element sweep (mag_sweep_reconciled)
                part_obj.sweep(ops | (cb.OP_LAYER_CHECK if has_layer else 0), fp_mode=mode, reconcile_mask=mark_mask)
                return part_obj.allreduce_stats()
            part_obj.sweep(ops | (cb.OP_LAYER_CHECK if has_layer else 0), fp_mode=mode)
            return part_obj.stats()
        return step

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    main = T.run(p, make_step(p, fp_mode, npr > 0), a.steps, a.warmup)
    clocks = sampler.stop() if rank == 0 else None
    st = main["stats"]
    launches = main["launches"]
    if world > 1:
        lt = torch.tensor([launches], device="cuda", dtype=torch.int64)
        dist.all_reduce(lt, op=dist.ReduceOp.SUM)
        launches = int(lt.item())
        # per-part owned counts: parts of the same parity see the same field on the same mesh (see field_coords), so
        # interior parts of one parity must report identical counts -- a size-independent check of the 8-part run
        mine = p.stats()
        per = torch.tensor([mine["n_split"], mine["n_collapse"], mine["n_bad"]], device="cuda", dtype=torch.int64)
        allp = [torch.zeros_like(per) for _ in range(world)]
        dist.all_gather(allp, per)
        per_part = [[int(x) for x in t.tolist()] for t in allp]
    else:
        per_part = None
    # where the multi-part step spends what a single part does not (device events on the sweep stream; the statistics call
    # ends in a host synchronisation, so it is timed on the host from the moment the device is idle)
    multi = None
    if world > 1:
        acc = np.zeros(4)
        reps = 5
        for _ in range(reps):
            e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            T.barrier()
            with torch.cuda.stream(stream):
                p.clear_flags()
                e[0].record(stream)
                p.sweep(ops | (cb.OP_LAYER_CHECK if npr else 0), fp_mode=fp_mode)
                e[1].record(stream)
                p.reconcile_edge_flags(mark_mask)
                e[2].record(stream)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            p.allreduce_stats()
            acc += np.array([e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2]), 1e3 * (time.perf_counter() - t0), 0.0])
        acc /= reps
        multi = {"sweep_ms": T.max_over_ranks(float(acc[0])), "reconcile_flags_ms": T.max_over_ranks(float(acc[1])),
                 "allreduce_stats_ms": T.max_over_ranks(float(acc[2])),
                 "what": "the three pieces run one after the other, per step, max over ranks: the sweep kernels; pack + grouped "
                         "ncclSend/Recv + merge of the part-boundary flag words (device events); all-gather of the statistics + copy + "
                         "host sync (host clock, device idle at start).  The timed steps overlap the second with the element sweep "
                         "(mag_sweep_reconciled)"}
    ents_rank = ne + nt + npr
    ents_all = ents_rank * world
    step_ms = main["ms_per_step"]
    value = ents_all / (step_ms * 1e-3)

    # ---- end to end through the public API with HOST (pinned) buffers
    e2e = None
    if a.e2e_steps > 0 and a.workload == "box":
        h_xyz, h_ev, h_tv = _pin(torch, xyz), _pin(torch, edge_v), _pin(torch, tet_v)
        h_own = _pin(torch, edge_owned)
        kind_id = {"iso": 1, "aniso": 2, "logm": 3}[a.field]
        if a.field == "iso":
            h_m = (_pin(torch, size), None)
        elif a.field == "aniso":
            h_m = (_pin(torch, h), _pin(torch, R))
        else:
            # the log-Euclidean field is built on the host from sizes + frames (libm log, as the reference): once, outside
            # the timed steps -- the reference also builds it once per size-field construction, not per sweep
            h_m = (None, _pin(torch, p.set_size_field_logm_from_frames(h, R, 0, want_logm=True)))
        nbytes = lambda ts: sum(t.numel() * t.element_size() for t in ts if t is not None)

        def time_host(fn, steps):
            fn()
            T.barrier()
            t0 = time.perf_counter()
            for _ in range(steps):
                fn()
            T.barrier()
            return T.max_over_ranks(time.perf_counter() - t0) / steps

        # (a) the part stays resident (the sweeps of one MeshAdapt iteration, or a solver re-evaluating its size field on a
        # fixed mesh): coordinates + size field + one mark byte per entity up, mark bytes + statistics down
        h_em = torch.zeros(ne, dtype=torch.uint8).pin_memory()
        h_lm = torch.zeros(nt, dtype=torch.uint8).pin_memory()
        o_em = torch.empty(ne, dtype=torch.uint8).pin_memory()
        o_lm = torch.empty(nt, dtype=torch.uint8).pin_memory()

        def resweep_step():
            s = p.resweep_host(xyz=h_xyz, kind=kind_id, field_a=h_m[0], field_b=h_m[1], edge_marks=h_em, elem_marks=h_lm,
                               out_edge_marks=o_em, out_elem_marks=o_lm, ops=ops, fp_mode=fp_mode)
            if world > 1:
                p.reconcile_edge_flags(mark_mask)
                s = p.allreduce_stats()
                if s["n_flag_mismatch"]:          # never on consistent inputs: the owner's bits replaced a copy's
                    p.mark_bytes(o_em, None)
            return s

        dt = time_host(resweep_step, a.e2e_steps)
        h2d = nbytes((h_xyz, h_em, h_lm) + tuple(h_m))
        d2h = nbytes((o_em, o_lm)) + 88
        e2e = {"value": ents_all / dt, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "steps": a.e2e_steps, "ms_per_step": 1e3 * dt,
               "what": "mag_resweep_host per step and rank, connectivity resident on the device: vertex coordinates + size field + "
                       "one mark byte per entity up from pinned host buffers, sweep, mark bytes + statistics down "
                       "(uploads / kernels / downloads on three streams)"}
        # (a') the other sweeps of the SAME MeshAdapt iteration: mesh, coordinates and size field are unchanged, only the mark
        # bytes travel (one byte per entity up, one down).  Reported beside the headline e2e, which re-sends coordinates + field.
        if not a.no_extras:
            def marks_step():
                s = p.resweep_host(xyz=None, kind=-1, edge_marks=h_em, elem_marks=h_lm, out_edge_marks=o_em, out_elem_marks=o_lm,
                                   ops=ops, fp_mode=fp_mode)
                if world > 1:
                    p.reconcile_edge_flags(mark_mask)
                    s = p.allreduce_stats()
                return s
            dtm = time_host(marks_step, a.e2e_steps)
            extras["e2e_marks_only"] = {
                "value": ents_all / dtm, "unit": UNIT, "ms_per_step": 1e3 * dtm,
                "h2d_bytes_per_step": int(nbytes((h_em, h_lm))), "d2h_bytes_per_step": int(nbytes((o_em, o_lm)) + 88),
                "what": "mag_resweep_host with coordinates and size field unchanged (NULL): mark bytes up, sweep, mark bytes + "
                        "statistics down -- every sweep of a MeshAdapt iteration after its first"}
        # (b) full re-export every step (a MeshAdapt iteration after the mesh changed): the round-1 end-to-end leg
        if not a.no_extras:
            h_ef = torch.zeros(ne, dtype=torch.int32).pin_memory()
            h_lf = torch.zeros(nt, dtype=torch.int32).pin_memory()
            o_len = torch.empty(ne, dtype=torch.float64).pin_memory()
            o_q = torch.empty(nt, dtype=torch.float64).pin_memory()
            o_ef = torch.empty(ne, dtype=torch.int32).pin_memory()
            o_lf = torch.empty(nt, dtype=torch.int32).pin_memory()

            def export_step():
                s = p.sweep_host(h_xyz, h_ev, h_tv, kind_id, h_m[0], h_m[1], edge_flags=h_ef, elem_flags=h_lf,
                                 edge_owned=h_own, out_lengths=o_len, out_qualities=o_q, out_edge_flags=o_ef,
                                 out_elem_flags=o_lf, ops=ops, fp_mode=fp_mode)
                if world > 1:
                    p.reconcile_edge_flags(mark_mask)
                    s = p.allreduce_stats()
                    if s["n_flag_mismatch"]:
                        p.flags(o_ef, None)
                return s

            dt = time_host(export_step, max(1, min(3, a.e2e_steps)))
            extras["e2e_full_export"] = {
                "value": ents_all / dt, "unit": UNIT, "ms_per_step": 1e3 * dt,
                "h2d_bytes_per_step": int(nbytes((h_xyz, h_ev, h_tv, h_ef, h_lf, h_own) + tuple(h_m))),
                "d2h_bytes_per_step": int(nbytes((o_len, o_q, o_ef, o_lf)) + 88),
                "what": "mag_sweep_host: mesh + size field + flag words up, lengths + qualities + flag words + statistics down"}
            del h_ef, h_lf, o_len, o_q, o_ef, o_lf

    # ---- the other configurations BASELINE.json names, a few steps each, in the same line (rank 0 reports)
    if extras is not None and a.workload == "box" and a.field == "aniso":
        xs = max(3, min(5, a.steps))
        ab = lambda f: algorithmic_bytes(nv, ne, nt, f)
        peak, _ = _peak()

        def brief(r, f):
            return {"ms_per_step": r["ms_per_step"], "value": ents_all / (r["ms_per_step"] * 1e-3),
                    "kernel_ms_all": {"vertex_pass": r["vertex_ms"], "edges": r["edge_ms"], "elements": r["elem_ms"]},
                    "step_frac": ab(f)["total"] / (r["ms_per_step"] * 1e-3) / 1e9 / peak,
                    "edge_kernel_frac": ab(f)["edge_kernel"] / (r["edge_ms"] * 1e-3) / 1e9 / peak,
                    "stats": {k: r["stats"][k] for k in ("n_split", "n_collapse", "n_bad", "n_near_threshold")}}
        # the bit-exact arithmetic on the same part
        if a.fp == "fast":
            extras["strict"] = brief(T.run(p, make_step(p, cb.FP_STRICT, False), xs, 1), "aniso")
            # MAG_FP_FAST_LISTED: near-threshold edges are listed and decided by the fast value instead of being re-evaluated in
            # strict arithmetic -- the exception the parity rule itself allows; the lattice puts 8.4 M z edges ON the threshold
            extras["fast_listed"] = brief(T.run(p, make_step(p, cb.FP_FAST_LISTED, False), xs, 2), "aniso")
            extras["fast_listed"]["what"] = ("MAG_FP_FAST_LISTED: the same sweep with near-threshold edges listed, not re-evaluated; "
                                             "n_collapse / n_split may differ from the headline's by at most n_near_threshold")
        # SURVEY 8d's perturbed variant: no edge sits exactly on a threshold
        if a.jitter == 0:
            xj = cb.fields.jitter(xyz0, 0.2 * hbar)
            hj, Rj = cb.fields.shock_rotating(field_coords(xj), hbar)
            p.set_coords(xj)
            p.set_size_field_aniso(hj, Rj)
            extras["jitter_0.2"] = brief(T.run(p, make_step(p, fp_mode, False), xs, 2), "aniso")
            p.set_coords(xyz)
        # ma::configure's default: the log-Euclidean field built from the same sizes + frames (ma/maInput.h:177,198)
        t0 = time.perf_counter()
        p.set_size_field_logm_from_frames(h, R, 0)
        p.synchronize()
        t_logm = time.perf_counter() - t0
        extras["logm"] = brief(T.run(p, make_step(p, fp_mode, False), xs, 2), "logm")
        extras["logm"]["field_build_s"] = t_logm
        extras["logm"]["workload"] = workload_name(n, "logm")
        if a.fp == "fast":
            extras["logm_fast_listed"] = brief(T.run(p, make_step(p, cb.FP_FAST_LISTED, False), xs, 2), "logm")
        p.set_size_field_aniso(h, R)
    p.close()
    del p

    # ---- BASELINE configs[4]: mixed tet / prism boundary-layer box, one slab part per GPU
    if extras is not None and a.workload == "box" and a.field == "aniso":
        mn, mk = a.mixed_n, a.layers
        mp = cb.boxmesh.mixed_slab_part(world * mn, mn, mn, mk, world, rank, wx=float(world))
        hm, Rm = cb.fields.shock_rotating(field_coords(mp["xyz"]), 1.0 / mn)
        q = cb.Part(local)
        q.set_stream(stream.cuda_stream)
        q.set_mesh(mp["xyz"], mp["edge_v"], mp["tet_v"], prism_v=mp["prism_v"], edge_owned=mp["edge_owned"])
        q.set_size_field_aniso(hm, Rm)
        if world > 1:
            connect(q, mp["links"])
        r = T.run(q, make_step(q, fp_mode, True), max(3, min(10, a.steps)), 2)
        m_ents = (len(mp["edge_v"]) + len(mp["tet_v"]) + len(mp["prism_v"])) * world
        ms = r["stats"]
        extras["mixed"] = {
            "workload": "mixed boundary-layer box %d^3 cells per part, bottom %d layers prisms (%d prisms + %d tets, %d edges per part), "
                        "rotating shock-layer AnisoSizeField; per step: flags cleared, ma::resetLayer on the device (LAYER closure + "
                        "freeze), edge lengths + marks, tet mean-ratio + isPrismOk per topology, statistics"
                        % (mn, mk, len(mp["prism_v"]), len(mp["tet_v"]), len(mp["edge_v"])),
            "ms_per_step": r["ms_per_step"], "value": m_ents / (r["ms_per_step"] * 1e-3), "unit": UNIT,
            "kernel_ms_all": {"vertex_pass": r["vertex_ms"], "edges": r["edge_ms"], "elements": r["elem_ms"]},
            "stats": {k: ms[k] for k in ("n_split", "n_collapse", "n_bad", "n_near_threshold", "n_flag_mismatch", "n_layer_unsafe")}}
        q.close()

    if rank == 0:
        roofline = _roofline(main, nv, ne, nt, a.field, a.fp, n, world, a.jitter)
        cpu = None
        if not a.no_cpu:
            cpu = cpu_baseline(a.cpu_n or 55)   # BASELINE configs[1] size: 998,250 tets, ~11 s on one core
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
                "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": (workload_name(n, a.field) if a.workload == "box" else
                                        "mixed boundary-layer box n=%d per part, %d prism layers (%d prisms, %d tets, %d edges)"
                                        % (n, a.layers, npr, nt, ne)),
                           "fp_mode": a.fp,
                           "ops": "lengths+mark_split+mark_collapse+qualities+mark_bad" + ("+layer_check+reset_layer" if npr else ""),
                           "parts": world, "partition": "x-slabs, one part per GPU" if world > 1 else "1 part",
                           "l2": "inputs_exceed_l2 (>=4 GB per step vs 126 MB L2), no flush needed",
                           "jitter": a.jitter, "mesh_generation_s": t_gen, "export_s": t_exp,
                           "field_upload_ms": 1e3 * t_field},
                "stats": {k: st[k] for k in ("n_split", "n_collapse", "n_bad", "n_near_threshold", "n_flag_mismatch",
                                             "min_quality", "max_length")},
                "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks}
        if per_part is not None:
            even = [c for i, c in enumerate(per_part) if i % 2 == 0 and 0 < i]
            odd = [c for i, c in enumerate(per_part) if i % 2 == 1]
            line["per_part_owned_counts"] = per_part
            line["parity_symmetry"] = "ok" if all(c == even[0] for c in even) and all(c == odd[0] for c in odd) else "FAILED"
        if multi is not None:
            line["multi_part_overhead"] = multi
        if pm is not None:
            line["parity_multi"] = pm["status"]
            line["parity_multi_detail"] = pm
        if extras:
            line["extra"] = extras
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    a = parse()
    if a.impl == "reference":
        run_reference_arm(a)
    else:
        run_b200(a)


if __name__ == "__main__":
    main()
