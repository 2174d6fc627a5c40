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
            "config": {"workload": workload_name(a.n, a.field), "cpu_sample": sample},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": P, "kind": kind, "sample": sample},
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

    # ---- this rank's part: x-slab `rank` of a (world*n) x n x n global box (weak scaling)
    t_gen = time.perf_counter()
    if world == 1:
        xyz, edge_v, tet_v = cb.boxmesh.kuhn_box(n, n, n)
        edge_owned, links = None, []
    else:
        part = cb.boxmesh.slab_part(world * n, n, n, world, rank, wx=float(world))
        xyz, edge_v, tet_v = part["xyz"], part["edge_v"], part["tet_v"]
        edge_owned, links = part["edge_owned"], part["links"]
    hbar = 1.0 / n
    if a.jitter > 0:
        xyz = cb.fields.jitter(xyz, a.jitter * hbar)
    # every slab is a unit cube; the field sees the triangle-wave coordinate u(x) = 1 - |1 - (x mod 2)|, a
    # continuous function of the GLOBAL position, so both copies of a shared vertex get bit-identical values
    fxyz = xyz
    if world > 1:
        fxyz = xyz.copy()
        fxyz[:, 0] = 1.0 - np.abs(1.0 - np.mod(xyz[:, 0], 2.0))
    if a.field == "iso":
        size = cb.fields.iso_linear(fxyz, hbar)
    else:
        h, R = cb.fields.shock_rotating(fxyz, hbar)
    nv, ne, nt = len(xyz), len(edge_v), len(tet_v)
    t_gen = time.perf_counter() - t_gen

    stream = torch.cuda.Stream()
    p = cb.Part(local)
    p.set_stream(stream.cuda_stream)
    p.set_mesh(xyz, edge_v, tet_v, edge_owned=edge_owned)
    if a.field == "iso":
        p.set_size_field_iso(size)
    elif a.field == "aniso":
        p.set_size_field_aniso(h, R)
    else:
        p.set_size_field_logm_from_frames(h, R, 0)
    if world > 1:
        uid = [cb.Part.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        p.comm_init(world, rank, uid[0])
        p.set_edge_links(links)
    mark_mask = cb.SPLIT | cb.COLLAPSE | cb.NEED_NOT_SPLIT | cb.NEED_NOT_COLLAPSE

    def step():
        p.clear_flags()                      # incoming flag words = 0 (async device memset)
        p.sweep(ops, fp_mode=fp_mode)
        if world > 1:
            p.reconcile_edge_flags(mark_mask)
            return p.allreduce_stats()
        return p.stats()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(a.warmup):
        st = step()
    launches0 = p.launch_count()
    p.timing_begin(a.steps)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    with torch.cuda.stream(stream):
        ev0.record(stream)
        for _ in range(a.steps):
            st = step()
        ev1.record(stream)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms_total = ev0.elapsed_time(ev1)
    kt = p.timing_read()                     # [steps, 3] vertex / edge / elem ms
    launches = p.launch_count() - launches0
    if world > 1:
        t = torch.tensor([ms_total], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
        lt = torch.tensor([launches], device="cuda", dtype=torch.int64)
        dist.all_reduce(lt, op=dist.ReduceOp.SUM)
        launches = int(lt.item())
    ents_rank = ne + nt
    ents_all = ents_rank * world
    value = ents_all * a.steps / (ms_total * 1e-3)

    # ---- end to end through the public API with HOST (pinned) buffers: full re-export every step
    e2e = None
    if a.e2e_steps > 0:
        def pin(x):
            t = torch.from_numpy(np.ascontiguousarray(x)).pin_memory()
            return t
        h_xyz, h_ev, h_tv = pin(xyz), pin(edge_v), pin(tet_v)
        h_own = pin(edge_owned) if edge_owned is not None else None
        h_ef = torch.zeros(ne, dtype=torch.int32).pin_memory()
        h_lf = torch.zeros(nt, dtype=torch.int32).pin_memory()
        if a.field == "iso":
            h_m = (pin(size),)
        else:
            h_m = (pin(h), pin(R))
        o_len = torch.empty(ne, dtype=torch.float64).pin_memory()
        o_q = torch.empty(nt, dtype=torch.float64).pin_memory()
        o_ef = torch.empty(ne, dtype=torch.int32).pin_memory()
        o_lf = torch.empty(nt, dtype=torch.int32).pin_memory()

        kind_id = {"iso": 1, "aniso": 2, "logm": 3}[a.field]
        if a.field == "logm":
            # the log-Euclidean field is built on the host from sizes + frames (libm log, as the reference): once, outside
            # the timed steps -- the reference also builds it once per size-field construction, not per sweep
            h_m = (None, pin(p.set_size_field_logm_from_frames(h, R, 0, want_logm=True)))
        elif a.field == "iso":
            h_m = (h_m[0], None)

        h2d = sum(t.numel() * t.element_size() for t in (h_xyz, h_ev, h_tv, h_ef, h_lf, h_own) + tuple(h_m) if t is not None)
        d2h = sum(t.numel() * t.element_size() for t in (o_len, o_q, o_ef, o_lf)) + 88

        def e2e_step():
            # one streamed call: export (mesh + size field + flag words), sweep, lengths / qualities / flags / statistics
            st = p.sweep_host(h_xyz, h_ev, h_tv, kind_id, h_m[0], h_m[1], edge_flags=h_ef, elem_flags=h_lf,
                              edge_owned=h_own, out_lengths=o_len, out_qualities=o_q, out_edge_flags=o_ef,
                              out_elem_flags=o_lf, ops=ops, fp_mode=fp_mode)
            if world > 1:
                p.reconcile_edge_flags(mark_mask)
                st = p.allreduce_stats()
                if st["n_flag_mismatch"]:      # never on consistent inputs: the owner's bits replaced a copy's
                    p.flags(o_ef, None)
            return st

        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(a.e2e_steps):
            e2e_step()
        barrier()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e = {"value": ents_all * a.e2e_steps / dt, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "steps": a.e2e_steps, "ms_per_step": 1e3 * dt / a.e2e_steps,
               "what": "mag_sweep_host per step and rank: mesh + size field + flag words up from pinned host buffers, sweep, "
                       "lengths + qualities + flags + statistics down, uploads / kernels / downloads streamed in slices"}

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(HERE, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        ab = algorithmic_bytes(nv, ne, nt, a.field)
        edge_ms, elem_ms, vert_ms = float(kt[:, 1].mean()), float(kt[:, 2].mean()), float(kt[:, 0].mean())
        dom = "k_edges" if edge_ms >= elem_ms else "k_tets"
        dom_ms = max(edge_ms, elem_ms)
        dom_bytes = ab["edge_kernel"] if dom == "k_edges" else ab["elem_kernel"]
        achieved = dom_bytes / (dom_ms * 1e-3) / 1e9
        step_ms = ms_total / a.steps
        # DRAM bytes per launch of that kernel from the committed ncu --set full capture (profiles/traffic.json,
        # written by scripts/profile_summary.py); only meaningful for the configuration that was profiled
        traffic, units = None, None
        try:
            if n == 203 and a.field == "aniso" and world == 1 and a.jitter == 0:
                tj = json.load(open(os.path.join(HERE, "profiles", "traffic.json")))
                # template arguments as ncu prints them: k_edges<KIND, FAST, VERT>, k_tets<KIND, FAST, USE_MAX>
                key = "%s<2, %d, %d>" % (dom, 1 if a.fp == "fast" else 0, 0 if dom == "k_edges" else 1)
                traffic = float(tj[key]["dram_bytes_per_launch"])
                # the path is not purely HBM-bound: utilisation of the other units under ncu
                units = {k: tj[key].get(k) for k in ("fp64_pipe_pct", "l1_data_pipe_pct", "dram_pct", "warps_active_pct")}
                units["source"] = tj[key].get("source")
        except Exception:
            traffic, units = None, None
        roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": int(dom_bytes), "kernel_ms": dom_ms,
                    "kernel_ms_all": {"vertex_pass": vert_ms, "edges": edge_ms, "elements": elem_ms},
                    "step_algorithmic_bytes": int(ab["total"]),
                    "step_frac": ab["total"] / (step_ms * 1e-3) / 1e9 / peak,
                    "unit_utilisation_ncu": units}
        cpu = None
        if not a.no_cpu:
            cpu = cpu_baseline(a.cpu_n or 55)   # BASELINE configs[1] size: 998,250 tets, ~11 s on one core
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
                "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload_name(n, a.field), "fp_mode": a.fp,
                           "ops": "lengths+mark_split+mark_collapse+qualities+mark_bad", "parts": world,
                           "partition": "x-slabs, one part per GPU" if world > 1 else "1 part",
                           "l2": "inputs_exceed_l2 (>=4 GB per step vs 126 MB L2), no flush needed",
                           "mesh_generation_s": t_gen},
                "stats": {k: st[k] for k in ("n_split", "n_collapse", "n_bad", "n_near_threshold", "n_flag_mismatch",
                                             "min_quality", "max_length")},
                "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks}
        print(json.dumps(line), flush=True)
    p.close()
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
