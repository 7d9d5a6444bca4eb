#!/usr/bin/env python
"""bench.py -- K-assembly throughput of the B200 path on BASELINE.json's config 1.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config tetra4|hexa8|...] [--size S]

A "step" is one full stiffness assembly of the workload mesh: NewGeneralAssembler + AddIsoparametric
(symbolic phase + numeric phase + deterministic row reduction), exactly what BenchmarkTetra4Assembly times
(benchmark_test.go:39-47).  Default workload: Tetra4 BCC cube, N=48: 1,299,456 elements / 684,723 dofs,
solids.Isotropic{E:200e9, Poisson:0.3}.Solid3D().

  value : elements/s with inputs (nodes, int32 connectivity) already resident in HBM, timed with CUDA events
          on the library's stream, L2 flushed between steps.
  e2e   : the same through the reference-facing host API with HOST buffers (pinned): H2D of nodes and int64
          connectivity, assembly, D2H of the full CSR K into pinned host arrays, all inside the timed region.
  roofline     : dominant kernel (k_noderows: B^T C B products fused with the deterministic row reduction) against the
                 measured HBM copy bandwidth of MEASURED_PEAKS.json; FP64 / copy peaks probed in the same process.
  reassemble_cached_pattern : numeric-only re-assembly on the cached symbolic plan (femb200_reassemble).
  cpu_baseline : the CPU oracle (C port of the reference algorithm, single thread like the reference) timed on
                 this box's host cores on the same workload.
  --impl reference : the oracle port timed alone (the Go reference cannot be built: no Go toolchain).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

WORKLOADS = {
    # name: (mesh kind, default size, element class, constituter, model dofs flag, oracle kind, oracle mode, description)
    "tetra4": ("tetra4", 48, "Tetra4", ("Isotropic", "Solid3D"), 7, 0, 0, "Tetra4 BCC cube N={s}: {ne} elements / {nd} dofs, Isotropic(E=200e9, nu=0.3).Solid3D"),
    "quad4": ("quad4", 2000, "Quad4", ("Isotropic", "Axisymmetric"), 3, 4, 1, "Quad4 axisymmetric {s}x{s}: {ne} elements / {nd} dofs"),
    "hexa8": ("hexa8", 100, "Hexa8", ("Isotropic", "Solid3D"), 7, 2, 0, "Hexa8 brick {s}^3: {ne} elements / {nd} dofs, 2x2x2 Gauss"),
    "tetra10": ("tetra10", 40, "Tetra10", ("Isotropic", "Solid3D"), 7, 1, 0, "Tetra10 BCC N={s}: {ne} elements / {nd} dofs, 4-point rule"),
    "hexa20": ("hexa20", 40, "Hexa20", ("Isotropic", "Solid3D"), 7, 3, 0, "Hexa20 brick {s}^3: {ne} elements / {nd} dofs, 3x3x3 Gauss"),
}
E_MOD, NU = 200e9, 0.3


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """SM clock / throttle-reason sampling DURING the timed region (B200_PROFILING.md clocks line).  The timed
    region is tens of milliseconds, so NVML is polled in-process (~1 ms per sample); nvidia-smi -lms is the
    fallback when the NVML binding is unavailable."""

    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu, self.sm, self.mx, self.reasons, self.stop_flag = gpu_index, [], [], set(), False
        self.proc = None
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            idx = gpu_index
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            if vis:
                try:
                    idx = int(vis.split(",")[gpu_index])
                except Exception:
                    idx = gpu_index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.nvml = pynvml
            self.mx.append(float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)))
        except Exception:
            self.nvml = None

    def _sample_nvml(self):
        n = self.nvml
        self.sm.append(float(n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM)))
        mask = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.h)) if hasattr(n, "nvmlDeviceGetCurrentClocksEventReasons") \
            else int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
        for name, bit in self.REASONS:
            if mask & bit:
                self.reasons.add(name)

    def run(self):
        if self.nvml is not None:
            while not self.stop_flag:
                try:
                    self._sample_nvml()
                except Exception:
                    break
                time.sleep(0.001)
            return
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                r = [x.strip() for x in line.split(",")]
                if len(r) >= 6 and r[0].replace(".", "").isdigit():
                    self.sm.append(float(r[0]))
                    self.mx.append(float(r[1]))
                    for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2:6]):
                        if v.lower().startswith("active"):
                            self.reasons.add(name)
                if self.stop_flag:
                    break
        except Exception:
            pass

    def finish(self):
        self.stop_flag = True
        if self.proc:
            try:
                self.proc.terminate()
            except Exception:
                pass
        self.join(timeout=2.0)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": max(self.mx) if self.mx else None,
                "reasons": sorted(self.reasons), "samples": len(self.sm),
                "source": "nvml polled in-process during the timed region" if self.nvml is not None else "nvidia-smi -lms 20"}


def make_workload(name, size):
    pkg = entry.load_package()
    from importlib import import_module
    meshes = import_module("gofem_b200.meshes")
    kind, dsize, ecls, (mcls, mmeth), flags, okind, omode, desc = WORKLOADS[name]
    s = size or dsize
    nodes, conn = meshes.mesh_for(kind, s)
    elemT = getattr(pkg.elements, ecls)()
    c = getattr(getattr(pkg.solids, mcls)(E_MOD, NU), mmeth)()
    ndn = bin(flags).count("1")
    return dict(pkg=pkg, nodes=nodes, conn=conn, elemT=elemT, c=c, flags=flags, ndn=ndn, okind=okind, omode=omode,
                desc=desc.format(s=s, ne=conn.shape[0], nd=nodes.shape[0] * ndn), size=s, name=name)


def algorithmic_bytes(w, nnz):
    """Per-launch algorithmic bytes of the dominant kernel k_rows: staged geometry records
    (one per element and quadrature point) + incidence list (n2e_t) + element->slot map (pos16) read, CSR values
    written.  Record = 4 doubles per node when a node carries 3 values (f rides in the 4th), else nn*NV + f padded even."""
    ne, nn = w["conn"].shape
    nv = {"tetra4": 3, "hexa8": 3, "tetra10": 3, "hexa20": 3, "quad4": 3}[w["name"]]
    nqp = len(w["elemT"].Quadrature()[1])
    rec = 4 * nn if nv == 3 else ((nn * nv + 2) & ~1)
    return 8 * rec * nqp * ne + 4 * nn * ne + 2 * nn * nn * ne + 8 * nnz


def compulsory_bytes(w, nnz, ndof):
    """SURVEY section 8(d) B_alg of one full call: int32 connectivity + node coordinates in, CSR (indptr int64 + indices int32 +
    values f64) out."""
    ne, nn = w["conn"].shape
    return 4 * nn * ne + 24 * w["nodes"].shape[0] + 8 * (ndof + 1) + 12 * nnz


def cpu_port_seconds(w, nodes, conn):
    O = entry.load_oracle()
    O.build()
    O.set_normaliser(False)          # the reference computes no test normaliser
    Cm, bk = O.constitutive(w["omode"], E_MOD, NU)
    og = O.GeneralAssembler(nodes, w["flags"])
    t0 = time.perf_counter()
    og.add_isoparametric(w["okind"], Cm, bk, conn)
    dt = time.perf_counter() - t0
    O.set_normaliser(True)
    return dt


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (oracle port; the Go original cannot be built here) timed
    alone on the host.  Single thread: the reference is one sequential loop (fem.go:140)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    name = args.config
    if args.size:
        size = args.size
    elif name == "tetra4":
        # bounded sample: about 90 s of CPU work for the whole run at ~1.1e5 elements/s (single thread), never more than the
        # workload itself; BCC cube N -> 12 N^2 (N-1) elements
        budget = 90.0 / max(1, args.steps + args.warmup) * 1.1e5
        size = 8
        while size < 48 and 12 * (size + 1) ** 2 * size <= budget:
            size += 1
    else:
        size = {"quad4": 700, "hexa8": 40, "tetra10": 16, "hexa20": 12}[name]
    w = make_workload(name, size)
    full = make_workload_desc(name, args.size)
    times = []
    for i in range(args.warmup + args.steps):
        dt = cpu_port_seconds(w, w["nodes"], w["conn"])
        if i >= args.warmup:
            times.append(dt)
    ne = w["conn"].shape[0]
    v = ne * len(times) / sum(times)
    go_probe = subprocess.run("go version", shell=True, capture_output=True, text=True)
    line = {"impl": "reference", "metric": "K assembly elements/s", "value": v, "unit": "elements/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": full, "sample": w["desc"]},
            "cpu_baseline": {"value": v, "unit": "elements/s", "cores": 1, "kind": "port",
                             "sample": f"{w['desc']} ({ne} elements per step); C port of the reference algorithm incl. full COO + sort-accumulate; go toolchain: "
                                       + ("present but modules unavailable offline" if go_probe.returncode == 0 else "absent")},
            "dofs_per_s": v * w["nodes"].shape[0] * w["ndn"] / ne,
            "e2e": {"value": v, "unit": "elements/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def make_workload_desc(name, size):
    kind, dsize, *_rest, desc = WORKLOADS[name]
    s = size or dsize
    if name == "tetra4":
        return desc.format(s=s, ne=12 * s * s * (s - 1), nd=3 * ((s + 1) ** 3 + s ** 3))
    return desc.format(s=s, ne="?", nd="?")


def run_multi_gpu(args, rank, world, local_rank):
    """N > 1: weak scaling.  Global mesh = BCC box of (48*N) x 48 x 48 cells (N x config 1); elements in cell-major
    (space-filling) order are split into N contiguous chunks, one per GPU; a node (row block) belongs to the lowest
    rank touching it.  A step = local assembly of the chunk (symbolic + numeric, inputs resident in HBM) + the
    exchange that leaves every GPU owning its row block of the distributed CSR K:
      exchange (headline, the north star's scheme): owned rows assembled from [own elements ; halo elements (pattern
        only)], interface rows assembled from the own elements touching foreign nodes, packed per owner, ONE NCCL
        all-to-all over NVLink (sizes fixed by the partition plan), added in place in ascending source rank;
      owner_computes (reported alongside): halo elements are assembled numerically by the owner as well -- redundant
        O(surface) work, no collective at all, rows bit-identical to the single-GPU result.
    Timed with CUDA events on the single stream everything runs on, max over ranks."""
    import ctypes as C
    import torch
    import torch.distributed as dist
    from importlib import import_module
    pkg = entry.load_package()
    D = import_module("gofem_b200.distributed")
    M = import_module("gofem_b200.meshes")
    dev = torch.device("cuda", local_rank)
    N = args.size or 48
    nodes, conn = M.bcc_tetra4_box(N * world, N, N, 1.0 / N)
    ne_total = conn.shape[0]
    bounds = D.chunk_bounds(ne_total, world)
    plan = D.ExchangePlan(nodes.shape[0], conn, bounds, rank, 3)
    iplan = D.InterfacePlan(plan, nodes.shape[0], conn, bounds, 3, dist)
    # rank-local node numbering (own + ghost nodes, ascending global id): every device array of a step is sized by
    # the rank's share of the mesh; K rows are local, K column indices global
    gids = np.sort(np.concatenate([np.arange(bounds[rank], bounds[rank + 1]), plan.halo_elems]))
    ln = D.LocalNumbering(nodes, conn[gids])
    lconn, nhalo = plan.local_connectivity(conn, bounds)            # [own ; halo] for the row-exchange variant
    my = ln.conn(lconn)
    own_conn = conn[bounds[rank]:bounds[rank + 1]]
    out_conn = ln.conn(own_conn[plan.send_elems]) if plan.send_elems.size else np.zeros((0, 4), dtype=np.int32)
    ghost = ln.conn(conn[gids])                                     # own + halo in ascending GLOBAL element index
    n_local = ln.nodes.shape[0]
    d_nodes = torch.from_numpy(ln.nodes).to(dev)
    d_l2g = torch.from_numpy(ln.l2g.astype(np.int32)).to(dev)
    d_conn = torch.from_numpy(my).to(dev)
    d_ghost = torch.from_numpy(ghost).to(dev)
    d_out = torch.from_numpy(out_conn).to(dev) if out_conn.size else torch.zeros(4, dtype=torch.int32, device=dev)
    m_own = torch.as_tensor(ln.node_mask(plan.own_node_mask), device=dev)
    m_send = torch.as_tensor(ln.node_mask(plan.send_node_mask), device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    elemT, c = pkg.elements.Tetra4(), pkg.solids.Isotropic(E_MOD, NU).Solid3D()
    # two library contexts per rank, one stream and one host thread each: the owned row block and the (small) set of
    # outgoing interface rows are assembled concurrently, the second call filling the gaps of the first
    from concurrent.futures import ThreadPoolExecutor
    holder = pkg.RawAssembler(None, 3, device=local_rank, nodes_device_ptr=d_nodes.data_ptr(), n_nodes=n_local)
    holder_b = pkg.RawAssembler(None, 3, device=local_rank, nodes_device_ptr=d_nodes.data_ptr(), n_nodes=n_local)
    L = pkg.lib()
    s0 = torch.cuda.current_stream()
    sa, sb = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    L.femb200_ctx_set_stream(holder.ctx, C.c_void_p(sa.cuda_stream))
    L.femb200_ctx_set_stream(holder_b.ctx, C.c_void_p(sb.cuda_stream))
    L.femb200_ctx_set_concurrent(holder.ctx, 1)
    L.femb200_ctx_set_concurrent(holder_b.ctx, 1)
    pool = ThreadPoolExecutor(max_workers=1)
    desc = holder.make_desc(elemT, c)
    px = D.PlannedExchange(iplan, dev, local=ln)

    class View:
        def __init__(self, h):
            self.h = h

        def csr_device(self):
            return tuple(v.value for v in _csr_dev(L, self.h))

        def total_dofs(self):
            return L.femb200_total_dofs(self.h)

        def nnz(self):
            return L.femb200_nnz(self.h)

    def new_assembler(mask, ctx=None):
        h = C.c_void_p()
        assert L.femb200_assembler_create(ctx or holder.ctx, C.c_void_p(d_nodes.data_ptr()), n_local, 3, 1, C.byref(h)) == 0
        L.femb200_assembler_set_node_mask_device(h, C.c_void_p(mask.data_ptr()))
        assert L.femb200_assembler_set_column_map_device(h, C.c_void_p(d_l2g.data_ptr()), nodes.shape[0]) == 0
        return h

    def assemble_outgoing():
        # interface rows to send: own elements touching foreign nodes, rows of foreign nodes only
        torch.cuda.set_device(local_rank)
        hb = new_assembler(m_send, holder_b.ctx)
        rc = L.femb200_add_isoparametric(hb, C.byref(desc), C.c_void_p(d_out.data_ptr()), 0, 1, out_conn.shape[0])
        assert rc == 0, rc
        return hb

    sub_t = {}

    def step_exchange_diag():
        """FEMB200_BENCH_SUBSTEPS=1: the same step serialised, host clock per sub-step (diagnostics only)."""
        def mark(name, t0):
            torch.cuda.synchronize()
            sub_t[name] = sub_t.get(name, 0.0) + 1e3 * (time.perf_counter() - t0)
            return time.perf_counter()
        t0 = time.perf_counter()
        ha = new_assembler(m_own)
        rc = L.femb200_add_isoparametric_halo(ha, C.byref(desc), C.c_void_p(d_conn.data_ptr()), 0, 1, my.shape[0], nhalo)
        t0 = mark("owned_block", t0)
        hb = assemble_outgoing()
        t0 = mark("outgoing_rows", t0)
        be_a, be_b = D.Femb200Backend(pkg, View(ha), dev), D.Femb200Backend(pkg, View(hb), dev)
        tt = [t0]

        def ap():
            s0.wait_stream(sb)
            tt[0] = mark("pack", tt[0])

        def ac():
            tt[0] = mark("all_to_all", tt[0])
            sa.wait_stream(s0)
        sent = px.run(be_a, be_b, dist, after_pack=ap, after_collective=ac)
        s0.wait_stream(sa)
        t0 = mark("add", tt[0])
        nnz = L.femb200_nnz(ha)
        L.femb200_assembler_destroy(ha)
        L.femb200_assembler_destroy(hb)
        mark("destroy", t0)
        return sent, nnz

    def step_exchange():
        if os.environ.get("FEMB200_BENCH_SUBSTEPS"):
            return step_exchange_diag()
        fut = pool.submit(assemble_outgoing)
        # owned row block: own elements + pattern-only halo, rows of owned nodes only
        ha = new_assembler(m_own)
        rc = L.femb200_add_isoparametric_halo(ha, C.byref(desc), C.c_void_p(d_conn.data_ptr()), 0, 1, my.shape[0], nhalo)
        assert rc == 0, rc
        hb = fut.result()
        # both calls have synchronised their streams; pack on sb -> all-to-all on s0 -> adds on sa -> s0 waits for sa
        sent = px.run(D.Femb200Backend(pkg, View(ha), dev), D.Femb200Backend(pkg, View(hb), dev), dist,
                      after_pack=lambda: s0.wait_stream(sb), after_collective=lambda: sa.wait_stream(s0))
        s0.wait_stream(sa)
        nnz = L.femb200_nnz(ha)
        L.femb200_assembler_destroy(ha)      # synchronises its stream: the exchange is complete when the step returns
        L.femb200_assembler_destroy(hb)
        return sent, nnz

    def step_owner_computes():
        ha = new_assembler(m_own)
        rc = L.femb200_add_isoparametric(ha, C.byref(desc), C.c_void_p(d_ghost.data_ptr()), 0, 1, ghost.shape[0])
        assert rc == 0, rc
        nnz = L.femb200_nnz(ha)
        L.femb200_assembler_destroy(ha)
        return 0, nnz

    warmup = max(args.warmup, 3)

    def timed(step):
        for _ in range(warmup):
            step()
            flush.fill_(1)
        torch.cuda.synchronize()
        dist.barrier()
        sampler = ClockSampler(local_rank)
        sampler.start()
        launches0 = L.femb200_kernel_launches(holder.ctx)
        times = []
        sent = nnz = 0
        sub_t.clear()
        for _ in range(args.steps):
            flush.fill_(1)
            torch.cuda.synchronize()
            dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            sent, nnz = step()
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
        launches = L.femb200_kernel_launches(holder.ctx) - launches0
        clocks = sampler.finish()
        t = torch.tensor([float(np.mean(times))], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        tot = torch.tensor([float(sent), float(nnz)], dtype=torch.float64, device=dev)
        dist.all_reduce(tot)
        return float(t.item()), tot[0].item(), tot[1].item(), int(launches), clocks

    ms_x, sent_b, nnz_x, launches_x, _ = timed(step_exchange)
    sub_x = dict(sub_t)
    mism = torch.tensor([float(px.mismatches())], dtype=torch.float64, device=dev)
    dist.all_reduce(mism)
    ms, _, nnz_t, launches, clocks = timed(step_owner_computes)
    halo_tot = torch.tensor([float(nhalo)], dtype=torch.float64, device=dev)
    dist.all_reduce(halo_tot)
    if rank == 0:
        value = ne_total / (ms * 1e-3)
        line = {"metric": "K assembly elements/s", "value": value, "unit": "elements/s", "n_gpus": world, "steps": args.steps, "warmup": warmup,
                "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": f"Tetra4 BCC box {N * world}x{N}x{N}: {ne_total} elements / {nodes.shape[0] * 3} dofs ({world} x config 1), Isotropic(E=200e9, nu=0.3).Solid3D",
                           "l2": "flushed between timed steps (256 MiB write)", "inputs": "nodes f64 + int32 connectivity chunk resident in HBM",
                           "multi_gpu": "contiguous cell-major element chunks, one per GPU; a node (row block) belongs to the lowest rank touching it; "
                                        "owner computes: every rank assembles its own elements plus the halo elements touching its rows "
                                        "(redundant O(surface) work, ascending global element order => rows bit-identical to 1 GPU), no data-path collective; "
                                        "the NCCL row exchange of partial interface rows is timed alongside (key 'row_exchange')"},
                "dofs_per_s": value * nodes.shape[0] * 3 / ne_total, "nnz_total": nnz_t, "redundant_halo_elements_total": halo_tot.item(),
                "row_exchange": {"value": ne_total / (ms_x * 1e-3), "unit": "elements/s", "ms_per_step": ms_x, "gpu_launches": launches_x,
                                 "interface_bytes_sent_total": sent_b, "nnz_total": nnz_x, "plan_mismatches": mism.item(),
                                 "note": "north-star scheme: owned rows from [own ; halo (pattern only)], interface rows from the own elements touching foreign "
                                         "nodes (second context, concurrent), ONE NCCL all-to-all with plan-time sizes, in-place add in ascending source rank"},
                "e2e": {"value": value, "unit": "elements/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                        "note": "multi-GPU result stays distributed on the devices (row blocks); host-buffer e2e is reported at N=1"},
                "gpu_launches": launches, "clocks": clocks, "roofline": None, "cpu_baseline": None}
        if sub_x:
            line["row_exchange"]["substeps_ms_rank0_serialised"] = {k: v / args.steps for k, v in sub_x.items()}
        print(json.dumps(line))
    dist.barrier()
    dist.destroy_process_group()


def _csr_dev(L, h):
    import ctypes as C
    a, b, c = C.c_void_p(), C.c_void_p(), C.c_void_p()
    L.femb200_csr_device(h, C.byref(a), C.byref(b), C.byref(c))
    return a, b, c


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--config", default="tetra4", choices=sorted(WORKLOADS))
    ap.add_argument("--size", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-reassemble", action="store_true", help="skip the numeric-only re-assembly timing (ncu launch lists)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"   # the version banner goes to stdout; the contract is ONE JSON line
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the assembly path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    warmup = max(args.warmup, 3)
    if world > 1:
        return run_multi_gpu(args, rank, world, local_rank)

    w = make_workload(args.config, args.size)
    pkg = w["pkg"]
    nodes, conn = w["nodes"], w["conn"]
    ne, nn = conn.shape
    ndof = nodes.shape[0] * w["ndn"]

    # ---- device-resident inputs (value) -----------------------------------------------------------
    d_nodes = torch.from_numpy(nodes).to(dev)
    d_conn = torch.from_numpy(conn.astype(np.int32)).to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2
    L = pkg.lib()

    def one_resident_step():
        ra = pkg.RawAssembler(None, w["ndn"], device=local_rank, nodes_device_ptr=d_nodes.data_ptr(), n_nodes=nodes.shape[0])
        desc = ra.make_desc(w["elemT"], w["c"])
        rc = ra.add(desc, conn_ptr=d_conn.data_ptr(), n_elem=ne, is64=False, on_device=True)
        if rc:
            raise RuntimeError(f"femb200_add_isoparametric -> {rc} {ra.last_error()}")
        return ra

    # one context reused across steps (the memory pool keeps blocks, like a long-lived Go process)
    ctx_holder = pkg.RawAssembler(None, w["ndn"], device=local_rank, nodes_device_ptr=d_nodes.data_ptr(), n_nodes=nodes.shape[0])
    import ctypes as C

    def resident_step_shared_ctx():
        h = C.c_void_p()
        rc = L.femb200_assembler_create(ctx_holder.ctx, C.c_void_p(d_nodes.data_ptr()), nodes.shape[0], w["ndn"], 1, C.byref(h))
        assert rc == 0
        rc = L.femb200_add_isoparametric(h, C.byref(desc0), C.c_void_p(d_conn.data_ptr()), 0, 1, ne)
        if rc:
            raise RuntimeError(f"femb200_add_isoparametric -> {rc}")
        ph = (C.c_float * pkg.N_TIMERS)()
        L.femb200_phase_ms(h, ph)
        nnz = L.femb200_nnz(h)
        L.femb200_assembler_destroy(h)
        return [float(x) for x in ph], nnz

    desc0 = ctx_holder.make_desc(w["elemT"], w["c"])
    for _ in range(warmup):
        resident_step_shared_ctx()
        flush.fill_(1)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = L.femb200_kernel_launches(ctx_holder.ctx)
    phases, wall = [], []
    for _ in range(args.steps):
        flush.fill_(1)                     # L2 flush between timed iterations (outside the timed region)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ph, nnz = resident_step_shared_ctx()   # returns after the library synchronised its stream
        wall.append(time.perf_counter() - t0)
        phases.append(ph)
    torch.cuda.synchronize()
    launches = L.femb200_kernel_launches(ctx_holder.ctx) - launches0
    clocks = sampler.finish()
    ph_mean = np.mean(np.array(phases), axis=0)
    dev_ms = float(ph_mean[6])             # CUDA-event time of the whole call on the library's stream
    t_dev = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
    dev_ms_max = float(t_dev.item())
    value = world * ne / (dev_ms_max * 1e-3)

    # ---- numeric-only re-assembly on the cached pattern (femb200_reassemble; SURVEY 8d "secondary" timing) ------------
    h_re = C.c_void_p()
    assert L.femb200_assembler_create(ctx_holder.ctx, C.c_void_p(d_nodes.data_ptr()), nodes.shape[0], w["ndn"], 1, C.byref(h_re)) == 0
    assert L.femb200_add_isoparametric(h_re, C.byref(desc0), C.c_void_p(d_conn.data_ptr()), 0, 1, ne) == 0
    re_ms = []
    for i in range(0 if args.no_reassemble else warmup + args.steps):
        flush.fill_(1)
        torch.cuda.synchronize()
        rc = L.femb200_reassemble(h_re, C.byref(desc0), None, 0)
        assert rc == 0, rc
        ph = (C.c_float * pkg.N_TIMERS)()
        L.femb200_phase_ms(h_re, ph)
        if i >= warmup:
            re_ms.append(float(ph[6]))
    L.femb200_assembler_destroy(h_re)
    re_ms = float(np.mean(re_ms)) if re_ms else float("nan")

    # ---- end to end through the host API with host buffers ------------------------------------------
    h_nodes = torch.from_numpy(nodes).pin_memory()
    h_conn = torch.from_numpy(conn).pin_memory()          # int64: a Go []int
    h_indptr = torch.empty(ndof + 1, dtype=torch.int64).pin_memory()
    h_indices = torch.empty(nnz, dtype=torch.int32).pin_memory()
    h_values = torch.empty(nnz, dtype=torch.float64).pin_memory()

    def e2e_step():
        h = C.c_void_p()
        rc = L.femb200_assembler_create(ctx_holder.ctx, C.c_void_p(h_nodes.data_ptr()), nodes.shape[0], w["ndn"], 0, C.byref(h))
        assert rc == 0
        rc = L.femb200_add_isoparametric(h, C.byref(desc0), C.c_void_p(h_conn.data_ptr()), 1, 0, ne)
        assert rc == 0, rc
        rc = L.femb200_csr_copy(h, C.c_void_p(h_indptr.data_ptr()), C.c_void_p(h_indices.data_ptr()), C.c_void_p(h_values.data_ptr()))
        assert rc == 0
        L.femb200_assembler_destroy(h)

    for _ in range(warmup):
        e2e_step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e2e_t = []
    for _ in range(args.steps):
        flush.fill_(1)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e2e_step()
        e2e_t.append(time.perf_counter() - t0)
    t_e2e = torch.tensor([float(np.mean(e2e_t))], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e_value = world * ne / float(t_e2e.item())
    h2d = nodes.nbytes + conn.nbytes
    d2h = h_indptr.numel() * 8 + h_indices.numel() * 4 + h_values.numel() * 8

    # ---- roofline of the dominant kernel ---------------------------------------------------------------
    # k_rows (B^T C B products fused with the deterministic row reduction) has the largest share of the step.
    # achieved = algorithmic bytes of one launch / its CUDA-event time ([9] of femb200_phase_ms: events recorded
    # on the library's stream right before and after the launch), averaged over the timed steps.
    peak, peak_src = measured_peaks()
    names = list(pkg.TIMER_NAMES)
    kern_ms = float(ph_mean[names.index("k_rows")])
    abytes = algorithmic_bytes(w, nnz)
    achieved = abytes / (kern_ms * 1e-3) / 1e9
    probe = (C.c_double * 2)()
    rc = L.femb200_probe_peaks(ctx_holder.ctx, 5, probe)
    roofline = {"bound": "hbm", "kernel": "k_rows (B^T C B per scalar CSR row fused with the deterministic segmented reduction)",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                "peak_source": peak_src, "algorithmic_bytes": abytes, "kernel_ms": kern_ms,
                "kernel_share_of_step": kern_ms / dev_ms,
                "algorithmic_bytes_def": "staged geometry records read + incidence list + element->slot map (pos16) read + CSR values written",
                "probe_this_run": {"fp64_dfma_tflops": float(probe[0]), "hbm_copy_gbs": float(probe[1])} if rc == 0 else None,
                "whole_step": {"compulsory_bytes": compulsory_bytes(w, nnz, ndof), "hbm_floor_ms": compulsory_bytes(w, nnz, ndof) / (peak * 1e9) * 1e3,
                               "frac_of_hbm_roofline": compulsory_bytes(w, nnz, ndof) / (peak * 1e9) * 1e3 / dev_ms}}
    traffic_file = os.path.join(ROOT, "profiles", "traffic_k_rows.json")
    if os.path.exists(traffic_file):
        roofline["traffic"] = json.load(open(traffic_file)).get(args.config)

    # ---- CPU baseline (rank 0, N=1 only) ---------------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        dt = cpu_port_seconds(w, nodes, conn)
        cpu = {"value": ne / dt, "unit": "elements/s", "cores": 1, "kind": "port",
               "sample": f"full workload once ({ne} elements, {dt:.1f} s): C port of the reference algorithm (dense per-element Ke, "
                         f"nd^2 COO triplets, sort + in-order accumulate), single thread like the reference's loop (fem.go:140); "
                         f"host cpus={os.cpu_count()}"}

    if rank == 0:
        line = {"metric": "K assembly elements/s", "value": value, "unit": "elements/s", "n_gpus": world, "steps": args.steps, "warmup": warmup,
                "ms_per_step": dev_ms_max, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": w["desc"], "l2": "flushed between timed steps (256 MiB write)", "inputs": "nodes f64 + int32 connectivity resident in HBM",
                           "multi_gpu": "replicas (one full mesh per rank)" if world > 1 else "single GPU"},
                "dofs_per_s": value * ndof / ne,
                "wall_ms_per_step_host_clock": 1e3 * float(np.mean(wall)),
                "phases_ms": dict(zip(pkg.TIMER_NAMES, [float(x) for x in ph_mean])),
                "e2e": {"value": e2e_value, "unit": "elements/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": 1e3 * float(t_e2e.item())},
                "reassemble_cached_pattern": {"value": ne / (re_ms * 1e-3), "unit": "elements/s", "ms_per_step": re_ms,
                                              "note": "femb200_reassemble: geometry kernel + node-row kernel on the cached symbolic plan (not the headline: "
                                                      "the reference rebuilds everything on every call)"},
                "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
                "published_go_reference": {"elements_per_s": 39380, "hardware": "AMD Ryzen 5 3400G, README.md:86 (33.0 s for this mesh)"}}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
