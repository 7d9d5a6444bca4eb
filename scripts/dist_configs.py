"""Every element type of BASELINE.json on N GPUs of one node through the C ABI (femb200_dist_*), one rank per GPU:

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 \
      scripts/dist_configs.py [config2 config3 config4 config5] [--steps 3] [--out gpurun_out/dist_configs.json]

  config2  Quad4 axisymmetric 2000 x 2000 (4 M elements)            config4  Tetra10 BCC N=95 (10.2 M elements, 4-point rule)
  config3  Hexa8 200^3 (8 M elements, 2x2x2 Gauss)                   config5  Hexa20 400 x 250 x 200 (20 M elements, 3x3x3 Gauss; 8 GPUs)

Rank 0 builds the mesh with the numpy generators (go-fem_b200/meshes.py) and broadcasts nodes + int32 connectivity over NCCL;
EVERY rank then holds the whole mesh on its device and partitions it there itself (femb200_dist_create: Morton order, lowest
rank owns a node, halo elements), so the plan needs no communication.  A step = femb200_dist_reset +
femb200_dist_add_isoparametric (owner computes, no data-path collective); CUDA events on the library stream, max over ranks.
Reported per config: elements/s, DOFs/s, the whole-step fraction of the SURVEY 8(d) roofline on N GPUs
(max(F_dense / (N * P64), B_compulsory / (N * BW))), the partition time, and the closed-form nnz check."""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402
import bench  # noqa: E402

CONFIGS = {
    # name: (mesh kind, mesh_for size or dims, element class, constituter method, model dofs, radial)
    "config2": ("quad4", 2000, "Quad4", "Axisymmetric", 2, True),
    "config3": ("hexa8", 200, "Hexa8", "Solid3D", 3, False),
    "config4": ("tetra10", 95, "Tetra10", "Solid3D", 3, False),
    "config5": ("hexa20", (400, 250, 200), "Hexa20", "Solid3D", 3, False),
}


def expected_nnz(name, nodes, size):
    if name == "config2":
        return 144_048_004
    if name == "config3":
        return 1_953_736_209
    if name == "config4":
        return 3_533_762_313
    n1, n2, n3 = size
    s1, s2, s3 = n1 + n2 + n3, n1 * n2 + n1 * n3 + n2 * n3, n1 * n2 * n3
    return 9 * (1 + 8 * s1 + 47 * s2 + 234 * s3)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("configs", nargs="*", default=["config2", "config3", "config4"])
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--out", default="gpurun_out/dist_configs.json")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from importlib import import_module
    rank, world, lrank = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lrank)
    dev = torch.device("cuda", lrank)
    dist.init_process_group("nccl", device_id=dev)
    pkg = entry.load_package()
    M = import_module("gofem_b200.meshes")
    L = pkg.lib()
    holder = pkg.RawAssembler(np.zeros((1, 3)), 3, device=lrank)
    ctx = holder.ctx
    peak_bw, _ = bench.measured_peaks()
    probe = (C.c_double * 2)()
    p64 = float(probe[0]) if L.femb200_probe_peaks(ctx, 3, probe) == 0 and probe[0] > 0 else 37.2
    results = {}
    for name in args.configs:
        kind, size, ecls, cmeth, mdpn, radial = CONFIGS[name]
        t0 = time.time()
        meta = torch.zeros(3, dtype=torch.int64, device=dev)
        if rank == 0:
            nodes, conn = (M.hexa20_grid(*size) if isinstance(size, tuple) else M.mesh_for(kind, size))
            meta = torch.tensor([nodes.shape[0], conn.shape[0], conn.shape[1]], dtype=torch.int64, device=dev)
        dist.broadcast(meta, 0)
        n_nodes, n_elem, nn = (int(v) for v in meta.tolist())
        d_nodes = torch.empty(n_nodes * 3, dtype=torch.float64, device=dev)
        d_conn = torch.empty(n_elem * nn, dtype=torch.int32, device=dev)
        if rank == 0:
            d_nodes.copy_(torch.from_numpy(np.ascontiguousarray(nodes).reshape(-1)))
            d_conn.copy_(torch.from_numpy(np.ascontiguousarray(conn.astype(np.int32)).reshape(-1)))
            del nodes, conn
        dist.broadcast(d_nodes, 0)
        dist.broadcast(d_conn, 0)
        torch.cuda.synchronize()
        gen_s = time.time() - t0
        d = pkg.DistRank(ctx, rank, world, mdpn, nn, nodes_ptr=d_nodes.data_ptr(), conn_ptr=d_conn.data_ptr(), n_nodes=n_nodes, n_elem=n_elem, order_mode=1)
        first_ms = d.info()["partition_ms"]
        d.close()
        d = pkg.DistRank(ctx, rank, world, mdpn, nn, nodes_ptr=d_nodes.data_ptr(), conn_ptr=d_conn.data_ptr(), n_nodes=n_nodes, n_elem=n_elem, order_mode=1)
        info = d.info()
        del d_conn
        torch.cuda.empty_cache()
        elemT = getattr(pkg.elements, ecls)()
        cst = getattr(pkg.solids.Isotropic(200e9, 0.3), cmeth)()
        tmp_ra = pkg.RawAssembler(np.zeros((1, 3)), mdpn, device=lrank)     # only for the descriptor tables (host mirror)
        desc = tmp_ra.make_desc(elemT, cst)
        tmp_ra.close()
        ms, phases = [], None
        for i in range(args.steps + 2):                     # two untimed passes (memory pool growth), then the timed ones
            d.reset()
            rc = d.add(desc)
            if rc:
                raise RuntimeError(f"{name}: femb200_dist_add_isoparametric -> {rc} {d.last_error()}")
            ph = (C.c_float * pkg.N_TIMERS)()
            L.femb200_phase_ms(d.assembler(), ph)
            if i > 1:
                ms.append(float(ph[6]))
                phases = [float(x) for x in ph]
        nnz_rank = d.info()["nnz"]
        t = torch.tensor([float(np.median(ms)), float(info["partition_ms"])], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        tot = torch.tensor([float(nnz_rank), float(info["halo_elems"]), float(info["local_elems"])], dtype=torch.float64, device=dev)
        dist.all_reduce(tot)
        d.close()
        del d_nodes
        torch.cuda.empty_cache()
        if rank == 0:
            step_ms = t[0].item()
            nqp = len(elemT.Quadrature()[1])
            dimC = cst.Constitutive().shape[0]
            dgeo = 2 if kind == "quad4" else 3
            f_tot = bench.dense_flops_per_element(dgeo, nn, mdpn, nqp, dimC, radial) * n_elem
            nnz = int(tot[0].item())
            b_tot = 4 * nn * n_elem + 24 * n_nodes + 8 * (n_nodes * mdpn + 1) + 12 * nnz
            roof_ms = max(f_tot / (world * p64 * 1e12), b_tot / (world * peak_bw * 1e9)) * 1e3
            results[name] = {"workload": f"{ecls} {size}: {n_elem} elements / {n_nodes * mdpn} dofs on {world} GPUs", "n_gpus": world,
                             "ms_per_step": step_ms, "elements_per_s": n_elem / (step_ms * 1e-3), "dofs_per_s": n_nodes * mdpn / (step_ms * 1e-3),
                             "nnz_total": nnz, "nnz_expected": expected_nnz(name, None, size), "nnz_ok": nnz == expected_nnz(name, None, size),
                             "roofline_8d": {"dense_flops": f_tot, "compulsory_bytes": b_tot, "roof_ms": roof_ms, "frac": roof_ms / step_ms,
                                             "bound": "fp64" if f_tot / p64 / 1e12 >= b_tot / peak_bw / 1e9 else "hbm", "p64_tflops": p64, "hbm_gbs": peak_bw},
                             "partition_ms_device_max": t[1].item(), "partition_ms_first_call_rank0": first_ms,
                             "halo_elements_total": int(tot[1].item()), "local_elements_total": int(tot[2].item()),
                             "mesh_build_and_broadcast_s": gen_s, "phases_ms_rank0": dict(zip(pkg.TIMER_NAMES, phases)), "steps": args.steps}
            print(name, json.dumps(results[name]), flush=True)
        dist.barrier()
    if rank == 0:
        os.makedirs(os.path.dirname(os.path.join(ROOT, args.out)), exist_ok=True)
        json.dump(results, open(os.path.join(ROOT, args.out), "w"), indent=1)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
