"""Large sharded configurations of BASELINE.json on N GPUs of one node (one rank per GPU, launched with torchrun):

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 \
      scripts/dist_config.py tetra4 204 [steps]            # "S": BCC cube N=204, 101,376,576 Tetra4 elements
  ... scripts/dist_config.py hexa20 400 250 200 [steps]    # config 5: 20,000,000 Hexa20 elements, 3x3x3 Gauss

Rank 0 generates the global mesh and the partition (contiguous cell-major element chunks, lowest-rank node ownership,
halo elements, rank-local node numbering) and scatters every rank's share; the other ranks never hold the global mesh.
Every rank then assembles its row block of the distributed CSR K (owner computes: own + halo elements numerically,
ascending global element order, no data-path collective).  Timed with CUDA events, max over ranks.
Checks: the global nnz equals the closed-form node-pair count of the mesh family (pattern), and on a sample of owned rows
rigid translations are in the null space (row sums per dof direction vanish) and the diagonal is positive (values)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402


def main():
    if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"
    import ctypes as C
    import torch
    import torch.distributed as dist
    from importlib import import_module
    rank, world, lrank = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lrank)
    dev = torch.device("cuda", lrank)
    dist.init_process_group("nccl", device_id=dev)
    pkg = entry.load_package()
    D = import_module("gofem_b200.distributed")
    M = import_module("gofem_b200.meshes")
    kind = sys.argv[1]
    if kind == "tetra4":
        n = int(sys.argv[2])
        steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
        desc_s = f"Tetra4 BCC cube N={n}"
        elemT = pkg.elements.Tetra4()
        exp_blocks = 30 * n ** 3 + 9 * n ** 2 - 15 * n - 23
    else:
        n1, n2, n3 = int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
        steps = int(sys.argv[5]) if len(sys.argv) > 5 else 3
        desc_s = f"Hexa20 brick {n1}x{n2}x{n3}, 3x3x3 Gauss"
        elemT = pkg.elements.Hexa20()
        exp_blocks = 1 + 8 * (n1 + n2 + n3) + 47 * (n1 * n2 + n1 * n3 + n2 * n3) + 234 * n1 * n2 * n3
    c = pkg.solids.Isotropic(200e9, 0.3).Solid3D()

    t0 = time.time()
    shares = [None] * world
    meta = [None]
    if rank == 0:
        if kind == "tetra4":
            nodes, conn = M.bcc_tetra4_box(n, n, n, 1.0 / n)
        else:
            nodes, conn = M.hexa20_grid(n1, n2, n3)
        conn = np.ascontiguousarray(conn.astype(np.int32))
        ne, nn = conn.shape
        nnodes = nodes.shape[0]
        bounds = D.chunk_bounds(ne, world)
        owner = np.full(nnodes, world, dtype=np.int8)
        for r in range(world - 1, -1, -1):                      # descending: the lowest rank writes last
            owner[conn[bounds[r]:bounds[r + 1]].ravel()] = r
        owner[owner == world] = 0
        for r in range(world):
            touch = (owner == r)[conn].any(axis=1)
            touch[bounds[r]:bounds[r + 1]] = True                # own elements + halo, ascending global index
            gids = np.nonzero(touch)[0]
            nhalo = int(gids.size - (bounds[r + 1] - bounds[r]))
            ln = D.LocalNumbering(nodes, conn[gids])
            shares[r] = dict(l2g=ln.l2g.astype(np.int32), nodes=ln.nodes, conn=ln.conn(conn[gids]), own=(owner[ln.l2g] == r).astype(np.uint8),
                             nhalo=nhalo, n_own=int(bounds[r + 1] - bounds[r]))
            del touch, gids, ln
        meta[0] = dict(ne=int(ne), nnodes=int(nnodes), nn=int(nn), host_s=time.time() - t0)
        del nodes, conn, owner
    mine = [None]
    dist.scatter_object_list(mine, shares if rank == 0 else None, src=0)
    dist.broadcast_object_list(meta, src=0)
    sh, meta = mine[0], meta[0]
    del shares
    ne_total, n_global_nodes = meta["ne"], meta["nnodes"]

    d_nodes = torch.from_numpy(sh["nodes"]).to(dev)
    d_conn = torch.from_numpy(sh["conn"]).to(dev)
    d_l2g = torch.from_numpy(sh["l2g"]).to(dev)
    d_own = torch.from_numpy(sh["own"]).to(dev)
    n_local, n_el = sh["nodes"].shape[0], sh["conn"].shape[0]
    L = pkg.lib()
    holder = pkg.RawAssembler(None, 3, device=lrank, nodes_device_ptr=d_nodes.data_ptr(), n_nodes=n_local)
    L.femb200_ctx_set_stream(holder.ctx, C.c_void_p(torch.cuda.current_stream().cuda_stream))
    desc = holder.make_desc(elemT, c)

    def assemble():
        h = C.c_void_p()
        assert L.femb200_assembler_create(holder.ctx, C.c_void_p(d_nodes.data_ptr()), n_local, 3, 1, C.byref(h)) == 0
        L.femb200_assembler_set_node_mask_device(h, C.c_void_p(d_own.data_ptr()))
        assert L.femb200_assembler_set_column_map_device(h, C.c_void_p(d_l2g.data_ptr()), n_global_nodes) == 0
        rc = L.femb200_add_isoparametric(h, C.byref(desc), C.c_void_p(d_conn.data_ptr()), 0, 1, n_el)
        assert rc == 0, (rc, L.femb200_last_cuda_error(holder.ctx))
        return h

    times, phases = [], []
    h = None
    for it in range(1 + steps):
        if h is not None:
            L.femb200_assembler_destroy(h)
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        h = assemble()
        e1.record()
        torch.cuda.synchronize()
        if it >= 1:
            times.append(e0.elapsed_time(e1))
            ph = (C.c_float * pkg.N_TIMERS)()
            L.femb200_phase_ms(h, ph)
            phases.append([float(v) for v in ph])
    # ---- checks on the last assembly -----------------------------------------------------------------------------
    nnz = L.femb200_nnz(h)
    ip_p, ix_p, va_p = C.c_void_p(), C.c_void_p(), C.c_void_p()
    L.femb200_csr_device(h, C.byref(ip_p), C.byref(ix_p), C.byref(va_p))
    indptr = D.device_tensor(ip_p.value, n_local * 3 + 1, "<i8", dev)
    indices = D.device_tensor(ix_p.value, nnz, "<i4", dev)
    values = D.device_tensor(va_p.value, nnz, "<f8", dev)
    own_rows = torch.nonzero(d_own.repeat_interleave(3)).flatten()
    g = torch.Generator(device="cpu"); g.manual_seed(5 + rank)
    pick = own_rows[torch.randperm(own_rows.numel(), generator=g)[:4000].to(dev)]
    worst, diag_ok, checked = 0.0, True, 0
    ipc = indptr.cpu().numpy()
    for r in pick.cpu().numpy():
        lo, hi = int(ipc[r]), int(ipc[r + 1])
        if hi == lo:
            continue
        cols = indices[lo:hi].cpu().numpy().astype(np.int64)
        vals = values[lo:hi].cpu().numpy()
        grow = int(sh["l2g"][r // 3]) * 3 + r % 3
        scale = np.abs(vals).sum()
        for k in range(3):
            worst = max(worst, abs(vals[cols % 3 == k].sum()) / scale)
        dpos = np.nonzero(cols == grow)[0]
        diag_ok &= bool(dpos.size == 1 and vals[dpos[0]] > 0)
        assert np.all(np.diff(cols) > 0)
        checked += 1
    stats = torch.tensor([float(nnz), float(sh["nhalo"]), float(checked)], dtype=torch.float64, device=dev)
    dist.all_reduce(stats)
    mx = torch.tensor([float(np.mean(times)), worst, 0.0 if diag_ok else 1.0, float(torch.cuda.max_memory_allocated(dev)) / 2 ** 30], dtype=torch.float64, device=dev)
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    phm = torch.tensor(np.mean(np.array(phases), axis=0), dtype=torch.float64, device=dev)
    dist.all_reduce(phm, op=dist.ReduceOp.MAX)
    if rank == 0:
        ms = mx[0].item()
        out = {"config": f"{desc_s}: {ne_total} elements / {n_global_nodes * 3} dofs, Isotropic(E=200e9, nu=0.3).Solid3D", "n_gpus": world,
               "steps": steps, "ms_per_step": ms, "elements_per_s": ne_total / (ms * 1e-3), "dofs_per_s": n_global_nodes * 3 / (ms * 1e-3),
               "nnz_total": int(stats[0].item()), "nnz_expected": int(9 * exp_blocks), "nnz_ok": int(stats[0].item()) == int(9 * exp_blocks),
               "redundant_halo_elements_total": int(stats[1].item()), "rows_sampled": int(stats[2].item()),
               "max_rigid_translation_rowsum_over_abs_rowsum": mx[1].item(), "diagonal_positive": mx[2].item() == 0.0,
               "scheme": "owner computes (own + halo elements, rank-local node numbering, no data-path collective)",
               "phases_ms_max_over_ranks": dict(zip(pkg.TIMER_NAMES, [round(float(v), 3) for v in phm.cpu().numpy()])),
               "host_partition_s_rank0": round(meta["host_s"], 1), "lib_device_mem_note": "scratch arena + K per rank; torch-side inputs extra"}
        print(json.dumps(out), flush=True)
    L.femb200_assembler_destroy(h)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
