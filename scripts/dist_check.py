"""Multi-GPU parity check (run under torchrun, one rank per GPU):
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/dist_check.py [kind size]
Each rank assembles its Morton-ordered element chunk on its GPU, interface rows are exchanged over NCCL, and the
union of the row blocks is compared with the CPU oracle's K on rank 0 (pattern bit-exact, values 1e-12)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, torch.distributed as dist
import __graft_entry__ as entry
from conftest import KINDS
from test_gpu_parity import elem_of, constituter_of

rank, world, lrank = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lrank)
dev = torch.device("cuda", lrank)
dist.init_process_group("nccl", device_id=dev)
pkg = entry.load_package()
from importlib import import_module
D = import_module("gofem_b200.distributed"); M = import_module("gofem_b200.meshes")
cases = [(sys.argv[1], int(sys.argv[2]))] if len(sys.argv) > 2 else [("tetra4", 10), ("hexa8", 8), ("quad4", 40), ("tetra10", 4)]
ok_all = True
for name, size in cases:
    nodes, conn = M.mesh_for(name, size)
    conn = conn[D.morton_order(nodes, conn)]
    k = KINDS[name]
    mdpn = bin(k["flags"]).count("1")
    bounds = D.chunk_bounds(conn.shape[0], world)
    plan = D.ExchangePlan(nodes.shape[0], conn, bounds, rank, mdpn)
    raw = pkg.RawAssembler(nodes, mdpn, device=lrank)
    desc = raw.make_desc(elem_of(pkg, name), constituter_of(pkg, k["mode"], 200e9, 0.3))
    import ctypes as C
    L = pkg.lib()
    own_conn = conn[bounds[rank]:bounds[rank + 1]]
    if os.environ.get("FEMB200_GENERIC_EXCHANGE"):
        rc = raw.add(desc, conn=own_conn)
        assert rc == 0, (rc, raw.last_error())
        be = D.Femb200Backend(pkg, raw, dev)
        sent = D.exchange_and_own(be, plan, dev, dist)
        raw2 = None
    else:
        m_own = torch.as_tensor(plan.own_node_mask, device=dev)
        m_send = torch.as_tensor(plan.send_node_mask, device=dev)
        L.femb200_assembler_set_node_mask_device(raw.h, C.c_void_p(m_own.data_ptr()))
        lconn, nhalo = plan.local_connectivity(conn, bounds)
        rc = raw.add(desc, conn=lconn, n_pattern_only=nhalo)
        assert rc == 0, (rc, raw.last_error())
        raw2 = pkg.RawAssembler(nodes, mdpn, device=lrank)
        L.femb200_assembler_set_node_mask_device(raw2.h, C.c_void_p(m_send.data_ptr()))
        rc = raw2.add(desc, conn=own_conn[plan.send_elems])
        assert rc == 0, (rc, raw2.last_error())
        be = D.Femb200Backend(pkg, raw, dev)
        if os.environ.get("FEMB200_RAGGED_EXCHANGE"):
            sent = D.exchange_into_owned(be, D.Femb200Backend(pkg, raw2, dev), plan, dev, dist)
        else:
            iplan = D.InterfacePlan(plan, nodes.shape[0], conn, bounds, mdpn, dist)
            px = D.PlannedExchange(iplan, dev)
            # the library's stream must be the stream NCCL is enqueued on
            for r_ in (raw, raw2):
                L.femb200_ctx_set_stream(r_.ctx, C.c_void_p(torch.cuda.current_stream().cuda_stream))
            sent = px.run(be, D.Femb200Backend(pkg, raw2, dev), dist)
            torch.cuda.synchronize()
            assert px.mismatches() == 0, px.mismatches()
    ptr, idx, val = D.gather_distributed_csr(be, plan, dist, dev)
    if rank == 0:
        O = entry.load_oracle()
        Cm, bk = O.constitutive(k["mode"], 200e9, 0.3)
        og = O.GeneralAssembler(nodes, k["flags"]); og.add_isoparametric(k["kind"], Cm, bk, conn)
        rp, ri, rv, ab = og.csr(with_absum=True)
        pat = np.array_equal(ptr, rp) and np.array_equal(idx, ri)
        nerr = (np.abs(val - rv) / np.maximum(np.maximum(np.abs(rv), ab), 1e-300)).max() if pat else float("inf")
        ok = pat and nerr <= 1e-12
        ok_all &= ok
        print(f"[dist_check] {name} size={size} world={world} nelem={conn.shape[0]} nnz={rv.size} pattern_exact={pat} max_norm_err={nerr:.2e} -> {'OK' if ok else 'FAIL'}", flush=True)
    # owner-computes variant: own + halo elements, all numeric, ascending global element index; no collective
    raw3 = pkg.RawAssembler(nodes, mdpn, device=lrank)
    m_own3 = torch.as_tensor(plan.own_node_mask, device=dev)
    L.femb200_assembler_set_node_mask_device(raw3.h, C.c_void_p(m_own3.data_ptr()))
    gids = np.sort(np.concatenate([np.arange(bounds[rank], bounds[rank + 1]), plan.halo_elems]))
    rc = raw3.add(desc, conn=conn[gids])
    assert rc == 0, (rc, raw3.last_error())
    ptr3, idx3, val3 = D.gather_distributed_csr(D.Femb200Backend(pkg, raw3, dev), plan, dist, dev)
    if rank == 0:
        pat3 = np.array_equal(ptr3, rp) and np.array_equal(idx3, ri)
        nerr3 = (np.abs(val3 - rv) / np.maximum(np.maximum(np.abs(rv), ab), 1e-300)).max() if pat3 else float("inf")
        ok3 = pat3 and nerr3 <= 1e-12
        ok_all &= ok3
        print(f"[dist_check] {name} owner-computes: pattern_exact={pat3} max_norm_err={nerr3:.2e} -> {'OK' if ok3 else 'FAIL'}", flush=True)
    # the same two schemes over the rank's LOCAL node numbering (what bench.py times): arrays sized by the rank's share
    ln = D.LocalNumbering(nodes, conn[gids])
    d_l2g = torch.from_numpy(ln.l2g.astype(np.int32)).to(dev)
    m_own_l = torch.as_tensor(ln.node_mask(plan.own_node_mask), device=dev)
    m_send_l = torch.as_tensor(ln.node_mask(plan.send_node_mask), device=dev)

    def local_asm(mask_t):
        r_ = pkg.RawAssembler(ln.nodes, mdpn, device=lrank)
        L.femb200_ctx_set_stream(r_.ctx, C.c_void_p(torch.cuda.current_stream().cuda_stream))
        L.femb200_assembler_set_node_mask_device(r_.h, C.c_void_p(mask_t.data_ptr()))
        assert L.femb200_assembler_set_column_map_device(r_.h, C.c_void_p(d_l2g.data_ptr()), nodes.shape[0]) == 0
        return r_

    raw4 = local_asm(m_own_l)
    rc = raw4.add(desc, conn=ln.conn(conn[gids]))
    assert rc == 0, (rc, raw4.last_error())
    ptr4, idx4, val4 = D.gather_distributed_csr_local(D.Femb200Backend(pkg, raw4, dev), ln, ln.node_mask(plan.own_node_mask), mdpn, nodes.shape[0], dist)
    raw5 = local_asm(m_own_l)
    lc, nh = plan.local_connectivity(conn, bounds)
    rc = raw5.add(desc, conn=ln.conn(lc), n_pattern_only=nh)
    assert rc == 0, (rc, raw5.last_error())
    raw6 = local_asm(m_send_l)
    rc = raw6.add(desc, conn=ln.conn(own_conn[plan.send_elems]) if plan.send_elems.size else np.zeros((0, conn.shape[1]), dtype=np.int32))
    assert rc == 0, (rc, raw6.last_error())
    iplan2 = D.InterfacePlan(plan, nodes.shape[0], conn, bounds, mdpn, dist)
    px2 = D.PlannedExchange(iplan2, dev, local=ln)
    px2.run(D.Femb200Backend(pkg, raw5, dev), D.Femb200Backend(pkg, raw6, dev), dist)
    torch.cuda.synchronize()
    assert px2.mismatches() == 0, px2.mismatches()
    ptr5, idx5, val5 = D.gather_distributed_csr_local(D.Femb200Backend(pkg, raw5, dev), ln, ln.node_mask(plan.own_node_mask), mdpn, nodes.shape[0], dist)
    if rank == 0:
        for tag, (pp, ii, vv) in (("owner-computes/local", (ptr4, idx4, val4)), ("row-exchange/local", (ptr5, idx5, val5))):
            patx = np.array_equal(pp, rp) and np.array_equal(ii, ri)
            nerrx = (np.abs(vv - rv) / np.maximum(np.maximum(np.abs(rv), ab), 1e-300)).max() if patx else float("inf")
            okx = patx and nerrx <= 1e-12
            ok_all &= okx
            print(f"[dist_check] {name} {tag}: pattern_exact={patx} max_norm_err={nerrx:.2e} -> {'OK' if okx else 'FAIL'}", flush=True)
    for r_ in (raw4, raw5, raw6):
        r_.close()
    raw.close()
    raw3.close()
    if raw2 is not None:
        raw2.close()
    dist.barrier()
dist.destroy_process_group()
if rank == 0 and not ok_all:
    sys.exit(1)
