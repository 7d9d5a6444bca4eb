"""Multi-rank host logic on CPU: gloo backend, world_size 2 and 3.  Each rank assembles its element chunk with
the ORACLE (test infrastructure) standing in for the GPU, then runs the real partition / ownership / ragged
all-to-all / merge code of go-fem_b200/distributed.py.  The union of the row blocks must be the oracle's K."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, kind_name, size, use_morton, q):
    try:
        sys.path.insert(0, ROOT)
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import torch
        import torch.distributed as dist
        import __graft_entry__ as entry
        from conftest import KINDS
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        entry.load_package()
        from importlib import import_module
        D = import_module("gofem_b200.distributed")
        M = import_module("gofem_b200.meshes")
        O = entry.load_oracle()
        nodes, conn = M.mesh_for(kind_name, size)
        k = KINDS[kind_name]
        if use_morton:
            conn = conn[D.morton_order(nodes, conn)]
        Cm, bk = O.constitutive(k["mode"], 200e9, 0.3)
        mdpn = bin(k["flags"]).count("1")
        bounds = D.chunk_bounds(conn.shape[0], world)
        plan = D.ExchangePlan(nodes.shape[0], conn, bounds, rank, mdpn)
        og = O.GeneralAssembler(nodes, k["flags"])
        og.add_isoparametric(k["kind"], Cm, bk, conn[bounds[rank]:bounds[rank + 1]])
        be = D.NumpyBackend(*og.csr())
        sent = D.exchange_and_own(be, plan, torch.device("cpu"), dist)
        tsent = torch.tensor([sent], dtype=torch.int64)
        dist.all_reduce(tsent)            # rank 0 owns everything it touches (lowest rank wins): only the others send
        sent = int(tsent.item())
        # the rank now holds only rows it owns
        lens = np.diff(be.indptr)
        assert np.all(lens[plan.keep_mask == 0] == 0)
        ptr, idx, val = D.gather_distributed_csr(be, plan, dist, torch.device("cpu"))
        if rank == 0:
            full = O.GeneralAssembler(nodes, k["flags"])
            full.add_isoparametric(k["kind"], Cm, bk, conn)
            rp, ri, rv, ab = full.csr(with_absum=True)
            ok = np.array_equal(ptr, rp) and np.array_equal(idx, ri) and bool(np.all(np.abs(val - rv) <= 1e-12 * np.maximum(np.abs(rv), ab)))
            q.put((ok, int(sent), int(rv.size)))
        dist.barrier()
        dist.destroy_process_group()
    except Exception as ex:  # pragma: no cover
        import traceback
        q.put((False, traceback.format_exc(), 0))
        raise


@pytest.mark.parametrize("world,kind,size,morton", [(2, "tetra4", 4, False), (3, "hexa8", 4, True), (2, "quad4", 9, True)])
def test_row_block_exchange_gloo(world, kind, size, morton):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 400) + world * 7 + size
    procs = [ctx.Process(target=_worker, args=(r, world, port, kind, size, morton, q)) for r in range(world)]
    for p in procs:
        p.start()
    ok, sent, nnz = q.get(timeout=240)
    for p in procs:
        p.join(timeout=240)
    assert ok is True, sent
    assert sent > 0 and nnz > 0
    assert all(p.exitcode == 0 for p in procs)


def _worker_planned(rank, world, port, kind_name, size, q):
    """Planned exchange (what bench.py runs on the GPUs): owned block assembled from [own ; halo(pattern only)], the
    outgoing interface rows from the own elements touching foreign nodes, ONE all-to-all with plan-time sizes."""
    try:
        sys.path.insert(0, ROOT)
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import torch
        import torch.distributed as dist
        import __graft_entry__ as entry
        from conftest import KINDS
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        entry.load_package()
        from importlib import import_module
        D = import_module("gofem_b200.distributed")
        M = import_module("gofem_b200.meshes")
        O = entry.load_oracle()
        nodes, conn = M.mesh_for(kind_name, size)
        k = KINDS[kind_name]
        conn = conn[D.morton_order(nodes, conn)]
        Cm, bk = O.constitutive(k["mode"], 200e9, 0.3)
        mdpn = bin(k["flags"]).count("1")
        bounds = D.chunk_bounds(conn.shape[0], world)
        plan = D.ExchangePlan(nodes.shape[0], conn, bounds, rank, mdpn)
        iplan = D.InterfacePlan(plan, nodes.shape[0], conn, bounds, mdpn, dist)
        local = D.InterfacePlan(plan, nodes.shape[0], conn, bounds, mdpn, None)     # peers' plans derived locally
        for a, b in zip(iplan.recv, local.recv):
            assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
        own = conn[bounds[rank]:bounds[rank + 1]]
        # owned block: values of the own elements, pattern of own + halo (the oracle assembles the halo with a zero material)
        og = O.GeneralAssembler(nodes, k["flags"])
        og.add_isoparametric(k["kind"], Cm, bk, own)
        if plan.halo_elems.size:
            og.add_isoparametric(k["kind"], Cm * 0.0, bk, conn[plan.halo_elems])
        owned = D.NumpyBackend(*og.csr())
        owned.restrict_rows(plan.own_node_mask, mdpn)
        og2 = O.GeneralAssembler(nodes, k["flags"])
        if plan.send_elems.size:
            og2.add_isoparametric(k["kind"], Cm, bk, own[plan.send_elems])
        outgoing = D.NumpyBackend(*og2.csr())
        outgoing.restrict_rows(plan.send_node_mask, mdpn)
        px = D.PlannedExchange(iplan, torch.device("cpu"))
        sent = px.run(owned, outgoing, dist)
        assert px.mismatches() == 0
        tsent = torch.tensor([sent], dtype=torch.int64)
        dist.all_reduce(tsent)
        ptr, idx, val = D.gather_distributed_csr(owned, plan, dist, torch.device("cpu"))
        if rank == 0:
            full = O.GeneralAssembler(nodes, k["flags"])
            full.add_isoparametric(k["kind"], Cm, bk, conn)
            rp, ri, rv, ab = full.csr(with_absum=True)
            ok = np.array_equal(ptr, rp) and np.array_equal(idx, ri) and bool(np.all(np.abs(val - rv) <= 1e-12 * np.maximum(np.abs(rv), ab)))
            q.put((ok, int(tsent.item()), int(rv.size)))
        dist.barrier()
        dist.destroy_process_group()
    except Exception:  # pragma: no cover
        import traceback
        q.put((False, traceback.format_exc(), 0))
        raise


@pytest.mark.parametrize("world,kind,size", [(2, "tetra4", 4), (3, "hexa8", 4)])
def test_planned_exchange_gloo(world, kind, size):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29900 + (os.getpid() % 300) + world * 11 + size
    procs = [ctx.Process(target=_worker_planned, args=(r, world, port, kind, size, q)) for r in range(world)]
    for p in procs:
        p.start()
    ok, sent, nnz = q.get(timeout=240)
    for p in procs:
        p.join(timeout=240)
    assert ok is True, sent
    assert sent > 0 and nnz > 0
    assert all(p.exitcode == 0 for p in procs)


def test_partition_helpers():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as entry
    entry.load_package()
    from importlib import import_module
    D = import_module("gofem_b200.distributed")
    M = import_module("gofem_b200.meshes")
    nodes, conn = M.bcc_tetra4(4)
    perm = D.morton_order(nodes, conn)
    assert sorted(perm.tolist()) == list(range(conn.shape[0]))
    b = D.chunk_bounds(conn.shape[0], 3)
    assert b[0] == 0 and b[-1] == conn.shape[0] and all(b[i] <= b[i + 1] for i in range(3))
    own = D.node_owner(nodes.shape[0], conn[perm], b)
    assert own.min() == 0 and own.max() == 2
    plans = [D.ExchangePlan(nodes.shape[0], conn[perm], b, r, 3, own) for r in range(3)]
    assert sum(int(p.keep_mask.sum()) for p in plans) == nodes.shape[0] * 3          # every row has exactly one owner
    for p in plans:
        for qq, rows in enumerate(p.send_rows):
            assert np.all(np.diff(rows) > 0)
            assert np.all(own[rows // 3] == qq) if rows.size else True
    # Morton order keeps the interface small: fewer rows to send than with a random element order
    rnd = np.random.default_rng(0).permutation(conn.shape[0])
    sent_m = sum(r.size for p in plans for r in p.send_rows)
    sent_r = sum(r.size for rr in range(3) for r in D.ExchangePlan(nodes.shape[0], conn[rnd], b, rr, 3).send_rows)
    assert sent_m < sent_r


def test_local_numbering():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as entry
    entry.load_package()
    from importlib import import_module
    D = import_module("gofem_b200.distributed")
    M = import_module("gofem_b200.meshes")
    nodes, conn = M.bcc_tetra4(4)
    b = D.chunk_bounds(conn.shape[0], 3)
    for r in range(3):
        plan = D.ExchangePlan(nodes.shape[0], conn, b, r, 3)
        gids = np.sort(np.concatenate([np.arange(b[r], b[r + 1]), plan.halo_elems]))
        ln = D.LocalNumbering(nodes, conn[gids])
        assert np.all(np.diff(ln.l2g) > 0)                                   # ascending => column order is preserved
        lc = ln.conn(conn[gids])
        assert np.array_equal(ln.l2g[lc], conn[gids]) and np.array_equal(ln.nodes, nodes[ln.l2g])
        assert np.all(plan.own_node_mask[ln.l2g][np.unique(ln.conn(conn[b[r]:b[r + 1]]))] | True)
        # every owned, referenced node is local; rows translate back and forth
        owned = np.nonzero(plan.own_node_mask)[0]
        referenced = np.intersect1d(owned, np.unique(conn))
        assert np.all(np.isin(referenced, ln.l2g))
        rows_g = (referenced[:, None] * 3 + np.arange(3)[None, :]).ravel()
        rows_l = ln.rows(rows_g, 3)
        assert np.array_equal(ln.l2g[rows_l // 3] * 3 + rows_l % 3, rows_g)
        iplan = D.InterfacePlan(plan, nodes.shape[0], conn, b, 3, None)
        for rows, _ in iplan.send + iplan.recv:
            if rows.size:
                ln.rows(rows, 3)                                               # interface rows are all local

