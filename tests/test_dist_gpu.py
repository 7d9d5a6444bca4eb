"""Multi-GPU path behind the C ABI (femb200_dist_* / femb200_group_*), exercised on ONE device: every rank's partition and
row block is computed one after the other on cuda:0 (the ranks are independent: owner computes, no collective), stitched
together and compared with the single-GPU K (bit for bit: same ascending element order per row) and with the CPU oracle.
Also: the device mesh generator against the numpy one."""
import ctypes as C

import numpy as np
import pytest

from conftest import compare_csr

pytestmark = pytest.mark.gpu


def _stitch(blocks, ndof):
    """row blocks [(global rows, ptr, indices, values)] -> one CSR; every scalar row must be owned exactly once"""
    lens = np.zeros(ndof, dtype=np.int64)
    seen = np.zeros(ndof, dtype=np.int32)
    for rows, ptr, _, _ in blocks:
        lens[rows] = ptr[1:] - ptr[:-1]
        seen[rows] += 1
    indptr = np.zeros(ndof + 1, dtype=np.int64)
    np.cumsum(lens, out=indptr[1:])
    indices = np.zeros(indptr[-1], dtype=np.int32)
    values = np.zeros(indptr[-1], dtype=np.float64)
    for rows, ptr, idx, val in blocks:
        for k, r in enumerate(rows):
            indices[indptr[r]:indptr[r + 1]] = idx[ptr[k]:ptr[k + 1]]
            values[indptr[r]:indptr[r + 1]] = val[ptr[k]:ptr[k + 1]]
    return indptr, indices, values, seen


@pytest.mark.parametrize("kind,size,world,order_mode", [("tetra4", 5, 3, 1), ("tetra4", 4, 2, 0), ("hexa8", 4, 4, 1), ("quad4", 9, 3, 1), ("tetra10", 2, 2, 1)])
def test_rank_row_blocks_equal_the_single_gpu_rows(pkg, O, meshes, kind, size, world, order_mode):
    from conftest import KINDS
    nodes, conn = meshes.mesh_for(kind, size)
    info = KINDS[kind]
    elemT = getattr(pkg.elements, {"tetra4": "Tetra4", "hexa8": "Hexa8", "quad4": "Quad4", "tetra10": "Tetra10"}[kind])()
    c = pkg.solids.Isotropic(200e9, 0.3).Axisymmetric() if kind == "quad4" else pkg.solids.Isotropic(200e9, 0.3).Solid3D()
    mdpn = bin(info["flags"]).count("1")
    single = pkg.RawAssembler(nodes, mdpn)
    desc = single.make_desc(elemT, c)
    assert single.add(desc, conn=conn) == 0
    ref = single.csr()
    ndof = single.total_dofs()
    blocks, chunk_total, owned_total, halo_total = [], 0, 0, 0
    for r in range(world):
        d = pkg.DistRank(single.ctx, r, world, mdpn, conn.shape[1], nodes=nodes, conn=conn, order_mode=order_mode)
        assert d.add(desc) == 0, d.last_error()
        i = d.info()
        chunk_total += i["chunk_elems"]; owned_total += i["owned_nodes"]; halo_total += i["halo_elems"]
        assert i["global_nodes"] == nodes.shape[0] and i["global_elems"] == conn.shape[0]
        assert i["local_elems"] <= i["chunk_elems"] + i["halo_elems"]
        blocks.append(d.row_block())
        d.close()
    assert chunk_total == conn.shape[0]                       # chunks partition the elements
    referenced = np.unique(conn).size
    assert owned_total == referenced                          # every referenced node has exactly one owner
    indptr, indices, values, seen = _stitch(blocks, ndof)
    ref_rows_nonempty = (ref[0][1:] - ref[0][:-1]) > 0
    assert np.all(seen[ref_rows_nonempty] == 1)
    assert np.array_equal(indptr, ref[0]) and np.array_equal(indices, ref[1])
    assert np.array_equal(values, ref[2]), "row blocks differ from the single-GPU rows (same element order: must be bit-identical)"
    # and the oracle
    Cm, bk = O.constitutive(info["mode"], 200e9, 0.3)
    og = O.GeneralAssembler(nodes, info["flags"])
    og.add_isoparametric(info["kind"], Cm, bk, conn)
    rp, ri, rv, ab = og.csr(with_absum=True)
    compare_csr((indptr, indices, values), (rp, ri, rv), ab)
    single.close()


def test_dist_reports_the_global_element_of_an_error(pkg, meshes):
    nodes, conn = meshes.bcc_tetra4(4)
    bad = conn.copy()
    for e in (500, 41):
        bad[e, [2, 3]] = bad[e, [3, 2]]
    holder = pkg.RawAssembler(nodes, 3)
    desc = holder.make_desc(pkg.elements.Tetra4(), pkg.solids.Isotropic(200e9, 0.3).Solid3D())
    lowest = []
    for r in range(3):
        d = pkg.DistRank(holder.ctx, r, 3, 3, 4, nodes=nodes, conn=bad, order_mode=1)
        rc = d.add(desc)
        if rc:
            k, e, q = d.last_error()
            assert k == 1 and e in (41, 500)
            lowest.append(e)
        d.close()
    assert min(lowest) == 41                                     # the reference's sequential loop would report #41
    holder.close()


def test_group_api_on_one_device_listed_twice(pkg, meshes):
    """femb200_group_*: one process, a host thread per device (here cuda:0 twice) -- what a cgo caller uses."""
    L = pkg.lib()
    nodes, conn = meshes.bcc_tetra4(4)
    single = pkg.RawAssembler(nodes, 3)
    desc = single.make_desc(pkg.elements.Tetra4(), pkg.solids.Isotropic(200e9, 0.3).Solid3D())
    assert single.add(desc, conn=conn) == 0
    ref = single.csr()
    devs = np.array([0, 0], dtype=np.int32)
    g = C.c_void_p()
    rc = L.femb200_group_create(devs.ctypes.data_as(C.POINTER(C.c_int32)), 2, nodes.ctypes.data, nodes.shape[0], 3, conn.ctypes.data, 1, conn.shape[0], 4, 1, C.byref(g))
    assert rc == 0
    assert L.femb200_group_size(g) == 2
    who = C.c_int32(-7)
    assert L.femb200_group_add_isoparametric(g, C.byref(desc), C.byref(who)) == 0 and who.value == -1
    blocks = []
    for r in range(2):
        d = pkg.DistRank(None, r, 2, 3, 4, handle=C.c_void_p(L.femb200_group_rank(g, r)))
        blocks.append(d.row_block())
    indptr, indices, values, _ = _stitch(blocks, single.total_dofs())
    assert np.array_equal(indptr, ref[0]) and np.array_equal(indices, ref[1]) and np.array_equal(values, ref[2])
    L.femb200_group_destroy(g)
    single.close()


@pytest.mark.parametrize("dims", [(3, 3, 3), (5, 2, 4), (1, 1, 1), (2, 7, 1)])
def test_device_mesh_generator_equals_the_numpy_one(pkg, meshes, dims):
    holder = pkg.RawAssembler(np.zeros((1, 3)), 3)
    nx, ny, nz = dims
    h = 1.0 / 7
    nodes, conn = meshes.bcc_tetra4_box(nx, ny, nz, h)
    dn, nn_, dc, ne = pkg.device_mesh_bcc_tetra4(holder.ctx, nx, ny, nz, h)
    assert nn_ == nodes.shape[0] and ne == conn.shape[0]
    got_nodes = np.zeros((nn_, 3))
    got_conn = np.zeros((max(ne, 1), 4), dtype=np.int32)
    L = pkg.lib()
    assert L.femb200_copy_to_host(holder.ctx, got_nodes.ctypes.data, dn, got_nodes.nbytes) == 0
    if ne:
        assert L.femb200_copy_to_host(holder.ctx, got_conn.ctypes.data, dc, ne * 16) == 0
    assert np.array_equal(got_nodes, nodes), "device BCC nodes differ from meshes.bcc_tetra4_box"
    assert np.array_equal(got_conn[:ne], conn.astype(np.int32)), "device BCC connectivity differs from meshes.bcc_tetra4_box"
    pkg.lib().femb200_mesh_free(holder.ctx, dn)
    pkg.lib().femb200_mesh_free(holder.ctx, dc)
    holder.close()


def test_rank_row_block_as_block_pattern(pkg, meshes):
    """femb200_bsr_copy on a rank's assembler: block columns are rank-LOCAL node ids; with the local->global map of
    femb200_dist_local_nodes they reproduce the GLOBAL scalar indices of femb200_csr_copy, values identical."""
    import ctypes as C
    nodes, conn = meshes.mesh_for("tetra4", 6)
    holder = pkg.RawAssembler(nodes, 3)
    desc = holder.make_desc(pkg.elements.Tetra4(), pkg.solids.Isotropic(200e9, 0.3).Solid3D())
    L = pkg.lib()
    for r in range(2):
        d = pkg.DistRank(holder.ctx, r, 2, 3, 4, nodes=nodes, conn=conn, order_mode=1)
        assert d.add(desc) == 0
        ga = d.assembler()
        n_local = d.info()["local_nodes"]
        nnz = L.femb200_nnz(ga)
        nblk = L.femb200_bsr_blocks(ga)
        assert nblk * 9 == nnz
        indptr, indices, values = np.zeros(3 * n_local + 1, np.int64), np.zeros(nnz, np.int32), np.zeros(nnz)
        assert L.femb200_csr_copy(ga, indptr.ctypes.data, indices.ctypes.data, values.ctypes.data) == 0
        blkptr, blkcols, bvals = np.zeros(n_local + 1, np.int64), np.zeros(nblk, np.int32), np.zeros(nnz)
        assert L.femb200_bsr_copy(ga, blkptr.ctypes.data, blkcols.ctypes.data, bvals.ctypes.data) == 0
        l2g, owned = d.local_nodes()
        assert np.array_equal(bvals, values)
        assert np.array_equal(indptr[::3], 9 * blkptr) and np.all((np.diff(blkptr) > 0) <= (owned != 0))
        gcols = (l2g[blkcols].astype(np.int64)[:, None] * 3 + np.arange(3)[None, :]).reshape(-1)
        idx = np.concatenate([np.tile(gcols[3 * blkptr[I]:3 * blkptr[I + 1]], 3) for I in range(n_local)])
        assert np.array_equal(idx.astype(np.int32), indices)
        d.close()
    holder.close()
