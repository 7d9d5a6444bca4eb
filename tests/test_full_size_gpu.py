"""Parity at BASELINE.json's FULL sizes for configs 2-5 (config 1 is compared with the oracle entry by entry in
test_gpu_parity.py::test_config1_full_size_properties).  The reference's all-triplets-in-RAM algorithm cannot hold these
(110 GB .. 1.7 TB of COO), so the oracle computes the K ROWS OF A SAMPLE OF NODES from their incident elements
only: every element containing a sampled node, in ascending element index, over a compacted node numbering -- the rows of
the sampled nodes are then exactly the rows the full assembly would produce (same contributions, same order).  The GPU
assembles the whole mesh; its sampled rows are read back row by row from the device CSR and compared: pattern bit-exact
(global column ids), values |a-b| <= 1e-12 * max(|b|, s_ij).  Plus the size-independent checks (closed-form nnz)."""
import ctypes as C

import numpy as np
import pytest

from conftest import KINDS

pytestmark = pytest.mark.gpu
TOL = 1e-12

# name: (mesh kind, size, element class, constituter method, sampled nodes, closed-form node pairs or None)
FULL = {
    "config2_quad4_axisym_2000x2000": ("quad4", 2000, "Quad4", "Axisymmetric", 10000),
    "config3_hexa8_200cubed": ("hexa8", 200, "Hexa8", "Solid3D", 6000),
    "config4_tetra10_bcc95": ("tetra10", 95, "Tetra10", "Solid3D", 3000),
    "config5_share_hexa20_110cubed": ("hexa20", 110, "Hexa20", "Solid3D", 1200),
}


def _device_rows(L, ctx, ga, rows):
    """(indptr of the whole K, {row: (indices, values)}) read from the device CSR without copying K."""
    a, b, c = C.c_void_p(), C.c_void_p(), C.c_void_p()
    L.femb200_csr_device(ga, C.byref(a), C.byref(b), C.byref(c))
    n = L.femb200_total_dofs(ga)
    indptr = np.zeros(n + 1, dtype=np.int64)
    assert L.femb200_copy_to_host(ctx, indptr.ctypes.data, a, indptr.nbytes) == 0
    out = {}
    for r in rows:
        lo, hi = int(indptr[r]), int(indptr[r + 1])
        ix = np.zeros(hi - lo, dtype=np.int32)
        v = np.zeros(hi - lo, dtype=np.float64)
        if hi > lo:
            assert L.femb200_copy_to_host(ctx, ix.ctypes.data, b.value + 4 * lo, ix.nbytes) == 0
            assert L.femb200_copy_to_host(ctx, v.ctypes.data, c.value + 8 * lo, v.nbytes) == 0
        out[int(r)] = (ix, v)
    return indptr, out


def _sampled_rows_oracle(O, info, nodes, conn, sample, mode_args):
    """Oracle rows of the sampled nodes: {global scalar row: (global cols, values, absum)}"""
    sel = np.isin(conn, sample).any(axis=1)
    sub = conn[sel]                                           # ascending element index is kept
    l2g = np.unique(sub)
    lconn = np.searchsorted(l2g, sub).astype(np.int64)
    Cm, bk = O.constitutive(info["mode"], *mode_args)
    og = O.GeneralAssembler(np.ascontiguousarray(nodes[l2g]), info["flags"])
    og.add_isoparametric(info["kind"], Cm, bk, lconn)
    rp, ri, rv, ab = og.csr(with_absum=True)
    mdpn = bin(info["flags"]).count("1")
    rows = {}
    for g in sample:
        ln = int(np.searchsorted(l2g, g))
        for i in range(mdpn):
            lr = ln * mdpn + i
            cols_l = ri[rp[lr]:rp[lr + 1]].astype(np.int64)
            cols_g = l2g[cols_l // mdpn].astype(np.int64) * mdpn + cols_l % mdpn
            rows[int(g) * mdpn + i] = (cols_g.astype(np.int32), rv[rp[lr]:rp[lr + 1]], ab[rp[lr]:rp[lr + 1]])
    return rows, int(sel.sum()), og.last_seconds()


@pytest.mark.parametrize("name", list(FULL))
def test_full_size_sampled_rows_match_the_oracle(pkg, O, meshes, name):
    kind, size, ecls, cmeth, nsample = FULL[name]
    info = KINDS[kind]
    nodes, conn = meshes.mesh_for(kind, size)
    mdpn = bin(info["flags"]).count("1")
    elemT = getattr(pkg.elements, ecls)()
    c = getattr(pkg.solids.Isotropic(200e9, 0.3), cmeth)()
    ra = pkg.RawAssembler(nodes, mdpn)
    desc = ra.make_desc(elemT, c)
    rc = ra.add(desc, conn=conn.astype(np.int32))
    assert rc == 0, ra.last_error()
    L = pkg.lib()
    nnz, ndof = ra.nnz(), ra.total_dofs()
    # closed-form pattern sizes of SURVEY 8(d)
    if kind == "quad4":
        assert conn.shape[0] == 4_000_000 and ndof == 8_008_002 and nnz == 144_048_004
    if kind == "hexa8":
        assert conn.shape[0] == 8_000_000 and ndof == 24_361_803 and nnz == 1_953_736_209
    if kind == "tetra10":
        N = size
        assert conn.shape[0] == 10_180_200 and nodes.shape[0] == 16 * N ** 3 + 6 * N ** 2 - 6 * N - 7
        assert nnz == 9 * (460 * N ** 3 - 192 * N ** 2 - 204 * N - 63) == 3_533_762_313
    if kind == "hexa20":
        n1 = n2 = n3 = size
        s1, s2, s3 = n1 + n2 + n3, n1 * n2 + n1 * n3 + n2 * n3, n1 * n2 * n3
        assert nodes.shape[0] == 1 + 2 * s1 + 3 * s2 + 4 * s3 and nnz == 9 * (1 + 8 * s1 + 47 * s2 + 234 * s3)
    # deterministic sample of referenced nodes: spread over the whole id range plus the first and last ones
    used = np.unique(conn[:: max(1, conn.shape[0] // (4 * nsample))])
    rng = np.random.default_rng(12345)
    sample = np.unique(np.concatenate([rng.choice(used, size=min(nsample, used.size), replace=False), used[:8], used[-8:]]))
    want, nsub, osec = _sampled_rows_oracle(O, info, nodes, conn, sample, (200e9, 0.3))
    indptr, got = _device_rows(L, ra.ctx, ra.h, sorted(want))
    assert indptr[-1] == nnz
    worst = 0.0
    for r, (cols, vals, ab) in want.items():
        gi, gv = got[r]
        assert np.array_equal(gi, cols), f"{name}: pattern of row {r} differs"
        scale = np.maximum(np.abs(vals), ab)
        err = np.abs(gv - vals)
        assert np.all(err <= TOL * scale), f"{name}: row {r}: max normalised error {(err / scale).max():.3e}"
        worst = max(worst, float((err / np.where(scale > 0, scale, 1.0)).max()))
    print(f"{name}: {conn.shape[0]} elements, nnz {nnz}, {len(want)} sampled rows from {nsub} incident elements "
          f"(oracle {osec:.1f} s): pattern exact, max normalised error {worst:.2e}; phases {ra.phase_ms()}")
    ra.close()


def test_config2_full_oracle(pkg, O, meshes):
    """Config 2 entirely through the oracle (4 M Quad4, 256 M triplets): every entry of K."""
    from conftest import compare_csr
    nodes, conn = meshes.mesh_for("quad4", 2000)
    ra = pkg.RawAssembler(nodes, 2)
    desc = ra.make_desc(pkg.elements.Quad4(), pkg.solids.Isotropic(200e9, 0.3).Axisymmetric())
    assert ra.add(desc, conn=conn.astype(np.int32)) == 0
    got = ra.csr()
    ra.close()
    Cm, bk = O.constitutive(1, 200e9, 0.3)
    og = O.GeneralAssembler(nodes, 3)
    og.add_isoparametric(O.QUAD4, Cm, bk, conn)
    rp, ri, rv, ab = og.csr(with_absum=True)
    nerr, rel = compare_csr(got, (rp, ri, rv), ab, TOL)
    print(f"config 2 vs oracle (all {rv.size} entries): max normalised err {nerr:.2e}, max rel err {rel:.2e}, oracle {og.last_seconds():.1f} s")
