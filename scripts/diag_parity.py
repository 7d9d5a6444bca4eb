"""Diagnostic: per element type, where do GPU and oracle differ?  python scripts/diag_parity.py [names...]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
pkg, O = g.load_package(), g.load_oracle()
from importlib import import_module
M = import_module("gofem_b200.meshes")
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from conftest import KINDS
from test_gpu_parity import elem_of, constituter_of

names = sys.argv[1:] or ["tetra4", "tetra10", "hexa8", "hexa20", "quad4", "quad8", "triangle3"]
sizes = dict(tetra4=6, tetra10=3, hexa8=5, hexa20=3, quad4=17, quad8=9, triangle3=13)
for name in names:
    nodes, conn = M.mesh_for(name, sizes[name])
    k = KINDS[name]
    ga = pkg.fem.NewGeneralAssembler(nodes, k["flags"])
    err = ga.AddIsoparametric(elem_of(pkg, name), constituter_of(pkg, k["mode"], 200e9, 0.3), conn.shape[0], conn)
    print(name, "err:", err, ga.phase_ms())
    gp, gi, gv = ga.Ksolid().csr()
    C, bk = O.constitutive(k["mode"], 200e9, 0.3)
    og = O.GeneralAssembler(nodes, k["flags"])
    og.add_isoparametric(k["kind"], C, bk, conn)
    rp, ri, rv, ab = og.csr(True)
    print("  pattern equal:", np.array_equal(gp, rp) and np.array_equal(gi, ri), "nnz", gv.size)
    scale = np.maximum(np.abs(rv), ab)
    nerr = np.abs(gv - rv) / np.where(scale > 0, scale, 1)
    bad = np.nonzero(nerr > 1e-12)[0]
    print("  bad entries:", bad.size, "max nerr", nerr.max())
    if bad.size:
        rows = np.searchsorted(rp, bad, side="right") - 1
        ndn = bin(k["flags"]).count("1")
        print("  bad rows (node ids):", np.unique(rows // ndn)[:20], "count", np.unique(rows // ndn).size, "of", nodes.shape[0])
        for b in bad[:8]:
            r = np.searchsorted(rp, b, side="right") - 1
            print(f"    row {r} col {gi[b]} gpu {gv[b]:.6e} ref {rv[b]:.6e} absum {ab[b]:.3e} rowlen {rp[r+1]-rp[r]}")
        rl = np.diff(rp)[::ndn] // ndn
        print("  row block lengths: max", rl.max(), "bad rows' lengths:", np.unique(rl[np.unique(rows // ndn)]))
