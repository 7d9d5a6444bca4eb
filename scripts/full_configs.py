"""BASELINE.json configs at FULL size on one B200: device phase timers + size-independent properties.
  python scripts/full_configs.py [tetra4 quad4 hexa8 tetra10 hexa20] [--json out.json]
Checks (no oracle at these sizes, see tests/ for parity at oracle-sized meshes):
  * nnz equals the closed-form node-pair count of the mesh family times ndn^2
  * bilinear symmetry  x2.(K x1) == x1.(K x2)  for random vectors (K is never transposed or copied to the host)
  * rigid translations are in the null space (3-D solids: x, y, z; axisymmetric: z only)
  * the diagonal of every non-empty row is positive; assembling twice is bitwise identical (sampled via K x)
"""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench

SIZES = {"tetra4": 48, "quad4": 2000, "hexa8": 200, "tetra10": 95, "hexa20": 110}


def expected_blocks(name, s):
    if name == "tetra4":
        return 30 * s ** 3 + 9 * s ** 2 - 15 * s - 23
    if name == "tetra10":
        return 460 * s ** 3 - 192 * s ** 2 - 204 * s - 63
    if name == "hexa8":
        return (3 * s + 1) ** 3
    if name == "quad4":
        return (3 * s + 1) ** 2
    if name == "hexa20":
        n = s
        return 1 + 8 * 3 * n + 47 * 3 * n * n + 234 * n ** 3
    return None


def run(name, size):
    import ctypes as C
    t0 = time.time()
    w = bench.make_workload(name, size)
    pkg, nodes, conn = w["pkg"], w["nodes"], w["conn"]
    ne, nn = conn.shape
    ndn = w["ndn"]
    ndof = nodes.shape[0] * ndn
    t_mesh = time.time() - t0
    L = pkg.lib()
    ra = pkg.RawAssembler(nodes, ndn)
    desc = ra.make_desc(w["elemT"], w["c"])
    conn32 = np.ascontiguousarray(conn.astype(np.int32))
    res = []
    ys = []
    rng = np.random.default_rng(7)
    x1, x2 = rng.standard_normal(ndof), rng.standard_normal(ndof)
    out = {"config": w["desc"], "elements": int(ne), "dofs": int(ndof), "mesh_s": round(t_mesh, 1)}
    for it in range(4):
        h = C.c_void_p()
        rc = L.femb200_assembler_create(ra.ctx, C.c_void_p(nodes.ctypes.data), nodes.shape[0], ndn, 0, C.byref(h))
        assert rc == 0, rc
        rc = L.femb200_add_isoparametric(h, C.byref(desc), C.c_void_p(conn32.ctypes.data), 0, 0, ne)
        assert rc == 0, (rc, L.femb200_last_cuda_error(ra.ctx))
        ph = (C.c_float * pkg.N_TIMERS)()
        L.femb200_phase_ms(h, ph)
        if it >= 1:
            res.append([float(v) for v in ph])
        nnz = L.femb200_nnz(h)
        if it in (2, 3):
            y = np.empty(ndof)
            assert L.femb200_csr_matvec(h, x1.ctypes.data_as(C.POINTER(C.c_double)), y.ctypes.data_as(C.POINTER(C.c_double))) == 0
            ys.append(y)
        if it == 3:
            y2 = np.empty(ndof)
            assert L.femb200_csr_matvec(h, x2.ctypes.data_as(C.POINTER(C.c_double)), y2.ctypes.data_as(C.POINTER(C.c_double))) == 0
            scale = np.abs(ys[0]).max()
            out["symmetry_rel"] = float(abs(x2 @ ys[0] - x1 @ y2) / (np.linalg.norm(x2) * np.linalg.norm(ys[0])))
            axes = [1] if name == "quad4" else list(range(ndn))      # axisymmetric: only z translation is rigid
            worst = 0.0
            for ax in axes:
                u = np.zeros(ndof); u[ax::ndn] = 1.0
                yr = np.empty(ndof)
                assert L.femb200_csr_matvec(h, u.ctypes.data_as(C.POINTER(C.c_double)), yr.ctypes.data_as(C.POINTER(C.c_double))) == 0
                # K u = 0 up to rounding: compare with |K| |u| ~ row abs sums (use K applied to ones of all axes as scale)
                worst = max(worst, float(np.abs(yr).max()))
            out["rigid_translation_residual_over_kx_scale"] = worst / scale
            out["deterministic_bitwise"] = bool(np.array_equal(ys[0], ys[1]))
        L.femb200_assembler_destroy(h)
    m = np.mean(res, axis=0)
    out["phases_ms"] = dict(zip(pkg.TIMER_NAMES, [round(float(v), 4) for v in m]))
    out["nnz"] = int(nnz)
    eb = expected_blocks(name, w["size"])
    out["nnz_expected"] = int(eb * ndn * ndn) if eb else None
    out["nnz_ok"] = (out["nnz"] == out["nnz_expected"]) if eb else None
    dev_ms = float(m[6])
    out["elements_per_s"] = ne / (dev_ms * 1e-3)
    out["dofs_per_s"] = ndof / (dev_ms * 1e-3)
    ra.close()
    return out


if __name__ == "__main__":
    argv = sys.argv[1:]
    jpath = None
    if "--json" in argv:
        jpath = argv[argv.index("--json") + 1]
        del argv[argv.index("--json"):argv.index("--json") + 2]
    args = [a for a in argv if not a.startswith("--")]
    names = args or ["tetra4", "quad4", "hexa8", "tetra10"]
    outs = []
    for n in names:
        nm, _, sz = n.partition(":")
        o = run(nm, int(sz) if sz else SIZES[nm])
        print(json.dumps(o), flush=True)
        outs.append(o)
    if jpath:
        json.dump(outs, open(jpath, "w"), indent=1)
