"""Pins the CPU oracle against every golden vector / invariant the reference's own tests hold for the
AddIsoparametric path.  CPU only."""
import json
import os

import numpy as np
import pytest

from conftest import go_rand_float64

GOLD = os.path.join(os.path.dirname(__file__), "golden")
EX = json.load(open(os.path.join(GOLD, "example_test_outputs.json")))


def _fmt_equal(name, K):
    fmt, G = EX[name]["format"], np.array(EX[name]["K"])
    F = np.array([[float(fmt % v) for v in row] for row in K])
    assert F.shape == G.shape
    assert np.array_equal(F, G), f"{name}: formatted K differs from the reference's Output block"


# ---- example_test.go:14-149 ---------------------------------------------------------------------
def test_example_quad4_axisymmetric(O):           # example_test.go:14-49
    C, bk = O.constitutive(O.ISO_AXISYM, 1000, 0.33)
    ga = O.GeneralAssembler([[20, 0, 0], [30, 0, 0], [30, 1, 0], [20, 1, 0]], 3)
    ga.add_isoparametric(O.QUAD4, C, bk, [[0, 1, 2, 3]])
    _fmt_equal("ExampleGeneralAssembler_AddIsoparametric_axisymmetric", ga.dense())


def test_example_tetra4_solid3d(O):               # example_test.go:52-86 (pins weight=1, exact zeros)
    C, bk = O.constitutive(O.ISO_SOLID3D, 200e9, 0.3)
    ga = O.GeneralAssembler([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]], 7)
    ga.add_isoparametric(O.TETRA4, C, bk, [[0, 1, 2, 3]])
    _fmt_equal("ExampleGeneralAssembler", ga.dense())


def test_example_quad4_plane_stress(O):           # example_test.go:88-120
    C, bk = O.constitutive(O.ISO_PLANESTRESS, 1, 0.3)
    ga = O.GeneralAssembler([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0]], 3)
    ga.add_isoparametric(O.QUAD4, C, bk, [[0, 1, 2, 3]])
    _fmt_equal("ExampleGeneralAssembler_AddIsoparametric_quad4PlaneStress", ga.dense())


def test_example_quad4_axisymmetric_thermal(O):   # example_test.go:122-149 (NodeDofs override, 1-dof model)
    C, bk = O.constitutive(O.COND_AXISYM, 45)
    ga = O.GeneralAssembler([[20, 0, 0], [30, 0, 0], [30, 1, 0], [20, 1, 0]], 1)
    ga.add_isoparametric(O.QUAD4, C, bk, [[0, 1, 2, 3]], node_dofs=1)
    _fmt_equal("ExampleGeneralAssembler_AddIsoparametric_axisymmetricThermal", ga.dense())


# ---- elements/elements3d_test.go:27-130, elements2d_test.go:27-126 --------------------------------
ELEMS3 = ["TETRA4", "TETRA10", "HEXA8", "HEXA20"]
ELEMS2 = ["QUAD8", "QUAD4", "TRIANGLE3"]


@pytest.mark.parametrize("name", ELEMS3 + ELEMS2)
def test_basis_kronecker(O, name):               # TestBasis3d / TestBasis2, tol 1e-16
    k = getattr(O, name)
    for i, n in enumerate(O.iso_nodes(k)):
        ff = O.basis(k, n)
        for j, v in enumerate(ff):
            assert abs(v - (1.0 if i == j else 0.0)) <= 1e-16


@pytest.mark.parametrize("name", ELEMS3 + ELEMS2)
def test_basis_diff_central_difference(O, name):  # TestBasisDiff3d / 2d: h=1e-4, tol 1e-11, 100 points
    k = getattr(O, name)
    nn = O.len_nodes(k)
    dims = 3 if name in ELEMS3 else 2
    h = 1e-4
    rng = np.random.default_rng(1)
    for _ in range(100):
        p = np.zeros(3)
        p[:dims] = rng.random(dims)
        got = O.basis_diff(k, p)
        for ax in range(dims):
            e = np.zeros(3)
            e[ax] = h / 2
            fd = (O.basis(k, p + e) - O.basis(k, p - e)) * (1 / h)
            g = got[ax * nn:(ax + 1) * nn]      # floats.EqualApprox = within 1e-11 absolute OR relative
            diff = np.abs(g - fd)
            assert np.all((diff <= 1e-11) | (diff <= 1e-11 * np.maximum(np.abs(g), np.abs(fd))))


@pytest.mark.parametrize("name", ELEMS3 + ELEMS2 + ["TRIANGLE6"])
def test_quadrature_sums(O, name):                # TestQuadrature3d / 2d: sum w = sum w*sum N = volume()
    k = getattr(O, name)
    pos, w = O.quadrature(k)
    assert len(pos) == len(w)
    vol = O.measure(k)
    assert abs(w.sum() - vol) <= 1e-11
    assert abs(sum(wi * O.basis(k, p).sum() for p, wi in zip(pos, w)) - vol) <= 1e-11


def test_quirks(O):
    # 8-Q1: tet weights are 6x textbook; 8-Q4: gaussQuad1D(4) has a == b; 8-Q5: Triangle6 unusable
    assert O.quadrature(O.TETRA4)[1].tolist() == [1.0]
    assert O.quadrature(O.TETRA10)[1].tolist() == [0.25] * 4
    pos, _ = O.quadrature(O.HEXA8, 4)
    xs = np.unique(np.round(pos[:, 0], 15))
    assert xs.size == 2 and np.allclose(np.abs(xs), 0.5216121828372476)
    assert O.basis_diff(O.TRIANGLE6, [0.1, 0.2, 0]).size == 8
    C, bk = O.constitutive(O.ISO_PLANESTRESS, 1, 0.3)
    ga = O.GeneralAssembler(np.zeros((6, 3)), 3)
    with pytest.raises(O.OracleError) as ei:
        ga.add_isoparametric(O.TRIANGLE6, C, bk, [[0, 1, 2, 3, 4, 5]])
    assert ei.value.kind == 10
    # a2: mapdofs loops to popcount(model) -> non-contiguous flags give a non-injective map
    assert O.mapdofs(0b111, 0b111) == [0, 1, 2]
    assert O.mapdofs(0b101, 0b101) == [0, 0]
    # fem.go:138 + assembler.go:148-150: elem dofs/node must equal model dofs/node ("bad length")
    C6, b6 = O.constitutive(O.ISO_SOLID3D, 1, 0.3)
    ga = O.GeneralAssembler(np.eye(4, 3), 63)
    with pytest.raises(O.OracleError) as ei:
        ga.add_isoparametric(O.TETRA4, C6, b6, [[0, 1, 2, 3]])
    assert ei.value.kind == 7


# ---- constitution/solids/patch_test.go ------------------------------------------------------------
def _washer():
    ri, lx, ly = 1, 0.24, 0.12
    r = go_rand_float64(8)
    nodes = [[ri, 0, 0], [ri + lx, 0, 0], [ri + lx, ly, 0], [ri, ly, 0],
             [ri + lx / 3 + (r[0] - 0.5) * lx / 4, ly / 3 + (r[1] - 0.5) * ly / 4, 0],
             [ri + lx * 2 / 3 - (r[2] - 0.5) * lx / 4, ly / 3 + (r[3] - 0.5) * ly / 4, 0],
             [ri + lx * 2 / 3 - (r[4] - 0.5) * lx / 4, ly * 2 / 3 - (r[5] - 0.5) * ly / 4, 0],
             [ri + lx / 3 + (r[6] - 0.5) * lx / 4, ly * 2 / 3 - (r[7] - 0.5) * ly / 4, 0]]
    elems = [[0, 1, 5, 4], [1, 2, 6, 5], [6, 2, 3, 7], [0, 4, 7, 3], [4, 5, 6, 7]]
    return np.array(nodes), np.array(elems), ri, lx, ly


def solve_patch(K, nodes, ndn, fixed_mask, u):
    n = K.shape[0]
    fixed = np.repeat(fixed_mask, ndn)
    free = ~fixed
    rhs = -K[np.ix_(free, fixed)] @ u[fixed]
    u = u.copy()
    u[free] = np.linalg.solve(K[np.ix_(free, free)], rhs)
    return u


def test_axisymmetric_patch(O):                   # patch_test.go:19-157
    nodes, elems, ri, lx, ly = _washer()
    C, bk = O.constitutive(O.ISO_AXISYM, 1e6, 0.25)
    ga = O.GeneralAssembler(nodes, 3)
    ga.add_isoparametric(O.QUAD4, C, bk, elems)
    K = ga.dense()
    ext = (nodes[:, 1] == ly) | (nodes[:, 1] == 0) | (nodes[:, 0] == ri) | (nodes[:, 0] == ri + lx)
    # case 1
    u = np.zeros(16)
    u[0::2] = np.where(ext, nodes[:, 0] / 1000, 0)
    u[1::2] = np.where(ext, (nodes[:, 1] + nodes[:, 0]) / 1000, 0)
    u = solve_patch(K, nodes, 2, ext, u)
    s = ga.strains(u, O.QUAD4, C, bk, elems)
    assert np.all(np.abs(s - 1e-3) / 1e-3 <= 3e-2)
    # case 2
    u = np.zeros(16)
    u[0::2] = np.where(ext, nodes[:, 0] / 100, 0)
    u[1::2] = np.where(ext, nodes[:, 1] / 100, 0)
    u = solve_patch(K, nodes, 2, ext, u)
    s = ga.strains(u, O.QUAD4, C, bk, elems)
    assert np.all(np.abs(s[..., :3] - 9.95e-3) / 9.95e-3 <= 3e-2)
    assert np.all(np.abs(s[..., 3]) <= 1e-8)


def test_axisymmetric_thermal_patch(O):           # patch_test.go:162-247
    nodes, elems, ri, lx, ly = _washer()
    C, bk = O.constitutive(O.COND_AXISYM, 1.0)
    ga = O.GeneralAssembler(nodes, 1)
    ga.add_isoparametric(O.QUAD4, C, bk, elems, node_dofs=1)
    K = ga.dense()
    top, bot = nodes[:, 1] == ly, nodes[:, 1] == 0
    T = np.where(top, 100.0, 0.0)
    T = solve_patch(K, nodes, 1, top | bot, T)
    exp = 100 * nodes[:, 1] / ly
    assert np.all(np.abs(T - exp) / (exp + 1e-2) <= 1e-2)


# ---- fixture meshes of the reference's examples -----------------------------------------------------
def test_ruc_hexa8_is_inverted_and_fixable(O):
    """examples/ruc ships Hexa8 connectivity with negative Jacobians: the reference's own main() must hit
    assembler_isoparametric.go:95-96 on element #0; swapping the two faces gives a valid mesh."""
    m = json.load(open(os.path.join(GOLD, "ruc_hexa8.json")))
    nodes, h8 = np.array(m["nodes"]), np.array(m["hexa8"])
    C, bk = O.constitutive(O.ISO_SOLID3D, 4.8e3, 0.34)
    ga = O.GeneralAssembler(nodes, 7)
    with pytest.raises(O.OracleError) as ei:
        ga.add_isoparametric(O.HEXA8, C, bk, h8)
    assert (ei.value.kind, ei.value.elem, ei.value.qp) == (1, 0, 0)
    assert ga.csr()[2].size == 0                    # K untouched on error (:118-121)
    ga.add_isoparametric(O.HEXA8, C, bk, h8[:, [4, 5, 6, 7, 0, 1, 2, 3]])
    K = ga.dense()
    assert np.allclose(K, K.T, rtol=0, atol=1e-9 * np.abs(K).max())
    for ax in range(3):                              # rigid translations are in the null space
        u = np.zeros(K.shape[0])
        u[ax::3] = 1
        assert np.abs(K @ u).max() <= 1e-9 * np.abs(K).max()


def test_bimetallic_quad8(O):
    m = json.load(open(os.path.join(GOLD, "bimetallic_quad8.json")))
    nodes, q8 = np.array(m["nodes"]), np.array(m["quad8"])
    C, bk = O.constitutive(O.ISO_PLANESTRAIN, 200e3, 0.3)
    ga = O.GeneralAssembler(nodes, 3)
    ga.add_isoparametric(O.QUAD8, C, bk, q8)
    K = ga.dense()
    assert np.allclose(K, K.T, rtol=0, atol=1e-9 * np.abs(K).max())
    for ax in range(2):
        u = np.zeros(K.shape[0])
        u[ax::2] = 1
        assert np.abs(K @ u).max() <= 1e-9 * np.abs(K).max()


def test_accumulate_two_calls_and_bcc_counts(O, meshes):
    # benchmark_test.go:28-51 (dofs, elems) pairs of README.md:82-86
    for N, dofs, elems in ((2, 105, 48), (8, 3723, 5376), (16, 27027, 46080)):
        n, c = meshes.bcc_tetra4(N)
        assert (n.shape[0] * 3, c.shape[0]) == (dofs, elems)
    n, c = meshes.bcc_tetra4(4)
    C, bk = O.constitutive(O.ISO_SOLID3D, 200e9, 0.3)
    a = O.GeneralAssembler(n, 7)
    a.add_isoparametric(O.TETRA4, C, bk, c)
    b = O.GeneralAssembler(n, 7)
    h = c.shape[0] // 2
    b.add_isoparametric(O.TETRA4, C, bk, c[:h])
    b.add_isoparametric(O.TETRA4, C, bk, c[h:])
    pa, ia, va = a.csr()
    pb, ib, vb = b.csr()
    assert np.array_equal(pa, pb) and np.array_equal(ia, ib)
    assert np.allclose(va, vb, rtol=1e-13, atol=1e-3)
    nblk = 30 * 4 ** 3 + 9 * 4 ** 2 - 15 * 4 - 23   # SURVEY 8d node-pair count
    assert va.size == 9 * nblk
    assert int((np.diff(pa) == 0).sum()) == 24      # the 8 unreferenced cube corners
