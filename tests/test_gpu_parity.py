"""Parity tests proper: the CUDA path (through the host mirror and the C ABI) against the CPU oracle on the
same inputs.  Bar: CSR pattern (indptr/indices) bit-exact, values within 1e-12 normalised per SURVEY 8-Q8
(|a-b| <= 1e-12*max(|b|, sum_e|Ke_e[i][j]|)).  Run on a B200: python -m pytest tests -m gpu."""
import json
import os

import numpy as np
import pytest

from conftest import KINDS, compare_csr, go_rand_float64

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
EX = json.load(open(os.path.join(GOLD, "example_test_outputs.json")))
TOL = 1e-12


def elem_of(pkg, name, **kw):
    return {"tetra4": pkg.elements.Tetra4, "tetra10": pkg.elements.Tetra10, "hexa8": pkg.elements.Hexa8, "hexa20": pkg.elements.Hexa20,
            "quad4": pkg.elements.Quad4, "quad8": pkg.elements.Quad8, "triangle3": pkg.elements.Triangle3}[name](**kw)


def constituter_of(pkg, mode, p0, p1=0.0):
    s = pkg.solids
    return [s.Isotropic(p0, p1).Solid3D, s.Isotropic(p0, p1).Axisymmetric, s.Isotropic(p0, p1).PlaneStess, s.Isotropic(p0, p1).PlaneStrain,
            s.IsotropicConductivity(p0).Solid3D, s.IsotropicConductivity(p0).Plane, s.IsotropicConductivity(p0).Axisymmetric][mode]()


def assemble_both(pkg, O, name, nodes, conn, mode, flags, p0=200e9, p1=0.3, quad_order=0, node_dofs=0):
    ga = pkg.fem.NewGeneralAssembler(nodes, flags)
    err = ga.AddIsoparametric(elem_of(pkg, name, QuadratureOrder=quad_order, NodeDofs=node_dofs), constituter_of(pkg, mode, p0, p1), conn.shape[0], conn)
    assert err is None, err
    C, bk = O.constitutive(mode, p0, p1)
    og = O.GeneralAssembler(nodes, flags)
    og.add_isoparametric(KINDS[name]["kind"], C, bk, conn, quad_order=quad_order, node_dofs=node_dofs)
    rp, ri, rv, ab = og.csr(with_absum=True)
    return ga, ga.Ksolid().csr(), (rp, ri, rv), ab


# ---- the reference's golden examples, through the GPU ------------------------------------------------
def _fmt_equal(name, K):
    fmt, G = EX[name]["format"], np.array(EX[name]["K"])
    F = np.array([[float(fmt % v) for v in row] for row in K])
    assert np.array_equal(F, G), name


def test_golden_examples_on_gpu(pkg):
    fem, el, so = pkg.fem, pkg.elements, pkg.solids
    q = [[20, 0, 0], [30, 0, 0], [30, 1, 0], [20, 1, 0]]
    ga = fem.NewGeneralAssembler(q, el.Quad4().Dofs())
    assert ga.AddIsoparametric(el.Quad4(), so.Isotropic(1000, 0.33).Axisymmetric(), 1, lambda i: [0, 1, 2, 3]) is None
    _fmt_equal("ExampleGeneralAssembler_AddIsoparametric_axisymmetric", ga.Ksolid().dense())
    ga = fem.NewGeneralAssembler([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]], el.Tetra4().Dofs())
    assert ga.AddIsoparametric(el.Tetra4(), so.Isotropic(200e9, 0.3).Solid3D(), 1, lambda i: [0, 1, 2, 3]) is None
    _fmt_equal("ExampleGeneralAssembler", ga.Ksolid().dense())
    ga = fem.NewGeneralAssembler([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0]], el.Quad4().Dofs())
    assert ga.AddIsoparametric(el.Quad4(), so.Isotropic(1, 0.3).PlaneStess(), 1, lambda i: [0, 1, 2, 3]) is None
    _fmt_equal("ExampleGeneralAssembler_AddIsoparametric_quad4PlaneStress", ga.Ksolid().dense())
    et = el.Quad4(NodeDofs=fem.DofPosX)
    ga = fem.NewGeneralAssembler(q, et.Dofs())
    assert ga.AddIsoparametric(et, so.IsotropicConductivity(45).Axisymmetric(), 1, lambda i: [0, 1, 2, 3]) is None
    _fmt_equal("ExampleGeneralAssembler_AddIsoparametric_axisymmetricThermal", ga.Ksolid().dense())


# ---- every element type of the north star against the oracle ------------------------------------------
@pytest.mark.parametrize("name,size", [("tetra4", 6), ("tetra10", 3), ("hexa8", 5), ("hexa20", 3), ("quad4", 17), ("quad8", 9), ("triangle3", 13)])
def test_element_types_match_oracle(pkg, O, meshes, name, size):
    nodes, conn = meshes.mesh_for(name, size)
    k = KINDS[name]
    ga, got, ref, ab = assemble_both(pkg, O, name, nodes, conn, k["mode"], k["flags"])
    nerr, rel = compare_csr(got, ref, ab, TOL)
    print(f"{name}: nelem={conn.shape[0]} nnz={got[2].size} max normalised err={nerr:.2e} max rel err={rel:.2e} phases={ga.phase_ms()}")


@pytest.mark.parametrize("name,mode,flags", [("quad4", 2, 3), ("quad8", 3, 3), ("triangle3", 2, 3), ("quad4", 3, 3)])
def test_plane_kinds_match_oracle(pkg, O, meshes, name, mode, flags):
    nodes, conn = meshes.mesh_for(name, 8)
    _, got, ref, ab = assemble_both(pkg, O, name, nodes, conn, mode, flags, 70e9, 0.25)
    compare_csr(got, ref, ab, TOL)


@pytest.mark.parametrize("name,mode,size", [("hexa8", 4, 4), ("tetra10", 4, 2), ("hexa20", 4, 2), ("quad4", 5, 9),
                                            ("quad8", 6, 6), ("triangle3", 6, 7), ("quad4", 6, 9)])
def test_thermal_kinds_match_oracle(pkg, O, meshes, name, mode, size):
    nodes, conn = meshes.mesh_for(name, size)
    _, got, ref, ab = assemble_both(pkg, O, name, nodes, conn, mode, 1, 45.0, 0.0, node_dofs=1)
    compare_csr(got, ref, ab, TOL)


def test_tetra4_cannot_be_thermal(pkg, O, meshes):
    """Tetra4.Dofs() ignores NodeDofs (tetra4.go:18-20), so a 1-dof thermal model is rejected by mapdofs."""
    nodes, conn = meshes.bcc_tetra4(2)
    ga = pkg.fem.NewGeneralAssembler(nodes, 1)
    err = ga.AddIsoparametric(pkg.elements.Tetra4(NodeDofs=1), pkg.solids.IsotropicConductivity(45.0).Solid3D(), len(conn), conn)
    assert err.msg == "model dofs 1 does not contain all element dofs 123"
    og = O.GeneralAssembler(nodes, 1)
    Cm, bk = O.constitutive(4, 45.0)
    with pytest.raises(O.OracleError) as ei:
        og.add_isoparametric(O.TETRA4, Cm, bk, conn, node_dofs=1)
    assert ei.value.kind == 5


@pytest.mark.parametrize("name,order", [("hexa8", 1), ("hexa8", 3), ("hexa8", 4), ("quad4", 1), ("quad4", 5), ("quad8", 2), ("triangle3", 3), ("hexa20", 2)])
def test_quadrature_orders(pkg, O, meshes, name, order):
    nodes, conn = meshes.mesh_for(name, 3)
    k = KINDS[name]
    _, got, ref, ab = assemble_both(pkg, O, name, nodes, conn, k["mode"], k["flags"], quad_order=order)
    compare_csr(got, ref, ab, TOL)


def test_global_accumulator_path_matches(pkg, O, meshes, monkeypatch):
    """Rows too long for shared memory fall back to accumulating straight into the CSR slots."""
    monkeypatch.setenv("FEMB200_FORCE_GLOBAL_ACC", "1")
    for name, size in (("tetra4", 4), ("hexa8", 3), ("quad4", 9)):
        nodes, conn = meshes.mesh_for(name, size)
        k = KINDS[name]
        _, got, ref, ab = assemble_both(pkg, O, name, nodes, conn, k["mode"], k["flags"])
        compare_csr(got, ref, ab, TOL)


@pytest.mark.parametrize("env", ["FEMB200_FORCE_THREAD_ROWS", "FEMB200_NO_HASH_PATTERN"])
def test_alternative_kernel_paths_match(pkg, O, meshes, monkeypatch, env):
    """The thread-per-row numeric kernel (used for nn > 8) and the sorted-insert pattern kernel (overflow path)
    stay parity-checked on the low-order elements too."""
    monkeypatch.setenv(env, "1")
    for name, size in (("tetra4", 5), ("hexa8", 3), ("quad4", 9), ("triangle3", 6), ("quad8", 4)):
        nodes, conn = meshes.mesh_for(name, size)
        k = KINDS[name]
        _, got, ref, ab = assemble_both(pkg, O, name, nodes, conn, k["mode"], k["flags"])
        compare_csr(got, ref, ab, TOL)


def test_collapsed_element_with_positive_volume_neighbours(pkg, O, meshes):
    """Quad4 with two coincident node ids (a triangle-shaped quad) is legal: detJ stays > 0 at the Gauss points.
    Two local nodes then map to the same column block; the group kernel serialises those lanes."""
    nodes, conn = meshes.quad4_grid(6, 6, jittered=False)
    conn = conn.copy()
    conn[7, 3] = conn[7, 2]          # node 3 collapsed onto node 2
    conn[20, 1] = conn[20, 0]
    ga = pkg.fem.NewGeneralAssembler(nodes, 3)
    err = ga.AddIsoparametric(pkg.elements.Quad4(), pkg.solids.Isotropic(1e5, 0.3).PlaneStess(), len(conn), conn)
    og = O.GeneralAssembler(nodes, 3)
    Cm, bk = O.constitutive(2, 1e5, 0.3)
    try:
        og.add_isoparametric(O.QUAD4, Cm, bk, conn)
    except O.OracleError as oe:
        assert err is not None and f"#{oe.elem}" in err.msg
        return
    assert err is None, err
    rp, ri, rv, ab = og.csr(with_absum=True)
    compare_csr(ga.Ksolid().csr(), (rp, ri, rv), ab, TOL)


def test_high_valence_node_long_rows(pkg, O):
    """A fan of 300 triangles around one hub node: the hub row has 301 blocks (spills the shared list)."""
    n = 300
    ang = np.linspace(0, 2 * np.pi, n, endpoint=False)
    nodes = np.zeros((n + 1, 3))
    nodes[1:, 0] = 5 + np.cos(ang)
    nodes[1:, 1] = 5 + np.sin(ang)
    nodes[0, :2] = 5
    conn = np.array([[0, 1 + i, 1 + (i + 1) % n] for i in range(n)], dtype=np.int64)
    _, got, ref, ab = assemble_both(pkg, O, "triangle3", nodes, conn, 1, 3, 1e5, 0.3)
    compare_csr(got, ref, ab, TOL)
    assert np.diff(got[0]).max() == 2 * (n + 1)


def test_fixture_meshes(pkg, O):
    m = json.load(open(os.path.join(GOLD, "ruc_hexa8.json")))
    nodes, h8 = np.array(m["nodes"]), np.array(m["hexa8"], dtype=np.int64)
    ga = pkg.fem.NewGeneralAssembler(nodes, 7)
    err = ga.AddIsoparametric(pkg.elements.Hexa8(), pkg.solids.Isotropic(4.8e3, 0.34).Solid3D(), len(h8), h8)
    assert err is not None and err.msg == "negative determinant of jacobian of element #0, Check node ordering"
    assert ga.Ksolid().NonZero() == 0                       # K untouched (assembler_isoparametric.go:118-121)
    fixed = h8[:, [4, 5, 6, 7, 0, 1, 2, 3]]
    _, got, ref, ab = assemble_both(pkg, O, "hexa8", nodes, fixed, 0, 7, 4.8e3, 0.34)
    compare_csr(got, ref, ab, TOL)
    m = json.load(open(os.path.join(GOLD, "bimetallic_quad8.json")))
    nodes, q8 = np.array(m["nodes"]), np.array(m["quad8"], dtype=np.int64)
    _, got, ref, ab = assemble_both(pkg, O, "quad8", nodes, q8, 3, 3, 200e3, 0.3)
    compare_csr(got, ref, ab, TOL)
    shifted = nodes.copy()
    shifted[:, 0] += 200
    _, got, ref, ab = assemble_both(pkg, O, "quad8", shifted, q8, 1, 3, 200e3, 0.3)
    compare_csr(got, ref, ab, TOL)


# ---- error behaviour -------------------------------------------------------------------------------

@pytest.mark.parametrize("elem", ["tetra4", "hexa8"])
def test_two_faults_follow_the_sequential_loop(pkg, meshes, elem):
    """fem.go:140 walks the elements in order: with an inverted element and an out-of-range node in ONE mesh the
    reference reports whichever comes first (the Jacobian error of #3 before the slice panic of #10, and vice versa).
    Tetra4 takes the fused path (checks on the second stream), Hexa8 the staged one (checks over the leading elements)."""
    if elem == "tetra4":
        nodes, conn = meshes.bcc_tetra4(3)
        et, swap = pkg.elements.Tetra4(), [2, 3]
    else:
        nodes, conn = meshes.hexa8_grid(4, 4, 4)
        et, swap = pkg.elements.Hexa8(), [0, 6]
    c = pkg.solids.Isotropic(200e9, 0.3).Solid3D()
    a = conn.copy()
    a[3, swap] = a[3, swap[::-1]]
    a[10, 1] = nodes.shape[0]
    err = pkg.fem.NewGeneralAssembler(nodes, 7).AddIsoparametric(et, c, len(a), a)
    assert err is not None and not err.is_panic and "element #3" in err.msg
    b = conn.copy()
    b[3, 1] = nodes.shape[0] + 5
    b[10, swap] = b[10, swap[::-1]]
    err = pkg.fem.NewGeneralAssembler(nodes, 7).AddIsoparametric(et, c, len(b), b)
    assert err is not None and err.is_panic and "#3" in err.msg
    # the same element both inverted and out of range: the gather panics before the Jacobian exists
    d = conn.copy()
    d[7, swap] = d[7, swap[::-1]]
    d[7, 2] = -1
    err = pkg.fem.NewGeneralAssembler(nodes, 7).AddIsoparametric(et, c, len(d), d)
    assert err is not None and err.is_panic and "#7" in err.msg

def test_errors_report_lowest_element_and_leave_k_untouched(pkg, O, meshes):
    nodes, conn = meshes.bcc_tetra4(4)
    bad = conn.copy()
    for e in (500, 37, 212):                                # swapping two vertices flips the Jacobian sign
        bad[e, [2, 3]] = bad[e, [3, 2]]
    ga = pkg.fem.NewGeneralAssembler(nodes, 7)
    c = pkg.solids.Isotropic(200e9, 0.3).Solid3D()
    assert ga.AddIsoparametric(pkg.elements.Tetra4(), c, len(conn) // 2, conn[: len(conn) // 2]) is None
    before = ga.Ksolid().csr()
    err = ga.AddIsoparametric(pkg.elements.Tetra4(), c, len(bad), bad)
    assert err.msg == "negative determinant of jacobian of element #37, Check node ordering"
    after = ga.Ksolid().csr()
    assert all(np.array_equal(a, b) for a, b in zip(before, after))
    og = O.GeneralAssembler(nodes, 7)
    Cm, bk = O.constitutive(0, 200e9, 0.3)
    with pytest.raises(O.OracleError) as ei:
        og.add_isoparametric(O.TETRA4, Cm, bk, bad)
    assert (ei.value.kind, ei.value.elem) == (1, 37)
    # degenerate element -> "zero determinant" (dJac < 1e-12 absolute)
    flat = conn.copy()
    flat[11, 3] = flat[11, 2]
    err = ga.AddIsoparametric(pkg.elements.Tetra4(), c, len(flat), flat)
    with pytest.raises(O.OracleError) as ei:
        og.add_isoparametric(O.TETRA4, Cm, bk, flat)
    # two identical Jacobian rows: LU gives exactly 0 or +-O(eps) depending on x*(1/x)==1 -- the GPU follows the
    # reference's LU (unfused), so it lands on the same side as the oracle
    assert (ei.value.kind, ei.value.elem) in ((2, 11), (1, 11))
    expect = {2: "zero determinant of jacobian of element #11, Check element shape for bad aspect ratio",
              1: "negative determinant of jacobian of element #11, Check node ordering"}[ei.value.kind]
    assert err.msg == expect
    # tiny but valid elements trip the ABSOLUTE 1e-12 threshold too (8-Q3)
    tiny = pkg.fem.NewGeneralAssembler(nodes * 1e-5, 7)
    err = tiny.AddIsoparametric(pkg.elements.Tetra4(), c, len(conn), conn)
    assert err is not None and err.msg.startswith("zero determinant of jacobian of element #0")
    # node index out of range -> the reference panics on the slice bound
    oob = conn.copy()
    oob[77, 1] = nodes.shape[0]
    err = ga.AddIsoparametric(pkg.elements.Tetra4(), c, len(oob), oob)
    assert err is not None and err.is_panic and "#77" in err.msg
    # axisymmetric element on the axis: r = 0 -> 1/r = Inf, scale stays finite (0) => no NaN-scale error,
    # but NaN coordinates give a NaN radius -> ":106 NaN scale value"
    qn, qc = meshes.quad4_grid(3, 3)
    qn[5, 0] = np.nan
    g2 = pkg.fem.NewGeneralAssembler(qn, 3)
    err = g2.AddIsoparametric(pkg.elements.Quad4(), pkg.solids.Isotropic(1e3, 0.3).Axisymmetric(), len(qc), qc)
    og2 = O.GeneralAssembler(qn, 3)
    Ca, bka = O.constitutive(1, 1e3, 0.3)
    with pytest.raises(O.OracleError) as ei:
        og2.add_isoparametric(O.QUAD4, Ca, bka, qc)
    assert err is not None
    if ei.value.kind == 4:
        assert err.msg == f"NaN scale value returned by SetStrainDisplacementMatrix at element #{ei.value.elem}, quad {ei.value.qp}"


def test_host_side_errors(pkg, meshes):
    nodes, conn = meshes.bcc_tetra4(2)
    fem, el, so = pkg.fem, pkg.elements, pkg.solids
    c = so.Isotropic(200e9, 0.3).Solid3D()
    ga = fem.NewGeneralAssembler(nodes, 7)
    err = ga.AddIsoparametric(el.Tetra4(), c, 3, lambda i: [0, 1, 2, 3] if i != 1 else [0, 1, 2])
    assert err.msg == "element #1 of 3 nodes expected to be of 4 nodes"                       # fem.go:143
    err = ga.AddIsoparametric(el.Tetra4(), c, len(conn), conn, orientation=[1, 0, 0])
    assert err.msg == "arbitrary constitutive orientation not implemented yet"                # :87
    err = fem.NewGeneralAssembler(nodes, fem.DofPosX | fem.DofPosY).AddIsoparametric(el.Tetra4(), c, len(conn), conn)
    assert err.msg == "model dofs 12 does not contain all element dofs 123"                   # fem.go:158
    err = fem.NewGeneralAssembler(nodes, fem.Dof6).AddIsoparametric(el.Tetra4(), c, len(conn), conn)
    assert err.is_panic and err.msg == "bad length"                                           # assembler.go:149
    err = ga.AddIsoparametric(el.Hexa8(QuadratureOrder=7), c, 0, np.zeros((0, 8)))
    assert err.msg == "bad quadrature result from isoparametric element"
    err = fem.NewGeneralAssembler(nodes, 3).AddIsoparametric(el.Triangle6(), so.Isotropic(1, 0.3).PlaneStess(), 0, np.zeros((0, 6)))
    assert err.is_panic                                                                        # 8-Q5
    err = ga.AddIsoparametric(el.Tetra4(), so.Isotropic(1.0, 0.5).Solid3D(), len(conn), conn)
    assert "condition number" in err.msg                                                       # isotropic.go:24-26
    assert ga.Ksolid().NonZero() == 0


# ---- assembler semantics ---------------------------------------------------------------------------
def test_repeated_calls_accumulate_and_mix_element_types(pkg, O, meshes):
    nodes, conn = meshes.hexa8_grid(4, 4, 4)
    ga = pkg.fem.NewGeneralAssembler(nodes, 7)
    og = O.GeneralAssembler(nodes, 7)
    Cm, bk = O.constitutive(0, 200e9, 0.3)
    h = conn.shape[0] // 3
    c1, c2 = pkg.solids.Isotropic(200e9, 0.3).Solid3D(), pkg.solids.Isotropic(70e9, 0.33).Solid3D()
    C2, _ = O.constitutive(0, 70e9, 0.33)
    assert ga.AddIsoparametric(pkg.elements.Hexa8(), c1, h, conn[:h]) is None
    og.add_isoparametric(O.HEXA8, Cm, bk, conn[:h])
    assert ga.AddIsoparametric(pkg.elements.Hexa8(), c2, conn.shape[0] - h, conn[h:]) is None
    og.add_isoparametric(O.HEXA8, C2, bk, conn[h:])
    # a third call with a different element type on the same nodes (tets over the first cells)
    tets = np.array([[conn[e, 0], conn[e, 1], conn[e, 3], conn[e, 4]] for e in range(10)], dtype=np.int64)
    assert ga.AddIsoparametric(pkg.elements.Tetra4(), c1, len(tets), tets) is None
    og.add_isoparametric(O.TETRA4, Cm, bk, tets)
    rp, ri, rv, ab = og.csr(with_absum=True)
    compare_csr(ga.Ksolid().csr(), (rp, ri, rv), ab, TOL)


def test_deterministic_bitwise(pkg, meshes):
    nodes, conn = meshes.bcc_tetra4(8)
    c = pkg.solids.Isotropic(200e9, 0.3).Solid3D()
    runs = []
    for _ in range(3):
        ga = pkg.fem.NewGeneralAssembler(nodes, 7)
        assert ga.AddIsoparametric(pkg.elements.Tetra4(), c, conn.shape[0], conn) is None
        runs.append(ga.Ksolid().csr())
    for r in runs[1:]:
        assert all(np.array_equal(a, b) for a, b in zip(runs[0], r))       # bit-identical values


def test_empty_and_degenerate_inputs(pkg):
    fem, el, so = pkg.fem, pkg.elements, pkg.solids
    nodes = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1.0], [9, 9, 9]])
    ga = fem.NewGeneralAssembler(nodes, 7)
    assert ga.TotalDofs() == 15
    assert ga.AddIsoparametric(el.Tetra4(), so.Isotropic(1, 0.3).Solid3D(), 0, np.zeros((0, 4))) is None
    ip, ix, v = ga.Ksolid().csr()
    assert ip.tolist() == [0] * 16 and ix.size == 0
    assert ga.AddIsoparametric(el.Tetra4(), so.Isotropic(1, 0.3).Solid3D(), 1, lambda i: [0, 1, 2, 3]) is None
    ip, ix, v = ga.Ksolid().csr()
    assert ip[12] == 144 and ip[15] == 144          # node 4 unreferenced -> three empty rows
    K = ga.Ksolid()
    assert K.At(0, 0) == v[0] and K.At(14, 14) == 0.0 and K.Dims() == (15, 15)
    assert fem.NewGeneralAssembler(np.zeros((0, 3)), 7).TotalDofs() == 0


def test_noncontiguous_model_dofs_quirk(pkg, O, meshes):
    """fem.go:155-172 maps DofPosX|DofPosZ to dofmap [0,0]: both element dofs land on the same global dof.
    The drop-in reproduces it (rows of the second dof stay empty)."""
    nodes, conn = meshes.quad4_grid(4, 4)
    flags = pkg.fem.DofPosX | pkg.fem.DofPosZ
    ga = pkg.fem.NewGeneralAssembler(nodes, flags)
    et = pkg.elements.Quad4(NodeDofs=flags)
    assert ga.AddIsoparametric(et, pkg.solids.Isotropic(1e3, 0.3).PlaneStess(), len(conn), conn) is None
    og = O.GeneralAssembler(nodes, flags)
    Cm, bk = O.constitutive(2, 1e3, 0.3)
    og.add_isoparametric(O.QUAD4, Cm, bk, conn, node_dofs=flags)
    rp, ri, rv, ab = og.csr(with_absum=True)
    compare_csr(ga.Ksolid().csr(), (rp, ri, rv), ab, TOL)
    assert np.all(np.diff(rp)[1::2] == 0)


def test_slot_map(pkg, meshes):
    nodes, conn = meshes.hexa8_grid(3, 3, 3)
    ga = pkg.fem.NewGeneralAssembler(nodes, 7)
    assert ga.AddIsoparametric(pkg.elements.Hexa8(), pkg.solids.Isotropic(1e3, 0.3).Solid3D(), len(conn), conn) is None
    indptr, indices, _ = ga.Ksolid().csr()
    slot = ga.slot_map(len(conn), 8)
    blkptr = indptr[::3] // 9
    for e in (0, 13, 26):
        for a in range(8):
            for b in range(8):
                s = slot[e, a, b]
                row = conn[e, a]
                assert blkptr[row] <= s < blkptr[row + 1]
                pos = s - blkptr[row]
                assert indices[indptr[3 * row] + 3 * pos] == 3 * conn[e, b]


def test_matvec_and_rigid_body_modes(pkg, meshes):
    nodes, conn = meshes.bcc_tetra4(6)
    ga = pkg.fem.NewGeneralAssembler(nodes, 7)
    assert ga.AddIsoparametric(pkg.elements.Tetra4(), pkg.solids.Isotropic(200e9, 0.3).Solid3D(), len(conn), conn) is None
    K = ga.Ksolid()
    ip, ix, v = K.csr()
    import scipy.sparse as sp
    A = sp.csr_matrix((v, ix, ip), shape=K.Dims())
    x = np.random.default_rng(0).standard_normal(A.shape[0])
    assert np.allclose(K.MulVec(x), A @ x, rtol=1e-12, atol=1e-12 * np.abs(v).max())
    for ax in range(3):
        u = np.zeros(A.shape[0])
        u[ax::3] = 1
        assert np.abs(K.MulVec(u)).max() <= 1e-9 * np.abs(v).max()


# ---- patch tests of the reference through the GPU -----------------------------------------------------
def _washer():
    ri, lx, ly = 1, 0.24, 0.12
    r = go_rand_float64(8)
    nodes = [[ri, 0, 0], [ri + lx, 0, 0], [ri + lx, ly, 0], [ri, ly, 0],
             [ri + lx / 3 + (r[0] - 0.5) * lx / 4, ly / 3 + (r[1] - 0.5) * ly / 4, 0],
             [ri + lx * 2 / 3 - (r[2] - 0.5) * lx / 4, ly / 3 + (r[3] - 0.5) * ly / 4, 0],
             [ri + lx * 2 / 3 - (r[4] - 0.5) * lx / 4, ly * 2 / 3 - (r[5] - 0.5) * ly / 4, 0],
             [ri + lx / 3 + (r[6] - 0.5) * lx / 4, ly * 2 / 3 - (r[7] - 0.5) * ly / 4, 0]]
    elems = [[0, 1, 5, 4], [1, 2, 6, 5], [6, 2, 3, 7], [0, 4, 7, 3], [4, 5, 6, 7]]
    return np.array(nodes), np.array(elems, dtype=np.int64), ri, lx, ly


def test_axisymmetric_patch_test_on_gpu(pkg, O):            # patch_test.go:19-157
    nodes, elems, ri, lx, ly = _washer()
    fem = pkg.fem
    mat = pkg.solids.Isotropic(1e6, 0.25).Axisymmetric()
    et = pkg.elements.Quad4()
    ga = fem.NewGeneralAssembler(nodes, et.Dofs())
    assert ga.AddIsoparametric(et, mat, len(elems), lambda i: elems[i]) is None
    K = ga.Ksolid().dense()
    fix = fem.NewFixity(fem.DofPosX | fem.DofPosY, len(nodes))
    u = np.zeros(16)
    for i, n in enumerate(nodes):
        if n[1] == ly or n[1] == 0 or n[0] == ri or n[0] == ri + lx:
            u[2 * i], u[2 * i + 1] = n[0] / 1000, (n[1] + n[0]) / 1000
            fix.Fix(i, fem.DofPosX | fem.DofPosY)
    free, fixed = fix.FreeDofs(), fix.FixedDofs()
    u[free] = np.linalg.solve(K[np.ix_(free, free)], -K[np.ix_(free, fixed)] @ u[fixed])
    s = ga.IsoparametricStrains(u, et, mat, elems)
    assert np.all(np.abs(s - 1e-3) / 1e-3 <= 3e-2)
    Cm, bk = O.constitutive(1, 1e6, 0.25)
    og = O.GeneralAssembler(nodes, 3)
    og.add_isoparametric(O.QUAD4, Cm, bk, elems)
    assert np.allclose(s, og.strains(u, O.QUAD4, Cm, bk, elems), rtol=1e-11, atol=1e-15)


def test_thermal_patch_test_on_gpu(pkg):                    # patch_test.go:162-247
    nodes, elems, ri, lx, ly = _washer()
    fem = pkg.fem
    et = pkg.elements.Quad4(NodeDofs=1)
    ga = fem.NewGeneralAssembler(nodes, 1)
    assert ga.AddIsoparametric(et, pkg.solids.IsotropicConductivity(1.0).Axisymmetric(), len(elems), lambda i: elems[i]) is None
    K = ga.Ksolid().dense()
    fix = fem.NewFixity(fem.DofPosX, len(nodes))
    T = np.zeros(8)
    for i, n in enumerate(nodes):
        if n[1] == ly:
            T[i] = 100
            fix.Fix(i, 1)
        elif n[1] == 0:
            fix.Fix(i, 1)
    free, fixed = fix.FreeDofs(), fix.FixedDofs()
    T[free] = np.linalg.solve(K[np.ix_(free, free)], -K[np.ix_(free, fixed)] @ T[fixed])
    exp = 100 * nodes[:, 1] / ly
    assert np.all(np.abs(T - exp) / (exp + 1e-2) <= 1e-2)


@pytest.mark.parametrize("name,size,mode,flags", [("hexa8", 4, 0, 7), ("tetra10", 2, 0, 7), ("quad8", 5, 3, 3), ("triangle3", 6, 1, 3)])
def test_strains_match_oracle(pkg, O, meshes, name, size, mode, flags):
    nodes, conn = meshes.mesh_for(name, size)
    ga = pkg.fem.NewGeneralAssembler(nodes, flags)
    u = np.random.default_rng(3).standard_normal(ga.TotalDofs()) * 1e-3
    s = ga.IsoparametricStrains(u, elem_of(pkg, name), constituter_of(pkg, mode, 1e5, 0.3), conn)
    Cm, bk = O.constitutive(mode, 1e5, 0.3)
    og = O.GeneralAssembler(nodes, flags)
    ref = og.strains(u, KINDS[name]["kind"], Cm, bk, conn)
    assert s.shape == ref.shape
    assert np.allclose(s, ref, rtol=1e-11, atol=1e-13 * np.abs(ref).max())


# ---- numeric-only re-assembly on the cached pattern (femb200_reassemble) -----------------------------
@pytest.mark.parametrize("name,size,mode", [("hexa8", 5, 0), ("tetra4", 6, 0), ("quad4", 12, 1), ("tetra10", 3, 0)])
def test_reassemble_numeric_only(pkg, O, meshes, name, size, mode):
    nodes, conn = meshes.mesh_for(name, size)
    flags = KINDS[name]["flags"]
    mdpn = bin(flags).count("1")
    ra = pkg.RawAssembler(nodes, mdpn)
    d1 = ra.make_desc(elem_of(pkg, name), constituter_of(pkg, mode, 200e9, 0.3))
    assert ra.add(d1, conn=conn) == 0
    ip, ix, v1 = ra.csr()
    # doubled Young modulus: every C entry doubles exactly => K doubles bitwise; pattern untouched
    d2 = ra.make_desc(elem_of(pkg, name), constituter_of(pkg, mode, 400e9, 0.3))
    assert ra.reassemble(d2) == 0
    ip2, ix2, v2 = ra.csr()
    assert np.array_equal(ip, ip2) and np.array_equal(ix, ix2) and np.array_equal(v2, 2.0 * v1)
    assert ra.phase_ms()["adjacency"] == 0.0 and ra.phase_ms()["k_rows"] > 0.0      # no symbolic phase ran
    # moved nodes: bitwise equal to a fresh assembly on the moved mesh
    rng = np.random.default_rng(3)
    nodes2 = nodes + 1e-3 * rng.standard_normal(nodes.shape) * (np.abs(nodes).max() / size)
    if nodes.shape[1] == 3 and name in ("quad4",):
        nodes2[:, 2] = nodes[:, 2]
    assert ra.reassemble(d1, nodes=nodes2) == 0
    _, _, v3 = ra.csr()
    fresh = pkg.RawAssembler(nodes2, mdpn)
    assert fresh.add(fresh.make_desc(elem_of(pkg, name), constituter_of(pkg, mode, 200e9, 0.3)), conn=conn) == 0
    fp, fi, fv = fresh.csr()
    assert np.array_equal(fp, ip) and np.array_equal(fi, ix) and np.array_equal(fv, v3)
    # an element error leaves K untouched (assembler_isoparametric.go:118-120)
    bad = nodes2.copy()
    e0 = conn[conn.shape[0] // 2]
    bad[e0[0]], bad[e0[1]] = nodes2[e0[1]].copy(), nodes2[e0[0]].copy()          # swap two nodes: inverted element(s)
    rc = ra.reassemble(d1, nodes=bad)
    assert rc in (pkg.ERR_NEG_DET, pkg.ERR_ZERO_DET) if hasattr(pkg, "ERR_NEG_DET") else rc in (1, 2)
    _, _, v4 = ra.csr()
    assert np.array_equal(v4, v3)
    # a second AddIsoparametric on the same assembler invalidates the cached plan for re-assembly
    assert ra.add(d1, conn=conn) == 0
    assert ra.reassemble(d1) == 20
    ra.close(); fresh.close()


# ---- multi-GPU building blocks exercised on one GPU -------------------------------------------------------------------
def _dev(t):
    import ctypes as C
    return C.c_void_p(t.data_ptr())


def test_pack_rows_and_add_entries_async(pkg, meshes):
    """femb200_csr_pack_rows_async / femb200_csr_add_entries_async (the planned interface exchange): packing rows of K and
    adding them back doubles exactly those rows; a plan that contradicts K is counted in the mismatch word."""
    import torch
    nodes, conn = meshes.mesh_for("tetra4", 5)
    ra = pkg.RawAssembler(nodes, 3)
    desc = ra.make_desc(pkg.elements.Tetra4(), pkg.solids.Isotropic(200e9, 0.3).Solid3D())
    assert ra.add(desc, conn=conn) == 0
    ip, ix, v0 = ra.csr()
    dev = torch.device("cuda", 0)
    rows = np.array([3, 4, 5, 60, 61, 62, 200], dtype=np.int64)
    lens = ip[rows + 1] - ip[rows]
    ptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    t_rows, t_ptr = torch.from_numpy(rows).to(dev), torch.from_numpy(ptr).to(dev)
    oi = torch.zeros(int(ptr[-1]), dtype=torch.int32, device=dev)
    ov = torch.zeros(int(ptr[-1]), dtype=torch.float64, device=dev)
    mism = torch.zeros(1, dtype=torch.int64, device=dev)
    L = pkg.lib()
    assert L.femb200_csr_pack_rows_async(ra.h, rows.size, _dev(t_rows), _dev(t_ptr), _dev(oi), _dev(ov), _dev(mism)) == 0
    assert L.femb200_sync(ra.ctx) == 0
    for w, r in enumerate(rows):
        assert np.array_equal(oi[ptr[w]:ptr[w + 1]].cpu().numpy(), ix[ip[r]:ip[r + 1]])
        assert np.array_equal(ov[ptr[w]:ptr[w + 1]].cpu().numpy(), v0[ip[r]:ip[r + 1]])
    assert L.femb200_csr_add_entries_async(ra.h, rows.size, _dev(t_rows), _dev(t_ptr), _dev(oi), _dev(ov), _dev(mism)) == 0
    assert L.femb200_sync(ra.ctx) == 0
    _, _, v1 = ra.csr()
    expect = v0.copy()
    for r in rows:
        expect[ip[r]:ip[r + 1]] *= 2.0
    assert np.array_equal(v1, expect) and int(mism.item()) == 0
    # a plan with a wrong row length, and an entry outside the pattern, are both counted
    bad_ptr = ptr + np.arange(ptr.size, dtype=np.int64)            # every row one entry longer than K holds
    t_bad = torch.from_numpy(bad_ptr).to(dev)
    oi2 = torch.zeros(int(bad_ptr[-1]), dtype=torch.int32, device=dev)
    ov2 = torch.zeros(int(bad_ptr[-1]), dtype=torch.float64, device=dev)
    assert L.femb200_csr_pack_rows_async(ra.h, rows.size, _dev(t_rows), _dev(t_bad), _dev(oi2), _dev(ov2), _dev(mism)) == 0
    assert L.femb200_sync(ra.ctx) == 0
    assert int(mism.item()) == rows.size
    mism.zero_()
    oi[0] = int(ix.max()) + 7            # a column that no row holds
    assert L.femb200_csr_add_entries_async(ra.h, rows.size, _dev(t_rows), _dev(t_ptr), _dev(oi), _dev(ov), _dev(mism)) == 0
    assert L.femb200_sync(ra.ctx) == 0
    assert int(mism.item()) == 1
    ra.close()


def test_column_map_and_node_mask(pkg, O, meshes):
    """Rank-local numbering: an assembler over a subset of the nodes (ascending global ids) with a column map produces the
    corresponding rows of the global K with GLOBAL column ids; a node mask restricts the assembled rows."""
    import torch
    nodes, conn = meshes.mesh_for("hexa8", 4)
    Cm, bk = O.constitutive(0, 200e9, 0.3)
    og = O.GeneralAssembler(nodes, 7)
    og.add_isoparametric(O.HEXA8 if hasattr(O, "HEXA8") else KINDS["hexa8"]["kind"], Cm, bk, conn)
    rp, ri, rv, ab = og.csr(with_absum=True)
    sel = np.arange(conn.shape[0] // 3, conn.shape[0])               # a chunk of the elements
    l2g = np.unique(conn[sel])
    lconn = np.searchsorted(l2g, conn[sel]).astype(np.int32)
    # rows whose node is touched ONLY by the selected elements are complete in the local K
    others = np.setdiff1d(np.arange(conn.shape[0]), sel)
    complete = ~np.isin(l2g, np.unique(conn[others]))
    dev = torch.device("cuda", 0)
    ra = pkg.RawAssembler(np.ascontiguousarray(nodes[l2g]), 3)
    d_l2g = torch.from_numpy(l2g.astype(np.int32)).to(dev)
    mask = torch.from_numpy(complete.astype(np.uint8)).to(dev)
    L = pkg.lib()
    assert L.femb200_assembler_set_column_map_device(ra.h, _dev(d_l2g), nodes.shape[0]) == 0
    L.femb200_assembler_set_node_mask_device(ra.h, _dev(mask))
    desc = ra.make_desc(pkg.elements.Hexa8(), pkg.solids.Isotropic(200e9, 0.3).Solid3D())
    assert ra.add(desc, conn=lconn) == 0
    ip, ix, v = ra.csr()
    assert complete.any() and (~complete).any()
    for ln_, gnode in enumerate(l2g):
        for k in range(3):
            lo, hi = ip[3 * ln_ + k], ip[3 * ln_ + k + 1]
            if not complete[ln_]:
                assert hi == lo                                               # masked out
                continue
            g = 3 * gnode + k
            assert np.array_equal(ix[lo:hi], ri[rp[g]:rp[g + 1]])              # global column ids, bit-exact pattern
            assert np.all(np.abs(v[lo:hi] - rv[rp[g]:rp[g + 1]]) <= TOL * np.maximum(np.abs(rv[rp[g]:rp[g + 1]]), ab[rp[g]:rp[g + 1]]))
    ra.close()


def test_csr_slice_free_fixed(pkg, meshes):
    """K(free,free) / K(free,fixed) extraction on the device (lap.Slice in patch_test.go:80,87 with Fixity's dof lists)."""
    import scipy.sparse as sp
    nodes, conn = meshes.mesh_for("hexa8", 5)
    ga = pkg.fem.NewGeneralAssembler(nodes, 7)
    assert ga.AddIsoparametricFlat(pkg.elements.Hexa8(), pkg.solids.Isotropic(200e9, 0.3).Solid3D(), conn) is None
    K = ga.Ksolid()
    ip, ix, v = K.csr()
    n = ip.size - 1
    A = sp.csr_matrix((v, ix, ip), shape=(n, n))
    fix = pkg.fem.NewFixity(7, nodes.shape[0])
    for node in np.nonzero(nodes[:, 0] < 1e-9)[0]:
        fix.Fix(int(node), 7)
    fix.Fix(int(nodes.shape[0] - 1), 2)
    free, fixed = np.array(fix.FreeDofs()), np.array(fix.FixedDofs())
    assert free.size + fixed.size == n and fixed.size > 0
    for rows, cols in ((free, free), (free, fixed), (fixed[::-1].copy(), free), (free[:0], free), (free, fixed[:0])):
        p, i, x = K.Slice(rows, cols)
        ref = A[rows][:, cols].tocsr()
        ref.sort_indices()
        assert np.array_equal(p, ref.indptr) and np.array_equal(i, ref.indices) and np.array_equal(x, ref.data)
    with pytest.raises(RuntimeError):
        K.Slice(free, fixed[::-1].copy())          # columns must be ascending


# ---- BASELINE config 1 at full size: size-independent properties ----------------------------------------
def test_config1_full_size_properties(pkg, O, meshes):
    """Tetra4 BCC N=48: 1,299,456 elements / 684,723 dofs (benchmark_test.go largest case).  The oracle is
    run on the same mesh too (tens of seconds): pattern bit-exact and values to 1e-12."""
    N = 48
    nodes, conn = meshes.bcc_tetra4(N)
    assert (nodes.shape[0] * 3, conn.shape[0]) == (684723, 1299456)
    ga = pkg.fem.NewGeneralAssembler(nodes, 7)
    assert ga.AddIsoparametricFlat(pkg.elements.Tetra4(), pkg.solids.Isotropic(200e9, 0.3).Solid3D(), conn) is None
    ip, ix, v = ga.Ksolid().csr()
    nblk = 30 * N ** 3 + 9 * N ** 2 - 15 * N - 23
    assert v.size == 9 * nblk == 30039777
    assert int((np.diff(ip) == 0).sum()) == 24
    assert np.all(np.diff(ix.astype(np.int64))[np.setdiff1d(np.arange(ix.size - 1), ip[1:-1] - 1)] > 0)   # ascending in rows
    import scipy.sparse as sp
    A = sp.csr_matrix((v, ix, ip), shape=(684723, 684723))
    scale = np.abs(v).max()
    assert abs(A - A.T).max() <= 1e-12 * scale                             # symmetry
    for ax in range(3):                                                     # rigid translations
        u = np.zeros(684723)
        u[ax::3] = 1
        assert np.abs(A @ u).max() <= 1e-9 * scale
    d = A.diagonal()
    assert np.all(d[np.diff(ip) > 0] > 0)
    print("config1 phases (ms):", ga.phase_ms())
    Cm, bk = O.constitutive(0, 200e9, 0.3)
    og = O.GeneralAssembler(nodes, 7)
    og.add_isoparametric(O.TETRA4, Cm, bk, conn)
    rp, ri, rv, ab = og.csr(with_absum=True)
    nerr, rel = compare_csr((ip, ix, v), (rp, ri, rv), ab, TOL)
    print(f"config1 vs oracle: max normalised err {nerr:.2e}, max rel err {rel:.2e}, oracle {og.last_seconds():.1f}s")


# ---- additive boundary entries: block-CSR hand-off, foreign triplets ------------------------------------------
@pytest.mark.parametrize("kind,size", [("tetra4", 4), ("quad4", 7)])
def test_bsr_copy_reproduces_the_scalar_csr(pkg, meshes, kind, size):
    """femb200_bsr_copy ships the block pattern (13 MB for config 1) instead of the scalar indices (120 MB): the scalar CSR
    rebuilt from (blkptr, blkcols) on the host must equal femb200_csr_copy bit for bit."""
    nodes, conn = meshes.mesh_for(kind, size)
    if kind == "tetra4":
        elemT, c, mdpn = pkg.elements.Tetra4(), pkg.solids.Isotropic(200e9, 0.3).Solid3D(), 3
    else:
        elemT, c, mdpn = pkg.elements.Quad4(), pkg.solids.Isotropic(200e9, 0.3).Axisymmetric(), 2
    ra = pkg.RawAssembler(nodes, mdpn)
    assert ra.add(ra.make_desc(elemT, c), conn=conn) == 0
    indptr, indices, values = ra.csr()
    blkptr, blkcols, bvals = ra.bsr()
    assert np.array_equal(bvals, values)
    n_nodes = nodes.shape[0]
    rl = np.diff(blkptr)
    ip = np.zeros(n_nodes * mdpn + 1, dtype=np.int64)
    ip[1:] = np.cumsum(np.repeat(rl * mdpn, mdpn))
    assert np.array_equal(ip, indptr)
    cols = (blkcols[:, None] * mdpn + np.arange(mdpn)[None, :]).reshape(-1)          # one block row of column ids per block
    idx = np.concatenate([np.tile(cols[mdpn * blkptr[I]:mdpn * blkptr[I + 1]], mdpn) for I in range(n_nodes)]) if n_nodes else cols
    assert np.array_equal(idx.astype(np.int32), indices)
    # a second call merges another pattern into K: the block view is no longer valid and says so
    assert ra.add(ra.make_desc(elemT, c), conn=conn[: len(conn) // 2]) == 0
    with pytest.raises(RuntimeError):
        ra.bsr()
    ra.close()


def test_add_coo_mixes_foreign_triplets_into_k(pkg, meshes):
    """assembler.go:225-231: AddElement3 adds its beam blocks into the same ksolid.  femb200_add_coo takes such triplets
    (any order, duplicates, entries outside the solid pattern) and K becomes the pattern union, checked against scipy."""
    import scipy.sparse as sp
    nodes, conn = meshes.bcc_tetra4(3)
    ra = pkg.RawAssembler(nodes, 3)
    assert ra.add(ra.make_desc(pkg.elements.Tetra4(), pkg.solids.Isotropic(200e9, 0.3).Solid3D()), conn=conn) == 0
    n = ra.total_dofs()
    K0 = sp.csr_matrix((ra.csr()[2], ra.csr()[1], ra.csr()[0]), shape=(n, n))
    rng = np.random.default_rng(5)
    m = 500
    rows = rng.integers(0, n, m)
    cols = rng.integers(0, n, m)
    rows[:50], cols[:50] = rows[50:100], cols[50:100]                                   # duplicates
    vals = rng.standard_normal(m) * 1e9
    assert ra.add_coo(rows, cols, vals) == 0
    indptr, indices, values = ra.csr()
    K1 = sp.csr_matrix((values, indices, indptr), shape=(n, n))
    want = (K0 + sp.coo_matrix((vals, (rows, cols)), shape=(n, n)).tocsr()).tocsr()
    assert abs(K1 - want).max() <= 1e-6 * abs(want).max()
    assert all(np.all(np.diff(indices[indptr[r]:indptr[r + 1]]) > 0) for r in range(0, n, 7))  # canonical: ascending, unique
    assert K1.nnz >= K0.nnz
    # the cached plan no longer matches K: numeric-only re-assembly must refuse
    assert ra.reassemble(ra.make_desc(pkg.elements.Tetra4(), pkg.solids.Isotropic(100e9, 0.3).Solid3D())) != 0
    # out-of-range triplet
    assert ra.add_coo([n], [0], [1.0]) != 0
    # triplets into an EMPTY assembler (beams only)
    rb = pkg.RawAssembler(nodes, 3)
    assert rb.add_coo(rows, cols, vals) == 0
    ip2, ix2, v2 = rb.csr()
    K2 = sp.csr_matrix((v2, ix2, ip2), shape=(n, n))
    assert abs(K2 - sp.coo_matrix((vals, (rows, cols)), shape=(n, n)).tocsr()).max() <= 1e-6 * abs(vals).max()
    ra.close(); rb.close()


def test_keep_rows_then_reassemble_is_refused(pkg, meshes):
    """after femb200_csr_keep_rows_device the cached plan addresses slots that no longer exist."""
    import torch
    nodes, conn = meshes.bcc_tetra4(3)
    ra = pkg.RawAssembler(nodes, 3)
    desc = ra.make_desc(pkg.elements.Tetra4(), pkg.solids.Isotropic(200e9, 0.3).Solid3D())
    assert ra.add(desc, conn=conn) == 0
    assert ra.reassemble(desc) == 0
    keep = torch.ones(ra.total_dofs(), dtype=torch.uint8, device="cuda")
    keep[::2] = 0
    assert pkg.lib().femb200_csr_keep_rows_device(ra.h, keep.data_ptr()) == 0
    assert ra.reassemble(desc) != 0
    with pytest.raises(RuntimeError):
        ra.bsr()
    ra.close()


# ---- paths the structured meshes never reach: long rows, many incidences per node, hint retry -------------------------
def _fan_triangles(nfan, r0=50.0):
    """A fan of `nfan` triangles around one centre node plus a second ring: the centre row has nfan + 1 node blocks and
    nfan incident elements (long-row copy-out of the fused kernel, > 32 incidences: sorted-insert pattern fallback)."""
    ang = 2 * np.pi * np.arange(nfan) / nfan
    ring1 = np.stack([r0 + 1.0 * np.cos(ang), 30 + 1.0 * np.sin(ang), np.zeros(nfan)], axis=1)
    ring2 = np.stack([r0 + 2.0 * np.cos(ang + 0.1), 30 + 2.0 * np.sin(ang + 0.1), np.zeros(nfan)], axis=1)
    nodes = np.concatenate([[[r0, 30.0, 0.0]], ring1, ring2], axis=0)
    tris = []
    for k in range(nfan):
        a, b = 1 + k, 1 + (k + 1) % nfan
        tris.append([0, a, b])                                     # inner fan (counter-clockwise)
        c, d = 1 + nfan + k, 1 + nfan + (k + 1) % nfan
        tris.append([a, c, b])
        tris.append([b, c, d])
    return nodes, np.array(tris, dtype=np.int64)


def _cone_tets(nfan):
    """`nfan` tetrahedra around a vertical axis (two apex nodes shared by all of them): the apex rows have nfan + 2 blocks."""
    ang = 2 * np.pi * np.arange(nfan) / nfan
    ring = np.stack([np.cos(ang), np.sin(ang), 0.1 * np.cos(3 * ang)], axis=1)
    nodes = np.concatenate([[[0, 0, -1.0], [0, 0, 1.0]], ring], axis=0)
    tets = [[0, 1, 2 + k, 2 + (k + 1) % nfan] for k in range(nfan)]
    nodes, tets = np.asarray(nodes, dtype=np.float64), np.array(tets, dtype=np.int64)
    p = nodes[tets]
    neg = np.linalg.det(p[:, 1:4, :] - p[:, 0:1, :]) < 0
    tets[neg, 2], tets[neg, 3] = tets[neg, 3].copy(), tets[neg, 2].copy()
    return nodes, tets


@pytest.mark.parametrize("case", ["tri_fan_20_plane", "tri_fan_40_axisym", "tri_fan_70_plane", "tet_cone_12", "tet_cone_30", "tet_cone_45"])
def test_long_rows_and_high_valence(pkg, O, case):
    kind, n = case.split("_")[0], int(case.split("_")[2])
    if kind == "tri":
        nodes, conn = _fan_triangles(n)
        axis = case.endswith("axisym")
        elemT, okind = pkg.elements.Triangle3(), O.TRIANGLE3
        c = pkg.solids.Isotropic(210e3, 0.3).Axisymmetric() if axis else pkg.solids.Isotropic(210e3, 0.3).PlaneStrain()
        flags, mode = 3, (1 if axis else 3)
        Cm, bk = O.constitutive(mode, 210e3, 0.3)
    else:
        nodes, conn = _cone_tets(n)
        elemT, okind, c, flags = pkg.elements.Tetra4(), O.TETRA4, pkg.solids.Isotropic(200e9, 0.3).Solid3D(), 7
        Cm, bk = O.constitutive(0, 200e9, 0.3)
    ga = pkg.fem.NewGeneralAssembler(nodes, flags)
    err = ga.AddIsoparametric(elemT, c, len(conn), conn)
    assert err is None, err.msg
    got = ga.Ksolid().csr()
    og = O.GeneralAssembler(nodes, flags)
    og.add_isoparametric(okind, Cm, bk, conn)
    rp, ri, rv, ab = og.csr(with_absum=True)
    compare_csr(got, (rp, ri, rv), ab, TOL)
    mdpn = bin(flags).count("1")
    longest = int(np.diff(rp).max()) // mdpn
    assert longest >= (n + 1 if kind == "tri" else n + 2)


def test_size_hint_of_another_mesh_is_caught_and_the_call_repeated(pkg, O):
    """The context sizes a call from its previous call with the same (nodes, elements, nn) signature.  Two different meshes
    with that signature but different row lengths / block counts must both come out right: every kernel guards its writes
    against the hinted capacities and a tripped guard repeats the call without the hint."""
    # mesh A: 40 triangles in two separate fans of 20; mesh B: one fan of 40 -- same node and element counts
    def two_fans():
        na, ca = _fan_triangles(20)
        keep = ca[::3]                                             # inner fans only
        nb = na.copy(); nb[:, 0] += 10.0
        nodes = np.concatenate([na, nb], axis=0)
        return nodes, np.concatenate([keep, keep + na.shape[0]], axis=0)

    def one_fan():
        n1, c1 = _fan_triangles(40)
        nodes = np.concatenate([n1, n1[:1] + 5.0], axis=0)         # pad to the same node count with an unreferenced node
        return nodes, c1[::3]

    (nA, cA), (nB, cB) = two_fans(), one_fan()
    assert nA.shape == nB.shape and cA.shape == cB.shape
    c = pkg.solids.Isotropic(210e3, 0.3).PlaneStrain()
    Cm, bk = O.constitutive(3, 210e3, 0.3)
    ra = pkg.RawAssembler(nA, 2)                                   # ONE context for all calls: the hint persists
    desc = ra.make_desc(pkg.elements.Triangle3(), c)
    L = pkg.lib()
    import ctypes as C
    for nodes, conn in ((nA, cA), (nA, cA), (nB, cB), (nB, cB), (nA, cA)):
        h = C.c_void_p()
        nodes = np.ascontiguousarray(nodes)
        assert L.femb200_assembler_create(ra.ctx, nodes.ctypes.data, nodes.shape[0], 2, 0, C.byref(h)) == 0
        c32 = np.ascontiguousarray(conn, dtype=np.int32)
        assert L.femb200_add_isoparametric(h, C.byref(desc), c32.ctypes.data, 0, 0, conn.shape[0]) == 0
        n, nnz = L.femb200_total_dofs(h), L.femb200_nnz(h)
        ip, ix, v = np.zeros(n + 1, dtype=np.int64), np.zeros(nnz, dtype=np.int32), np.zeros(nnz)
        assert L.femb200_csr_copy(h, ip.ctypes.data, ix.ctypes.data, v.ctypes.data) == 0
        L.femb200_assembler_destroy(h)
        og = O.GeneralAssembler(nodes, 3)
        og.add_isoparametric(O.TRIANGLE3, Cm, bk, conn)
        rp, ri, rv, ab = og.csr(with_absum=True)
        compare_csr((ip, ix, v), (rp, ri, rv), ab, TOL)
    ra.close()


# ---- the element-block path of Hexa8 / Tetra10 / Hexa20 (csrc/keblocks.cu) ---------------------------------------------
class _GeneralSolid:
    """An IsoConstituter with a user-supplied 6 x 6 matrix and the XYZ strain-displacement filler."""

    def __init__(self, Cm):
        self.Cm = np.ascontiguousarray(Cm, dtype=np.float64)

    def Constitutive(self):
        return self.Cm

    def BKind(self):
        return 0


@pytest.mark.parametrize("name,size", [("hexa8", 4), ("tetra10", 2), ("hexa20", 2)])
@pytest.mark.parametrize("cmat", ["unsymmetric_full", "symmetric_full", "orthotropic"])
def test_general_constitutive_matrix_on_larger_elements(pkg, O, meshes, name, size, cmat):
    """The block kernel pulls the quadrature sum inside the contraction with C and takes short cuts for a symmetric C
    (mirrored blocks by transposition) and for a C without normal/shear coupling (structural zeros skipped): all three
    variants against the oracle's dense B^T C B with the same matrix."""
    rng = np.random.default_rng(11)
    Cm = rng.uniform(0.5, 2.0, (6, 6)) * 1e3 + np.diag([8e3] * 6)
    if cmat == "symmetric_full":
        Cm = 0.5 * (Cm + Cm.T)
    elif cmat == "orthotropic":
        Cm = 0.5 * (Cm + Cm.T)
        Cm[:3, 3:] = 0.0
        Cm[3:, :3] = 0.0
        Cm[3:, 3:] = np.diag(np.diag(Cm[3:, 3:]))
    nodes, conn = meshes.mesh_for(name, size)
    ra = pkg.RawAssembler(nodes, 3)
    desc = ra.make_desc(elem_of(pkg, name), _GeneralSolid(Cm))
    assert ra.add(desc, conn=conn) == 0
    og = O.GeneralAssembler(nodes, 7)
    og.add_isoparametric(KINDS[name]["kind"], np.ascontiguousarray(Cm), 0, conn)
    rp, ri, rv, ab = og.csr(with_absum=True)
    compare_csr(ra.csr(), (rp, ri, rv), ab, TOL)
    ra.close()


@pytest.mark.parametrize("name", ["hexa8", "hexa20"])
def test_collapsed_larger_elements(pkg, O, meshes, name):
    """Bricks collapsed to wedges by REPEATING node ids (one face shrinks to an edge; a Hexa20 then names one node three
    times): legal -- detJ stays positive at the Gauss points -- and several local nodes share a column block, so the row gather
    adds their blocks one local node at a time, in the order of the reference's COO stream."""
    nodes, conn0 = meshes.mesh_for(name, 3, jittered=False)
    conn = conn0.copy()
    for e in (4, 13):
        X = nodes[conn0[e]]
        ymin, xmin = X[:, 1].min(), X[:, 0].min()
        for m in range(conn.shape[1]):
            if abs(X[m, 1] - ymin) < 1e-9 and X[m, 0] > xmin + 1e-9:
                tgt = [k for k in range(conn.shape[1]) if abs(X[k, 1] - ymin) < 1e-9 and abs(X[k, 2] - X[m, 2]) < 1e-9 and abs(X[k, 0] - xmin) < 1e-9][0]
                conn[e, m] = conn0[e, tgt]
    assert len(set(conn[4])) < conn.shape[1]
    ga = pkg.fem.NewGeneralAssembler(nodes, 7)
    err = ga.AddIsoparametric(elem_of(pkg, name), pkg.solids.Isotropic(1e5, 0.3).Solid3D(), len(conn), conn)
    assert err is None, err
    og = O.GeneralAssembler(nodes, 7)
    Cm, bk = O.constitutive(0, 1e5, 0.3)
    og.add_isoparametric(KINDS[name]["kind"], Cm, bk, conn)
    rp, ri, rv, ab = og.csr(with_absum=True)
    compare_csr(ga.Ksolid().csr(), (rp, ri, rv), ab, TOL)


def test_ctx_trim_releases_the_scratch_arena(pkg, meshes):
    """femb200_ctx_trim hands the per-call scratch arena (here: the element-matrix staging of a Hexa20 mesh) back to the device;
    K stays valid and the next assembly on the same context reproduces it bit for bit."""
    import torch
    nodes, conn = meshes.mesh_for("hexa20", 12)
    ra = pkg.RawAssembler(nodes, 3)
    desc = ra.make_desc(pkg.elements.Hexa20(), pkg.solids.Isotropic(200e9, 0.3).Solid3D())
    assert ra.add(desc, conn=conn) == 0
    k1 = [a.copy() for a in ra.csr()]
    torch.cuda.synchronize()
    free0 = torch.cuda.mem_get_info()[0]
    assert pkg.lib().femb200_ctx_trim(ra.ctx) == 0
    free1 = torch.cuda.mem_get_info()[0]
    assert free1 - free0 >= conn.shape[0] * 3600 * 8          # at least the staging array came back
    k1b = ra.csr()
    assert all(np.array_equal(a, b) for a, b in zip(k1, k1b))
    import ctypes as C
    L = pkg.lib()
    h = C.c_void_p()                                           # a second assembler on the SAME (trimmed) context
    nodes_c = np.ascontiguousarray(nodes, dtype=np.float64)
    assert L.femb200_assembler_create(ra.ctx, C.c_void_p(nodes_c.ctypes.data), nodes_c.shape[0], 3, 0, C.byref(h)) == 0
    conn32 = np.ascontiguousarray(conn, dtype=np.int32)
    assert L.femb200_add_isoparametric(h, C.byref(desc), C.c_void_p(conn32.ctypes.data), 0, 0, conn32.shape[0]) == 0
    nnz = L.femb200_nnz(h)
    ip, ix, v = np.empty(3 * nodes.shape[0] + 1, np.int64), np.empty(nnz, np.int32), np.empty(nnz, np.float64)
    assert L.femb200_csr_copy(h, C.c_void_p(ip.ctypes.data), C.c_void_p(ix.ctypes.data), C.c_void_p(v.ctypes.data)) == 0
    assert np.array_equal(ip, k1[0]) and np.array_equal(ix, k1[1]) and np.array_equal(v, k1[2])
    L.femb200_assembler_destroy(h)
    ra.close()


@pytest.mark.parametrize("name,size,mode,flags", [("tetra4", 5, 0, 7), ("hexa8", 4, 0, 7), ("tetra10", 2, 0, 7), ("hexa20", 2, 0, 7),
                                                  ("quad4", 9, 1, 3), ("quad8", 5, 3, 3)])
def test_pattern_only_elements(pkg, O, meshes, name, size, mode, flags):
    """femb200_add_isoparametric_halo: the trailing n_pattern_only elements contribute their sparsity pattern and no values (a
    rank of the row-exchange variant lists its halo that way).  Every numeric path stops a row's sum at its first
    pattern-only element: the fused kernel (Tetra4, Quad4), the element-block gather (Hexa8, Tetra10, Hexa20) and the
    node-row kernel (Quad8).  Expectation: the oracle with the halo assembled from a zero material."""
    nodes, conn = meshes.mesh_for(name, size)
    n_own = (2 * conn.shape[0]) // 3
    own, halo = conn[:n_own], conn[n_own:]
    mdpn = bin(flags).count("1")
    ra = pkg.RawAssembler(nodes, mdpn)
    desc = ra.make_desc(elem_of(pkg, name), constituter_of(pkg, mode, 200e9, 0.3))
    assert ra.add(desc, conn=conn, n_pattern_only=halo.shape[0]) == 0
    Cm, bk = O.constitutive(mode, 200e9, 0.3)
    og = O.GeneralAssembler(nodes, flags)
    og.add_isoparametric(KINDS[name]["kind"], Cm, bk, own)
    og.add_isoparametric(KINDS[name]["kind"], Cm * 0.0, bk, halo)
    rp, ri, rv, ab = og.csr(with_absum=True)
    # the zero-material pass adds nothing to the normaliser of the halo-only entries: give those an absolute scale
    got = ra.csr()
    assert np.array_equal(got[0], rp) and np.array_equal(got[1], ri)
    scale = np.maximum(np.abs(rv), ab)
    assert np.all(np.abs(got[2] - rv) <= TOL * np.where(scale > 0, scale, 1.0))
    assert np.all(got[2][scale == 0] == 0.0)
    ra.close()
