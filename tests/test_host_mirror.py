"""CPU tests of the host layer above the C ABI: the library loads and exports every declared symbol,
and the C++ mirror of packages elements / solids / fem evaluates the same tables as the oracle's literal
restatement of the Go sources.  No GPU compute is called here."""
import numpy as np
import pytest

ELEMS = [("Tetra4", "TETRA4"), ("Tetra10", "TETRA10"), ("Hexa8", "HEXA8"), ("Hexa20", "HEXA20"),
         ("Quad4", "QUAD4"), ("Quad8", "QUAD8"), ("Triangle3", "TRIANGLE3")]


def test_library_exports_every_declared_symbol(pkg):
    L = pkg.lib()
    names = pkg.exported_symbols()
    assert len(names) >= 35
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    assert L.femb200_abi_version() == 1


def test_product_does_not_reference_the_oracle():
    import os
    root = os.path.join(os.path.dirname(__file__), "..", "go-fem_b200")
    for dp, _, files in os.walk(root):
        if "build" in dp.split(os.sep):
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h", ".go")):
                src = open(os.path.join(dp, f), errors="ignore").read()
                assert "femoracle" not in src and "from oracle" not in src and "import oracle" not in src, f


@pytest.mark.parametrize("cls,oname", ELEMS)
def test_element_tables_match_oracle(pkg, O, cls, oname):
    e = getattr(pkg.elements, cls)()
    k = getattr(O, oname)
    assert e.LenNodes() == O.len_nodes(k)
    assert e.Dofs() == O.elem_dofs(k, 0)
    assert np.array_equal(e.IsoparametricNodes(), O.iso_nodes(k))
    pos, w = e.Quadrature()
    opos, ow = O.quadrature(k)
    assert pos.shape == opos.shape
    assert np.allclose(pos, opos, rtol=0, atol=2e-16) and np.allclose(w, ow, rtol=4e-16, atol=0)
    rng = np.random.default_rng(7)
    pts = list(opos) + [rng.uniform(-1, 1, 3) for _ in range(20)]
    for p in pts:
        assert np.allclose(e.Basis(p), O.basis(k, p), rtol=0, atol=5e-15)
        assert np.allclose(e.BasisDiff(p), O.basis_diff(k, p), rtol=0, atol=2e-14)


@pytest.mark.parametrize("cls,oname,orders", [("Hexa8", "HEXA8", [1, 2, 3, 4, 5, 6]), ("Hexa20", "HEXA20", [2, 3, 4]),
                                              ("Quad4", "QUAD4", [1, 2, 3, 4, 5, 6]), ("Quad8", "QUAD8", [2, 3, 4]),
                                              ("Triangle3", "TRIANGLE3", [1, 2, 3, 4])])
def test_quadrature_orders_match_oracle(pkg, O, cls, oname, orders):
    for q in orders:
        pos, w = getattr(pkg.elements, cls)(QuadratureOrder=q).Quadrature()
        opos, ow = O.quadrature(getattr(O, oname), q)
        assert pos.shape == opos.shape
        assert np.allclose(pos, opos, rtol=0, atol=3e-16) and np.allclose(w, ow, rtol=1e-15, atol=0)


def test_strings_and_node_dofs(pkg):
    el = pkg.elements
    assert el.Tetra4().String() == "TETRA4" and el.Hexa20().String() == "HEXA20"
    assert el.Quad4().String() == "Quad43d(order=2)" and el.Quad8(QuadratureOrder=4).String() == "QUAD8(order=4)"
    assert el.Triangle3().String() == "TRI6(order=1)"            # sic, triangles.go:78
    assert el.Tetra4(NodeDofs=1).Dofs() == 7                      # Tetra4 ignores NodeDofs (tetra4.go:18-20)
    assert el.Quad4(NodeDofs=1).Dofs() == 1 and el.Hexa8(NodeDofs=1).Dofs() == 1


@pytest.mark.parametrize("mode,args", [(0, (200e9, 0.3)), (1, (1000, 0.33)), (2, (1, 0.3)), (3, (5.0, 0.2)), (4, (45.0,)), (5, (45.0,)), (6, (45.0,))])
def test_constitutive_matches_oracle(pkg, O, mode, args):
    s = pkg.solids
    iso = [lambda E, nu: s.Isotropic(E, nu).Solid3D(), lambda E, nu: s.Isotropic(E, nu).Axisymmetric(),
           lambda E, nu: s.Isotropic(E, nu).PlaneStess(), lambda E, nu: s.Isotropic(E, nu).PlaneStrain(),
           lambda K: s.IsotropicConductivity(K).Solid3D(), lambda K: s.IsotropicConductivity(K).Plane(),
           lambda K: s.IsotropicConductivity(K).Axisymmetric()]
    c = iso[mode](*args)
    Co, bko = O.constitutive(mode, *args)
    assert np.array_equal(c.Constitutive(), Co) and c.BKind() == bko


def test_isotropic_incompressible_is_an_error(pkg, O):
    with pytest.raises(pkg.GoError):
        pkg.solids.Isotropic(1.0, 0.5).Solid3D().Constitutive()   # isotropic.go:24-26
    with pytest.raises(O.OracleError):
        O.constitutive(O.ISO_SOLID3D, 1.0, 0.5)


def test_fixity(pkg):
    fem = pkg.fem
    f = fem.NewFixity(fem.DofPosX | fem.DofPosY, 4)
    f.Fix(0, fem.DofPosX | fem.DofPosY)
    f.Fix(2, fem.DofPosY | fem.DofPosZ)                          # Z is not in the model: ignored
    assert f.FixedDofs().tolist() == [0, 1, 5]
    assert f.FreeDofs().tolist() == [2, 3, 4, 6, 7]
    f.Free(0, fem.DofPosX)
    assert f.FixedDofs().tolist() == [1, 5]


def test_no_gpu_means_loud_failure(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError):
        pkg.fem.NewGeneralAssembler(np.zeros((4, 3)), 7)
    with pytest.raises(RuntimeError):
        pkg.RawAssembler(np.zeros((4, 3)), 3)


def test_mesh_generators(meshes, O):
    # every generated mesh has positive Jacobians in the reference's node order (oracle accepts it)
    cases = [("tetra4", 3, O.TETRA4, O.ISO_SOLID3D, 7), ("tetra10", 2, O.TETRA10, O.ISO_SOLID3D, 7), ("hexa8", 3, O.HEXA8, O.ISO_SOLID3D, 7),
             ("hexa20", 2, O.HEXA20, O.ISO_SOLID3D, 7), ("quad4", 5, O.QUAD4, O.ISO_AXISYM, 3), ("quad8", 4, O.QUAD8, O.ISO_AXISYM, 3),
             ("triangle3", 4, O.TRIANGLE3, O.ISO_AXISYM, 3)]
    for name, size, k, mode, flags in cases:
        nodes, conn = meshes.mesh_for(name, size)
        assert conn.dtype == np.int64 and conn.min() >= 0 and conn.max() <= nodes.shape[0] - 1
        C, bk = O.constitutive(mode, 200e9, 0.3)
        ga = O.GeneralAssembler(nodes, flags)
        ga.add_isoparametric(k, C, bk, conn)
        assert ga.csr()[2].size > 0


def test_go_shim_binds_only_declared_symbols(pkg):
    """The (uncompiled: no Go toolchain in the image) cgo shim under go-fem_b200/gofem may only call entry points that
    include/femb200.h declares and libfemb200.so exports, and it must bind the hot-path entry points."""
    import os
    import re
    ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    declared = set(pkg.exported_symbols())
    used = set()
    for dirpath, _, files in os.walk(os.path.join(ROOT, "go-fem_b200", "gofem")):
        for f in files:
            if f.endswith(".go"):
                used |= set(re.findall(r"\bC\.(femb200_\w+)\(", open(os.path.join(dirpath, f)).read()))
    assert used, "no cgo calls found"
    assert used <= declared, sorted(used - declared)
    for must in ("femb200_ctx_create", "femb200_assembler_create", "femb200_add_isoparametric", "femb200_last_error", "femb200_csr_copy",
                 "femb200_nnz", "femb200_total_dofs", "femb200_reassemble"):
        assert must in used, must



def test_shape_function_tables_equal_the_oracle_bitwise():
    """The host mirror evaluates Basis / BasisDiff / Quadrature with the reference's own term order (elements/*.go), so the
    tables that cross the C ABI are bit-identical to the oracle's (= Go's): a one-ulp difference in dN shows up as 1e-13..1e-12
    of normalised error in K once |X|/h reaches 100 (found at config 3 / config 4 size)."""
    import numpy as np
    import __graft_entry__ as entry
    from conftest import KINDS
    O = entry.load_oracle()
    O.build()
    pkg = entry.load_package()
    rng = np.random.default_rng(1)
    el = pkg.elements
    for name, cls in (("tetra4", el.Tetra4), ("tetra10", el.Tetra10), ("hexa8", el.Hexa8), ("hexa20", el.Hexa20), ("quad4", el.Quad4),
                      ("quad8", el.Quad8), ("triangle3", el.Triangle3)):
        kind = KINDS[name]["kind"]
        for order in (0, 1, 2, 3, 4):
            e = cls(QuadratureOrder=order)
            pos, w = e.Quadrature()
            if len(w) == 0:
                continue
            opos, ow = O.quadrature(kind, order)
            assert np.array_equal(np.array(pos), np.array(opos).reshape(np.array(pos).shape)), (name, order)
            assert np.array_equal(np.array(w), np.array(ow)), (name, order)
            for p in list(pos) + [tuple(rng.uniform(-0.9, 0.9, 3)) for _ in range(10)]:
                p = tuple(float(x) for x in p)
                a = np.array(e.BasisDiff(p))
                assert np.array_equal(a, np.array(O.basis_diff(kind, p)).reshape(a.shape)), (name, order, p)
                a = np.array(e.Basis(p))
                assert np.array_equal(a, np.array(O.basis(kind, p)).reshape(a.shape)), (name, order, p)
