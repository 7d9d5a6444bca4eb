"""ctypes front-end of the CPU ORACLE (oracle/libfemoracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the cpu_baseline /
--impl reference legs of bench.py.  The product package (go-fem_b200/) never imports this.
The oracle is a C restatement of soypat/go-fem's AddIsoparametric path; see femoracle.h.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libfemoracle.so")

TETRA4, TETRA10, HEXA8, HEXA20, QUAD4, QUAD8, TRIANGLE3, TRIANGLE6 = range(8)
B_XYZ, B_PLANE, B_AXISYM, B_THERMAL, B_THERMAL_AXISYM = range(5)
ISO_SOLID3D, ISO_AXISYM, ISO_PLANESTRESS, ISO_PLANESTRAIN, COND_SOLID3D, COND_PLANE, COND_AXISYM = range(7)
ERR_NAMES = {0: "ok", 1: "neg_det", 2: "zero_det", 3: "solve", 4: "nan_scale", 5: "dofs", 6: "empty_dofmap",
             7: "bad_length", 8: "node_oob", 9: "constitutive", 10: "element", 11: "quadrature", 12: "dims"}


def build(force=False):
    srcs = [os.path.join(_HERE, f) for f in ("femoracle_elements.c", "femoracle_assemble.c", "femoracle.h")]
    if not force and os.path.exists(_LIB) and all(os.path.getmtime(_LIB) >= os.path.getmtime(s) for s in srcs):
        return _LIB
    subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return _LIB


_lib = None
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_i64p = C.POINTER(C.c_int64)
_i32p = C.POINTER(C.c_int32)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            build()
        L = C.CDLL(_LIB)
        L.femo_elem_measure.restype = C.c_double
        L.femo_new_general_assembler.restype = C.c_void_p
        L.femo_new_general_assembler.argtypes = [_dp, C.c_int64, C.c_int]
        L.femo_free.argtypes = [C.c_void_p]
        L.femo_total_dofs.restype = C.c_int64
        L.femo_total_dofs.argtypes = [C.c_void_p]
        L.femo_nnz.restype = C.c_int64
        L.femo_nnz.argtypes = [C.c_void_p]
        L.femo_last_seconds.restype = C.c_double
        L.femo_last_seconds.argtypes = [C.c_void_p]
        L.femo_csr.argtypes = [C.c_void_p, _i64p, _i32p, _dp]
        L.femo_csr_absum.argtypes = [C.c_void_p, _dp]
        L.femo_add_isoparametric.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, _dp, C.c_int, C.c_int,
                                             _i64p, C.c_int64, _i64p, _ip]
        L.femo_element_matrices.argtypes = [_dp, C.c_int64, C.c_int, C.c_int, C.c_int, _dp, C.c_int, C.c_int,
                                            _i64p, C.c_int64, _dp, _i64p, _ip]
        L.femo_isoparametric_strains.argtypes = [C.c_void_p, _dp, C.c_int, C.c_int, C.c_int, _dp, C.c_int, C.c_int,
                                                 _i64p, C.c_int64, _dp, _i64p, _ip]
        L.femo_elem_basis.argtypes = [C.c_int, _dp, _dp]
        L.femo_elem_basis_diff.argtypes = [C.c_int, _dp, _dp]
        L.femo_elem_iso_nodes.argtypes = [C.c_int, _dp]
        L.femo_elem_quadrature.argtypes = [C.c_int, C.c_int, _dp, _dp, C.c_int]
        L.femo_constitutive.argtypes = [C.c_int, C.c_double, C.c_double, _dp, _ip, _ip]
        L.femo_mapdofs.argtypes = [C.c_int, C.c_int, _ip, C.c_int]
        _lib = L
    return _lib


def _d(a):
    return a.ctypes.data_as(_dp)


class OracleError(Exception):
    def __init__(self, kind, elem=-1, qp=-1):
        self.kind, self.elem, self.qp = kind, elem, qp
        super().__init__(f"oracle error {ERR_NAMES.get(kind, kind)} elem={elem} qp={qp}")


def set_normaliser(on):
    """Test-only s_ij bookkeeping on/off (off for timed cpu_baseline runs)."""
    lib().femo_set_normaliser(int(on))


def len_nodes(kind):
    return lib().femo_elem_len_nodes(kind)


def elem_dofs(kind, node_dofs=0):
    return lib().femo_elem_dofs(kind, node_dofs)


def measure(kind):
    return lib().femo_elem_measure(kind)


def iso_nodes(kind):
    nn = len_nodes(kind)
    out = np.zeros((nn, 3))
    lib().femo_elem_iso_nodes(kind, _d(out))
    return out


def basis(kind, v):
    v = np.ascontiguousarray(v, dtype=np.float64)
    out = np.zeros(32)
    n = lib().femo_elem_basis(kind, _d(v), _d(out))
    return out[:n].copy()


def basis_diff(kind, v):
    v = np.ascontiguousarray(v, dtype=np.float64)
    out = np.zeros(96)
    n = lib().femo_elem_basis_diff(kind, _d(v), _d(out))
    return out[:n].copy()


def quadrature(kind, order=0):
    pos = np.zeros((216, 3))
    w = np.zeros(216)
    n = lib().femo_elem_quadrature(kind, order, _d(pos), _d(w), 216)
    if n < 0:
        raise OracleError(-n)
    return pos[:n].copy(), w[:n].copy()


def constitutive(mode, p0, p1=0.0):
    Cm = np.zeros(36)
    dimC, bkind = C.c_int(0), C.c_int(0)
    rc = lib().femo_constitutive(mode, p0, p1, _d(Cm), C.byref(dimC), C.byref(bkind))
    if rc:
        raise OracleError(rc)
    return Cm[: dimC.value ** 2].reshape(dimC.value, dimC.value).copy(), bkind.value


def mapdofs(model, elem):
    out = (C.c_int * 16)()
    n = lib().femo_mapdofs(model, elem, out, 16)
    if n < 0:
        raise OracleError(-n)
    return list(out[:n])


class GeneralAssembler:
    """Oracle twin of fem.GeneralAssembler (assembler.go:13-37)."""

    def __init__(self, nodes, model_dofs):
        nodes = np.ascontiguousarray(nodes, dtype=np.float64).reshape(-1, 3)
        self.n_nodes = nodes.shape[0]
        self._h = lib().femo_new_general_assembler(_d(nodes), self.n_nodes, int(model_dofs))

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                lib().femo_free(self._h)
                self._h = None
        except Exception:  # interpreter shutdown
            pass

    def total_dofs(self):
        return lib().femo_total_dofs(self._h)

    def add_isoparametric(self, kind, Cmat, bkind, conn, quad_order=0, node_dofs=0):
        Cmat = np.ascontiguousarray(Cmat, dtype=np.float64)
        conn = np.ascontiguousarray(conn, dtype=np.int64)
        nn = len_nodes(kind)
        nelem = conn.size // nn if nn > 0 else 0
        ee, qq = C.c_int64(-1), C.c_int(-1)
        rc = lib().femo_add_isoparametric(self._h, kind, quad_order, node_dofs, _d(Cmat), Cmat.shape[0], bkind,
                                          conn.ctypes.data_as(_i64p), nelem, C.byref(ee), C.byref(qq))
        if rc:
            raise OracleError(rc, ee.value, qq.value)

    def last_seconds(self):
        return lib().femo_last_seconds(self._h)

    def csr(self, with_absum=False):
        n = self.total_dofs()
        nnz = lib().femo_nnz(self._h)
        indptr = np.zeros(n + 1, dtype=np.int64)
        indices = np.zeros(nnz, dtype=np.int32)
        values = np.zeros(nnz, dtype=np.float64)
        lib().femo_csr(self._h, indptr.ctypes.data_as(_i64p), indices.ctypes.data_as(_i32p), _d(values))
        if with_absum:
            ab = np.zeros(nnz)
            lib().femo_csr_absum(self._h, _d(ab))
            return indptr, indices, values, ab
        return indptr, indices, values

    def dense(self):
        indptr, indices, values = self.csr()
        n = self.total_dofs()
        K = np.zeros((n, n))
        for r in range(n):
            K[r, indices[indptr[r]:indptr[r + 1]]] = values[indptr[r]:indptr[r + 1]]
        return K

    def strains(self, u, kind, Cmat, bkind, conn, quad_order=0, node_dofs=0):
        Cmat = np.ascontiguousarray(Cmat, dtype=np.float64)
        conn = np.ascontiguousarray(conn, dtype=np.int64)
        u = np.ascontiguousarray(u, dtype=np.float64)
        nn = len_nodes(kind)
        nelem = conn.size // nn
        nqp = quadrature(kind, quad_order)[1].size
        out = np.zeros((nelem, nqp, Cmat.shape[0]))
        ee, qq = C.c_int64(-1), C.c_int(-1)
        rc = lib().femo_isoparametric_strains(self._h, _d(u), kind, quad_order, node_dofs, _d(Cmat), Cmat.shape[0],
                                              bkind, conn.ctypes.data_as(_i64p), nelem, _d(out), C.byref(ee), C.byref(qq))
        if rc:
            raise OracleError(rc, ee.value, qq.value)
        return out


def element_matrices(nodes, kind, Cmat, bkind, conn, quad_order=0, node_dofs=0):
    nodes = np.ascontiguousarray(nodes, dtype=np.float64).reshape(-1, 3)
    Cmat = np.ascontiguousarray(Cmat, dtype=np.float64)
    conn = np.ascontiguousarray(conn, dtype=np.int64)
    nn = len_nodes(kind)
    nelem = conn.size // nn
    ndn = bin(elem_dofs(kind, node_dofs)).count("1")
    nd = nn * ndn
    Ke = np.zeros((nelem, nd, nd))
    ee, qq = C.c_int64(-1), C.c_int(-1)
    rc = lib().femo_element_matrices(_d(nodes), nodes.shape[0], kind, quad_order, node_dofs, _d(Cmat), Cmat.shape[0],
                                     bkind, conn.ctypes.data_as(_i64p), nelem, _d(Ke), C.byref(ee), C.byref(qq))
    if rc:
        raise OracleError(rc, ee.value, qq.value)
    return Ke
