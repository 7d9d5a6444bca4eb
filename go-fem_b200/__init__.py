"""gofem_b200 -- Python harness over libfemb200.so (B200-native isoparametric stiffness assembly).

The product is the shared library: hand-written sm_100a CUDA kernels behind the C ABI of
include/femb200.h, plus the C++ host mirror of the go-fem API (host/gofem.hpp).  This module only
binds them with ctypes so that tests and bench.py read like the reference's own tests:

    ga = fem.NewGeneralAssembler(nodes, fem.DofPos)
    err = ga.AddIsoparametric(elements.Tetra4(), solids.Isotropic(E=200e9, Poisson=0.3).Solid3D(), n, get_element)
    K = ga.Ksolid()

There is no CPU fallback: every assembly call goes to the GPU through the library, and loading fails
loudly when libfemb200.so is missing.  (The directory is named go-fem_b200; load it with
__graft_entry__.load_package() which registers it as module `gofem_b200`.)
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libfemb200.so")

N_TIMERS = 10
TIMER_NAMES = ["upload_validate", "adjacency", "row_pattern", "csr_expand", "numeric_rows", "merge", "total_device", "launches", "k_geometry", "k_rows"]

_dp = C.POINTER(C.c_double)
_i64p = C.POINTER(C.c_int64)
_i32p = C.POINTER(C.c_int32)
_u16p = C.POINTER(C.c_uint16)


class ElemDesc(C.Structure):
    """struct femb200_elem_desc (include/femb200.h)"""
    _fields_ = [("d", C.c_int32), ("nn", C.c_int32), ("ndn", C.c_int32), ("nqp", C.c_int32), ("dimC", C.c_int32),
                ("bkind", C.c_int32), ("model_dofs_per_node", C.c_int32), ("reserved", C.c_int32),
                ("Npg", _dp), ("dNpg", _dp), ("wpg", _dp), ("C", _dp), ("dofmap", _i32p)]


_lib = None


def lib():
    """Loads libfemb200.so; raises if it has not been built (no fallback of any kind)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: run `python go-fem_b200/build.py` (nvcc, sm_100a). "
                           "There is no CPU fallback for the assembly path.")
    L = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    L.femb200_abi_version.restype = C.c_int
    L.femb200_ctx_create.argtypes = [C.c_int, C.POINTER(vp)]
    L.femb200_ctx_destroy.argtypes = [vp]
    L.femb200_ctx_set_stream.argtypes = [vp, vp]
    L.femb200_ctx_trim.argtypes = [vp]
    L.femb200_ctx_set_concurrent.argtypes = [vp, C.c_int32]
    L.femb200_last_cuda_error.argtypes = [vp]
    L.femb200_last_cuda_error.restype = C.c_char_p
    L.femb200_assembler_create.argtypes = [vp, vp, C.c_int64, C.c_int32, C.c_int32, C.POINTER(vp)]
    L.femb200_assembler_destroy.argtypes = [vp]
    L.femb200_add_isoparametric.argtypes = [vp, C.POINTER(ElemDesc), vp, C.c_int32, C.c_int32, C.c_int64]
    L.femb200_last_error.argtypes = [vp, _i32p, _i64p, _i32p]
    L.femb200_reassemble.argtypes = [vp, C.POINTER(ElemDesc), vp, C.c_int32]
    L.femb200_total_dofs.argtypes = [vp]
    L.femb200_total_dofs.restype = C.c_int64
    L.femb200_nnz.argtypes = [vp]
    L.femb200_nnz.restype = C.c_int64
    L.femb200_csr_copy.argtypes = [vp, vp, vp, vp]
    L.femb200_csr_device.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]
    L.femb200_bsr_blocks.argtypes = [vp]
    L.femb200_bsr_blocks.restype = C.c_int64
    L.femb200_bsr_copy.argtypes = [vp, vp, vp, vp]
    L.femb200_add_coo.argtypes = [vp, C.c_int64, _i64p, _i64p, _dp]
    L.femb200_csr_at.argtypes = [vp, C.c_int64, C.c_int64, _dp]
    L.femb200_csr_matvec.argtypes = [vp, _dp, _dp]
    L.femb200_csr_slice.argtypes = [vp, C.c_int64, _i64p, C.c_int64, _i64p, _i64p, _i32p, _dp, _i64p]
    L.femb200_csr_accumulate_rows_device.argtypes = [vp, C.c_int64, vp, vp, vp, vp]
    L.femb200_csr_keep_rows_device.argtypes = [vp, vp]
    L.femb200_assembler_set_node_mask_device.argtypes = [vp, vp]
    L.femb200_assembler_set_column_map_device.argtypes = [vp, vp, C.c_int64]
    L.femb200_csr_add_entries_device.argtypes = [vp, C.c_int64, vp, vp, vp, vp]
    L.femb200_csr_pack_rows_async.argtypes = [vp, C.c_int64, vp, vp, vp, vp, vp]
    L.femb200_csr_add_entries_async.argtypes = [vp, C.c_int64, vp, vp, vp, vp, vp]
    L.femb200_add_isoparametric_halo.argtypes = [vp, C.POINTER(ElemDesc), vp, C.c_int32, C.c_int32, C.c_int64, C.c_int64]
    L.femb200_last_slot_map.argtypes = [vp, _i64p]
    L.femb200_isoparametric_strains.argtypes = [vp, C.POINTER(ElemDesc), vp, C.c_int32, C.c_int32, C.c_int64, _dp, _dp]
    L.femb200_phase_ms.argtypes = [vp, C.POINTER(C.c_float)]
    L.femb200_kernel_launches.argtypes = [vp]
    L.femb200_kernel_launches.restype = C.c_int64
    L.femb200_sync.argtypes = [vp]
    L.femb200_probe_peaks.argtypes = [vp, C.c_int32, _dp]
    L.femb200_dist_create.argtypes = [vp, C.c_int32, C.c_int32, vp, C.c_int32, C.c_int64, C.c_int32, vp, C.c_int32, C.c_int32, C.c_int64, C.c_int32, C.c_int32, C.POINTER(vp)]
    L.femb200_dist_destroy.argtypes = [vp]
    L.femb200_dist_add_isoparametric.argtypes = [vp, C.POINTER(ElemDesc)]
    L.femb200_dist_last_error.argtypes = [vp, _i32p, _i64p, _i32p]
    L.femb200_dist_reset.argtypes = [vp]
    L.femb200_dist_assembler.argtypes = [vp]
    L.femb200_dist_assembler.restype = vp
    L.femb200_dist_info.argtypes = [vp, _i64p, C.POINTER(C.c_float)]
    L.femb200_dist_local_nodes.argtypes = [vp, vp, vp]
    L.femb200_group_create.argtypes = [_i32p, C.c_int32, vp, C.c_int64, C.c_int32, vp, C.c_int32, C.c_int64, C.c_int32, C.c_int32, C.POINTER(vp)]
    L.femb200_group_destroy.argtypes = [vp]
    L.femb200_group_size.argtypes = [vp]
    L.femb200_group_rank.argtypes = [vp, C.c_int32]
    L.femb200_group_rank.restype = vp
    L.femb200_group_add_isoparametric.argtypes = [vp, C.POINTER(ElemDesc), _i32p]
    L.femb200_mesh_bcc_tetra4.argtypes = [vp, C.c_int32, C.c_int32, C.c_int32, C.c_double, C.c_int32, C.POINTER(vp), _i64p, C.POINTER(vp), _i64p]
    L.femb200_mesh_free.argtypes = [vp, vp]
    L.femb200_copy_to_host.argtypes = [vp, vp, vp, C.c_int64]
    # host mirror
    L.gofem_elem_info.argtypes = [C.c_int, C.c_int, C.c_int, _i32p, _i32p, _dp, _i32p, _dp, _dp, C.c_char_p, C.c_int]
    L.gofem_elem_basis.argtypes = [C.c_int, _dp, _dp, _dp, _i32p]
    L.gofem_constitutive.argtypes = [C.c_int, C.c_double, C.c_double, _dp, _i32p, _i32p]
    L.gofem_new_general_assembler.argtypes = [C.c_int, _dp, C.c_int64, C.c_int, C.POINTER(vp)]
    L.gofem_free.argtypes = [vp]
    L.gofem_add_isoparametric.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, _i64p, _i32p, C.c_int64, _dp]
    L.gofem_add_isoparametric_flat.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, _i64p, C.c_int64]
    L.gofem_isoparametric_strains.argtypes = [vp, _dp, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, _i64p, C.c_int64, _dp]
    L.gofem_last_error.argtypes = [vp, _i32p]
    L.gofem_last_error.restype = C.c_char_p
    L.gofem_total_dofs.argtypes = [vp]
    L.gofem_total_dofs.restype = C.c_int64
    L.gofem_ksolid_nnz.argtypes = [vp]
    L.gofem_ksolid_nnz.restype = C.c_int64
    L.gofem_ksolid_at.argtypes = [vp, C.c_int64, C.c_int64]
    L.gofem_ksolid_at.restype = C.c_double
    L.gofem_ksolid_csr.argtypes = [vp, _i64p, _i32p, _dp]
    L.gofem_ksolid_mulvec.argtypes = [vp, _dp, _dp]
    L.gofem_device_assembler.argtypes = [vp]
    L.gofem_device_assembler.restype = vp
    L.gofem_device_ctx.argtypes = [vp]
    L.gofem_device_ctx.restype = vp
    L.gofem_fixity_dofs.argtypes = [C.c_int, C.c_int64, _u16p, C.c_int, _i64p]
    L.gofem_fixity_dofs.restype = C.c_int64
    _lib = L
    return L


def exported_symbols():
    """Every symbol the two headers declare (checked by the CPU test suite against the built .so)."""
    import re
    names = []
    for h in ("femb200.h", "femb200_host.h"):
        src = open(os.path.join(_HERE, "..", "include", h)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names += re.findall(r"\b((?:femb200|gofem)_\w+)\s*\(", src)
    return sorted(set(names))


def _csr_slice(handle, rows, cols):
    rows = np.ascontiguousarray(rows, dtype=np.int64)
    cols = np.ascontiguousarray(cols, dtype=np.int64)
    nnz = C.c_int64(0)
    L = lib()
    rp, cp = rows.ctypes.data_as(_i64p), cols.ctypes.data_as(_i64p)
    rc = L.femb200_csr_slice(handle, rows.size, rp, cols.size, cp, None, None, None, C.byref(nnz))
    if rc:
        raise RuntimeError(f"femb200_csr_slice -> {rc}")
    indptr = np.zeros(rows.size + 1, dtype=np.int64)
    indices = np.zeros(nnz.value, dtype=np.int32)
    values = np.zeros(nnz.value, dtype=np.float64)
    rc = L.femb200_csr_slice(handle, rows.size, rp, cols.size, cp, indptr.ctypes.data_as(_i64p), indices.ctypes.data_as(_i32p), _d(values), C.byref(nnz))
    if rc:
        raise RuntimeError(f"femb200_csr_slice -> {rc}")
    return indptr, indices, values


def _d(a):
    return a.ctypes.data_as(_dp)


class GoError(Exception):
    """A Go `error` (or panic) returned by the mirrored API."""

    def __init__(self, msg, is_panic=False):
        super().__init__(msg)
        self.msg, self.is_panic = msg, is_panic


# ----------------------------------------------------------------------------------------------- fem
class fem:
    DofPosX, DofPosY, DofPosZ, DofRotX, DofRotY, DofRotZ = 1, 2, 4, 8, 16, 32   # fem.go:70-82
    DofPos, DofRot, Dof6 = 7, 56, 63                                            # fem.go:86-90

    @staticmethod
    def NewGeneralAssembler(nodes, modelDofs, device=0):
        return GeneralAssembler(nodes, modelDofs, device)

    @staticmethod
    def NewFixity(modelDofs, numNodes):
        return Fixity(modelDofs, numNodes)


class _Elem:
    KIND = -1

    def __init__(self, QuadratureOrder=0, NodeDofs=0):
        self.QuadratureOrder, self.NodeDofs = int(QuadratureOrder), int(NodeDofs)

    def _info(self):
        L = lib()
        nn, fl, nqp = C.c_int32(), C.c_int32(), C.c_int32()
        iso = np.zeros(60)
        qp = np.zeros(216 * 3)
        qw = np.zeros(216)
        name = C.create_string_buffer(64)
        L.gofem_elem_info(self.KIND, self.QuadratureOrder, self.NodeDofs, C.byref(nn), C.byref(fl), _d(iso), C.byref(nqp), _d(qp), _d(qw), name, 64)
        return nn.value, fl.value, iso[: 3 * nn.value].reshape(-1, 3).copy(), qp[: 3 * nqp.value].reshape(-1, 3).copy(), qw[: nqp.value].copy(), name.value.decode()

    def LenNodes(self):
        return self._info()[0]

    def Dofs(self):
        return self._info()[1]

    def IsoparametricNodes(self):
        return self._info()[2]

    def Quadrature(self):
        i = self._info()
        return i[3], i[4]

    def String(self):
        return self._info()[5]

    def _basis(self, v):
        v = np.ascontiguousarray(v, dtype=np.float64)
        N, dN, n = np.zeros(32), np.zeros(96), C.c_int32()
        lib().gofem_elem_basis(self.KIND, _d(v), _d(N), _d(dN), C.byref(n))
        return N[: self.LenNodes()].copy(), dN[: n.value].copy()

    def Basis(self, v):
        return self._basis(v)[0]

    def BasisDiff(self, v):
        return self._basis(v)[1]


class elements:
    class Tetra4(_Elem):
        KIND = 0

    class Tetra10(_Elem):
        KIND = 1

    class Hexa8(_Elem):
        KIND = 2

    class Hexa20(_Elem):
        KIND = 3

    class Quad4(_Elem):
        KIND = 4

    class Quad8(_Elem):
        KIND = 5

    class Triangle3(_Elem):
        KIND = 6

    class Triangle6(_Elem):
        KIND = 7


class _Constituter:
    def __init__(self, ckind, p0, p1=0.0):
        self.ckind, self.p0, self.p1 = ckind, float(p0), float(p1)

    def Constitutive(self):
        Cm, dimC, bk = np.zeros(36), C.c_int32(), C.c_int32()
        rc = lib().gofem_constitutive(self.ckind, self.p0, self.p1, _d(Cm), C.byref(dimC), C.byref(bk))
        if rc:
            raise GoError("matrix singular or near-singular")
        return Cm[: dimC.value ** 2].reshape(dimC.value, dimC.value).copy()

    def BKind(self):
        Cm, dimC, bk = np.zeros(36), C.c_int32(), C.c_int32()
        lib().gofem_constitutive(self.ckind, self.p0, self.p1, _d(Cm), C.byref(dimC), C.byref(bk))
        return bk.value


class solids:
    class Isotropic:
        def __init__(self, E, Poisson):
            self.E, self.Poisson = E, Poisson

        def Solid3D(self):
            return _Constituter(0, self.E, self.Poisson)

        def Axisymmetric(self):
            return _Constituter(1, self.E, self.Poisson)

        def PlaneStess(self):  # sic (isotropic.go:55)
            return _Constituter(2, self.E, self.Poisson)

        def PlaneStrain(self):
            return _Constituter(3, self.E, self.Poisson)

    class IsotropicConductivity:
        def __init__(self, K):
            self.K = K

        def Solid3D(self):
            return _Constituter(4, self.K)

        def Plane(self):
            return _Constituter(5, self.K)

        def Axisymmetric(self):
            return _Constituter(6, self.K)


class Sparse:
    """The lap.Sparse view returned by Ksolid(): Dims / At / NonZero / MulVec + the CSR arrays."""

    def __init__(self, ga):
        self._ga = ga

    def Dims(self):
        n = self._ga.TotalDofs()
        return n, n

    def NonZero(self):
        return lib().gofem_ksolid_nnz(self._ga._h)

    def At(self, i, j):
        return lib().gofem_ksolid_at(self._ga._h, i, j)

    def csr(self):
        n = self._ga.TotalDofs()
        nnz = self.NonZero()
        indptr = np.zeros(n + 1, dtype=np.int64)
        indices = np.zeros(nnz, dtype=np.int32)
        values = np.zeros(nnz, dtype=np.float64)
        lib().gofem_ksolid_csr(self._ga._h, indptr.ctypes.data_as(_i64p), indices.ctypes.data_as(_i32p), _d(values))
        return indptr, indices, values

    def dense(self):
        indptr, indices, values = self.csr()
        n = self._ga.TotalDofs()
        K = np.zeros((n, n))
        for r in range(n):
            K[r, indices[indptr[r]:indptr[r + 1]]] = values[indptr[r]:indptr[r + 1]]
        return K

    def Slice(self, rows, cols):
        """lap.Slice(K, rows, cols) on the device (femb200_csr_slice): CSR (indptr, indices, values) of K[rows][:, cols]
        with column indices = positions in `cols` (ascending, unique)."""
        return _csr_slice(lib().gofem_device_assembler(self._ga._h), rows, cols)

    def MulVec(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.zeros_like(x)
        lib().gofem_ksolid_mulvec(self._ga._h, _d(x), _d(y))
        return y


class GeneralAssembler:
    """fem.GeneralAssembler (assembler.go:13-37) backed by the GPU library through the C++ host mirror."""

    def __init__(self, nodes, modelDofs, device=0):
        nodes = np.ascontiguousarray(nodes, dtype=np.float64).reshape(-1, 3)
        self._nodes = nodes
        self._h = C.c_void_p()
        rc = lib().gofem_new_general_assembler(device, _d(nodes), nodes.shape[0], int(modelDofs), C.byref(self._h))
        if rc:
            msg = lib().gofem_last_error(self._h, None).decode() if self._h else "gofem_new_general_assembler failed"
            raise RuntimeError(msg)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().gofem_free(self._h)
            self._h = None

    def close(self):
        self.__del__()

    def TotalDofs(self):
        return lib().gofem_total_dofs(self._h)

    def Ksolid(self):
        return Sparse(self)

    def _err(self):
        p = C.c_int32()
        msg = lib().gofem_last_error(self._h, C.byref(p)).decode()
        return GoError(msg, bool(p.value)) if msg else None

    def AddIsoparametric(self, elemT, c, Nelem, getElement, orientation=None):
        """getElement: callable i -> sequence of node ids (called Nelem times, like the reference), or an
        int array [Nelem, nn] served row by row through the C++ callback.  Returns None or a GoError."""
        lens = None
        if callable(getElement):
            rows = [np.asarray(getElement(i), dtype=np.int64).ravel() for i in range(Nelem)]
            ln = np.array([r.size for r in rows], dtype=np.int32)
            conn = np.concatenate(rows) if rows else np.zeros(0, dtype=np.int64)
            if Nelem and not np.all(ln == elemT.LenNodes()):
                lens = ln
        else:
            conn = np.ascontiguousarray(getElement, dtype=np.int64).ravel()
        conn = np.ascontiguousarray(conn, dtype=np.int64)
        ori = None if orientation is None else np.ascontiguousarray(orientation, dtype=np.float64)
        lib().gofem_add_isoparametric(self._h, elemT.KIND, elemT.QuadratureOrder, elemT.NodeDofs, c.ckind, c.p0, c.p1,
                                      conn.ctypes.data_as(_i64p), None if lens is None else lens.ctypes.data_as(_i32p), Nelem,
                                      None if ori is None else _d(ori))
        return self._err()

    def AddIsoparametricFlat(self, elemT, c, conn):
        conn = np.ascontiguousarray(conn, dtype=np.int64)
        nelem = conn.size // elemT.LenNodes()
        lib().gofem_add_isoparametric_flat(self._h, elemT.KIND, elemT.QuadratureOrder, elemT.NodeDofs, c.ckind, c.p0, c.p1,
                                           conn.ctypes.data_as(_i64p), nelem)
        return self._err()

    def IsoparametricStrains(self, displacements, elemT, c, conn):
        conn = np.ascontiguousarray(conn, dtype=np.int64)
        u = np.ascontiguousarray(displacements, dtype=np.float64)
        nelem = conn.size // elemT.LenNodes()
        nqp = elemT.Quadrature()[1].size
        dimC = c.Constitutive().shape[0]
        out = np.zeros((nelem, nqp, dimC))
        lib().gofem_isoparametric_strains(self._h, _d(u), u.size, elemT.KIND, elemT.QuadratureOrder, elemT.NodeDofs, c.ckind, c.p0, c.p1,
                                          conn.ctypes.data_as(_i64p), nelem, _d(out))
        err = self._err()
        if err:
            raise err
        return out

    # ---- device-level extras (femb200.h) -------------------------------------------------------
    def device_handle(self):
        return lib().gofem_device_assembler(self._h)

    def device_ctx(self):
        return lib().gofem_device_ctx(self._h)

    def phase_ms(self):
        out = (C.c_float * N_TIMERS)()
        lib().femb200_phase_ms(self.device_handle(), out)
        return dict(zip(TIMER_NAMES, [float(v) for v in out]))

    def slot_map(self, nelem, nn):
        out = np.zeros((nelem, nn, nn), dtype=np.int64)
        rc = lib().femb200_last_slot_map(self.device_handle(), out.ctypes.data_as(_i64p))
        if rc:
            raise RuntimeError(f"femb200_last_slot_map -> {rc}")
        return out


class Fixity:
    """fem.Fixity (fem.go:175-232)."""

    def __init__(self, modelDofs, numNodes):
        self.modelDofs, self.flags = int(modelDofs), np.zeros(numNodes, dtype=np.uint16)

    def Fix(self, i, fixed):
        self.flags[i] |= (fixed & self.modelDofs)

    def Free(self, i, freed):
        self.flags[i] &= ~np.uint16(freed & self.modelDofs)

    def _get(self, fixed):
        n = lib().gofem_fixity_dofs(self.modelDofs, self.flags.size, self.flags.ctypes.data_as(_u16p), int(fixed), None)
        out = np.zeros(n, dtype=np.int64)
        lib().gofem_fixity_dofs(self.modelDofs, self.flags.size, self.flags.ctypes.data_as(_u16p), int(fixed), out.ctypes.data_as(_i64p))
        return out

    def FreeDofs(self):
        return self._get(False)

    def FixedDofs(self):
        return self._get(True)


# ----------------------------------------------------------------------------------------------- raw C ABI
class RawAssembler:
    """Thin object over the femb200_* C ABI with caller-provided tables and (optionally) device-resident
    inputs -- what the cgo shim does.  Used by bench.py for the inputs-in-HBM timing."""

    def __init__(self, nodes, mdpn, device=0, nodes_device_ptr=None, n_nodes=None):
        L = lib()
        self.ctx = C.c_void_p()
        rc = L.femb200_ctx_create(device, C.byref(self.ctx))
        if rc:
            raise RuntimeError(f"femb200_ctx_create({device}) -> {rc}: no usable sm_100 GPU (no CPU fallback)")
        self.h = C.c_void_p()
        if nodes_device_ptr is not None:
            rc = L.femb200_assembler_create(self.ctx, C.c_void_p(nodes_device_ptr), n_nodes, mdpn, 1, C.byref(self.h))
        else:
            nodes = np.ascontiguousarray(nodes, dtype=np.float64).reshape(-1, 3)
            self._nodes = nodes
            rc = L.femb200_assembler_create(self.ctx, C.c_void_p(nodes.ctypes.data), nodes.shape[0], mdpn, 0, C.byref(self.h))
        if rc:
            raise RuntimeError(f"femb200_assembler_create -> {rc}: {L.femb200_last_cuda_error(self.ctx).decode()}")
        self.mdpn = mdpn

    def make_desc(self, elemT, c, dofmap=None):
        """Evaluates the once-per-call tables through the host mirror (what the Go shim does in Go)."""
        nn = elemT.LenNodes()
        pos, w = elemT.Quadrature()
        N = np.ascontiguousarray(np.stack([elemT.Basis(p) for p in pos]))
        dN = np.ascontiguousarray(np.stack([elemT.BasisDiff(p) for p in pos]))
        Cm = np.ascontiguousarray(c.Constitutive())
        d = dN.shape[1] // nn
        ndn = bin(elemT.Dofs()).count("1")
        dm = np.ascontiguousarray(np.arange(ndn) if dofmap is None else dofmap, dtype=np.int32)
        desc = ElemDesc(d, nn, ndn, len(w), Cm.shape[0], c.BKind(), self.mdpn, 0, _d(N), _d(dN), _d(np.ascontiguousarray(w)), _d(Cm), dm.ctypes.data_as(_i32p))
        desc._keep = (N, dN, w, Cm, dm)
        return desc

    def add(self, desc, conn=None, conn_ptr=None, n_elem=None, is64=False, on_device=False, n_pattern_only=0):
        if conn is not None:
            conn = np.ascontiguousarray(conn)
            is64 = conn.dtype == np.int64
            if not is64:
                conn = np.ascontiguousarray(conn, dtype=np.int32)
            n_elem = conn.size // desc.nn
            conn_ptr = conn.ctypes.data
        return lib().femb200_add_isoparametric_halo(self.h, C.byref(desc), C.c_void_p(conn_ptr), int(is64), int(on_device), n_elem, n_pattern_only)

    def last_error(self):
        k, e, q = C.c_int32(), C.c_int64(), C.c_int32()
        lib().femb200_last_error(self.h, C.byref(k), C.byref(e), C.byref(q))
        return k.value, e.value, q.value

    def slice(self, rows, cols):
        return _csr_slice(self.h, rows, cols)

    def reassemble(self, desc, nodes=None, nodes_device_ptr=None):
        """Numeric-only re-assembly on the cached pattern (femb200_reassemble)."""
        if nodes_device_ptr is not None:
            return lib().femb200_reassemble(self.h, C.byref(desc), C.c_void_p(nodes_device_ptr), 1)
        if nodes is not None:
            nodes = np.ascontiguousarray(nodes, dtype=np.float64)
            return lib().femb200_reassemble(self.h, C.byref(desc), C.c_void_p(nodes.ctypes.data), 0)
        return lib().femb200_reassemble(self.h, C.byref(desc), None, 0)

    def nnz(self):
        return lib().femb200_nnz(self.h)

    def total_dofs(self):
        return lib().femb200_total_dofs(self.h)

    def csr(self):
        n, nnz = self.total_dofs(), self.nnz()
        indptr = np.zeros(n + 1, dtype=np.int64)
        indices = np.zeros(nnz, dtype=np.int32)
        values = np.zeros(nnz, dtype=np.float64)
        rc = lib().femb200_csr_copy(self.h, indptr.ctypes.data, indices.ctypes.data, values.ctypes.data)
        if rc:
            raise RuntimeError(f"femb200_csr_copy -> {rc}")
        return indptr, indices, values

    def bsr(self):
        """femb200_bsr_copy: (blkptr, blkcols, values) -- K at node-block granularity."""
        nblk = lib().femb200_bsr_blocks(self.h)
        if nblk < 0:
            raise RuntimeError("femb200_bsr_copy: K is not a single identity-map AddIsoparametric call")
        n_nodes = self.total_dofs() // self.mdpn
        blkptr = np.zeros(n_nodes + 1, dtype=np.int64)
        blkcols = np.zeros(nblk, dtype=np.int32)
        values = np.zeros(self.nnz(), dtype=np.float64)
        rc = lib().femb200_bsr_copy(self.h, blkptr.ctypes.data, blkcols.ctypes.data, values.ctypes.data)
        if rc:
            raise RuntimeError(f"femb200_bsr_copy -> {rc}")
        return blkptr, blkcols, values

    def add_coo(self, rows, cols, vals):
        rows = np.ascontiguousarray(rows, dtype=np.int64)
        cols = np.ascontiguousarray(cols, dtype=np.int64)
        vals = np.ascontiguousarray(vals, dtype=np.float64)
        return lib().femb200_add_coo(self.h, rows.size, rows.ctypes.data_as(_i64p), cols.ctypes.data_as(_i64p), _d(vals))

    def csr_copy_into(self, indptr_ptr, indices_ptr, values_ptr):
        return lib().femb200_csr_copy(self.h, indptr_ptr, indices_ptr, values_ptr)

    def csr_device(self):
        a, b, c = C.c_void_p(), C.c_void_p(), C.c_void_p()
        lib().femb200_csr_device(self.h, C.byref(a), C.byref(b), C.byref(c))
        return a.value, b.value, c.value

    def phase_ms(self):
        out = (C.c_float * N_TIMERS)()
        lib().femb200_phase_ms(self.h, out)
        return dict(zip(TIMER_NAMES, [float(v) for v in out]))

    def launches(self):
        return lib().femb200_kernel_launches(self.ctx)

    def close(self):
        if self.h:
            lib().femb200_assembler_destroy(self.h)
            self.h = None
        if self.ctx:
            lib().femb200_ctx_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ----------------------------------------------------------------------------------------------- multi-GPU (C ABI)
class DistRank:
    """femb200_dist: rank `rank` of `world` on the context's device -- device-side partition + owner-computes row block.
    nodes / conn may be host arrays or device pointers (nodes_ptr / conn_ptr with n_nodes / n_elem)."""

    def __init__(self, ctx, rank, world, mdpn, nn, nodes=None, conn=None, nodes_ptr=None, conn_ptr=None, n_nodes=None, n_elem=None,
                 conn_is64=False, order_mode=1, handle=None):
        L = lib()
        self.ctx, self.mdpn, self.nn, self.owned_handle = ctx, mdpn, nn, handle is None
        if handle is not None:
            self.h = handle
        else:
            self.h = C.c_void_p()
            if nodes is not None:
                nodes = np.ascontiguousarray(nodes, dtype=np.float64).reshape(-1, 3)
                conn = np.ascontiguousarray(conn)
                conn_is64 = conn.dtype == np.int64
                if not conn_is64:
                    conn = np.ascontiguousarray(conn, dtype=np.int32)
                self._keep = (nodes, conn)
                rc = L.femb200_dist_create(ctx, rank, world, nodes.ctypes.data, 0, nodes.shape[0], mdpn, conn.ctypes.data, int(conn_is64), 0,
                                           conn.size // nn, nn, order_mode, C.byref(self.h))
            else:
                rc = L.femb200_dist_create(ctx, rank, world, nodes_ptr, 1, n_nodes, mdpn, conn_ptr, int(conn_is64), 1, n_elem, nn, order_mode, C.byref(self.h))
            if rc:
                raise RuntimeError(f"femb200_dist_create -> {rc}")

    def info(self):
        out = np.zeros(8, dtype=np.int64)
        ms = C.c_float()
        lib().femb200_dist_info(self.h, out.ctypes.data_as(_i64p), C.byref(ms))
        keys = ("local_nodes", "owned_nodes", "local_elems", "chunk_elems", "halo_elems", "nnz", "global_nodes", "global_elems")
        d = dict(zip(keys, [int(v) for v in out]))
        d["partition_ms"] = float(ms.value)
        return d

    def add(self, desc):
        return lib().femb200_dist_add_isoparametric(self.h, C.byref(desc))

    def last_error(self):
        k, e, q = C.c_int32(), C.c_int64(), C.c_int32()
        lib().femb200_dist_last_error(self.h, C.byref(k), C.byref(e), C.byref(q))
        return k.value, e.value, q.value

    def reset(self):
        return lib().femb200_dist_reset(self.h)

    def assembler(self):
        return lib().femb200_dist_assembler(self.h)

    def local_nodes(self):
        n = self.info()["local_nodes"]
        l2g, owned = np.zeros(n, dtype=np.int32), np.zeros(n, dtype=np.uint8)
        lib().femb200_dist_local_nodes(self.h, l2g.ctypes.data, owned.ctypes.data)
        return l2g, owned

    def row_block(self):
        """(global scalar rows ascending, indptr, indices (global), values) of the rows this rank owns -- host copy, tests only."""
        L = lib()
        ga = self.assembler()
        n, nnz = L.femb200_total_dofs(ga), L.femb200_nnz(ga)
        indptr = np.zeros(n + 1, dtype=np.int64)
        indices = np.zeros(nnz, dtype=np.int32)
        values = np.zeros(nnz, dtype=np.float64)
        rc = L.femb200_csr_copy(ga, indptr.ctypes.data, indices.ctypes.data, values.ctypes.data)
        if rc:
            raise RuntimeError(f"femb200_csr_copy -> {rc}")
        l2g, owned = self.local_nodes()
        m = self.mdpn
        lrows = (np.nonzero(owned)[0][:, None] * m + np.arange(m)[None, :]).ravel()
        grows = (l2g[np.nonzero(owned)[0]].astype(np.int64)[:, None] * m + np.arange(m)[None, :]).ravel()
        lens = indptr[lrows + 1] - indptr[lrows]
        ptr = np.zeros(lrows.size + 1, dtype=np.int64)
        np.cumsum(lens, out=ptr[1:])
        take = np.concatenate([np.arange(indptr[r], indptr[r + 1]) for r in lrows]) if lrows.size else np.zeros(0, dtype=np.int64)
        # rows of nodes this rank does not own must be empty
        other = np.ones(n, dtype=bool)
        other[lrows] = False
        assert np.all((indptr[1:] - indptr[:-1])[other] == 0), "a non-owned row holds entries"
        return grows, ptr, indices[take], values[take]

    def close(self):
        if self.h and self.owned_handle:
            lib().femb200_dist_destroy(self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def device_mesh_bcc_tetra4(ctx, nx, ny, nz, h, jittered=True):
    """femb200_mesh_bcc_tetra4 -> (nodes device pointer, n_nodes, conn device pointer (int32), n_elem); free with lib().femb200_mesh_free."""
    dn, dc = C.c_void_p(), C.c_void_p()
    nn_, ne = C.c_int64(), C.c_int64()
    rc = lib().femb200_mesh_bcc_tetra4(ctx, nx, ny, nz, float(h), int(jittered), C.byref(dn), C.byref(nn_), C.byref(dc), C.byref(ne))
    if rc:
        raise RuntimeError(f"femb200_mesh_bcc_tetra4 -> {rc}")
    return dn.value, nn_.value, dc.value, ne.value
