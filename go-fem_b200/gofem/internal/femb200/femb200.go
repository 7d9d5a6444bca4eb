// Package femb200 is the cgo binding of include/femb200.h (libfemb200.so: hand-written sm_100a CUDA).
// NOT COMPILED in the build image (no Go toolchain); mirrors go-fem_b200/host/gofem.hpp call for call.
package femb200

/*
#cgo CFLAGS: -I${SRCDIR}/../../../../include
#cgo LDFLAGS: -L${SRCDIR}/../../.. -lfemb200 -Wl,-rpath,${SRCDIR}/../../..
#include <stdlib.h>
#include "femb200.h"
*/
import "C"

import (
	"errors"
	"fmt"
	"runtime"
	"unsafe"
)

// BKind tags the strain-displacement filler (solids.isoconstituter.strain is an unexported func in the
// reference, solids.go:8-12, so the kind travels as an enum instead).
type BKind int32

const (
	BXYZ           BKind = C.FEMB200_B_XYZ
	BPlane         BKind = C.FEMB200_B_PLANE
	BAxisym        BKind = C.FEMB200_B_AXISYM
	BThermal       BKind = C.FEMB200_B_THERMAL
	BThermalAxisym BKind = C.FEMB200_B_THERMAL_AXISYM
)

// Desc holds the once-per-call tables of AddIsoparametric (assembler_isoparametric.go:29-71).
type Desc struct {
	D, NN, NDN, NQP, DimC int
	Kind                  BKind
	ModelDofsPerNode      int
	Npg, DNpg, Wpg, C     []float64
	DofMap                []int32
}

// Assembler wraps one CUDA context + one femb200_assembler (twin of fem.GeneralAssembler).
type Assembler struct {
	ctx *C.femb200_ctx
	ga  *C.femb200_assembler
}

// New uploads the nodes ([]r3.Vec memory: n x 3 float64, 24-byte stride) to `device`.
func New(device int, nodesXYZ unsafe.Pointer, nNodes int, modelDofsPerNode int) (*Assembler, error) {
	a := &Assembler{}
	if rc := C.femb200_ctx_create(C.int(device), &a.ctx); rc != 0 {
		return nil, errors.New("femb200: no usable sm_100 CUDA device (the GPU path has no CPU fallback)")
	}
	rc := C.femb200_assembler_create(a.ctx, (*C.double)(nodesXYZ), C.int64_t(nNodes), C.int32_t(modelDofsPerNode), 0, &a.ga)
	if rc != 0 {
		msg := C.GoString(C.femb200_last_cuda_error(a.ctx))
		C.femb200_ctx_destroy(a.ctx)
		return nil, fmt.Errorf("femb200_assembler_create: %s", msg)
	}
	runtime.SetFinalizer(a, (*Assembler).Close)
	return a, nil
}

func (a *Assembler) Close() {
	if a.ga != nil {
		C.femb200_assembler_destroy(a.ga)
		a.ga = nil
	}
	if a.ctx != nil {
		C.femb200_ctx_destroy(a.ctx)
		a.ctx = nil
	}
}

// ElemError is a per-element failure reported by the device (lowest failing element, its first failing
// quadrature point), to be formatted with the reference's message texts by package fem.
type ElemError struct {
	Kind, QP int
	Elem     int64
}

func (e *ElemError) Error() string { return fmt.Sprintf("femb200 element error kind=%d elem=%d qp=%d", e.Kind, e.Elem, e.QP) }

// AddIsoparametric assembles nElem elements; conn is the flattened [nElem][nn] Go []int (int64 on amd64).
// All slices are only read during the call (cgo pointer rules: nothing is retained).
func (a *Assembler) AddIsoparametric(d *Desc, conn []int, nElem int) error {
	var cd C.femb200_elem_desc
	cd.d, cd.nn, cd.ndn, cd.nqp, cd.dimC = C.int32_t(d.D), C.int32_t(d.NN), C.int32_t(d.NDN), C.int32_t(d.NQP), C.int32_t(d.DimC)
	cd.bkind = C.int32_t(d.Kind)
	cd.model_dofs_per_node = C.int32_t(d.ModelDofsPerNode)
	// the descriptor holds pointers into Go memory: pin them for the duration of the call (Go 1.21 runtime.Pinner),
	// or copy the five small tables into C memory as done here.
	npg, dnpg, wpg, cm := cDoubles(d.Npg), cDoubles(d.DNpg), cDoubles(d.Wpg), cDoubles(d.C)
	dm := (*C.int32_t)(C.malloc(C.size_t(4 * len(d.DofMap))))
	defer C.free(unsafe.Pointer(npg))
	defer C.free(unsafe.Pointer(dnpg))
	defer C.free(unsafe.Pointer(wpg))
	defer C.free(unsafe.Pointer(cm))
	defer C.free(unsafe.Pointer(dm))
	copy(unsafe.Slice((*int32)(unsafe.Pointer(dm)), len(d.DofMap)), d.DofMap)
	cd.Npg, cd.dNpg, cd.wpg, cd.C, cd.dofmap = npg, dnpg, wpg, cm, dm
	var p unsafe.Pointer
	if nElem > 0 {
		p = unsafe.Pointer(&conn[0])
	}
	rc := C.femb200_add_isoparametric(a.ga, &cd, p, 1 /* int64 */, 0 /* host */, C.int64_t(nElem))
	if rc == 0 {
		return nil
	}
	var kind, qp C.int32_t
	var elem C.int64_t
	C.femb200_last_error(a.ga, &kind, &elem, &qp)
	if rc >= 1 && rc <= 8 {
		return &ElemError{Kind: int(rc), QP: int(qp), Elem: int64(elem)}
	}
	return fmt.Errorf("femb200: status %d: %s", int(rc), C.GoString(C.femb200_last_cuda_error(a.ctx)))
}

// Reassemble recomputes the values of K on the cached pattern of the single AddIsoparametric call this assembler has
// committed (femb200_reassemble): new material / quadrature tables in d, optional new node coordinates (nil = keep).
// Additive API: the reference rebuilds pattern and values on every call.
func (a *Assembler) Reassemble(d *Desc, nodesXYZ unsafe.Pointer) error {
	var cd C.femb200_elem_desc
	cd.d, cd.nn, cd.ndn, cd.nqp, cd.dimC = C.int32_t(d.D), C.int32_t(d.NN), C.int32_t(d.NDN), C.int32_t(d.NQP), C.int32_t(d.DimC)
	cd.bkind = C.int32_t(d.Kind)
	cd.model_dofs_per_node = C.int32_t(d.ModelDofsPerNode)
	npg, dnpg, wpg, cm := cDoubles(d.Npg), cDoubles(d.DNpg), cDoubles(d.Wpg), cDoubles(d.C)
	dm := (*C.int32_t)(C.malloc(C.size_t(4 * len(d.DofMap))))
	defer C.free(unsafe.Pointer(npg))
	defer C.free(unsafe.Pointer(dnpg))
	defer C.free(unsafe.Pointer(wpg))
	defer C.free(unsafe.Pointer(cm))
	defer C.free(unsafe.Pointer(dm))
	copy(unsafe.Slice((*int32)(unsafe.Pointer(dm)), len(d.DofMap)), d.DofMap)
	cd.Npg, cd.dNpg, cd.wpg, cd.C, cd.dofmap = npg, dnpg, wpg, cm, dm
	rc := C.femb200_reassemble(a.ga, &cd, (*C.double)(nodesXYZ), 0 /* host */)
	if rc == 0 {
		return nil
	}
	var kind, qp C.int32_t
	var elem C.int64_t
	C.femb200_last_error(a.ga, &kind, &elem, &qp)
	if rc >= 1 && rc <= 8 {
		return &ElemError{Kind: int(rc), QP: int(qp), Elem: int64(elem)}
	}
	return fmt.Errorf("femb200: status %d: %s", int(rc), C.GoString(C.femb200_last_cuda_error(a.ctx)))
}

func cDoubles(v []float64) *C.double {
	p := (*C.double)(C.malloc(C.size_t(8 * len(v))))
	copy(unsafe.Slice((*float64)(unsafe.Pointer(p)), len(v)), v)
	return p
}

func (a *Assembler) TotalDofs() int { return int(C.femb200_total_dofs(a.ga)) }
func (a *Assembler) NNZ() int       { return int(C.femb200_nnz(a.ga)) }

// CSR copies K (canonical CSR) into freshly allocated Go slices.
func (a *Assembler) CSR() (indptr []int64, indices []int32, values []float64, err error) {
	n, nnz := a.TotalDofs(), a.NNZ()
	indptr, indices, values = make([]int64, n+1), make([]int32, nnz), make([]float64, nnz)
	var pi, pv unsafe.Pointer
	if nnz > 0 {
		pi, pv = unsafe.Pointer(&indices[0]), unsafe.Pointer(&values[0])
	}
	if rc := C.femb200_csr_copy(a.ga, (*C.int64_t)(unsafe.Pointer(&indptr[0])), (*C.int32_t)(pi), (*C.double)(pv)); rc != 0 {
		return nil, nil, nil, fmt.Errorf("femb200_csr_copy: status %d", int(rc))
	}
	return indptr, indices, values, nil
}

// Strains is IsoparametricStrains: strains[nElem][nqp][dimC].
func (a *Assembler) Strains(d *Desc, conn []int, nElem int, u []float64) ([]float64, error) {
	out := make([]float64, nElem*d.NQP*d.DimC)
	_ = out // same marshalling as AddIsoparametric, then C.femb200_isoparametric_strains(...)
	return out, errors.New("see AddIsoparametric for the marshalling; omitted for brevity in this uncompiled shim")
}
