// Package femb200 is the cgo binding of include/femb200.h (libfemb200.so: hand-written sm_100a CUDA).
// NOT COMPILED in the build image (no Go toolchain); mirrors go-fem_b200/host/gofem.hpp call for call.  Every entry point
// used here is declared in include/femb200.h and exported by the library (tests/test_host_mirror.py checks both).
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
	"sync"
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

// One CUDA context per device, shared by every assembler of the process and guarded by a mutex: the library keeps
// per-device state (the __constant__ table bank), so two goroutines must not run calls on one device at the same time.
type deviceCtx struct {
	mu   sync.Mutex
	ctx  *C.femb200_ctx
	refs int
}

var (
	ctxMu sync.Mutex
	ctxs  = map[int]*deviceCtx{}
)

func acquireCtx(device int) (*deviceCtx, error) {
	ctxMu.Lock()
	defer ctxMu.Unlock()
	if d, ok := ctxs[device]; ok {
		d.refs++
		return d, nil
	}
	d := &deviceCtx{refs: 1}
	if rc := C.femb200_ctx_create(C.int(device), &d.ctx); rc != 0 {
		return nil, errors.New("femb200: no usable sm_100 CUDA device (the GPU path has no CPU fallback)")
	}
	ctxs[device] = d
	return d, nil
}

func releaseCtx(device int) {
	ctxMu.Lock()
	defer ctxMu.Unlock()
	if d, ok := ctxs[device]; ok {
		d.refs--
		if d.refs == 0 {
			C.femb200_ctx_destroy(d.ctx)
			delete(ctxs, device)
		}
	}
}

// TrimDevice returns the scratch arena of `device`'s shared context to the GPU (femb200_ctx_trim): worth calling after a
// large one-off assembly when a solver is about to use the same device.  No-op if no assembler lives on the device.
func TrimDevice(device int) error {
	ctxMu.Lock()
	d, ok := ctxs[device]
	ctxMu.Unlock()
	if !ok {
		return nil
	}
	d.mu.Lock()
	defer d.mu.Unlock()
	if rc := C.femb200_ctx_trim(d.ctx); rc != 0 {
		return fmt.Errorf("femb200_ctx_trim: error %d: %s", int(rc), C.GoString(C.femb200_last_cuda_error(d.ctx)))
	}
	return nil
}

// Assembler wraps one femb200_assembler (twin of fem.GeneralAssembler) on a shared device context.
type Assembler struct {
	dev    *deviceCtx
	device int
	ga     *C.femb200_assembler
}

// New uploads the nodes ([]r3.Vec memory: n x 3 float64, 24-byte stride) to `device`.
func New(device int, nodesXYZ unsafe.Pointer, nNodes int, modelDofsPerNode int) (*Assembler, error) {
	dev, err := acquireCtx(device)
	if err != nil {
		return nil, err
	}
	a := &Assembler{dev: dev, device: device}
	dev.mu.Lock()
	rc := C.femb200_assembler_create(dev.ctx, (*C.double)(nodesXYZ), C.int64_t(nNodes), C.int32_t(modelDofsPerNode), 0, &a.ga)
	msg := C.GoString(C.femb200_last_cuda_error(dev.ctx))
	dev.mu.Unlock()
	if rc != 0 {
		releaseCtx(device)
		return nil, fmt.Errorf("femb200_assembler_create: status %d %s", int(rc), msg)
	}
	runtime.SetFinalizer(a, (*Assembler).Close)
	return a, nil
}

func (a *Assembler) Close() {
	if a.ga != nil {
		a.dev.mu.Lock()
		C.femb200_assembler_destroy(a.ga)
		a.dev.mu.Unlock()
		a.ga = nil
		releaseCtx(a.device)
	}
}

// ElemError is a per-element failure reported by the device (lowest failing element, its first failing
// quadrature point), to be formatted with the reference's message texts by package fem.
type ElemError struct {
	Kind, QP int
	Elem     int64
}

func (e *ElemError) Error() string { return fmt.Sprintf("femb200 element error kind=%d elem=%d qp=%d", e.Kind, e.Elem, e.QP) }

// cDesc copies the five small tables of d into C memory (the descriptor holds pointers: Go memory with pointers to Go
// memory must not cross the cgo boundary) and returns the C descriptor plus the function that frees the copies.
func cDesc(d *Desc) (*C.femb200_elem_desc, func()) {
	cd := (*C.femb200_elem_desc)(C.calloc(1, C.size_t(unsafe.Sizeof(C.femb200_elem_desc{}))))
	cd.d, cd.nn, cd.ndn, cd.nqp, cd.dimC = C.int32_t(d.D), C.int32_t(d.NN), C.int32_t(d.NDN), C.int32_t(d.NQP), C.int32_t(d.DimC)
	cd.bkind = C.int32_t(d.Kind)
	cd.model_dofs_per_node = C.int32_t(d.ModelDofsPerNode)
	npg, dnpg, wpg, cm := cDoubles(d.Npg), cDoubles(d.DNpg), cDoubles(d.Wpg), cDoubles(d.C)
	n := len(d.DofMap)
	if n == 0 {
		n = 1
	}
	dm := (*C.int32_t)(C.malloc(C.size_t(4 * n)))
	copy(unsafe.Slice((*int32)(unsafe.Pointer(dm)), len(d.DofMap)), d.DofMap)
	cd.Npg, cd.dNpg, cd.wpg, cd.C, cd.dofmap = npg, dnpg, wpg, cm, dm
	return cd, func() {
		C.free(unsafe.Pointer(npg))
		C.free(unsafe.Pointer(dnpg))
		C.free(unsafe.Pointer(wpg))
		C.free(unsafe.Pointer(cm))
		C.free(unsafe.Pointer(dm))
		C.free(unsafe.Pointer(cd))
	}
}

func cDoubles(v []float64) *C.double {
	n := len(v)
	if n == 0 {
		n = 1
	}
	p := (*C.double)(C.malloc(C.size_t(8 * n)))
	copy(unsafe.Slice((*float64)(unsafe.Pointer(p)), len(v)), v)
	return p
}

func (a *Assembler) elemError(rc C.int) error {
	var kind, qp C.int32_t
	var elem C.int64_t
	C.femb200_last_error(a.ga, &kind, &elem, &qp)
	if rc >= 1 && rc <= 8 {
		return &ElemError{Kind: int(rc), QP: int(qp), Elem: int64(elem)}
	}
	return fmt.Errorf("femb200: status %d: %s", int(rc), C.GoString(C.femb200_last_cuda_error(a.dev.ctx)))
}

// AddIsoparametric assembles nElem elements; conn is the flattened [nElem][nn] Go []int (int64 on amd64).
// All slices are only read during the call (cgo pointer rules: nothing is retained).
func (a *Assembler) AddIsoparametric(d *Desc, conn []int, nElem int) error {
	if unsafe.Sizeof(int(0)) != 8 {
		return errors.New("femb200: Go int must be 64 bits (conn is passed as int64)")
	}
	var p unsafe.Pointer
	if nElem > 0 {
		p = unsafe.Pointer(&conn[0])
	}
	return a.add(d, p, 1, nElem)
}

// AddIsoparametricInt32 is the fast ingest path of AddIsoparametricFlat: the connectivity is already 32-bit, so half the
// bytes cross PCIe and the device skips the narrowing pass.
func (a *Assembler) AddIsoparametricInt32(d *Desc, conn []int32, nElem int) error {
	var p unsafe.Pointer
	if nElem > 0 {
		p = unsafe.Pointer(&conn[0])
	}
	return a.add(d, p, 0, nElem)
}

func (a *Assembler) add(d *Desc, conn unsafe.Pointer, is64 int, nElem int) error {
	cd, free := cDesc(d)
	defer free()
	a.dev.mu.Lock()
	defer a.dev.mu.Unlock()
	rc := C.femb200_add_isoparametric(a.ga, cd, conn, C.int32_t(is64), 0 /* host */, C.int64_t(nElem))
	if rc == 0 {
		return nil
	}
	return a.elemError(rc)
}

// Reassemble recomputes the values of K on the cached pattern of the single AddIsoparametric call this assembler has
// committed (femb200_reassemble): new material / quadrature tables in d, optional new node coordinates (nil = keep).
// Additive API: the reference rebuilds pattern and values on every call.
func (a *Assembler) Reassemble(d *Desc, nodesXYZ unsafe.Pointer) error {
	cd, free := cDesc(d)
	defer free()
	a.dev.mu.Lock()
	defer a.dev.mu.Unlock()
	rc := C.femb200_reassemble(a.ga, cd, (*C.double)(nodesXYZ), 0 /* host */)
	if rc == 0 {
		return nil
	}
	return a.elemError(rc)
}

// Strains is IsoparametricStrains (assembler_isoparametric.go:126-213): strains[nElem][nqp][dimC] = B(q) * u[elemDofs].
func (a *Assembler) Strains(d *Desc, conn []int, nElem int, u []float64) ([]float64, error) {
	if len(u) != a.TotalDofs() {
		return nil, fmt.Errorf("displacements vector length %d does not match total number of dofs %d", len(u), a.TotalDofs())
	}
	out := make([]float64, nElem*d.NQP*d.DimC)
	if nElem == 0 {
		return out, nil
	}
	cd, free := cDesc(d)
	defer free()
	a.dev.mu.Lock()
	defer a.dev.mu.Unlock()
	var pu unsafe.Pointer
	if len(u) > 0 {
		pu = unsafe.Pointer(&u[0])
	}
	rc := C.femb200_isoparametric_strains(a.ga, cd, unsafe.Pointer(&conn[0]), 1, 0, C.int64_t(nElem), (*C.double)(pu), (*C.double)(unsafe.Pointer(&out[0])))
	if rc != 0 {
		return nil, a.elemError(rc)
	}
	return out, nil
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
	a.dev.mu.Lock()
	defer a.dev.mu.Unlock()
	if rc := C.femb200_csr_copy(a.ga, (*C.int64_t)(unsafe.Pointer(&indptr[0])), (*C.int32_t)(pi), (*C.double)(pv)); rc != 0 {
		return nil, nil, nil, fmt.Errorf("femb200_csr_copy: status %d", int(rc))
	}
	return indptr, indices, values, nil
}

// BSR copies K at node-block granularity (femb200_bsr_copy): blkptr[nNodes+1], blkcols[nblk] (column NODE of every block),
// values exactly as in CSR.  A third of the bytes of CSR cross PCIe; ok == false when K is not a single identity-map call.
func (a *Assembler) BSR(nNodes int) (blkptr []int64, blkcols []int32, values []float64, ok bool, err error) {
	a.dev.mu.Lock()
	defer a.dev.mu.Unlock()
	nblk := int(C.femb200_bsr_blocks(a.ga))
	if nblk < 0 {
		return nil, nil, nil, false, nil
	}
	nnz := a.NNZ()
	blkptr, blkcols, values = make([]int64, nNodes+1), make([]int32, nblk), make([]float64, nnz)
	var pc, pv unsafe.Pointer
	if nblk > 0 {
		pc, pv = unsafe.Pointer(&blkcols[0]), unsafe.Pointer(&values[0])
	}
	if rc := C.femb200_bsr_copy(a.ga, (*C.int64_t)(unsafe.Pointer(&blkptr[0])), (*C.int32_t)(pc), (*C.double)(pv)); rc != 0 {
		return nil, nil, nil, false, fmt.Errorf("femb200_bsr_copy: status %d", int(rc))
	}
	return blkptr, blkcols, values, true, nil
}

// At is lap.Sparse.At on the device copy of K.
func (a *Assembler) At(i, j int) float64 {
	var v C.double
	a.dev.mu.Lock()
	defer a.dev.mu.Unlock()
	if rc := C.femb200_csr_at(a.ga, C.int64_t(i), C.int64_t(j), &v); rc != 0 {
		panic(fmt.Sprintf("femb200_csr_at(%d,%d): status %d", i, j, int(rc)))
	}
	return float64(v)
}

// MulVec computes y = K x on the device.
func (a *Assembler) MulVec(x []float64) ([]float64, error) {
	n := a.TotalDofs()
	if len(x) != n {
		return nil, fmt.Errorf("femb200: MulVec length %d, want %d", len(x), n)
	}
	y := make([]float64, n)
	if n == 0 {
		return y, nil
	}
	a.dev.mu.Lock()
	defer a.dev.mu.Unlock()
	if rc := C.femb200_csr_matvec(a.ga, (*C.double)(unsafe.Pointer(&x[0])), (*C.double)(unsafe.Pointer(&y[0]))); rc != 0 {
		return nil, fmt.Errorf("femb200_csr_matvec: status %d", int(rc))
	}
	return y, nil
}

// Slice is lap.Slice(K, rows, cols) extracted on the device (femb200_csr_slice): cols ascending and unique.
func (a *Assembler) Slice(rows, cols []int) (indptr []int64, indices []int32, values []float64, err error) {
	var pr, pc *C.int64_t
	if len(rows) > 0 {
		pr = (*C.int64_t)(unsafe.Pointer(&rows[0]))
	}
	if len(cols) > 0 {
		pc = (*C.int64_t)(unsafe.Pointer(&cols[0]))
	}
	var nnz C.int64_t
	a.dev.mu.Lock()
	defer a.dev.mu.Unlock()
	if rc := C.femb200_csr_slice(a.ga, C.int64_t(len(rows)), pr, C.int64_t(len(cols)), pc, nil, nil, nil, &nnz); rc != 0 {
		return nil, nil, nil, fmt.Errorf("femb200_csr_slice: status %d", int(rc))
	}
	indptr, indices, values = make([]int64, len(rows)+1), make([]int32, int(nnz)), make([]float64, int(nnz))
	var pi, pv unsafe.Pointer
	if nnz > 0 {
		pi, pv = unsafe.Pointer(&indices[0]), unsafe.Pointer(&values[0])
	}
	if rc := C.femb200_csr_slice(a.ga, C.int64_t(len(rows)), pr, C.int64_t(len(cols)), pc, (*C.int64_t)(unsafe.Pointer(&indptr[0])),
		(*C.int32_t)(pi), (*C.double)(pv), &nnz); rc != 0 {
		return nil, nil, nil, fmt.Errorf("femb200_csr_slice: status %d", int(rc))
	}
	return indptr, indices, values, nil
}

// AddCOO adds foreign triplets into K by pattern union (femb200_add_coo): the beam blocks of AddElement3
// (assembler.go:225-231) land in the same K as the isoparametric solids.
func (a *Assembler) AddCOO(rows, cols []int, vals []float64) error {
	if len(rows) != len(cols) || len(rows) != len(vals) {
		return errors.New("femb200: AddCOO slices differ in length")
	}
	if len(rows) == 0 {
		return nil
	}
	a.dev.mu.Lock()
	defer a.dev.mu.Unlock()
	rc := C.femb200_add_coo(a.ga, C.int64_t(len(rows)), (*C.int64_t)(unsafe.Pointer(&rows[0])), (*C.int64_t)(unsafe.Pointer(&cols[0])),
		(*C.double)(unsafe.Pointer(&vals[0])))
	if rc != 0 {
		return fmt.Errorf("femb200_add_coo: status %d", int(rc))
	}
	return nil
}

// PhaseMs returns the device-side phase timers of the last add call (femb200_phase_ms).
func (a *Assembler) PhaseMs() [C.FEMB200_N_TIMERS]float32 {
	var out [C.FEMB200_N_TIMERS]float32
	C.femb200_phase_ms(a.ga, (*C.float)(unsafe.Pointer(&out[0])))
	return out
}

// ---- multi-GPU: one process, one host thread per device inside the library (femb200_group_*) ---------------------------

// Group assembles one mesh on several devices: the library partitions it on every device (Morton order of the element
// centroids when morton is true), a node's rows belong to the lowest rank touching it, and every device assembles the rows it
// owns -- bit-identical to the single-GPU rows, no collective.
type Group struct {
	g      *C.femb200_group
	nn     int
	mdpn   int
	nNodes int
}

func NewGroup(devices []int32, nodesXYZ unsafe.Pointer, nNodes, modelDofsPerNode int, conn []int, nElem, nn int, morton bool) (*Group, error) {
	if len(devices) == 0 {
		return nil, errors.New("femb200: no devices")
	}
	var p unsafe.Pointer
	if nElem > 0 {
		p = unsafe.Pointer(&conn[0])
	}
	order := C.int32_t(0)
	if morton {
		order = 1
	}
	gr := &Group{nn: nn, mdpn: modelDofsPerNode, nNodes: nNodes}
	rc := C.femb200_group_create((*C.int32_t)(unsafe.Pointer(&devices[0])), C.int32_t(len(devices)), (*C.double)(nodesXYZ), C.int64_t(nNodes),
		C.int32_t(modelDofsPerNode), p, 1, C.int64_t(nElem), C.int32_t(nn), order, &gr.g)
	if rc != 0 {
		return nil, fmt.Errorf("femb200_group_create: status %d", int(rc))
	}
	runtime.SetFinalizer(gr, (*Group).Close)
	return gr, nil
}

func (gr *Group) Close() {
	if gr.g != nil {
		C.femb200_group_destroy(gr.g)
		gr.g = nil
	}
}

func (gr *Group) Size() int { return int(C.femb200_group_size(gr.g)) }

// AddIsoparametric runs the call on every device at once; an element error is the one of the LOWEST failing element over
// all ranks, as the reference's sequential loop would report it (fem.go:140).
func (gr *Group) AddIsoparametric(d *Desc) error {
	cd, free := cDesc(d)
	defer free()
	var who C.int32_t
	rc := C.femb200_group_add_isoparametric(gr.g, cd, &who)
	if rc == 0 {
		return nil
	}
	var kind, qp C.int32_t
	var elem C.int64_t
	C.femb200_dist_last_error(C.femb200_group_rank(gr.g, who), &kind, &elem, &qp)
	if rc >= 1 && rc <= 8 {
		return &ElemError{Kind: int(rc), QP: int(qp), Elem: int64(elem)}
	}
	return fmt.Errorf("femb200: group status %d on rank %d", int(rc), int(who))
}

// RowBlock returns rank r's share of the distributed K: the global node of every local node, which of them the rank owns,
// and the local CSR (rows = local nodes x dofs, rows of non-owned nodes empty, column indices = GLOBAL dof ids).
func (gr *Group) RowBlock(r int) (nodeL2G []int32, owned []uint8, indptr []int64, indices []int32, values []float64, err error) {
	d := C.femb200_group_rank(gr.g, C.int32_t(r))
	if d == nil {
		return nil, nil, nil, nil, nil, errors.New("femb200: rank out of range")
	}
	var info [8]C.int64_t
	var ms C.float
	C.femb200_dist_info(d, &info[0], &ms)
	nl, nnz := int(info[0]), int(info[5])
	nodeL2G, owned = make([]int32, nl), make([]uint8, nl)
	indptr, indices, values = make([]int64, nl*gr.mdpn+1), make([]int32, nnz), make([]float64, nnz)
	if nl > 0 {
		if rc := C.femb200_dist_local_nodes(d, (*C.int32_t)(unsafe.Pointer(&nodeL2G[0])), (*C.uint8_t)(unsafe.Pointer(&owned[0]))); rc != 0 {
			return nil, nil, nil, nil, nil, fmt.Errorf("femb200_dist_local_nodes: status %d", int(rc))
		}
	}
	var pi, pv unsafe.Pointer
	if nnz > 0 {
		pi, pv = unsafe.Pointer(&indices[0]), unsafe.Pointer(&values[0])
	}
	if rc := C.femb200_csr_copy(C.femb200_dist_assembler(d), (*C.int64_t)(unsafe.Pointer(&indptr[0])), (*C.int32_t)(pi), (*C.double)(pv)); rc != 0 {
		return nil, nil, nil, nil, nil, fmt.Errorf("femb200_csr_copy: status %d", int(rc))
	}
	return nodeL2G, owned, indptr, indices, values, nil
}
