// Package fem: the reference's assembler API (fem.go, assembler.go, assembler_isoparametric.go) with the
// isoparametric stiffness assembly running on a B200 through internal/femb200.  Signatures are the
// reference's, verbatim.  NOT COMPILED in the build image; the C++ mirror host/gofem.hpp is the tested twin.
package fem

import (
	"errors"
	"fmt"
	"math/bits"
	"strconv"
	"unsafe"

	"github.com/soypat/go-fem/internal/femb200"
	"github.com/soypat/lap"
	"gonum.org/v1/gonum/mat"
	"gonum.org/v1/gonum/spatial/r3"
)

type Element interface {
	LenNodes() int
	Dofs() DofsFlag
}

type Isoparametric interface {
	Element
	IsoparametricNodes() []r3.Vec
	Quadrature() (positions []r3.Vec, weights []float64)
	Basis(r3.Vec) []float64
	BasisDiff(r3.Vec) []float64
}

type Constituter interface {
	Constitutive() (mat.Matrix, error)
}

type IsoConstituter interface {
	Constituter
	SetStrainDisplacementMatrix(dstB, elemNodes, dN *mat.Dense, N *mat.VecDense) (scale float64)
}

// DeviceKinded is implemented by the constituters of package solids: it names the device-side B filler.
// A user-defined IsoConstituter (arbitrary Go code) cannot run on the GPU and is rejected: no CPU fallback.
type DeviceKinded interface{ DeviceBKind() int32 }

type DofsFlag uint16

const (
	DofPosX DofsFlag = 1 << iota
	DofPosY
	DofPosZ
	DofRotX
	DofRotY
	DofRotZ
)
const (
	DofPos = DofPosX | DofPosY | DofPosZ
	DofRot = DofRotX | DofRotY | DofRotZ
	Dof6   = DofPos | DofRot
)

func (d DofsFlag) Count() int          { return bits.OnesCount16(uint16(d)) }
func (d DofsFlag) Has(q DofsFlag) bool { return d&q == q }
func (d DofsFlag) String() (s string) {
	if d == 0 {
		return "none"
	}
	for i := 0; d != 0; i++ {
		if d&1 != 0 {
			s += strconv.Itoa(i + 1)
		} else {
			s += "-"
		}
		d >>= 1
	}
	return s
}

type GeneralAssembler struct {
	dev    *femb200.Assembler
	ksolid lap.Sparse // filled lazily from the device CSR by Ksolid()
	dirty  bool
	nodes  []r3.Vec
	dofs   DofsFlag
}

// NewGeneralAssembler: same signature as the reference (assembler.go:21).
func NewGeneralAssembler(nodes []r3.Vec, modelDofs DofsFlag) *GeneralAssembler {
	var p unsafe.Pointer
	if len(nodes) > 0 {
		p = unsafe.Pointer(&nodes[0]) // []r3.Vec is n x {X,Y,Z float64}: no Go pointers inside, safe to pass
	}
	dev, err := femb200.New(0, p, len(nodes), modelDofs.Count())
	if err != nil {
		panic(err) // no GPU: fail loudly
	}
	total := len(nodes) * modelDofs.Count()
	return &GeneralAssembler{dev: dev, ksolid: *lap.NewSparse(total, total), dofs: modelDofs, nodes: nodes}
}

func (ga *GeneralAssembler) TotalDofs() int { return len(ga.nodes) * ga.dofs.Count() }

// Ksolid returns K as *lap.Sparse, built from the device CSR with ONE Accumulate of the nnz unique,
// sorted triplets (instead of Nelem*nd^2 as in the reference).
func (ga *GeneralAssembler) Ksolid() *lap.Sparse {
	if ga.dirty {
		indptr, indices, values, err := ga.dev.CSR()
		if err != nil {
			panic(err)
		}
		n := ga.TotalDofs()
		ga.ksolid = *lap.NewSparse(n, n)
		acc := lap.NewSparseAccum(len(values))
		for r := 0; r < n; r++ {
			for k := indptr[r]; k < indptr[r+1]; k++ {
				acc.I[k], acc.J[k], acc.V[k] = r, int(indices[k]), values[k]
			}
		}
		ga.ksolid.Accumulate(acc)
		ga.dirty = false
	}
	return &ga.ksolid
}

// AddIsoparametric: same signature and error texts as the reference (assembler_isoparametric.go:28-123).
func (ga *GeneralAssembler) AddIsoparametric(elemT Isoparametric, c IsoConstituter, Nelem int, getElement func(i int) (elem []int, xC, yC r3.Vec)) error {
	if elemT == nil || getElement == nil {
		panic("nil argument to AddIsoparametric2")
	}
	C, err := c.Constitutive()
	if err != nil {
		return err
	}
	dimC, _c := C.Dims()
	if _c != dimC {
		return fmt.Errorf("expected constitutive matrix to be square, got %dx%d", dimC, _c)
	}
	nn := elemT.LenNodes()
	d := len(elemT.BasisDiff(r3.Vec{})) / nn
	ndn := elemT.Dofs().Count()
	upg, wpg := elemT.Quadrature()
	if len(upg) == 0 || len(upg) != len(wpg) {
		return fmt.Errorf("bad quadrature result from isoparametric element")
	}
	desc := &femb200.Desc{D: d, NN: nn, NDN: ndn, NQP: len(upg), DimC: dimC, ModelDofsPerNode: ga.dofs.Count(), Wpg: wpg}
	for _, pg := range upg {
		dN := elemT.BasisDiff(pg)
		if len(dN) != d*nn {
			panic("mat: bad matrix dimensions") // what mat.NewDense(d, nn, dN) does in the reference (:70)
		}
		desc.Npg = append(desc.Npg, elemT.Basis(pg)...)
		desc.DNpg = append(desc.DNpg, dN...)
	}
	for i := 0; i < dimC; i++ {
		for j := 0; j < dimC; j++ {
			desc.C = append(desc.C, C.At(i, j))
		}
	}
	if d < 1 || d > 3 {
		return errors.New("dimsPerNode must be 1, 2 or 3")
	}
	dofMapping, err := mapdofs(ga.dofs, elemT.Dofs())
	if err != nil {
		return err
	}
	if ga.dofs.Count() != ndn {
		panic("bad length") // fem.go:138 + assembler.go:148-150
	}
	for _, v := range dofMapping {
		desc.DofMap = append(desc.DofMap, int32(v))
	}
	dk, ok := c.(DeviceKinded)
	if !ok {
		return errors.New("femb200: user-defined IsoConstituter cannot run on the GPU (no CPU fallback)")
	}
	desc.Kind = femb200.BKind(dk.DeviceBKind())
	// fem.go:140-144: one callback per element, flattened for the device
	conn := make([]int, Nelem*nn)
	for i := 0; i < Nelem; i++ {
		element, x, y := getElement(i)
		if len(element) != nn {
			return fmt.Errorf("element #%d of %d nodes expected to be of %d nodes", i, len(element), nn)
		}
		if x != (r3.Vec{}) || y != (r3.Vec{}) {
			return errors.New("arbitrary constitutive orientation not implemented yet")
		}
		copy(conn[i*nn:], element)
	}
	err = ga.dev.AddIsoparametric(desc, conn, Nelem)
	ga.dirty = true
	var ee *femb200.ElemError
	if errors.As(err, &ee) {
		switch ee.Kind {
		case 1:
			return fmt.Errorf("negative determinant of jacobian of element #%d, Check node ordering", ee.Elem)
		case 2:
			return fmt.Errorf("zero determinant of jacobian of element #%d, Check element shape for bad aspect ratio", ee.Elem)
		case 3:
			return fmt.Errorf("error calculating element #%d form factor: matrix singular or near-singular", ee.Elem)
		case 4:
			return fmt.Errorf("NaN scale value returned by SetStrainDisplacementMatrix at element #%d, quad %d", ee.Elem, ee.QP)
		case 8:
			panic(fmt.Sprintf("runtime error: index out of range (element #%d)", ee.Elem))
		}
	}
	return err
}

func mapdofs(modelDofs, elemDofs DofsFlag) (dofs []int, err error) {
	numModelDofs := modelDofs.Count()
	if !modelDofs.Has(elemDofs) {
		return nil, fmt.Errorf("model dofs %s does not contain all element dofs %s", modelDofs.String(), elemDofs.String())
	}
	dofMapping := make([]int, elemDofs.Count())
	idm := 0
	for i := 0; i < numModelDofs; i++ { // sic: the reference's loop bound (only flags contiguous from bit 0 map correctly)
		if modelDofs.Has(1<<i) && elemDofs.Has(1<<i) {
			dofMapping[idm] = i
			idm++
		}
	}
	if idm == 0 || len(dofMapping) == 0 {
		return nil, errors.New("element has empty dof mapping")
	}
	return dofMapping, nil
}
