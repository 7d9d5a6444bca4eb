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

// descriptor builds the once-per-call tables (assembler_isoparametric.go:29-71) that cross the C ABI.
func (ga *GeneralAssembler) descriptor(elemT Isoparametric, c IsoConstituter) (*femb200.Desc, error) {
	C, err := c.Constitutive()
	if err != nil {
		return nil, err
	}
	dimC, _c := C.Dims()
	if _c != dimC {
		return nil, fmt.Errorf("expected constitutive matrix to be square, got %dx%d", dimC, _c)
	}
	nn := elemT.LenNodes()
	d := len(elemT.BasisDiff(r3.Vec{})) / nn
	ndn := elemT.Dofs().Count()
	upg, wpg := elemT.Quadrature()
	if len(upg) == 0 || len(upg) != len(wpg) {
		return nil, fmt.Errorf("bad quadrature result from isoparametric element")
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
		return nil, errors.New("dimsPerNode must be 1, 2 or 3")
	}
	dofMapping, err := mapdofs(ga.dofs, elemT.Dofs())
	if err != nil {
		return nil, err
	}
	if ga.dofs.Count() != ndn {
		panic("bad length") // fem.go:138 + assembler.go:148-150
	}
	for _, v := range dofMapping {
		desc.DofMap = append(desc.DofMap, int32(v))
	}
	dk, ok := c.(DeviceKinded)
	if !ok {
		return nil, errors.New("femb200: user-defined IsoConstituter cannot run on the GPU (no CPU fallback)")
	}
	desc.Kind = femb200.BKind(dk.DeviceBKind())
	return desc, nil
}

// flatten calls getElement Nelem times (fem.go:140-144), checks the element length and the (unsupported) orientation
// vectors in element order, and returns the [Nelem][nn] connectivity.
func flatten(nn, Nelem int, getElement func(i int) (elem []int, xC, yC r3.Vec)) ([]int, error) {
	conn := make([]int, Nelem*nn)
	for i := 0; i < Nelem; i++ {
		element, x, y := getElement(i)
		if len(element) != nn {
			return nil, fmt.Errorf("element #%d of %d nodes expected to be of %d nodes", i, len(element), nn)
		}
		if x != (r3.Vec{}) || y != (r3.Vec{}) {
			return nil, errors.New("arbitrary constitutive orientation not implemented yet")
		}
		copy(conn[i*nn:], element)
	}
	return conn, nil
}

// deviceError formats a device-side element failure with the reference's texts (:96, :98, :102, :106).
func deviceError(err error) error {
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
			panic(fmt.Sprintf("runtime error: index out of range (element #%d)", ee.Elem)) // assembler.go:137: the reference panics on the slice bound
		}
	}
	return err
}

// AddIsoparametric: same signature and error texts as the reference (assembler_isoparametric.go:28-123).
func (ga *GeneralAssembler) AddIsoparametric(elemT Isoparametric, c IsoConstituter, Nelem int, getElement func(i int) (elem []int, xC, yC r3.Vec)) error {
	if elemT == nil || getElement == nil {
		panic("nil argument to AddIsoparametric2")
	}
	desc, err := ga.descriptor(elemT, c)
	if err != nil {
		return err
	}
	conn, err := flatten(desc.NN, Nelem, getElement)
	if err != nil {
		return err
	}
	err = ga.dev.AddIsoparametric(desc, conn, Nelem)
	ga.dirty = true
	return deviceError(err)
}

// AddIsoparametricFlat is the fast ingest path (additive API, SURVEY 8f-3): the connectivity is already a flat
// [Nelem][nn] int32 slice, so no per-element closure call and half the bytes over PCIe.
func (ga *GeneralAssembler) AddIsoparametricFlat(elemT Isoparametric, c IsoConstituter, conn []int32) error {
	if elemT == nil {
		panic("nil argument to AddIsoparametric2")
	}
	desc, err := ga.descriptor(elemT, c)
	if err != nil {
		return err
	}
	if len(conn)%desc.NN != 0 {
		return fmt.Errorf("flat connectivity of %d entries is not a multiple of %d nodes", len(conn), desc.NN)
	}
	err = ga.dev.AddIsoparametricInt32(desc, conn, len(conn)/desc.NN)
	ga.dirty = true
	return deviceError(err)
}

// Reassemble recomputes the values of K on the pattern of the one AddIsoparametric call made so far (additive API): new
// material or quadrature, and new node coordinates when nodes is not nil (same node count).
func (ga *GeneralAssembler) Reassemble(elemT Isoparametric, c IsoConstituter, nodes []r3.Vec) error {
	desc, err := ga.descriptor(elemT, c)
	if err != nil {
		return err
	}
	var p unsafe.Pointer
	if nodes != nil {
		if len(nodes) != len(ga.nodes) {
			return errors.New("femb200: Reassemble needs the same number of nodes")
		}
		p = unsafe.Pointer(&nodes[0])
	}
	err = ga.dev.Reassemble(desc, p)
	if err == nil && nodes != nil {
		ga.nodes = nodes
	}
	ga.dirty = true
	return deviceError(err)
}

// IsoparametricStrains: same signature as the reference (assembler_isoparametric.go:126-213).  The strains of every
// element and quadrature point are computed on the device in one call; strainCallback then runs in element order.
func (ga *GeneralAssembler) IsoparametricStrains(displacements lap.Vector, elemT Isoparametric, c IsoConstituter, Nelem int, getElement func(i int) (elem []int, xC, yC r3.Vec), strainCallback func(iele int, strains []float64)) error {
	nDisp := displacements.Len()
	if nDisp != ga.TotalDofs() {
		return fmt.Errorf("displacements vector length %d does not match total number of dofs %d", nDisp, ga.TotalDofs())
	}
	if elemT == nil || getElement == nil || strainCallback == nil {
		panic("nil argument to AddIsoparametric2")
	}
	desc, err := ga.descriptor(elemT, c)
	if err != nil {
		return err
	}
	conn, err := flatten(desc.NN, Nelem, getElement)
	if err != nil {
		return err
	}
	u := make([]float64, nDisp)
	for i := range u {
		u[i] = displacements.AtVec(i)
	}
	strains, err := ga.dev.Strains(desc, conn, Nelem, u)
	if err != nil {
		return deviceError(err)
	}
	per := desc.NQP * desc.DimC
	for iele := 0; iele < Nelem; iele++ {
		strainCallback(iele, strains[iele*per:(iele+1)*per])
	}
	return nil
}

type elementDofCallback = func(elemIdx int, elemNodes []float64, elemDofs []int) error

// ForEachElement: the reference's host-side iterator (assembler_isoparametric.go:19-21, fem.go:120-153), unchanged: it
// runs user code per element and never touches the device.
func (ga *GeneralAssembler) ForEachElement(elemT Element, spatialDimsPerNode, Nelem int, getElement func(i int) []int, elemCallback elementDofCallback) error {
	if elemT == nil || getElement == nil || elemCallback == nil {
		panic("nil argument to AddIsoparametric2")
	} else if spatialDimsPerNode < 1 || spatialDimsPerNode > 3 {
		return errors.New("dimsPerNode must be 1, 2 or 3")
	}
	dofMapping, err := mapdofs(ga.dofs, elemT.Dofs())
	if err != nil {
		return err
	}
	nn := elemT.LenNodes()
	mdpn := ga.dofs.Count()
	elemNodBacking := make([]float64, spatialDimsPerNode*nn)
	elemDofs := make([]int, mdpn*nn)
	for iele := 0; iele < Nelem; iele++ {
		element := getElement(iele)
		if len(element) != nn {
			return fmt.Errorf("element #%d of %d nodes expected to be of %d nodes", iele, len(element), nn)
		}
		storeElemNode(elemNodBacking, ga.nodes, element, spatialDimsPerNode)
		storeElemDofs(elemDofs, element, dofMapping, mdpn)
		if err := elemCallback(iele, elemNodBacking, elemDofs); err != nil {
			return err
		}
	}
	return nil
}

func storeElemNode(dst []float64, allNodes []r3.Vec, elem []int, dims int) { // assembler.go:118-145
	if len(dst) != dims*len(elem) {
		panic("bad length")
	}
	for i := range elem {
		n := allNodes[elem[i]]
		switch dims {
		case 1:
			dst[i] = n.X
		case 2:
			dst[2*i], dst[2*i+1] = n.X, n.Y
		case 3:
			dst[3*i], dst[3*i+1], dst[3*i+2] = n.X, n.Y, n.Z
		default:
			panic("unhandled dimension")
		}
	}
}

func storeElemDofs(dst, elem, dofmapping []int, modelDofsPerNode int) { // assembler.go:147-158
	if len(dofmapping)*len(elem) != len(dst) {
		panic("bad length")
	}
	for inod, node := range elem {
		offset := inod * len(dofmapping)
		dofstart := modelDofsPerNode * node
		for j, dofoffset := range dofmapping {
			dst[offset+j] = dofstart + dofoffset
		}
	}
}

// ElementInfo (assembler.go:39-77).
type ElementInfo struct {
	nodes []r3.Vec
	dofs  [][]int
}

func (einfo ElementInfo) Nodes() []r3.Vec { return append([]r3.Vec{}, einfo.nodes...) }
func (einfo ElementInfo) Dofs() [][]int   { return einfo.dofs }

func (ga *GeneralAssembler) ElementInfo(elemT Element, element []int) (einfo ElementInfo, _ error) {
	dofmapping, err := ga.DofMapping(elemT)
	if err != nil {
		return ElementInfo{}, err
	}
	modelDofs := ga.dofs.Count()
	einfo.nodes = make([]r3.Vec, len(element))
	backingDofs := make([]int, len(element)*elemT.Dofs().Count())
	einfo.dofs = make([][]int, len(element))
	storeElemDofs(backingDofs, element, dofmapping, modelDofs)
	for i := range element {
		if element[i] < 0 || element[i] >= len(ga.nodes) {
			return ElementInfo{}, errors.New("element node index out of bounds")
		}
		einfo.nodes[i] = ga.nodes[element[i]]
		einfo.dofs[i] = backingDofs[i*len(dofmapping) : (i+1)*len(dofmapping)]
	}
	return einfo, nil
}

// DofMapping (assembler.go:97-116): the same loop as mapdofs, quirk included.
func (ga *GeneralAssembler) DofMapping(e Element) ([]int, error) { return mapdofs(ga.dofs, e.Dofs()) }

// CSR returns K as canonical CSR straight from the device (additive API: for large models this skips lap.Sparse).
func (ga *GeneralAssembler) CSR() (indptr []int64, indices []int32, values []float64, err error) { return ga.dev.CSR() }

// Slice is lap.Slice(ga.Ksolid(), rows, cols) extracted on the device (cols ascending and unique, e.g. Fixity.FreeDofs()).
func (ga *GeneralAssembler) Slice(rows, cols []int) (indptr []int64, indices []int32, values []float64, err error) {
	return ga.dev.Slice(rows, cols)
}

// AddTriplets adds foreign (row, col, value) triplets into K (what AddElement3 does with its beam blocks,
// assembler.go:225-231): pattern union on the device.
func (ga *GeneralAssembler) AddTriplets(rows, cols []int, vals []float64) error {
	ga.dirty = true
	return ga.dev.AddCOO(rows, cols, vals)
}

// Fixity (fem.go:175-232), host-side and unchanged.
type Fixity struct {
	nodes     []DofsFlag
	modelDofs DofsFlag
}

func NewFixity(modelDofs DofsFlag, numNodes int) Fixity {
	return Fixity{nodes: make([]DofsFlag, numNodes), modelDofs: modelDofs}
}

func (f Fixity) Fix(nodeIdx int, fixed DofsFlag)  { f.nodes[nodeIdx] |= (fixed & f.modelDofs) }
func (f Fixity) Free(nodeIdx int, freed DofsFlag) { f.nodes[nodeIdx] &^= (freed & f.modelDofs) }
func (f Fixity) FreeDofs() []int                  { return f.fixity(false) }
func (f Fixity) FixedDofs() []int                 { return f.fixity(true) }

func (f Fixity) fixity(getFixed bool) []int {
	dofsPerNode := f.modelDofs.Count()
	var dofIdx []int
	for i := 0; i < 16; i++ {
		if f.modelDofs&(1<<i) != 0 {
			dofIdx = append(dofIdx, i)
		}
	}
	indices := make([]int, 0, len(f.nodes)*dofsPerNode/2)
	for i, n := range f.nodes {
		for j, dof := range dofIdx {
			if getFixed == (n&(1<<dof) != 0) {
				indices = append(indices, i*dofsPerNode+j)
			}
		}
	}
	return indices
}

// AssembleOnDevices is the multi-GPU form of NewGeneralAssembler + AddIsoparametric (additive API): one process, the
// library runs a host thread per device, partitions the mesh on the devices (Morton order) and leaves every device owning
// a row block of the distributed K (femb200_group_*; rows bit-identical to the single-GPU K).  The returned group hands
// out the row blocks (RowBlock) in the caller's global DOF numbering.
func AssembleOnDevices(devices []int32, nodes []r3.Vec, modelDofs DofsFlag, elemT Isoparametric, c IsoConstituter, Nelem int, getElement func(i int) (elem []int, xC, yC r3.Vec)) (*femb200.Group, error) {
	tmp := &GeneralAssembler{dofs: modelDofs, nodes: nodes}
	desc, err := tmp.descriptor(elemT, c)
	if err != nil {
		return nil, err
	}
	conn, err := flatten(desc.NN, Nelem, getElement)
	if err != nil {
		return nil, err
	}
	var p unsafe.Pointer
	if len(nodes) > 0 {
		p = unsafe.Pointer(&nodes[0])
	}
	g, err := femb200.NewGroup(devices, p, len(nodes), modelDofs.Count(), conn, Nelem, desc.NN, true)
	if err != nil {
		return nil, err
	}
	if err := g.AddIsoparametric(desc); err != nil {
		g.Close()
		return nil, deviceError(err)
	}
	return g, nil
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
