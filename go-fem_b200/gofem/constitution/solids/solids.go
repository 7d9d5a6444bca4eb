// Package solids: Isotropic / IsotropicConductivity factories as in the reference (isotropic.go, thermal.go),
// each constituter tagged with the device-side B filler (DeviceBKind).  NOT COMPILED in the build image.
package solids

import (
	"math"

	"github.com/soypat/go-fem"
	"gonum.org/v1/gonum/mat"
)

type isoconstituter struct {
	C     mat.Matrix
	bkind int32 // FEMB200_B_*
	err   error
}

func (c isoconstituter) Constitutive() (mat.Matrix, error) { return c.C, c.err }
func (c isoconstituter) DeviceBKind() int32                { return c.bkind }

// SetStrainDisplacementMatrix is never called on the GPU path (the fillers run on the device); it is kept so
// the type satisfies fem.IsoConstituter and host-side post-processing code keeps compiling.
func (c isoconstituter) SetStrainDisplacementMatrix(dstB, elemNod, dN *mat.Dense, N *mat.VecDense) float64 {
	panic("solids: strain-displacement fillers run on the device")
}

type Isotropic struct{ E, Poisson float64 }

func (m Isotropic) ShearModulus() float64 { return m.E / (2 + 2*m.Poisson) }
func (m Isotropic) lameParam2() float64   { return m.E * m.Poisson / ((1 + m.Poisson) * (1 - 2*m.Poisson)) }

func (m Isotropic) Constitutive() (mat.Matrix, error) {
	G, l := m.ShearModulus(), m.lameParam2()
	if math.IsNaN(l) || math.IsInf(l, 0) {
		return nil, mat.Condition((1 + m.Poisson) * (1 - 2*m.Poisson))
	}
	return mat.NewDense(6, 6, []float64{
		l + 2*G, l, l, 0, 0, 0,
		l, l + 2*G, l, 0, 0, 0,
		l, l, l + 2*G, 0, 0, 0,
		0, 0, 0, G, 0, 0,
		0, 0, 0, 0, G, 0,
		0, 0, 0, 0, 0, G}), nil
}

func (m Isotropic) Solid3D() fem.IsoConstituter {
	c, err := m.Constitutive()
	return isoconstituter{C: c, err: err, bkind: 0}
}

func (m Isotropic) Axisymmetric() fem.IsoConstituter {
	nu := m.Poisson
	f, g := nu/(1-nu), (1-2*nu)/(2*(1-nu))
	factor := m.E * (1 - nu) / ((1 + nu) * (1 - 2*nu))
	f, g = f*factor, g*factor
	return isoconstituter{C: mat.NewDense(4, 4, []float64{factor, f, f, 0, f, factor, f, 0, f, f, factor, 0, 0, 0, 0, g}), bkind: 2}
}

func (m Isotropic) PlaneStess() fem.IsoConstituter { // sic
	factor := m.E / (1 - m.Poisson*m.Poisson)
	nu := m.Poisson
	return isoconstituter{C: mat.NewDense(3, 3, []float64{factor, factor * nu, 0, factor * nu, factor, 0, 0, 0, factor * (1 - nu) / 2}), bkind: 1}
}

func (m Isotropic) PlaneStrain() fem.IsoConstituter {
	nu := m.Poisson
	factor := m.E / (1 + nu) / (1 - 2*nu)
	nuf := nu * factor
	return isoconstituter{C: mat.NewDense(3, 3, []float64{factor * (1 - nu), nuf, nuf, nuf, factor * (1 - nu), nuf, nuf, nuf, factor * (1 - 2*nu)}), bkind: 1}
}

type IsotropicConductivity struct{ K float64 }

func (k IsotropicConductivity) Solid3D() fem.IsoConstituter {
	return isoconstituter{C: mat.NewDiagDense(3, []float64{k.K, k.K, k.K}), bkind: 3}
}
func (k IsotropicConductivity) Plane() fem.IsoConstituter {
	return isoconstituter{C: mat.NewDiagDense(2, []float64{k.K, k.K}), bkind: 3}
}
func (k IsotropicConductivity) Axisymmetric() fem.IsoConstituter {
	return isoconstituter{C: mat.NewDiagDense(2, []float64{k.K, k.K}), bkind: 4}
}
