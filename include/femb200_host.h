/*
 * femb200_host.h -- C view of the C++ host mirror (go-fem_b200/host/gofem.hpp).
 *
 * NOT what a Go build binds (that is femb200.h).  The Go toolchain is absent from the build image, so
 * the host layer of the reference (packages fem, elements, constitution/solids) is mirrored in C++;
 * these wrappers let the Python test/bench harness drive that mirror exactly as the reference's own
 * tests drive the Go API: NewGeneralAssembler / AddIsoparametric(elemT, constituter, Nelem, getElement)
 * / Ksolid.  Element and constituter arguments are passed by kind + fields.
 */
#ifndef FEMB200_HOST_H
#define FEMB200_HOST_H
#include <stdint.h>
#include "femb200.h"
#ifdef __cplusplus
extern "C" {
#endif

/* elements.<Type>{QuadratureOrder, NodeDofs} (elements/<type>.go) */
enum gofem_elem_kind {
  GOFEM_TETRA4 = 0, GOFEM_TETRA10 = 1, GOFEM_HEXA8 = 2, GOFEM_HEXA20 = 3,
  GOFEM_QUAD4 = 4, GOFEM_QUAD8 = 5, GOFEM_TRIANGLE3 = 6, GOFEM_TRIANGLE6 = 7
};
/* solids.Isotropic{E,Poisson}.{Solid3D,Axisymmetric,PlaneStess,PlaneStrain}() and
 * solids.IsotropicConductivity{K}.{Solid3D,Plane,Axisymmetric}() */
enum gofem_constituter_kind {
  GOFEM_ISO_SOLID3D = 0, GOFEM_ISO_AXISYM = 1, GOFEM_ISO_PLANESTRESS = 2, GOFEM_ISO_PLANESTRAIN = 3,
  GOFEM_COND_SOLID3D = 4, GOFEM_COND_PLANE = 5, GOFEM_COND_AXISYM = 6
};

typedef struct gofem_ga gofem_ga;

/* Element tables as the mirror evaluates them (for the table-parity tests): returns 0 or -1.
 * Outputs sized for the largest case: iso_nodes[20*3], N/dN evaluated at `v`, quadrature pos[216*3], w[216]. */
int gofem_elem_info(int kind, int quad_order, int node_dofs, int32_t *len_nodes, int32_t *dofs_flag, double *iso_nodes,
                    int32_t *nqp, double *qpos, double *qw, char *name, int name_cap);
int gofem_elem_basis(int kind, const double v[3], double *N, double *dN, int32_t *len_dN);
int gofem_constitutive(int ckind, double p0, double p1, double *C, int32_t *dimC, int32_t *bkind);

/* fem.NewGeneralAssembler(nodes, modelDofs) on CUDA device `device` */
int gofem_new_general_assembler(int device, const double *nodes_xyz, int64_t n_nodes, int model_dofs_flag, gofem_ga **out);
void gofem_free(gofem_ga *ga);
/* ga.AddIsoparametric(elemT, c, Nelem, getElement): getElement(i) is served from conn[i*nn..] (or from
 * `lens`-ragged rows when elem_lens != NULL, to exercise the node-count error), with xC = yC = orient. */
int gofem_add_isoparametric(gofem_ga *ga, int elem_kind, int quad_order, int node_dofs, int ckind, double p0, double p1,
                            const int64_t *conn, const int32_t *elem_lens, int64_t n_elem, const double orient_x[3]);
/* additive flat-connectivity fast path (no per-element callback) */
int gofem_add_isoparametric_flat(gofem_ga *ga, int elem_kind, int quad_order, int node_dofs, int ckind, double p0, double p1,
                                 const int64_t *conn, int64_t n_elem);
int gofem_isoparametric_strains(gofem_ga *ga, const double *u, int64_t n_u, int elem_kind, int quad_order, int node_dofs,
                                int ckind, double p0, double p1, const int64_t *conn, int64_t n_elem, double *strains);
/* Go-style error of the last call: "" == nil; is_panic != 0 where the reference panics */
const char *gofem_last_error(const gofem_ga *ga, int32_t *is_panic);
int64_t gofem_total_dofs(const gofem_ga *ga);
/* ga.Ksolid(): Dims / At / NonZero / CSR arrays / MulVec */
int64_t gofem_ksolid_nnz(gofem_ga *ga);
double gofem_ksolid_at(gofem_ga *ga, int64_t i, int64_t j);
int gofem_ksolid_csr(gofem_ga *ga, int64_t *indptr, int32_t *indices, double *values);
int gofem_ksolid_mulvec(gofem_ga *ga, const double *x, double *y);
/* the device handles underneath (for femb200_* calls: timers, slot map, device CSR) */
femb200_assembler *gofem_device_assembler(gofem_ga *ga);
femb200_ctx *gofem_device_ctx(gofem_ga *ga);
/* fem.Fixity: returns count; out may be NULL to query */
int64_t gofem_fixity_dofs(int model_dofs_flag, int64_t n_nodes, const uint16_t *fixed_flags, int get_fixed, int64_t *out);

#ifdef __cplusplus
}
#endif
#endif
