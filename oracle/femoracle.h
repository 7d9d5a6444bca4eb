/*
 * femoracle.h -- CPU ORACLE for the go-fem isoparametric stiffness-assembly path.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a plain-C, single-threaded restatement of the
 * reference algorithm (soypat/go-fem: assembler_isoparametric.go, assembler.go, fem.go,
 * elements/<type>.go, constitution/solids/<file>.go).  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may link or call it.  Nothing under
 * go-fem_b200/ (the product) includes, links or dlopens anything in oracle/.
 *
 * Parity status (see DESIGN.md "Oracle"):
 *   - pinned against the reference's own golden vectors: example_test.go:39-48 (Quad4 axisym),
 *     :72-85 (Tetra4 3D), :110-119 (Quad4 plane stress), :143-148 (Quad4 axisym thermal),
 *     the shape-function invariants of elements/elements{2,3}d_test.go and the two patch tests
 *     of constitution/solids/patch_test.go.
 *   - the Go reference cannot be compiled here (no Go toolchain, deps not vendored), so
 *     gonum v0.11.0 (mat.Det / Solve / Mul) and soypat/lap v0.1.3 (Sparse.Accumulate) are
 *     restated from their published algorithms: LU with partial pivoting, det = exp(sum log|u_ii|)*sign,
 *     inner-index-ascending products without FMA.  lap's storage order is "parity unpinned";
 *     the canonical CSR defined in SURVEY.md 8(c) is used instead.
 */
#ifndef FEMORACLE_H
#define FEMORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* element kinds (elements/<type>.go) */
enum {
  FEMO_TETRA4 = 0, FEMO_TETRA10 = 1, FEMO_HEXA8 = 2, FEMO_HEXA20 = 3,
  FEMO_QUAD4 = 4, FEMO_QUAD8 = 5, FEMO_TRIANGLE3 = 6, FEMO_TRIANGLE6 = 7
};
/* strain-displacement fillers (constitution/solids/straindisplacement.go, thermal.go) */
enum {
  FEMO_B_XYZ = 0, FEMO_B_PLANE = 1, FEMO_B_AXISYM = 2, FEMO_B_THERMAL = 3, FEMO_B_THERMAL_AXISYM = 4
};
/* solids.Isotropic factories (isotropic.go) and IsotropicConductivity (thermal.go) */
enum {
  FEMO_ISO_SOLID3D = 0, FEMO_ISO_AXISYM = 1, FEMO_ISO_PLANESTRESS = 2, FEMO_ISO_PLANESTRAIN = 3,
  FEMO_COND_SOLID3D = 4, FEMO_COND_PLANE = 5, FEMO_COND_AXISYM = 6
};
/* error kinds of AddIsoparametric (assembler_isoparametric.go:95-107, fem.go:142-159) */
enum {
  FEMO_OK = 0,
  FEMO_ERR_NEG_DET = 1,     /* :96  "negative determinant of jacobian of element #%d, Check node ordering" */
  FEMO_ERR_ZERO_DET = 2,    /* :98  "zero determinant of jacobian of element #%d, ..." */
  FEMO_ERR_SOLVE = 3,       /* :102 "error calculating element #%d form factor: ..." (cond > 1e16) */
  FEMO_ERR_NAN_SCALE = 4,   /* :106 "NaN scale value returned ... at element #%d, quad %d" */
  FEMO_ERR_DOFS = 5,        /* fem.go:158 model dofs does not contain all element dofs */
  FEMO_ERR_EMPTY_DOFMAP = 6,/* fem.go:169 */
  FEMO_ERR_BAD_LENGTH = 7,  /* assembler.go:149 panic("bad length"): elem dofs/node != model dofs/node */
  FEMO_ERR_NODE_OOB = 8,    /* Go slice bounds panic at assembler.go:137 */
  FEMO_ERR_CONSTITUTIVE = 9,/* isotropic.go:24-26 NaN/Inf lambda */
  FEMO_ERR_ELEMENT = 10,    /* Triangle6: BasisDiff length 8 for 6 nodes -> mat.NewDense panic (:70) */
  FEMO_ERR_QUADRATURE = 11, /* quadrature.go:108 degree not implemented / triangles.go:160-165 */
  FEMO_ERR_DIMS = 12        /* fem.go:124 dimsPerNode must be 1, 2 or 3 */
};

int femo_elem_len_nodes(int kind);
int femo_elem_default_dofs(int kind);               /* DofsFlag bitset */
int femo_elem_dofs(int kind, int node_dofs);        /* Dofs() honouring NodeDofs (Tetra4/Triangle6 ignore it) */
int femo_elem_iso_nodes(int kind, double *out3);    /* [nn][3] */
int femo_elem_basis(int kind, const double v[3], double *N);              /* returns len */
int femo_elem_basis_diff(int kind, const double v[3], double *dN);        /* returns len ([d][nn] row major) */
int femo_elem_quadrature(int kind, int order, double *pos3, double *w, int cap); /* returns nqp or -err */
double femo_elem_measure(int kind);                 /* volume()/area() of the tests */

int femo_constitutive(int mode, double p0, double p1, double *C, int *dimC, int *bkind);
int femo_mapdofs(int model_flags, int elem_flags, int *dofmap, int cap);  /* returns ndn or -err */

/* ---- sparse K in canonical CSR ---- */
typedef struct femo_assembler femo_assembler;
femo_assembler *femo_new_general_assembler(const double *nodes3, int64_t n_nodes, int model_flags);
void femo_free(femo_assembler *ga);
int64_t femo_total_dofs(const femo_assembler *ga);
/* AddIsoparametric. conn is [nelem][nn] int64 (Go int). err_elem/err_qp receive the failing element/qp. */
int femo_add_isoparametric(femo_assembler *ga, int kind, int quad_order, int node_dofs,
                           const double *C, int dimC, int bkind,
                           const int64_t *conn, int64_t nelem, int64_t *err_elem, int *err_qp);
int64_t femo_nnz(const femo_assembler *ga);
void femo_csr(const femo_assembler *ga, int64_t *indptr, int32_t *indices, double *values);
/* per-entry normaliser s_ij, same layout as values: the sum over elements and quadrature
 * points of |dJac*w*scale| * (|B|^T |C| |B|)_ij, i.e. the sum of the absolute values of every product that
 * enters K_ij.  Rounding-level differences between two correct implementations are bounded by a small
 * multiple of eps * s_ij; entries that cancel (inside Ke or across elements) have |K_ij| << s_ij. */
void femo_csr_absum(const femo_assembler *ga, double *absum);
void femo_set_normaliser(int on); /* 0: skip the test-only normaliser (timed cpu_baseline runs) */

/* element stiffness matrices only: Ke[nelem][nd][nd] (assembler_isoparametric.go:89-113) */
int femo_element_matrices(const double *nodes3, int64_t n_nodes, int kind, int quad_order, int node_dofs,
                          const double *C, int dimC, int bkind, const int64_t *conn, int64_t nelem,
                          double *Ke, int64_t *err_elem, int *err_qp);
/* IsoparametricStrains (assembler_isoparametric.go:126-213): strains[nelem][nqp][dimC] */
int femo_isoparametric_strains(const femo_assembler *ga, const double *u, int kind, int quad_order, int node_dofs,
                               const double *C, int dimC, int bkind, const int64_t *conn, int64_t nelem,
                               double *strains, int64_t *err_elem, int *err_qp);
/* timing hook for bench.py: seconds spent in the last femo_add_isoparametric (whole call) */
double femo_last_seconds(const femo_assembler *ga);

#ifdef __cplusplus
}
#endif
#endif
