/*
 * femoracle_assemble.c -- ORACLE (test infrastructure, see femoracle.h).
 * Restates, operation by operation, the reference hot path of soypat/go-fem:
 *   fem.NewGeneralAssembler            assembler.go:21-28
 *   (*GeneralAssembler).AddIsoparametric  assembler_isoparametric.go:28-123
 *   forEachElement / mapdofs           fem.go:120-172
 *   storeElemNode / storeElemDofs / assembleElement   assembler.go:79-92,118-158
 *   solids.Isotropic / IsotropicConductivity / B fillers  constitution/solids/{isotropic,thermal,straindisplacement}.go
 * plus the third-party pieces that are NOT under /root/reference (restated from their published
 * algorithms, unverifiable here): gonum v0.11.0 mat.Det (LU + exp(sum log|u_ii|)*sign), Dense.Solve
 * (LU + getrs, Condition error if cond_1 > 1e16), Dense.Mul (inner index ascending, no FMA) and
 * soypat/lap v0.1.3 Sparse.Accumulate (duplicates summed) -- whose storage is replaced by the
 * canonical CSR of SURVEY.md 8(c): all emitted (row,col) pairs kept (explicit zeros too), rows and
 * columns ascending, duplicates summed in ascending element index, then call order.
 * Build with -ffp-contract=off.
 */
#include "femoracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define MAXNN 20
#define MAXD 3
#define MAXNDN 6
#define MAXQP 216
#define MAXDIMC 6
#define MAXND (MAXNN * MAXNDN)

/* ------------------------------------------------------------------ constitutive matrices */
/* isotropic.go:21-35,37-42 (Solid3D), :84-102 (Axisymmetric), :55-67 (PlaneStess), :69-82 (PlaneStrain);
 * thermal.go:12-58.  p0 = E or K, p1 = Poisson. */
int femo_constitutive(int mode, double p0, double p1, double *C, int *dimC, int *bkind) {
  double E = p0, nu = p1;
  switch (mode) {
  case FEMO_ISO_SOLID3D: {
    double G = E / (2 + 2 * nu);                               /* isotropic.go:45-47 */
    double lambda = E * nu / ((1 + nu) * (1 - 2 * nu));        /* isotropic.go:51-53 */
    if (isnan(lambda) || isinf(lambda)) return FEMO_ERR_CONSTITUTIVE;
    memset(C, 0, sizeof(double) * 36);
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) C[i * 6 + j] = (i == j) ? lambda + 2 * G : lambda;
    C[3 * 6 + 3] = G; C[4 * 6 + 4] = G; C[5 * 6 + 5] = G;
    *dimC = 6; *bkind = FEMO_B_XYZ;
    return 0;
  }
  case FEMO_ISO_AXISYM: {
    double f = nu / (1 - nu);
    double g = (1 - 2 * nu) / (2 * (1 - nu));
    double factor = E * (1 - nu) / ((1 + nu) * (1 - 2 * nu));
    f *= factor; g *= factor;
    double d[16] = {factor, f, f, 0, f, factor, f, 0, f, f, factor, 0, 0, 0, 0, g};
    memcpy(C, d, sizeof d);
    *dimC = 4; *bkind = FEMO_B_AXISYM;
    return 0;
  }
  case FEMO_ISO_PLANESTRESS: {
    double factor = E / (1 - nu * nu);
    double d[9] = {factor, factor * nu, 0, factor * nu, factor, 0, 0, 0, factor * (1 - nu) / 2};
    memcpy(C, d, sizeof d);
    *dimC = 3; *bkind = FEMO_B_PLANE;
    return 0;
  }
  case FEMO_ISO_PLANESTRAIN: {
    double factor = E / (1 + nu) / (1 - 2 * nu);
    double nuf = nu * factor;
    double d[9] = {factor * (1 - nu), nuf, nuf, nuf, factor * (1 - nu), nuf, nuf, nuf, factor * (1 - 2 * nu)};
    memcpy(C, d, sizeof d);
    *dimC = 3; *bkind = FEMO_B_PLANE;
    return 0;
  }
  case FEMO_COND_SOLID3D: {
    double d[9] = {p0, 0, 0, 0, p0, 0, 0, 0, p0};
    memcpy(C, d, sizeof d);
    *dimC = 3; *bkind = FEMO_B_THERMAL;
    return 0;
  }
  case FEMO_COND_PLANE: case FEMO_COND_AXISYM: {
    double d[4] = {p0, 0, 0, p0};
    memcpy(C, d, sizeof d);
    *dimC = 2; *bkind = (mode == FEMO_COND_PLANE) ? FEMO_B_THERMAL : FEMO_B_THERMAL_AXISYM;
    return 0;
  }
  }
  return FEMO_ERR_CONSTITUTIVE;
}

/* ------------------------------------------------------------------ dof mapping */
static int popcount16(int x) { int c = 0; for (int i = 0; i < 16; i++) c += (x >> i) & 1; return c; }

/* fem.go:155-172.  NOTE the loop bound is popcount(model), not 16: only flags contiguous from bit 0 map
 * correctly; unfilled entries stay 0 exactly as make([]int, n) leaves them. */
int femo_mapdofs(int model, int elem, int *dofmap, int cap) {
  int num_model = popcount16(model);
  if ((model & elem) != elem) return -FEMO_ERR_DOFS;
  int n = popcount16(elem);
  if (n > cap) return -FEMO_ERR_DOFS;
  for (int i = 0; i < n; i++) dofmap[i] = 0;
  int idm = 0;
  for (int i = 0; i < num_model; i++)
    if ((model & (1 << i)) && (elem & (1 << i))) dofmap[idm++] = i;
  if (idm == 0 || n == 0) return -FEMO_ERR_EMPTY_DOFMAP;
  return n;
}

/* ------------------------------------------------------------------ gonum-style small dense LA */
/* Dgetf2-style LU with partial pivoting on an n x n row-major matrix (n<=3): first max |.| pivot,
 * column scaled by the reciprocal of the pivot, rank-1 update.  Returns swaps count via piv. */
static void lu_factor(int n, double *a, int *piv) {
  for (int j = 0; j < n; j++) {
    int jp = j; double mx = fabs(a[j * n + j]);
    for (int i = j + 1; i < n; i++) if (fabs(a[i * n + j]) > mx) { mx = fabs(a[i * n + j]); jp = i; }
    piv[j] = jp;
    if (a[jp * n + j] != 0) {
      if (jp != j) for (int k = 0; k < n; k++) { double t = a[j * n + k]; a[j * n + k] = a[jp * n + k]; a[jp * n + k] = t; }
      double r = 1 / a[j * n + j];
      for (int i = j + 1; i < n; i++) a[i * n + j] *= r;
    }
    for (int i = j + 1; i < n; i++) {
      double l = a[i * n + j];
      for (int k = j + 1; k < n; k++) a[i * n + k] = a[i * n + k] - l * a[j * n + k];
    }
  }
}

/* mat.Det = exp(LogDet)*sign: assembler_isoparametric.go:94 */
static double det_gonum(int n, const double *jac) {
  double a[9]; int piv[3];
  memcpy(a, jac, sizeof(double) * n * n);
  lu_factor(n, a, piv);
  double logdet = 0, sign = 1;
  for (int i = 0; i < n; i++) {
    double v = a[i * n + i];
    if (v < 0) sign = -sign;
    if (piv[i] != i) sign = -sign;
    logdet += log(fabs(v));
  }
  return exp(logdet) * sign;
}

/* dNxy.Solve(jac, dN): assembler_isoparametric.go:100.  Returns cond_1 estimate (exact 1-norm
 * condition here; gonum uses Dgecon's estimate) so the caller can raise the Condition error. */
static double solve_gonum(int n, const double *jac, int nrhs, const double *b, double *x) {
  double a[9]; int piv[3];
  memcpy(a, jac, sizeof(double) * n * n);
  double anorm = 0;
  for (int j = 0; j < n; j++) { double s = 0; for (int i = 0; i < n; i++) s += fabs(jac[i * n + j]); if (s > anorm) anorm = s; }
  lu_factor(n, a, piv);
  memcpy(x, b, sizeof(double) * n * nrhs);
  for (int i = 0; i < n; i++) if (piv[i] != i)
    for (int k = 0; k < nrhs; k++) { double t = x[i * nrhs + k]; x[i * nrhs + k] = x[piv[i] * nrhs + k]; x[piv[i] * nrhs + k] = t; }
  for (int i = 1; i < n; i++)                 /* L (unit) forward */
    for (int k = 0; k < i; k++) { double l = a[i * n + k]; if (l != 0) for (int c = 0; c < nrhs; c++) x[i * nrhs + c] -= l * x[k * nrhs + c]; }
  for (int i = n - 1; i >= 0; i--) {          /* U backward */
    for (int k = i + 1; k < n; k++) { double u = a[i * n + k]; if (u != 0) for (int c = 0; c < nrhs; c++) x[i * nrhs + c] -= u * x[k * nrhs + c]; }
    double r = 1 / a[i * n + i];
    for (int c = 0; c < nrhs; c++) x[i * nrhs + c] *= r;
  }
  /* ||A^-1||_1 from the solved identity */
  double inv[9], e[9];
  for (int i = 0; i < n * n; i++) e[i] = 0;
  for (int i = 0; i < n; i++) e[i * n + i] = 1;
  memcpy(inv, e, sizeof(double) * n * n);
  for (int i = 0; i < n; i++) if (piv[i] != i)
    for (int k = 0; k < n; k++) { double t = inv[i * n + k]; inv[i * n + k] = inv[piv[i] * n + k]; inv[piv[i] * n + k] = t; }
  for (int i = 1; i < n; i++) for (int k = 0; k < i; k++) for (int c = 0; c < n; c++) inv[i * n + c] -= a[i * n + k] * inv[k * n + c];
  for (int i = n - 1; i >= 0; i--) {
    for (int k = i + 1; k < n; k++) for (int c = 0; c < n; c++) inv[i * n + c] -= a[i * n + k] * inv[k * n + c];
    for (int c = 0; c < n; c++) inv[i * n + c] /= a[i * n + i];
  }
  double inorm = 0;
  for (int j = 0; j < n; j++) { double s = 0; for (int i = 0; i < n; i++) s += fabs(inv[i * n + j]); if (s > inorm) inorm = s; }
  return anorm * inorm;
}

/* ------------------------------------------------------------------ per-call tables (prologue) */
typedef struct {
  int kind, d, nn, ndn, nd, nqp, dimC, bkind;
  double N[MAXQP][MAXNN];            /* Npg  : assembler_isoparametric.go:66-71 */
  double dN[MAXQP][MAXD * MAXNN];    /* dNpg : [d][nn] row-major */
  double w[MAXQP];
  double C[MAXDIMC * MAXDIMC];
  int dofmap[MAXNDN];
  int model_dofs_per_node;
} tables;

static int build_tables(tables *t, int model_flags, int kind, int quad_order, int node_dofs,
                        const double *C, int dimC, int bkind) {
  memset(t, 0, sizeof *t);
  int nn = femo_elem_len_nodes(kind);
  if (nn < 0) return FEMO_ERR_ELEMENT;
  double zero[3] = {0, 0, 0}, tmp[MAXD * MAXNN + 8];
  int lendn = femo_elem_basis_diff(kind, zero, tmp);
  int d = lendn / nn;                                    /* :40 integer division */
  if (d * nn != lendn) return FEMO_ERR_ELEMENT;          /* Triangle6: NewDense(1,6,<8>) panics at :70 */
  int eflags = femo_elem_dofs(kind, node_dofs);
  int ndn = popcount16(eflags);                          /* :44 */
  if (dimC < 1 || dimC > MAXDIMC || ndn > MAXNDN) return FEMO_ERR_CONSTITUTIVE;
  static double pos[MAXQP * 3];
  int nqp = femo_elem_quadrature(kind, quad_order, pos, t->w, MAXQP);   /* :56 */
  if (nqp < 0) return -nqp;
  if (nqp == 0) return FEMO_ERR_QUADRATURE;              /* :60-62 */
  for (int q = 0; q < nqp; q++) {
    femo_elem_basis(kind, &pos[3 * q], t->N[q]);
    femo_elem_basis_diff(kind, &pos[3 * q], t->dN[q]);
  }
  if (d < 1 || d > 3) return FEMO_ERR_DIMS;              /* fem.go:123-125 */
  int r = femo_mapdofs(model_flags, eflags, t->dofmap, MAXNDN);  /* fem.go:127 */
  if (r < 0) return -r;
  t->model_dofs_per_node = popcount16(model_flags);
  /* fem.go:138 allocates elemDofs with the MODEL dof count; storeElemDofs panics "bad length"
   * (assembler.go:148-150) unless it equals the element dof count. */
  if (t->model_dofs_per_node != ndn) return FEMO_ERR_BAD_LENGTH;
  t->kind = kind; t->d = d; t->nn = nn; t->ndn = ndn; t->nd = nn * ndn; t->nqp = nqp;
  t->dimC = dimC; t->bkind = bkind;
  memcpy(t->C, C, sizeof(double) * dimC * dimC);         /* Cd.Copy(C) :63 */
  return 0;
}

/* B fillers.  B is dimC x nd row-major, zero wherever the filler does not write (8-Q10). */
static double fill_B(const tables *t, double *B, const double *X /*[nn][d]*/, const double *dNxy /*[d][nn]*/,
                     const double *N) {
  int nn = t->nn, nd = t->nd;
  switch (t->bkind) {
  case FEMO_B_XYZ: /* straindisplacement.go:5-26 */
    for (int i = 0; i < nn; i++) {
      double dN0 = dNxy[i], dN1 = dNxy[nn + i], dN2 = dNxy[2 * nn + i];
      B[0 * nd + i * 3] = dN0; B[1 * nd + i * 3 + 1] = dN1; B[2 * nd + i * 3 + 2] = dN2;
      B[3 * nd + i * 3] = dN1; B[3 * nd + i * 3 + 1] = dN0;
      B[4 * nd + i * 3 + 1] = dN2; B[4 * nd + i * 3 + 2] = dN1;
      B[5 * nd + i * 3] = dN2; B[5 * nd + i * 3 + 2] = dN0;
    }
    return 1;
  case FEMO_B_PLANE: /* straindisplacement.go:28-40 */
    for (int i = 0; i < nn; i++) {
      double g0 = dNxy[i], g1 = dNxy[nn + i];
      B[0 * nd + i * 2] = g0; B[1 * nd + i * 2 + 1] = g1; B[2 * nd + i * 2] = g1; B[2 * nd + i * 2 + 1] = g0;
    }
    return 1;
  case FEMO_B_AXISYM: { /* straindisplacement.go:42-58 */
    double radius = 0;
    for (int i = 0; i < nn; i++) radius += X[i * t->d] * N[i];     /* mat.Dot(col 0, N) */
    double rInverse = 1 / radius;
    for (int i = 0; i < nn; i++) {
      double Ni = N[i], Ndr = dNxy[i], Ndz = dNxy[nn + i];
      B[0 * nd + i * 2] = Ndr; B[1 * nd + i * 2] = Ni * rInverse; B[2 * nd + i * 2 + 1] = Ndz;
      B[3 * nd + i * 2] = Ndz; B[3 * nd + i * 2 + 1] = Ndr;
    }
    return radius;
  }
  case FEMO_B_THERMAL: /* thermal.go:18-21,31-34: B.Copy(dN) */
    for (int r = 0; r < t->d; r++) for (int i = 0; i < nn; i++) B[r * nd + i] = dNxy[r * nn + i];
    return 1;
  case FEMO_B_THERMAL_AXISYM: { /* thermal.go:43-56 */
    double radius = 0;
    for (int i = 0; i < nn; i++) radius += X[i * t->d] * N[i];
    double rInverse = 1 / radius;
    for (int i = 0; i < nn; i++) {
      double Nir = N[i] * rInverse, Ndr = dNxy[i], Ndz = dNxy[nn + i];
      B[0 * nd + i] = Nir + Ndr; B[1 * nd + i] = Ndz;
    }
    return radius;
  }
  }
  return NAN;
}

/* geometry of one quadrature point: assembler_isoparametric.go:92-107.  Returns error kind. */
static int qp_geometry(const tables *t, int q, const double *X, double *B, double *dJac, double *scale) {
  int d = t->d, nn = t->nn;
  double jac[9], dNxy[MAXD * MAXNN];
  const double *dN = t->dN[q];
  for (int i = 0; i < d; i++)                      /* jac.Mul(dN, elemNod) :93 */
    for (int j = 0; j < d; j++) {
      double s = 0;
      for (int a = 0; a < nn; a++) s += dN[i * nn + a] * X[a * d + j];
      jac[i * d + j] = s;
    }
  *dJac = det_gonum(d, jac);                        /* :94 */
  if (*dJac < 0) return FEMO_ERR_NEG_DET;           /* :95-96 */
  else if (*dJac < 1e-12) return FEMO_ERR_ZERO_DET; /* :97-98 */
  double cond = solve_gonum(d, jac, nn, dN, dNxy);  /* :100 */
  if (cond > 1e16) return FEMO_ERR_SOLVE;           /* :101-103 (mat.ConditionTolerance) */
  *scale = fill_B(t, B, X, dNxy, t->N[q]);          /* :104 */
  if (isnan(*scale)) return FEMO_ERR_NAN_SCALE;     /* :105-107 */
  return 0;
}

/* test-only: |B| for the normaliser, from |J^-1| |dN| and |N| |1/r| */
static void abs_B(const tables *t, int q, const double *X, double *Ba) {
  int d = t->d, nn = t->nn, nd = t->nd;
  double jac[9], inv[9], eye[9], ga[MAXD * MAXNN];
  const double *dN = t->dN[q];
  for (int i = 0; i < d; i++)
    for (int j = 0; j < d; j++) {
      double s = 0;
      for (int a = 0; a < nn; a++) s += dN[i * nn + a] * X[a * d + j];
      jac[i * d + j] = s;
      eye[i * d + j] = (i == j);
    }
  solve_gonum(d, jac, d, eye, inv);
  for (int i = 0; i < d; i++)
    for (int b = 0; b < nn; b++) {
      double s = 0;
      for (int j = 0; j < d; j++) s += fabs(inv[i * d + j]) * fabs(dN[j * nn + b]);
      ga[i * nn + b] = s;
    }
  double radius = 0;
  for (int i = 0; i < nn; i++) radius += X[i * d] * t->N[q][i];
  double ri = fabs(1 / radius);
  memset(Ba, 0, sizeof(double) * t->dimC * nd);
  for (int i = 0; i < nn; i++) {
    double Na = fabs(t->N[q][i]);
    switch (t->bkind) {
    case FEMO_B_XYZ:
      Ba[0 * nd + i * 3] = ga[i]; Ba[1 * nd + i * 3 + 1] = ga[nn + i]; Ba[2 * nd + i * 3 + 2] = ga[2 * nn + i];
      Ba[3 * nd + i * 3] = ga[nn + i]; Ba[3 * nd + i * 3 + 1] = ga[i];
      Ba[4 * nd + i * 3 + 1] = ga[2 * nn + i]; Ba[4 * nd + i * 3 + 2] = ga[nn + i];
      Ba[5 * nd + i * 3] = ga[2 * nn + i]; Ba[5 * nd + i * 3 + 2] = ga[i];
      break;
    case FEMO_B_PLANE:
      Ba[0 * nd + i * 2] = ga[i]; Ba[1 * nd + i * 2 + 1] = ga[nn + i]; Ba[2 * nd + i * 2] = ga[nn + i]; Ba[2 * nd + i * 2 + 1] = ga[i];
      break;
    case FEMO_B_AXISYM:
      Ba[0 * nd + i * 2] = ga[i]; Ba[1 * nd + i * 2] = Na * ri; Ba[2 * nd + i * 2 + 1] = ga[nn + i];
      Ba[3 * nd + i * 2] = ga[nn + i]; Ba[3 * nd + i * 2 + 1] = ga[i];
      break;
    case FEMO_B_THERMAL:
      for (int r = 0; r < d; r++) Ba[r * nd + i] = ga[r * nn + i];
      break;
    case FEMO_B_THERMAL_AXISYM:
      Ba[0 * nd + i] = Na * ri + ga[i]; Ba[1 * nd + i] = ga[nn + i];
      break;
    }
  }
}

/* Ke of one element: assembler_isoparametric.go:89-113 */
static int element_ke(const tables *t, const double *X, double *B, double *aux1, double *aux2, double *Ke, int *err_qp,
                      double *KeAbs) {
  int nd = t->nd, dimC = t->dimC;
  memset(Ke, 0, sizeof(double) * nd * nd);                   /* Ke.Zero() :89 */
  if (KeAbs) memset(KeAbs, 0, sizeof(double) * nd * nd);
  for (int q = 0; q < t->nqp; q++) {
    double dJac, scale;
    int e = qp_geometry(t, q, X, B, &dJac, &scale);
    if (e) { *err_qp = q; return e; }
    /* aux1 = B^T C (nd x dimC): dgemm Trans/NoTrans, l outermost, zero multipliers skipped :109 */
    memset(aux1, 0, sizeof(double) * nd * dimC);
    for (int l = 0; l < dimC; l++)
      for (int i = 0; i < nd; i++) {
        double tmp = B[l * nd + i];
        if (tmp != 0) for (int j = 0; j < dimC; j++) aux1[i * dimC + j] += tmp * t->C[l * dimC + j];
      }
    /* aux2 = aux1 B (nd x nd): NoTrans/NoTrans, for i, for l, axpy :110 */
    memset(aux2, 0, sizeof(double) * nd * nd);
    for (int i = 0; i < nd; i++)
      for (int l = 0; l < dimC; l++) {
        double tmp = aux1[i * dimC + l];
        if (tmp != 0) for (int j = 0; j < nd; j++) aux2[i * nd + j] += tmp * B[l * nd + j];
      }
    double f = dJac * t->w[q] * scale;                        /* :111 left-to-right */
    for (int k = 0; k < nd * nd; k++) aux2[k] = f * aux2[k];
    for (int k = 0; k < nd * nd; k++) Ke[k] = Ke[k] + aux2[k]; /* :112 */
    if (KeAbs) {
      /* test-only normaliser (not in the reference): sum of |terms| of the same products,
       * KeAbs += |f| * |B|^T |C| |B| with |B| built from |J^-1| |dN| (so cancellation inside
       * dNxy = J^-1 dN is covered too) -- the scale against which rounding differences are graded */
      double Ba[MAXDIMC * MAXND];
      abs_B(t, q, X, Ba);
      for (int i = 0; i < nd; i++)
        for (int l = 0; l < dimC; l++) {
          double a1 = 0;
          for (int m = 0; m < dimC; m++) a1 += Ba[m * nd + i] * fabs(t->C[m * dimC + l]);
          if (a1 != 0) for (int j = 0; j < nd; j++) KeAbs[i * nd + j] += fabs(f) * a1 * Ba[l * nd + j];
        }
    }
  }
  return 0;
}

/* storeElemNode assembler.go:118-145 */
static void store_elem_node(double *dst, const double *nodes3, const int64_t *elem, int nn, int d) {
  for (int i = 0; i < nn; i++) for (int k = 0; k < d; k++) dst[i * d + k] = nodes3[3 * elem[i] + k];
}

/* ------------------------------------------------------------------ assembler + canonical CSR */
static int g_track_normaliser = 1;   /* tests: on; bench cpu_baseline: off (the reference computes no such thing) */
void femo_set_normaliser(int on) { g_track_normaliser = on; }
struct femo_assembler {
  const double *nodes; double *nodes_own;
  int64_t n_nodes; int model_flags; int64_t ndof;
  int64_t *indptr; int32_t *indices; double *values; double *absum; int64_t nnz;
  double last_seconds;
};

femo_assembler *femo_new_general_assembler(const double *nodes3, int64_t n_nodes, int model_flags) {
  femo_assembler *ga = (femo_assembler *)calloc(1, sizeof *ga);
  ga->nodes_own = (double *)malloc(sizeof(double) * 3 * (n_nodes > 0 ? n_nodes : 1));
  memcpy(ga->nodes_own, nodes3, sizeof(double) * 3 * n_nodes);
  ga->nodes = ga->nodes_own;
  ga->n_nodes = n_nodes; ga->model_flags = model_flags;
  ga->ndof = n_nodes * popcount16(model_flags);              /* assembler.go:22 */
  ga->indptr = (int64_t *)calloc(ga->ndof + 1, sizeof(int64_t));
  return ga;
}
void femo_free(femo_assembler *ga) {
  if (!ga) return;
  free(ga->nodes_own); free(ga->indptr); free(ga->indices); free(ga->values); free(ga->absum); free(ga);
}
int64_t femo_total_dofs(const femo_assembler *ga) { return ga->ndof; }
int64_t femo_nnz(const femo_assembler *ga) { return ga->nnz; }
double femo_last_seconds(const femo_assembler *ga) { return ga->last_seconds; }
void femo_csr(const femo_assembler *ga, int64_t *indptr, int32_t *indices, double *values) {
  memcpy(indptr, ga->indptr, sizeof(int64_t) * (ga->ndof + 1));
  if (ga->nnz) { memcpy(indices, ga->indices, sizeof(int32_t) * ga->nnz); memcpy(values, ga->values, sizeof(double) * ga->nnz); }
}
void femo_csr_absum(const femo_assembler *ga, double *absum) { if (ga->nnz) memcpy(absum, ga->absum, sizeof(double) * ga->nnz); }

typedef struct { int32_t col; int32_t seq_hi; double v; double va; } rowent; /* seq kept implicitly by stable order */

static int cmp_col_stable(const void *a, const void *b) {
  const rowent *x = (const rowent *)a, *y = (const rowent *)b;
  if (x->col != y->col) return x->col < y->col ? -1 : 1;
  return x->seq_hi < y->seq_hi ? -1 : (x->seq_hi > y->seq_hi);
}

/* lap.Sparse.Accumulate stand-in: sum the triplets (I,J,V), in emission order per (row,col), into the
 * canonical CSR, on top of the existing content (assembler_isoparametric.go:121). */
static void accumulate(femo_assembler *ga, const int64_t *I, const int64_t *J, const double *V, const double *Vabs, int64_t n) {
  int64_t ndof = ga->ndof;
  int64_t *cnt = (int64_t *)calloc(ndof + 1, sizeof(int64_t));
  for (int64_t k = 0; k < n; k++) cnt[I[k] + 1]++;
  for (int64_t r = 0; r < ndof; r++) cnt[r + 1] += cnt[r];
  int64_t *fill = (int64_t *)malloc(sizeof(int64_t) * (ndof + 1));
  memcpy(fill, cnt, sizeof(int64_t) * (ndof + 1));
  int32_t *bc = (int32_t *)malloc(sizeof(int32_t) * (n ? n : 1));
  double *bv = (double *)malloc(sizeof(double) * (n ? n : 1));
  double *ba = (double *)malloc(sizeof(double) * (n ? n : 1));
  for (int64_t k = 0; k < n; k++) { int64_t p = fill[I[k]]++; bc[p] = (int32_t)J[k]; bv[p] = V[k]; ba[p] = Vabs ? Vabs[k] : 0.0; } /* stable by row */
  free(fill);
  /* per row: stable sort by column, reduce duplicates in order, merge with the old row */
  int64_t maxrow = 0;
  for (int64_t r = 0; r < ndof; r++) if (cnt[r + 1] - cnt[r] > maxrow) maxrow = cnt[r + 1] - cnt[r];
  rowent *buf = (rowent *)malloc(sizeof(rowent) * (maxrow ? maxrow : 1));
  int64_t cap = ga->nnz + n;
  int64_t *nptr = (int64_t *)calloc(ndof + 1, sizeof(int64_t));
  int32_t *nidx = (int32_t *)malloc(sizeof(int32_t) * (cap ? cap : 1));
  double *nval = (double *)malloc(sizeof(double) * (cap ? cap : 1));
  double *nabs = (double *)malloc(sizeof(double) * (cap ? cap : 1));
  int64_t out = 0;
  for (int64_t r = 0; r < ndof; r++) {
    int64_t m = cnt[r + 1] - cnt[r];
    for (int64_t k = 0; k < m; k++) { buf[k].col = bc[cnt[r] + k]; buf[k].seq_hi = (int32_t)k; buf[k].v = bv[cnt[r] + k]; buf[k].va = ba[cnt[r] + k]; }
    if (m > 1) qsort(buf, (size_t)m, sizeof(rowent), cmp_col_stable);
    int64_t o = ga->indptr[r], oe = ga->indptr[r + 1], k = 0;
    nptr[r] = out;
    while (o < oe || k < m) {
      int32_t c;
      if (k >= m) c = ga->indices[o]; else if (o >= oe) c = buf[k].col; else c = ga->indices[o] < buf[k].col ? ga->indices[o] : buf[k].col;
      double s = 0, sa = 0; int have_new = 0;
      while (k < m && buf[k].col == c) { s = have_new ? s + buf[k].v : buf[k].v; sa += buf[k].va; have_new = 1; k++; }
      double val = s, ab = sa;
      if (o < oe && ga->indices[o] == c) { val = have_new ? ga->values[o] + s : ga->values[o]; ab = ga->absum[o] + sa; o++; }
      nidx[out] = c; nval[out] = val; nabs[out] = ab; out++;
    }
  }
  nptr[ndof] = out;
  free(buf); free(bc); free(bv); free(ba); free(cnt);
  free(ga->indptr); free(ga->indices); free(ga->values); free(ga->absum);
  ga->indptr = nptr; ga->indices = nidx; ga->values = nval; ga->absum = nabs; ga->nnz = out;
}

int femo_add_isoparametric(femo_assembler *ga, int kind, int quad_order, int node_dofs,
                           const double *C, int dimC, int bkind,
                           const int64_t *conn, int64_t nelem, int64_t *err_elem, int *err_qp) {
  struct timespec t0, t1;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  tables *t = (tables *)malloc(sizeof(tables));
  if (err_elem) *err_elem = -1;
  if (err_qp) *err_qp = -1;
  int rc = build_tables(t, ga->model_flags, kind, quad_order, node_dofs, C, dimC, bkind);
  if (rc) { free(t); return rc; }
  int nn = t->nn, nd = t->nd, ndn = t->ndn;
  int64_t nval = (int64_t)nd * nd;
  /* lap.NewSparseAccum(NvalPerElem*Nelem) :78-79 -- the full COO buffer, as the reference allocates it */
  int64_t total = nval * nelem;
  int64_t *I = (int64_t *)malloc(sizeof(int64_t) * (total ? total : 1));
  int64_t *J = (int64_t *)malloc(sizeof(int64_t) * (total ? total : 1));
  double *V = (double *)malloc(sizeof(double) * (total ? total : 1));
  const int track = g_track_normaliser;
  double *Vabs = (double *)malloc(sizeof(double) * (track && total ? total : 1));
  double *KeAbs = track ? (double *)malloc(sizeof(double) * nd * nd) : 0;
  double X[MAXNN * MAXD], B[MAXDIMC * MAXND], *aux1 = (double *)malloc(sizeof(double) * nd * MAXDIMC),
         *aux2 = (double *)malloc(sizeof(double) * nd * nd), *Ke = (double *)malloc(sizeof(double) * nd * nd);
  int64_t dofs[MAXND];
  memset(B, 0, sizeof B);                                       /* allocated once, never re-zeroed :52 */
  for (int64_t e = 0; e < nelem && !rc; e++) {                  /* fem.go:140-151 */
    const int64_t *elem = conn + e * nn;
    for (int i = 0; i < nn; i++) if (elem[i] < 0 || elem[i] >= ga->n_nodes) { rc = FEMO_ERR_NODE_OOB; if (err_elem) *err_elem = e; break; }
    if (rc) break;
    store_elem_node(X, ga->nodes, elem, nn, t->d);
    for (int a = 0; a < nn; a++)                                /* storeElemDofs assembler.go:147-158 */
      for (int k = 0; k < ndn; k++) dofs[a * ndn + k] = (int64_t)t->model_dofs_per_node * elem[a] + t->dofmap[k];
    int q = -1;
    rc = element_ke(t, X, B, aux1, aux2, Ke, &q, KeAbs);
    if (rc) { if (err_elem) *err_elem = e; if (err_qp) *err_qp = q; break; }
    int64_t off = e * nval;                                     /* assembleElement assembler.go:79-92 */
    for (int i = 0; i < nd; i++)
      for (int j = 0; j < nd; j++) { V[off + i * nd + j] = Ke[i * nd + j]; if (track) Vabs[off + i * nd + j] = KeAbs[i * nd + j]; I[off + i * nd + j] = dofs[i]; J[off + i * nd + j] = dofs[j]; }
  }
  if (!rc) accumulate(ga, I, J, V, track ? Vabs : 0, total);                      /* :118-121: K untouched on error */
  free(I); free(J); free(V); free(Vabs); free(KeAbs); free(aux1); free(aux2); free(Ke); free(t);
  clock_gettime(CLOCK_MONOTONIC, &t1);
  ga->last_seconds = (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
  return rc;
}

int femo_element_matrices(const double *nodes3, int64_t n_nodes, int kind, int quad_order, int node_dofs,
                          const double *C, int dimC, int bkind, const int64_t *conn, int64_t nelem,
                          double *KeOut, int64_t *err_elem, int *err_qp) {
  tables *t = (tables *)malloc(sizeof(tables));
  int eflags = femo_elem_dofs(kind, node_dofs);
  int rc = build_tables(t, eflags, kind, quad_order, node_dofs, C, dimC, bkind);
  if (rc) { free(t); return rc; }
  int nn = t->nn, nd = t->nd;
  double X[MAXNN * MAXD], B[MAXDIMC * MAXND];
  double *aux1 = (double *)malloc(sizeof(double) * nd * MAXDIMC), *aux2 = (double *)malloc(sizeof(double) * nd * nd);
  memset(B, 0, sizeof B);
  for (int64_t e = 0; e < nelem && !rc; e++) {
    const int64_t *elem = conn + e * nn;
    for (int i = 0; i < nn; i++) if (elem[i] < 0 || elem[i] >= n_nodes) rc = FEMO_ERR_NODE_OOB;
    if (rc) { if (err_elem) *err_elem = e; break; }
    store_elem_node(X, nodes3, elem, nn, t->d);
    int q = -1;
    rc = element_ke(t, X, B, aux1, aux2, KeOut + e * (int64_t)nd * nd, &q, 0);
    if (rc) { if (err_elem) *err_elem = e; if (err_qp) *err_qp = q; }
  }
  free(aux1); free(aux2); free(t);
  return rc;
}

/* assembler_isoparametric.go:126-213: strain_qp = B * u[elemDofs] */
int femo_isoparametric_strains(const femo_assembler *ga, const double *u, int kind, int quad_order, int node_dofs,
                               const double *C, int dimC, int bkind, const int64_t *conn, int64_t nelem,
                               double *strains, int64_t *err_elem, int *err_qp) {
  tables *t = (tables *)malloc(sizeof(tables));
  int rc = build_tables(t, ga->model_flags, kind, quad_order, node_dofs, C, dimC, bkind);
  if (rc) { free(t); return rc; }
  int nn = t->nn, nd = t->nd, ndn = t->ndn;
  double X[MAXNN * MAXD], B[MAXDIMC * MAXND];
  int64_t dofs[MAXND];
  memset(B, 0, sizeof B);
  for (int64_t e = 0; e < nelem && !rc; e++) {
    const int64_t *elem = conn + e * nn;
    for (int i = 0; i < nn; i++) if (elem[i] < 0 || elem[i] >= ga->n_nodes) rc = FEMO_ERR_NODE_OOB;
    if (rc) { if (err_elem) *err_elem = e; break; }
    store_elem_node(X, ga->nodes, elem, nn, t->d);
    for (int a = 0; a < nn; a++)
      for (int k = 0; k < ndn; k++) dofs[a * ndn + k] = (int64_t)t->model_dofs_per_node * elem[a] + t->dofmap[k];
    for (int q = 0; q < t->nqp; q++) {
      double dJac, scale;
      rc = qp_geometry(t, q, X, B, &dJac, &scale);
      if (rc) { if (err_elem) *err_elem = e; if (err_qp) *err_qp = q; break; }
      for (int r = 0; r < t->dimC; r++) {                       /* vd.MulVec(B, u_e) :206-207 */
        double s = 0;
        for (int j = 0; j < nd; j++) s += B[r * nd + j] * u[dofs[j]];
        strains[(e * t->nqp + q) * t->dimC + r] = s;
      }
    }
  }
  free(t);
  return rc;
}
