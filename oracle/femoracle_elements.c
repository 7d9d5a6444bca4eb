/*
 * femoracle_elements.c -- ORACLE (test infrastructure, see femoracle.h).
 * Restates the reference-element tables of soypat/go-fem `elements/<type>.go`:
 * node positions, shape functions, their derivatives and the quadrature rules.
 * Each function cites the reference lines it follows.  Floating expressions keep the
 * reference's operand order; Go's exact untyped-constant arithmetic is reproduced with
 * __float128 constant folding (Q suffix) followed by one rounding to double.
 * Build with -ffp-contract=off (Go on amd64 never fuses a*b+c).
 */
#include "femoracle.h"
#include <string.h>

#define X v[0]
#define Y v[1]
#define Z v[2]

int femo_elem_len_nodes(int kind) {
  switch (kind) {
  case FEMO_TETRA4: return 4;      /* tetra4.go:22 */
  case FEMO_TETRA10: return 10;    /* tetra10.go:26 */
  case FEMO_HEXA8: return 8;       /* hexa8.go:22 */
  case FEMO_HEXA20: return 20;     /* hexa20.go:22 */
  case FEMO_QUAD4: return 4;       /* quad4.go:32 */
  case FEMO_QUAD8: return 8;       /* quad8.go:33 */
  case FEMO_TRIANGLE3: return 3;   /* triangles.go:32 */
  case FEMO_TRIANGLE6: return 6;   /* triangles.go:98 */
  }
  return -1;
}

int femo_elem_default_dofs(int kind) {
  switch (kind) {
  case FEMO_TETRA4: case FEMO_TETRA10: case FEMO_HEXA8: case FEMO_HEXA20:
    return 0x7;                    /* fem.DofPos */
  case FEMO_QUAD4: case FEMO_QUAD8: case FEMO_TRIANGLE3: case FEMO_TRIANGLE6:
    return 0x3;                    /* DofPosX|DofPosY */
  }
  return 0;
}

/* Dofs(): NodeDofs override honoured by every type except Tetra4 (tetra4.go:18-20) and
 * Triangle6 (triangles.go:93-95, no such field). */
int femo_elem_dofs(int kind, int node_dofs) {
  if (kind == FEMO_TETRA4 || kind == FEMO_TRIANGLE6 || node_dofs == 0) return femo_elem_default_dofs(kind);
  return node_dofs;
}

double femo_elem_measure(int kind) {
  switch (kind) {
  case FEMO_TETRA4: case FEMO_TETRA10: return 1;   /* tetra4.go:55, tetra10.go:117 */
  case FEMO_HEXA8: case FEMO_HEXA20: return 8;     /* hexa8.go:113, hexa20.go:181 */
  case FEMO_QUAD4: case FEMO_QUAD8: return 4;      /* quad4.go:93, quad8.go:109 */
  case FEMO_TRIANGLE3: case FEMO_TRIANGLE6: return 0.5; /* triangles.go:80,156 */
  }
  return 0;
}

int femo_elem_iso_nodes(int kind, double *o) {
  static const double t10[10][3] = { /* tetra10.go:29-42 (first four = tetra4.go:25-32) */
    {0,0,0},{1,0,0},{0,1,0},{0,0,1},{0.5,0,0},{0,0.5,0},{0,0,0.5},{0.5,0.5,0},{0,0.5,0.5},{0.5,0,0.5}};
  static const double h8[8][3] = {   /* hexa8.go:32-43 */
    {-1,-1,-1},{1,-1,-1},{1,1,-1},{-1,1,-1},{-1,-1,1},{1,-1,1},{1,1,1},{-1,1,1}};
  static const double h20[20][3] = { /* hexa20.go:33-56 */
    {-1,-1,-1},{-1,1,-1},{1,1,-1},{1,-1,-1},{-1,-1,1},{-1,1,1},{1,1,1},{1,-1,1},
    {-1,0,-1},{0,1,-1},{1,0,-1},{0,-1,-1},{-1,0,1},{0,1,1},{1,0,1},{0,-1,1},
    {-1,-1,0},{-1,1,0},{1,1,0},{1,-1,0}};
  static const double q8[8][3] = {   /* quad8.go:36-47 (first four = quad4.go:35-42) */
    {-1,-1,0},{1,-1,0},{1,1,0},{-1,1,0},{0,-1,0},{1,0,0},{0,1,0},{-1,0,0}};
  static const double t6[6][3] = {   /* triangles.go:101-111 (first three = :35-42) */
    {0,0,0},{1,0,0},{0,1,0},{0.5,0.5,0},{0,0.5,0},{0.5,0,0}};
  int nn = femo_elem_len_nodes(kind);
  const double *src = 0;
  switch (kind) {
  case FEMO_TETRA4: case FEMO_TETRA10: src = &t10[0][0]; break;
  case FEMO_HEXA8: src = &h8[0][0]; break;
  case FEMO_HEXA20: src = &h20[0][0]; break;
  case FEMO_QUAD4: case FEMO_QUAD8: src = &q8[0][0]; break;
  case FEMO_TRIANGLE3: case FEMO_TRIANGLE6: src = &t6[0][0]; break;
  default: return -1;
  }
  memcpy(o, src, sizeof(double) * 3 * nn);
  return nn;
}

int femo_elem_basis(int kind, const double v[3], double *N) {
  switch (kind) {
  case FEMO_TETRA4: /* tetra4.go:35-38 */
    N[0] = 1 - X - Y - Z; N[1] = X; N[2] = Y; N[3] = Z;
    return 4;
  case FEMO_TETRA10: /* tetra10.go:45-58 */
    N[0] = 2*X*X + 4*X*Y + 4*X*Z - 3*X + 2*Y*Y + 4*Y*Z - 3*Y + 2*Z*Z - 3*Z + 1;
    N[1] = 2*X*X - X;
    N[2] = 2*Y*Y - Y;
    N[3] = 2*Z*Z - Z;
    N[4] = 4*X - 4*X*Y - 4*X*Z - 4*X*X;
    N[5] = 4*Y - 4*X*Y - 4*Y*Z - 4*Y*Y;
    N[6] = 4*Z - 4*X*Z - 4*Y*Z - 4*Z*Z;
    N[7] = 4 * X * Y;
    N[8] = 4 * Y * Z;
    N[9] = 4 * X * Z;
    return 10;
  case FEMO_HEXA8: /* hexa8.go:46-57 */
    N[0] = (Z*Y - Y - X - Z + Z*X + Y*X - Z*Y*X + 1) / 8;
    N[1] = (X - Y - Z + Z*Y - Z*X - Y*X + Z*Y*X + 1) / 8;
    N[2] = (Y - Z + X - Z*Y - Z*X + Y*X - Z*Y*X + 1) / 8;
    N[3] = (Y - Z - X - Z*Y + Z*X - Y*X + Z*Y*X + 1) / 8;
    N[4] = (Z - Y - X - Z*Y - Z*X + Y*X + Z*Y*X + 1) / 8;
    N[5] = (Z - Y + X - Z*Y + Z*X - Y*X - Z*Y*X + 1) / 8;
    N[6] = (Z + Y + X + Z*Y + Z*X + Y*X + Z*Y*X + 1) / 8;
    N[7] = (Z + Y - X + Z*Y - Z*X - Y*X - Z*Y*X + 1) / 8;
    return 8;
  case FEMO_HEXA20: { /* hexa20.go:59-86 */
    double x2 = X * X, y2 = Y * Y, z2 = Z * Z;
    N[0] = ((Y - 1) * (X - 1) * (Z - 1) * (Y + X + Z + 2)) / 8;
    N[1] = -((Y + 1) * (X - 1) * (Z - 1) * (X - Y + Z + 2)) / 8;
    N[2] = -((Y + 1) * (X + 1) * (Z - 1) * (Y + X - Z - 2)) / 8;
    N[3] = -((Y - 1) * (X + 1) * (Z - 1) * (Y - X + Z + 2)) / 8;
    N[4] = -((Y - 1) * (X - 1) * (Z + 1) * (Y + X - Z + 2)) / 8;
    N[5] = -((Y + 1) * (X - 1) * (Z + 1) * (Y - X + Z - 2)) / 8;
    N[6] = ((Y + 1) * (X + 1) * (Z + 1) * (Y + X + Z - 2)) / 8;
    N[7] = ((Y - 1) * (X + 1) * (Z + 1) * (Y - X - Z + 2)) / 8;
    N[8] = -((y2 - 1) * (X - 1) * (Z - 1)) / 4;
    N[9] = ((x2 - 1) * (Y + 1) * (Z - 1)) / 4;
    N[10] = ((y2 - 1) * (X + 1) * (Z - 1)) / 4;
    N[11] = -((x2 - 1) * (Y - 1) * (Z - 1)) / 4;
    N[12] = ((y2 - 1) * (X - 1) * (Z + 1)) / 4;
    N[13] = -((x2 - 1) * (Y + 1) * (Z + 1)) / 4;
    N[14] = -((y2 - 1) * (X + 1) * (Z + 1)) / 4;
    N[15] = ((x2 - 1) * (Y - 1) * (Z + 1)) / 4;
    N[16] = -((z2 - 1) * (Y - 1) * (X - 1)) / 4;
    N[17] = ((z2 - 1) * (Y + 1) * (X - 1)) / 4;
    N[18] = -((z2 - 1) * (Y + 1) * (X + 1)) / 4;
    N[19] = ((z2 - 1) * (Y - 1) * (X + 1)) / 4;
    return 20;
  }
  case FEMO_QUAD4: { /* quad4.go:46-54 */
    double x = X, y = Y;
    N[0] = 0.25 * (1 - x) * (1 - y);
    N[1] = 0.25 * (1 + x) * (1 - y);
    N[2] = 0.25 * (1 + x) * (1 + y);
    N[3] = 0.25 * (1 - x) * (1 + y);
    return 4;
  }
  case FEMO_QUAD8: { /* quad8.go:51-63 */
    double x = X, y = Y;
    N[0] = 0.25 * (1 - x) * (1 - y) * (-1 - x - y);
    N[1] = 0.25 * (1 + x) * (1 - y) * (-1 + x - y);
    N[2] = 0.25 * (1 + x) * (1 + y) * (-1 + x + y);
    N[3] = 0.25 * (1 - x) * (1 + y) * (-1 - x + y);
    N[4] = 0.5 * (1 - x) * (1 + x) * (1 - y);
    N[5] = 0.5 * (1 + x) * (1 + y) * (1 - y);
    N[6] = 0.5 * (1 - x) * (1 + x) * (1 + y);
    N[7] = 0.5 * (1 - x) * (1 - y) * (1 + y);
    return 8;
  }
  case FEMO_TRIANGLE3: { /* triangles.go:46-53 */
    double x = X, y = Y;
    N[0] = 1 - x - y; N[1] = x; N[2] = y;
    return 3;
  }
  case FEMO_TRIANGLE6: { /* triangles.go:114-124 (fails the Kronecker test in the reference too) */
    double x = X, y = Y;
    N[0] = 1 - 3*x - 3*y + 2*x*x + 4*x*y + 2*y*y;
    N[1] = 4*x - 4*x*x - 4*x*y;
    N[2] = -x + 2*x*x;
    N[3] = 4 * x * y;
    N[4] = -y + 2*y*y;
    N[5] = 4*y - 4*x*y - 4*y*y;
    return 6;
  }
  }
  return -1;
}

int femo_elem_basis_diff(int kind, const double v[3], double *d) {
  switch (kind) {
  case FEMO_TETRA4: { /* tetra4.go:40-46 */
    static const double c[12] = {-1, 1, 0, 0, -1, 0, 1, 0, -1, 0, 0, 1};
    memcpy(d, c, sizeof c);
    return 12;
  }
  case FEMO_TETRA10: /* tetra10.go:60-96 */
    d[0] = 4*X + 4*Y + 4*Z - 3;
    d[1] = 4*X - 1;
    d[2] = 0; d[3] = 0;
    d[4] = 4 - 4*Y - 4*Z - 8*X;
    d[5] = -4 * Y;
    d[6] = -4 * Z;
    d[7] = 4 * Y;
    d[8] = 0;
    d[9] = 4 * Z;
    d[10] = 4*X + 4*Y + 4*Z - 3;
    d[11] = 0;
    d[12] = 4*Y - 1;
    d[13] = 0;
    d[14] = -4 * X;
    d[15] = 4 - 8*Y - 4*Z - 4*X;
    d[16] = -4 * Z;
    d[17] = 4 * X;
    d[18] = 4 * Z;
    d[19] = 0;
    d[20] = 4*X + 4*Y + 4*Z - 3;
    d[21] = 0; d[22] = 0;
    d[23] = 4*Z - 1;
    d[24] = -4 * X;
    d[25] = -4 * Y;
    d[26] = 4 - 4*Y - 8*Z - 4*X;
    d[27] = 0;
    d[28] = 4 * Y;
    d[29] = 4 * X;
    return 30;
  case FEMO_HEXA8: /* hexa8.go:67-96 */
    d[0] = (Z + Y - (Z * Y) - 1) / 8;
    d[1] = ((Z * Y) - Y - Z + 1) / 8;
    d[2] = (Y - Z - (Z * Y) + 1) / 8;
    d[3] = (Z - Y + (Z * Y) - 1) / 8;
    d[4] = (Y - Z + (Z * Y) - 1) / 8;
    d[5] = (Z - Y - (Z * Y) + 1) / 8;
    d[6] = (Z + Y + (Z * Y) + 1) / 8;
    d[7] = (-Z - Y - (Z * Y) - 1) / 8;
    d[8] = (Z + X - (Z * X) - 1) / 8;
    d[9] = (Z - X + (Z * X) - 1) / 8;
    d[10] = (X - Z - (Z * X) + 1) / 8;
    d[11] = ((Z * X) - X - Z + 1) / 8;
    d[12] = (X - Z + (Z * X) - 1) / 8;
    d[13] = (-Z - X - (Z * X) - 1) / 8;
    d[14] = (Z + X + (Z * X) + 1) / 8;
    d[15] = (Z - X - (Z * X) + 1) / 8;
    d[16] = (Y + X - (Y * X) - 1) / 8;
    d[17] = (Y - X + (Y * X) - 1) / 8;
    d[18] = (-Y - X - (Y * X) - 1) / 8;
    d[19] = (X - Y + (Y * X) - 1) / 8;
    d[20] = ((Y * X) - X - Y + 1) / 8;
    d[21] = (X - Y - (Y * X) + 1) / 8;
    d[22] = (Y + X + (Y * X) + 1) / 8;
    d[23] = (Y - X - (Y * X) + 1) / 8;
    return 24;
  case FEMO_HEXA20: { /* hexa20.go:96-163 */
    double x2 = X * X, y2 = Y * Y, z2 = Z * Z;
    d[0] = ((Y - 1) * (Z - 1) * (Y + 2*X + Z + 1)) / 8;
    d[1] = -((Y + 1) * (Z - 1) * (2*X - Y + Z + 1)) / 8;
    d[2] = -((Y + 1) * (Z - 1) * (Y + 2*X - Z - 1)) / 8;
    d[3] = -((Y - 1) * (Z - 1) * (Y - 2*X + Z + 1)) / 8;
    d[4] = -((Y - 1) * (Z + 1) * (Y + 2*X - Z + 1)) / 8;
    d[5] = -((Y + 1) * (Z + 1) * (Y - 2*X + Z - 1)) / 8;
    d[6] = ((Y + 1) * (Z + 1) * (Y + 2*X + Z - 1)) / 8;
    d[7] = ((Y - 1) * (Z + 1) * (Y - 2*X - Z + 1)) / 8;
    d[8] = -((y2 - 1) * (Z - 1)) / 4;
    d[9] = (X * (Y + 1) * (Z - 1)) / 2;
    d[10] = ((y2 - 1) * (Z - 1)) / 4;
    d[11] = -(X * (Y - 1) * (Z - 1)) / 2;
    d[12] = ((y2 - 1) * (Z + 1)) / 4;
    d[13] = -(X * (Y + 1) * (Z + 1)) / 2;
    d[14] = -((y2 - 1) * (Z + 1)) / 4;
    d[15] = (X * (Y - 1) * (Z + 1)) / 2;
    d[16] = -((z2 - 1) * (Y - 1)) / 4;
    d[17] = ((z2 - 1) * (Y + 1)) / 4;
    d[18] = -((z2 - 1) * (Y + 1)) / 4;
    d[19] = ((z2 - 1) * (Y - 1)) / 4;
    d[20] = ((X - 1) * (Z - 1) * (2*Y + X + Z + 1)) / 8;
    d[21] = -((X - 1) * (Z - 1) * (X - 2*Y + Z + 1)) / 8;
    d[22] = -((X + 1) * (Z - 1) * (2*Y + X - Z - 1)) / 8;
    d[23] = -((X + 1) * (Z - 1) * (2*Y - X + Z + 1)) / 8;
    d[24] = -((X - 1) * (Z + 1) * (2*Y + X - Z + 1)) / 8;
    d[25] = -((X - 1) * (Z + 1) * (2*Y - X + Z - 1)) / 8;
    d[26] = ((X + 1) * (Z + 1) * (2*Y + X + Z - 1)) / 8;
    d[27] = ((X + 1) * (Z + 1) * (2*Y - X - Z + 1)) / 8;
    d[28] = -(Y * (X - 1) * (Z - 1)) / 2;
    d[29] = ((x2 - 1) * (Z - 1)) / 4;
    d[30] = (Y * (X + 1) * (Z - 1)) / 2;
    d[31] = -((x2 - 1) * (Z - 1)) / 4;
    d[32] = (Y * (X - 1) * (Z + 1)) / 2;
    d[33] = -((x2 - 1) * (Z + 1)) / 4;
    d[34] = -(Y * (X + 1) * (Z + 1)) / 2;
    d[35] = ((x2 - 1) * (Z + 1)) / 4;
    d[36] = -((z2 - 1) * (X - 1)) / 4;
    d[37] = ((z2 - 1) * (X - 1)) / 4;
    d[38] = -((z2 - 1) * (X + 1)) / 4;
    d[39] = ((z2 - 1) * (X + 1)) / 4;
    d[40] = ((Y - 1) * (X - 1) * (Y + X + 2*Z + 1)) / 8;
    d[41] = -((Y + 1) * (X - 1) * (X - Y + 2*Z + 1)) / 8;
    d[42] = -((Y + 1) * (X + 1) * (Y + X - 2*Z - 1)) / 8;
    d[43] = -((Y - 1) * (X + 1) * (Y - X + 2*Z + 1)) / 8;
    d[44] = -((Y - 1) * (X - 1) * (Y + X - 2*Z + 1)) / 8;
    d[45] = -((Y + 1) * (X - 1) * (Y - X + 2*Z - 1)) / 8;
    d[46] = ((Y + 1) * (X + 1) * (Y + X + 2*Z - 1)) / 8;
    d[47] = ((Y - 1) * (X + 1) * (Y - X - 2*Z + 1)) / 8;
    d[48] = -((y2 - 1) * (X - 1)) / 4;
    d[49] = ((x2 - 1) * (Y + 1)) / 4;
    d[50] = ((y2 - 1) * (X + 1)) / 4;
    d[51] = -((x2 - 1) * (Y - 1)) / 4;
    d[52] = ((y2 - 1) * (X - 1)) / 4;
    d[53] = -((x2 - 1) * (Y + 1)) / 4;
    d[54] = -((y2 - 1) * (X + 1)) / 4;
    d[55] = ((x2 - 1) * (Y - 1)) / 4;
    d[56] = -(Z * (Y - 1) * (X - 1)) / 2;
    d[57] = (Z * (Y + 1) * (X - 1)) / 2;
    d[58] = -(Z * (Y + 1) * (X + 1)) / 2;
    d[59] = (Z * (Y - 1) * (X + 1)) / 2;
    return 60;
  }
  case FEMO_QUAD4: { /* quad4.go:57-71 */
    double x = X, y = Y;
    d[0] = y/4 - 1./4;
    d[1] = 1./4 - y/4;
    d[2] = y/4 + 1./4;
    d[3] = -y/4 - 1./4;
    d[4] = x/4 - 1./4;
    d[5] = -x/4 - 1./4;
    d[6] = x/4 + 1./4;
    d[7] = 1./4 - x/4;
    return 8;
  }
  case FEMO_QUAD8: { /* quad8.go:66-88 */
    double x = X, y = Y;
    d[0] = -(x/4-1./4)*(y-1) - ((y-1)*(x+y+1))/4;
    d[1] = ((y-1)*(y-x+1))/4 - (x/4+1./4)*(y-1);
    d[2] = (x/4+1./4)*(y+1) + ((y+1)*(x+y-1))/4;
    d[3] = (x/4-1./4)*(y+1) + ((y+1)*(x-y+1))/4;
    d[4] = ((x+1)*(y-1))/2 + (x/2-1./2)*(y-1);
    d[5] = -((y - 1) * (y + 1)) / 2;
    d[6] = -((x+1)*(y+1))/2 - (x/2-1./2)*(y+1);
    d[7] = ((y - 1) * (y + 1)) / 2;
    d[8] = -(x/4-1./4)*(y-1) - (x/4-1./4)*(x+y+1);
    d[9] = (x/4+1./4)*(y-x+1) + (x/4+1./4)*(y-1);
    d[10] = (x/4+1./4)*(y+1) + (x/4+1./4)*(x+y-1);
    d[11] = (x/4-1./4)*(x-y+1) - (x/4-1./4)*(y+1);
    d[12] = (x/2 - 1./2) * (x + 1);
    d[13] = -(x/2+1./2)*(y-1) - (x/2+1./2)*(y+1);
    d[14] = -(x/2 - 1./2) * (x + 1);
    d[15] = (x/2-1./2)*(y-1) + (x/2-1./2)*(y+1);
    return 16;
  }
  case FEMO_TRIANGLE3: { /* triangles.go:56-65 */
    static const double c[6] = {-1, 1, 0, -1, 0, 1};
    memcpy(d, c, sizeof c);
    return 6;
  }
  case FEMO_TRIANGLE6: { /* triangles.go:127-139: EIGHT values for six nodes (reference bug, 8-Q5) */
    double x = X, y = Y;
    d[0] = -3 + 4*x + 4*y;
    d[1] = 4 - 8*x - 4*y;
    d[2] = -1 + 4*x;
    d[3] = 4 * y;
    d[4] = -4 * y;
    d[5] = -4 + 4*x + 8*y;
    d[6] = 4 * x;
    d[7] = -4 * x;
    return 8;
  }
  }
  return -1;
}

/* quadrature.go:50-111 -- 1-D Gauss-Legendre, n = 1..6.  n=4 keeps the reference's a==b (8-Q4). */
static int gauss1d(int n, double *x, double *w) {
#define SQRT3 1.7320508075688772935274463415058723669428052538103806280558069794Q
#define SQRT5 2.2360679774997896964091736687312762354406183596115257242708972454Q
  switch (n) {
  case 1: x[0] = 0; w[0] = 2; return 0;
  case 2: { const double a = (double)(SQRT3 / 3);
    x[0] = -a; x[1] = a; w[0] = 1; w[1] = 1; return 0; }
  case 3: { const double a = (double)(SQRT3 / SQRT5);
    x[0] = -a; x[1] = 0; x[2] = a;
    w[0] = (double)(5.0Q / 9.0Q); w[1] = (double)(8.0Q / 9.0Q); w[2] = (double)(5.0Q / 9.0Q); return 0; }
  case 4: {
#define SQRT30 5.4772255750516611345696978280080213395274469499798325422689444973Q
    const double a = (double)0.5216121828372476321853091190703617791381046921172457657794949152Q;
    const double b = (double)0.5216121828372476321853091190703617791381046921172457657794949152Q;
    const double wa = (double)((18 + SQRT30) / 36), wb = (double)((18 - SQRT30) / 36);
    x[0] = -b; x[1] = -a; x[2] = a; x[3] = b;
    w[0] = wb; w[1] = wa; w[2] = wa; w[3] = wb; return 0; }
  case 5: {
#define SQRT70 8.3666002653407554797817202578518748939281536929867219981119154308Q
    const double a = (double)0.538469310105683091036314420700208804967286606905559956202231627059471Q;
    const double b = (double)0.9061798459386639927976268782993929651256519107625308628737622865Q;
    const double wa = (double)((322 + 13 * SQRT70) / 900), wb = (double)((322 - 13 * SQRT70) / 900);
    x[0] = -b; x[1] = -a; x[2] = 0; x[3] = a; x[4] = b;
    w[0] = wb; w[1] = wa; w[2] = (double)(128.0Q / 225.0Q); w[3] = wa; w[4] = wb; return 0; }
  case 6: {
    const double a = 0.932469514203152, b = 0.661209386466265, c = 0.238619186083197;
    const double wa = 0.171324492379170, wb = 0.360761573048139, wc = 0.467913934572691;
    x[0] = -a; x[1] = -b; x[2] = -c; x[3] = c; x[4] = b; x[5] = a;
    w[0] = wa; w[1] = wb; w[2] = wc; w[3] = wc; w[4] = wb; w[5] = wa; return 0; }
  }
  return -1; /* "degree %d not implemented for Gauss quadrature" */
}

/* quadrature.go:22-48: tensor rule, x outermost, z innermost, w = wx*wy*wz */
static int uniform_gauss(int nx, int ny, int nz, double *pos, double *w, int cap) {
  double x[6], wx[6], y[6], wy[6], z[6], wz[6];
  if (gauss1d(nx, x, wx) || gauss1d(ny, y, wy) || gauss1d(nz, z, wz)) return -FEMO_ERR_QUADRATURE;
  if (nx * ny * nz > cap) return -FEMO_ERR_QUADRATURE;
  int count = 0;
  for (int i = 0; i < nx; i++)
    for (int j = 0; j < ny; j++)
      for (int k = 0; k < nz; k++) {
        w[count] = wx[i] * wy[j] * wz[k];
        pos[3 * count] = x[i]; pos[3 * count + 1] = y[j]; pos[3 * count + 2] = z[k];
        count++;
      }
  return count;
}

/* quadrature.go:9-18: 3-D rule with nz=1 (z=0, wz=2), then weights halved */
static int uniform_gauss2d(int nx, int ny, double *pos, double *w, int cap) {
  int n = uniform_gauss(nx, ny, 1, pos, w, cap);
  for (int i = 0; i < n; i++) w[i] /= 2;
  return n;
}

/* triangles.go:158-212: tabulated rules, weights * 0.5 */
static int triangle_quads(int order, double *pos, double *w, int cap) {
  static const double w1[] = {1};
  static const double p1[] = {1.0 / 3, 1.0 / 3};
  static const double w2[] = {1.0 / 3, 1.0 / 3, 1.0 / 3};
  static const double p2[] = {1.0 / 6, 2.0 / 3, 1.0 / 6, 1.0 / 6, 2.0 / 3, 1.0 / 6};
  static const double w3[] = {-0.5625, 0.520833333333333, 0.520833333333333, 0.520833333333333};
  static const double p3[] = {1.0 / 3, 1.0 / 3, 0.2, 0.6, 0.2, 0.2, 0.6, 0.2};
  static const double w4[] = {0.223381589678011, 0.223381589678011, 0.223381589678011,
                              0.109951743655322, 0.109951743655322, 0.109951743655322};
  static const double p4[] = {0.445948490915965, 0.108103018168070, 0.445948490915965, 0.445948490915965,
                              0.108103018168070, 0.445948490915965, 0.091576213509771, 0.816847572980459,
                              0.091576213509771, 0.091576213509771, 0.816847572980459, 0.091576213509771};
  const double *ws[] = {w1, w2, w3, w4}, *ps[] = {p1, p2, p3, p4};
  const int ns[] = {1, 3, 4, 6};
  if (order < 1 || order > 4) return -FEMO_ERR_QUADRATURE; /* panics in the reference */
  int n = ns[order - 1];
  if (n > cap) return -FEMO_ERR_QUADRATURE;
  for (int i = 0; i < n; i++) {
    pos[3 * i] = ps[order - 1][2 * i]; pos[3 * i + 1] = ps[order - 1][2 * i + 1]; pos[3 * i + 2] = 0;
    w[i] = ws[order - 1][i] * 0.5;
  }
  return n;
}

int femo_elem_quadrature(int kind, int order, double *pos, double *w, int cap) {
  switch (kind) {
  case FEMO_TETRA4: /* tetra4.go:48-50: one point, weight 1 (8-Q1) */
    if (cap < 1) return -FEMO_ERR_QUADRATURE;
    pos[0] = .25; pos[1] = .25; pos[2] = .25; w[0] = 1;
    return 1;
  case FEMO_TETRA10: { /* tetra10.go:98-112 */
    const double a = (double)((5 + 3 * 2.2360679774997896964091736687312762354Q) / 20);
    const double b = (double)((5 - 2.2360679774997896964091736687312762354Q) / 20);
    const double p[12] = {a, b, b, b, b, b, b, b, a, b, a, b};
    if (cap < 4) return -FEMO_ERR_QUADRATURE;
    memcpy(pos, p, sizeof p);
    for (int i = 0; i < 4; i++) w[i] = 0.25;
    return 4;
  }
  case FEMO_HEXA8: /* hexa8.go:98-108: order==0 -> 2 */
    if (order == 0) order = 2;
    return uniform_gauss(order, order, order, pos, w, cap);
  case FEMO_HEXA20: /* hexa20.go:166-176: order==0 -> 3 */
    if (order == 0) order = 3;
    return uniform_gauss(order, order, order, pos, w, cap);
  case FEMO_QUAD4: /* quad4.go:74-88: order<=0 -> 2 */
    if (order <= 0) order = 2;
    return uniform_gauss2d(order, order, pos, w, cap);
  case FEMO_QUAD8: /* quad8.go:91-104: order<=0 -> 3 */
    if (order <= 0) order = 3;
    return uniform_gauss2d(order, order, pos, w, cap);
  case FEMO_TRIANGLE3: /* triangles.go:68-76: order<=0 -> 1 */
    if (order <= 0) order = 1;
    return triangle_quads(order, pos, w, cap);
  case FEMO_TRIANGLE6: /* triangles.go:142-152: order<=0 -> 2 */
    if (order <= 0) order = 2;
    return triangle_quads(order, pos, w, cap);
  }
  return -FEMO_ERR_ELEMENT;
}
