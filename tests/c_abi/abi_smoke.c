/* abi_smoke.c -- a plain C translation unit compiled against include/femb200.h and linked to libfemb200.so: what a cgo
 * (or any FFI) binding sees.  Assembles the single Tetra4 of the reference's ExampleGeneralAssembler
 * (example_test.go:52-86: nodes (0,0,0) (1,0,0) (0,1,0) (0,0,1), solids.Isotropic{E: 200e9, Poisson: 0.3}.Solid3D())
 * through the C ABI only -- tables written out by hand exactly as elements/tetra4.go and constitution/solids/isotropic.go
 * define them -- and prints the dense 12x12 K with "%.5g" like the example's Output block, one row per line.  The pytest
 * wrapper (tests/test_c_abi.py) compares the text with the golden vector; header drift (a changed struct layout or
 * signature) breaks either the compilation or the numbers.
 *   gcc -std=c99 -Wall -Werror -Iinclude tests/c_abi/abi_smoke.c -Lgo-fem_b200 -lfemb200 -Wl,-rpath,$PWD/go-fem_b200 -o tests/c_abi/abi_smoke
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "femb200.h"

#define CHECK(call)                                                                 \
  do {                                                                              \
    int rc_ = (call);                                                               \
    if (rc_ != 0) { fprintf(stderr, "%s -> %d\n", #call, rc_); return 10 + rc_; }   \
  } while (0)

int main(int argc, char **argv) {
  if (argc > 1 && strcmp(argv[1], "--abi") == 0) { /* CPU-only use: the library loads and reports its ABI version */
    printf("%d %d\n", femb200_abi_version(), FEMB200_ABI_VERSION);
    return femb200_abi_version() == FEMB200_ABI_VERSION ? 0 : 1;
  }
  const double nodes[4 * 3] = {0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0, 1};
  const int64_t conn[4] = {0, 1, 2, 3};             /* a Go []int */
  /* tetra4.go:23-50: N = (1-x-y-z, x, y, z) at (1/4,1/4,1/4); BasisDiff rows d/dx, d/dy, d/dz; one point, weight 1 */
  const double Npg[4] = {0.25, 0.25, 0.25, 0.25};
  const double dNpg[3 * 4] = {-1, 1, 0, 0, -1, 0, 1, 0, -1, 0, 0, 1};
  const double wpg[1] = {1.0};
  /* isotropic.go:13-34 */
  const double E = 200e9, nu = 0.3;
  const double G = E / (2 + 2 * nu), lam = E * nu / ((1 + nu) * (1 - 2 * nu));
  double Cm[36];
  memset(Cm, 0, sizeof Cm);
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) Cm[i * 6 + j] = lam;
    Cm[i * 6 + i] = lam + 2 * G;
    Cm[(3 + i) * 6 + 3 + i] = G;
  }
  const int32_t dofmap[3] = {0, 1, 2};
  femb200_elem_desc d;
  memset(&d, 0, sizeof d);
  d.d = 3; d.nn = 4; d.ndn = 3; d.nqp = 1; d.dimC = 6; d.bkind = FEMB200_B_XYZ; d.model_dofs_per_node = 3;
  d.Npg = Npg; d.dNpg = dNpg; d.wpg = wpg; d.C = Cm; d.dofmap = dofmap;

  femb200_ctx *ctx = NULL;
  femb200_assembler *ga = NULL;
  CHECK(femb200_ctx_create(0, &ctx));
  CHECK(femb200_assembler_create(ctx, nodes, 4, 3, 0, &ga));
  CHECK(femb200_add_isoparametric(ga, &d, conn, 1, 0, 1));
  const int64_t n = femb200_total_dofs(ga), nnz = femb200_nnz(ga);
  if (n != 12 || nnz != 144) { fprintf(stderr, "dims %lld nnz %lld\n", (long long)n, (long long)nnz); return 2; }
  int64_t indptr[13];
  int32_t indices[144];
  double values[144], K[12][12];
  CHECK(femb200_csr_copy(ga, indptr, indices, values));
  memset(K, 0, sizeof K);
  for (int r = 0; r < 12; ++r)
    for (int64_t k = indptr[r]; k < indptr[r + 1]; ++k) K[r][indices[k]] = values[k];
  /* block view must agree: one block row per node, 4 blocks each */
  int64_t blkptr[5];
  int32_t blkcols[16];
  if (femb200_bsr_blocks(ga) != 16) return 3;
  CHECK(femb200_bsr_copy(ga, blkptr, blkcols, NULL));
  for (int I = 0; I < 4; ++I)
    if (blkptr[I + 1] - blkptr[I] != 4 || blkcols[blkptr[I] + 2] != 2) return 4;
  /* lap.Sparse.At through the ABI */
  double k00 = 0;
  CHECK(femb200_csr_at(ga, 0, 0, &k00));
  if (k00 != K[0][0]) return 5;
  for (int r = 0; r < 12; ++r) {
    for (int c = 0; c < 12; ++c) printf("%s%.5g", c ? " " : "", K[r][c]);
    printf("\n");
  }
  /* an inverted element: error code, lowest element, K untouched */
  const int64_t bad[4] = {0, 2, 1, 3};
  int rc = femb200_add_isoparametric(ga, &d, bad, 1, 0, 1);
  int32_t kind = 0, qp = 0;
  int64_t elem = -1;
  femb200_last_error(ga, &kind, &elem, &qp);
  if (rc != FEMB200_ERR_NEG_DET || kind != FEMB200_ERR_NEG_DET || elem != 0 || femb200_nnz(ga) != 144) return 6;
  femb200_assembler_destroy(ga);
  femb200_ctx_destroy(ctx);
  return 0;
}
