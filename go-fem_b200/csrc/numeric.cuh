// numeric.cuh -- device helpers shared by the numeric kernels (numeric.cu, srows.cu): strain-displacement block
// structure, table access, Jacobian determinant / inverse, per-quadrature-point geometry, staged record layout.
#pragma once
#include "common.cuh"
#include <type_traits>

namespace femb200 {

static __constant__ double c_tab[kTabDoubles]; // per translation unit; only numeric.cu uploads and reads its copy

// ---- strain-displacement block structure ---------------------------------------------------------
// Node "values" v[] and the selector sel(k,i): B_node[k][i] = v[sel] (or 0 when sel < 0).
//  XYZ            v = (gx, gy, gz)              straindisplacement.go:5-26
//  PLANE          v = (gx, gy)                  straindisplacement.go:28-40
//  AXISYM         v = (gr, gz, N/r)             straindisplacement.go:42-58
//  THERMAL        v = (g0..g_{d-1})             thermal.go:18-21,31-34 (B = dN)
//  THERMAL_AXISYM v = (N/r + gr, gz)            thermal.go:43-56
template <int BKIND, int D> struct BT;
template <> struct BT<FEMB200_B_XYZ, 3> {
  static constexpr int NDN = 3, DIMC = 6, NV = 3; static constexpr bool RADIAL = false;
  __host__ __device__ static constexpr int sel(int k, int i) {
    constexpr int t[6][3] = {{0, -1, -1}, {-1, 1, -1}, {-1, -1, 2}, {1, 0, -1}, {-1, 2, 1}, {2, -1, 0}};
    return t[k][i];
  }
};
template <> struct BT<FEMB200_B_PLANE, 2> {
  static constexpr int NDN = 2, DIMC = 3, NV = 2; static constexpr bool RADIAL = false;
  __host__ __device__ static constexpr int sel(int k, int i) {
    constexpr int t[3][2] = {{0, -1}, {-1, 1}, {1, 0}};
    return t[k][i];
  }
};
template <> struct BT<FEMB200_B_AXISYM, 2> {
  static constexpr int NDN = 2, DIMC = 4, NV = 3; static constexpr bool RADIAL = true;
  __host__ __device__ static constexpr int sel(int k, int i) {
    constexpr int t[4][2] = {{0, -1}, {2, -1}, {-1, 1}, {1, 0}};
    return t[k][i];
  }
};
template <int D> struct BT<FEMB200_B_THERMAL, D> {
  static constexpr int NDN = 1, DIMC = D, NV = D; static constexpr bool RADIAL = false;
  __host__ __device__ static constexpr int sel(int k, int i) { return k; }
};
template <> struct BT<FEMB200_B_THERMAL_AXISYM, 2> {
  static constexpr int NDN = 1, DIMC = 2, NV = 2; static constexpr bool RADIAL = true;
  __host__ __device__ static constexpr int sel(int k, int i) { return k; }
};

// ---- table access: __constant__ (uniform broadcast) or global fallback for oversized tables -------
template <bool TAB_CONST> struct Tab {
  const double *g;
  int offW, offN, offDN;
  __device__ __forceinline__ double at(int i) const {
    if constexpr (TAB_CONST) return c_tab[i]; else return __ldg(g + i);
  }
  template <int DIMC> __device__ __forceinline__ double C(int k, int c) const { return at(k * DIMC + c); }
  __device__ __forceinline__ double w(int q) const { return at(offW + q); }
  template <int NN> __device__ __forceinline__ double N(int q, int b) const { return at(offN + q * NN + b); }
  template <int D, int NN> __device__ __forceinline__ double dN(int q, int i, int b) const { return at(offDN + (q * D + i) * NN + b); }
};

// ---- d x d determinant and inverse --------------------------------------------------------------
// The determinant follows the reference's algorithm class: LU with partial pivoting (gonum mat.Det ->
// Dgetf2, first-max pivot, multipliers = a * (1/pivot)), det = sign * prod(u_ii), so that exactly singular
// Jacobians (collapsed elements: two identical rows) give EXACTLY 0 -> "zero determinant", and the
// sign of a degenerate element is not decided by FMA rounding noise.  The inverse is the adjugate
// scaled by 1/det (differences to the reference's LU solve are O(eps * cond)).
// LuFac: the factors of that LU (unit lower multipliers below the diagonal, U on and above), the row interchanges and the
// reciprocals of the pivots -- what gonum's Dense.Solve (Dgetrf + Dgetrs) works with.
template <int D> struct LuFac {
  double a[D][D];
  double rdiag[D];
  int p[D];
};

template <int D> __device__ __forceinline__ double lu_factor(const double (&Jin)[D][D], LuFac<D> &f) {
  double (&A)[D][D] = f.a;
#pragma unroll
  for (int i = 0; i < D; ++i)
#pragma unroll
    for (int j = 0; j < D; ++j) A[i][j] = Jin[i][j];
  double det = 1.0;
#pragma unroll
  for (int j = 0; j < D; ++j) {
    // first index of the largest |A[i][j]|, i >= j; swap it into row j (predicated, static indices)
    double best = fabs(A[j][j]);
    int p = j;
#pragma unroll
    for (int i = j + 1; i < D; ++i) {
      const double v = fabs(A[i][j]);
      if (v > best) { best = v; p = i; }
    }
    f.p[j] = p;
#pragma unroll
    for (int i = j + 1; i < D; ++i)
      if (p == i) {
#pragma unroll
        for (int k = 0; k < D; ++k) { const double t = A[j][k]; A[j][k] = A[i][k]; A[i][k] = t; }
        det = -det;
      }
    const double piv = A[j][j];
    det = __dmul_rn(det, piv);
    f.rdiag[j] = 1.0 / piv;
    if (piv != 0.0) {
      const double r = f.rdiag[j];
#pragma unroll
      for (int i = j + 1; i < D; ++i) {
        const double l = __dmul_rn(A[i][j], r);
        A[i][j] = l;
#pragma unroll
        for (int k = j + 1; k < D; ++k) A[i][k] = __dsub_rn(A[i][k], __dmul_rn(l, A[j][k])); // no FMA: same rounding as the reference's LU
      }
    }
  }
  return det;
}

template <int D> __device__ __forceinline__ double lu_det(const double (&Jin)[D][D]) {
  LuFac<D> f;
  return lu_factor<D>(Jin, f);
}

// x <- J^-1 x with the factors above, operation by operation as Dgetrs does it (row interchanges, unit-lower forward
// substitution, upper backward substitution with a multiplication by the reciprocal pivot; no FMA):
// dNxy.Solve(jac, dN) of assembler_isoparametric.go:100 column by column.  An adjugate inverse differs from this by
// O(eps*cond) of the LARGEST term -- 1e-10 relative on gradient components that nearly cancel (measured on the Tetra10
// benchmark mesh), which showed up as 9.6e-13 normalised error in K; with the same operations the gradients agree.
template <int D> __device__ __forceinline__ void lu_solve(const LuFac<D> &f, double (&x)[D]);
template <> __device__ __forceinline__ void lu_solve<1>(const LuFac<1> &f, double (&x)[1]) { x[0] = __dmul_rn(x[0], f.rdiag[0]); }
template <> __device__ __forceinline__ void lu_solve<2>(const LuFac<2> &f, double (&x)[2]) {
  const bool s01 = f.p[0] == 1;
  double x0 = s01 ? x[1] : x[0], x1 = s01 ? x[0] : x[1];
  x1 = __dsub_rn(x1, __dmul_rn(f.a[1][0], x0));            // (Dtrsm skips zero multipliers: the same value up to the sign of a zero)
  x1 = __dmul_rn(x1, f.rdiag[1]);
  x0 = __dsub_rn(x0, __dmul_rn(f.a[0][1], x1));
  x0 = __dmul_rn(x0, f.rdiag[0]);
  x[0] = x0; x[1] = x1;
}
template <> __device__ __forceinline__ void lu_solve<3>(const LuFac<3> &f, double (&x)[3]) {
  // scalars and selects only: an indexed swap of x[] sent the vector to local memory
  const bool s01 = f.p[0] == 1, s02 = f.p[0] == 2, s12 = f.p[1] == 2;
  const double y0 = s01 ? x[1] : (s02 ? x[2] : x[0]);      // rows 0 <-> p[0]
  const double y1 = s01 ? x[0] : x[1], y2 = s02 ? x[0] : x[2];
  double x0 = y0, x1 = s12 ? y2 : y1, x2 = s12 ? y1 : y2;  // rows 1 <-> p[1]
  x1 = __dsub_rn(x1, __dmul_rn(f.a[1][0], x0));            // (Dtrsm skips zero multipliers: the same value up to the sign of a zero)
  x2 = __dsub_rn(x2, __dmul_rn(f.a[2][0], x0));
  x2 = __dsub_rn(x2, __dmul_rn(f.a[2][1], x1));
  x2 = __dmul_rn(x2, f.rdiag[2]);
  x1 = __dsub_rn(x1, __dmul_rn(f.a[1][2], x2));
  x1 = __dmul_rn(x1, f.rdiag[1]);
  x0 = __dsub_rn(x0, __dmul_rn(f.a[0][1], x1));
  x0 = __dsub_rn(x0, __dmul_rn(f.a[0][2], x2));
  x0 = __dmul_rn(x0, f.rdiag[0]);
  x[0] = x0; x[1] = x1; x[2] = x2;
}

// adjugate inverse scaled by 1/det (only for the condition check and the Strains helper: O(eps*cond) off the LU solve)
template <int D> __device__ __forceinline__ double inv_det(const double (&J)[D][D], double (&inv)[D][D], double det);
template <> __device__ __forceinline__ double inv_det<1>(const double (&J)[1][1], double (&inv)[1][1], double det) {
  inv[0][0] = 1.0 / J[0][0];
  return det;
}
template <> __device__ __forceinline__ double inv_det<2>(const double (&J)[2][2], double (&inv)[2][2], double det) {
  const double r = 1.0 / det;
  inv[0][0] = J[1][1] * r; inv[0][1] = -J[0][1] * r;
  inv[1][0] = -J[1][0] * r; inv[1][1] = J[0][0] * r;
  return det;
}
template <> __device__ __forceinline__ double inv_det<3>(const double (&J)[3][3], double (&inv)[3][3], double det) {
  const double r = 1.0 / det;
  inv[0][0] = (J[1][1] * J[2][2] - J[1][2] * J[2][1]) * r;
  inv[1][0] = (J[1][2] * J[2][0] - J[1][0] * J[2][2]) * r;
  inv[2][0] = (J[1][0] * J[2][1] - J[1][1] * J[2][0]) * r;
  inv[0][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) * r;
  inv[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) * r;
  inv[2][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) * r;
  inv[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) * r;
  inv[1][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) * r;
  inv[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) * r;
  return det;
}
template <int D> __device__ __forceinline__ double norm1(const double (&A)[D][D]) {
  double m = 0;
#pragma unroll
  for (int j = 0; j < D; ++j) {
    double s = 0;
#pragma unroll
    for (int i = 0; i < D; ++i) s += fabs(A[i][j]);
    m = fmax(m, s);
  }
  return m;
}

// Geometry of one quadrature point.  Returns the error kind (0 = ok) exactly where the reference
// would have returned it (assembler_isoparametric.go:95-107).
template <int D, int NN, bool RADIAL, class TabT>
__device__ __forceinline__ int qp_geometry(const TabT &tab, int q, const double (&X)[NN][D], LuFac<D> &fac,
                                           double &detJ, double &rinv, double &scale) {
  double J[D][D];
#pragma unroll
  for (int i = 0; i < D; ++i)
#pragma unroll
    for (int j = 0; j < D; ++j) J[i][j] = 0.0;
#pragma unroll
  for (int b = 0; b < NN; ++b)
#pragma unroll
    for (int i = 0; i < D; ++i) {
      const double dn = tab.template dN<D, NN>(q, i, b);
#pragma unroll
      for (int j = 0; j < D; ++j) J[i][j] = __dadd_rn(J[i][j], __dmul_rn(dn, X[b][j])); // jac.Mul(dN, X): node-ascending, unfused (:93)
    }
  detJ = lu_factor<D>(J, fac);
  if (detJ < 0) return FEMB200_ERR_NEG_DET;
  if (detJ < 1e-12) return FEMB200_ERR_ZERO_DET;
  {
    double inv[D][D];
    inv_det<D>(J, inv, detJ);
    if (norm1<D>(J) * norm1<D>(inv) > 1e16) return FEMB200_ERR_SOLVE;
  }
  scale = 1.0;
  rinv = 0.0;
  if constexpr (RADIAL) {
    double r = 0.0;
#pragma unroll
    for (int b = 0; b < NN; ++b) r = fma(X[b][0], tab.template N<NN>(q, b), r);
    rinv = 1.0 / r;
    scale = r;
    if (isnan(r)) return FEMB200_ERR_NAN_SCALE;
  }
  return 0;
}

// v[] of one node at quadrature point q (see BT above)
template <int BKIND, int D, int NN, class TabT>
__device__ __forceinline__ void node_vals(const TabT &tab, int q, int b, const LuFac<D> &fac, double rinv,
                                          double (&v)[BT<BKIND, D>::NV]) {
  double g[D];
#pragma unroll
  for (int i = 0; i < D; ++i) g[i] = tab.template dN<D, NN>(q, i, b);
  lu_solve<D>(fac, g);                                      // column b of dNxy.Solve(jac, dN) (:100)
  if constexpr (BKIND == FEMB200_B_AXISYM) {
    v[0] = g[0]; v[1] = g[1]; v[2] = tab.template N<NN>(q, b) * rinv;
  } else if constexpr (BKIND == FEMB200_B_THERMAL_AXISYM) {
    v[0] = tab.template N<NN>(q, b) * rinv + g[0]; v[1] = g[1];
  } else {
#pragma unroll
    for (int i = 0; i < D; ++i) v[i] = g[i];
  }
}

// E = f * B_a^T C   (NDN x DIMC), only the structural non-zeros of B_a are touched
template <int BKIND, int D, class TabT>
__device__ __forceinline__ void bt_c(const TabT &tab, const double (&va)[BT<BKIND, D>::NV], double f,
                                     double (&E)[BT<BKIND, D>::NDN][BT<BKIND, D>::DIMC]) {
  using B = BT<BKIND, D>;
  double vs[B::NV];
#pragma unroll
  for (int i = 0; i < B::NV; ++i) vs[i] = va[i] * f;
#pragma unroll
  for (int i = 0; i < B::NDN; ++i)
#pragma unroll
    for (int c = 0; c < B::DIMC; ++c) {
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < B::DIMC; ++k)
        if (B::sel(k, i) >= 0) s = fma(vs[B::sel(k, i)], tab.template C<B::DIMC>(k, c), s);
      E[i][c] = s;
    }
}

// blk = E * B_b  (NDN x NDN)
template <int BKIND, int D>
__device__ __forceinline__ void e_b(const double (&E)[BT<BKIND, D>::NDN][BT<BKIND, D>::DIMC],
                                    const double (&vb)[BT<BKIND, D>::NV],
                                    double (&blk)[BT<BKIND, D>::NDN * BT<BKIND, D>::NDN]) {
  using B = BT<BKIND, D>;
#pragma unroll
  for (int i = 0; i < B::NDN; ++i)
#pragma unroll
    for (int j = 0; j < B::NDN; ++j) {
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < B::DIMC; ++k)
        if (B::sel(k, j) >= 0) s = fma(E[i][k], vb[B::sel(k, j)], s);
      blk[i * B::NDN + j] = s;
    }
}


template <int BKIND, int D, int NN> struct Stag {
  using B = BT<BKIND, D>;
  static constexpr bool FIN = (B::NV == 3);                 // f stored in every node's 4th slot
  static constexpr int NVP = FIN ? 4 : B::NV;               // doubles per node
  static constexpr int REC = FIN ? NN * 4 : ((NN * B::NV + 2) & ~1);
  static constexpr int FOFF = NN * B::NV;                   // !FIN: position of f
};

__device__ __forceinline__ void ld256(const double *p, double (&v)[4]) {
  asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p));
}
__device__ __forceinline__ void st256(double *p, double a, double b, double c, double d) {
  asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// ---- mbarrier + bulk (TMA 1-D) copy helpers ---------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes),
               "r"(smem_u32(bar)) : "memory");
}

// Staged geometry for the element-block kernel (keblocks.cu): per (element, point) the node gradients as three arrays
// gx[NNP] gy[NNP] gz[NNP] (NNP = nn rounded up to 4: every array 32-byte aligned) followed by [f 0 0 0] -- in shared memory a
// lane's two adjacent nodes are one conflict-free 128-bit load per component.
template <int NN> struct StagSoa {
  static constexpr int NNP = (NN + 3) & ~3;
  static constexpr int QREC = 3 * NNP + 4;
};

struct RowArgs {
  int64_t n_nodes;
  int64_t n_numeric;          // elements [n_numeric, n_elem) are pattern-only (halo): no values
  const int32_t *n2e_ptr;
  const uint32_t *n2e_t;
  const uint16_t *pos16;
  const int64_t *blkptr;
  const double *stag;
  const uint8_t *degen;
  const int32_t *order;       // visiting order of the node rows (NULL: identity)
  int64_t v0, v1;             // visiting range of this launch
  double *values;
  int64_t ninc;               // n_elem * nn (length of n2e_t)
  int nqp;
  int has_halo;               // some incidences belong to pattern-only elements (>= n_numeric): the incidence list must be consulted
  double C[kMaxDimC * kMaxDimC]; // constitutive matrix (row-major dimC x dimC): kernel-parameter space = constant bank
};

// srows.cu: the scalar-row kernel (thread per scalar CSR row); -2 = does not apply, keep the node-row kernel
int srows_staged(femb200_ctx *ctx, const ElemShape &sh, const RowArgs &a, const int32_t *order, int64_t v0, int64_t v1, int64_t rowlen_cap);

int srows_fused(femb200_ctx *ctx, const ElemShape &sh, const RowArgs &a, const double *d_nodes, const int32_t *blkcols,
                const double *h_tab, const int32_t *order, int64_t v0, int64_t v1, int64_t rowlen_cap, double4 *nodes4, int64_t cap_blk,
                const int32_t *rowcols, int cols_stride);

// keblocks.cu: element blocks + ordered row gather for Hexa8 / Tetra10 / Hexa20 solids (XYZ filler)
size_t ke_doubles_per_element(int nn);
size_t ke_blocks_smem(int nn, int nqp);
int ke_elem_blocks(femb200_ctx *ctx, int nn, int64_t n_elem, const double *stag, int nqp, double *ke, const double *C36);
int ke_rows_gather(femb200_ctx *ctx, int nn, const RowArgs &a, const double *ke, int64_t rowlen_cap);

} // namespace femb200
