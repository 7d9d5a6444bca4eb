// numeric.cu -- numeric phase + deterministic row reduction, fused in ONE kernel.
//
// Reference semantics (assembler_isoparametric.go:85-117, per element, per quadrature point):
//   jac = dN*X (:93); dJac = det(jac) with the <0 / <1e-12 checks (:94-99); dNxy = jac^-1 dN (:100);
//   B filled by the constituter (:104, straindisplacement.go / thermal.go); Ke += B^T C B * dJac*w*scale (:109-112);
//   then nd^2 COO triplets per element (assembler.go:79-92) summed by lap.Sparse.Accumulate (:121).
//
// B200 design: ROW-centric gather.  One thread owns one node-block row I of K.  It walks the elements
// incident to node I in ascending element index (n2e lists of the symbolic phase), recomputes that
// element's geometry in FP64 registers, forms only block-row a of Ke (a = local index of I in e):
//     Ke[a][b] = sum_q (B_a^T C) B_b * dJac*w*scale      (ndn x ndn block, B sparsity exploited)
// and adds it to the row's accumulators in shared memory (interleaved by thread => bank-conflict free).
// No COO stream, no Ke staging in HBM, no atomics: every K entry is a fixed-order (ascending element)
// sum, i.e. a deterministic segmented reduction whose segments are the CSR rows.  HBM traffic is the
// compulsory one: connectivity + coordinates (L2 resident) in, CSR values out.
#include "numeric.cuh"
#include <cstdlib>
#include <cstring>

namespace femb200 {

struct NumArgs {
  int64_t n_nodes;
  int64_t n_numeric;          // elements [n_numeric, n_elem) are pattern-only (halo): no values
  const double *nodes;        // [n_nodes][3]
  const int32_t *conn;        // [n_elem][NN]
  const int32_t *n2e_ptr;     // [n_nodes+1]
  const uint32_t *n2e_t;      // [n_elem*NN]
  const uint16_t *pos16;      // [n_elem*NN][NN]
  const int64_t *blkptr;      // [n_nodes+1]
  double *values;             // CSR values of this call
  unsigned long long *err;
  const double *tab_g;
  const double *h_tab;        // host copy of the table blob (C first)
  int offW, offN, offDN, nqp;
  int nu;                     // distinct dofmap values (== NDN when identity)
  int64_t n_elem_total;       // numeric + pattern-only elements of the call
  int dof2u[kMaxNdn];
};

// SMEM_ACC: row accumulators in shared memory R[(p*NB2 + k)*T + tid] (requires identity dofmap);
// otherwise they are the final CSR slots in global memory (pre-zeroed), which also handles the
// reference's non-injective dofmap quirk (fem.go:155-172).
// ACC_REGS: the whole block-row of one element is summed over quadrature points in registers first.
template <int D, int NN, int BKIND, bool SMEM_ACC, bool ACC_REGS, bool TAB_CONST, int TPB>
__global__ void __launch_bounds__(TPB)
k_numeric_rows(const NumArgs a) {
  using B = BT<BKIND, D>;
  constexpr int NDN = B::NDN, NB2 = NDN * NDN, NV = B::NV;
  extern __shared__ double s_R[];
  const int T = blockDim.x, tid = threadIdx.x;
  const int64_t I = (int64_t)blockIdx.x * T + tid;
  if (I >= a.n_nodes) return;
  Tab<TAB_CONST> tab{a.tab_g, a.offW, a.offN, a.offDN};
  const int beg = a.n2e_ptr[I], end = a.n2e_ptr[I + 1];
  const int64_t b0 = a.blkptr[I];
  const int rl = (int)(a.blkptr[I + 1] - b0);
  if (rl == 0) return;                 // unreferenced node, or a row this rank does not assemble
  const int nu = SMEM_ACC ? NDN : a.nu;
  double *const vrow = a.values + (int64_t)nu * nu * b0; // this block row's slice of CSR values

  if constexpr (SMEM_ACC) {
    for (int m = 0; m < rl * NB2; ++m) s_R[m * T + tid] = 0.0;
  }
  auto add_block = [&](int p, const double(&blk)[NB2]) {
    if constexpr (SMEM_ACC) {
#pragma unroll
      for (int k = 0; k < NB2; ++k) s_R[(p * NB2 + k) * T + tid] += blk[k];
    } else {
#pragma unroll
      for (int i = 0; i < NDN; ++i)
#pragma unroll
        for (int j = 0; j < NDN; ++j)
          vrow[(int64_t)a.dof2u[i] * rl * nu + (int64_t)p * nu + a.dof2u[j]] += blk[i * NDN + j];
    }
  };

  for (int k = beg; k < end; ++k) {
    const uint32_t t = a.n2e_t[k];
    const int64_t e = t / NN;
    if (e >= a.n_numeric) continue;
    const int la = (int)(t % NN);
    double X[NN][D];
    {
      const int32_t *ce = a.conn + e * NN;
#pragma unroll
      for (int b = 0; b < NN; ++b) {
        const int64_t nb = __ldg(ce + b);
#pragma unroll
        for (int j = 0; j < D; ++j) X[b][j] = __ldg(a.nodes + nb * 3 + j);
      }
    }
    const uint16_t *pk = a.pos16 + (int64_t)k * NN;
    double acc[ACC_REGS ? NN : 1][NB2];
    if constexpr (ACC_REGS) {
#pragma unroll
      for (int b = 0; b < NN; ++b)
#pragma unroll
        for (int m = 0; m < NB2; ++m) acc[b][m] = 0.0;
    }
    bool bad = false;
    for (int q = 0; q < a.nqp; ++q) {
      LuFac<D> fac;
      double detJ, rinv, scale;
      const int ek = qp_geometry<D, NN, B::RADIAL>(tab, q, X, fac, detJ, rinv, scale);
      if (ek) {
        atomicMin(a.err, pack_error(e, q, ek));
        bad = true;
        break;
      }
      const double f = detJ * tab.w(q) * scale; // (dJac*w)*scale, left to right (:111)
      double va[NV], E[NDN][B::DIMC];
      node_vals<BKIND, D, NN>(tab, q, la, fac, rinv, va);
      bt_c<BKIND, D>(tab, va, f, E);
#pragma unroll
      for (int b = 0; b < NN; ++b) {
        double vb[NV], blk[NB2];
        node_vals<BKIND, D, NN>(tab, q, b, fac, rinv, vb);
        e_b<BKIND, D>(E, vb, blk);
        if constexpr (ACC_REGS) {
#pragma unroll
          for (int m = 0; m < NB2; ++m) acc[b][m] += blk[m];
        } else {
          add_block(pk[b] & pos_mask(NN), blk);
        }
      }
    }
    if constexpr (ACC_REGS) {
      if (!bad) {
#pragma unroll
        for (int b = 0; b < NN; ++b) add_block(pk[b] & pos_mask(NN), acc[b]);
      }
    }
  }

  if constexpr (SMEM_ACC) {
    // scalar CSR layout of the block row: scalar row i holds rl*NDN entries, entry (p, j) at p*NDN + j
    for (int i = 0; i < NDN; ++i)
      for (int p = 0; p < rl; ++p)
#pragma unroll
        for (int j = 0; j < NDN; ++j)
          vrow[(int64_t)i * rl * NDN + p * NDN + j] = s_R[(p * NB2 + i * NDN + j) * T + tid];
  }
}

// =====================================================================================================
// Two-kernel numeric path (default for every element type):
//
//  k_geometry  -- "one batched per-element kernel": one thread per element evaluates, per quadrature
//                 point, jac = dN*X, det (LU), the reference's error checks, J^-1 dN and the scale, and stages
//                 the node values v_b[q] (see BT) plus f[q] = dJac*w*scale.  Record layout per (e, q):
//                    NV == 3 :  [v0 v1 v2 f] per node  (32-byte aligned: one 256-bit load per node, f rides along)
//                    else    :  [v_0 .. v_{nn-1} | f | pad]
//                 Errors are detected here, in element order (atomicMin on the packed error word); elements with
//                 a repeated node are flagged (degen[e]) for the ordered slow path of the row kernel.
//  k_rows      -- deterministic row reduction fused with the B^T C B products (see the comment on the kernel).
// =====================================================================================================


// STORE = false: the reference's per-element checks only (error word), nothing is staged -- the companion of the fused
// scalar-row kernel (srows.cu), which recomputes the geometry it needs.
// SOA: the record layout of StagSoa<NN> (numeric.cuh) for the element-block kernel instead of Stag's node-major one.
template <int D, int NN, int BKIND, bool TAB_CONST, bool STORE = true, bool SOA = false>
__global__ void __launch_bounds__(128, (NN <= 4 && D == 3 ? (STORE ? 4 : 6) : 1)) // linear tetrahedra: checks only 6 blocks per SM, staging (opt-in path) 4
k_geometry(int64_t n_elem, const double *__restrict__ nodes, const int32_t *__restrict__ conn, double *__restrict__ stag,
           uint8_t *__restrict__ degen, unsigned long long *err, const double *tab_g, int offW, int offN, int offDN, int nqp,
           int64_t e_off = 0) {   // e_off: index of element 0 of this launch in the call (chunked launches; error word only)
  using B = BT<BKIND, D>;
  using S = Stag<BKIND, D, NN>;
  constexpr int NV = B::NV, REC = S::REC;
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_elem) return;
  Tab<TAB_CONST> tab{tab_g, offW, offN, offDN};
  double X[NN][D];
  int32_t nid[NN];
  if constexpr (NN % 4 == 0) {
#pragma unroll
    for (int v4 = 0; v4 < NN / 4; ++v4) {
      const int4 w = __ldg(reinterpret_cast<const int4 *>(conn + e * NN) + v4);
      nid[4 * v4] = w.x; nid[4 * v4 + 1] = w.y; nid[4 * v4 + 2] = w.z; nid[4 * v4 + 3] = w.w;
    }
  } else {
#pragma unroll
    for (int b = 0; b < NN; ++b) nid[b] = __ldg(conn + e * NN + b);
  }
  bool dup = false;
#pragma unroll
  for (int b = 0; b < NN; ++b) {
#pragma unroll
    for (int c = 0; c < b; ++c) dup = dup || (nid[b] == nid[c]);
#pragma unroll
    for (int j = 0; j < D; ++j) X[b][j] = __ldg(nodes + (int64_t)nid[b] * 3 + j);
  }
  if constexpr (STORE) degen[e] = dup ? 1 : 0;
  for (int q = 0; q < nqp; ++q) {
    LuFac<D> fac;
      double detJ, rinv, scale;
    const int ek = qp_geometry<D, NN, B::RADIAL>(tab, q, X, fac, detJ, rinv, scale);
    constexpr int RECQ = SOA ? StagSoa<NN>::QREC : REC;
    double *rec = stag + (e * nqp + q) * RECQ;
    if (ek) {
      atomicMin(err, pack_error(e + e_off, q, ek));
      if constexpr (STORE)
        for (int m = 0; m < RECQ; ++m) rec[m] = 0.0;
      return; // the reference stops at the first failing quadrature point of the element (:95-107)
    }
    if constexpr (!STORE) continue;
    const double f = detJ * tab.w(q) * scale; // (dJac*w)*scale, left to right (:111)
    if constexpr (SOA) {
      static_assert(!SOA || NV == 3, "three gradient components");
      constexpr int NNP = StagSoa<NN>::NNP;
#pragma unroll
      for (int v4 = 0; v4 < NNP / 4; ++v4) {                // four nodes at a time: one 256-bit store per component
        double g[3][4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int b = 4 * v4 + k;
          if (b < NN) {
            double vb[NV];
            node_vals<BKIND, D, NN>(tab, q, b, fac, rinv, vb);
            g[0][k] = vb[0]; g[1][k] = vb[NV > 1 ? 1 : 0]; g[2][k] = vb[NV > 2 ? 2 : 0];
          } else {
            g[0][k] = g[1][k] = g[2][k] = 0.0;
          }
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) st256(rec + c * NNP + 4 * v4, g[c][0], g[c][1], g[c][2], g[c][3]);
      }
      st256(rec + 3 * NNP, f, 0.0, 0.0, 0.0);
    } else if constexpr (S::FIN) {
#pragma unroll
      for (int b = 0; b < NN; ++b) {
        double vb[NV];
        node_vals<BKIND, D, NN>(tab, q, b, fac, rinv, vb);
        st256(rec + 4 * b, vb[0], vb[1], vb[2], f);
      }
    } else {
      double rv[REC];
#pragma unroll
      for (int b = 0; b < NN; ++b) {
        double vb[NV];
        node_vals<BKIND, D, NN>(tab, q, b, fac, rinv, vb);
#pragma unroll
        for (int v = 0; v < NV; ++v) rv[b * NV + v] = vb[v];
      }
      rv[S::FOFF] = f;
      if constexpr (S::FOFF + 1 < REC) rv[S::FOFF + 1] = 0.0;
#pragma unroll
      for (int c2 = 0; c2 < REC / 2; ++c2) reinterpret_cast<double2 *>(rec)[c2] = make_double2(rv[2 * c2], rv[2 * c2 + 1]);
    }
  }
}


// ---- the node-row kernel (default) ----------------------------------------------------------------------
// A group of G lanes owns one NODE-BLOCK row I of K, i.e. the NDN scalar CSR rows of node I at once.  Compared
// with k_rows (one scalar row per group) every staged record is loaded once instead of NDN times, the scalar
// row index i is a compile-time loop (so column i of B_a is known statically and its rows of C are read
// straight from the constant bank as DFMA operands: no registers spent on C), and the block needs NDN times
// fewer threads for the same shared-memory accumulators.  Everything else is as in k_rows: lane g takes the
// element nodes b = g, g+G, ..., keeps their partial sums over the quadrature points in FP64 registers, and
// after the quadrature loop adds its (distinct) column blocks into the row accumulators in shared memory;
// elements are visited in ascending index by the whole group => deterministic segmented reduction, no atomics.
// Loads run one step ahead of the arithmetic through two register buffers (A/B, swapped by unrolling the step
// loop by two), incidence ids three elements ahead, L2 prefetch of the staged record two elements ahead.
// Accumulators: entry m of node slot n at s_R[m * NPBP + n], NPBP odd => the per-lane read-modify-writes and
// the warp-cooperative, fully coalesced copy-out of the finished rows are both bank-conflict free.
template <int BKIND, int D, int I_, int M_> struct KRow { // row k of B with sel(k, I_) == M_  (-1: none)
  static constexpr int value() {
    using B = BT<BKIND, D>;
    int kr = -1;
    for (int k = 0; k < B::DIMC; ++k)
      if (B::sel(k, I_) == M_) kr = k;
    return kr;
  }
};

template <int BKIND, int D, int I_, int M_>
__device__ __forceinline__ void e_row_add(const RowArgs &a, double u, double (&E)[BT<BKIND, D>::DIMC]) {
  constexpr int KR = KRow<BKIND, D, I_, M_>::value();
  constexpr int DIMC = BT<BKIND, D>::DIMC;
  if constexpr (KR >= 0) {
#pragma unroll
    for (int c = 0; c < DIMC; ++c) E[c] = fma(u, a.C[KR * DIMC + c], E[c]);
  }
}

// PACKED: every node row of the block gets exactly its own rl*NDN*NDN accumulators (offsets by a warp scan) instead of
// max_rowlen interleaved ones -- quadratic meshes mix rows of very different lengths, and the FP64-bound elements
// do few read-modify-writes per flop, so a few bank conflicts cost nothing while the smaller footprint buys occupancy.
template <int D, int NN, int BKIND, int G, int TPB, bool PACKED>
__global__ void __launch_bounds__(TPB)
k_noderows(const RowArgs a) {
  using B = BT<BKIND, D>;
  using S = Stag<BKIND, D, NN>;
  constexpr int NDN = B::NDN, NV = B::NV, DIMC = B::DIMC, NVP = S::NVP, REC = S::REC;
  static_assert(TPB % 32 == 0 && 32 % G == 0, "G divides the warp");
  constexpr int NPB = TPB / G;              // node rows per block
  constexpr int NPBP = PACKED ? 1 : (NPB | 1); // stride of one row's accumulators (interleaved: odd, conflict free)
  constexpr int NPW = 32 / G;               // node rows per warp
  constexpr int NB = (NN + G - 1) / G;      // element nodes per lane
  extern __shared__ double s_R[];
  const int tid = threadIdx.x, lane = tid & 31;
  const int g = tid % G, nloc = tid / G;
  const int64_t vis = a.v0 + (int64_t)blockIdx.x * NPB + nloc;
  const int64_t I = vis < a.v1 ? (a.order ? (int64_t)__ldg(a.order + vis) : vis) : a.n_nodes;
  const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (lane & ~(G - 1)));
  int beg = 0, end = 0, rl = 0;
  int64_t b0 = 0;
  if (I < a.n_nodes) {
    beg = __ldg(a.n2e_ptr + I); end = __ldg(a.n2e_ptr + I + 1);
    b0 = __ldg(a.blkptr + I);
    rl = (int)(__ldg(a.blkptr + I + 1) - b0);
  }
  if (rl == 0) end = beg;                   // unreferenced node, or a row this rank does not assemble
  const int rowsz = rl * NDN;               // scalar row length
  const int tot = rowsz * NDN;              // entries of the node-block row: [i][p][j]
  int soff = nloc;                          // interleaved: entry m of row slot n at s_R[m * NPBP + n]
  if constexpr (PACKED) {                   // packed: rows back to back, offsets = exclusive scan over the warp's rows
    static_assert(!PACKED || TPB == 32, "packed accumulators: one warp per block");
    const int mine = (g == 0) ? tot : 0;
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    soff = __shfl_sync(0xffffffffu, incl - mine, lane & ~(G - 1));
  }
  double *const sacc = s_R + soff;
  for (int m = g; m < tot; m += G) sacc[m * NPBP] = 0.0;
  if constexpr (G > 1) __syncwarp(gmask);

  const int64_t recsz = (int64_t)a.nqp * REC;
  struct Buf { double va[NVP]; double f; double vb[NB][NVP]; };
  auto load_rec = [&](const double *rec, int la, Buf &o) {
    if constexpr (S::FIN) {
      ld256(rec + 4 * la, o.va);
      o.f = 0.0; // rides in va[3]
    } else {
      const double2 w = __ldg(reinterpret_cast<const double2 *>(rec) + la);
      o.va[0] = w.x; o.va[1] = w.y;
      o.f = __ldg(rec + S::FOFF);
    }
#pragma unroll
    for (int s = 0; s < NB; ++s) {
      const int b = g + G * s;
      if (NN % G == 0 || b < NN) {
        if constexpr (S::FIN) {
          ld256(rec + 4 * b, o.vb[s]);
        } else {
          const double2 w = __ldg(reinterpret_cast<const double2 *>(rec) + b);
          o.vb[s][0] = w.x; o.vb[s][1] = w.y;
        }
      } else {
#pragma unroll
        for (int m = 0; m < NVP; ++m) o.vb[s][m] = 0.0;
      }
    }
  };
  auto load_pos = [&](int k, int (&pos)[NB]) {
    const uint16_t *pk = a.pos16 + (int64_t)k * NN;
    if constexpr (G == 1 && NN % 4 == 0) {
#pragma unroll
      for (int v4 = 0; v4 < NN / 4; ++v4) {
        const uint2 w = __ldg(reinterpret_cast<const uint2 *>(pk) + v4);
        pos[4 * v4] = w.x & pos_mask(NN); pos[4 * v4 + 1] = w.x >> 16; pos[4 * v4 + 2] = w.y & 0xffff; pos[4 * v4 + 3] = w.y >> 16;
      }
    } else {
#pragma unroll
      for (int s = 0; s < NB; ++s) pos[s] = (g + G * s < NN) ? (int)(__ldg(pk + g + G * s) & pos_mask(NN)) : 0;
    }
  };

  // ---- pipeline state ------------------------------------------------------------------------------------
  // incidence lists are ascending in element index and pattern-only (halo) elements come last: the first halo
  // incidence ends the row's numeric work
  int kend = end, k = beg, q = 0, la = 0;
  uint32_t tA = beg < end ? __ldg(a.n2e_t + beg) : 0u;         // id of incidence k
  uint32_t tB = beg + 1 < end ? __ldg(a.n2e_t + beg + 1) : 0u; // k+1
  uint32_t tC = beg + 2 < end ? __ldg(a.n2e_t + beg + 2) : 0u; // k+2
  uint32_t tD = 0u;                                            // k+3, in flight
  if (beg < end && (int64_t)(tA / NN) >= a.n_numeric) kend = beg;
  bool more = false;
  const double *rec = a.stag;
  Buf bufA, bufB;
  int pos[NB], posn[NB];
#pragma unroll
  for (int s = 0; s < NB; ++s) { pos[s] = 0; posn[s] = 0; }
  bool dg = false, dgn = false;
  double acc[NDN][NB][NDN];
#pragma unroll
  for (int i = 0; i < NDN; ++i)
#pragma unroll
    for (int s = 0; s < NB; ++s)
#pragma unroll
      for (int j = 0; j < NDN; ++j) acc[i][s][j] = 0.0;

  // bookkeeping at the start of element k (ids tA, tB, tC are in registers)
  auto begin_element = [&]() {
    la = (int)(tA % NN);
    if (k + 3 < end) tD = __ldg(a.n2e_t + k + 3);
    if (k + 2 < end) {
      const char *pr = reinterpret_cast<const char *>(a.stag + (int64_t)(tC / NN) * recsz);
#pragma unroll
      for (int ln = 0; ln < (REC * 8 + 127) / 128 && ln < 4; ++ln) prefetch_l2(pr + 128 * ln);
    }
    // the slot map and the incidence list are streamed: touch the line two lines ahead when a new line starts
    if (((k * NN * 2) & 127) < NN * 2 && (int64_t)k * NN + 128 < a.ninc * NN) prefetch_l2(a.pos16 + (int64_t)k * NN + 128);
    if ((k & 31) == 0 && k + 64 < a.ninc) prefetch_l2(a.n2e_t + k + 64);
    more = (k + 1 < kend) && ((int64_t)(tB / NN) < a.n_numeric);
    if (k + 1 < kend && !more) kend = k + 1;
  };
  if (k < kend) {
    rec = a.stag + (int64_t)(tA / NN) * recsz;
    load_rec(rec, (int)(tA % NN), bufA);
    load_pos(k, pos);
    dg = __ldg(a.degen + tA / NN) != 0;
    if (beg + 1 < end) prefetch_l2(a.stag + (int64_t)(tB / NN) * recsz);
    begin_element();
  }

  auto step = [&](Buf &cur, Buf &nxt) {
    // 1. issue the loads of the next step
    const bool last_q = (q + 1 == a.nqp);
    const double *recn = rec;
    if (!last_q) {
      load_rec(rec + (int64_t)(q + 1) * REC, la, nxt);
    } else if (more) {
      recn = a.stag + (int64_t)(tB / NN) * recsz;
      load_rec(recn, (int)(tB % NN), nxt);
      load_pos(k + 1, posn);
      dgn = __ldg(a.degen + tB / NN) != 0;
    }
    // 2. row i of every block (a, b) of this quadrature point, i = 0..NDN-1
    const double f = S::FIN ? cur.va[NVP - 1] : cur.f;
    double u[NV];
#pragma unroll
    for (int m = 0; m < NV; ++m) u[m] = cur.va[m] * f;
    auto row = [&](auto ic) {
      constexpr int I_ = decltype(ic)::value;
      double E[DIMC];
#pragma unroll
      for (int c = 0; c < DIMC; ++c) E[c] = 0.0;
      e_row_add<BKIND, D, I_, 0>(a, u[0], E);
      if constexpr (NV > 1) e_row_add<BKIND, D, I_, (NV > 1 ? 1 : 0)>(a, u[NV > 1 ? 1 : 0], E);
      if constexpr (NV > 2) e_row_add<BKIND, D, I_, (NV > 2 ? 2 : 0)>(a, u[NV > 2 ? 2 : 0], E);
#pragma unroll
      for (int s = 0; s < NB; ++s) {
        if (NN % G == 0 || g + G * s < NN) {
#pragma unroll
          for (int j = 0; j < NDN; ++j) {
            double sj = acc[I_][s][j];
#pragma unroll
            for (int kk = 0; kk < DIMC; ++kk)
              if (B::sel(kk, j) >= 0) sj = fma(E[kk], cur.vb[s][B::sel(kk, j)], sj);
            acc[I_][s][j] = sj;
          }
        }
      }
    };
    row(std::integral_constant<int, 0>{});
    if constexpr (NDN > 1) row(std::integral_constant<int, (NDN > 1 ? 1 : 0)>{});
    if constexpr (NDN > 2) row(std::integral_constant<int, (NDN > 2 ? 2 : 0)>{});
    if (!last_q) { ++q; return; }
    // 3. end of the element: ordered accumulation into the shared row accumulators.  The nn column blocks of
    // one element are distinct unless the element is collapsed (a repeated node, flagged by k_geometry).
    if (!dg) {
#pragma unroll
      for (int i = 0; i < NDN; ++i) {
        double old[NB][NDN];
        double *const srow = sacc + i * rowsz * NPBP;
#pragma unroll
        for (int s = 0; s < NB; ++s)
#pragma unroll
          for (int j = 0; j < NDN; ++j) old[s][j] = srow[(pos[s] * NDN + j) * NPBP];
#pragma unroll
        for (int s = 0; s < NB; ++s)
          if (NN % G == 0 || g + G * s < NN) {
#pragma unroll
            for (int j = 0; j < NDN; ++j) srow[(pos[s] * NDN + j) * NPBP] = old[s][j] + acc[i][s][j];
          }
      }
    } else {
      // collapsed element: element nodes in ascending b, one at a time (same order as the reference's COO stream)
#pragma unroll
      for (int s = 0; s < NB; ++s) {
#pragma unroll 1
        for (int gg = 0; gg < G; ++gg) {
          if (gg == g && g + G * s < NN) {
#pragma unroll
            for (int i = 0; i < NDN; ++i)
#pragma unroll
              for (int j = 0; j < NDN; ++j) sacc[(i * rowsz + pos[s] * NDN + j) * NPBP] += acc[i][s][j];
          }
          if constexpr (G > 1) __syncwarp(gmask);
        }
      }
    }
    if constexpr (G > 1) __syncwarp(gmask);
#pragma unroll
    for (int i = 0; i < NDN; ++i)
#pragma unroll
      for (int s = 0; s < NB; ++s)
#pragma unroll
        for (int j = 0; j < NDN; ++j) acc[i][s][j] = 0.0;
    // advance to element k+1
    ++k; q = 0;
    tA = tB; tB = tC; tC = tD;
    rec = recn;
#pragma unroll
    for (int s = 0; s < NB; ++s) pos[s] = posn[s];
    dg = dgn;
    if (k < kend) begin_element();
  };
  while (k < kend) {
    step(bufA, bufB);
    if (!(k < kend)) break;
    step(bufB, bufA);
  }

  // ---- copy-out: the node rows of one warp are contiguous in the CSR values; every lane of the warp helps
  // with every row => 256-byte coalesced stores
  __syncwarp();
  double *const vbase = a.values + (int64_t)NDN * NDN * b0;
#pragma unroll 1
  for (int n = 0; n < NPW; ++n) {
    const int tot_n = __shfl_sync(0xffffffffu, tot, n * G);
    const unsigned long long vb_n = __shfl_sync(0xffffffffu, (unsigned long long)vbase, n * G);
    const int soff_n = __shfl_sync(0xffffffffu, soff, n * G);
    double *const vrow = reinterpret_cast<double *>(vb_n);
    const double *const src = s_R + soff_n;
    for (int m = lane; m < tot_n; m += 32) vrow[m] = src[m * NPBP];
  }
}

// lanes per node row: default per element size, FEMB200_G overrides (tuning); larger G = less shared memory per block
template <int NN> struct GroupOf { static constexpr int G = NN >= 8 ? 4 : (NN >= 4 ? 2 : 1); };

// one launch of the node-row kernel over the visiting range [v0, v1) with accumulators for rows of <= rowlen_cap blocks
template <int D, int NN, int BKIND, int G>
static int launch_noderows(femb200_ctx *ctx, const RowArgs &a0, const SymbolicResult *sym, const int32_t *order, int64_t v0, int64_t v1,
                           int64_t rowlen_cap) {
  using B = BT<BKIND, D>;
  constexpr int TPN = 32;                        // one warp per block
  constexpr int NPB = TPN / G;
  constexpr bool PACKED = NN == 8 || NN == 20;   // elements with many quadrature points per read-modify-write (see the kernel)
  constexpr int SLOT = NPB == 32 ? 3 : (NPB == 16 ? 2 : (NPB == 8 ? 1 : 0)); // group_need index for NPB rows
  static_assert(NPB >= 4 || !PACKED, "packed accumulators need at least 4 rows per warp");
  const size_t smem = PACKED ? (size_t)sym->group_need[SLOT] * B::NDN * B::NDN * sizeof(double)
                             : (size_t)rowlen_cap * B::NDN * B::NDN * sizeof(double) * (NPB | 1);
  if (smem > ctx->smem_optin) return -2;
  auto nrows = k_noderows<D, NN, BKIND, G, TPN, PACKED>;
  FEMB_CUDA_OK(ctx, cudaFuncSetAttribute(nrows, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smem_optin));
  RowArgs a = a0;
  a.order = order; a.v0 = v0; a.v1 = v1;
  if (v1 > v0) {
    nrows<<<grid_for(v1 - v0, NPB), TPN, smem, ctx->stream>>>(a);
    ctx->launches++;
  }
  FEMB_CUDA_OK(ctx, cudaGetLastError());
  return 0;
}

template <int D, int NN, int BKIND, bool TAB_CONST>
static int launch_dof_T(femb200_assembler *ga, const ElemShape &sh, const double *d_tab, const double *h_tab, const int32_t *conn32,
                        int64_t n_elem, SymbolicResult *sym, int64_t n_elem_total) {
  femb200_ctx *ctx = ga->ctx;
  using B = BT<BKIND, D>;
  using S = Stag<BKIND, D, NN>;
  constexpr int G0 = GroupOf<NN>::G;
  constexpr int G1 = (2 * G0 <= NN) ? 2 * G0 : G0, G2 = (4 * G0 <= NN && NN < 8) ? 4 * G0 : G1; // fallbacks with smaller blocks
  int gsel = getenv("FEMB200_G") ? atoi(getenv("FEMB200_G")) : G0;
  if (gsel != G0 && gsel != G1 && gsel != G2) gsel = G0;
  {
    // does any variant fit the shared-memory accumulators?  (the smallest block is the one with the most lanes per row;
    // a single row needs max_rowlen blocks whatever the layout)
    const size_t smem_min = (size_t)sym->max_rowlen * B::NDN * B::NDN * sizeof(double) * ((NN == 8 || NN == 20) ? 1 : ((32 / G2) | 1));
    if (smem_min > ctx->smem_optin) return -2; // caller falls back to global accumulators
  }
  TabLayout tl = tab_layout(sh.d, sh.nn, sh.nqp, sh.dimC);
  double *stag = nullptr;
  uint8_t *degen = nullptr;
  int rc;
  if constexpr (NN <= 4) {
    // default for elements with <= 4 nodes: the fused scalar-row kernel (srows.cu).  k_geometry runs without stores, as
    // the reference's per-element checks only; FEMB200_ROWS=noderows keeps the staged node-row kernel.
    const char *rows_env0 = getenv("FEMB200_ROWS");
    const bool want_fused = !(rows_env0 && (!strcmp(rows_env0, "noderows") || !strcmp(rows_env0, "srow")));
    if (want_fused && sym->blkcols && sh.nqp <= 9) {
      double4 *nodes4 = nullptr;
      if ((rc = scratch_alloc(ctx, &nodes4, (size_t)ga->n_nodes))) return rc;
      FEMB_CUDA_OK(ctx, cudaEventRecord(ctx->ev[9], ctx->stream));
      if (n_elem && !ga->check_done) {                           // normally done already, on the second stream (api.cu)
        k_geometry<D, NN, BKIND, TAB_CONST, false><<<grid_for(n_elem, 128), 128, 0, ctx->stream>>>(n_elem, ga->d_nodes, conn32, nullptr, nullptr, ctx->d_err, d_tab,
                                                                                                 tl.offW, tl.offN, tl.offDN, sh.nqp);
        ctx->launches++;
      }
      FEMB_CUDA_OK(ctx, cudaEventRecord(ctx->ev[10], ctx->stream));
      RowArgs a;
      a.n_nodes = ga->n_nodes; a.n_numeric = n_elem; a.n2e_ptr = sym->n2e_ptr; a.n2e_t = sym->n2e_t; a.pos16 = sym->pos16; a.blkptr = sym->blkptr;
      a.stag = nullptr; a.degen = nullptr; a.order = nullptr; a.v0 = 0; a.v1 = ga->n_nodes; a.values = sym->csr.values; a.nqp = sh.nqp; a.ninc = n_elem * NN;
      a.has_halo = n_elem < n_elem_total ? 1 : 0;
      for (int i = 0; i < kMaxDimC * kMaxDimC; ++i) a.C[i] = i < sh.dimC * sh.dimC ? h_tab[i] : 0.0;
      rc = srows_fused(ctx, sh, a, ga->d_nodes, sym->blkcols, h_tab, nullptr, 0, ga->n_nodes, sym->max_rowlen, nodes4, sym->cap_blk,
                       sym->rowcols, sym->cols_stride);
      FEMB_CUDA_OK(ctx, cudaEventRecord(ctx->ev[11], ctx->stream));
      if (rc != -2) { ga->split_numeric = true; return rc; }
    }
  }
  if (sym->speculative) return -3; // hinted sizes are only safe with the guarded fused kernel: the caller repeats the call without hints
  if constexpr (D == 3 && BKIND == FEMB200_B_XYZ && NN > 4) {
    // default for Hexa8 / Tetra10 / Hexa20 solids: element blocks + ordered row gather (keblocks.cu).  The geometry is staged
    // for one chunk of elements at a time (the chunk's records are consumed by the block kernel right behind it; computing
    // the geometry inside the block kernel was measured: 27 of 64 threads busy on long dependent chains, 1.7x slower).
    // FEMB200_ROWS=noderows keeps the row-centric kernel, which also takes over when the element-matrix staging does not fit.
    const char *rows_env0 = getenv("FEMB200_ROWS");
    const size_t ke_bytes = (size_t)n_elem * ke_doubles_per_element(NN) * sizeof(double);
    const int64_t chunk = getenv("FEMB200_KE_CHUNK") ? atoll(getenv("FEMB200_KE_CHUNK")) : (int64_t)1 << 18;
    const int64_t nchunk = n_elem < chunk ? (n_elem > 0 ? n_elem : 1) : chunk;
    const size_t stag_bytes = (size_t)nchunk * sh.nqp * StagSoa<NN>::QREC * sizeof(double);
    bool fits = (size_t)sym->max_rowlen * 9 * sizeof(double) <= ctx->smem_optin && ke_blocks_smem(NN, sh.nqp) + 1024 <= ctx->smem_optin;
    if (fits) {
      const size_t req[3] = {stag_bytes, (size_t)n_elem, ke_bytes};             // the order of the allocations below
      const size_t extra = scratch_extra_needed(ctx, req, 3);
      if (extra) {
        size_t free_b = 0, total_b = 0;
        uint64_t reserved = 0, used = 0;
        FEMB_CUDA_OK(ctx, cudaMemGetInfo(&free_b, &total_b));
        cudaMemPoolGetAttribute(ctx->pool, cudaMemPoolAttrReservedMemCurrent, &reserved);
        cudaMemPoolGetAttribute(ctx->pool, cudaMemPoolAttrUsedMemCurrent, &used);
        const double avail = (double)free_b + (double)(reserved > used ? reserved - used : 0);
        fits = (double)extra < 0.85 * avail;
      }
    }
    if (fits && !(rows_env0 && (!strcmp(rows_env0, "noderows") || !strcmp(rows_env0, "srow")))) {
      double *ke = nullptr;
      // the small arrays first: a new arena chunk is at least half the size of the previous one, so a request that FOLLOWS the
      // element-matrix staging (tens of GB) would reserve half as much again
      if ((rc = scratch_alloc(ctx, &stag, stag_bytes / sizeof(double)))) return rc;
      if ((rc = scratch_alloc(ctx, &degen, (size_t)n_elem))) return rc;
      if ((rc = scratch_alloc(ctx, &ke, ke_bytes / sizeof(double)))) return rc;
      FEMB_CUDA_OK(ctx, cudaEventRecord(ctx->ev[9], ctx->stream));
      for (int64_t c0 = 0; c0 < n_elem; c0 += nchunk) {
        const int64_t nc = n_elem - c0 < nchunk ? n_elem - c0 : nchunk;
        k_geometry<D, NN, BKIND, TAB_CONST, true, true><<<grid_for(nc, 128), 128, 0, ctx->stream>>>(nc, ga->d_nodes, conn32 + c0 * NN, stag, degen + c0, ctx->d_err,
                                                                                                  d_tab, tl.offW, tl.offN, tl.offDN, sh.nqp, c0);
        ctx->launches++;
        if ((rc = ke_elem_blocks(ctx, NN, nc, stag, sh.nqp, ke + c0 * (int64_t)ke_doubles_per_element(NN), h_tab))) return rc;
      }
      FEMB_CUDA_OK(ctx, cudaEventRecord(ctx->ev[10], ctx->stream));
      RowArgs a;
      a.n_nodes = ga->n_nodes; a.n_numeric = n_elem; a.n2e_ptr = sym->n2e_ptr; a.n2e_t = sym->n2e_t; a.pos16 = sym->pos16; a.blkptr = sym->blkptr;
      a.stag = nullptr; a.degen = degen; a.order = nullptr; a.v0 = 0; a.v1 = ga->n_nodes; a.values = sym->csr.values; a.nqp = sh.nqp; a.ninc = n_elem * NN;
      a.has_halo = n_elem < n_elem_total ? 1 : 0;
      for (int i = 0; i < kMaxDimC * kMaxDimC; ++i) a.C[i] = 0.0;
      rc = ke_rows_gather(ctx, NN, a, ke, sym->max_rowlen);
      FEMB_CUDA_OK(ctx, cudaEventRecord(ctx->ev[11], ctx->stream));
      ga->split_numeric = true;
      return rc;
    }
  }
  rc = scratch_alloc(ctx, &stag, (size_t)n_elem * sh.nqp * S::REC);
  if (rc) return rc;
  if ((rc = scratch_alloc(ctx, &degen, (size_t)n_elem))) return rc;
  FEMB_CUDA_OK(ctx, cudaEventRecord(ctx->ev[9], ctx->stream));
  if (n_elem) {
    k_geometry<D, NN, BKIND, TAB_CONST><<<grid_for(n_elem, 128), 128, 0, ctx->stream>>>(n_elem, ga->d_nodes, conn32, stag, degen, ctx->d_err, d_tab,
                                                                                      tl.offW, tl.offN, tl.offDN, sh.nqp);
    ctx->launches++;
  }
  FEMB_CUDA_OK(ctx, cudaEventRecord(ctx->ev[10], ctx->stream));
  RowArgs a;
  a.n_nodes = ga->n_nodes; a.n_numeric = n_elem; a.n2e_ptr = sym->n2e_ptr; a.n2e_t = sym->n2e_t; a.pos16 = sym->pos16; a.blkptr = sym->blkptr;
  a.stag = stag; a.degen = degen; a.order = nullptr; a.v0 = 0; a.v1 = ga->n_nodes; a.values = sym->csr.values; a.nqp = sh.nqp; a.ninc = n_elem * NN;
  a.has_halo = n_elem < n_elem_total ? 1 : 0;
  for (int i = 0; i < kMaxDimC * kMaxDimC; ++i) a.C[i] = i < sh.dimC * sh.dimC ? h_tab[i] : 0.0;
  rc = -2;
  const int64_t nn_ = a.n_nodes;
  constexpr bool PACKED = NN == 8 || NN == 20;
  const char *rows_env = getenv("FEMB200_ROWS");
  if (rows_env && !strcmp(rows_env, "srow")) {
    if (sym->order && sym->n_short > 0 && sym->n_short < nn_) {
      rc = srows_staged(ctx, sh, a, sym->order, 0, sym->n_short, sym->max_rowlen / 2);
      if (rc == 0) rc = srows_staged(ctx, sh, a, sym->order, sym->n_short, nn_, sym->max_rowlen);
    } else {
      rc = srows_staged(ctx, sh, a, nullptr, 0, nn_, sym->max_rowlen);
    }
  }
  if (rc == -2)
  if (!PACKED && sym->order && sym->n_short > 0 && sym->n_short < nn_ && gsel == G0) {
    // two launches (see k_class_flags): short rows (rowlen <= max/2) with half the accumulators per row, then the long
    // (and empty) rows.  (More lanes per row for the long class was measured: no difference.)
    rc = launch_noderows<D, NN, BKIND, G0>(ctx, a, sym, sym->order, 0, sym->n_short, sym->max_rowlen / 2);
    if (rc == 0) rc = launch_noderows<D, NN, BKIND, G0>(ctx, a, sym, sym->order, sym->n_short, nn_, sym->max_rowlen);
  }
  if (rc == -2 && gsel == G0) rc = launch_noderows<D, NN, BKIND, G0>(ctx, a, sym, nullptr, 0, nn_, sym->max_rowlen);
  if (rc == -2 && gsel <= G1) rc = launch_noderows<D, NN, BKIND, G1>(ctx, a, sym, nullptr, 0, nn_, sym->max_rowlen);
  if (rc == -2) rc = launch_noderows<D, NN, BKIND, G2>(ctx, a, sym, nullptr, 0, nn_, sym->max_rowlen);
  FEMB_CUDA_OK(ctx, cudaEventRecord(ctx->ev[11], ctx->stream));
  ga->split_numeric = true;
  return rc;
}

// ---- launch plumbing -------------------------------------------------------------------------------
template <int D, int NN, int BKIND, bool SMEM_ACC, bool TAB_CONST>
static int launch_rows_T(femb200_ctx *ctx, const NumArgs &a, int max_rowlen) {
  using B = BT<BKIND, D>;
  constexpr int NB2 = B::NDN * B::NDN;
  constexpr bool ACC_REGS = (NN * NB2 <= 36);
  constexpr int TPB = 128;
  auto kern = k_numeric_rows<D, NN, BKIND, SMEM_ACC, ACC_REGS, TAB_CONST, TPB>;
  int T = TPB;
  size_t smem = 0;
  if (SMEM_ACC) {
    const size_t per_thread = (size_t)max_rowlen * NB2 * sizeof(double);
    FEMB_CUDA_OK(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smem_optin));
    int best_threads = 0, bestT = 0;
    const int cands[4] = {128, 96, 64, 32};
    for (int c = 0; c < 4; ++c) {
      size_t need = per_thread * cands[c];
      if (need > ctx->smem_optin) continue;
      int nb = 0;
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, cands[c], need) != cudaSuccess) continue;
      if (nb * cands[c] > best_threads) { best_threads = nb * cands[c]; bestT = cands[c]; }
    }
    if (bestT == 0) return -2; // does not fit: caller falls back to global accumulators
    T = bestT;
    smem = per_thread * T;
  }
  kern<<<grid_for(a.n_nodes, T), T, smem, ctx->stream>>>(a);
  ctx->launches++;
  FEMB_CUDA_OK(ctx, cudaGetLastError());
  return 0;
}

template <int D, int NN, int BKIND>
static int launch_rows(femb200_assembler *ga, const ElemShape &sh, NumArgs &a, bool tab_const, SymbolicResult *sym, int64_t n_elem) {
  femb200_ctx *ctx = ga->ctx;
  using B = BT<BKIND, D>;
  if (sh.ndn != B::NDN || sh.dimC != B::DIMC) return FEMB200_ERR_UNSUPPORTED;
  int rc = -2;
  const bool force_global = getenv("FEMB200_FORCE_GLOBAL_ACC") != nullptr;
  const bool force_thread_rows = getenv("FEMB200_FORCE_THREAD_ROWS") != nullptr;
  if (sh.identity && !force_global && !force_thread_rows) {
    rc = tab_const ? launch_dof_T<D, NN, BKIND, true>(ga, sh, a.tab_g, a.h_tab, a.conn, n_elem, sym, a.n_elem_total)
                   : launch_dof_T<D, NN, BKIND, false>(ga, sh, a.tab_g, a.h_tab, a.conn, n_elem, sym, a.n_elem_total);
    if (rc != -2) return rc;
  }
  if (sh.identity && !force_global)
    rc = tab_const ? launch_rows_T<D, NN, BKIND, true, true>(ctx, a, sym->max_rowlen)
                   : launch_rows_T<D, NN, BKIND, true, false>(ctx, a, sym->max_rowlen);
  if (rc == -2) {
    FEMB_CUDA_OK(ctx, cudaMemsetAsync(a.values, 0, sizeof(double) * (size_t)sym->csr.nnz, ctx->stream));
    rc = tab_const ? launch_rows_T<D, NN, BKIND, false, true>(ctx, a, sym->max_rowlen)
                   : launch_rows_T<D, NN, BKIND, false, false>(ctx, a, sym->max_rowlen);
  }
  return rc;
}

template <int D, int BKIND>
static int dispatch_nn(femb200_assembler *ga, const ElemShape &sh, NumArgs &a, bool tc, SymbolicResult *sym, int64_t n_elem) {
  if (D == 3) {
    switch (sh.nn) {
    case 4: return launch_rows<D, 4, BKIND>(ga, sh, a, tc, sym, n_elem);
    case 8: return launch_rows<D, 8, BKIND>(ga, sh, a, tc, sym, n_elem);
    case 10: return launch_rows<D, 10, BKIND>(ga, sh, a, tc, sym, n_elem);
    case 20: return launch_rows<D, 20, BKIND>(ga, sh, a, tc, sym, n_elem);
    }
  } else {
    switch (sh.nn) {
    case 3: return launch_rows<D, 3, BKIND>(ga, sh, a, tc, sym, n_elem);
    case 4: return launch_rows<D, 4, BKIND>(ga, sh, a, tc, sym, n_elem);
    case 8: return launch_rows<D, 8, BKIND>(ga, sh, a, tc, sym, n_elem);
    }
  }
  return FEMB200_ERR_UNSUPPORTED;
}

bool numeric_supported(const ElemShape &sh) {
  const bool nn3 = sh.nn == 4 || sh.nn == 8 || sh.nn == 10 || sh.nn == 20;
  const bool nn2 = sh.nn == 3 || sh.nn == 4 || sh.nn == 8;
  switch (sh.bkind) {
  case FEMB200_B_XYZ: return sh.d == 3 && sh.ndn == 3 && sh.dimC == 6 && nn3;
  case FEMB200_B_PLANE: return sh.d == 2 && sh.ndn == 2 && sh.dimC == 3 && nn2;
  case FEMB200_B_AXISYM: return sh.d == 2 && sh.ndn == 2 && sh.dimC == 4 && nn2;
  case FEMB200_B_THERMAL: return sh.ndn == 1 && sh.dimC == sh.d && ((sh.d == 3 && nn3) || (sh.d == 2 && nn2));
  case FEMB200_B_THERMAL_AXISYM: return sh.d == 2 && sh.ndn == 1 && sh.dimC == 2 && nn2;
  }
  return false;
}

int numeric_phase(femb200_assembler *ga, const ElemShape &sh, const double *d_tab, const double *h_tab, int tab_doubles,
                  const int32_t *conn32, int64_t n_elem, SymbolicResult *sym, int64_t n_elem_total) {
  femb200_ctx *ctx = ga->ctx;
  if (!numeric_supported(sh)) return FEMB200_ERR_UNSUPPORTED;
  if (sym->speculative && !(sh.identity && sh.nn <= 4)) return -3; // hinted sizes need the guarded fused kernel
  const bool tab_const = tab_doubles <= kTabDoubles && !ctx->concurrent; // __constant__ tables are per device: one call at a time
  if (tab_const)
    FEMB_CUDA_OK(ctx, cudaMemcpyToSymbolAsync(c_tab, h_tab, sizeof(double) * tab_doubles, 0, cudaMemcpyHostToDevice, ctx->stream));
  TabLayout tl = tab_layout(sh.d, sh.nn, sh.nqp, sh.dimC);
  NumArgs a;
  a.n_nodes = ga->n_nodes;
  a.n_numeric = n_elem;
  a.nodes = ga->d_nodes;
  a.conn = conn32;
  a.n2e_ptr = sym->n2e_ptr;
  a.n2e_t = sym->n2e_t;
  a.pos16 = sym->pos16;
  a.blkptr = sym->blkptr;
  a.values = sym->csr.values;
  a.err = ctx->d_err;
  a.tab_g = d_tab;
  a.h_tab = h_tab;
  a.offW = tl.offW; a.offN = tl.offN; a.offDN = tl.offDN; a.nqp = sh.nqp;
  a.nu = sh.nu;
  a.n_elem_total = n_elem_total;
  for (int i = 0; i < kMaxNdn; ++i) a.dof2u[i] = sh.dof2u[i];
  switch (sh.bkind) {
  case FEMB200_B_XYZ: return dispatch_nn<3, FEMB200_B_XYZ>(ga, sh, a, tab_const, sym, n_elem);
  case FEMB200_B_PLANE: return dispatch_nn<2, FEMB200_B_PLANE>(ga, sh, a, tab_const, sym, n_elem);
  case FEMB200_B_AXISYM: return dispatch_nn<2, FEMB200_B_AXISYM>(ga, sh, a, tab_const, sym, n_elem);
  case FEMB200_B_THERMAL:
    return sh.d == 3 ? dispatch_nn<3, FEMB200_B_THERMAL>(ga, sh, a, tab_const, sym, n_elem)
                     : dispatch_nn<2, FEMB200_B_THERMAL>(ga, sh, a, tab_const, sym, n_elem);
  case FEMB200_B_THERMAL_AXISYM: return dispatch_nn<2, FEMB200_B_THERMAL_AXISYM>(ga, sh, a, tab_const, sym, n_elem);
  }
  return FEMB200_ERR_UNSUPPORTED;
}


// ---- per-element checks only (assembler_isoparametric.go:95-107), for the fused numeric path ------------------------------
// Runs k_geometry without stores on `st` (the context's second stream: beside the symbolic phase).  Tables from global memory.
template <int D, int NN, int BKIND>
static int launch_check(femb200_assembler *ga, const ElemShape &sh, const double *d_tab, const int32_t *conn32, int64_t n_elem, cudaStream_t st) {
  femb200_ctx *ctx = ga->ctx;
  using B = BT<BKIND, D>;
  if (sh.ndn != B::NDN || sh.dimC != B::DIMC) return FEMB200_ERR_UNSUPPORTED;
  TabLayout tl = tab_layout(sh.d, sh.nn, sh.nqp, sh.dimC);
  if (n_elem) {
    k_geometry<D, NN, BKIND, false, false><<<grid_for(n_elem, 128), 128, 0, st>>>(n_elem, ga->d_nodes, conn32, nullptr, nullptr, ctx->d_err, d_tab,
                                                                                 tl.offW, tl.offN, tl.offDN, sh.nqp);
    ctx->launches++;
  }
  FEMB_CUDA_OK(ctx, cudaGetLastError());
  return 0;
}

template <int D, int BKIND>
static int check_nn(femb200_assembler *ga, const ElemShape &sh, const double *d_tab, const int32_t *conn32, int64_t n_elem, cudaStream_t st) {
#define FEMB_C(NNV) return launch_check<D, NNV, BKIND>(ga, sh, d_tab, conn32, n_elem, st)
  if constexpr (D == 3) {
    switch (sh.nn) { case 4: FEMB_C(4); case 8: FEMB_C(8); case 10: FEMB_C(10); case 20: FEMB_C(20); }
  } else {
    switch (sh.nn) { case 3: FEMB_C(3); case 4: FEMB_C(4); case 8: FEMB_C(8); }
  }
#undef FEMB_C
  return FEMB200_ERR_UNSUPPORTED;
}

int check_phase(femb200_assembler *ga, const ElemShape &sh, const double *d_tab, const int32_t *conn32, int64_t n_elem, cudaStream_t st) {
  if (!numeric_supported(sh)) return FEMB200_ERR_UNSUPPORTED;
  switch (sh.bkind) {
  case FEMB200_B_XYZ: return check_nn<3, FEMB200_B_XYZ>(ga, sh, d_tab, conn32, n_elem, st);
  case FEMB200_B_PLANE: return check_nn<2, FEMB200_B_PLANE>(ga, sh, d_tab, conn32, n_elem, st);
  case FEMB200_B_AXISYM: return check_nn<2, FEMB200_B_AXISYM>(ga, sh, d_tab, conn32, n_elem, st);
  case FEMB200_B_THERMAL:
    return sh.d == 3 ? check_nn<3, FEMB200_B_THERMAL>(ga, sh, d_tab, conn32, n_elem, st) : check_nn<2, FEMB200_B_THERMAL>(ga, sh, d_tab, conn32, n_elem, st);
  case FEMB200_B_THERMAL_AXISYM: return check_nn<2, FEMB200_B_THERMAL_AXISYM>(ga, sh, d_tab, conn32, n_elem, st);
  }
  return FEMB200_ERR_UNSUPPORTED;
}

// ---- IsoparametricStrains (assembler_isoparametric.go:126-213): one thread per element -----------
template <int D, int NN, int BKIND, bool TAB_CONST>
__global__ void k_strains(int64_t n_elem, const double *__restrict__ nodes, const int32_t *__restrict__ conn,
                          const double *__restrict__ u, double *__restrict__ strains, unsigned long long *err,
                          const double *tab_g, int offW, int offN, int offDN, int nqp, int mdpn,
                          const int32_t *__restrict__ dofmap) {
  using B = BT<BKIND, D>;
  constexpr int NDN = B::NDN, NV = B::NV, DIMC = B::DIMC;
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_elem) return;
  Tab<TAB_CONST> tab{tab_g, offW, offN, offDN};
  double X[NN][D];
  int64_t nd[NN];
#pragma unroll
  for (int b = 0; b < NN; ++b) {
    nd[b] = conn[e * NN + b];
#pragma unroll
    for (int j = 0; j < D; ++j) X[b][j] = nodes[nd[b] * 3 + j];
  }
  for (int q = 0; q < nqp; ++q) {
    LuFac<D> fac;
      double detJ, rinv, scale;
    const int ek = qp_geometry<D, NN, B::RADIAL>(tab, q, X, fac, detJ, rinv, scale);
    if (ek) { atomicMin(err, pack_error(e, q, ek)); return; }
    double s[DIMC];
#pragma unroll
    for (int k = 0; k < DIMC; ++k) s[k] = 0.0;
#pragma unroll
    for (int b = 0; b < NN; ++b) {
      double vb[NV];
      node_vals<BKIND, D, NN>(tab, q, b, fac, rinv, vb);
#pragma unroll
      for (int i = 0; i < NDN; ++i) {
        const double ui = u[nd[b] * mdpn + dofmap[i]];
#pragma unroll
        for (int k = 0; k < DIMC; ++k)
          if (B::sel(k, i) >= 0) s[k] = fma(vb[B::sel(k, i)], ui, s[k]);
      }
    }
#pragma unroll
    for (int k = 0; k < DIMC; ++k) strains[(e * nqp + q) * DIMC + k] = s[k];
  }
}

template <int D, int NN, int BKIND>
static int launch_strains(femb200_ctx *ctx, bool tc, int64_t n_elem, const double *nodes, const int32_t *conn,
                          const double *u, double *strains, const double *tab_g, const TabLayout &tl, int nqp, int mdpn,
                          const int32_t *dofmap) {
  if (n_elem == 0) return 0;
  if (tc)
    k_strains<D, NN, BKIND, true><<<grid_for(n_elem, 128), 128, 0, ctx->stream>>>(n_elem, nodes, conn, u, strains, ctx->d_err, tab_g, tl.offW, tl.offN, tl.offDN, nqp, mdpn, dofmap);
  else
    k_strains<D, NN, BKIND, false><<<grid_for(n_elem, 128), 128, 0, ctx->stream>>>(n_elem, nodes, conn, u, strains, ctx->d_err, tab_g, tl.offW, tl.offN, tl.offDN, nqp, mdpn, dofmap);
  ctx->launches++;
  FEMB_CUDA_OK(ctx, cudaGetLastError());
  return 0;
}

template <int D, int BKIND>
static int strains_nn(femb200_ctx *ctx, const ElemShape &sh, bool tc, int64_t n_elem, const double *nodes, const int32_t *conn,
                      const double *u, double *strains, const double *tab_g, const TabLayout &tl, const int32_t *dofmap) {
#define FEMB_S(NNV) return launch_strains<D, NNV, BKIND>(ctx, tc, n_elem, nodes, conn, u, strains, tab_g, tl, sh.nqp, sh.mdpn, dofmap)
  if (D == 3) {
    switch (sh.nn) { case 4: FEMB_S(4); case 8: FEMB_S(8); case 10: FEMB_S(10); case 20: FEMB_S(20); }
  } else {
    switch (sh.nn) { case 3: FEMB_S(3); case 4: FEMB_S(4); case 8: FEMB_S(8); }
  }
#undef FEMB_S
  return FEMB200_ERR_UNSUPPORTED;
}

int strains_phase(femb200_assembler *ga, const ElemShape &sh, const double *d_tab, const int32_t *conn32,
                  int64_t n_elem, const int32_t *d_dofmap, const double *d_u, double *d_strains) {
  femb200_ctx *ctx = ga->ctx;
  if (!numeric_supported(sh)) return FEMB200_ERR_UNSUPPORTED;
  TabLayout tl = tab_layout(sh.d, sh.nn, sh.nqp, sh.dimC);
  // strains always read the tables from global memory (tiny kernel, keeps c_tab free for assembly)
  const bool tc = false;
  switch (sh.bkind) {
  case FEMB200_B_XYZ: return strains_nn<3, FEMB200_B_XYZ>(ctx, sh, tc, n_elem, ga->d_nodes, conn32, d_u, d_strains, d_tab, tl, d_dofmap);
  case FEMB200_B_PLANE: return strains_nn<2, FEMB200_B_PLANE>(ctx, sh, tc, n_elem, ga->d_nodes, conn32, d_u, d_strains, d_tab, tl, d_dofmap);
  case FEMB200_B_AXISYM: return strains_nn<2, FEMB200_B_AXISYM>(ctx, sh, tc, n_elem, ga->d_nodes, conn32, d_u, d_strains, d_tab, tl, d_dofmap);
  case FEMB200_B_THERMAL:
    return sh.d == 3 ? strains_nn<3, FEMB200_B_THERMAL>(ctx, sh, tc, n_elem, ga->d_nodes, conn32, d_u, d_strains, d_tab, tl, d_dofmap)
                     : strains_nn<2, FEMB200_B_THERMAL>(ctx, sh, tc, n_elem, ga->d_nodes, conn32, d_u, d_strains, d_tab, tl, d_dofmap);
  case FEMB200_B_THERMAL_AXISYM: return strains_nn<2, FEMB200_B_THERMAL_AXISYM>(ctx, sh, tc, n_elem, ga->d_nodes, conn32, d_u, d_strains, d_tab, tl, d_dofmap);
  }
  return FEMB200_ERR_UNSUPPORTED;
}

} // namespace femb200
