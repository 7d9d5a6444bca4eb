// srows.cu -- the scalar-row kernel: numeric phase + deterministic segmented reduction, one THREAD per scalar CSR row.
//
// Reference semantics: Ke += B^T C B * dJac*w*scale per quadrature point (assembler_isoparametric.go:109-112), then the
// nd^2 COO triplets of every element (assembler.go:79-92) summed by lap.Sparse.Accumulate (:121).  Here a thread owns
// scalar row (node I, dof i) of K and walks the elements incident to node I in ascending element index, so every K entry
// is a fixed-order sum: a deterministic segmented reduction whose segments are the CSR rows, no atomics, no COO stream.
//
// What this kernel changes against k_noderows (numeric.cu), and why (ncu, profiles/r1x_k_noderows_full.md: that kernel
// sits at 88 % of the L1 data pipe, 29 M of its 103 M wavefronts are shared-memory bank-conflict replays):
//  * the accumulators of a thread are PRIVATE (entry m at s_acc[m * TPB + tid]): every read-modify-write hits bank
//    tid, whatever the slot => no bank conflicts by construction, no group synchronisation, collapsed elements
//    (repeated node) need no special path because a thread's updates are sequential;
//  * the element's nodes are visited in ROTATED order (b' = 0 is the row's own node), so the diagonal block is a
//    compile-time target and lives in registers for the whole row: a quarter of the read-modify-writes of a
//    linear tetrahedron disappear;
//  * E_i = f * B_a[:, i]^T C for the thread's own i only (a lane-private NV x DIMC slice of C in registers): no
//    redundant E across the lanes of a node row;
//  * the finished rows leave through a per-warp staging buffer (own accumulators -> final layout -> coalesced stores).
//
// Front ends (how a thread gets the node values v_b[q] and f[q] = dJac*w*scale of an incident element):
//  STAGED  the records written by k_geometry (numeric.cu), one 256-bit load per node;
//  FUSED   (linear elements) the coordinates of the row's column nodes are cached in shared memory once per row and
//          the Jacobian is inverted in the kernel, its rows split across the lanes of the node row -- no staged
//          records, no geometry kernel: HBM traffic is the slot map in, the CSR values out.
#include "numeric.cuh"
#include <cstdlib>

namespace femb200 {

// lane-private slice of C: Ci[m][c] = C[k][c] for the row k of B with sel(k, i) == m (zero when there is none)
template <int BKIND, int D>
__device__ __forceinline__ void load_ci(const double *C, int i, double (&Ci)[BT<BKIND, D>::NV][BT<BKIND, D>::DIMC]) {
  using B = BT<BKIND, D>;
#pragma unroll
  for (int m = 0; m < B::NV; ++m)
#pragma unroll
    for (int c = 0; c < B::DIMC; ++c) Ci[m][c] = 0.0;
#pragma unroll
  for (int ii = 0; ii < B::NDN; ++ii)
    if (ii == i) {
#pragma unroll
      for (int k = 0; k < B::DIMC; ++k)
        if (B::sel(k, ii) >= 0) {
#pragma unroll
          for (int c = 0; c < B::DIMC; ++c) Ci[B::sel(k, ii)][c] = C[k * B::DIMC + c];
        }
    }
}

// ---- copy-out: own accumulators -> staging in the final layout -> coalesced stores -----------------------------------
// node-block row of node I = NDN scalar rows back to back: entry (i, p, j) at i*rowsz + p*NDN + j.  A pass takes as many
// whole node rows of the warp as fit in the staging buffer.  (A thread's accumulators sit in ONE bank, so they cannot be
// read by consecutive lanes; the detour through the staging buffer costs two shared-memory passes and keeps the global
// stores fully coalesced.)
template <int NDN, int TPB>
__device__ __forceinline__ void srow_copy_out(const double *acc, double *s_stage, int stage_cap, bool act, int lane, int r, int i, int rowsz,
                                              double *vbase) {
  __syncwarp();
  const int tot = NDN * rowsz;                                   // doubles of my node row
  int incl = (act && i == 0) ? tot : 0;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t2 = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t2;
  }
  const int leader = r * NDN;
  const int gend = __shfl_sync(0xffffffffu, incl, act ? leader : 0);
  const int goff = gend - tot;                                   // offset of my node row in the warp's output
  const int wtot = __shfl_sync(0xffffffffu, incl, 31);
  int base = 0;
  while (base < wtot) {
    const bool inpass = act && rowsz > 0 && goff >= base && goff + tot - base <= stage_cap;
    const unsigned pm = __ballot_sync(0xffffffffu, inpass);
    if (pm == 0) break; // a single node row larger than the staging buffer: excluded at launch
    if (inpass) {
      double *const dst = s_stage + (goff - base) + i * rowsz;
      const int sk = (rowsz & 1) ? 0 : (lane % rowsz);           // skew the write order of even-length rows across banks
      for (int m = 0; m < rowsz; ++m) {
        int mm = m + sk;
        if (mm >= rowsz) mm -= rowsz;
        dst[mm] = acc[mm * TPB];
      }
    }
    __syncwarp();
    unsigned leaders = pm;
    int pass_end = base;
    while (leaders) {
      const int l0 = __ffs(leaders) - 1;
      leaders &= ~(((1u << NDN) - 1u) << l0);                    // the NDN lanes of that node row
      const int tot_n = __shfl_sync(0xffffffffu, tot, l0);
      const int goff_n = __shfl_sync(0xffffffffu, goff, l0);
      const unsigned long long vb_n = __shfl_sync(0xffffffffu, (unsigned long long)vbase, l0);
      double *const vrow = reinterpret_cast<double *>(vb_n);
      const double *const src = s_stage + (goff_n - base);
      for (int m = lane; m < tot_n; m += 32) vrow[m] = src[m];
      pass_end = goff_n + tot_n;
    }
    __syncwarp();
    base = pass_end;
  }
}

// In-place variant for warp-private accumulators acc_w[m * 32 + lane] (the warp owns a contiguous region of cap_m * 32
// doubles): every lane pulls its whole scalar row into registers (rowsz <= RCAP), the region is then re-used as the staging
// buffer in the final layout and written out by the whole warp.  Two shared-memory passes instead of the four of
// srow_copy_out and no second buffer.  Rows are padded to an odd length in the staging so that consecutive lanes fall
// into different banks.
template <int NDN, int RCAP>
__device__ __forceinline__ void srow_copy_out_inplace(double *acc_w, bool act, int lane, int i, int rowsz, double *vrow_mine) {
  double v[RCAP];
#pragma unroll
  for (int m = 0; m < RCAP; ++m) v[m] = (m < rowsz) ? acc_w[m * 32 + lane] : 0.0;
  const int padded = act ? (rowsz | 1) : 0;
  int incl = padded;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t2 = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t2;
  }
  const int soff = incl - padded;
  __syncwarp();
#pragma unroll
  for (int m = 0; m < RCAP; ++m)
    if (m < rowsz) acc_w[soff + m] = v[m];
  __syncwarp();
  // every scalar row of the warp, two per iteration (one per half warp)
  const int half = lane >> 4, l16 = lane & 15;
#pragma unroll 1
  for (int o = 0; o < 32; o += 2) {
    const int src_lane = o + half;
    const int n = __shfl_sync(0xffffffffu, rowsz, src_lane);
    const int so = __shfl_sync(0xffffffffu, soff, src_lane);
    const unsigned long long vp = __shfl_sync(0xffffffffu, (unsigned long long)vrow_mine, src_lane);
    double *const dst = reinterpret_cast<double *>(vp);
    for (int m = l16; m < n; m += 16) dst[m] = acc_w[so + m];
  }
}

// Long rows (more than 48 entries per scalar row): 16 entries of every scalar row at a time through a small staging buffer
// (32 x 17 doubles, the dead coordinate cache): own accumulators -> staging (odd stride: conflict free) -> 128-byte
// segments of the final rows, two scalar rows per iteration.
__device__ __forceinline__ void srow_copy_out_chunked(const double *acc_w, double *stage, int lane, int rowsz, double *vrow_mine) {
  int mx = rowsz;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  const int half = lane >> 4, l16 = lane & 15;
  for (int c0 = 0; c0 < mx; c0 += 16) {
#pragma unroll
    for (int mm = 0; mm < 16; ++mm)
      if (c0 + mm < rowsz) stage[lane * 17 + mm] = acc_w[(c0 + mm) * 32 + lane];
    __syncwarp();
#pragma unroll 1
    for (int o = 0; o < 32; o += 2) {
      const int src_lane = o + half;
      const int n = __shfl_sync(0xffffffffu, rowsz, src_lane);
      const unsigned long long vp = __shfl_sync(0xffffffffu, (unsigned long long)vrow_mine, src_lane);
      if (c0 + l16 < n) reinterpret_cast<double *>(vp)[c0 + l16] = stage[src_lane * 17 + l16];
    }
    __syncwarp();
  }
}

// (bulk asynchronous copies (TMA) + mbarrier helpers: numeric.cuh; here they stage the slot-map words of the next eight
// incident elements on the opt-in FEMB200_STAGE_POS path)

struct SRowExtra {
  int cap_m;        // accumulator entries per thread (rl_cap * NDN)
  int stage_cap;    // doubles of copy-out staging per warp
};

template <int D, int NN, int BKIND, int TPB>
__global__ void __launch_bounds__(TPB)
k_srows(const RowArgs a, const SRowExtra x) {
  using B = BT<BKIND, D>;
  using S = Stag<BKIND, D, NN>;
  constexpr int NDN = B::NDN, NV = B::NV, DIMC = B::DIMC, REC = S::REC;
  constexpr int NPW = 32 / NDN;             // node rows per warp
  constexpr int NSR = NPW * NDN;            // active lanes per warp
  constexpr int NWARP = TPB / 32;
  extern __shared__ double s_mem[];
  double *const s_acc = s_mem;                                       // [cap_m][TPB]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double *const s_stage = s_mem + (size_t)x.cap_m * TPB + (size_t)warp * x.stage_cap;
  const bool act = lane < NSR;
  const int r = lane / NDN, i = lane - r * NDN;
  const int64_t vis = a.v0 + ((int64_t)blockIdx.x * NWARP + warp) * NPW + r;
  const int64_t I = (act && vis < a.v1) ? (a.order ? (int64_t)__ldg(a.order + vis) : vis) : a.n_nodes;
  int beg = 0, end = 0, rl = 0;
  int64_t b0 = 0;
  if (I < a.n_nodes) {
    beg = __ldg(a.n2e_ptr + I); end = __ldg(a.n2e_ptr + I + 1);
    b0 = __ldg(a.blkptr + I);
    rl = (int)(__ldg(a.blkptr + I + 1) - b0);
  }
  if (rl == 0) end = beg;                   // unreferenced node, or a row this rank does not assemble
  const int rowsz = rl * NDN;               // entries of my scalar row
  double *const acc = s_acc + tid;
  for (int m = 0; m < rowsz; ++m) acc[m * TPB] = 0.0;

  double Ci[NV][DIMC];
  load_ci<BKIND, D>(a.C, i, Ci);
  double diag[NDN];
#pragma unroll
  for (int j = 0; j < NDN; ++j) diag[j] = 0.0;
  int pdiag = 0;
  const int64_t recsz = (int64_t)a.nqp * REC;

  // incidence lists are ascending in element index and pattern-only (halo) elements come last
  uint32_t tn = beg < end ? __ldg(a.n2e_t + beg) : 0u;
  for (int k = beg; k < end; ++k) {
    const uint32_t t = tn;
    const int64_t e = t / NN;
    if (e >= a.n_numeric) break;
    const int la = (int)(t - (uint32_t)e * NN);
    if (k + 1 < end) {
      tn = __ldg(a.n2e_t + k + 1);
      prefetch_l2(a.stag + (int64_t)(tn / NN) * recsz);
    }
    // slot positions, rotated: pr[b'] = slot of node (la + b') mod NN
    int pr[NN];
    if constexpr (NN == 4) {
      const uint2 w = __ldg(reinterpret_cast<const uint2 *>(a.pos16 + (int64_t)k * NN));
      const unsigned long long w64 = ((unsigned long long)w.y << 32) | w.x;
#pragma unroll
      for (int bp = 0; bp < NN; ++bp) pr[bp] = (int)((w64 >> (16 * ((la + bp) & 3))) & kPosMask);
    } else {
#pragma unroll
      for (int bp = 0; bp < NN; ++bp) {
        int b = la + bp;
        if (b >= NN) b -= NN;
        pr[bp] = (int)(__ldg(a.pos16 + (int64_t)k * NN + b) & pos_mask(NN));
      }
    }
    double blk[NN][NDN];
#pragma unroll
    for (int bp = 0; bp < NN; ++bp)
#pragma unroll
      for (int j = 0; j < NDN; ++j) blk[bp][j] = 0.0;
    const double *rec = a.stag + e * recsz;
    for (int q = 0; q < a.nqp; ++q, rec += REC) {
      double va[4], f;
      if constexpr (S::FIN) {
        ld256(rec + 4 * la, va);
        f = va[3];
      } else {
        const double2 w = __ldg(reinterpret_cast<const double2 *>(rec) + la);
        va[0] = w.x; va[1] = w.y; va[2] = 0.0; va[3] = 0.0;
        f = __ldg(rec + S::FOFF);
      }
      double E[DIMC];
#pragma unroll
      for (int c = 0; c < DIMC; ++c) E[c] = 0.0;
#pragma unroll
      for (int m = 0; m < NV; ++m) {
        const double u = va[m] * f;
#pragma unroll
        for (int c = 0; c < DIMC; ++c) E[c] = fma(u, Ci[m][c], E[c]);
      }
#pragma unroll
      for (int bp = 0; bp < NN; ++bp) {
        int b = la + bp;
        if (b >= NN) b -= NN;
        double vb[4];
        if constexpr (S::FIN) {
          if (bp == 0) { vb[0] = va[0]; vb[1] = va[1]; vb[2] = va[2]; }
          else ld256(rec + 4 * b, vb);
        } else {
          if (bp == 0) { vb[0] = va[0]; vb[1] = va[1]; }
          else {
            const double2 w = __ldg(reinterpret_cast<const double2 *>(rec) + b);
            vb[0] = w.x; vb[1] = w.y;
          }
        }
#pragma unroll
        for (int j = 0; j < NDN; ++j) {
          double s = blk[bp][j];
#pragma unroll
          for (int kk = 0; kk < DIMC; ++kk)
            if (B::sel(kk, j) >= 0) s = fma(E[kk], vb[B::sel(kk, j)], s);
          blk[bp][j] = s;
        }
      }
    }
    // ordered accumulation: diagonal block in registers, the others into the thread's private accumulators
#pragma unroll
    for (int j = 0; j < NDN; ++j) diag[j] += blk[0][j];
    pdiag = pr[0];
#pragma unroll
    for (int bp = 1; bp < NN; ++bp) {
      double *const s = acc + (size_t)(pr[bp] * NDN) * TPB;
#pragma unroll
      for (int j = 0; j < NDN; ++j) s[j * TPB] += blk[bp][j];
    }
  }
  if (rowsz) {
#pragma unroll
    for (int j = 0; j < NDN; ++j) acc[(pdiag * NDN + j) * TPB] += diag[j];
  }

  srow_copy_out<NDN, TPB>(acc, s_stage, x.stage_cap, act, lane, r, i, rowsz, a.values + (int64_t)NDN * NDN * b0);
}

template <int D, int NN, int BKIND>
static int launch_srows_T(femb200_ctx *ctx, const RowArgs &a0, const int32_t *order, int64_t v0, int64_t v1, int64_t rowlen_cap) {
  using B = BT<BKIND, D>;
  constexpr int TPB = 128, NPW = 32 / B::NDN, NWARP = TPB / 32;
  SRowExtra x;
  x.cap_m = (int)rowlen_cap * B::NDN;
  const int node_tot = (int)rowlen_cap * B::NDN * B::NDN;
  // staging: at least one node row; two by default (more passes cost little, shared memory costs occupancy)
  int stage = 2 * node_tot;
  if (getenv("FEMB200_SROW_STAGE")) stage = atoi(getenv("FEMB200_SROW_STAGE")) * node_tot;
  if (stage < node_tot) stage = node_tot;
  x.stage_cap = (stage + 1) & ~1;
  const size_t smem = ((size_t)x.cap_m * TPB + (size_t)NWARP * x.stage_cap) * sizeof(double);
  if (smem > ctx->smem_optin) return -2;
  auto kern = k_srows<D, NN, BKIND, TPB>;
  FEMB_CUDA_OK(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smem_optin));
  RowArgs a = a0;
  a.order = order; a.v0 = v0; a.v1 = v1;
  if (v1 > v0) {
    kern<<<grid_for(v1 - v0, NWARP * NPW), TPB, smem, ctx->stream>>>(a, x);
    ctx->launches++;
  }
  FEMB_CUDA_OK(ctx, cudaGetLastError());
  return 0;
}

template <int D, int BKIND>
static int srows_nn(femb200_ctx *ctx, int nn, const RowArgs &a, const int32_t *order, int64_t v0, int64_t v1, int64_t cap) {
  if constexpr (D == 3) {
    switch (nn) {
    case 4: return launch_srows_T<D, 4, BKIND>(ctx, a, order, v0, v1, cap);
    case 8: return launch_srows_T<D, 8, BKIND>(ctx, a, order, v0, v1, cap);
    case 10: return launch_srows_T<D, 10, BKIND>(ctx, a, order, v0, v1, cap);
    }
  } else {
    switch (nn) {
    case 3: return launch_srows_T<D, 3, BKIND>(ctx, a, order, v0, v1, cap);
    case 4: return launch_srows_T<D, 4, BKIND>(ctx, a, order, v0, v1, cap);
    case 8: return launch_srows_T<D, 8, BKIND>(ctx, a, order, v0, v1, cap);
    }
  }
  return -2;
}

// =====================================================================================================================
// FUSED front end (elements with <= 4 nodes: Tetra4, Triangle3, Quad4): no staged records.
//   * once per node row, the coordinates of the row's column nodes (its CSR block columns) are cached in shared memory
//     -- the slot map pos16 then doubles as the gather index of every incident element's coordinates;
//   * per incident element and quadrature point the thread recomputes jac = dN*X (:93), its inverse and determinant,
//     dNxy = jac^-1 dN (:100) and the constituter's node values (:104) in registers, from tables that live in
//     kernel-parameter space (constant bank operands), then forms its row of B_a^T C B_b * dJac*w*scale (:109-112).
// The reference's error checks (:95-107) are not repeated here: k_check (numeric.cu) evaluates them once per element
// with the reference's own determinant algorithm, in element order, before this kernel's result is committed.
// HBM traffic of the launch: incidence list + slot map + block columns in, CSR values out (coordinates stay in L2).
// =====================================================================================================================
constexpr int kFusedMaxQp = 9;
struct FusedTab {                       // [q][..] tables of the call (assembler_isoparametric.go:66-71)
  double w[kFusedMaxQp];
  double N[kFusedMaxQp * 4];            // [q][b]
  double dN[kFusedMaxQp * 3 * 4];       // [q][i][b]
};
struct FusedIn {
  const double4 *nodes4;                // [n_nodes] (x, y, z, 0)
  const int32_t *blkcols;               // ascending column nodes of row I at blkcols[blkptr[I] ...] (re-assembly), or
  const int32_t *rowcols;               // at rowcols[cols_stride * n2e_ptr[I] ...] (first assembly: the pattern kernel's lists)
  int cols_stride;
  int rl_cap;
  int64_t cap_blk;                      // capacity (node blocks) of values / blkcols: guards a call sized from a hint
  unsigned long long *guard;            // set when a row does not fit the hinted sizes (the call is then repeated)
};

// a*b - c*d with both products rounded BEFORE the subtraction (never contracted into an FMA): when the two products are
// equal -- the symmetric cancellations behind the structural zeros of K on axis-aligned or axisymmetric meshes -- the
// result is exactly 0, as in the reference's unfused arithmetic; fma(a, b, -(c*d)) would leave the rounding residual of c*d.
__device__ __forceinline__ double diff_of_products(double a, double b, double c, double d) { return __dsub_rn(__dmul_rn(a, b), __dmul_rn(c, d)); }

template <int D> __device__ __forceinline__ double inv_lean(const double (&J)[D][D], double (&inv)[D][D]);
template <> __device__ __forceinline__ double inv_lean<2>(const double (&J)[2][2], double (&inv)[2][2]) {
  const double det = diff_of_products(J[0][0], J[1][1], J[0][1], J[1][0]);
  const double r = 1.0 / det;
  inv[0][0] = J[1][1] * r; inv[0][1] = -J[0][1] * r;
  inv[1][0] = -J[1][0] * r; inv[1][1] = J[0][0] * r;
  return det;
}
template <> __device__ __forceinline__ double inv_lean<3>(const double (&J)[3][3], double (&inv)[3][3]) {
  // columns of the inverse = cross products of the rows of J, over det = r0 . (r1 x r2)
  double c[3][3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const int k1 = (k + 1) % 3, k2 = (k + 2) % 3;
    c[k][0] = diff_of_products(J[k1][1], J[k2][2], J[k1][2], J[k2][1]);
    c[k][1] = diff_of_products(J[k1][2], J[k2][0], J[k1][0], J[k2][2]);
    c[k][2] = diff_of_products(J[k1][0], J[k2][1], J[k1][1], J[k2][0]);
  }
  const double det = fma(J[0][0], c[0][0], fma(J[0][1], c[0][1], J[0][2] * c[0][2]));
  const double r = 1.0 / det;
#pragma unroll
  for (int k = 0; k < 3; ++k)
#pragma unroll
    for (int j = 0; j < 3; ++j) inv[j][k] = c[k][j] * r;
  return det;
}

// SIMPLEX: the dN table is the linear simplex one (rows e_k - e_0, one quadrature point), detected on the host from the
// table values.  Then jac rows are edge vectors, the columns of jac^-1 are cross products of edges over the determinant and
// dNxy follows without a matrix product; the element's nodes are visited in ROTATED order (b' = 0 is the row's own
// node), which makes the diagonal block a compile-time target: it stays in registers for the whole row and only NN-1
// blocks per element go through shared memory.  (A cyclic shift of 4 vertices is an odd permutation: the determinant
// changes sign with the parity of the shift, the gradients do not.)
template <int D, int NN, int BKIND, int NQP, bool SIMPLEX, int TPB, bool STAGE_POS>
__global__ void __launch_bounds__(TPB)
k_srows_fused(const RowArgs a, const SRowExtra x, const FusedIn fi, const FusedTab tb) {
  using B = BT<BKIND, D>;
  constexpr int NDN = B::NDN, NV = B::NV, DIMC = B::DIMC;
  constexpr int NPW = 32 / NDN, NSR = NPW * NDN, NWARP = TPB / 32;
  static_assert(NN <= 4, "fused front end: tables sized for <= 4 nodes");
  static_assert(!SIMPLEX || (NN == D + 1 && NQP == 1 && !B::RADIAL && NV == D), "simplex path: linear simplex, one point, B = gradients");
  extern __shared__ double s_mem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // per warp: accumulators acc_w[m * 32 + lane] (bank = lane: conflict free whatever the slot), then the coordinate cache
  double *const acc_w = s_mem + (size_t)warp * ((size_t)x.cap_m * 32 + x.stage_cap + (STAGE_POS ? 2 * NPW * 8 + 2 : 0));
  double *const s_stage = acc_w + (size_t)x.cap_m * 32;
  const bool act = lane < NSR;
  const int r = lane / NDN, i = lane - r * NDN;
  const int64_t vis = a.v0 + ((int64_t)blockIdx.x * NWARP + warp) * NPW + r;
  const int64_t I = (act && vis < a.v1) ? (a.order ? (int64_t)__ldg(a.order + vis) : vis) : a.n_nodes;
  int beg = 0, end = 0, rl = 0;
  int64_t b0 = 0;
  if (I < a.n_nodes) {
    beg = __ldg(a.n2e_ptr + I); end = __ldg(a.n2e_ptr + I + 1);
    b0 = __ldg(a.blkptr + I);
    rl = (int)(__ldg(a.blkptr + I + 1) - b0);
  }
  if (rl > fi.rl_cap || b0 + rl > fi.cap_blk) {              // sizes came from a hint and this row does not fit
    if (i == 0) atomicOr(fi.guard, 1ull);
    rl = 0;
  }
  if (rl == 0) end = beg;
  const int rowsz = rl * NDN;
  double *const acc = acc_w + lane;
  constexpr int AS = 32;                                         // accumulator stride
  for (int m = 0; m < rowsz; ++m) acc[m * AS] = 0.0;

  // coordinate cache of my node row: xs[p][0..D) = coordinates of column node p (the NDN lanes of the row share the
  // gathers); the simplex path caches the edge vectors X[p] - X[I] instead, which is all it ever needs
  double *const xs = s_stage + (size_t)r * fi.rl_cap * D;
  if (rl) {
    double X0[4] = {0.0, 0.0, 0.0, 0.0};
    if constexpr (SIMPLEX) ld256(reinterpret_cast<const double *>(fi.nodes4 + I), X0);
    const int32_t *cols = fi.rowcols ? fi.rowcols + (int64_t)fi.cols_stride * beg : fi.blkcols + b0;
    for (int p = i; p < rl; p += NDN) {
      double c4[4];
      ld256(reinterpret_cast<const double *>(fi.nodes4 + __ldg(cols + p)), c4);
#pragma unroll
      for (int j = 0; j < D; ++j) xs[p * D + j] = c4[j] - X0[j];
    }
  }
  __syncwarp();

  double Ci[NV][DIMC];
  load_ci<BKIND, D>(a.C, i, Ci);
  const int nqp = NQP > 0 ? NQP : a.nqp;
  double diag[NDN];
#pragma unroll
  for (int j = 0; j < NDN; ++j) diag[j] = 0.0;
  int pdiag = 0;

  // bookkeeping runs one element ahead of the arithmetic: (t, slot positions) of incidence k+1 are in flight during k
  auto load_pos = [&](int k, unsigned long long &w64) {
    if constexpr (NN == 4) {
      const uint2 w = __ldg(reinterpret_cast<const uint2 *>(a.pos16 + (int64_t)k * NN));
      w64 = ((unsigned long long)w.y << 32) | w.x;
    } else {
      w64 = 0;
#pragma unroll
      for (int b = 0; b < NN; ++b) w64 |= (unsigned long long)__ldg(a.pos16 + (int64_t)k * NN + b) << (16 * b);
    }
  };
  // pattern-only (halo) elements sit at the end of every incidence list: find the end of the numeric part once
  if (a.has_halo) {
    while (end > beg && (int64_t)(__ldg(a.n2e_t + end - 1) / NN) >= a.n_numeric) --end;
  }
  // one incident element of my row, given its slot-map word
  auto element = [&](const unsigned long long w64) {
    const int la = (int)((w64 >> kPosLaShift) & 3u);
    if constexpr (SIMPLEX) {
      // rotated order: node b' of this visit = element node (la + b') mod NN; b' = 0 is node I itself
      int pr[NN];
#pragma unroll
      for (int bp = 0; bp < NN; ++bp) {
        int b = la + bp;
        if (b >= NN) b -= NN;
        pr[bp] = (int)((w64 >> (16 * b)) & kPosMask);
      }
      double ed[D][D];                                           // edge vectors e_k = X[k] - X[0], k = 1..D (cached)
#pragma unroll
      for (int kk = 0; kk < D; ++kk) {
        const double *xp = xs + pr[kk + 1] * D;
#pragma unroll
        for (int j = 0; j < D; ++j) ed[kk][j] = xp[j];
      }
      // c[b'] = det * grad N_b' (unscaled): cross products of edges (3-D) / rotated edges (2-D); c[0] = -(sum of the others)
      double c[NN][D], det;
      if constexpr (D == 3) {
#pragma unroll
        for (int kk = 0; kk < 3; ++kk) {
          const int k1 = (kk + 1) % 3, k2 = (kk + 2) % 3;
          c[kk + 1][0] = diff_of_products(ed[k1][1], ed[k2][2], ed[k1][2], ed[k2][1]);
          c[kk + 1][1] = diff_of_products(ed[k1][2], ed[k2][0], ed[k1][0], ed[k2][2]);
          c[kk + 1][2] = diff_of_products(ed[k1][0], ed[k2][1], ed[k1][1], ed[k2][0]);
        }
        det = fma(ed[0][0], c[1][0], fma(ed[0][1], c[1][1], ed[0][2] * c[1][2]));
      } else {
        det = diff_of_products(ed[0][0], ed[1][1], ed[0][1], ed[1][0]);
        c[1][0] = ed[1][1]; c[1][1] = -ed[1][0];
        c[2][0] = -ed[0][1]; c[2][1] = ed[0][0];
      }
      const double rd = 1.0 / det;                               // off the critical path: only the final scaling of E needs it
#pragma unroll
      for (int j = 0; j < D; ++j) {
        double sg = c[1][j];
#pragma unroll
        for (int kk = 2; kk < NN; ++kk) sg += c[kk][j];
        c[0][j] = -sg;
      }
      // u = f * grad N_a = (det_nat * w) * c[0] / det = +-w * c[0]: a cyclic shift of 4 vertices by an odd count flips the
      // orientation (det = -det_nat), a shift of 3 vertices never does
      double ws = tb.w[0];                                       // (dJac*w)*scale with scale = 1 (:111)
      if (D == 3 && (la & 1)) ws = -ws;
      double E[DIMC];
#pragma unroll
      for (int cc = 0; cc < DIMC; ++cc) E[cc] = 0.0;
#pragma unroll
      for (int m = 0; m < NV; ++m) {
        const double u = c[0][m] * ws;
#pragma unroll
        for (int cc = 0; cc < DIMC; ++cc) E[cc] = fma(u, Ci[m][cc], E[cc]);
      }
#pragma unroll
      for (int cc = 0; cc < DIMC; ++cc) E[cc] *= rd;             // E = f * B_a[:, i]^T C / det: the blocks below use the unscaled c
#pragma unroll
      for (int bp = 0; bp < NN; ++bp) {
        double bl[NDN];
#pragma unroll
        for (int j = 0; j < NDN; ++j) {
          double sb = 0.0;
#pragma unroll
          for (int kk = 0; kk < DIMC; ++kk)
            if (B::sel(kk, j) >= 0) sb = fma(E[kk], c[bp][B::sel(kk, j)], sb);
          bl[j] = sb;
        }
        if (bp == 0) {
#pragma unroll
          for (int j = 0; j < NDN; ++j) diag[j] += bl[j];
        } else {
          double *const sp = acc + (pr[bp] * NDN) * AS;
#pragma unroll
          for (int j = 0; j < NDN; ++j) sp[j * AS] += bl[j];
        }
      }
      pdiag = pr[0];
    } else {
      int pos[NN];
#pragma unroll
      for (int b = 0; b < NN; ++b) pos[b] = (int)((w64 >> (16 * b)) & kPosMask);
      double X[NN][D];
#pragma unroll
      for (int b = 0; b < NN; ++b) {
        const double *xp = xs + pos[b] * D;
#pragma unroll
        for (int j = 0; j < D; ++j) X[b][j] = xp[j];
      }
      double blk[NN][NDN];
#pragma unroll
      for (int b = 0; b < NN; ++b)
#pragma unroll
        for (int j = 0; j < NDN; ++j) blk[b][j] = 0.0;
#pragma unroll
      for (int q = 0; q < (NQP > 0 ? NQP : kFusedMaxQp); ++q) {
        if (NQP == 0 && q >= nqp) break;
        double J[D][D], inv[D][D];
#pragma unroll
        for (int ii = 0; ii < D; ++ii)
#pragma unroll
          for (int j = 0; j < D; ++j) {
            // jac.Mul(dN, X) (:93) exactly as the reference rounds it: node-ascending, multiply and add unfused.  With
            // coordinates far from the origin (|X| / h ~ 2000 in config 2) the entries of J carry a relative rounding
            // error of ~|X|/h * eps; only the SAME roundings keep K within 1e-12 of the reference there.
            double sj = 0.0;
#pragma unroll
            for (int b = 0; b < NN; ++b) sj = __dadd_rn(sj, __dmul_rn(tb.dN[(q * D + ii) * NN + b], X[b][j]));
            J[ii][j] = sj;
          }
        const double detJ = inv_lean<D>(J, inv);
        double scale = 1.0, rinv = 0.0;
        if constexpr (B::RADIAL) {
          double rr = 0.0;
#pragma unroll
          for (int b = 0; b < NN; ++b) rr = fma(X[b][0], tb.N[q * NN + b], rr);
          rinv = 1.0 / rr;
          scale = rr;
        }
        const double f = detJ * tb.w[q] * scale; // (dJac*w)*scale, left to right (:111)
        double v[NN][NV];
#pragma unroll
        for (int b = 0; b < NN; ++b) {
          double g[D];
#pragma unroll
          for (int ii = 0; ii < D; ++ii) {
            double sg = 0.0;
#pragma unroll
            for (int j = 0; j < D; ++j) sg = fma(inv[ii][j], tb.dN[(q * D + j) * NN + b], sg);
            g[ii] = sg;
          }
          if constexpr (BKIND == FEMB200_B_AXISYM) {
            v[b][0] = g[0]; v[b][1] = g[1]; v[b][2] = tb.N[q * NN + b] * rinv;
          } else if constexpr (BKIND == FEMB200_B_THERMAL_AXISYM) {
            v[b][0] = tb.N[q * NN + b] * rinv + g[0]; v[b][1] = g[1];
          } else {
#pragma unroll
            for (int ii = 0; ii < D; ++ii) v[b][ii] = g[ii];
          }
        }
        double E[DIMC];
#pragma unroll
        for (int c = 0; c < DIMC; ++c) E[c] = 0.0;
#pragma unroll
        for (int m = 0; m < NV; ++m) {
          double vam = v[0][m];
#pragma unroll
          for (int b = 1; b < NN; ++b) vam = (la == b) ? v[b][m] : vam;
          const double u = vam * f;
#pragma unroll
          for (int c = 0; c < DIMC; ++c) E[c] = fma(u, Ci[m][c], E[c]);
        }
#pragma unroll
        for (int b = 0; b < NN; ++b)
#pragma unroll
          for (int j = 0; j < NDN; ++j) {
            double sb = blk[b][j];
#pragma unroll
            for (int kk = 0; kk < DIMC; ++kk)
              if (B::sel(kk, j) >= 0) sb = fma(E[kk], v[b][B::sel(kk, j)], sb);
            blk[b][j] = sb;
          }
      }
#pragma unroll
      for (int b = 0; b < NN; ++b) {
        double *const sp = acc + (pos[b] * NDN) * AS;
#pragma unroll
        for (int j = 0; j < NDN; ++j) sp[j * AS] += blk[b][j];
      }
    }
  };
  if constexpr (STAGE_POS) {
    // Slot-map words through shared memory: the words of a row are contiguous (8 bytes per incident element), so the next
    // EIGHT steps of every row of the warp arrive as one 64-byte bulk asynchronous copy per row (cp.async.bulk, completion
    // on an mbarrier), double buffered.  The step loop is warp-uniform; a row whose list starts at an odd incidence begins
    // one slot into its first (16-byte aligned) chunk.
    static_assert(!STAGE_POS || NN == 4, "staged slot words: 8 bytes per incidence");
    uint64_t *const posbuf = reinterpret_cast<uint64_t *>(s_stage + x.stage_cap);      // [2][NPW][8]
    uint64_t *const bars = posbuf + 2 * NPW * 8;                                        // [2]
    const int off = beg & 1;
    const int nsteps = (end > beg) ? (end - beg) + off : 0;
    int maxs = nsteps;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) maxs = max(maxs, __shfl_xor_sync(0xffffffffu, maxs, o));
    const int nchunks = (maxs + 7) >> 3;
    if (lane == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    const uint16_t *const src0 = a.pos16 + (int64_t)(beg - off) * NN;
    auto issue = [&](int c) {
      const bool prod = act && i == 0 && 8 * c < nsteps;
      const unsigned np = __popc(__ballot_sync(0xffffffffu, prod));
      if (lane == 0) mbar_expect_tx(&bars[c & 1], np * 64u);
      __syncwarp();
      if (prod) bulk_g2s(posbuf + ((c & 1) * NPW + r) * 8, src0 + (int64_t)c * 8 * NN, 64u, &bars[c & 1]);
    };
    if (nchunks > 0) issue(0);
    for (int c = 0; c < nchunks; ++c) {
      mbar_wait(&bars[c & 1], (uint32_t)((c >> 1) & 1));
      __syncwarp();                                               // every lane is done with chunk c-1: its buffer can be refilled
      if (c + 1 < nchunks) issue(c + 1);
      const uint64_t *const mine = posbuf + ((c & 1) * NPW + (act ? r : 0)) * 8;
#pragma unroll 1
      for (int j = 0; j < 8; ++j) {
        const int sp = 8 * c + j;
        if (sp >= off && sp < nsteps) element(mine[j]);
      }
    }
  } else {
    unsigned long long wn = 0ull;
    if (beg < end) load_pos(beg, wn);
    for (int k = beg; k < end; ++k) {
      const unsigned long long w64 = wn;
      if (k + 1 < end) load_pos(k + 1, wn);
      element(w64);
    }
  }
  if constexpr (SIMPLEX) {
    if (rowsz) {
#pragma unroll
      for (int j = 0; j < NDN; ++j) acc[(pdiag * NDN + j) * AS] += diag[j];
    }
  }
  __syncwarp();
  double *const vrow_mine = a.values + (int64_t)NDN * NDN * b0 + (int64_t)i * rowsz;
  if (x.cap_m <= 24) srow_copy_out_inplace<NDN, 24>(acc_w, act, lane, i, rowsz, vrow_mine);
  else if (x.cap_m <= 48) srow_copy_out_inplace<NDN, 48>(acc_w, act, lane, i, rowsz, vrow_mine);
  else srow_copy_out_chunked(acc_w, s_stage, lane, rowsz, vrow_mine);
}

__global__ void k_pad_nodes(int64_t n, const double *__restrict__ nodes, double4 *__restrict__ out) {
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v < n) out[v] = make_double4(nodes[3 * v], nodes[3 * v + 1], nodes[3 * v + 2], 0.0);
}

template <int D, int NN, int BKIND, int TPB>
static int launch_fused_T(femb200_ctx *ctx, const ElemShape &sh, const RowArgs &a0, const FusedIn &fi0, const double *h_tab,
                          const int32_t *order, int64_t v0, int64_t v1, int64_t rowlen_cap) {
  using B = BT<BKIND, D>;
  constexpr int NPW = 32 / B::NDN, NWARP = TPB / 32;
  if (sh.nqp > kFusedMaxQp) return -2;
  SRowExtra x;
  x.cap_m = (int)rowlen_cap * B::NDN;
  const int node_tot = (int)rowlen_cap * B::NDN * B::NDN;
  (void)node_tot;
  int cache = NPW * (int)rowlen_cap * D;                         // coordinate cache (doubles) behind the warp's accumulators
  if (cache < 32) cache = 32;                                    // the in-place copy-out pads every scalar row to an odd length
  if (x.cap_m > 48 && cache < 32 * 17) cache = 32 * 17;          // long rows: chunked copy-out staging
  x.stage_cap = (cache + 1) & ~1;
  // measured (config 1 / config 2): 0.263 vs 0.230 ms and 1.477 vs 1.400 ms WITH the staging -- the warp-uniform chunk loop and
  // the mbarrier waits cost more than the 8 wavefronts per step the bulk copies save; opt-in for experiments only
  const bool stage_pos = NN == 4 && getenv("FEMB200_STAGE_POS") != nullptr;
  const size_t pos_doubles = stage_pos ? (2 * NPW * 8 + 2) : 0;   // staged slot words + two mbarriers per warp
  const size_t smem = (size_t)NWARP * ((size_t)x.cap_m * 32 + x.stage_cap + pos_doubles) * sizeof(double);
  if (smem > ctx->smem_optin) return -2;
  FusedTab tb;
  TabLayout tl = tab_layout(sh.d, sh.nn, sh.nqp, sh.dimC);
  for (int q = 0; q < sh.nqp; ++q) {
    tb.w[q] = h_tab[tl.offW + q];
    for (int b = 0; b < NN; ++b) tb.N[q * NN + b] = h_tab[tl.offN + q * NN + b];
    for (int ii = 0; ii < D; ++ii)
      for (int b = 0; b < NN; ++b) tb.dN[(q * D + ii) * NN + b] = h_tab[tl.offDN + (q * D + ii) * NN + b];
  }
  // linear simplex table?  dN[i][0] = -1, dN[i][i+1] = +1, zero elsewhere (tetra4.go:33-37, triangles.go:38-41), one point
  bool simplex = (NN == D + 1) && sh.nqp == 1 && !getenv("FEMB200_NO_SIMPLEX");
  for (int ii = 0; ii < D && simplex; ++ii)
    for (int b = 0; b < NN; ++b) {
      const double want = b == 0 ? -1.0 : (b == ii + 1 ? 1.0 : 0.0);
      if (tb.dN[ii * NN + b] != want) simplex = false;
    }
  RowArgs a = a0;
  a.order = order; a.v0 = v0; a.v1 = v1;
  FusedIn fi = fi0;
  fi.rl_cap = (int)rowlen_cap;
  const unsigned grid = grid_for(v1 - v0, NWARP * NPW);
  auto go = [&](auto kern) -> int {
    FEMB_CUDA_OK(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smem_optin));
    if (v1 > v0) {
      kern<<<grid, TPB, smem, ctx->stream>>>(a, x, fi, tb);
      ctx->launches++;
    }
    FEMB_CUDA_OK(ctx, cudaGetLastError());
    return 0;
  };
  if constexpr (NN == 4) {
    if (stage_pos) {
      if constexpr (NN == D + 1 && !B::RADIAL && B::NV == D) {
        if (simplex) return go(k_srows_fused<D, NN, BKIND, 1, true, TPB, true>);
      }
      if (sh.nqp == 1) return go(k_srows_fused<D, NN, BKIND, 1, false, TPB, true>);
      if (sh.nqp == 4) return go(k_srows_fused<D, NN, BKIND, 4, false, TPB, true>);
      return go(k_srows_fused<D, NN, BKIND, 0, false, TPB, true>);
    }
  }
  if constexpr (NN == D + 1 && !B::RADIAL && B::NV == D) {
    if (simplex) return go(k_srows_fused<D, NN, BKIND, 1, true, TPB, false>);
  }
  if (sh.nqp == 1) return go(k_srows_fused<D, NN, BKIND, 1, false, TPB, false>);
  if (sh.nqp == 4) return go(k_srows_fused<D, NN, BKIND, 4, false, TPB, false>);
  return go(k_srows_fused<D, NN, BKIND, 0, false, TPB, false>);
}

// fused front end for elements with <= 4 nodes; returns -2 when it does not apply
int srows_fused(femb200_ctx *ctx, const ElemShape &sh, const RowArgs &a, const double *d_nodes, const int32_t *blkcols,
                const double *h_tab, const int32_t *order, int64_t v0, int64_t v1, int64_t rowlen_cap, double4 *nodes4, int64_t cap_blk,
                const int32_t *rowcols, int cols_stride) {
  if (sh.nn > 4) return -2;
  k_pad_nodes<<<grid_for(a.n_nodes, 256), 256, 0, ctx->stream>>>(a.n_nodes, d_nodes, nodes4);
  ctx->launches++;
  FusedIn fi;
  fi.nodes4 = nodes4; fi.blkcols = blkcols; fi.rowcols = rowcols; fi.cols_stride = cols_stride; fi.rl_cap = 0;
  fi.cap_blk = cap_blk; fi.guard = (unsigned long long *)(ctx->d_scalars + 7);
  const int tpb = getenv("FEMB200_SROW_TPB") ? atoi(getenv("FEMB200_SROW_TPB")) : 32; // one warp per block: finest shared-memory granularity (measured 2-3 % over 64)
#define FEMB_F(Dv, NNv, BK)                                                                                          \
  return tpb == 64    ? launch_fused_T<Dv, NNv, BK, 64>(ctx, sh, a, fi, h_tab, order, v0, v1, rowlen_cap)            \
         : tpb == 128 ? launch_fused_T<Dv, NNv, BK, 128>(ctx, sh, a, fi, h_tab, order, v0, v1, rowlen_cap)           \
                      : launch_fused_T<Dv, NNv, BK, 32>(ctx, sh, a, fi, h_tab, order, v0, v1, rowlen_cap)
  switch (sh.bkind) {
  case FEMB200_B_XYZ: if (sh.nn == 4) FEMB_F(3, 4, FEMB200_B_XYZ); break;
  case FEMB200_B_PLANE: if (sh.nn == 4) FEMB_F(2, 4, FEMB200_B_PLANE); if (sh.nn == 3) FEMB_F(2, 3, FEMB200_B_PLANE); break;
  case FEMB200_B_AXISYM: if (sh.nn == 4) FEMB_F(2, 4, FEMB200_B_AXISYM); if (sh.nn == 3) FEMB_F(2, 3, FEMB200_B_AXISYM); break;
  case FEMB200_B_THERMAL:
    if (sh.d == 3 && sh.nn == 4) FEMB_F(3, 4, FEMB200_B_THERMAL);
    if (sh.d == 2 && sh.nn == 4) FEMB_F(2, 4, FEMB200_B_THERMAL);
    if (sh.d == 2 && sh.nn == 3) FEMB_F(2, 3, FEMB200_B_THERMAL);
    break;
  case FEMB200_B_THERMAL_AXISYM: if (sh.nn == 4) FEMB_F(2, 4, FEMB200_B_THERMAL_AXISYM); if (sh.nn == 3) FEMB_F(2, 3, FEMB200_B_THERMAL_AXISYM); break;
  }
#undef FEMB_F
  return -2;
}

// staged front end; returns -2 when this kernel does not apply (caller keeps the node-row kernel)
int srows_staged(femb200_ctx *ctx, const ElemShape &sh, const RowArgs &a, const int32_t *order, int64_t v0, int64_t v1, int64_t rowlen_cap) {
  switch (sh.bkind) {
  case FEMB200_B_XYZ: return srows_nn<3, FEMB200_B_XYZ>(ctx, sh.nn, a, order, v0, v1, rowlen_cap);
  case FEMB200_B_PLANE: return srows_nn<2, FEMB200_B_PLANE>(ctx, sh.nn, a, order, v0, v1, rowlen_cap);
  case FEMB200_B_AXISYM: return srows_nn<2, FEMB200_B_AXISYM>(ctx, sh.nn, a, order, v0, v1, rowlen_cap);
  case FEMB200_B_THERMAL:
    return sh.d == 3 ? srows_nn<3, FEMB200_B_THERMAL>(ctx, sh.nn, a, order, v0, v1, rowlen_cap)
                     : srows_nn<2, FEMB200_B_THERMAL>(ctx, sh.nn, a, order, v0, v1, rowlen_cap);
  case FEMB200_B_THERMAL_AXISYM: return srows_nn<2, FEMB200_B_THERMAL_AXISYM>(ctx, sh.nn, a, order, v0, v1, rowlen_cap);
  }
  return -2;
}

} // namespace femb200
