// keblocks.cu -- numeric phase for the larger 3-D solid elements (Hexa8, Tetra10, Hexa20 with the XYZ strain-displacement
// filler, straindisplacement.go:5-26): element blocks first, ordered row gather second.
//
// For elements with many nodes and quadrature points the row-centric kernels re-read the staged geometry of an element once
// per incident node row (20 x 17 KB for a Hexa20) and cannot use K_ba = K_ab-transposed.  Here
//
//   k_elem_blocks  computes every 3 x 3 block K_ab of the element matrix ONCE, through the gradient moment
//                     M_ab[c][d] = sum_q f_q g_a^q[c] g_b^q[d]              (f_q = dJac*w*scale, g = J^-1 dN, both from k_geometry)
//                     K_ab[i][j] = sum_{c,d} C[v(i,c)][v(j,d)] M_ab[c][d]   (v(i,c): the strain row in which column i of B_a holds g_a[c])
//                  which is B_a^T C B_b summed over the quadrature points with the sum over q pulled inside -- the work per
//                  block and point drops from 27 + multiply-adds to 9, and a block pair (a, b), (b, a) shares M.  A lane owns a
//                  2 x 2 tile of node pairs (36 accumulators; six 128-bit shared-memory loads feed 42 DFMA per point); the blocks go to a
//                  staging array ke[e][a][b][3][3] in HBM.
//   k_rows_gather  one warp per node-block row I: the row's incident (element, local node) pairs are visited in ascending
//                  order (the order of the reference's COO stream after its stable sort), the nn blocks K_a* of each -- one
//                  contiguous nn*72-byte read -- are added into the row's shared-memory accumulators at the slots of the
//                  element->slot map, and the finished row is written to the CSR values: a deterministic segmented reduction
//                  without atomics, as on every other path.
//
// The element matrix staging costs 8*(3 nn)^2 bytes per element written and read once (HBM bound: 57.6 KB per Hexa20 element
// against 264 KB of staged-record re-reads and twice the multiply-adds on the row-centric path).
#include "numeric.cuh"
#include <cstdlib>

namespace femb200 {

size_t ke_blocks_smem(int nn, int nqp);

namespace {

// strain row k with B_a[k][i] = g_a[c]   (from the selector table of BT<XYZ,3>)
__host__ __device__ constexpr int voigt_row(int i, int c) {
  using B = BT<FEMB200_B_XYZ, 3>;
  int kr = -1;
  for (int k = 0; k < B::DIMC; ++k)
    if (B::sel(k, i) == c) kr = k;
  return kr;
}

template <int NN> struct KeShape {
  static_assert(NN % 2 == 0, "node pairs");
  static constexpr int NH = NN / 2;                                   // node pairs
  static constexpr int NBP = NH * (NH + 1) / 2;                       // 2 x 2 tiles A <= B
  static constexpr int T = 64;                                        // threads per block of k_elem_blocks
  static constexpr int EPG = T / NBP;                                 // elements per block: 6 (Hexa8), 4 (Tetra10), 1 (Hexa20)
  static constexpr int ITEMS = EPG * NBP;                             // 60 / 60 / 55 of 64 lanes busy
  static constexpr int ENT = NN * 9;                                  // doubles of one block row K_a*
  static constexpr int KE = NN * ENT;                                 // doubles of one element matrix
  static_assert(EPG >= 1, "tile list fits one block");
};

struct KeParams {
  double C[36];
  int symc;     // C == C^T exactly: K_ba is the transpose of K_ab
};

// ORTHO: C couples no normal strain with a shear strain and its shear block is diagonal (checked exactly on the host; every
// isotropic / orthotropic material): the products with the structural zeros are skipped -- adding +-0 changes no sum.
__host__ __device__ constexpr bool ortho_nonzero(int k, int l) { return (k < 3 && l < 3) || k == l; }

template <bool TRANSPOSED, bool ORTHO>
__device__ __forceinline__ void contract(const KeParams &p, const double (&M)[9], double (&K)[9]) {
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      double s = 0.0;
#pragma unroll
      for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          // K_ab[i][j] += C[v(i,c)][v(j,d)] M[c][d];   K_ba[j][i] += C[v(j,d)][v(i,c)] M[c][d]
          const int kr = TRANSPOSED ? voigt_row(j, d) : voigt_row(i, c), kc = TRANSPOSED ? voigt_row(i, c) : voigt_row(j, d);
          if (!ORTHO || ortho_nonzero(kr, kc)) s = fma(p.C[kr * 6 + kc], M[c * 3 + d], s);
        }
      if constexpr (TRANSPOSED) K[j * 3 + i] = s; else K[i * 3 + j] = s;
    }
}

__device__ __forceinline__ void store_block(double *dst, const double (&K)[9]) {
#pragma unroll
  for (int m = 0; m < 9; ++m) dst[m] = K[m];
}

// A block of 64 threads takes EPG consecutive elements; thread = one 2 x 2 tile of node pairs of one of them.  The group's
// staged geometry (contiguous: consecutive elements) arrives by ONE bulk asynchronous copy (TMA 1-D, mbarrier) -- plain
// loads were latency bound, every quadrature point a fresh DRAM line -- and is read as 128-bit shared-memory loads, a lane's
// two adjacent nodes per component (StagSoa layout: conflict free).  The finished blocks are collected in the same shared
// memory in the layout of the staging array and leave as one contiguous, fully coalesced stream (per-lane 72-byte block
// stores were measured first: one sector per lane and instruction, 6x slower than the HBM bound).
template <int NN, bool ORTHO>
__global__ void __launch_bounds__(KeShape<NN>::T)
k_elem_blocks(int64_t n_elem, const double *__restrict__ stag, int nqp, double *__restrict__ ke, const KeParams p) {
  using S = KeShape<NN>;
  constexpr int NNP = StagSoa<NN>::NNP, QREC = StagSoa<NN>::QREC;
  extern __shared__ __align__(128) double s_buf[];  // geometry records [EPG][nqp][QREC], then the blocks [EPG][NN][NN][3][3]
  __shared__ uint64_t bar;
  const int item = threadIdx.x;
  const int64_t e0 = (int64_t)blockIdx.x * S::EPG;
  const int64_t left = n_elem - e0;
  const int ng = (int)(left < S::EPG ? left : S::EPG);
  const int recsz = nqp * QREC;
  if (item == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    mbar_expect_tx(&bar, (uint32_t)(ng * recsz * 8));
    bulk_g2s(s_buf, stag + e0 * recsz, (uint32_t)(ng * recsz * 8), &bar);
  }
  const int el = item / S::NBP, blk = item - el * S::NBP;
  const bool act = item < S::ITEMS && el < ng;
  int A = 0, rem = blk;
  while (rem >= S::NH - A) { rem -= S::NH - A; ++A; }
  const int Bq = A + rem;
  const double *rec = s_buf + (act ? el : 0) * recsz; // idle lanes read a valid record and drop the result
  const double *pa = rec + 2 * A, *pb = rec + 2 * Bq; // nodes 2A, 2A+1 and 2B, 2B+1 of component 0
  double M[4][9];
#pragma unroll
  for (int t = 0; t < 4; ++t)
#pragma unroll
    for (int m = 0; m < 9; ++m) M[t][m] = 0.0;
  __syncthreads();                                    // the barrier is initialised
  mbar_wait(&bar, 0);
#pragma unroll 3
  for (int q = 0; q < nqp; ++q) {
    double2 ga[3], gb[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      ga[c] = *reinterpret_cast<const double2 *>(pa + q * QREC + c * NNP);
      gb[c] = *reinterpret_cast<const double2 *>(pb + q * QREC + c * NNP);
    }
    const double f = rec[q * QREC + 3 * NNP];
    double u1[3], u2[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) { u1[c] = ga[c].x * f; u2[c] = ga[c].y * f; }
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        M[0][c * 3 + d] = fma(u1[c], gb[d].x, M[0][c * 3 + d]);
        M[1][c * 3 + d] = fma(u1[c], gb[d].y, M[1][c * 3 + d]);
        M[2][c * 3 + d] = fma(u2[c], gb[d].x, M[2][c * 3 + d]);
        M[3][c * 3 + d] = fma(u2[c], gb[d].y, M[3][c * 3 + d]);
      }
  }
  __syncthreads();                                    // every record has been read: the buffer now collects the blocks
  if (act) {
    double *const kel = s_buf + el * S::KE;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int na = 2 * A + (t >> 1), nb = 2 * Bq + (t & 1);
      double K[9];
      contract<false, ORTHO>(p, M[t], K);
      store_block(kel + (na * NN + nb) * 9, K);
      if (A != Bq) {                       // the mirrored block (nb, na); on a diagonal tile all four pairs are direct
        double Kt[9];
        if (p.symc) {
#pragma unroll
          for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) Kt[j * 3 + i] = K[i * 3 + j];
        } else {
          contract<true, ORTHO>(p, M[t], Kt);
        }
        store_block(kel + (nb * NN + na) * 9, Kt);
      }
    }
  }
  __syncthreads();
  const int n2 = ng * (S::KE / 2);
  double2 *const dst = reinterpret_cast<double2 *>(ke + e0 * (int64_t)S::KE);
  const double2 *const src = reinterpret_cast<const double2 *>(s_buf);
  for (int i = threadIdx.x; i < n2; i += S::T) dst[i] = src[i];
}

// one warp per node-block row; the WPB warps of a block are independent (more resident warps than the 32 blocks an SM takes)
template <int NN, int U, int WPB>
__global__ void __launch_bounds__(32 * WPB)
k_rows_gather(const RowArgs a, const double *__restrict__ ke, int acc_stride) {
  using S = KeShape<NN>;
  constexpr int ENT = S::ENT, NPL = (ENT + 31) / 32;
  extern __shared__ double s_all[];        // per warp [rl][3][3]
  const int lane = threadIdx.x & 31;
  double *const s_acc = s_all + (threadIdx.x >> 5) * acc_stride;
  const int64_t vis = a.v0 + (int64_t)blockIdx.x * WPB + (threadIdx.x >> 5);
  if (vis >= a.v1) return;
  const int64_t I = a.order ? (int64_t)__ldg(a.order + vis) : vis;
  const int beg = __ldg(a.n2e_ptr + I), end = __ldg(a.n2e_ptr + I + 1);
  const int64_t b0 = __ldg(a.blkptr + I);
  const int rl = (int)(__ldg(a.blkptr + I + 1) - b0);
  if (rl == 0) return;                     // unreferenced node, or a row this rank does not assemble
  const int tot = rl * 9;
  for (int m = lane; m < tot; m += 32) s_acc[m] = 0.0;
  int eb[NPL], eo[NPL];                    // this lane's entries of a block row: local node b and offset inside the block
#pragma unroll
  for (int r = 0; r < NPL; ++r) {
    const int idx = lane + 32 * r;
    eb[r] = idx < ENT ? idx / 9 : 0;
    eo[r] = idx < ENT ? idx % 9 : -1;
  }
  __syncwarp();
#pragma unroll 1
  for (int k0 = beg; k0 < end; k0 += 32) {
    const int kq = k0 + lane;
    const uint32_t t = kq < end ? __ldg(a.n2e_t + kq) : 0xffffffffu;
    const bool num = kq < end && (int64_t)(t / NN) < a.n_numeric;   // lists ascend in element index, pattern-only elements last
    const int cnt = __popc(__ballot_sync(0xffffffffu, num));
    const int dg = num ? (int)__ldg(a.degen + t / NN) : 0;
#pragma unroll 1
    for (int kk = 0; kk < cnt; kk += U) {
      double v[U][NPL];
      int slot[U], dgu[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int ku = kk + u < cnt ? kk + u : cnt - 1;
        const uint32_t tk = __shfl_sync(0xffffffffu, t, ku);
        dgu[u] = __shfl_sync(0xffffffffu, dg, ku);
        const double *src = ke + (int64_t)tk * ENT;                  // tk = e*nn + a: block row a of element e
        slot[u] = lane < NN ? (int)(__ldg(a.pos16 + (int64_t)(k0 + ku) * NN + lane) & pos_mask(NN)) : 0;
#pragma unroll
        for (int r = 0; r < NPL; ++r) v[u][r] = (eo[r] >= 0) ? __ldg(src + lane + 32 * r) : 0.0;
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (kk + u < cnt) {
          if (!dgu[u]) {
#pragma unroll
            for (int r = 0; r < NPL; ++r) {
              const int sb = __shfl_sync(0xffffffffu, slot[u], eb[r]);
              if (eo[r] >= 0) s_acc[sb * 9 + eo[r]] += v[u][r];
            }
          } else {
            // collapsed element (a repeated node): two local nodes share a slot -- local nodes one at a time, ascending
#pragma unroll 1
            for (int b = 0; b < NN; ++b) {
              const int sb = __shfl_sync(0xffffffffu, slot[u], b);
#pragma unroll
              for (int r = 0; r < NPL; ++r)
                if (eo[r] >= 0 && eb[r] == b) s_acc[sb * 9 + eo[r]] += v[u][r];
              __syncwarp();
            }
          }
        }
        __syncwarp();
      }
    }
  }
  // copy-out: scalar row i of the node row = [p][j] over the row's blocks
  double *const vrow = a.values + 9 * b0;
  const int rowsz = rl * 3;
#pragma unroll
  for (int i = 0; i < 3; ++i)
    for (int x = lane; x < rowsz; x += 32) {
      const int pcol = x / 3, j = x - 3 * pcol;
      vrow[i * rowsz + x] = s_acc[pcol * 9 + i * 3 + j];
    }
}

template <int NN, bool ORTHO>
int elem_blocks_T(femb200_ctx *ctx, int64_t n_elem, const double *stag, int nqp, double *ke, const KeParams &p) {
  using S = KeShape<NN>;
  if (n_elem <= 0) return 0;
  const size_t smem = ke_blocks_smem(NN, nqp);
  if (smem + 1024 > ctx->smem_optin) return -2;
  auto kern = k_elem_blocks<NN, ORTHO>;
  FEMB_CUDA_OK(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<grid_for(n_elem, S::EPG), S::T, smem, ctx->stream>>>(n_elem, stag, nqp, ke, p);
  ctx->launches++;
  FEMB_CUDA_OK(ctx, cudaGetLastError());
  return 0;
}

template <int NN, int WPB>
int rows_gather_W(femb200_ctx *ctx, const RowArgs &a, const double *ke, int64_t rowlen_cap) {
  constexpr int U = NN == 20 ? 2 : 4;
  const int acc_stride = (int)rowlen_cap * 9;
  const size_t smem = (size_t)acc_stride * WPB * sizeof(double);
  if (smem > ctx->smem_optin) return -2;
  auto kern = k_rows_gather<NN, U, WPB>;
  FEMB_CUDA_OK(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smem_optin));
  if (a.v1 > a.v0) {
    kern<<<grid_for(a.v1 - a.v0, WPB), 32 * WPB, smem, ctx->stream>>>(a, ke, acc_stride);
    ctx->launches++;
  }
  FEMB_CUDA_OK(ctx, cudaGetLastError());
  return 0;
}

template <int NN>
int rows_gather_T(femb200_ctx *ctx, const RowArgs &a, const double *ke, int64_t rowlen_cap) {
  // two warps per block for short rows (measured: Hexa8 1.44 -> 1.34 ms per 10^6 elements; no gain or a loss for the long rows of
  // Tetra10 / Hexa20, whose accumulators would leave fewer resident warps)
  int wpb = rowlen_cap <= 27 ? 2 : 1;
  if (getenv("FEMB200_GATHER_WPB")) wpb = atoi(getenv("FEMB200_GATHER_WPB"));
  if (wpb >= 4) { const int rc = rows_gather_W<NN, 4>(ctx, a, ke, rowlen_cap); if (rc != -2) return rc; }
  if (wpb >= 2) { const int rc = rows_gather_W<NN, 2>(ctx, a, ke, rowlen_cap); if (rc != -2) return rc; }
  return rows_gather_W<NN, 1>(ctx, a, ke, rowlen_cap);
}

} // namespace

size_t ke_doubles_per_element(int nn) { return (size_t)nn * nn * 9; }

template <int NN> static size_t blocks_smem_T(int nqp) {
  using S = KeShape<NN>;
  const size_t smem_rec = (size_t)S::EPG * nqp * StagSoa<NN>::QREC * sizeof(double), smem_ke = (size_t)S::EPG * S::KE * sizeof(double);
  return smem_rec > smem_ke ? smem_rec : smem_ke;
}
// dynamic shared memory of k_elem_blocks (0: element type not handled here)
size_t ke_blocks_smem(int nn, int nqp) {
  switch (nn) {
  case 8: return blocks_smem_T<8>(nqp);
  case 10: return blocks_smem_T<10>(nqp);
  case 20: return blocks_smem_T<20>(nqp);
  }
  return 0;
}

static void make_params(const double *C36, KeParams *p, bool *ortho) {
  p->symc = 1;
  *ortho = true;
  for (int k = 0; k < 6; ++k)
    for (int l = 0; l < 6; ++l) {
      p->C[k * 6 + l] = C36[k * 6 + l];
      if (C36[k * 6 + l] != C36[l * 6 + k]) p->symc = 0;
      if (!ortho_nonzero(k, l) && C36[k * 6 + l] != 0.0) *ortho = false;
    }
}

int ke_elem_blocks(femb200_ctx *ctx, int nn, int64_t n_elem, const double *stag, int nqp, double *ke, const double *C36) {
  KeParams p;
  bool ortho;
  make_params(C36, &p, &ortho);
  switch (nn) {
  case 8: return ortho ? elem_blocks_T<8, true>(ctx, n_elem, stag, nqp, ke, p) : elem_blocks_T<8, false>(ctx, n_elem, stag, nqp, ke, p);
  case 10: return ortho ? elem_blocks_T<10, true>(ctx, n_elem, stag, nqp, ke, p) : elem_blocks_T<10, false>(ctx, n_elem, stag, nqp, ke, p);
  case 20: return ortho ? elem_blocks_T<20, true>(ctx, n_elem, stag, nqp, ke, p) : elem_blocks_T<20, false>(ctx, n_elem, stag, nqp, ke, p);
  }
  return -2;
}

int ke_rows_gather(femb200_ctx *ctx, int nn, const RowArgs &a, const double *ke, int64_t rowlen_cap) {
  switch (nn) {
  case 8: return rows_gather_T<8>(ctx, a, ke, rowlen_cap);
  case 10: return rows_gather_T<10>(ctx, a, ke, rowlen_cap);
  case 20: return rows_gather_T<20>(ctx, a, ke, rowlen_cap);
  }
  return -2;
}

} // namespace femb200
