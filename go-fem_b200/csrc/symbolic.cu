// symbolic.cu -- on-device symbolic phase: node->element adjacency, node-block row pattern,
// row-local slot positions and the scalar CSR pattern (indptr / indices).
//
// Replaces, for the whole call, what the reference does implicitly by emitting nd^2 COO triplets per
// element (assembler.go:79-92) and letting lap.Sparse.Accumulate merge duplicates
// (assembler_isoparametric.go:78-79,121).  Nothing is emitted per dof pair here: the pattern is built at
// NODE-BLOCK granularity (nn^2 pairs per element instead of nd^2) and expanded once.
//
//   conn[e][a]                      --histogram+scan-->  n2e_ptr[node]
//   t = e*nn+a                      --atomic cursor-->   n2e_t grouped by node (lists sorted ascending by the
//                                                        pattern kernel => deterministic, ascending element)
//   per row node I                  --sorted insert-->   column nodes of row I (ascending), rowlen[I]
//   second sweep over the incidences ---------------->   pos16[k][b] = position of node b of e in row I
//   scan(rowlen) -> blkptr;  expand -> indptr (int64), indices (int32)
#include "common.cuh"
#include <cub/cub.cuh>
#include <type_traits>
#include <cstring>

namespace femb200 {

// ------------------------------------------------------------------------------------------------
// cnt != NULL: the incidence histogram of the adjacency (k_count_incidence) is taken in the same pass; AGG as there.
template <class IntT, bool AGG>
__global__ void k_narrow_validate(const IntT *__restrict__ in, int32_t *__restrict__ out, int64_t n, int nn,
                                  int64_t n_nodes, unsigned long long *err, int32_t *__restrict__ cnt) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  long long v = (long long)in[t];
  if (v < 0 || v >= n_nodes) {
    atomicMin(err, pack_error(t / nn, 0, 0)); // kind 0 in the device word = node out of range: beats every geometry error of the SAME
                                              // element (the reference panics while gathering the coordinates, before the Jacobian)
    v = 0;
  }
  out[t] = (int32_t)v;
  if (cnt) {
    const int32_t node = (int32_t)v;
    if constexpr (AGG) {
      const unsigned peers = __match_any_sync(__activemask(), node);
      if ((threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&cnt[node], __popc(peers));
    } else {
      atomicAdd(&cnt[node], 1);
    }
  }
}

int narrow_and_validate_conn(femb200_assembler *ga, const void *d_in, bool is64, int64_t n, int nn, int32_t *out, int32_t *cnt, bool agg) {
  femb200_ctx *ctx = ga->ctx;
  if (n == 0) return 0;
  const unsigned g = grid_for(n, 256);
  if (is64) {
    if (agg) k_narrow_validate<long long, true><<<g, 256, 0, ctx->stream>>>((const long long *)d_in, out, n, nn, ga->n_nodes, ctx->d_err, cnt);
    else k_narrow_validate<long long, false><<<g, 256, 0, ctx->stream>>>((const long long *)d_in, out, n, nn, ga->n_nodes, ctx->d_err, cnt);
  } else {
    if (agg) k_narrow_validate<int32_t, true><<<g, 256, 0, ctx->stream>>>((const int32_t *)d_in, out, n, nn, ga->n_nodes, ctx->d_err, cnt);
    else k_narrow_validate<int32_t, false><<<g, 256, 0, ctx->stream>>>((const int32_t *)d_in, out, n, nn, ga->n_nodes, ctx->d_err, cnt);
  }
  ctx->launches++;
  FEMB_CUDA_OK(ctx, cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------------
// Warp-aggregated: neighbouring elements share nodes, so the lanes of a warp that hit the same node (match.any) send ONE
// atomic with their count -- the histogram and the fill are bound by L2 atomic throughput.
// (match.any costs a step per distinct value in the warp: it pays for linear tetrahedra -- 8 consecutive elements share
// most of their nodes, config 1: adjacency 0.112 -> 0.072 ms -- and loses for quads / hexahedra / quadratic elements.)
template <bool AGG>
__global__ void k_count_incidence(const int32_t *__restrict__ conn, int64_t n, int32_t *__restrict__ cnt) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const int32_t node = conn[t];
  if constexpr (AGG) {
    const unsigned peers = __match_any_sync(__activemask(), node);
    if ((threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&cnt[node], __popc(peers));
  } else {
    atomicAdd(&cnt[node], 1);
  }
}

// Incidence lists by atomic cursor: n2e_t[n2e_ptr[node] + slot] = t.  The slot order depends on the atomics'
// arrival order; every list is sorted (ascending t = ascending element) by the row-pattern kernel that
// consumes it, which is what makes the later row reduction deterministic.
template <bool AGG>
__global__ void k_fill_incidence(const int32_t *__restrict__ conn, int64_t n, const int32_t *__restrict__ n2e_ptr,
                                 int32_t *__restrict__ cursor, uint32_t *__restrict__ n2e_t) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const int32_t node = conn[t];
  int slot;
  if constexpr (AGG) {
    const int lane = threadIdx.x & 31;
    const unsigned peers = __match_any_sync(__activemask(), node);
    const int leader = __ffs(peers) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(&cursor[node], __popc(peers));
    base = __shfl_sync(peers, base, leader);
    slot = base + __popc(peers & ((1u << lane) - 1u));
  } else {
    slot = atomicAdd(&cursor[node], 1);
  }
  n2e_t[n2e_ptr[node] + slot] = (uint32_t)t;
}

// One thread per row node: builds the ascending list of distinct column nodes of the node-block row.
// The working list lives in shared memory, interleaved by thread (conflict free), and spills to the
// thread's private slice of `tmpcols` (capacity nn * valence >= any possible row length) on overflow.
template <int NN>
__global__ void __launch_bounds__(128)
k_row_pattern(int64_t n_nodes, const int32_t *__restrict__ n2e_ptr, uint32_t *__restrict__ n2e_t,
              const int32_t *__restrict__ conn, int32_t *__restrict__ tmpcols, int64_t *__restrict__ rowlen,
              uint16_t *__restrict__ pos16, int cap_s, unsigned long long *err, int only_flagged,
              const uint8_t *__restrict__ node_mask) {
  extern __shared__ int32_t s_list[];
  const int T = blockDim.x, tid = threadIdx.x;
  int64_t I = (int64_t)blockIdx.x * T + tid;
  if (I >= n_nodes) return;
  if (I == 0) rowlen[n_nodes] = 0;                           // tail entry of the scan input
  if (node_mask && !node_mask[I]) { rowlen[I] = 0; return; } // row not assembled on this rank
  if (only_flagged && rowlen[I] >= 0) return; // already done by k_row_pattern_group
  const int beg = n2e_ptr[I], end = n2e_ptr[I + 1];
  for (int a = beg + 1; a < end; ++a) { // incidence list into ascending element order (insertion sort, in place)
    const uint32_t v = n2e_t[a];
    int j = a - 1;
    while (j >= beg && n2e_t[j] > v) { n2e_t[j + 1] = n2e_t[j]; --j; }
    n2e_t[j + 1] = v;
  }
  int32_t *gl = tmpcols + (int64_t)NN * beg;
  bool in_s = true;
  int cnt = 0;
#define LST(i) (in_s ? s_list[(i) * T + tid] : gl[(i)])
#define LST_SET(i, v)                                   \
  do {                                                  \
    if (in_s) s_list[(i) * T + tid] = (v); else gl[(i)] = (v); \
  } while (0)
  for (int k = beg; k < end; ++k) {
    const int64_t e = n2e_t[k] / NN;
    for (int b = 0; b < NN; ++b) {
      const int32_t J = conn[e * NN + b];
      int lo = 0, hi = cnt;
      while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (LST(mid) < J) lo = mid + 1; else hi = mid;
      }
      if (lo < cnt && LST(lo) == J) continue;
      if (in_s && cnt == cap_s) { // spill the shared list to the global slice and continue there
        for (int m = 0; m < cnt; ++m) gl[m] = s_list[m * T + tid];
        in_s = false;
      }
      for (int m = cnt; m > lo; --m) LST_SET(m, LST(m - 1));
      LST_SET(lo, J);
      ++cnt;
    }
  }
  if (in_s) for (int m = 0; m < cnt; ++m) gl[m] = s_list[m * T + tid];
  rowlen[I] = cnt;
  // second sweep: row-local position of every (incidence, local node) = the element->slot map
  for (int k = beg; k < end; ++k) {
    const int64_t e = n2e_t[k] / NN;
    for (int b = 0; b < NN; ++b) {
      const int32_t J = conn[e * NN + b];
      int lo = 0, hi = cnt;
      while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (LST(mid) < J) lo = mid + 1; else hi = mid;
      }
      pos16[(int64_t)k * NN + b] = (uint16_t)(lo | ((NN <= 4 && b == 0) ? (int)((n2e_t[k] % NN) << kPosLaShift) : 0));
    }
  }
#undef LST
#undef LST_SET
}

// Fast path of the row pattern: a group of G = 8 lanes per row node.
//   1. the lanes load the node's (<= 32) incidences in one go and sort them by rank counting in shared
//      memory (ascending t = ascending element); the sorted list is written back to n2e_t,
//   2. every lane gathers the connectivity of its incidences (all gathers of the row are in flight at once --
//      one memory round trip per row instead of one per incident element),
//   3. the candidate column nodes go into a small open-addressing hash table in shared memory (atomicCAS),
//   4. the distinct keys are compacted (ballot + popc) and ranked by counting => ascending column list,
//      each table slot learns its rank,
//   5. every (incidence, local node) is turned into its row-local slot position with one probe
//      (the element -> CSR-slot map), written with vector stores.
// Rows with more than 32 incidences or more than CAP distinct columns are flagged (rowlen = -1) and redone
// by the sorted-insert kernel k_row_pattern.
template <int NN, int HS, int IPL, int TPB>
__global__ void __launch_bounds__(TPB)
k_row_pattern_group(int64_t n_nodes, const int32_t *__restrict__ n2e_ptr, uint32_t *__restrict__ n2e_t,
                    const int32_t *__restrict__ conn, int32_t *__restrict__ tmpcols, int64_t *__restrict__ rowlen,
                    uint16_t *__restrict__ pos16, const uint8_t *__restrict__ node_mask, int only_flagged) {
  constexpr int G = 8, MAXINC = G * IPL; // IPL incidences per lane: rows with more than 8*IPL incidences take the fallback
  constexpr int CAP = HS * 3 / 4;
  constexpr int LOG = HS == 32 ? 5 : (HS == 64 ? 6 : 7);
  constexpr int GW = HS + HS + MAXINC + ((CAP + 4 + 3) & ~3); // 32-bit words of shared memory per group (every part a multiple of 4: 16-byte loads)
  extern __shared__ int32_t s_mem[];
  const int tid = threadIdx.x, l = tid & (G - 1);
  int32_t *keys = s_mem + (tid / G) * GW;
  int32_t *rk = keys + HS;
  uint32_t *lst = reinterpret_cast<uint32_t *>(rk + HS);
  int32_t *ul = reinterpret_cast<int32_t *>(lst + MAXINC);
  const int64_t I = ((int64_t)blockIdx.x * blockDim.x + tid) / G;
  if (I >= n_nodes) return;
  if (I == 0 && l == 0) rowlen[n_nodes] = 0;                              // tail entry of the scan input
  if (only_flagged && rowlen[I] >= 0) return;                             // done by the previous (cheaper) pass
  if (node_mask && !node_mask[I]) { if (l == 0) rowlen[I] = 0; return; } // row not assembled on this rank
  const int gshift = (tid & 31) & ~(G - 1);
  const unsigned gmask = 0xffu << gshift;
  const int beg = n2e_ptr[I], end = n2e_ptr[I + 1];
  const int m = end - beg;
  if (m > MAXINC) { if (l == 0) rowlen[I] = -1; return; }

  // 1. incidences, sorted (rank counting against the whole list, four list entries per shared-memory load)
  uint32_t tv[IPL];
#pragma unroll
  for (int s = 0; s < IPL; ++s) {
    const int idx = l + G * s;
    tv[s] = idx < m ? n2e_t[beg + idx] : 0xffffffffu;
    lst[idx] = tv[s];                                            // MAXINC entries: the tail is padding that ranks above everything
  }
  __syncwarp(gmask);
  // (A look-before-sorting test -- is the list already ascending? -- was measured: the atomic fill leaves almost no list in
  // order, not even the warp-aggregated one, so the look is pure overhead.)
  {
    int rr[IPL];
#pragma unroll
    for (int s = 0; s < IPL; ++s) rr[s] = 0;
    const uint4 *l4 = reinterpret_cast<const uint4 *>(lst);
    for (int j = 0; j < (m + 3) / 4; ++j) {
      const uint4 w = l4[j];
#pragma unroll
      for (int s = 0; s < IPL; ++s) rr[s] += (w.x < tv[s]) + (w.y < tv[s]) + (w.z < tv[s]) + (w.w < tv[s]);
    }
#pragma unroll
    for (int s = 0; s < IPL; ++s)
      if (l + G * s < m) {
        rk[rr[s]] = (int32_t)tv[s];
        n2e_t[beg + rr[s]] = tv[s];
      }
    __syncwarp(gmask);
#pragma unroll
    for (int s = 0; s < IPL; ++s) tv[s] = (l + G * s < m) ? (uint32_t)rk[l + G * s] : 0xffffffffu;
    __syncwarp(gmask);
  }

  // 2. connectivity of my incidences
  int32_t cj[IPL][NN];
#pragma unroll
  for (int s = 0; s < IPL; ++s) {
    if (l + G * s < m) {
      const int32_t *ce = conn + (int64_t)(tv[s] / NN) * NN;
      if constexpr (NN % 4 == 0) {
#pragma unroll
        for (int v4 = 0; v4 < NN / 4; ++v4) {
          const int4 w = __ldg(reinterpret_cast<const int4 *>(ce) + v4);
          cj[s][4 * v4] = w.x; cj[s][4 * v4 + 1] = w.y; cj[s][4 * v4 + 2] = w.z; cj[s][4 * v4 + 3] = w.w;
        }
      } else {
#pragma unroll
        for (int b = 0; b < NN; ++b) cj[s][b] = __ldg(ce + b);
      }
    }
  }

  // 3. hash insert (shared-memory atomicCAS; a lock-step variant with plain stores and a re-read after a group barrier was
  // measured at twice the time of this kernel: two barriers per probe cost more than the atomic)
  for (int h = l; h < HS; h += G) keys[h] = -1;
  __syncwarp(gmask);
  bool ovf = false;
#pragma unroll
  for (int s = 0; s < IPL; ++s) {
    if (l + G * s < m) {
#pragma unroll
      for (int b = 0; b < NN; ++b) {
        const int32_t J = cj[s][b];
        unsigned h = ((unsigned)J * 0x9E3779B1u) >> (32 - LOG);
        int probe = 0;
        for (; probe < HS; ++probe) {
          // most candidates are duplicates (every neighbour shows up in 4-6 incident elements): look before the atomic
          int32_t cur = *reinterpret_cast<volatile int32_t *>(&keys[h]);
          if (cur == -1) cur = atomicCAS(&keys[h], -1, J);
          if (cur == -1 || cur == J) break;
          h = (h + 1) & (HS - 1);
        }
        ovf = ovf || (probe == HS);
      }
    }
  }
  __syncwarp(gmask);

  // 4. compact the distinct keys, rank them
  // large, mostly empty tables (Tetra10: 20-65 of 128 slots): rank the compacted list, not every slot.  Measured: Tetra10 row
  // pattern 1.21 -> 1.03 ms per 748,800 elements; Hexa20 (50-81 of 128 slots) 2 % slower, Hexa8 equal: not used there.
  constexpr bool RANK_COMPACT = (HS / G > 4) && (NN * 8 < HS);
  int cnt = 0;
#pragma unroll
  for (int s = 0; s < HS / G; ++s) {
    const int32_t key = keys[s * G + l];
    const bool has = key >= 0;
    const unsigned bal = (__ballot_sync(gmask, has) >> gshift) & 0xffu;
    const int off = cnt + __popc(bal & ((1u << l) - 1u));
    if (has && off < CAP) ul[off] = key;
    if constexpr (RANK_COMPACT) rk[s * G + l] = has ? off : 0;   // slot -> position in the compacted list (rank follows below)
    cnt += __popc(bal);
  }
  ovf = __any_sync(gmask, ovf) || cnt > CAP;
  if (ovf) { if (l == 0) rowlen[I] = -1; return; }
  if (l < 4 && cnt + l < CAP + 4) ul[cnt + l] = 0x7fffffff;       // padding for the four-wide rank loop (ul holds CAP + 4 words)
  __syncwarp(gmask);
  int32_t *gcols = tmpcols + (int64_t)NN * beg;
  if constexpr (RANK_COMPACT) {
    // lane l ranks the compacted keys l, l+8, ... (four at a time against the whole list): cnt/8 keys per lane instead of
    // the HS/8 slots of the table
    constexpr int NCH = (CAP + 4 * G - 1) / (4 * G);
    int rr[NCH][4];
    int32_t key[NCH][4];
    const int4 *u4 = reinterpret_cast<const int4 *>(ul);
    const int nj = (cnt + 3) / 4;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
#pragma unroll
      for (int i = 0; i < 4; ++i) { rr[c][i] = 0; key[c][i] = 0x7fffffff; }
      if (c * 4 * G < cnt) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int x = c * 4 * G + l + G * i;
          if (x < cnt) key[c][i] = ul[x];
        }
        for (int j = 0; j < nj; ++j) {
          const int4 w = u4[j];
#pragma unroll
          for (int i = 0; i < 4; ++i) rr[c][i] += (w.x < key[c][i]) + (w.y < key[c][i]) + (w.z < key[c][i]) + (w.w < key[c][i]);
        }
      }
    }
    __syncwarp(gmask);                                             // every lane has read the keys: the list now takes the ranks
#pragma unroll
    for (int c = 0; c < NCH; ++c)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int x = c * 4 * G + l + G * i;
        if (x < cnt) { ul[x] = rr[c][i]; gcols[rr[c][i]] = key[c][i]; }
      }
    __syncwarp(gmask);
#pragma unroll
    for (int s = 0; s < HS / G; ++s)
      if (keys[s * G + l] >= 0) rk[s * G + l] = ul[rk[s * G + l]];
  } else {
    int32_t key[HS / G];
    int rr[HS / G];
#pragma unroll
    for (int s = 0; s < HS / G; ++s) { key[s] = keys[s * G + l]; rr[s] = 0; }
    const int4 *u4 = reinterpret_cast<const int4 *>(ul);
    for (int j = 0; j < (cnt + 3) / 4; ++j) {
      const int4 w = u4[j];
#pragma unroll
      for (int s = 0; s < HS / G; ++s) rr[s] += (w.x < key[s]) + (w.y < key[s]) + (w.z < key[s]) + (w.w < key[s]);
    }
#pragma unroll
    for (int s = 0; s < HS / G; ++s)
      if (key[s] >= 0) {
        rk[s * G + l] = rr[s];
        gcols[rr[s]] = key[s];
      }
  }
  __syncwarp(gmask);

  // 5. element -> slot positions
#pragma unroll
  for (int s = 0; s < IPL; ++s) {
    const int idx = l + G * s;
    if (idx < m) {
      int pb[NN];
#pragma unroll
      for (int b = 0; b < NN; ++b) {
        const int32_t J = cj[s][b];
        unsigned h = ((unsigned)J * 0x9E3779B1u) >> (32 - LOG);
        while (keys[h] != J) h = (h + 1) & (HS - 1);
        pb[b] = rk[h];
      }
      if constexpr (NN <= 4) pb[0] |= (int)((tv[s] % NN) << kPosLaShift);
      uint16_t *pk = pos16 + (int64_t)(beg + idx) * NN;
      if constexpr (NN % 4 == 0) {
#pragma unroll
        for (int v4 = 0; v4 < NN / 4; ++v4) {
          uint2 w;
          w.x = (unsigned)pb[4 * v4] | ((unsigned)pb[4 * v4 + 1] << 16);
          w.y = (unsigned)pb[4 * v4 + 2] | ((unsigned)pb[4 * v4 + 3] << 16);
          reinterpret_cast<uint2 *>(pk)[v4] = w;
        }
      } else {
#pragma unroll
        for (int b = 0; b < NN; ++b) pk[b] = (uint16_t)pb[b];
      }
    }
  }
  if (l == 0) rowlen[I] = cnt;
}

// need[s] = max over aligned groups of (4 << s) consecutive rows of sum(rowlen): shared-memory budget of a block of the
// packed node-row kernel.  Block-level max in shared memory, then one atomic per block and slot.
__global__ void __launch_bounds__(256)
k_group_need(int64_t n_nodes, const int64_t *__restrict__ rowlen, unsigned long long *__restrict__ need) {
  __shared__ int s_max[4];
  if (threadIdx.x < 4) s_max[threadIdx.x] = 0;
  __syncthreads();
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  int len = v < n_nodes ? (int)rowlen[v] : 0;
  int m[4];
  len += __shfl_xor_sync(0xffffffffu, len, 1);
  len += __shfl_xor_sync(0xffffffffu, len, 2);
  m[0] = len;
  len += __shfl_xor_sync(0xffffffffu, len, 4);
  m[1] = len;
  len += __shfl_xor_sync(0xffffffffu, len, 8);
  m[2] = len;
  len += __shfl_xor_sync(0xffffffffu, len, 16);
  m[3] = len;
#pragma unroll
  for (int s = 0; s < 4; ++s) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m[s] = max(m[s], __shfl_xor_sync(0xffffffffu, m[s], o));
    if (lane == 0) atomicMax(&s_max[s], m[s]);
  }
  __syncthreads();
  if (threadIdx.x < 4) atomicMax(need + threadIdx.x, (unsigned long long)s_max[threadIdx.x]);
}

// Two-class visiting order for meshes whose rows differ a lot in length (quadratic tetrahedra: vertex rows are 2-3x
// longer than mid-edge rows): rows with 0 < rowlen <= max/2 are "short".  The numeric kernel is launched once per class
// with the shared memory that class needs, which doubles the resident warps of the short class.
__global__ void k_class_flags(int64_t n_nodes, const int64_t *__restrict__ rowlen, const int64_t *__restrict__ d_max,
                              uint8_t *__restrict__ flag) {
  const int64_t I = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (I >= n_nodes) return;
  const int64_t rl = rowlen[I];
  flag[I] = (rl > 0 && 2 * rl <= *d_max) ? 1 : 0;
}

// scalar row pointers: row (I, m), m in [0, mdpn)
__global__ void k_indptr(int64_t n_nodes, int mdpn, int nu, unsigned umask, const int64_t *__restrict__ blkptr,
                         int64_t *__restrict__ indptr) {
  int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t ndof = n_nodes * mdpn;
  if (r > ndof) return;
  if (r == ndof) { indptr[r] = (int64_t)nu * nu * blkptr[n_nodes]; return; }
  int64_t I = r / mdpn;
  int m = (int)(r % mdpn);
  int before = __popc(umask & ((1u << m) - 1u));
  int64_t rl = blkptr[I + 1] - blkptr[I];
  indptr[r] = (int64_t)nu * nu * blkptr[I] + (int64_t)before * rl * nu;
}

// One warp per node row: expands the ascending column-node list of the row into the nu scalar CSR rows of the node.
// The nu*nu*rl column indices of a node are contiguous in `indices` (scalar row ur at offset ur*rl*nu), so lane l
// writes entries l, l+32, ...: fully coalesced 128-byte stores, the column nodes come from L1.
// LPN lanes per node row: a warp for long rows, 8 lanes when every scalar row fits 32 entries (2-D meshes).
template <int NU, int LPN>
__global__ void __launch_bounds__(256)
k_expand(int64_t n_nodes, int nn, int mdpn, int nu_rt, const int *__restrict__ umap_packed, const int32_t *__restrict__ n2e_ptr,
         const int64_t *__restrict__ blkptr, const int32_t *__restrict__ tmpcols, int32_t *__restrict__ indices,
         const int32_t *__restrict__ col_map, int32_t *__restrict__ blkcols, int64_t cap_blk, unsigned long long *__restrict__ guard) {
  const int nu = NU > 0 ? NU : nu_rt;
  const int lane = threadIdx.x & (LPN - 1);
  const int64_t I = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / LPN;
  if (I >= n_nodes) return;
  const int64_t b0 = blkptr[I];
  const int rl = (int)(blkptr[I + 1] - b0);
  if (rl == 0) return;
  if (b0 + rl > cap_blk) { if (lane == 0) atomicOr(guard, 1ull); return; } // hinted capacity exceeded: the call is repeated
  const int32_t *cols = tmpcols + (int64_t)nn * n2e_ptr[I];
  const int rowsz = rl * nu;
  int um[NU > 0 ? NU : kMaxNdn];
#pragma unroll
  for (int u = 0; u < (NU > 0 ? NU : kMaxNdn); ++u) um[u] = u < nu ? umap_packed[u] : 0;
  int32_t *out = indices + (int64_t)nu * nu * b0;
  for (int idx = lane; idx < rowsz; idx += LPN) {
    const int p = idx / nu, uc = idx - p * nu;
    int off = um[0];
#pragma unroll
    for (int u = 1; u < (NU > 0 ? NU : kMaxNdn); ++u) off = (uc == u) ? um[u] : off;
    const int32_t J = cols[p];
    if (uc == 0) blkcols[b0 + p] = J;
    const int32_t c = (col_map ? col_map[J] : J) * mdpn + off;
    for (int ur = 0; ur < nu; ++ur) out[(int64_t)ur * rowsz + idx] = c;
  }
}

// Expansion for the common case (identity dof map, rows of <= 32 node blocks): a warp takes EIGHT consecutive node rows.
// The row metadata of all eight is fetched by parallel lanes and every row's column list sits in a register of lane p,
// so a warp pays two memory round trips per eight rows instead of three per row (k_expand is bound by exactly that
// latency); it also writes the scalar row pointers of its rows (k_indptr) and the compact block-column list of the plan.
template <int NU>
__global__ void __launch_bounds__(256)
k_expand8(int64_t n_nodes, int nn, const int32_t *__restrict__ n2e_ptr, const int64_t *__restrict__ blkptr,
          const int32_t *__restrict__ tmpcols, int32_t *__restrict__ indices, int64_t *__restrict__ indptr,
          const int32_t *__restrict__ col_map, int32_t *__restrict__ blkcols, int64_t cap_blk, unsigned long long *__restrict__ guard,
          int64_t *__restrict__ scalars) {
  constexpr int R = 8;
  const int lane = threadIdx.x & 31;
  const int64_t I0 = (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5) * R;
  if (I0 >= n_nodes) return;
  const int nrow = (int)(n_nodes - I0 < R ? n_nodes - I0 : R);
  int64_t bp = 0;
  int np = 0;
  if (lane <= nrow) bp = __ldg(blkptr + I0 + lane);
  if (lane < nrow) np = __ldg(n2e_ptr + I0 + lane);
  const int64_t bp_next = __shfl_down_sync(0xffffffffu, bp, 1);
  const int rl_mine = lane < nrow ? (int)(bp_next - bp) : 0;     // lane r: length of row I0 + r
  // scalar row pointers: row (I0 + r, i) starts at NU*NU*bp_r + i*rl_r*NU
  if (lane < nrow) {
#pragma unroll
    for (int i = 0; i < NU; ++i) indptr[(I0 + lane) * NU + i] = (int64_t)NU * NU * bp + (int64_t)i * rl_mine * NU;
  }
  if (I0 + nrow == n_nodes && lane == nrow) {
    indptr[n_nodes * NU] = (int64_t)NU * NU * bp;
    scalars[kScalarNblk] = bp;                                   // node blocks of the call
  }
  {                                                              // max row length of the call: warp max, one atomic per warp
    int mx = rl_mine;
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    // look before the atomic: after the first few warps the maximum is already there (one same-address atomic per warp cost
    // Quad4 0.13 ms)
    if (lane == 0 && mx > 0 && (int64_t)mx > *reinterpret_cast<volatile int64_t *>(scalars + kScalarMaxRowlen))
      atomicMax(reinterpret_cast<unsigned long long *>(scalars + kScalarMaxRowlen), (unsigned long long)mx);
  }
  const bool over = lane < nrow && bp + rl_mine > cap_blk;
  if (__any_sync(0xffffffffu, over)) { if (lane == 0) atomicOr(guard, 1ull); return; } // hinted capacity exceeded: the call is repeated
  int32_t colreg[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int rl = __shfl_sync(0xffffffffu, rl_mine, r);
    const int npr = __shfl_sync(0xffffffffu, np, r);
    colreg[r] = (r < nrow && lane < rl) ? __ldg(tmpcols + (int64_t)nn * npr + lane) : 0;
  }
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int rl = __shfl_sync(0xffffffffu, rl_mine, r);
    const int64_t b0 = __shfl_sync(0xffffffffu, bp, r);
    if (r >= nrow || rl == 0) continue;
    if (lane < rl) blkcols[b0 + lane] = colreg[r];
    const int32_t cg = col_map ? __ldg(col_map + colreg[r]) : colreg[r];
    const int rowsz = rl * NU;
    int32_t *out = indices + (int64_t)NU * NU * b0;
    for (int base = 0; base < rowsz; base += 32) {               // warp-uniform trip count: every lane takes part in the shuffle
      const int idx = base + lane;
      const int p = idx / NU, uc = idx - p * NU;
      const int32_t c = __shfl_sync(0xffffffffu, cg, p & 31) * NU + uc;
      if (idx < rowsz) {
#pragma unroll
        for (int ur = 0; ur < NU; ++ur) out[ur * rowsz + idx] = c;
      }
    }
  }
}

template <int NN>
static int launch_row_pattern(femb200_ctx *ctx, int64_t n_nodes, const int32_t *n2e_ptr, uint32_t *n2e_t,
                              const int32_t *conn, int32_t *tmpcols, int64_t *rowlen, uint16_t *pos16, const uint8_t *node_mask) {
  const bool no_hash = getenv("FEMB200_NO_HASH_PATTERN") != nullptr;
  if (!no_hash) {
    constexpr int HS = NN <= 4 ? 32 : (NN <= 8 ? 64 : 128);
    // incidences per lane and block size.  Every incidence keeps nn connectivity entries in registers, so large elements
    // run a first pass with ONE incidence per lane (rows with <= 8 incident elements: every node of a structured
    // hexahedral grid, the mid-edge nodes of quadratic tetrahedra) in small blocks, and a second pass with four per lane
    // only over the rows the first one flagged (vertex nodes of tetrahedral meshes: ~24 incident elements).
    constexpr int T = NN >= 8 ? 128 : 256;
    auto launch = [&](auto ipl_c, int only_flagged) -> int {
      constexpr int IPL = decltype(ipl_c)::value;
      constexpr int GW = HS + HS + 8 * IPL + ((HS * 3 / 4 + 4 + 3) & ~3);
      const size_t smem = (size_t)(T / 8) * GW * sizeof(int32_t);
      auto kern = k_row_pattern_group<NN, HS, IPL, T>;
      FEMB_CUDA_OK(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      kern<<<grid_for(n_nodes * 8, T), T, smem, ctx->stream>>>(n_nodes, n2e_ptr, n2e_t, conn, tmpcols, rowlen, pos16, node_mask, only_flagged);
      ctx->launches++;
      return 0;
    };
    int rc;
    if constexpr (NN >= 8) {
      if ((rc = launch(std::integral_constant<int, 1>{}, 0))) return rc;
      if ((rc = launch(std::integral_constant<int, 4>{}, 1))) return rc;
    } else {
      if ((rc = launch(std::integral_constant<int, 4>{}, 0))) return rc;
    }
  }
  const int T = 128;
  // sorted-insert fallback: every row when the group path is disabled, else only the rows it flagged
  int cap_s = NN <= 4 ? 48 : (NN <= 8 ? 64 : 96);
  size_t smem = (size_t)cap_s * T * sizeof(int32_t);
  FEMB_CUDA_OK(ctx, cudaFuncSetAttribute(k_row_pattern<NN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_row_pattern<NN><<<grid_for(n_nodes, T), T, smem, ctx->stream>>>(n_nodes, n2e_ptr, n2e_t, conn, tmpcols, rowlen, pos16, cap_s, ctx->d_err, no_hash ? 0 : 1, node_mask);
  ctx->launches++;
  FEMB_CUDA_OK(ctx, cudaGetLastError());
  return 0;
}

static int ensure_cub_tmp(femb200_ctx *ctx, size_t bytes) {
  if (bytes <= ctx->cub_tmp_bytes) return 0;
  if (ctx->cub_tmp) cudaFreeAsync(ctx->cub_tmp, ctx->stream);
  ctx->cub_tmp = nullptr;
  ctx->cub_tmp_bytes = 0;
  FEMB_CUDA_OK(ctx, cudaMallocAsync(&ctx->cub_tmp, bytes, ctx->pool, ctx->stream));
  ctx->cub_tmp_bytes = bytes;
  return 0;
}

int symbolic_phase(femb200_assembler *ga, const ElemShape &sh, const int32_t *conn32, int64_t n_elem,
                   SymbolicResult *out, bool use_hint, int fork_at, const std::function<int()> *fork_fn) {
  femb200_ctx *ctx = ga->ctx;
  cudaStream_t st = ctx->stream;
  const int nn = sh.nn;
  const int64_t n_nodes = ga->n_nodes;
  const int64_t ninc = n_elem * nn;
  if (ninc >= (int64_t)1 << 31 || n_nodes >= ((int64_t)1 << 31) / (sh.mdpn > 0 ? sh.mdpn : 1)) return FEMB200_ERR_TOO_LARGE;
  int rc;

  // ---- node -> element adjacency -------------------------------------------------------------
  int32_t *cnt = out->precount;                                  // histogram taken while narrowing the connectivity (api.cu)?
  const bool counted = cnt != nullptr;
  if (!counted && (rc = scratch_alloc(ctx, &cnt, 2 * (n_nodes + 1)))) return rc; // histogram | fill cursors
  int32_t *cursor = cnt + (n_nodes + 1);
  if ((rc = dev_alloc(ctx, &out->n2e_ptr, n_nodes + 1))) return rc;
  if ((rc = dev_alloc(ctx, &out->n2e_t, ninc))) return rc;
  if (!counted) FEMB_CUDA_OK(ctx, cudaMemsetAsync(cnt, 0, sizeof(int32_t) * 2 * (n_nodes + 1), st));
  if (ninc && !counted) {
    if (nn == 4 && sh.d == 3) k_count_incidence<true><<<grid_for(ninc, 256), 256, 0, st>>>(conn32, ninc, cnt);
    else k_count_incidence<false><<<grid_for(ninc, 256), 256, 0, st>>>(conn32, ninc, cnt);
    ctx->launches++;
  }
  size_t tmp1 = 0, tmp3 = 0, tmp4 = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tmp1, cnt, out->n2e_ptr, (int)(n_nodes + 1), st);
  int64_t *rowlen = nullptr;
  if ((rc = scratch_alloc(ctx, &rowlen, n_nodes + 1))) return rc;
  if ((rc = dev_alloc(ctx, &out->blkptr, n_nodes + 1))) return rc;
  cub::DeviceScan::ExclusiveSum(nullptr, tmp3, rowlen, out->blkptr, (int)(n_nodes + 1), st);
  cub::DeviceReduce::Max(nullptr, tmp4, rowlen, ctx->d_scalars, (int)(n_nodes + 1), st);
  size_t tmpmax = tmp1 > tmp3 ? tmp1 : tmp3;
  if (tmp4 > tmpmax) tmpmax = tmp4;
  if ((rc = ensure_cub_tmp(ctx, tmpmax + 256))) return rc;
  size_t tb = ctx->cub_tmp_bytes;
  FEMB_CUDA_OK(ctx, cub::DeviceScan::ExclusiveSum(ctx->cub_tmp, tb, cnt, out->n2e_ptr, (int)(n_nodes + 1), st));
  ctx->launches += 2;
  if (ninc) {
    if (nn == 4 && sh.d == 3) k_fill_incidence<true><<<grid_for(ninc, 256), 256, 0, st>>>(conn32, ninc, out->n2e_ptr, cursor, out->n2e_t);
    else k_fill_incidence<false><<<grid_for(ninc, 256), 256, 0, st>>>(conn32, ninc, out->n2e_ptr, cursor, out->n2e_t);
    ctx->launches++;
  }
  FEMB_CUDA_OK(ctx, cudaEventRecord(ctx->ev[2], st));
  if (fork_at == 2 && fork_fn && (rc = (*fork_fn)())) return rc;

  // ---- node-block row pattern ----------------------------------------------------------------
  int32_t *tmpcols = nullptr;
  if ((rc = scratch_alloc(ctx, &tmpcols, ninc * nn))) return rc;
  out->rowcols = tmpcols; out->cols_stride = nn;
  if ((rc = dev_alloc(ctx, &out->pos16, ninc * nn + 64))) return rc; // + 64: the numeric kernel stages whole 64-byte chunks of slot words
  switch (nn) {
  case 3: rc = launch_row_pattern<3>(ctx, n_nodes, out->n2e_ptr, out->n2e_t, conn32, tmpcols, rowlen, out->pos16, ga->d_node_mask); break;
  case 4: rc = launch_row_pattern<4>(ctx, n_nodes, out->n2e_ptr, out->n2e_t, conn32, tmpcols, rowlen, out->pos16, ga->d_node_mask); break;
  case 8: rc = launch_row_pattern<8>(ctx, n_nodes, out->n2e_ptr, out->n2e_t, conn32, tmpcols, rowlen, out->pos16, ga->d_node_mask); break;
  case 10: rc = launch_row_pattern<10>(ctx, n_nodes, out->n2e_ptr, out->n2e_t, conn32, tmpcols, rowlen, out->pos16, ga->d_node_mask); break;
  case 20: rc = launch_row_pattern<20>(ctx, n_nodes, out->n2e_ptr, out->n2e_t, conn32, tmpcols, rowlen, out->pos16, ga->d_node_mask); break;
  default: rc = FEMB200_ERR_UNSUPPORTED;
  }
  if (rc) return rc;
  // per-group shared-memory need of the numeric phase (rows are visited in node order: a visiting order by first
  // incident element was measured and rejected -- it mixes rows with 5 and with 24 incident elements in one warp)
  if (n_nodes && (nn == 8 || nn == 20)) { // the element sizes whose numeric kernel uses packed accumulators
    k_group_need<<<grid_for(n_nodes, 256), 256, 0, st>>>(n_nodes, rowlen, (unsigned long long *)(ctx->d_scalars + 2));
    ctx->launches++;
  }
  tb = ctx->cub_tmp_bytes;
  FEMB_CUDA_OK(ctx, cub::DeviceScan::ExclusiveSum(ctx->cub_tmp, tb, rowlen, out->blkptr, (int)(n_nodes + 1), st));
  ctx->launches += 2;
  // a call sized from a hint needs no scalar before its end: k_expand8 leaves max(rowlen) and the block count in the
  // per-call scalars, which the caller reads back once; otherwise: CUB max + the small copies the host waits for
  const bool scalars_from_expand = use_hint && sh.identity && sh.nu <= 3 && ctx->hint.max_rowlen <= 32 && nn != 10 && !getenv("FEMB200_NO_EXPAND8");
  if (!scalars_from_expand) {
    tb = ctx->cub_tmp_bytes;
    FEMB_CUDA_OK(ctx, cub::DeviceReduce::Max(ctx->cub_tmp, tb, rowlen, ctx->d_scalars, (int)(n_nodes + 1), st));
    ctx->launches += 2;
  }
  if (nn == 10 && n_nodes && !getenv("FEMB200_ONE_CLASS")) { // two-class visiting order (see k_class_flags)
    uint8_t *flag = nullptr;
    if ((rc = scratch_alloc(ctx, &flag, n_nodes))) return rc;
    if ((rc = dev_alloc(ctx, &out->order, n_nodes))) return rc;
    k_class_flags<<<grid_for(n_nodes, 256), 256, 0, st>>>(n_nodes, rowlen, ctx->d_scalars, flag);
    size_t tmp6 = 0;
    cub::CountingInputIterator<int32_t> ids(0);
    cub::DevicePartition::Flagged(nullptr, tmp6, ids, flag, out->order, ctx->d_scalars + 6, (int)n_nodes, st);
    if ((rc = ensure_cub_tmp(ctx, tmp6 + 256))) return rc;
    tb = ctx->cub_tmp_bytes;
    FEMB_CUDA_OK(ctx, cub::DevicePartition::Flagged(ctx->cub_tmp, tb, ids, flag, out->order, ctx->d_scalars + 6, (int)n_nodes, st));
    ctx->launches += 3;
  }
  if (!use_hint) {
    // host needs nblk (allocation sizes), max row length (numeric launch shape) and the error word
    FEMB_CUDA_OK(ctx, cudaMemcpyAsync(ctx->h_scalars, ctx->d_scalars, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    FEMB_CUDA_OK(ctx, cudaMemcpyAsync(ctx->h_scalars + 1, out->blkptr + n_nodes, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    FEMB_CUDA_OK(ctx, cudaMemcpyAsync(ctx->h_scalars + 2, ctx->d_scalars + 2, 5 * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    FEMB_CUDA_OK(ctx, cudaMemcpyAsync(ctx->h_err, ctx->d_err, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
  } else if (!scalars_from_expand) {
    FEMB_CUDA_OK(ctx, cudaMemcpyAsync(ctx->d_scalars + kScalarNblk, out->blkptr + n_nodes, sizeof(int64_t), cudaMemcpyDeviceToDevice, st));
  }
  FEMB_CUDA_OK(ctx, cudaEventRecord(ctx->ev[3], st));
  if (fork_at == 3 && fork_fn && (rc = (*fork_fn)())) return rc;
  if (use_hint) {
    // sizes from the previous call with this mesh signature; verified by the caller after the call's only synchronisation
    out->speculative = true;
    out->max_rowlen = ctx->hint.max_rowlen;
    out->nblk = ctx->hint.nblk;
  } else {
    FEMB_CUDA_OK(ctx, cudaStreamSynchronize(st));
    out->max_rowlen = (int)ctx->h_scalars[0];
    out->nblk = ctx->h_scalars[1];
    for (int s = 0; s < 4; ++s) out->group_need[s] = n_nodes ? ctx->h_scalars[2 + s] : 0;
    if (out->order) out->n_short = ctx->h_scalars[6];
    if (*ctx->h_err != kNoError) return -1; // caller decodes the error word
    if (out->max_rowlen > (int)pos_mask(nn)) return FEMB200_ERR_TOO_LARGE;
  }
  out->cap_blk = out->nblk;

  // ---- scalar CSR pattern ----------------------------------------------------------------------
  const int nu = sh.nu;
  const int64_t nnz = (int64_t)nu * nu * out->nblk;
  const int64_t ndof = n_nodes * sh.mdpn;
  if ((int64_t)out->max_rowlen * nu >= ((int64_t)1 << 31)) return FEMB200_ERR_TOO_LARGE;
  out->csr.n_rows = ndof;
  out->csr.nnz = nnz;
  if ((rc = dev_alloc(ctx, &out->csr.indptr, ndof + 1))) return rc;
  if ((rc = dev_alloc(ctx, &out->csr.indices, nnz))) return rc;
  if ((rc = dev_alloc(ctx, &out->csr.values, nnz))) return rc;
  if ((rc = dev_alloc(ctx, &out->blkcols, (size_t)out->nblk))) return rc;
  unsigned umask = 0;
  for (int i = 0; i < nu; ++i) umask |= 1u << sh.umap[i];
  const int *d_umap = out->d_umap;                              // normally uploaded with the tables (api.cu)
  if (!d_umap) {
    int *um = nullptr;
    if ((rc = scratch_alloc(ctx, &um, kMaxNdn))) return rc;
    FEMB_CUDA_OK(ctx, cudaMemcpyAsync(um, sh.umap, sizeof(int) * kMaxNdn, cudaMemcpyHostToDevice, st));
    d_umap = um;
  }
  const bool expand8 = sh.identity && nu <= 3 && out->max_rowlen <= 32 && out->nblk > 0 && !getenv("FEMB200_NO_EXPAND8");
  if (expand8) {
    const unsigned eg = grid_for(((n_nodes + 7) / 8) * 32, 256);
    unsigned long long *guard = (unsigned long long *)(ctx->d_scalars + 7);
    // The numeric kernel of the fused path reads the row's column nodes from the pattern kernel's lists (rowcols), not from
    // blkcols, so the expansion (120 MB of scalar indices, HBM bound) has no consumer inside the call and CAN run on the second
    // stream beside the numeric kernel.  Measured (config 1): the step goes from 0.535 to 0.524 ms, but the numeric kernel
    // itself stretches from 0.232 to 0.272 ms -- the two kernels share SM slots rather than overlap -- so it is opt-in
    // (FEMB200_FORK_EXPAND=1) and the default keeps the kernels, and their roofline timings, apart.
    cudaStream_t es = st;
    if (ctx->stream2 && sh.nn <= 4 && getenv("FEMB200_FORK_EXPAND")) {
      FEMB_CUDA_OK(ctx, cudaEventRecord(ctx->ev_fork2, st));
      FEMB_CUDA_OK(ctx, cudaStreamWaitEvent(ctx->stream2, ctx->ev_fork2, 0));
      es = ctx->stream2;
      out->expand_forked = true;
    }
    switch (nu) {
    case 1: k_expand8<1><<<eg, 256, 0, es>>>(n_nodes, nn, out->n2e_ptr, out->blkptr, tmpcols, out->csr.indices, out->csr.indptr, ga->d_col_map, out->blkcols, out->cap_blk, guard, ctx->d_scalars); break;
    case 2: k_expand8<2><<<eg, 256, 0, es>>>(n_nodes, nn, out->n2e_ptr, out->blkptr, tmpcols, out->csr.indices, out->csr.indptr, ga->d_col_map, out->blkcols, out->cap_blk, guard, ctx->d_scalars); break;
    default: k_expand8<3><<<eg, 256, 0, es>>>(n_nodes, nn, out->n2e_ptr, out->blkptr, tmpcols, out->csr.indices, out->csr.indptr, ga->d_col_map, out->blkcols, out->cap_blk, guard, ctx->d_scalars); break;
    }
    ctx->launches++;
    if (out->expand_forked) FEMB_CUDA_OK(ctx, cudaEventRecord(ctx->ev_join2, ctx->stream2));
  } else {
  k_indptr<<<grid_for(ndof + 1, 256), 256, 0, st>>>(n_nodes, sh.mdpn, nu, umask, out->blkptr, out->csr.indptr);
  ctx->launches++;
  if (out->nblk) {
    const bool small = (int64_t)out->max_rowlen * nu <= 32;       // short rows: 8 lanes per node row, 4 rows per warp
    const unsigned eg = grid_for(n_nodes * (small ? 8 : 32), 256);
#define FEMB_EXPAND(NUV)                                                                                                          \
  do {                                                                                                                          \
    if (small) k_expand<NUV, 8><<<eg, 256, 0, st>>>(n_nodes, nn, sh.mdpn, nu, d_umap, out->n2e_ptr, out->blkptr, tmpcols, out->csr.indices, ga->d_col_map, out->blkcols, out->cap_blk, (unsigned long long *)(ctx->d_scalars + 7)); \
    else k_expand<NUV, 32><<<eg, 256, 0, st>>>(n_nodes, nn, sh.mdpn, nu, d_umap, out->n2e_ptr, out->blkptr, tmpcols, out->csr.indices, ga->d_col_map, out->blkcols, out->cap_blk, (unsigned long long *)(ctx->d_scalars + 7));   \
  } while (0)
    switch (nu) {
    case 1: FEMB_EXPAND(1); break;
    case 2: FEMB_EXPAND(2); break;
    case 3: FEMB_EXPAND(3); break;
    default: FEMB_EXPAND(0); break;
    }
#undef FEMB_EXPAND
    ctx->launches++;
  }
  }
  FEMB_CUDA_OK(ctx, cudaGetLastError());
  FEMB_CUDA_OK(ctx, cudaEventRecord(ctx->ev[4], st));
  return 0;
}

// ------------------------------------------------------------------------------------------------
// element -> slot map query (north_star subsystem 1): slot[e][a][b] = blkptr[node_a] + pos16
__global__ void k_slot_map(int64_t ninc, int nn, const int32_t *__restrict__ n2e_ptr_unused, const uint32_t *__restrict__ n2e_t,
                           const uint16_t *__restrict__ pos16, const int64_t *__restrict__ blkptr,
                           const int32_t *__restrict__ node_of_k, int64_t *__restrict__ slot) {
  int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= ninc) return;
  const uint32_t t = n2e_t[k]; // = e*nn + a
  const int64_t base = blkptr[node_of_k[k]];
  for (int b = 0; b < nn; ++b) slot[(int64_t)t * nn + b] = base + (pos16[k * nn + b] & pos_mask(nn));
}

__global__ void k_node_of_k(int64_t n_nodes, const int32_t *__restrict__ n2e_ptr, int32_t *__restrict__ node_of_k) {
  int64_t I = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (I >= n_nodes) return;
  for (int k = n2e_ptr[I]; k < n2e_ptr[I + 1]; ++k) node_of_k[k] = (int32_t)I;
}

int slot_map(femb200_assembler *ga, int64_t *d_slot) {
  femb200_ctx *ctx = ga->ctx;
  LastPlan &pl = ga->plan;
  if (!pl.n2e_t) return FEMB200_ERR_BAD_ARG;
  const int64_t ninc = pl.n_elem * pl.nn;
  int32_t *node_of_k = nullptr;
  int rc;
  if ((rc = dev_alloc(ctx, &node_of_k, ninc))) return rc;
  k_node_of_k<<<grid_for(ga->n_nodes, 256), 256, 0, ctx->stream>>>(ga->n_nodes, pl.n2e_ptr, node_of_k);
  if (ninc) k_slot_map<<<grid_for(ninc, 256), 256, 0, ctx->stream>>>(ninc, pl.nn, pl.n2e_ptr, pl.n2e_t, pl.pos16, pl.blkptr, node_of_k, d_slot);
  ctx->launches += 2;
  FEMB_CUDA_OK(ctx, cudaGetLastError());
  dev_free(ctx, node_of_k);
  return 0;
}

} // namespace femb200
