// dist.cu -- multi-GPU assembly behind the C ABI: device-side partition + owner-computes row blocks.
//
// Reference: one assembler, one call (assembler.go:21-31, assembler_isoparametric.go:28).  Here rank r of `world` (one
// process per GPU, or one host thread per GPU inside femb200_group_*) ends up owning a row block of the distributed CSR K
// in the caller's global DOF numbering (dof = mdpn*node + k, assembler.go:151-156).  Everything that used to be host numpy
// (go-fem_b200/distributed.py: morton_order, node_owner, ExchangePlan.halo_elems, LocalNumbering) runs on the device:
//
//   1. element order: given, or Morton (Z-curve) keys of the element centroids (21 bits per axis) + CUB radix sort
//   2. the rank's chunk = positions [ne*r/world, ne*(r+1)/world) of that order
//   3. owner[node]  = lowest rank with an element touching the node            (atomicMin over all incidences)
//   4. the rank's elements = every element with a node it owns (its own chunk's + the halo of higher ranks), kept in
//      ASCENDING ORIGINAL element index => each owned row is summed in the same order as on one GPU: bit-identical rows
//   5. rank-local node numbering (ascending global id) + gathered coordinates / connectivity / owned-row mask
//   6. the single-GPU path (femb200_add_isoparametric) over the local mesh with the node mask and the column map
//
// No data-path collective: the path shards into independent units (rows), the redundant work is the O(surface) halo.
// Every rank runs steps 1-5 over the WHOLE connectivity (1.6 GB for the 101 M-element mesh): no communication for the plan.
#include "common.cuh"
#include <cub/cub.cuh>
#include <thread>
#include <vector>

using namespace femb200;

struct femb200_dist {
  femb200_ctx *ctx = nullptr;
  int rank = 0, world = 1, nn = 0, mdpn = 0;
  int64_t n_nodes = 0, n_elem = 0;          // global mesh
  int64_t n_local_nodes = 0, n_owned_nodes = 0, n_local_elems = 0, n_own_chunk = 0, n_halo = 0;
  double *lnodes = nullptr;                 // [n_local_nodes][3]
  int32_t *lconn = nullptr;                 // [n_local_elems][nn] local node ids, ascending global element index
  int32_t *l2g = nullptr;                   // [n_local_nodes] global node of a local node (ascending)
  uint8_t *lmask = nullptr;                 // [n_local_nodes] 1 = this rank owns the node's rows
  int32_t *elist = nullptr;                 // [n_local_elems] global element index of a local element
  femb200_assembler *ga = nullptr;          // local assembler (rows local, columns global)
  float partition_ms = 0.f;
};

namespace {

__device__ __forceinline__ unsigned long long spread21(unsigned long long v) {
  v &= 0x1FFFFFull;
  v = (v | (v << 32)) & 0x1F00000000FFFFull;
  v = (v | (v << 16)) & 0x1F0000FF0000FFull;
  v = (v | (v << 8)) & 0x100F00F00F00F00Full;
  v = (v | (v << 4)) & 0x10C30C30C30C30C3ull;
  v = (v | (v << 2)) & 0x1249249249249249ull;
  return v;
}

// bounding box of the nodes: block reduction + ordered-integer atomics (doubles compare like their sign-folded bits)
__device__ __forceinline__ unsigned long long dkey(double x) {
  const unsigned long long b = (unsigned long long)__double_as_longlong(x);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double dunkey(unsigned long long k) {
  const unsigned long long b = (k >> 63) ? (k & 0x7FFFFFFFFFFFFFFFull) : ~k;
  return __longlong_as_double((long long)b);
}
__global__ void k_bbox(int64_t n, const double *__restrict__ nodes, unsigned long long *__restrict__ box /* [6] min xyz, max xyz (keys) */) {
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
  if (v < n)
    for (int j = 0; j < 3; ++j) lo[j] = hi[j] = nodes[3 * v + j];
  for (int j = 0; j < 3; ++j) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo[j] = fmin(lo[j], __shfl_xor_sync(0xffffffffu, lo[j], o));
      hi[j] = fmax(hi[j], __shfl_xor_sync(0xffffffffu, hi[j], o));
    }
    if ((threadIdx.x & 31) == 0) {
      atomicMin(box + j, dkey(lo[j]));
      atomicMax(box + 3 + j, dkey(hi[j]));
    }
  }
}

__global__ void k_morton_keys(int64_t ne, int nn, const int32_t *__restrict__ conn, const double *__restrict__ nodes,
                              const unsigned long long *__restrict__ box, unsigned long long *__restrict__ keys, int32_t *__restrict__ ids) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= ne) return;
  double c[3] = {0, 0, 0};
  for (int a = 0; a < nn; ++a) {
    const int64_t nd = conn[e * nn + a];
    for (int j = 0; j < 3; ++j) c[j] += nodes[3 * nd + j];
  }
  unsigned long long key = 0;
  for (int j = 0; j < 3; ++j) {
    const double lo = dunkey(box[j]), hi = dunkey(box[3 + j]);
    const double span = hi > lo ? hi - lo : 1.0;
    double q = (c[j] / nn - lo) / span * 2097151.0;
    q = fmin(fmax(q, 0.0), 2097151.0);
    key |= spread21((unsigned long long)q) << j;
  }
  keys[e] = key;
  ids[e] = (int32_t)e;
}

__global__ void k_iota(int64_t n, int32_t *__restrict__ a) {
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v < n) a[v] = (int32_t)v;
}

__device__ __forceinline__ int rank_of_pos(int64_t pos, int64_t ne, int world) {
  // chunk q = positions [ne*q/world, ne*(q+1)/world)  (integer division): q = ceil((pos+1)*world/ne) - 1
  return (int)(((pos + 1) * world + ne - 1) / ne - 1);
}

// owner[node] = lowest rank touching it; erank[e] = rank of element e's chunk
__global__ void k_owner(int64_t ne, int nn, int world, const int32_t *__restrict__ order, const int32_t *__restrict__ conn,
                        int32_t *__restrict__ owner, uint8_t *__restrict__ erank) {
  const int64_t pos = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (pos >= ne) return;
  const int64_t e = order[pos];
  const int q = rank_of_pos(pos, ne, world);
  erank[e] = (uint8_t)q;
  for (int a = 0; a < nn; ++a) atomicMin(owner + conn[e * nn + a], q);
}

// flag[e] = element has a node this rank owns; counts own-chunk / halo members; marks the nodes of flagged elements
__global__ void k_flag_elems(int64_t ne, int nn, int rank, const int32_t *__restrict__ conn, const int32_t *__restrict__ owner,
                             const uint8_t *__restrict__ erank, uint8_t *__restrict__ flag, uint8_t *__restrict__ mark,
                             unsigned long long *__restrict__ counts /* [0] own chunk (all), [1] halo (flagged, other chunk) */) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  bool mine = false, chunk = false, f = false;
  if (e < ne) {
    for (int a = 0; a < nn; ++a) mine = mine || (owner[conn[e * nn + a]] == rank);
    chunk = erank[e] == rank;
    f = mine;
    flag[e] = f ? 1 : 0;
    if (f)
      for (int a = 0; a < nn; ++a) mark[conn[e * nn + a]] = 1;
  }
  const unsigned bc = __ballot_sync(0xffffffffu, chunk), bh = __ballot_sync(0xffffffffu, f && !chunk);
  if ((threadIdx.x & 31) == 0) {
    if (bc) atomicAdd(counts, (unsigned long long)__popc(bc));
    if (bh) atomicAdd(counts + 1, (unsigned long long)__popc(bh));
  }
}

__global__ void k_local_nodes(int64_t n_local, const int32_t *__restrict__ l2g, const double *__restrict__ nodes, const int32_t *__restrict__ owner,
                              int rank, double *__restrict__ lnodes, uint8_t *__restrict__ lmask, unsigned long long *__restrict__ n_owned) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  bool own = false;
  if (i < n_local) {
    const int64_t g = l2g[i];
    for (int j = 0; j < 3; ++j) lnodes[3 * i + j] = nodes[3 * g + j];
    own = owner[g] == rank;
    lmask[i] = own ? 1 : 0;
  }
  const unsigned b = __ballot_sync(0xffffffffu, own);
  if ((threadIdx.x & 31) == 0 && b) atomicAdd(n_owned, (unsigned long long)__popc(b));
}

__global__ void k_local_conn(int64_t n_loc, int nn, const int32_t *__restrict__ elist, const int32_t *__restrict__ conn,
                             const int32_t *__restrict__ g2l, int32_t *__restrict__ lconn) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_loc * nn) return;
  const int64_t k = t / nn;
  const int a = (int)(t - k * nn);
  lconn[t] = g2l[conn[(int64_t)elist[k] * nn + a]];
}

// Device temporaries of the partition: bump-allocated from the context's persistent scratch arena and released as a whole.
// (They were ~18 stream-ordered pool allocations of different sizes per call; interleaved with the long-lived arrays of the
// assembler they kept the pool growing and re-mapping: identical partitions took 3 .. 230 ms by the host clock and the
// following assembly 1 .. 30 ms.)
struct Tmp {
  femb200_ctx *ctx;
  explicit Tmp(femb200_ctx *c) : ctx(c) {}
  template <class T> int get(T **out, size_t n) { return scratch_alloc(ctx, out, n); }
  ~Tmp() { scratch_release(ctx); }
};

} // namespace

extern "C" {

int femb200_dist_create(femb200_ctx *ctx, int32_t rank, int32_t world, const double *nodes, int32_t nodes_on_device, int64_t n_nodes,
                        int32_t mdpn, const void *conn, int32_t conn_is_int64, int32_t conn_on_device, int64_t n_elem, int32_t nn,
                        int32_t order_mode, femb200_dist **out) {
  if (!ctx || !out || world < 1 || world > 255 || rank < 0 || rank >= world || n_nodes < 0 || n_elem < 0 || nn < 1 || nn > kMaxNN || mdpn < 1 ||
      (n_nodes > 0 && !nodes) || (n_elem > 0 && !conn) || n_elem * nn >= ((int64_t)1 << 31) || n_nodes >= ((int64_t)1 << 31))
    return FEMB200_ERR_BAD_ARG;
  *out = nullptr;
  cudaStream_t st = ctx->stream;
  FEMB_CUDA_OK(ctx, cudaSetDevice(ctx->device));
  cudaEvent_t e0, e1;
  FEMB_CUDA_OK(ctx, cudaEventCreate(&e0));
  FEMB_CUDA_OK(ctx, cudaEventCreate(&e1));
  Tmp tmp(ctx);
  int rc;
  // ---- inputs on the device ------------------------------------------------------------------------------------------
  const double *d_nodes = nodes;
  if (!nodes_on_device) {
    double *dn = nullptr;
    if ((rc = tmp.get(&dn, (size_t)n_nodes * 3))) return rc;
    FEMB_CUDA_OK(ctx, cudaMemcpyAsync(dn, nodes, sizeof(double) * 3 * n_nodes, cudaMemcpyHostToDevice, st));
    d_nodes = dn;
  }
  const int64_t ninc = n_elem * nn;
  int32_t *conn32 = nullptr;
  if ((rc = tmp.get(&conn32, (size_t)ninc))) return rc;
  {
    const void *d_in = conn;
    if (!conn_on_device && ninc) {
      char *sb = nullptr;
      if ((rc = tmp.get(&sb, (size_t)ninc * (conn_is_int64 ? 8 : 4)))) return rc;
      FEMB_CUDA_OK(ctx, cudaMemcpyAsync(sb, conn, (size_t)ninc * (conn_is_int64 ? 8 : 4), cudaMemcpyHostToDevice, st));
      d_in = sb;
    }
    femb200_assembler probe; // only n_nodes / ctx are read by the validator
    probe.ctx = ctx; probe.n_nodes = n_nodes;
    *ctx->h_err = kNoError;
    if (int rcr = reset_call_scalars(ctx)) return rcr;
    if ((rc = narrow_and_validate_conn(&probe, d_in, conn_is_int64 != 0, ninc, nn, conn32))) return rc;
    FEMB_CUDA_OK(ctx, cudaMemcpyAsync(ctx->h_err, ctx->d_err, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    FEMB_CUDA_OK(ctx, cudaStreamSynchronize(st));
    if (*ctx->h_err != kNoError) return FEMB200_ERR_NODE_OOB; // the reference would panic on the slice bound (assembler.go:137)
  }
  FEMB_CUDA_OK(ctx, cudaEventRecord(e0, st));

  // ---- 1. element order ------------------------------------------------------------------------------------------------
  int32_t *order = nullptr;
  if ((rc = tmp.get(&order, (size_t)n_elem))) return rc;
  if (order_mode == 1 && n_elem) {
    unsigned long long *box = nullptr, *keys = nullptr, *keys2 = nullptr;
    int32_t *ids = nullptr;
    if ((rc = tmp.get(&box, 6))) return rc;
    if ((rc = tmp.get(&keys, (size_t)n_elem))) return rc;
    if ((rc = tmp.get(&keys2, (size_t)n_elem))) return rc;
    if ((rc = tmp.get(&ids, (size_t)n_elem))) return rc;
    FEMB_CUDA_OK(ctx, cudaMemsetAsync(box, 0xff, 3 * sizeof(unsigned long long), st));
    FEMB_CUDA_OK(ctx, cudaMemsetAsync(box + 3, 0, 3 * sizeof(unsigned long long), st));
    k_bbox<<<grid_for(n_nodes, 256), 256, 0, st>>>(n_nodes, d_nodes, box);
    k_morton_keys<<<grid_for(n_elem, 256), 256, 0, st>>>(n_elem, nn, conn32, d_nodes, box, keys, ids);
    size_t tb = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tb, keys, keys2, ids, order, (int)n_elem, 0, 63, st);
    char *stmp = nullptr;
    if ((rc = tmp.get(&stmp, tb + 256))) return rc;
    FEMB_CUDA_OK(ctx, cub::DeviceRadixSort::SortPairs(stmp, tb, keys, keys2, ids, order, (int)n_elem, 0, 63, st)); // stable: ties keep element order
    ctx->launches += 10;
  } else if (n_elem) {
    k_iota<<<grid_for(n_elem, 256), 256, 0, st>>>(n_elem, order);
    ctx->launches++;
  }

  // ---- 2./3. chunks and node ownership ---------------------------------------------------------------------------------
  int32_t *owner = nullptr, *g2l = nullptr;
  uint8_t *erank = nullptr, *flag = nullptr, *mark = nullptr;
  unsigned long long *counts = nullptr;
  if ((rc = tmp.get(&owner, (size_t)n_nodes + 1))) return rc;
  if ((rc = tmp.get(&g2l, (size_t)n_nodes + 1))) return rc;
  if ((rc = tmp.get(&erank, (size_t)n_elem + 1))) return rc;
  if ((rc = tmp.get(&flag, (size_t)n_elem + 1))) return rc;
  if ((rc = tmp.get(&mark, (size_t)n_nodes + 1))) return rc;
  if ((rc = tmp.get(&counts, 4))) return rc;
  FEMB_CUDA_OK(ctx, cudaMemsetAsync(owner, 0x7f, sizeof(int32_t) * (n_nodes + 1), st)); // > any rank: unreferenced nodes belong to nobody (empty rows)
  FEMB_CUDA_OK(ctx, cudaMemsetAsync(mark, 0, (size_t)n_nodes + 1, st));
  FEMB_CUDA_OK(ctx, cudaMemsetAsync(counts, 0, 4 * sizeof(unsigned long long), st));
  if (n_elem) {
    k_owner<<<grid_for(n_elem, 256), 256, 0, st>>>(n_elem, nn, world, order, conn32, owner, erank);
    // ---- 4. this rank's elements (ascending original index) and the nodes they touch ------------------------------------
    k_flag_elems<<<grid_for(n_elem, 256), 256, 0, st>>>(n_elem, nn, rank, conn32, owner, erank, flag, mark, counts);
    ctx->launches += 2;
  }
  femb200_dist *d = new femb200_dist();
  d->ctx = ctx; d->rank = rank; d->world = world; d->nn = nn; d->mdpn = mdpn; d->n_nodes = n_nodes; d->n_elem = n_elem;
  auto bail = [&](int code) { femb200_dist_destroy(d); return code; };
  // sizes: flagged elements, marked nodes
  int64_t *d_cnt = nullptr;
  if ((rc = tmp.get(&d_cnt, 2))) return bail(rc);
  int32_t *elist_full = nullptr, *l2g_full = nullptr;
  if ((rc = tmp.get(&elist_full, (size_t)n_elem + 1))) return bail(rc);
  if ((rc = tmp.get(&l2g_full, (size_t)n_nodes + 1))) return bail(rc);
  {
    cub::CountingInputIterator<int32_t> ids(0);
    size_t t1 = 0, t2 = 0, t3 = 0;
    cub::DeviceSelect::Flagged(nullptr, t1, ids, flag, elist_full, d_cnt, (int)n_elem, st);
    cub::DeviceSelect::Flagged(nullptr, t2, ids, mark, l2g_full, d_cnt + 1, (int)n_nodes, st);
    cub::DeviceScan::ExclusiveSum(nullptr, t3, mark, g2l, (int)(n_nodes + 1), st);
    size_t tb = t1 > t2 ? t1 : t2;
    if (t3 > tb) tb = t3;
    char *stmp = nullptr;
    if ((rc = tmp.get(&stmp, tb + 256))) return bail(rc);
    cudaMemsetAsync(d_cnt, 0, 2 * sizeof(int64_t), st);
    size_t a1 = tb + 256;
    if (cub::DeviceSelect::Flagged(stmp, a1, ids, flag, elist_full, d_cnt, (int)n_elem, st) != cudaSuccess) return bail(FEMB200_ERR_CUDA);
    a1 = tb + 256;
    if (cub::DeviceSelect::Flagged(stmp, a1, ids, mark, l2g_full, d_cnt + 1, (int)n_nodes, st) != cudaSuccess) return bail(FEMB200_ERR_CUDA);
    a1 = tb + 256;
    if (cub::DeviceScan::ExclusiveSum(stmp, a1, mark, g2l, (int)(n_nodes + 1), st) != cudaSuccess) return bail(FEMB200_ERR_CUDA);
    ctx->launches += 6;
  }
  int64_t h_cnt[2] = {0, 0};
  unsigned long long h_counts[4] = {0, 0, 0, 0};
  if (cudaMemcpyAsync(h_cnt, d_cnt, sizeof h_cnt, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
      cudaMemcpyAsync(h_counts, counts, sizeof h_counts, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
      cudaStreamSynchronize(st) != cudaSuccess)
    return bail(FEMB200_ERR_CUDA);
  d->n_local_elems = h_cnt[0];
  d->n_local_nodes = h_cnt[1];
  d->n_own_chunk = (int64_t)h_counts[0];
  d->n_halo = (int64_t)h_counts[1];

  // ---- 5. rank-local arrays ----------------------------------------------------------------------------------------------
  if ((rc = dev_alloc(ctx, &d->lnodes, (size_t)d->n_local_nodes * 3))) return bail(rc);
  if ((rc = dev_alloc(ctx, &d->lconn, (size_t)d->n_local_elems * nn))) return bail(rc);
  if ((rc = dev_alloc(ctx, &d->l2g, (size_t)d->n_local_nodes))) return bail(rc);
  if ((rc = dev_alloc(ctx, &d->lmask, (size_t)d->n_local_nodes))) return bail(rc);
  if ((rc = dev_alloc(ctx, &d->elist, (size_t)d->n_local_elems))) return bail(rc);
  unsigned long long *n_owned = counts + 2;
  if (d->n_local_nodes) {
    cudaMemcpyAsync(d->l2g, l2g_full, sizeof(int32_t) * d->n_local_nodes, cudaMemcpyDeviceToDevice, st);
    k_local_nodes<<<grid_for(d->n_local_nodes, 256), 256, 0, st>>>(d->n_local_nodes, d->l2g, d_nodes, owner, rank, d->lnodes, d->lmask, n_owned);
    ctx->launches++;
  }
  if (d->n_local_elems) {
    cudaMemcpyAsync(d->elist, elist_full, sizeof(int32_t) * d->n_local_elems, cudaMemcpyDeviceToDevice, st);
    k_local_conn<<<grid_for(d->n_local_elems * nn, 256), 256, 0, st>>>(d->n_local_elems, nn, d->elist, conn32, g2l, d->lconn);
    ctx->launches++;
  }
  cudaEventRecord(e1, st);
  if (cudaMemcpyAsync(h_counts, counts, sizeof h_counts, cudaMemcpyDeviceToHost, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess)
    return bail(FEMB200_ERR_CUDA);
  d->n_owned_nodes = (int64_t)h_counts[2];
  cudaEventElapsedTime(&d->partition_ms, e0, e1);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  if (cudaGetLastError() != cudaSuccess) return bail(FEMB200_ERR_CUDA);

  // ---- 6. the local assembler: rows local, column indices global ---------------------------------------------------------
  if ((rc = femb200_assembler_create(ctx, d->lnodes, d->n_local_nodes, mdpn, 1, &d->ga))) return bail(rc);
  femb200_assembler_set_node_mask_device(d->ga, d->lmask);
  if ((rc = femb200_assembler_set_column_map_device(d->ga, d->l2g, n_nodes))) return bail(rc);
  *out = d;
  return 0;
}

void femb200_dist_destroy(femb200_dist *d) {
  if (!d) return;
  femb200_ctx *ctx = d->ctx;
  cudaSetDevice(ctx->device);
  if (d->ga) femb200_assembler_destroy(d->ga);
  dev_free(ctx, d->lnodes); dev_free(ctx, d->lconn); dev_free(ctx, d->l2g); dev_free(ctx, d->lmask); dev_free(ctx, d->elist);
  cudaStreamSynchronize(ctx->stream);
  delete d;
}

int femb200_dist_add_isoparametric(femb200_dist *d, const femb200_elem_desc *desc) {
  if (!d || !desc) return FEMB200_ERR_BAD_ARG;
  if (desc->nn != d->nn) return FEMB200_ERR_BAD_ARG;
  return femb200_add_isoparametric(d->ga, desc, d->lconn, 0, 1, d->n_local_elems);
}

int femb200_dist_reset(femb200_dist *d) {
  if (!d) return FEMB200_ERR_BAD_ARG;
  if (d->ga) femb200_assembler_destroy(d->ga);
  d->ga = nullptr;
  int rc = femb200_assembler_create(d->ctx, d->lnodes, d->n_local_nodes, d->mdpn, 1, &d->ga);
  if (rc) return rc;
  femb200_assembler_set_node_mask_device(d->ga, d->lmask);
  return femb200_assembler_set_column_map_device(d->ga, d->l2g, d->n_nodes);
}

femb200_assembler *femb200_dist_assembler(femb200_dist *d) { return d ? d->ga : nullptr; }

int femb200_dist_info(const femb200_dist *d, int64_t out[8], float *partition_ms) {
  if (!d || !out) return FEMB200_ERR_BAD_ARG;
  out[0] = d->n_local_nodes; out[1] = d->n_owned_nodes; out[2] = d->n_local_elems; out[3] = d->n_own_chunk; out[4] = d->n_halo;
  out[5] = d->ga ? femb200_nnz(d->ga) : 0; out[6] = d->n_nodes; out[7] = d->n_elem;
  if (partition_ms) *partition_ms = d->partition_ms;
  return 0;
}

int femb200_dist_local_nodes(const femb200_dist *d, int32_t *l2g, uint8_t *owned) {
  if (!d) return FEMB200_ERR_BAD_ARG;
  femb200_ctx *ctx = d->ctx;
  FEMB_CUDA_OK(ctx, cudaSetDevice(ctx->device));
  if (l2g && d->n_local_nodes) FEMB_CUDA_OK(ctx, cudaMemcpyAsync(l2g, d->l2g, sizeof(int32_t) * d->n_local_nodes, cudaMemcpyDeviceToHost, ctx->stream));
  if (owned && d->n_local_nodes) FEMB_CUDA_OK(ctx, cudaMemcpyAsync(owned, d->lmask, d->n_local_nodes, cudaMemcpyDeviceToHost, ctx->stream));
  FEMB_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  return 0;
}

// Lowest failing element of the last femb200_dist_add_isoparametric in GLOBAL numbering (the local call reports a local index).
int femb200_dist_last_error(const femb200_dist *d, int32_t *kind, int64_t *elem, int32_t *qp) {
  if (!d || !d->ga) return FEMB200_ERR_BAD_ARG;
  int64_t le = -1;
  int rc = femb200_last_error(d->ga, kind, &le, qp);
  if (rc) return rc;
  if (elem) {
    *elem = le;
    if (le >= 0 && le < d->n_local_elems) {
      int32_t g = -1;
      femb200_ctx *ctx = d->ctx;
      FEMB_CUDA_OK(ctx, cudaSetDevice(ctx->device));
      FEMB_CUDA_OK(ctx, cudaMemcpy(&g, d->elist + le, sizeof g, cudaMemcpyDeviceToHost));
      *elem = g;
    }
  }
  return 0;
}

// ---- one process, several devices (what a cgo caller uses): one context + one host thread per device -------------------
struct femb200_group {
  std::vector<femb200_ctx *> ctx;
  std::vector<femb200_dist *> dist;
  std::vector<int> rc;
};

int femb200_group_create(const int32_t *devices, int32_t ndev, const double *nodes, int64_t n_nodes, int32_t mdpn, const void *conn,
                         int32_t conn_is_int64, int64_t n_elem, int32_t nn, int32_t order_mode, femb200_group **out) {
  if (!devices || ndev < 1 || ndev > 255 || !out) return FEMB200_ERR_BAD_ARG;
  *out = nullptr;
  femb200_group *g = new femb200_group();
  g->ctx.assign(ndev, nullptr);
  g->dist.assign(ndev, nullptr);
  g->rc.assign(ndev, 0);
  std::vector<std::thread> th;
  for (int r = 0; r < ndev; ++r)
    th.emplace_back([&, r]() {
      int rc = femb200_ctx_create(devices[r], &g->ctx[r]);
      bool shared = false; // the same device listed twice (tests): the per-device __constant__ tables must not be shared
      for (int q = 0; q < ndev; ++q) shared = shared || (q != r && devices[q] == devices[r]);
      if (!rc && shared) rc = femb200_ctx_set_concurrent(g->ctx[r], 1);
      if (!rc) rc = femb200_dist_create(g->ctx[r], r, ndev, nodes, 0, n_nodes, mdpn, conn, conn_is_int64, 0, n_elem, nn, order_mode, &g->dist[r]);
      g->rc[r] = rc;
    });
  for (auto &t : th) t.join();
  for (int r = 0; r < ndev; ++r)
    if (g->rc[r]) { int rc = g->rc[r]; femb200_group_destroy(g); return rc; }
  *out = g;
  return 0;
}

void femb200_group_destroy(femb200_group *g) {
  if (!g) return;
  for (auto *d : g->dist) femb200_dist_destroy(d);
  for (auto *c : g->ctx) femb200_ctx_destroy(c);
  delete g;
}

int32_t femb200_group_size(const femb200_group *g) { return g ? (int32_t)g->dist.size() : 0; }
femb200_dist *femb200_group_rank(femb200_group *g, int32_t r) { return (g && r >= 0 && r < (int32_t)g->dist.size()) ? g->dist[r] : nullptr; }

// AddIsoparametric on every device at once.  Returns 0, or the status of the rank that holds the LOWEST failing element
// (the sequential reference reports that one, fem.go:140); *rank_out tells which rank to ask for the details.
int femb200_group_add_isoparametric(femb200_group *g, const femb200_elem_desc *desc, int32_t *rank_out) {
  if (!g || !desc) return FEMB200_ERR_BAD_ARG;
  const int n = (int)g->dist.size();
  std::vector<std::thread> th;
  for (int r = 0; r < n; ++r) th.emplace_back([&, r]() { g->rc[r] = femb200_dist_add_isoparametric(g->dist[r], desc); });
  for (auto &t : th) t.join();
  int best = -1;
  int64_t best_elem = INT64_MAX;
  for (int r = 0; r < n; ++r)
    if (g->rc[r]) {
      int64_t e = -1;
      femb200_dist_last_error(g->dist[r], nullptr, &e, nullptr);
      if (best < 0 || (e >= 0 && e < best_elem)) { best = r; best_elem = e >= 0 ? e : best_elem; }
    }
  if (rank_out) *rank_out = best;
  return best < 0 ? 0 : g->rc[best];
}

} // extern "C"
