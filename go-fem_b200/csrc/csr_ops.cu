// csr_ops.cu -- operations on the device-resident K (canonical CSR).
//   csr_merge_into : K += A with pattern union  (repeated AddIsoparametric calls accumulate into the
//                    same lap.Sparse, assembler_isoparametric.go:121)
//   csr_matvec     : y = K x                   (forces.MulVec(K,u), constitution/solids/patch_test.go:92)
#include "common.cuh"
#include <cub/cub.cuh>
#include <algorithm>
#include <vector>

namespace femb200 {

// one thread per row: size of the union of two ascending index lists
__global__ void k_union_count(int64_t n_rows, const int64_t *__restrict__ pa, const int32_t *__restrict__ ia,
                              const int64_t *__restrict__ pb, const int32_t *__restrict__ ib, int64_t *__restrict__ cnt) {
  int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r > n_rows) return;
  if (r == n_rows) { cnt[r] = 0; return; }
  int64_t a = pa[r], ae = pa[r + 1], b = pb[r], be = pb[r + 1], c = 0;
  while (a < ae && b < be) {
    int32_t x = ia[a], y = ib[b];
    a += (x <= y);
    b += (y <= x);
    ++c;
  }
  cnt[r] = c + (ae - a) + (be - b);
}

// values: old + new where both exist (call order: existing K first, then this call's sum)
__global__ void k_union_fill(int64_t n_rows, const int64_t *__restrict__ pa, const int32_t *__restrict__ ia,
                             const double *__restrict__ va, const int64_t *__restrict__ pb, const int32_t *__restrict__ ib,
                             const double *__restrict__ vb, const int64_t *__restrict__ pc, int32_t *__restrict__ ic,
                             double *__restrict__ vc) {
  int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rows) return;
  int64_t a = pa[r], ae = pa[r + 1], b = pb[r], be = pb[r + 1], o = pc[r];
  while (a < ae || b < be) {
    int32_t x = a < ae ? ia[a] : INT32_MAX, y = b < be ? ib[b] : INT32_MAX;
    if (x == y) { ic[o] = x; vc[o] = va[a] + vb[b]; ++a; ++b; }
    else if (x < y) { ic[o] = x; vc[o] = va[a]; ++a; }
    else { ic[o] = y; vc[o] = vb[b]; ++b; }
    ++o;
  }
}

int csr_merge_into(femb200_assembler *ga, DeviceCsr *add) {
  femb200_ctx *ctx = ga->ctx;
  cudaStream_t st = ctx->stream;
  DeviceCsr &K = ga->K;
  if (K.nnz == 0) { // first call: adopt
    dev_free(ctx, K.indptr); dev_free(ctx, K.indices); dev_free(ctx, K.values);
    K = *add;
    *add = DeviceCsr();
    return 0;
  }
  const int64_t n = K.n_rows;
  int rc;
  DeviceCsr out;
  out.n_rows = n;
  int64_t *cnt = nullptr;
  if ((rc = dev_alloc(ctx, &cnt, n + 1))) return rc;
  if ((rc = dev_alloc(ctx, &out.indptr, n + 1))) return rc;
  k_union_count<<<grid_for(n + 1, 128), 128, 0, st>>>(n, K.indptr, K.indices, add->indptr, add->indices, cnt);
  size_t tb = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tb, cnt, out.indptr, (int)(n + 1), st);
  if (tb > ctx->cub_tmp_bytes) {
    if (ctx->cub_tmp) cudaFreeAsync(ctx->cub_tmp, st);
    ctx->cub_tmp = nullptr; ctx->cub_tmp_bytes = 0;
    FEMB_CUDA_OK(ctx, cudaMallocAsync(&ctx->cub_tmp, tb + 256, ctx->pool, st));
    ctx->cub_tmp_bytes = tb + 256;
  }
  tb = ctx->cub_tmp_bytes;
  FEMB_CUDA_OK(ctx, cub::DeviceScan::ExclusiveSum(ctx->cub_tmp, tb, cnt, out.indptr, (int)(n + 1), st));
  FEMB_CUDA_OK(ctx, cudaMemcpyAsync(ctx->h_scalars, out.indptr + n, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  FEMB_CUDA_OK(ctx, cudaStreamSynchronize(st));
  out.nnz = ctx->h_scalars[0];
  if ((rc = dev_alloc(ctx, &out.indices, out.nnz))) return rc;
  if ((rc = dev_alloc(ctx, &out.values, out.nnz))) return rc;
  k_union_fill<<<grid_for(n, 128), 128, 0, st>>>(n, K.indptr, K.indices, K.values, add->indptr, add->indices, add->values,
                                                 out.indptr, out.indices, out.values);
  ctx->launches += 4;
  FEMB_CUDA_OK(ctx, cudaGetLastError());
  dev_free(ctx, cnt);
  dev_free(ctx, K.indptr); dev_free(ctx, K.indices); dev_free(ctx, K.values);
  dev_free(ctx, add->indptr); dev_free(ctx, add->indices); dev_free(ctx, add->values);
  *add = DeviceCsr();
  K = out;
  ga->plan.matches_K = false; // K's pattern is now a union: the cached symbolic plan no longer addresses its slots
  return 0;
}

// K += triplets (femb200_add_coo; assembler.go:225-231).  The triplets come from the host (a few 12x12 beam blocks in the
// reference's use): they are ordered by (row, col) with a stable sort, duplicates summed in input order, and the resulting
// small CSR is merged into K on the device by the same pattern-union kernels repeated AddIsoparametric calls use.
int csr_add_coo(femb200_assembler *ga, int64_t n, const int64_t *h_rows, const int64_t *h_cols, const double *h_vals) {
  femb200_ctx *ctx = ga->ctx;
  cudaStream_t st = ctx->stream;
  const int64_t nr = ga->K.n_rows;
  std::vector<int64_t> perm(n);
  for (int64_t k = 0; k < n; ++k) perm[k] = k;
  std::stable_sort(perm.begin(), perm.end(), [&](int64_t a, int64_t b) {
    return h_rows[a] != h_rows[b] ? h_rows[a] < h_rows[b] : h_cols[a] < h_cols[b];
  });
  std::vector<int64_t> indptr(nr + 1, 0);
  std::vector<int32_t> indices;
  std::vector<double> values;
  indices.reserve(n); values.reserve(n);
  for (int64_t k = 0; k < n; ++k) {
    const int64_t t = perm[k];
    if (k > 0 && h_rows[perm[k - 1]] == h_rows[t] && h_cols[perm[k - 1]] == h_cols[t]) { values.back() += h_vals[t]; continue; }
    indices.push_back((int32_t)h_cols[t]);
    values.push_back(h_vals[t]);
    indptr[h_rows[t] + 1]++;
  }
  for (int64_t r = 0; r < nr; ++r) indptr[r + 1] += indptr[r];
  DeviceCsr add;
  add.n_rows = nr;
  add.nnz = (int64_t)indices.size();
  int rc;
  if ((rc = dev_alloc(ctx, &add.indptr, (size_t)nr + 1))) return rc;
  if ((rc = dev_alloc(ctx, &add.indices, (size_t)add.nnz))) return rc;
  if ((rc = dev_alloc(ctx, &add.values, (size_t)add.nnz))) return rc;
  FEMB_CUDA_OK(ctx, cudaMemcpyAsync(add.indptr, indptr.data(), sizeof(int64_t) * (nr + 1), cudaMemcpyHostToDevice, st));
  FEMB_CUDA_OK(ctx, cudaMemcpyAsync(add.indices, indices.data(), sizeof(int32_t) * add.nnz, cudaMemcpyHostToDevice, st));
  FEMB_CUDA_OK(ctx, cudaMemcpyAsync(add.values, values.data(), sizeof(double) * add.nnz, cudaMemcpyHostToDevice, st));
  FEMB_CUDA_OK(ctx, cudaStreamSynchronize(st)); // the host vectors die with this frame
  ga->plan.matches_K = false; // foreign entries: the cached symbolic plan no longer describes K
  return csr_merge_into(ga, &add);
}

// ---- multi-GPU pack / unpack ----------------------------------------------------------------------------
__global__ void k_scatter_lens(int64_t n, const int64_t *__restrict__ rows, const int64_t *__restrict__ indptr, int64_t *__restrict__ len) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) len[rows[i]] = indptr[i + 1] - indptr[i];
}

int csr_accumulate_rows(femb200_assembler *ga, int64_t n_rows, const int64_t *rows, const int64_t *indptr, const int32_t *indices,
                        const double *values) {
  femb200_ctx *ctx = ga->ctx;
  cudaStream_t st = ctx->stream;
  if (n_rows == 0) return 0;
  const int64_t n = ga->K.n_rows;
  int rc;
  DeviceCsr add;
  add.n_rows = n;
  int64_t *len = nullptr;
  if ((rc = dev_alloc(ctx, &len, n + 1))) return rc;
  if ((rc = dev_alloc(ctx, &add.indptr, n + 1))) return rc;
  FEMB_CUDA_OK(ctx, cudaMemsetAsync(len, 0, sizeof(int64_t) * (n + 1), st));
  k_scatter_lens<<<grid_for(n_rows, 256), 256, 0, st>>>(n_rows, rows, indptr, len);
  size_t tb = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tb, len, add.indptr, (int)(n + 1), st);
  if (tb > ctx->cub_tmp_bytes) {
    if (ctx->cub_tmp) cudaFreeAsync(ctx->cub_tmp, st);
    ctx->cub_tmp = nullptr; ctx->cub_tmp_bytes = 0;
    FEMB_CUDA_OK(ctx, cudaMallocAsync(&ctx->cub_tmp, tb + 256, ctx->pool, st));
    ctx->cub_tmp_bytes = tb + 256;
  }
  tb = ctx->cub_tmp_bytes;
  FEMB_CUDA_OK(ctx, cub::DeviceScan::ExclusiveSum(ctx->cub_tmp, tb, len, add.indptr, (int)(n + 1), st));
  FEMB_CUDA_OK(ctx, cudaMemcpyAsync(ctx->h_scalars, indptr + n_rows, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  FEMB_CUDA_OK(ctx, cudaStreamSynchronize(st));
  add.nnz = ctx->h_scalars[0];
  if ((rc = dev_alloc(ctx, &add.indices, add.nnz))) return rc;
  if ((rc = dev_alloc(ctx, &add.values, add.nnz))) return rc;
  // rows ascending => the packed order is already the CSR order of the embedded matrix
  FEMB_CUDA_OK(ctx, cudaMemcpyAsync(add.indices, indices, sizeof(int32_t) * add.nnz, cudaMemcpyDeviceToDevice, st));
  FEMB_CUDA_OK(ctx, cudaMemcpyAsync(add.values, values, sizeof(double) * add.nnz, cudaMemcpyDeviceToDevice, st));
  ctx->launches += 3;
  dev_free(ctx, len);
  if (ga->K.nnz == 0 && add.nnz == 0) { dev_free(ctx, add.indptr); dev_free(ctx, add.indices); dev_free(ctx, add.values); return 0; }
  ga->plan.matches_K = false; // foreign rows enter K (also when K was empty and simply adopts them)
  return csr_merge_into(ga, &add);
}

// in-place add of received interface rows: warp per received row, one binary search per entry
__global__ void k_add_entries(int64_t n_rows, const int64_t *__restrict__ rows, const int64_t *__restrict__ ptr0,
                              const int32_t *__restrict__ idx, const double *__restrict__ val, const int64_t *__restrict__ kptr,
                              const int32_t *__restrict__ kidx, double *__restrict__ kval, unsigned long long *err, int check_only) {
  const int lane = threadIdx.x & 31;
  const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (w >= n_rows) return;
  const int64_t r = rows[w];
  const int64_t lo0 = kptr[r], hi0 = kptr[r + 1];
  for (int64_t k = ptr0[w] + lane; k < ptr0[w + 1]; k += 32) {
    const int32_t c = idx[k];
    int64_t lo = lo0, hi = hi0;
    while (lo < hi) {
      const int64_t mid = (lo + hi) >> 1;
      if (kidx[mid] < c) lo = mid + 1; else hi = mid;
    }
    if (lo < hi0 && kidx[lo] == c) { if (!check_only) kval[lo] += val[k]; }   // (row, col) unique within a piece: no two lanes share a slot
    else atomicMin(err, (unsigned long long)r);
  }
}

int csr_add_entries(femb200_assembler *ga, int64_t n_rows, const int64_t *rows, const int64_t *indptr, const int32_t *indices,
                    const double *values) {
  femb200_ctx *ctx = ga->ctx;
  cudaStream_t st = ctx->stream;
  if (n_rows == 0) return 0;
  if (ga->K.nnz == 0) return FEMB200_ERR_PATTERN;
  *ctx->h_err = kNoError;
  if (int rcr = reset_call_scalars(ctx)) return rcr;
  // pass 1 checks that every entry is in the pattern (K untouched on failure), pass 2 adds
  k_add_entries<<<grid_for(n_rows * 32, 256), 256, 0, st>>>(n_rows, rows, indptr, indices, values, ga->K.indptr, ga->K.indices, ga->K.values, ctx->d_err, 1);
  ctx->launches++;
  FEMB_CUDA_OK(ctx, cudaMemcpyAsync(ctx->h_err, ctx->d_err, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
  FEMB_CUDA_OK(ctx, cudaStreamSynchronize(st));
  if (*ctx->h_err != kNoError) return FEMB200_ERR_PATTERN;
  k_add_entries<<<grid_for(n_rows * 32, 256), 256, 0, st>>>(n_rows, rows, indptr, indices, values, ga->K.indptr, ga->K.indices, ga->K.values, ctx->d_err, 0);
  ctx->launches++;
  FEMB_CUDA_OK(ctx, cudaStreamSynchronize(st));
  return 0;
}

// ---- planned interface exchange (distributed.py): asynchronous pack and add, no host synchronisation ----------
// 8 lanes per row: copies row rows[w] of K into the packed arrays at the offsets the partition plan fixed; a row
// whose length differs from the plan bumps *mismatch and is truncated / zero padded.
__global__ void k_pack_rows(int64_t n_rows, const int64_t *__restrict__ rows, const int64_t *__restrict__ optr,
                            const int64_t *__restrict__ kptr, const int32_t *__restrict__ kidx, const double *__restrict__ kval,
                            int32_t *__restrict__ oidx, double *__restrict__ oval, unsigned long long *mismatch) {
  const int l = threadIdx.x & 7;
  const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
  if (w >= n_rows) return;
  const int64_t r = rows[w];
  const int64_t s = kptr[r], m = kptr[r + 1] - s, d = optr[w], cap = optr[w + 1] - d;
  if (m != cap && l == 0 && mismatch) atomicAdd(mismatch, 1ull);
  for (int64_t k = l; k < cap; k += 8) {
    oidx[d + k] = k < m ? kidx[s + k] : -1;
    oval[d + k] = k < m ? kval[s + k] : 0.0;
  }
}

__global__ void k_add_entries_async(int64_t n_rows, const int64_t *__restrict__ rows, const int64_t *__restrict__ ptr0,
                                    const int32_t *__restrict__ idx, const double *__restrict__ val, const int64_t *__restrict__ kptr,
                                    const int32_t *__restrict__ kidx, double *__restrict__ kval, unsigned long long *mismatch) {
  const int l = threadIdx.x & 7;
  const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
  if (w >= n_rows) return;
  const int64_t r = rows[w];
  const int64_t lo0 = kptr[r], hi0 = kptr[r + 1];
  for (int64_t k = ptr0[w] + l; k < ptr0[w + 1]; k += 8) {
    const int32_t c = idx[k];
    int64_t lo = lo0, hi = hi0;
    while (lo < hi) {
      const int64_t mid = (lo + hi) >> 1;
      if (kidx[mid] < c) lo = mid + 1; else hi = mid;
    }
    if (lo < hi0 && kidx[lo] == c) kval[lo] += val[k];   // (row, col) unique within a piece: no two lanes share a slot
    else if (mismatch) atomicAdd(mismatch, 1ull);
  }
}

int csr_pack_rows_async(femb200_assembler *ga, int64_t n_rows, const int64_t *rows, const int64_t *out_indptr, int32_t *out_indices,
                        double *out_values, unsigned long long *mismatch) {
  femb200_ctx *ctx = ga->ctx;
  if (n_rows == 0) return 0;
  k_pack_rows<<<grid_for(n_rows * 8, 256), 256, 0, ctx->stream>>>(n_rows, rows, out_indptr, ga->K.indptr, ga->K.indices, ga->K.values,
                                                                   out_indices, out_values, mismatch);
  ctx->launches++;
  FEMB_CUDA_OK(ctx, cudaGetLastError());
  return 0;
}

int csr_add_entries_async(femb200_assembler *ga, int64_t n_rows, const int64_t *rows, const int64_t *indptr, const int32_t *indices,
                          const double *values, unsigned long long *mismatch) {
  femb200_ctx *ctx = ga->ctx;
  if (n_rows == 0) return 0;
  k_add_entries_async<<<grid_for(n_rows * 8, 256), 256, 0, ctx->stream>>>(n_rows, rows, indptr, indices, values, ga->K.indptr, ga->K.indices,
                                                                           ga->K.values, mismatch);
  ctx->launches++;
  FEMB_CUDA_OK(ctx, cudaGetLastError());
  return 0;
}

__global__ void k_keep_lens(int64_t n, const int64_t *__restrict__ ptr, const uint8_t *__restrict__ keep, int64_t *__restrict__ len) {
  int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r > n) return;
  len[r] = (r < n && keep[r]) ? ptr[r + 1] - ptr[r] : 0;
}
// warp per row copy
__global__ void k_keep_copy(int64_t n, const int64_t *__restrict__ ptr, const int32_t *__restrict__ idx, const double *__restrict__ val,
                            const uint8_t *__restrict__ keep, const int64_t *__restrict__ nptr, int32_t *__restrict__ nidx, double *__restrict__ nval) {
  const int lane = threadIdx.x & 31;
  const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (r >= n || !keep[r]) return;
  const int64_t s = ptr[r], d = nptr[r], m = ptr[r + 1] - s;
  for (int64_t k = lane; k < m; k += 32) { nidx[d + k] = idx[s + k]; nval[d + k] = val[s + k]; }
}

int csr_keep_rows(femb200_assembler *ga, const uint8_t *keep) {
  femb200_ctx *ctx = ga->ctx;
  cudaStream_t st = ctx->stream;
  DeviceCsr &K = ga->K;
  const int64_t n = K.n_rows;
  if (K.nnz == 0 || n == 0) return 0;
  int rc;
  DeviceCsr out;
  out.n_rows = n;
  int64_t *len = nullptr;
  if ((rc = dev_alloc(ctx, &len, n + 1))) return rc;
  if ((rc = dev_alloc(ctx, &out.indptr, n + 1))) return rc;
  k_keep_lens<<<grid_for(n + 1, 256), 256, 0, st>>>(n, K.indptr, keep, len);
  size_t tb = ctx->cub_tmp_bytes;
  size_t need = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, need, len, out.indptr, (int)(n + 1), st);
  if (need > ctx->cub_tmp_bytes) {
    if (ctx->cub_tmp) cudaFreeAsync(ctx->cub_tmp, st);
    ctx->cub_tmp = nullptr; ctx->cub_tmp_bytes = 0;
    FEMB_CUDA_OK(ctx, cudaMallocAsync(&ctx->cub_tmp, need + 256, ctx->pool, st));
    ctx->cub_tmp_bytes = need + 256;
  }
  tb = ctx->cub_tmp_bytes;
  FEMB_CUDA_OK(ctx, cub::DeviceScan::ExclusiveSum(ctx->cub_tmp, tb, len, out.indptr, (int)(n + 1), st));
  FEMB_CUDA_OK(ctx, cudaMemcpyAsync(ctx->h_scalars, out.indptr + n, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  FEMB_CUDA_OK(ctx, cudaStreamSynchronize(st));
  out.nnz = ctx->h_scalars[0];
  if ((rc = dev_alloc(ctx, &out.indices, out.nnz))) return rc;
  if ((rc = dev_alloc(ctx, &out.values, out.nnz))) return rc;
  k_keep_copy<<<grid_for(n * 32, 256), 256, 0, st>>>(n, K.indptr, K.indices, K.values, keep, out.indptr, out.indices, out.values);
  ctx->launches += 4;
  FEMB_CUDA_OK(ctx, cudaGetLastError());
  dev_free(ctx, len);
  dev_free(ctx, K.indptr); dev_free(ctx, K.indices); dev_free(ctx, K.values);
  K = out;
  ga->plan.matches_K = false; // rows were dropped: the plan's block offsets are stale
  return 0;
}

// warp per row, shuffle reduction (fixed lane order => deterministic)
__global__ void k_matvec(int64_t n_rows, const int64_t *__restrict__ ptr, const int32_t *__restrict__ idx,
                         const double *__restrict__ val, const double *__restrict__ x, double *__restrict__ y) {
  const int lane = threadIdx.x & 31;
  const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (r >= n_rows) return;
  double s = 0.0;
  for (int64_t k = ptr[r] + lane; k < ptr[r + 1]; k += 32) s = fma(val[k], x[idx[k]], s);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
  if (lane == 0) y[r] = s;
}

int csr_matvec(femb200_assembler *ga, const double *d_x, double *d_y) {
  femb200_ctx *ctx = ga->ctx;
  const int64_t n = ga->K.n_rows;
  if (n == 0) return 0;
  if (ga->K.nnz == 0) {
    FEMB_CUDA_OK(ctx, cudaMemsetAsync(d_y, 0, sizeof(double) * n, ctx->stream));
    return 0;
  }
  k_matvec<<<grid_for(n * 32, 256), 256, 0, ctx->stream>>>(n, ga->K.indptr, ga->K.indices, ga->K.values, d_x, d_y);
  ctx->launches++;
  FEMB_CUDA_OK(ctx, cudaGetLastError());
  return 0;
}


// ---- roofline probes (bench.py): FP64 vector peak and HBM copy bandwidth ----------------------------
__global__ void __launch_bounds__(256) k_probe_dfma(double *out, int iters, double seed) {
  double a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = seed + i + threadIdx.x;
  const double m = 1.0000001, c = 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u)
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = fma(a[i], m, c);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i];
  if (s == 12345.678) out[blockIdx.x * blockDim.x + threadIdx.x] = s; // never true: keeps the chains alive
}

__global__ void __launch_bounds__(256) k_probe_copy(const double4 *__restrict__ in, double4 *__restrict__ out, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) out[i] = in[i];
}

int probe_peaks(femb200_ctx *ctx, int reps, double out[2]) {
  cudaStream_t st = ctx->stream;
  if (reps < 1) reps = 1;
  double *buf = nullptr;
  const size_t n_d = (size_t)1 << 27; // 2 x 512 MiB halves
  int rc = dev_alloc(ctx, &buf, n_d);
  if (rc) return rc;
  FEMB_CUDA_OK(ctx, cudaMemsetAsync(buf, 0, n_d * sizeof(double), st));
  const int iters = 4096, blocks = ctx->sm_count * 8;
  float best_f = 1e30f, best_c = 1e30f, ms = 0;
  for (int r = 0; r < reps + 1; ++r) {
    FEMB_CUDA_OK(ctx, cudaEventRecord(ctx->ev[7], st));
    k_probe_dfma<<<blocks, 256, 0, st>>>(buf, iters, 1.0);
    FEMB_CUDA_OK(ctx, cudaEventRecord(ctx->ev[8], st));
    FEMB_CUDA_OK(ctx, cudaStreamSynchronize(st));
    FEMB_CUDA_OK(ctx, cudaEventElapsedTime(&ms, ctx->ev[7], ctx->ev[8]));
    if (r && ms < best_f) best_f = ms;
    FEMB_CUDA_OK(ctx, cudaEventRecord(ctx->ev[7], st));
    k_probe_copy<<<ctx->sm_count * 16, 256, 0, st>>>((const double4 *)buf, (double4 *)(buf + n_d / 2), (int64_t)(n_d / 8));
    FEMB_CUDA_OK(ctx, cudaEventRecord(ctx->ev[8], st));
    FEMB_CUDA_OK(ctx, cudaStreamSynchronize(st));
    FEMB_CUDA_OK(ctx, cudaEventElapsedTime(&ms, ctx->ev[7], ctx->ev[8]));
    if (r && ms < best_c) best_c = ms;
  }
  ctx->launches += 2 * (reps + 1);
  out[0] = 2.0 * 64.0 * iters * (double)blocks * 256.0 / (best_f * 1e-3) / 1e12;
  out[1] = (double)(n_d * sizeof(double)) / (best_c * 1e-3) / 1e9; // n_d/2 doubles read + n_d/2 written
  dev_free(ctx, buf);
  return 0;
}


// ---- K(rows, cols) extraction (lap.Slice(K, free, free) of the patch tests; Fixity.FreeDofs / FixedDofs) -------------
__global__ void k_slice_colmap(int64_t n_cols, const int64_t *__restrict__ cols, int32_t *__restrict__ colmap) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c < n_cols) colmap[cols[c]] = (int32_t)c;
}
// 8 lanes per selected row: count (fill == 0) or write (fill == 1) the entries whose column is selected
__global__ void k_slice_rows(int64_t n_rows, const int64_t *__restrict__ rows, const int32_t *__restrict__ colmap,
                             const int64_t *__restrict__ kptr, const int32_t *__restrict__ kidx, const double *__restrict__ kval,
                             int64_t *__restrict__ cnt, const int64_t *__restrict__ optr, int32_t *__restrict__ oidx,
                             double *__restrict__ oval, int fill) {
  const int l = threadIdx.x & 7;
  const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
  if (w >= n_rows) return;
  const unsigned gmask = 0xffu << ((threadIdx.x & 31) & ~7);
  const int64_t r = rows[w];
  const int64_t lo = kptr[r], hi = kptr[r + 1];
  int64_t base = fill ? optr[w] : 0;
  int64_t total = 0;
  for (int64_t k0 = lo; k0 < hi; k0 += 8) { // ordered compaction: ballot + popc keeps the columns ascending
    const int64_t k = k0 + l;
    const int32_t nc = k < hi ? colmap[kidx[k]] : -1;
    const unsigned bal = (__ballot_sync(gmask, nc >= 0) >> ((threadIdx.x & 31) & ~7)) & 0xffu;
    if (fill && nc >= 0) {
      const int64_t o = base + __popc(bal & ((1u << l) - 1u));
      oidx[o] = nc;
      oval[o] = kval[k];
    }
    base += __popc(bal);
    total += __popc(bal);
  }
  if (!fill && l == 0) cnt[w] = total;
}

int csr_slice(femb200_assembler *ga, int64_t n_rows, const int64_t *h_rows, int64_t n_cols, const int64_t *h_cols, int64_t *h_indptr,
              int32_t *h_indices, double *h_values, int64_t *nnz_out) {
  femb200_ctx *ctx = ga->ctx;
  cudaStream_t st = ctx->stream;
  const DeviceCsr &K = ga->K;
  ScratchScope scope(ctx);
  int64_t *d_rows = nullptr, *d_cols = nullptr, *cnt = nullptr, *optr = nullptr;
  int32_t *colmap = nullptr;
  int rc;
  if ((rc = scratch_alloc(ctx, &d_rows, n_rows))) return rc;
  if ((rc = scratch_alloc(ctx, &d_cols, n_cols))) return rc;
  if ((rc = scratch_alloc(ctx, &cnt, n_rows + 1))) return rc;
  if ((rc = scratch_alloc(ctx, &optr, n_rows + 1))) return rc;
  if ((rc = scratch_alloc(ctx, &colmap, K.n_rows))) return rc;
  if (n_rows) FEMB_CUDA_OK(ctx, cudaMemcpyAsync(d_rows, h_rows, sizeof(int64_t) * n_rows, cudaMemcpyHostToDevice, st));
  if (n_cols) FEMB_CUDA_OK(ctx, cudaMemcpyAsync(d_cols, h_cols, sizeof(int64_t) * n_cols, cudaMemcpyHostToDevice, st));
  FEMB_CUDA_OK(ctx, cudaMemsetAsync(colmap, 0xff, sizeof(int32_t) * (K.n_rows ? K.n_rows : 1), st));
  FEMB_CUDA_OK(ctx, cudaMemsetAsync(cnt, 0, sizeof(int64_t) * (n_rows + 1), st));
  if (n_cols) k_slice_colmap<<<grid_for(n_cols, 256), 256, 0, st>>>(n_cols, d_cols, colmap);
  if (n_rows && K.nnz)
    k_slice_rows<<<grid_for(n_rows * 8, 256), 256, 0, st>>>(n_rows, d_rows, colmap, K.indptr, K.indices, K.values, cnt, nullptr, nullptr, nullptr, 0);
  size_t need = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, need, cnt, optr, (int)(n_rows + 1), st);
  if (need > ctx->cub_tmp_bytes) {
    if (ctx->cub_tmp) cudaFreeAsync(ctx->cub_tmp, st);
    ctx->cub_tmp = nullptr; ctx->cub_tmp_bytes = 0;
    FEMB_CUDA_OK(ctx, cudaMallocAsync(&ctx->cub_tmp, need + 256, ctx->pool, st));
    ctx->cub_tmp_bytes = need + 256;
  }
  size_t tb = ctx->cub_tmp_bytes;
  FEMB_CUDA_OK(ctx, cub::DeviceScan::ExclusiveSum(ctx->cub_tmp, tb, cnt, optr, (int)(n_rows + 1), st));
  FEMB_CUDA_OK(ctx, cudaMemcpyAsync(ctx->h_scalars, optr + n_rows, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  FEMB_CUDA_OK(ctx, cudaStreamSynchronize(st));
  const int64_t nnz = ctx->h_scalars[0];
  ctx->launches += 4;
  if (nnz_out) *nnz_out = nnz;
  if (!h_indptr && !h_indices && !h_values) return 0; // size query
  int32_t *oidx = nullptr;
  double *oval = nullptr;
  if ((rc = scratch_alloc(ctx, &oidx, nnz))) return rc;
  if ((rc = scratch_alloc(ctx, &oval, nnz))) return rc;
  if (n_rows && K.nnz) {
    k_slice_rows<<<grid_for(n_rows * 8, 256), 256, 0, st>>>(n_rows, d_rows, colmap, K.indptr, K.indices, K.values, cnt, optr, oidx, oval, 1);
    ctx->launches++;
  }
  if (h_indptr) FEMB_CUDA_OK(ctx, cudaMemcpyAsync(h_indptr, optr, sizeof(int64_t) * (n_rows + 1), cudaMemcpyDeviceToHost, st));
  if (h_indices && nnz) FEMB_CUDA_OK(ctx, cudaMemcpyAsync(h_indices, oidx, sizeof(int32_t) * nnz, cudaMemcpyDeviceToHost, st));
  if (h_values && nnz) FEMB_CUDA_OK(ctx, cudaMemcpyAsync(h_values, oval, sizeof(double) * nnz, cudaMemcpyDeviceToHost, st));
  FEMB_CUDA_OK(ctx, cudaStreamSynchronize(st));
  return 0;
}

} // namespace femb200
