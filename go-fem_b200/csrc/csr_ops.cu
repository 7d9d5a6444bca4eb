// csr_ops.cu -- operations on the device-resident K (canonical CSR).
//   csr_merge_into : K += A with pattern union  (repeated AddIsoparametric calls accumulate into the
//                    same lap.Sparse, assembler_isoparametric.go:121)
//   csr_matvec     : y = K x                   (forces.MulVec(K,u), constitution/solids/patch_test.go:92)
#include "common.cuh"
#include <cub/cub.cuh>

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

} // namespace femb200
