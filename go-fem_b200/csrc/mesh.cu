// mesh.cu -- synthetic benchmark meshes generated on the device (no reference counterpart; benchmarks and tests only).
//
// femb200_mesh_bcc_tetra4 builds, node for node and element for element, what go-fem_b200/meshes.py bcc_tetra4_box builds
// with numpy on the host (the mesh family of the reference's BenchmarkTetra4Assembly, benchmark_test.go:34-35: only the
// (dofs, elems) counts of manigold's BCC mesher are pinned, README.md:82-86):
//   nodes    corner nodes (nx+1)(ny+1)(nz+1), x-major, then the cell centres nx*ny*nz; interior nodes jittered by
//            0.15*h*(u-0.5) per axis, u = splitmix64(0x9E3779B97F4A7C15 ^ (3*node + axis)) / 2^64
//   elements 4 tets (centre-, centre+, face edge k) per interior cell face, cell-major order (key = 12*cell + 4*axis + k),
//            last two vertices swapped where the Jacobian determinant is negative (tetra4.go:40-46 orientation)
// The 101 M-element scaling mesh (N = 204) takes milliseconds here and minutes in numpy.
#include "common.cuh"
#include <cub/cub.cuh>

namespace {

__device__ __forceinline__ unsigned long long splitmix64(unsigned long long x) {
  unsigned long long z = x + 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

__global__ void k_bcc_nodes(int nx, int ny, int nz, double h, int jittered, double *__restrict__ nodes) {
  const int64_t Mx = nx + 1, My = ny + 1, Mz = nz + 1;
  const int64_t ncorner = Mx * My * Mz, n = ncorner + (int64_t)nx * ny * nz;
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= n) return;
  double p[3];
  if (v < ncorner) {
    const int64_t i = v / (My * Mz), j = (v / Mz) % My, k = v % Mz;
    p[0] = (double)i * h; p[1] = (double)j * h; p[2] = (double)k * h;
  } else {
    const int64_t c = v - ncorner;
    const int64_t i = c / ((int64_t)ny * nz), j = (c / nz) % ny, k = c % nz;
    p[0] = ((double)i + 0.5) * h; p[1] = ((double)j + 0.5) * h; p[2] = ((double)k + 0.5) * h;
  }
  if (jittered) {
    const double ext[3] = {(double)nx * h, (double)ny * h, (double)nz * h};
    bool interior = true;
    for (int a = 0; a < 3; ++a) interior = interior && (p[a] > 1e-12) && (p[a] < ext[a] - 1e-12);
    if (interior)
      for (int a = 0; a < 3; ++a) {
        const double u = __ull2double_rn(splitmix64(0x9E3779B97F4A7C15ull ^ ((unsigned long long)v * 3ull + (unsigned long long)a))) / 18446744073709551616.0;
        p[a] = __dadd_rn(p[a], __dmul_rn(__dmul_rn(0.15, h), (u - 0.5))); // amp * h * (u - 0.5), left to right like numpy
      }
  }
  nodes[3 * v] = p[0]; nodes[3 * v + 1] = p[1]; nodes[3 * v + 2] = p[2];
}

// tets per cell: 4 for every axis along which the cell is not the last one (the face towards the next cell)
__global__ void k_bcc_cell_counts(int nx, int ny, int nz, int32_t *__restrict__ cnt) {
  const int64_t nc = (int64_t)nx * ny * nz;
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c > nc) return;
  if (c == nc) { cnt[c] = 0; return; }
  const int i = (int)(c / ((int64_t)ny * nz)), j = (int)((c / nz) % ny), k = (int)(c % nz);
  cnt[c] = 4 * ((i < nx - 1) + (j < ny - 1) + (k < nz - 1));
}

__global__ void k_bcc_tets(int nx, int ny, int nz, const int64_t *__restrict__ off, const double *__restrict__ nodes, int32_t *__restrict__ conn) {
  const int64_t nc = (int64_t)nx * ny * nz;
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nc) return;
  const int64_t My = ny + 1, Mz = nz + 1;
  const int64_t ncorner = (int64_t)(nx + 1) * My * Mz;
  const int i = (int)(c / ((int64_t)ny * nz)), j = (int)((c / nz) % ny), k = (int)(c % nz);
  auto cid = [&](int a, int b, int cc) -> int64_t { return ((int64_t)a * My + b) * Mz + cc; };
  auto mid = [&](int a, int b, int cc) -> int64_t { return ncorner + ((int64_t)a * ny + b) * nz + cc; };
  int64_t e = off[c];
  for (int axis = 0; axis < 3; ++axis) {
    int64_t lo, hi, f[4];
    if (axis == 0) {
      if (i >= nx - 1) continue;
      const int A = i + 1;                 // face between cell A-1 = i and cell A
      lo = mid(A - 1, j, k); hi = mid(A, j, k);
      f[0] = cid(A, j, k); f[1] = cid(A, j + 1, k); f[2] = cid(A, j + 1, k + 1); f[3] = cid(A, j, k + 1);
    } else if (axis == 1) {
      if (j >= ny - 1) continue;
      const int A = j + 1;
      lo = mid(i, A - 1, k); hi = mid(i, A, k);
      f[0] = cid(i, A, k); f[1] = cid(i, A, k + 1); f[2] = cid(i + 1, A, k + 1); f[3] = cid(i + 1, A, k);
    } else {
      if (k >= nz - 1) continue;
      const int A = k + 1;
      lo = mid(i, j, A - 1); hi = mid(i, j, A);
      f[0] = cid(i, j, A); f[1] = cid(i + 1, j, A); f[2] = cid(i + 1, j + 1, A); f[3] = cid(i, j + 1, A);
    }
    for (int q = 0; q < 4; ++q, ++e) {
      int64_t v[4] = {lo, hi, f[q], f[(q + 1) & 3]};
      double p[4][3];
      for (int a = 0; a < 4; ++a)
        for (int d = 0; d < 3; ++d) p[a][d] = nodes[3 * v[a] + d];
      double r[3][3];
      for (int a = 0; a < 3; ++a)
        for (int d = 0; d < 3; ++d) r[a][d] = p[a + 1][d] - p[0][d];
      const double det = r[0][0] * (r[1][1] * r[2][2] - r[1][2] * r[2][1]) - r[0][1] * (r[1][0] * r[2][2] - r[1][2] * r[2][0]) +
                         r[0][2] * (r[1][0] * r[2][1] - r[1][1] * r[2][0]);
      if (det < 0) { const int64_t t = v[2]; v[2] = v[3]; v[3] = t; }
      for (int a = 0; a < 4; ++a) conn[4 * e + a] = (int32_t)v[a];
    }
  }
}

} // namespace

extern "C" {

int femb200_mesh_bcc_tetra4(femb200_ctx *ctx, int32_t nx, int32_t ny, int32_t nz, double h, int32_t jittered, double **d_nodes,
                            int64_t *n_nodes, int32_t **d_conn, int64_t *n_elem) {
  using namespace femb200;
  if (!ctx || nx < 1 || ny < 1 || nz < 1 || !d_nodes || !d_conn || !n_nodes || !n_elem) return FEMB200_ERR_BAD_ARG;
  FEMB_CUDA_OK(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const int64_t nc = (int64_t)nx * ny * nz;
  const int64_t nn_ = (int64_t)(nx + 1) * (ny + 1) * (nz + 1) + nc;
  const int64_t ne = 4 * ((int64_t)(nx - 1) * ny * nz + (int64_t)nx * (ny - 1) * nz + (int64_t)nx * ny * (nz - 1));
  if (nn_ >= ((int64_t)1 << 31) || 4 * ne >= ((int64_t)1 << 33)) return FEMB200_ERR_TOO_LARGE;
  double *nodes = nullptr;
  int32_t *conn = nullptr, *cnt = nullptr;
  int64_t *off = nullptr;
  FEMB_CUDA_OK(ctx, cudaMalloc((void **)&nodes, sizeof(double) * 3 * nn_));
  FEMB_CUDA_OK(ctx, cudaMalloc((void **)&conn, sizeof(int32_t) * 4 * (ne > 0 ? ne : 1)));
  FEMB_CUDA_OK(ctx, cudaMalloc((void **)&cnt, sizeof(int32_t) * (nc + 1)));
  FEMB_CUDA_OK(ctx, cudaMalloc((void **)&off, sizeof(int64_t) * (nc + 1)));
  k_bcc_nodes<<<grid_for(nn_, 256), 256, 0, st>>>(nx, ny, nz, h, jittered, nodes);
  k_bcc_cell_counts<<<grid_for(nc + 1, 256), 256, 0, st>>>(nx, ny, nz, cnt);
  size_t tb = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tb, cnt, off, (int)(nc + 1), st);
  void *tmp = nullptr;
  FEMB_CUDA_OK(ctx, cudaMalloc(&tmp, tb + 256));
  FEMB_CUDA_OK(ctx, cub::DeviceScan::ExclusiveSum(tmp, tb, cnt, off, (int)(nc + 1), st));
  k_bcc_tets<<<grid_for(nc, 128), 128, 0, st>>>(nx, ny, nz, off, nodes, conn);
  ctx->launches += 5;
  FEMB_CUDA_OK(ctx, cudaGetLastError());
  FEMB_CUDA_OK(ctx, cudaStreamSynchronize(st));
  cudaFree(tmp); cudaFree(cnt); cudaFree(off);
  *d_nodes = nodes; *d_conn = conn; *n_nodes = nn_; *n_elem = ne;
  return 0;
}

int femb200_copy_to_host(femb200_ctx *ctx, void *dst, const void *src, int64_t bytes) {
  using namespace femb200;
  if (!ctx || bytes < 0 || (bytes > 0 && (!dst || !src))) return FEMB200_ERR_BAD_ARG;
  FEMB_CUDA_OK(ctx, cudaSetDevice(ctx->device));
  FEMB_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  if (bytes) FEMB_CUDA_OK(ctx, cudaMemcpy(dst, src, (size_t)bytes, cudaMemcpyDeviceToHost));
  return 0;
}

void femb200_mesh_free(femb200_ctx *ctx, void *p) {
  if (!ctx || !p) return;
  cudaSetDevice(ctx->device);
  cudaFree(p);
}

} // extern "C"
