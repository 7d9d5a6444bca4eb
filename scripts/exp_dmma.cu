// exp_dmma.cu -- the experiment BASELINE.json's north star asks for: "FP64 tensor-core MMA is used only for the larger
// Hexa20/Tetra10 B^T D B products, and only if it beats register-blocked DFMA".
//   1. peak: independent mma.sync.m8n8k4.f64 chains per warp on all SMs  vs  independent DFMA chains
//   2. the product itself, operands in shared memory as in a per-element kernel: a 32 x 32 block of
//      Ke += (B^T C f)[32 x 6] * B[6 x 32] per quadrature point (dimC = 6), 27 points per element (Hexa20),
//        DMMA: K padded to 8 (two m8n8k4 steps), 4 x 4 tiles per warp, fragments loaded from shared memory
//        DFMA: every lane owns a 4 x 8 sub-block (32 accumulators), operands broadcast from shared memory, K = 6 exactly
//      reported as USEFUL TFLOP/s (2 * 32 * 32 * 6 flops per block and point).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/_build/exp_dmma scripts/exp_dmma.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ void dmma884(double (&d)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d[0]), "+d"(d[1]) : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(256) k_peak_dmma(int iters, double *out) {
  double acc[8][2];
  for (int t = 0; t < 8; ++t) acc[t][0] = acc[t][1] = 0.0;
  const double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int t = 0; t < 8; ++t) dmma884(acc[t], a, b);
  double s = 0;
  for (int t = 0; t < 8; ++t) s += acc[t][0] + acc[t][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) k_peak_dfma(int iters, double *out) {
  double acc[16];
  for (int t = 0; t < 16; ++t) acc[t] = 0.0;
  const double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9;
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int t = 0; t < 16; ++t) acc[t] = fma(acc[t], a, b);
  double s = 0;
  for (int t = 0; t < 16; ++t) s += acc[t];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ---- the product: one warp = one 32 x 32 block, NQP points, operands staged in shared memory --------------------------
constexpr int NQP = 27, NST = 3, M = 32, KC = 6, KP = 8; // NST points staged per warp and cycled (shared-memory budget)

template <bool USE_DMMA>
__global__ void __launch_bounds__(128) k_btdb(int nblocks_per_warp, const double *__restrict__ gA, const double *__restrict__ gB, double *__restrict__ out) {
  // per warp: A[q][M][KP] (B^T C f, K padded with zeros) and B[q][KP][M]
  extern __shared__ double sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double *sA = sm + (size_t)warp * (2 * NST * M * KP);
  double *sB = sA + NST * M * KP;
  for (int i = lane; i < NST * M * KP; i += 32) { sA[i] = gA[i]; sB[i] = gB[i]; }
  __syncwarp();
  double total = 0.0;
  for (int blk = 0; blk < nblocks_per_warp; ++blk) {
    if constexpr (USE_DMMA) {
      double acc[4][4][2];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
      const int r = lane >> 2, c = lane & 3;
      for (int q = 0; q < NQP; ++q) {
        const double *A = sA + (q % NST) * M * KP, *Bm = sB + (q % NST) * KP * M;
#pragma unroll
        for (int ks = 0; ks < KP / 4; ++ks) {
          double af[4], bf[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) af[i] = A[(8 * i + r) * KP + 4 * ks + c];        // A tile i: row r, col c
#pragma unroll
          for (int j = 0; j < 4; ++j) bf[j] = Bm[(4 * ks + c) * M + 8 * j + r];        // B tile j: row c, col r
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) dmma884(acc[i][j], af[i], bf[j]);
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) total += acc[i][j][0] + acc[i][j][1];
    } else {
      // lane owns rows 4*(lane/4) .. +3 and columns 8*(lane%4) .. +7
      double acc[4][8];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.0;
      const int r0 = 4 * (lane >> 2), c0 = 8 * (lane & 3);
      for (int q = 0; q < NQP; ++q) {
        const double *A = sA + (q % NST) * M * KP, *Bm = sB + (q % NST) * KP * M;
#pragma unroll
        for (int k = 0; k < KC; ++k) {
          double a[4], b[8];
#pragma unroll
          for (int i = 0; i < 4; ++i) a[i] = A[(r0 + i) * KP + k];
#pragma unroll
          for (int j = 0; j < 8; j += 2) {
            const double2 w = *reinterpret_cast<const double2 *>(Bm + k * M + c0 + j);
            b[j] = w.x; b[j + 1] = w.y;
          }
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) total += acc[i][j];
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = total;
}

template <class F> static float best_ms(F f, int reps = 5) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int i = 0; i < reps; ++i) {
    CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  CK(cudaGetLastError());
  return best;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  const int sms = p.multiProcessorCount;
  double *out; CK(cudaMalloc(&out, sizeof(double) * sms * 8 * 256));
  const int iters = 20000;
  {
    const int grid = sms * 8;
    float ms = best_ms([&] { k_peak_dmma<<<grid, 256>>>(iters, out); });
    const double fl = (double)grid * 8 /*warps*/ * iters * 8 /*mma*/ * 512.0;
    printf("peak DMMA m8n8k4 : %.2f TFLOP/s (%.3f ms)\n", fl / ms / 1e9, ms);
    ms = best_ms([&] { k_peak_dfma<<<grid, 256>>>(iters, out); });
    const double fl2 = (double)grid * 256 * iters * 16 * 2.0;
    printf("peak DFMA        : %.2f TFLOP/s (%.3f ms)\n", fl2 / ms / 1e9, ms);
  }
  {
    const size_t n = (size_t)NST * M * KP;
    double *hA = (double *)malloc(n * 8), *hB = (double *)malloc(n * 8);
    for (size_t i = 0; i < n; ++i) { hA[i] = ((i % KP) < KC) ? 1e-3 * (i % 97) : 0.0; hB[i] = ((i / M) % KP < KC) ? 1e-3 * (i % 89) : 0.0; }
    double *dA, *dB; CK(cudaMalloc(&dA, n * 8)); CK(cudaMalloc(&dB, n * 8));
    CK(cudaMemcpy(dA, hA, n * 8, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, hB, n * 8, cudaMemcpyHostToDevice));
    const int T = 128, warps = T / 32, per_warp = 400;
    const size_t smem = (size_t)warps * 2 * n * 8;
    CK(cudaFuncSetAttribute(k_btdb<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(k_btdb<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int nb1 = 0, nb2 = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb1, k_btdb<true>, T, smem));
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb2, k_btdb<false>, T, smem));
    const int grid = sms * (nb1 < nb2 ? nb1 : nb2);
    const double useful = (double)grid * warps * per_warp * NQP * 2.0 * M * M * KC;
    float ms = best_ms([&] { k_btdb<true><<<grid, T, smem>>>(per_warp, dA, dB, out); });
    printf("B^T D B 32x32 block, dimC=6, 27 points, operands in shared memory (%d blocks/SM):\n", nb1 < nb2 ? nb1 : nb2);
    printf("  DMMA (K padded to 8): %.2f useful TFLOP/s (%.3f ms)\n", useful / ms / 1e9, ms);
    ms = best_ms([&] { k_btdb<false><<<grid, T, smem>>>(per_warp, dA, dB, out); });
    printf("  DFMA (4x8 per lane) : %.2f useful TFLOP/s (%.3f ms)\n", useful / ms / 1e9, ms);
  }
  return 0;
}
