// exp_tma.cu -- micro-benchmark: how should a warp fetch one 128-byte staged record per node row and step?
//   V0  every lane of a 3-lane group loads the whole record with four 256-bit LDGs (same addresses inside a group)
//   V1  one lane per group issues a 128-byte cp.async.bulk (TMA) into a per-warp ring, mbarrier wait, LDS.128 reads
//   V2  the warp loads the 10 records cooperatively (one 32-byte sector per lane), STS into the ring, LDS.128 reads
// Access pattern mimics the node-row kernel on config 1: 228,241 rows, 24 incident elements per row, records of
// elements near 6*row (+- neighbour cells).  Prints ms per variant and M records/s.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/_build/exp_tma scripts/exp_tma.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int NPW = 10;          // node rows per warp
constexpr int RECB = 128;        // record bytes
constexpr int SLOTB = 144;       // ring slot stride (16-byte pad => conflict-free broadcast reads)
constexpr int STEPS = 24;

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
__device__ __forceinline__ void ld256(const double *p, double (&v)[4]) {
  asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p));
}

__device__ __forceinline__ int64_t elem_of(int64_t row, int step, int64_t n_elem) {
  // 24 incident elements: 6 neighbour cells x 4, offsets like a 48^3 BCC mesh in cell-major element order
  const int cell = step >> 2, j = step & 3;
  const int64_t mag = cell < 2 ? 12 : (cell < 4 ? 12 * 48 : 12 * 48 * 48);
  const int64_t off = cell == 0 ? 0 : ((cell & 1) ? -mag : mag);
  int64_t e = row * 6 - (row * 6) % 12 + off + j * 3 + (row % 3);
  if (e < 0) e += n_elem;
  if (e >= n_elem) e -= n_elem;
  return e;
}

template <int V, int S>
__global__ void __launch_bounds__(128) k_fetch(const double *__restrict__ stag, int64_t n_elem, int64_t n_rows, double *__restrict__ out) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const int r = lane / 3, i = lane % 3;
  const bool act = lane < NPW * 3;
  const int64_t row = ((int64_t)blockIdx.x * nw + warp) * NPW + (act ? r : 0);
  unsigned char *ring = smem + (size_t)warp * (S * NPW * SLOTB + 64);
  uint64_t *bars = reinterpret_cast<uint64_t *>(ring + S * NPW * SLOTB);
  double acc = 0.0;
  if (V == 0) {
    if (act && row < n_rows) {
      double va[4], vb[4], vc[4], vd[4];
#pragma unroll 2
      for (int s = 0; s < STEPS; ++s) {
        const double *rec = stag + elem_of(row, s, n_elem) * 16;
        ld256(rec, va); ld256(rec + 4, vb); ld256(rec + 8, vc); ld256(rec + 12, vd);
        acc += va[i] * vb[0] + vc[1] * vd[3] + va[3];
      }
    }
  } else if (V == 1) {
    if (lane == 0)
      for (int s = 0; s < S; ++s) mbar_init(&bars[s], 1);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    const bool rowok = act && row < n_rows;
    const bool prod = rowok && i == 0;
    const unsigned nprod = __popc(__ballot_sync(0xffffffffu, prod));
    // prologue: S stages in flight
    for (int s = 0; s < S && s < STEPS; ++s) {
      if (lane == 0 && nprod) mbar_expect_tx(&bars[s], nprod * RECB);
      __syncwarp();
      if (prod) bulk_g2s(ring + (s * NPW + r) * SLOTB, stag + elem_of(row, s, n_elem) * 16, RECB, &bars[s]);
    }
    for (int s = 0; s < STEPS; ++s) {
      const int st = s % S;
      const uint32_t par = (s / S) & 1;
      if (nprod) mbar_wait(&bars[st], par);
      if (rowok) {
        const double2 *rec = reinterpret_cast<const double2 *>(ring + (st * NPW + r) * SLOTB);
        double2 w[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) w[c] = rec[c];
        const double va[4] = {w[0].x, w[0].y, w[1].x, w[1].y};
        acc += va[i] * w[2].x + w[4].y * w[7].y + va[3];
      }
      __syncwarp();
      if (s + S < STEPS) {
        if (lane == 0 && nprod) mbar_expect_tx(&bars[st], nprod * RECB);
        __syncwarp();
        if (prod) bulk_g2s(ring + (st * NPW + r) * SLOTB, stag + elem_of(row, s + S, n_elem) * 16, RECB, &bars[st]);
      }
    }
  } else {
    // cooperative: 40 sectors per step, lane l takes sector l and (l < 8) sector 32 + l; register prefetch one step ahead
    const int64_t row0 = ((int64_t)blockIdx.x * nw + warp) * NPW;
    const bool rowok = act && row < n_rows;
    auto sector_ptr = [&](int sec, int s) -> const double * {
      const int rr = sec >> 2, q = sec & 3;
      int64_t rw = row0 + rr;
      if (rw >= n_rows) rw = n_rows - 1;
      return stag + elem_of(rw, s, n_elem) * 16 + q * 4;
    };
    double a0[4], a1[4] = {0, 0, 0, 0};
    ld256(sector_ptr(lane, 0), a0);
    if (lane < 8) ld256(sector_ptr(32 + lane, 0), a1);
    for (int s = 0; s < STEPS; ++s) {
      double2 *d0 = reinterpret_cast<double2 *>(ring + ((lane >> 2) * SLOTB + (lane & 3) * 32));
      d0[0] = make_double2(a0[0], a0[1]); d0[1] = make_double2(a0[2], a0[3]);
      if (lane < 8) {
        double2 *d1 = reinterpret_cast<double2 *>(ring + (((32 + lane) >> 2) * SLOTB + (lane & 3) * 32));
        d1[0] = make_double2(a1[0], a1[1]); d1[1] = make_double2(a1[2], a1[3]);
      }
      if (s + 1 < STEPS) {
        ld256(sector_ptr(lane, s + 1), a0);
        if (lane < 8) ld256(sector_ptr(32 + lane, s + 1), a1);
      }
      __syncwarp();
      if (rowok) {
        const double2 *rec = reinterpret_cast<const double2 *>(ring + r * SLOTB);
        double2 w[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) w[c] = rec[c];
        const double va[4] = {w[0].x, w[0].y, w[1].x, w[1].y};
        acc += va[i] * w[2].x + w[4].y * w[7].y + va[3];
      }
      __syncwarp();
    }
  }
  if (act && row < n_rows) out[row * 3 + i] = acc;
}

template <int V, int S>
static void run(const char *name, const double *stag, int64_t n_elem, int64_t n_rows, double *out, int smem_extra_kb) {
  const int T = 128, nw = T / 32;
  const size_t smem = (size_t)nw * (S * NPW * SLOTB + 64) + (size_t)smem_extra_kb * 1024;
  CK(cudaFuncSetAttribute(k_fetch<V, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  const unsigned grid = (unsigned)((n_rows + nw * NPW - 1) / (nw * NPW));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float best = 1e9f;
  int nb = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_fetch<V, S>, T, smem));
  for (int it = 0; it < 5; ++it) {
    CK(cudaEventRecord(e0));
    k_fetch<V, S><<<grid, T, smem>>>(stag, n_elem, n_rows, out);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  CK(cudaGetLastError());
  double h[3]; CK(cudaMemcpy(h, out + 3000, sizeof h, cudaMemcpyDeviceToHost));
  printf("%-28s stages %d smem/block %6.1f KB blocks/SM %2d : %.4f ms  %.1f M records/s  (check %.6g)\n", name, S, smem / 1024.0, nb, best,
         n_rows * (double)STEPS / best / 1e3, h[0] + h[1] + h[2]);
}

int main() {
  const int64_t n_elem = 1299456, n_rows = 228241;
  double *stag, *out;
  CK(cudaMalloc(&stag, n_elem * 16 * sizeof(double)));
  CK(cudaMalloc(&out, n_rows * 3 * sizeof(double)));
  {
    double *h = (double *)malloc(n_elem * 16 * sizeof(double));
    for (int64_t k = 0; k < n_elem * 16; ++k) h[k] = (double)((k * 2654435761u) % 1000) * 1e-3;
    CK(cudaMemcpy(stag, h, n_elem * 16 * sizeof(double), cudaMemcpyHostToDevice));
    free(h);
  }
  for (int extra : {0, 16, 40}) {
    printf("--- extra dynamic shared memory per block: %d KB\n", extra);
    run<0, 1>("V0 4x LDG.256 per lane", stag, n_elem, n_rows, out, extra);
    run<1, 2>("V1 cp.async.bulk 128 B", stag, n_elem, n_rows, out, extra);
    run<1, 4>("V1 cp.async.bulk 128 B", stag, n_elem, n_rows, out, extra);
    run<1, 8>("V1 cp.async.bulk 128 B", stag, n_elem, n_rows, out, extra);
    run<2, 1>("V2 cooperative LDG+STS", stag, n_elem, n_rows, out, extra);
  }
  return 0;
}
