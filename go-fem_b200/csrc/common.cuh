// common.cuh -- internal structures of libfemb200 (not part of the ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>
#include <functional>
#include "../../include/femb200.h"

namespace femb200 {

// ---- limits of the device tables -------------------------------------------------------------
constexpr int kMaxNN = 20;       // Hexa20
constexpr int kMaxD = 3;
constexpr int kMaxNdn = 6;       // DofsFlag has 6 named bits (fem.go:70-82)
constexpr int kMaxDimC = 6;
constexpr int kTabDoubles = 5632; // 44 KB of __constant__ table space (N | dN | w | C)
// Slot map entries pos16[k][b] (row-local block position of node b of the element of incidence k).  For elements with
// <= 4 nodes the two top bits of entry b = 0 also carry the local index of the row's own node in the element, so the
// fused scalar-row kernel reads ONE 8-byte word per incident element and never touches the incidence list.
// Per-call device scalars (ctx->d_scalars, mirrored in pinned ctx->h_scalars and read back with one copy)
constexpr int kScalarMaxRowlen = 0, kScalarNblk = 1, kScalarGroupNeed = 2 /* ..5 */, kScalarNShort = 6, kScalarGuard = 7, kScalarErr = 8, kNumScalars = 16;
constexpr size_t kStageBytes = 64 * 1024; // pinned staging buffer of the per-call tables
constexpr int kPosLaShift = 14;
constexpr unsigned kPosMask = 0x3fffu;
__host__ __device__ inline unsigned pos_mask(int nn) { return nn <= 4 ? kPosMask : 0xffffu; }

// Layout of the per-call table blob (doubles): [C dimC*dimC][w nqp][N nqp*nn][dN nqp*d*nn]
struct TabLayout {
  int offC, offW, offN, offDN, total;
};
inline TabLayout tab_layout(int d, int nn, int nqp, int dimC) {
  TabLayout t;
  t.offC = 0;
  t.offW = dimC * dimC;
  t.offN = t.offW + nqp;
  t.offDN = t.offN + nqp * nn;
  t.total = t.offDN + nqp * d * nn;
  return t;
}

// Device error word: min over (elem << 16 | qp << 4 | kind) so the LOWEST failing element wins,
// then its first failing quadrature point (sequential semantics of fem.go:140-151).
constexpr unsigned long long kNoError = ~0ull;
__host__ __device__ inline unsigned long long pack_error(long long elem, int qp, int kind) {
  return ((unsigned long long)elem << 16) | ((unsigned long long)(qp & 0xfff) << 4) | (unsigned long long)(kind & 0xf);
}

struct DeviceCsr {
  int64_t n_rows = 0;
  int64_t nnz = 0;
  int64_t *indptr = nullptr;  // [n_rows+1]
  int32_t *indices = nullptr; // [nnz]
  double *values = nullptr;   // [nnz]
};

// Per-call symbolic plan kept for the slot-map query.
struct LastPlan {
  int nn = 0, nu = 0;
  int64_t n_elem = 0;
  int32_t *n2e_ptr = nullptr;   // [n_nodes+1] incidence offsets
  uint32_t *n2e_t = nullptr;    // [n_elem*nn] t = e*nn + a, grouped by node, ascending
  uint16_t *pos16 = nullptr;    // [n_elem*nn][nn] row-local block position
  int64_t *blkptr = nullptr;    // [n_nodes+1] node-block row offsets (of THIS call's pattern)
  // kept for femb200_reassemble (numeric-only re-assembly on the cached pattern)
  int32_t *conn32 = nullptr;    // [n_elem*nn] validated connectivity
  int32_t *order = nullptr;     // two-class visiting order (see SymbolicResult), or NULL
  int64_t n_short = 0;
  int64_t n_numeric = 0;        // elements [n_numeric, n_elem) are pattern-only
  int32_t *blkcols = nullptr;   // [nblk] column node of every node block (row I: blkcols[blkptr[I] ...], ascending)
  int64_t nblk = 0;             // node blocks of this call's pattern: K.nnz must equal nu*nu*nblk while matches_K holds
  int d = 0, ndn = 0, bkind = 0, max_rowlen = 0;
  int64_t group_need[4] = {0, 0, 0, 0};
  bool identity = false;
  bool matches_K = false;       // K's pattern is exactly this call's pattern (no other call has been merged in)
};

} // namespace femb200

struct femb200_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = true;
  bool concurrent = false;                  // other contexts may run calls on this device at the same time
  cudaMemPool_t pool = nullptr;
  int sm_count = 148;
  size_t smem_optin = 0;
  int64_t launches = 0;
  std::string last_cuda_error;
  void *cub_tmp = nullptr;
  size_t cub_tmp_bytes = 0;
  unsigned long long *d_err = nullptr;      // device error word (alias of d_scalars + kScalarErr)
  int64_t *d_scalars = nullptr;             // small device scratch (8 x int64)
  int64_t *h_scalars = nullptr;             // pinned host mirror
  unsigned long long *h_err = nullptr;      // pinned (alias of h_scalars + kScalarErr)
  char *h_stage = nullptr;                  // pinned staging of the per-call tables
  cudaEvent_t ev[12] = {};
  // second stream: the per-element checks of the fused numeric path run beside the symbolic phase
  cudaStream_t stream2 = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;       // per-element checks beside the row-pattern kernel
  cudaEvent_t ev_fork2 = nullptr, ev_join2 = nullptr;     // CSR expansion beside the numeric kernel
  // Sizes of the previous successful call with the same mesh signature.  A call that finds a matching hint sizes its CSR
  // arrays and launch shapes from it and does NOT wait for the device between the symbolic and the numeric phase; the
  // kernels guard every write against the hinted capacities and the call is repeated the slow way if a guard tripped.
  struct SizeHint {
    bool valid = false;
    int64_t n_nodes = -1, n_elem = -1;
    int nn = 0, mdpn = 0, nu = 0;
    int64_t nblk = 0;
    int max_rowlen = 0;
  } hint;
  // per-call scratch arena: temporaries of one add call are carved out of (normally) ONE persistent chunk, so a
  // repeated assembly performs no allocation for them and the memory pool is not churned by multi-GB temporaries
  struct ScratchChunk { char *p; size_t size; };
  std::vector<ScratchChunk> scratch;
  size_t scratch_off = 0, scratch_used = 0, scratch_hint = 0;
};

struct femb200_assembler {
  femb200_ctx *ctx = nullptr;
  const double *d_nodes = nullptr; // [n_nodes][3]
  bool own_nodes = true;
  int64_t n_nodes = 0;
  int32_t mdpn = 0;
  const uint8_t *d_node_mask = nullptr; // optional: only rows of nodes with mask != 0 are assembled (row-block restriction)
  const int32_t *d_col_map = nullptr;   // optional: local node -> global node (monotone); K.indices then hold GLOBAL dof ids
  femb200::DeviceCsr K;
  femb200::LastPlan plan;
  int32_t err_kind = 0;
  int64_t err_elem = -1;
  int32_t err_qp = -1;
  float phase_ms[FEMB200_N_TIMERS] = {};
  bool split_numeric = false; // ev[9..11] bracket k_geometry / k_rows of the last call
  bool check_done = false;    // the per-element checks of this call already ran (second stream)
};

namespace femb200 {

#define FEMB_CUDA_OK(ctx, expr)                                                         \
  do {                                                                                  \
    cudaError_t _e = (expr);                                                            \
    if (_e != cudaSuccess) {                                                            \
      (ctx)->last_cuda_error = std::string(#expr) + ": " + cudaGetErrorString(_e);      \
      return FEMB200_ERR_CUDA;                                                          \
    }                                                                                   \
  } while (0)

template <class T> inline int dev_alloc(femb200_ctx *ctx, T **p, size_t n) {
  *p = nullptr;
  if (n == 0) n = 1;
  FEMB_CUDA_OK(ctx, cudaMallocAsync((void **)p, n * sizeof(T), ctx->pool, ctx->stream));
  return 0;
}
template <class T> inline void dev_free(femb200_ctx *ctx, T *&p) {
  if (p) cudaFreeAsync((void *)p, ctx->stream);
  p = nullptr;
}
// bump allocation from the context's scratch arena (256-byte aligned); released as a whole by scratch_release
template <class T> inline int scratch_alloc(femb200_ctx *ctx, T **p, size_t n) {
  *p = nullptr;
  if (n == 0) n = 1;
  const size_t bytes = (n * sizeof(T) + 255) & ~(size_t)255;
  if (ctx->scratch.empty() || ctx->scratch_off + bytes > ctx->scratch.back().size) {
    size_t sz = bytes > ctx->scratch_hint ? bytes : ctx->scratch_hint;
    if (!ctx->scratch.empty() && sz < ctx->scratch.back().size / 2) sz = ctx->scratch.back().size / 2; // geometric growth within a call
    void *q = nullptr;
    FEMB_CUDA_OK(ctx, cudaMallocAsync(&q, sz, ctx->pool, ctx->stream));
    ctx->scratch.push_back({(char *)q, sz});
    ctx->scratch_off = 0;
  }
  *p = reinterpret_cast<T *>(ctx->scratch.back().p + ctx->scratch_off);
  ctx->scratch_off += bytes;
  ctx->scratch_used += bytes;
  return 0;
}
// bytes of NEW device memory the arena would have to reserve for the given requests, in order (mirrors scratch_alloc)
inline size_t scratch_extra_needed(const femb200_ctx *ctx, const size_t *bytes_each, int n) {
  size_t cap = ctx->scratch.empty() ? 0 : ctx->scratch.back().size, off = ctx->scratch.empty() ? 0 : ctx->scratch_off, extra = 0;
  bool empty = ctx->scratch.empty();
  for (int i = 0; i < n; ++i) {
    const size_t bytes = ((bytes_each[i] ? bytes_each[i] : 1) + 255) & ~(size_t)255;
    if (empty || off + bytes > cap) {
      size_t sz = bytes > ctx->scratch_hint ? bytes : ctx->scratch_hint;
      if (!empty && sz < cap / 2) sz = cap / 2;
      extra += sz;
      cap = sz; off = 0; empty = false;
    }
    off += bytes;
  }
  return extra;
}
inline void scratch_release(femb200_ctx *ctx, bool drop_all = false) {
  if (ctx->scratch.size() > 1 || drop_all) { // consolidate: the next call gets one chunk that fits everything
    for (auto &c : ctx->scratch) cudaFreeAsync(c.p, ctx->stream);
    ctx->scratch.clear();
  }
  if (ctx->scratch_used > ctx->scratch_hint) ctx->scratch_hint = ctx->scratch_used;
  ctx->scratch_off = 0;
  ctx->scratch_used = 0;
}
struct ScratchScope {
  femb200_ctx *ctx;
  explicit ScratchScope(femb200_ctx *c) : ctx(c) {}
  ~ScratchScope() { scratch_release(ctx); }
};

inline unsigned grid_for(int64_t n, int block) { return (unsigned)((n + block - 1) / block); }

// ---- phases (symbolic.cu / numeric.cu / csr_ops.cu) -------------------------------------------
struct ElemShape {
  int d, nn, ndn, nqp, dimC, bkind;
  int nu;                 // number of distinct dofmap values
  int umap[kMaxNdn];      // sorted distinct dofmap values U
  int dof2u[kMaxNdn];     // dofmap[i] -> index into U
  bool identity;          // dofmap == [0..ndn) and nu == ndn == mdpn
  int mdpn;
};

struct SymbolicResult {
  int32_t *n2e_ptr = nullptr;
  uint32_t *n2e_t = nullptr;
  uint16_t *pos16 = nullptr;
  int64_t *blkptr = nullptr;
  int64_t nblk = 0;
  int max_rowlen = 0;
  int32_t *order = nullptr;   // two-class visiting order of the node rows (short rows first), or NULL
  int64_t n_short = 0;        // rows [0, n_short) of `order` have rowlen <= max_rowlen / 2
  int64_t group_need[4] = {0, 0, 0, 0}; // max over aligned groups of 4 / 8 / 16 / 32 consecutive rows of sum(rowlen)
  int32_t *blkcols = nullptr;       // [nblk] column node of every node block
  const int *d_umap = nullptr;      // distinct-dof map on the device, uploaded with the tables (or NULL: symbolic_phase uploads it)
  int32_t *precount = nullptr;      // [2*(n_nodes+1)] incidence histogram | zeroed fill cursors, taken while narrowing (or NULL)
  bool expand_forked = false;       // k_expand8 was queued on the second stream: the caller joins ev_join2 before it reads K
  const int32_t *rowcols = nullptr; // ascending column nodes of row I at rowcols[cols_stride * n2e_ptr[I] ...] (call scratch)
  int cols_stride = 0;
  bool speculative = false;         // sizes came from ctx->hint: nblk / max_rowlen are capacities until the call's final sync
  int64_t cap_blk = 0;              // capacity (node blocks) of indices / values / blkcols
  DeviceCsr csr; // indptr + indices filled, values allocated (uninitialised)
};

// conn32: device [n_elem*nn] validated int32 connectivity.
// fork_at / fork_fn: the caller's side-stream work (per-element checks) is started after the adjacency (2) or after the row
// pattern (3) has been queued
int symbolic_phase(femb200_assembler *ga, const ElemShape &sh, const int32_t *conn32, int64_t n_elem,
                   SymbolicResult *out, bool use_hint, int fork_at = 0, const std::function<int()> *fork_fn = nullptr);
int numeric_phase(femb200_assembler *ga, const ElemShape &sh, const double *d_tab, const double *h_tab, int tab_doubles,
                  const int32_t *conn32, int64_t n_elem, SymbolicResult *sym, int64_t n_elem_total);
bool numeric_supported(const ElemShape &sh);
int reset_call_scalars(femb200_ctx *ctx);
int check_phase(femb200_assembler *ga, const ElemShape &sh, const double *d_tab, const int32_t *conn32, int64_t n_elem, cudaStream_t st);
int csr_merge_into(femb200_assembler *ga, DeviceCsr *add); // K += add (consumes add)
int csr_accumulate_rows(femb200_assembler *ga, int64_t n_rows, const int64_t *rows, const int64_t *indptr, const int32_t *indices,
                        const double *values);
int csr_keep_rows(femb200_assembler *ga, const uint8_t *keep);
int csr_add_coo(femb200_assembler *ga, int64_t n, const int64_t *h_rows, const int64_t *h_cols, const double *h_vals);
int csr_add_entries(femb200_assembler *ga, int64_t n_rows, const int64_t *rows, const int64_t *indptr, const int32_t *indices,
                    const double *values);
int csr_matvec(femb200_assembler *ga, const double *d_x, double *d_y);
int csr_pack_rows_async(femb200_assembler *ga, int64_t n_rows, const int64_t *rows, const int64_t *out_indptr, int32_t *out_indices,
                        double *out_values, unsigned long long *mismatch);
int csr_add_entries_async(femb200_assembler *ga, int64_t n_rows, const int64_t *rows, const int64_t *indptr, const int32_t *indices,
                          const double *values, unsigned long long *mismatch);
int strains_phase(femb200_assembler *ga, const ElemShape &sh, const double *d_tab, const int32_t *conn32,
                  int64_t n_elem, const int32_t *d_dofmap, const double *d_u, double *d_strains);
int narrow_and_validate_conn(femb200_assembler *ga, const void *d_conn_in, bool is64, int64_t n, int nn,
                             int32_t *conn32_out, int32_t *cnt = nullptr, bool agg = false);
int slot_map(femb200_assembler *ga, int64_t *d_slot);
int probe_peaks(femb200_ctx *ctx, int reps, double out[2]);
int csr_slice(femb200_assembler *ga, int64_t n_rows, const int64_t *h_rows, int64_t n_cols, const int64_t *h_cols, int64_t *h_indptr,
              int32_t *h_indices, double *h_values, int64_t *nnz_out);

} // namespace femb200
