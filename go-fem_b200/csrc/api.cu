// api.cu -- the extern "C" surface declared in include/femb200.h.
#include "common.cuh"
#include <cstring>
#include <vector>
#include <functional>

using namespace femb200;

namespace femb200 {
__global__ void k_init_call(int64_t *scalars) {
  if (threadIdx.x < kNumScalars) scalars[threadIdx.x] = threadIdx.x == kScalarErr ? (int64_t)kNoError : 0;
}
// zeroes the per-call scalars and arms the error word (one tiny launch instead of a memset per scalar)
int reset_call_scalars(femb200_ctx *ctx) {
  k_init_call<<<1, 32, 0, ctx->stream>>>(ctx->d_scalars);
  ctx->launches++;
  FEMB_CUDA_OK(ctx, cudaGetLastError());
  return 0;
}
} // namespace femb200

extern "C" {

int femb200_abi_version(void) { return FEMB200_ABI_VERSION; }

int femb200_ctx_create(int device, femb200_ctx **out) {
  if (!out) return FEMB200_ERR_BAD_ARG;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return FEMB200_ERR_CUDA;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return FEMB200_ERR_CUDA;
  if (prop.major != 10) return FEMB200_ERR_CUDA; // sm_100a cubin only: no other architecture can run it
  femb200_ctx *ctx = new femb200_ctx();
  ctx->device = device;
  ctx->sm_count = prop.multiProcessorCount;
  ctx->smem_optin = prop.sharedMemPerBlockOptin;
#define CK(e)                                                          \
  do {                                                                 \
    if ((e) != cudaSuccess) { delete ctx; return FEMB200_ERR_CUDA; }   \
  } while (0)
  CK(cudaSetDevice(device));
  CK(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
  CK(cudaDeviceGetDefaultMemPool(&ctx->pool, device));
  uint64_t thr = UINT64_MAX; // keep freed blocks in the pool: repeated assemblies never hit cudaMalloc
  CK(cudaMemPoolSetAttribute(ctx->pool, cudaMemPoolAttrReleaseThreshold, &thr));
  // one small device block for every per-call scalar (common.cuh: kScalar*), its pinned mirror, and a pinned staging buffer
  // for the per-call tables: a call reads all of them back with ONE copy and uploads its tables with ONE copy
  CK(cudaMalloc((void **)&ctx->d_scalars, kNumScalars * sizeof(int64_t)));
  CK(cudaMallocHost((void **)&ctx->h_scalars, kNumScalars * sizeof(int64_t)));
  ctx->d_err = reinterpret_cast<unsigned long long *>(ctx->d_scalars + kScalarErr);
  ctx->h_err = reinterpret_cast<unsigned long long *>(ctx->h_scalars + kScalarErr);
  CK(cudaMallocHost((void **)&ctx->h_stage, kStageBytes));
  for (auto &e : ctx->ev) CK(cudaEventCreate(&e));
  CK(cudaStreamCreateWithFlags(&ctx->stream2, cudaStreamNonBlocking));
  CK(cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&ctx->ev_fork2, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&ctx->ev_join2, cudaEventDisableTiming));
#undef CK
  *out = ctx;
  return 0;
}

void femb200_ctx_destroy(femb200_ctx *ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  if (ctx->cub_tmp) cudaFreeAsync(ctx->cub_tmp, ctx->stream);
  scratch_release(ctx, true);
  cudaStreamSynchronize(ctx->stream);
  cudaFree(ctx->d_scalars);
  cudaFreeHost(ctx->h_scalars);
  cudaFreeHost(ctx->h_stage);
  for (auto &e : ctx->ev) if (e) cudaEventDestroy(e);
  if (ctx->stream2) { cudaStreamSynchronize(ctx->stream2); cudaStreamDestroy(ctx->stream2); }
  if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
  if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
  if (ctx->ev_fork2) cudaEventDestroy(ctx->ev_fork2);
  if (ctx->ev_join2) cudaEventDestroy(ctx->ev_join2);
  if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

int femb200_ctx_trim(femb200_ctx *ctx) {
  if (!ctx) return FEMB200_ERR_BAD_ARG;
  FEMB_CUDA_OK(ctx, cudaSetDevice(ctx->device));
  FEMB_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  scratch_release(ctx, true);
  ctx->scratch_hint = 0;
  if (ctx->cub_tmp) { cudaFreeAsync(ctx->cub_tmp, ctx->stream); ctx->cub_tmp = nullptr; ctx->cub_tmp_bytes = 0; }
  FEMB_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  if (ctx->pool) FEMB_CUDA_OK(ctx, cudaMemPoolTrimTo(ctx->pool, 0));
  return 0;
}

int femb200_ctx_set_stream(femb200_ctx *ctx, void *s) {
  if (!ctx) return FEMB200_ERR_BAD_ARG;
  cudaStreamSynchronize(ctx->stream);
  if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
  ctx->stream = (cudaStream_t)s;
  ctx->own_stream = false;
  return 0;
}

int femb200_ctx_set_concurrent(femb200_ctx *ctx, int32_t on) {
  if (!ctx) return FEMB200_ERR_BAD_ARG;
  ctx->concurrent = on != 0;
  return 0;
}

const char *femb200_last_cuda_error(const femb200_ctx *ctx) { return ctx ? ctx->last_cuda_error.c_str() : "no context"; }
int64_t femb200_kernel_launches(const femb200_ctx *ctx) { return ctx ? ctx->launches : 0; }
int femb200_sync(femb200_ctx *ctx) {
  if (!ctx) return FEMB200_ERR_BAD_ARG;
  FEMB_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  return 0;
}

int femb200_assembler_create(femb200_ctx *ctx, const double *nodes, int64_t n_nodes, int32_t mdpn,
                             int32_t on_device, femb200_assembler **out) {
  if (!ctx || !out || n_nodes < 0 || mdpn < 1 || mdpn > 16 || (n_nodes > 0 && !nodes)) return FEMB200_ERR_BAD_ARG;
  *out = nullptr;
  FEMB_CUDA_OK(ctx, cudaSetDevice(ctx->device));
  femb200_assembler *ga = new femb200_assembler();
  ga->ctx = ctx;
  ga->n_nodes = n_nodes;
  ga->mdpn = mdpn;
  if (on_device) {
    ga->d_nodes = nodes;
    ga->own_nodes = false;
  } else {
    double *d = nullptr;
    int rc = dev_alloc(ctx, &d, (size_t)n_nodes * 3);
    if (rc) { delete ga; return rc; }
    if (n_nodes) {
      cudaError_t e = cudaMemcpyAsync(d, nodes, sizeof(double) * 3 * n_nodes, cudaMemcpyHostToDevice, ctx->stream);
      if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream); // caller memory may be released after return
      if (e != cudaSuccess) { ctx->last_cuda_error = cudaGetErrorString(e); delete ga; return FEMB200_ERR_CUDA; }
    }
    ga->d_nodes = d;
  }
  ga->K.n_rows = n_nodes * mdpn; // assembler.go:22
  int rc = dev_alloc(ctx, &ga->K.indptr, (size_t)ga->K.n_rows + 1);
  if (rc) { delete ga; return rc; }
  cudaMemsetAsync(ga->K.indptr, 0, sizeof(int64_t) * (ga->K.n_rows + 1), ctx->stream);
  *out = ga;
  return 0;
}

static void free_plan(femb200_assembler *ga) {
  femb200_ctx *ctx = ga->ctx;
  dev_free(ctx, ga->plan.n2e_ptr);
  dev_free(ctx, ga->plan.n2e_t);
  dev_free(ctx, ga->plan.pos16);
  dev_free(ctx, ga->plan.blkcols);
  dev_free(ctx, ga->plan.blkptr);
  dev_free(ctx, ga->plan.conn32);
  dev_free(ctx, ga->plan.order);
  ga->plan = LastPlan();
}

void femb200_assembler_destroy(femb200_assembler *ga) {
  if (!ga) return;
  femb200_ctx *ctx = ga->ctx;
  cudaSetDevice(ctx->device);
  free_plan(ga);
  dev_free(ctx, ga->K.indptr); dev_free(ctx, ga->K.indices); dev_free(ctx, ga->K.values);
  if (ga->own_nodes) { double *d = const_cast<double *>(ga->d_nodes); dev_free(ctx, d); }
  cudaStreamSynchronize(ctx->stream);
  delete ga;
}

// descriptor -> ElemShape + packed table blob [C | w | N | dN]
static int prepare(const femb200_assembler *ga, const femb200_elem_desc *d, ElemShape *sh, std::vector<double> *blob) {
  if (!d || !d->Npg || !d->dNpg || !d->wpg || !d->C || !d->dofmap) return FEMB200_ERR_BAD_ARG;
  if (d->d < 1 || d->d > kMaxD || d->nn < 1 || d->nn > kMaxNN || d->ndn < 1 || d->ndn > kMaxNdn || d->nqp < 1 ||
      d->dimC < 1 || d->dimC > kMaxDimC || d->model_dofs_per_node != ga->mdpn)
    return FEMB200_ERR_BAD_ARG;
  sh->d = d->d; sh->nn = d->nn; sh->ndn = d->ndn; sh->nqp = d->nqp; sh->dimC = d->dimC; sh->bkind = d->bkind;
  sh->mdpn = ga->mdpn;
  // distinct dofmap values, ascending (U), and dof -> index in U
  bool seen[16] = {};
  for (int i = 0; i < d->ndn; ++i) {
    if (d->dofmap[i] < 0 || d->dofmap[i] >= ga->mdpn) return FEMB200_ERR_BAD_ARG;
    seen[d->dofmap[i]] = true;
  }
  sh->nu = 0;
  for (int i = 0; i < kMaxNdn; ++i) { sh->umap[i] = 0; sh->dof2u[i] = 0; }
  for (int m = 0; m < ga->mdpn && m < 16; ++m)
    if (seen[m]) {
      if (sh->nu >= kMaxNdn) return FEMB200_ERR_BAD_ARG;
      sh->umap[sh->nu++] = m;
    }
  sh->identity = (sh->nu == d->ndn) && (d->ndn == ga->mdpn);
  for (int i = 0; i < d->ndn; ++i) {
    for (int k = 0; k < sh->nu; ++k) if (sh->umap[k] == d->dofmap[i]) sh->dof2u[i] = k;
    if (d->dofmap[i] != i) sh->identity = false;
  }
  TabLayout tl = tab_layout(d->d, d->nn, d->nqp, d->dimC);
  blob->assign(tl.total, 0.0);
  std::memcpy(blob->data() + tl.offC, d->C, sizeof(double) * d->dimC * d->dimC);
  std::memcpy(blob->data() + tl.offW, d->wpg, sizeof(double) * d->nqp);
  std::memcpy(blob->data() + tl.offN, d->Npg, sizeof(double) * d->nqp * d->nn);
  std::memcpy(blob->data() + tl.offDN, d->dNpg, sizeof(double) * d->nqp * d->d * d->nn);
  return numeric_supported(*sh) ? 0 : FEMB200_ERR_UNSUPPORTED;
}

static int decode_error(femb200_assembler *ga) {
  unsigned long long w = *ga->ctx->h_err;
  if (w == kNoError) return 0;
  ga->err_kind = (int)(w & 0xf);
  if (ga->err_kind == 0) ga->err_kind = FEMB200_ERR_NODE_OOB; // device encoding, see k_narrow_validate
  ga->err_qp = (int)((w >> 4) & 0xfff);
  ga->err_elem = (int64_t)(w >> 16);
  if (ga->err_kind == FEMB200_ERR_NODE_OOB) ga->err_qp = -1;
  return ga->err_kind;
}

// uploads (if needed), narrows to int32 and validates the connectivity
static int stage_conn(femb200_assembler *ga, const void *conn, bool is64, bool on_device, int64_t n, int nn,
                      int32_t **conn32, void **staged, int32_t *cnt = nullptr, bool agg = false) {
  femb200_ctx *ctx = ga->ctx;
  *staged = nullptr;
  int rc;
  if ((rc = dev_alloc(ctx, conn32, (size_t)n))) return rc; // pool, not scratch: the plan keeps it (femb200_reassemble)
  const void *d_in = conn;
  if (!on_device && n) {
    size_t bytes = (size_t)n * (is64 ? 8 : 4);
    char *sb = nullptr;
    if ((rc = scratch_alloc(ctx, &sb, bytes))) return rc;
    *staged = sb;
    FEMB_CUDA_OK(ctx, cudaMemcpyAsync(*staged, conn, bytes, cudaMemcpyHostToDevice, ctx->stream));
    d_in = *staged;
  }
  return narrow_and_validate_conn(ga, d_in, is64, n, nn, *conn32, cnt, agg);
}

int femb200_add_isoparametric(femb200_assembler *ga, const femb200_elem_desc *desc, const void *conn,
                              int32_t conn_is_int64, int32_t conn_on_device, int64_t n_elem) {
  return femb200_add_isoparametric_halo(ga, desc, conn, conn_is_int64, conn_on_device, n_elem, 0);
}

// One add call.  Host synchronisations: ONE (at the end) when the context holds a size hint for this mesh signature and the
// fused numeric path applies; otherwise one more between the symbolic and the numeric phase (allocation sizes).
static int add_isoparametric_impl(femb200_assembler *ga, const femb200_elem_desc *desc, const void *conn, int32_t conn_is_int64,
                                  int32_t conn_on_device, int64_t n_elem, int64_t n_pattern_only, bool allow_hint) {
  femb200_ctx *ctx = ga->ctx;
  cudaStream_t st = ctx->stream;
  FEMB_CUDA_OK(ctx, cudaSetDevice(ctx->device));
  ga->err_kind = 0; ga->err_elem = -1; ga->err_qp = -1;
  ElemShape sh;
  std::vector<double> blob;
  int rc = prepare(ga, desc, &sh, &blob);
  if (rc) { ga->err_kind = rc; return rc; }
  for (float &f : ga->phase_ms) f = 0.f;
  ga->split_numeric = false;
  ga->check_done = false;
  const int64_t launches0 = ctx->launches;

  const char *rows_env = getenv("FEMB200_ROWS");
  const bool fused_ok = sh.identity && sh.nn <= 4 && sh.nqp <= 9 && !(rows_env && (!strcmp(rows_env, "noderows") || !strcmp(rows_env, "srow")));
  const femb200_ctx::SizeHint &h = ctx->hint;
  const bool use_hint = allow_hint && fused_ok && h.valid && h.n_nodes == ga->n_nodes && h.n_elem == n_elem && h.nn == sh.nn &&
                        h.mdpn == ga->mdpn && h.nu == sh.nu && !getenv("FEMB200_NO_HINT");

  FEMB_CUDA_OK(ctx, cudaEventRecord(ctx->ev[0], st));
  *ctx->h_err = kNoError;
  if ((rc = reset_call_scalars(ctx))) { ga->err_kind = rc; return rc; }
  int32_t *conn32 = nullptr;
  void *staged = nullptr;
  double *d_tab = nullptr;
  SymbolicResult sym;
  ScratchScope scratch_scope(ctx); // staged, d_tab, symbolic / numeric temporaries: released on every return path
  auto cleanup = [&]() { dev_free(ctx, conn32); };
  auto drop_sym = [&]() {
    dev_free(ctx, sym.n2e_ptr); dev_free(ctx, sym.n2e_t); dev_free(ctx, sym.pos16); dev_free(ctx, sym.blkptr); dev_free(ctx, sym.order); dev_free(ctx, sym.blkcols);
    dev_free(ctx, sym.csr.indptr); dev_free(ctx, sym.csr.indices); dev_free(ctx, sym.csr.values);
  };
  auto fail = [&](int code) {
    cleanup(); drop_sym();
    ga->err_kind = code;
    cudaStreamSynchronize(st);
    if (ctx->stream2) cudaStreamSynchronize(ctx->stream2);
    return code;
  };
  // the incidence histogram of the adjacency is taken in the pass that narrows / validates the connectivity
  int32_t *precount = nullptr;
  rc = scratch_alloc(ctx, &precount, (size_t)2 * (ga->n_nodes + 1));
  if (!rc && cudaMemsetAsync(precount, 0, sizeof(int32_t) * 2 * (ga->n_nodes + 1), st) != cudaSuccess) { ctx->last_cuda_error = "cudaMemsetAsync"; rc = FEMB200_ERR_CUDA; }
  if (!rc) rc = stage_conn(ga, conn, conn_is_int64 != 0, conn_on_device != 0, n_elem * sh.nn, sh.nn, &conn32, &staged, precount, sh.nn == 4 && sh.d == 3);
  sym.precount = precount;
  // tables [C | w | N | dN] and the distinct-dof map behind them: ONE upload from the pinned staging buffer
  const size_t tab_slots = blob.size() + (kMaxNdn * sizeof(int) + sizeof(double) - 1) / sizeof(double);
  if (!rc) rc = scratch_alloc(ctx, &d_tab, tab_slots);
  if (!rc) {
    cudaError_t e;
    if (tab_slots * sizeof(double) <= kStageBytes) {
      std::memcpy(ctx->h_stage, blob.data(), sizeof(double) * blob.size());
      std::memcpy(ctx->h_stage + sizeof(double) * blob.size(), sh.umap, sizeof(int) * kMaxNdn);
      e = cudaMemcpyAsync(d_tab, ctx->h_stage, sizeof(double) * tab_slots, cudaMemcpyHostToDevice, st);
      sym.d_umap = reinterpret_cast<const int *>(d_tab + blob.size());
    } else {
      e = cudaMemcpyAsync(d_tab, blob.data(), sizeof(double) * blob.size(), cudaMemcpyHostToDevice, st);
    }
    if (e != cudaSuccess) { ctx->last_cuda_error = cudaGetErrorString(e); rc = FEMB200_ERR_CUDA; }
  }
  if (!rc && cudaEventRecord(ctx->ev[1], st) != cudaSuccess) { ctx->last_cuda_error = "cudaEventRecord"; rc = FEMB200_ERR_CUDA; }
  if (rc) return fail(rc);

  // the reference's per-element checks (:95-107) do not depend on the pattern: on the fused path they run on the second
  // stream, beside the symbolic phase (FP64 work beside atomics / integer work)
  bool forked = false;
  std::function<int()> fork_fn = [&]() -> int {
    cudaError_t e = cudaEventRecord(ctx->ev_fork, st);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(ctx->stream2, ctx->ev_fork, 0);
    if (e != cudaSuccess) { ctx->last_cuda_error = cudaGetErrorString(e); return FEMB200_ERR_CUDA; }
    int r2 = check_phase(ga, sh, d_tab, conn32, n_elem - n_pattern_only, ctx->stream2);
    if (r2 == FEMB200_ERR_UNSUPPORTED) return 0;
    if (r2) return r2;
    e = cudaEventRecord(ctx->ev_join, ctx->stream2);
    if (e != cudaSuccess) { ctx->last_cuda_error = cudaGetErrorString(e); return FEMB200_ERR_CUDA; }
    forked = true;
    ga->check_done = true;
    return 0;
  };
  int fork_at = 0;
  if (fused_ok && ctx->stream2 && n_elem - n_pattern_only > 0 && !getenv("FEMB200_NO_FORK")) {
    fork_at = getenv("FEMB200_FORK_AT") ? atoi(getenv("FEMB200_FORK_AT")) : 2; // beside the row-pattern kernel (integer work)
    if (fork_at == 1 && (rc = fork_fn())) return fail(rc);
  }

  rc = symbolic_phase(ga, sh, conn32, n_elem, &sym, use_hint, fork_at, &fork_fn); // synchronises once unless sized from the hint
  if (rc == -1) {
    // a node index is out of range.  The reference's loop is sequential (fem.go:140): an element BEFORE the offending one
    // may fail its Jacobian checks first, so those checks run over the elements in front of it and the lowest word wins.
    rc = decode_error(ga);
    if (rc == FEMB200_ERR_NODE_OOB && !forked && ga->err_elem > 0) {
      int rc2 = check_phase(ga, sh, d_tab, conn32, ga->err_elem < n_elem - n_pattern_only ? ga->err_elem : n_elem - n_pattern_only, st);
      if (rc2 == 0) {
        cudaError_t e = cudaMemcpyAsync(ctx->h_err, ctx->d_err, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e == cudaSuccess) rc = decode_error(ga);
      }
    } else if (forked) {
      cudaError_t e = cudaStreamSynchronize(ctx->stream2); // the forked checks cover every element already
      if (e == cudaSuccess) e = cudaMemcpyAsync(ctx->h_err, ctx->d_err, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st);
      if (e == cudaSuccess) e = cudaStreamSynchronize(st);
      if (e == cudaSuccess) rc = decode_error(ga);
    }
  }
  if (rc) return fail(rc);

  rc = numeric_phase(ga, sh, d_tab, blob.data(), (int)blob.size(), conn32, n_elem - n_pattern_only, &sym, n_elem);
  if (rc == -3) { // hinted sizes unusable on the path this call takes: again, the slow way
    fail(0);
    return add_isoparametric_impl(ga, desc, conn, conn_is_int64, conn_on_device, n_elem, n_pattern_only, false);
  }
  if (!rc) {
    cudaError_t e = cudaSuccess;
    if (forked) e = cudaStreamWaitEvent(st, ctx->ev_join, 0);
    if (e == cudaSuccess && sym.expand_forked) e = cudaStreamWaitEvent(st, ctx->ev_join2, 0);
    if (e == cudaSuccess) // every scalar of the call (sizes, guard flag, error word) in one copy
      e = cudaMemcpyAsync(ctx->h_scalars, ctx->d_scalars, kNumScalars * sizeof(int64_t), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaEventRecord(ctx->ev[5], st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st); // blob / caller buffers are released after this point
    if (e != cudaSuccess) { ctx->last_cuda_error = cudaGetErrorString(e); rc = FEMB200_ERR_CUDA; }
  }
  if (!rc) rc = decode_error(ga);
  if (rc) return fail(rc); // K untouched (:118-120)
  if (sym.speculative) {
    const int64_t real_maxrl = ctx->h_scalars[0], real_nblk = ctx->h_scalars[1];
    if (ctx->h_scalars[kScalarGuard] != 0 || real_maxrl > sym.max_rowlen || real_nblk > sym.cap_blk || real_maxrl > (int64_t)pos_mask(sh.nn)) {
      ctx->hint.valid = false; // a guard tripped: the mesh is not the one the hint came from
      fail(0);
      return add_isoparametric_impl(ga, desc, conn, conn_is_int64, conn_on_device, n_elem, n_pattern_only, false);
    }
    sym.max_rowlen = (int)real_maxrl;
    sym.nblk = real_nblk;
    sym.csr.nnz = (int64_t)sh.nu * sh.nu * real_nblk;
  }
  {
    ctx->hint.valid = true;
    ctx->hint.n_nodes = ga->n_nodes; ctx->hint.n_elem = n_elem; ctx->hint.nn = sh.nn; ctx->hint.mdpn = ga->mdpn; ctx->hint.nu = sh.nu;
    ctx->hint.nblk = sym.nblk; ctx->hint.max_rowlen = sym.max_rowlen;
  }

  // commit: K += this call (assembler_isoparametric.go:121); keep the plan for the slot-map query
  free_plan(ga);
  ga->plan.nn = sh.nn; ga->plan.nu = sh.nu; ga->plan.n_elem = n_elem;
  ga->plan.n2e_ptr = sym.n2e_ptr; ga->plan.n2e_t = sym.n2e_t; ga->plan.pos16 = sym.pos16; ga->plan.blkptr = sym.blkptr;
  ga->plan.blkcols = sym.blkcols;
  ga->plan.conn32 = conn32; conn32 = nullptr;
  ga->plan.order = sym.order; ga->plan.n_short = sym.n_short;
  ga->plan.n_numeric = n_elem - n_pattern_only;
  ga->plan.nblk = sym.nblk;
  ga->plan.d = sh.d; ga->plan.ndn = sh.ndn; ga->plan.bkind = sh.bkind; ga->plan.max_rowlen = sym.max_rowlen; ga->plan.identity = sh.identity;
  for (int s = 0; s < 4; ++s) ga->plan.group_need[s] = sym.group_need[s];
  const bool first = (ga->K.nnz == 0);
  ga->plan.matches_K = first;
  rc = csr_merge_into(ga, &sym.csr); // first call on the assembler: adopts the arrays, no device work
  cleanup();
  if (!first) { // a union with earlier calls ran on the device: its time belongs to the call
    FEMB_CUDA_OK(ctx, cudaEventRecord(ctx->ev[6], st));
    FEMB_CUDA_OK(ctx, cudaStreamSynchronize(st));
  }
  if (rc) { ga->err_kind = rc; return rc; }
  float ms = 0;
  const int pairs[5][2] = {{0, 1}, {1, 2}, {2, 3}, {3, 4}, {4, 5}};
  for (int i = 0; i < 5; ++i)
    if (cudaEventElapsedTime(&ms, ctx->ev[pairs[i][0]], ctx->ev[pairs[i][1]]) == cudaSuccess) ga->phase_ms[i] = ms;
  if (!first && cudaEventElapsedTime(&ms, ctx->ev[5], ctx->ev[6]) == cudaSuccess) ga->phase_ms[5] = ms;
  if (cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[first ? 5 : 6]) == cudaSuccess) ga->phase_ms[6] = ms;
  ga->phase_ms[7] = (float)(ctx->launches - launches0);
  if (ga->split_numeric) {
    if (cudaEventElapsedTime(&ms, ctx->ev[9], ctx->ev[10]) == cudaSuccess) ga->phase_ms[8] = ms;
    if (cudaEventElapsedTime(&ms, ctx->ev[10], ctx->ev[11]) == cudaSuccess) ga->phase_ms[9] = ms;
  }
  return 0;
}

int femb200_add_isoparametric_halo(femb200_assembler *ga, const femb200_elem_desc *desc, const void *conn,
                                   int32_t conn_is_int64, int32_t conn_on_device, int64_t n_elem, int64_t n_pattern_only) {
  if (!ga || n_elem < 0 || n_pattern_only < 0 || n_pattern_only > n_elem || (n_elem > 0 && !conn)) return FEMB200_ERR_BAD_ARG;
  return add_isoparametric_impl(ga, desc, conn, conn_is_int64, conn_on_device, n_elem, n_pattern_only, true);
}

int femb200_reassemble(femb200_assembler *ga, const femb200_elem_desc *desc, const double *nodes, int32_t nodes_on_device) {
  if (!ga) return FEMB200_ERR_BAD_ARG;
  femb200_ctx *ctx = ga->ctx;
  cudaStream_t st = ctx->stream;
  FEMB_CUDA_OK(ctx, cudaSetDevice(ctx->device));
  LastPlan &pl = ga->plan;
  if (!pl.conn32 || !pl.matches_K || ga->K.nnz == 0) return FEMB200_ERR_BAD_ARG; // exactly one committed call on this assembler
  if (ga->K.nnz != (int64_t)pl.nu * pl.nu * pl.nblk) return FEMB200_ERR_BAD_ARG;   // K was reshaped behind the plan's back
  ga->err_kind = 0; ga->err_elem = -1; ga->err_qp = -1;
  ElemShape sh;
  std::vector<double> blob;
  int rc = prepare(ga, desc, &sh, &blob);
  if (!rc && (sh.nn != pl.nn || sh.ndn != pl.ndn || sh.d != pl.d || sh.nu != pl.nu || sh.identity != pl.identity)) rc = FEMB200_ERR_BAD_ARG;
  if (rc) { ga->err_kind = rc; return rc; }
  for (float &f : ga->phase_ms) f = 0.f;
  ga->split_numeric = false;
  ga->check_done = false;
  const int64_t launches0 = ctx->launches;
  FEMB_CUDA_OK(ctx, cudaEventRecord(ctx->ev[0], st));
  // new coordinates (same node count): nonlinear / moving-mesh re-assembly.  They replace the old ones only if the call
  // succeeds, so a failed re-assembly leaves the assembler exactly as it was.
  const double *old_nodes = ga->d_nodes;
  const bool old_own = ga->own_nodes;
  double *fresh_nodes = nullptr;
  if (nodes) {
    if (nodes_on_device) {
      ga->d_nodes = nodes;
      ga->own_nodes = false;
    } else {
      if ((rc = dev_alloc(ctx, &fresh_nodes, (size_t)ga->n_nodes * 3))) return rc;
      cudaError_t e = cudaMemcpyAsync(fresh_nodes, nodes, sizeof(double) * 3 * ga->n_nodes, cudaMemcpyHostToDevice, st);
      if (e != cudaSuccess) { ctx->last_cuda_error = cudaGetErrorString(e); dev_free(ctx, fresh_nodes); return FEMB200_ERR_CUDA; }
      ga->d_nodes = fresh_nodes;
      ga->own_nodes = true;
    }
  }
  auto restore_nodes = [&]() {
    dev_free(ctx, fresh_nodes);
    ga->d_nodes = old_nodes;
    ga->own_nodes = old_own;
  };
  // from here on the assembler is mutated (node pointer swapped): every failure leaves through fail(), which frees the
  // temporaries and restores the old coordinates, so that a failed re-assembly leaves the assembler exactly as it was
  double *d_tab = nullptr, *newvals = nullptr;
  ScratchScope scratch_scope(ctx);
  auto fail = [&](int code, cudaError_t ce = cudaSuccess) {
    if (ce != cudaSuccess) ctx->last_cuda_error = cudaGetErrorString(ce);
    dev_free(ctx, newvals);
    restore_nodes();
    ga->err_kind = code;
    cudaStreamSynchronize(st);
    return code;
  };
  cudaError_t ce;
  *ctx->h_err = kNoError;
  if ((rc = reset_call_scalars(ctx))) return fail(rc);
  if ((rc = scratch_alloc(ctx, &d_tab, blob.size()))) return fail(rc);
  if ((ce = cudaMemcpyAsync(d_tab, blob.data(), sizeof(double) * blob.size(), cudaMemcpyHostToDevice, st)) != cudaSuccess) return fail(FEMB200_ERR_CUDA, ce);
  if ((rc = dev_alloc(ctx, &newvals, (size_t)ga->K.nnz))) return fail(rc);
  SymbolicResult sym;
  sym.n2e_ptr = pl.n2e_ptr; sym.n2e_t = pl.n2e_t; sym.pos16 = pl.pos16; sym.blkptr = pl.blkptr; sym.blkcols = pl.blkcols;
  sym.max_rowlen = pl.max_rowlen; sym.nblk = pl.nblk; sym.cap_blk = pl.nblk;
  sym.order = pl.order; sym.n_short = pl.n_short;
  for (int s = 0; s < 4; ++s) sym.group_need[s] = pl.group_need[s];
  sym.csr.n_rows = ga->K.n_rows; sym.csr.nnz = ga->K.nnz; sym.csr.indptr = ga->K.indptr; sym.csr.indices = ga->K.indices;
  sym.csr.values = newvals;
  if ((ce = cudaEventRecord(ctx->ev[4], st)) != cudaSuccess) return fail(FEMB200_ERR_CUDA, ce);
  rc = numeric_phase(ga, sh, d_tab, blob.data(), (int)blob.size(), pl.conn32, pl.n_numeric, &sym, pl.n_elem);
  if (!rc) {
    ce = cudaMemcpyAsync(ctx->h_err, ctx->d_err, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st);
    if (ce == cudaSuccess) ce = cudaEventRecord(ctx->ev[5], st);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);
    if (ce != cudaSuccess) return fail(FEMB200_ERR_CUDA, ce);
  }
  if (!rc) rc = decode_error(ga);
  if (rc) return fail(rc); // K untouched
  if (nodes && old_own) { double *d = const_cast<double *>(old_nodes); dev_free(ctx, d); }
  dev_free(ctx, ga->K.values);
  ga->K.values = newvals;
  float ms = 0;
  if (cudaEventElapsedTime(&ms, ctx->ev[4], ctx->ev[5]) == cudaSuccess) ga->phase_ms[4] = ms;
  if (cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[5]) == cudaSuccess) ga->phase_ms[6] = ms;
  ga->phase_ms[7] = (float)(ctx->launches - launches0);
  if (ga->split_numeric) {
    if (cudaEventElapsedTime(&ms, ctx->ev[9], ctx->ev[10]) == cudaSuccess) ga->phase_ms[8] = ms;
    if (cudaEventElapsedTime(&ms, ctx->ev[10], ctx->ev[11]) == cudaSuccess) ga->phase_ms[9] = ms;
  }
  return 0;
}

int femb200_last_error(const femb200_assembler *ga, int32_t *kind, int64_t *elem, int32_t *qp) {
  if (!ga) return FEMB200_ERR_BAD_ARG;
  if (kind) *kind = ga->err_kind;
  if (elem) *elem = ga->err_elem;
  if (qp) *qp = ga->err_qp;
  return 0;
}

int64_t femb200_total_dofs(const femb200_assembler *ga) { return ga ? ga->K.n_rows : 0; }
int64_t femb200_nnz(const femb200_assembler *ga) { return ga ? ga->K.nnz : 0; }

int femb200_csr_copy(const femb200_assembler *ga, int64_t *indptr, int32_t *indices, double *values) {
  if (!ga) return FEMB200_ERR_BAD_ARG;
  femb200_ctx *ctx = ga->ctx;
  cudaStream_t st = ctx->stream;
  FEMB_CUDA_OK(ctx, cudaSetDevice(ctx->device));
  const DeviceCsr &K = ga->K;
  if (indptr) FEMB_CUDA_OK(ctx, cudaMemcpyAsync(indptr, K.indptr, sizeof(int64_t) * (K.n_rows + 1), cudaMemcpyDeviceToHost, st));
  if (indices && K.nnz) FEMB_CUDA_OK(ctx, cudaMemcpyAsync(indices, K.indices, sizeof(int32_t) * K.nnz, cudaMemcpyDeviceToHost, st));
  if (values && K.nnz) FEMB_CUDA_OK(ctx, cudaMemcpyAsync(values, K.values, sizeof(double) * K.nnz, cudaMemcpyDeviceToHost, st));
  FEMB_CUDA_OK(ctx, cudaStreamSynchronize(st));
  return 0;
}

int64_t femb200_bsr_blocks(const femb200_assembler *ga) {
  if (!ga || !ga->plan.matches_K || !ga->plan.identity || !ga->plan.blkcols) return -1;
  return ga->plan.nblk;
}

int femb200_bsr_copy(const femb200_assembler *ga, int64_t *blkptr, int32_t *blkcols, double *values) {
  if (!ga) return FEMB200_ERR_BAD_ARG;
  if (femb200_bsr_blocks(ga) < 0) return FEMB200_ERR_UNSUPPORTED;
  femb200_ctx *ctx = ga->ctx;
  cudaStream_t st = ctx->stream;
  FEMB_CUDA_OK(ctx, cudaSetDevice(ctx->device));
  const LastPlan &pl = ga->plan;
  if (blkptr) FEMB_CUDA_OK(ctx, cudaMemcpyAsync(blkptr, pl.blkptr, sizeof(int64_t) * (ga->n_nodes + 1), cudaMemcpyDeviceToHost, st));
  if (blkcols && pl.nblk) FEMB_CUDA_OK(ctx, cudaMemcpyAsync(blkcols, pl.blkcols, sizeof(int32_t) * pl.nblk, cudaMemcpyDeviceToHost, st));
  if (values && ga->K.nnz) FEMB_CUDA_OK(ctx, cudaMemcpyAsync(values, ga->K.values, sizeof(double) * ga->K.nnz, cudaMemcpyDeviceToHost, st));
  FEMB_CUDA_OK(ctx, cudaStreamSynchronize(st));
  return 0;
}

int femb200_csr_device(const femb200_assembler *ga, const int64_t **indptr, const int32_t **indices, const double **values) {
  if (!ga) return FEMB200_ERR_BAD_ARG;
  if (indptr) *indptr = ga->K.indptr;
  if (indices) *indices = ga->K.indices;
  if (values) *values = ga->K.values;
  return 0;
}

__global__ void k_at(const int64_t *ptr, const int32_t *idx, const double *val, int64_t i, int32_t j, double *out) {
  double v = 0.0;
  int64_t lo = ptr[i], hi = ptr[i + 1];
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    if (idx[mid] < j) lo = mid + 1; else hi = mid;
  }
  if (lo < ptr[i + 1] && idx[lo] == j) v = val[lo];
  *out = v;
}

int femb200_csr_at(const femb200_assembler *ga, int64_t i, int64_t j, double *out) {
  if (!ga || !out || i < 0 || j < 0 || i >= ga->K.n_rows || j >= ga->K.n_rows) return FEMB200_ERR_BAD_ARG;
  femb200_ctx *ctx = ga->ctx;
  *out = 0.0;
  if (ga->K.nnz == 0) return 0;
  FEMB_CUDA_OK(ctx, cudaSetDevice(ctx->device));
  double *d = (double *)ctx->d_scalars;
  k_at<<<1, 1, 0, ctx->stream>>>(ga->K.indptr, ga->K.indices, ga->K.values, i, (int32_t)j, d);
  ctx->launches++;
  FEMB_CUDA_OK(ctx, cudaMemcpyAsync(ctx->h_scalars, d, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  FEMB_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  std::memcpy(out, ctx->h_scalars, sizeof(double));
  return 0;
}

int femb200_csr_slice(const femb200_assembler *cga, int64_t n_rows, const int64_t *rows, int64_t n_cols, const int64_t *cols,
                      int64_t *indptr, int32_t *indices, double *values, int64_t *nnz) {
  femb200_assembler *ga = const_cast<femb200_assembler *>(cga);
  if (!ga || n_rows < 0 || n_cols < 0 || (n_rows > 0 && !rows) || (n_cols > 0 && !cols)) return FEMB200_ERR_BAD_ARG;
  for (int64_t i = 0; i < n_rows; ++i)
    if (rows[i] < 0 || rows[i] >= ga->K.n_rows) return FEMB200_ERR_BAD_ARG;
  for (int64_t i = 0; i < n_cols; ++i)
    if (cols[i] < 0 || cols[i] >= ga->K.n_rows || (i > 0 && cols[i] <= cols[i - 1])) return FEMB200_ERR_BAD_ARG; // ascending, unique
  if (ga->d_col_map) return FEMB200_ERR_UNSUPPORTED; // rank-local assemblers hold global column ids
  FEMB_CUDA_OK(ga->ctx, cudaSetDevice(ga->ctx->device));
  return csr_slice(ga, n_rows, rows, n_cols, cols, indptr, indices, values, nnz);
}

int femb200_csr_matvec(const femb200_assembler *cga, const double *x, double *y) {
  femb200_assembler *ga = const_cast<femb200_assembler *>(cga);
  if (!ga || !x || !y) return FEMB200_ERR_BAD_ARG;
  femb200_ctx *ctx = ga->ctx;
  const int64_t n = ga->K.n_rows;
  FEMB_CUDA_OK(ctx, cudaSetDevice(ctx->device));
  double *dx = nullptr, *dy = nullptr;
  int rc;
  if ((rc = dev_alloc(ctx, &dx, (size_t)n))) return rc;
  if ((rc = dev_alloc(ctx, &dy, (size_t)n))) return rc;
  FEMB_CUDA_OK(ctx, cudaMemcpyAsync(dx, x, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
  rc = csr_matvec(ga, dx, dy);
  if (!rc) FEMB_CUDA_OK(ctx, cudaMemcpyAsync(y, dy, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
  dev_free(ctx, dx); dev_free(ctx, dy);
  FEMB_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  return rc;
}

int femb200_add_coo(femb200_assembler *ga, int64_t n, const int64_t *rows, const int64_t *cols, const double *vals) {
  if (!ga || n < 0 || (n > 0 && (!rows || !cols || !vals))) return FEMB200_ERR_BAD_ARG;
  for (int64_t k = 0; k < n; ++k)
    if (rows[k] < 0 || rows[k] >= ga->K.n_rows || cols[k] < 0 || cols[k] >= ga->K.n_rows) return FEMB200_ERR_BAD_ARG;
  if (n == 0) return 0;
  FEMB_CUDA_OK(ga->ctx, cudaSetDevice(ga->ctx->device));
  int rc = csr_add_coo(ga, n, rows, cols, vals);
  cudaStreamSynchronize(ga->ctx->stream);
  return rc;
}

int femb200_csr_accumulate_rows_device(femb200_assembler *ga, int64_t n_rows, const int64_t *rows, const int64_t *indptr,
                                       const int32_t *indices, const double *values) {
  if (!ga || n_rows < 0 || (n_rows > 0 && (!rows || !indptr))) return FEMB200_ERR_BAD_ARG;
  FEMB_CUDA_OK(ga->ctx, cudaSetDevice(ga->ctx->device));
  int rc = csr_accumulate_rows(ga, n_rows, rows, indptr, indices, values);
  cudaStreamSynchronize(ga->ctx->stream);
  return rc;
}

int femb200_csr_add_entries_device(femb200_assembler *ga, int64_t n_rows, const int64_t *rows, const int64_t *indptr,
                                   const int32_t *indices, const double *values) {
  if (!ga || n_rows < 0 || (n_rows > 0 && (!rows || !indptr))) return FEMB200_ERR_BAD_ARG;
  FEMB_CUDA_OK(ga->ctx, cudaSetDevice(ga->ctx->device));
  return csr_add_entries(ga, n_rows, rows, indptr, indices, values);
}

int femb200_csr_pack_rows_async(femb200_assembler *ga, int64_t n_rows, const int64_t *rows, const int64_t *out_indptr,
                                int32_t *out_indices, double *out_values, uint64_t *mismatch) {
  if (!ga || n_rows < 0 || (n_rows > 0 && (!rows || !out_indptr))) return FEMB200_ERR_BAD_ARG;
  FEMB_CUDA_OK(ga->ctx, cudaSetDevice(ga->ctx->device));
  if (n_rows > 0 && ga->K.nnz == 0) return FEMB200_ERR_PATTERN;
  return csr_pack_rows_async(ga, n_rows, rows, out_indptr, out_indices, out_values, (unsigned long long *)mismatch);
}

int femb200_csr_add_entries_async(femb200_assembler *ga, int64_t n_rows, const int64_t *rows, const int64_t *indptr,
                                  const int32_t *indices, const double *values, uint64_t *mismatch) {
  if (!ga || n_rows < 0 || (n_rows > 0 && (!rows || !indptr))) return FEMB200_ERR_BAD_ARG;
  FEMB_CUDA_OK(ga->ctx, cudaSetDevice(ga->ctx->device));
  if (n_rows > 0 && ga->K.nnz == 0) return FEMB200_ERR_PATTERN;
  return csr_add_entries_async(ga, n_rows, rows, indptr, indices, values, (unsigned long long *)mismatch);
}

int femb200_assembler_set_node_mask_device(femb200_assembler *ga, const uint8_t *node_mask) {
  if (!ga) return FEMB200_ERR_BAD_ARG;
  ga->d_node_mask = node_mask;
  return 0;
}

int femb200_assembler_set_column_map_device(femb200_assembler *ga, const int32_t *node_l2g, int64_t n_global_nodes) {
  if (!ga || n_global_nodes < 0 || n_global_nodes * ga->mdpn >= ((int64_t)1 << 31)) return FEMB200_ERR_BAD_ARG;
  ga->d_col_map = node_l2g;
  return 0;
}

int femb200_csr_keep_rows_device(femb200_assembler *ga, const uint8_t *keep) {
  if (!ga || !keep) return FEMB200_ERR_BAD_ARG;
  FEMB_CUDA_OK(ga->ctx, cudaSetDevice(ga->ctx->device));
  int rc = csr_keep_rows(ga, keep);
  cudaStreamSynchronize(ga->ctx->stream);
  return rc;
}

int femb200_last_slot_map(const femb200_assembler *cga, int64_t *slot) {
  femb200_assembler *ga = const_cast<femb200_assembler *>(cga);
  if (!ga || !slot || !ga->plan.n2e_t) return FEMB200_ERR_BAD_ARG;
  femb200_ctx *ctx = ga->ctx;
  FEMB_CUDA_OK(ctx, cudaSetDevice(ctx->device));
  const size_t n = (size_t)ga->plan.n_elem * ga->plan.nn * ga->plan.nn;
  int64_t *d = nullptr;
  int rc;
  if ((rc = dev_alloc(ctx, &d, n))) return rc;
  rc = slot_map(ga, d);
  if (!rc && n) FEMB_CUDA_OK(ctx, cudaMemcpyAsync(slot, d, sizeof(int64_t) * n, cudaMemcpyDeviceToHost, ctx->stream));
  dev_free(ctx, d);
  FEMB_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  return rc;
}

int femb200_isoparametric_strains(femb200_assembler *ga, const femb200_elem_desc *desc, const void *conn,
                                  int32_t conn_is_int64, int32_t conn_on_device, int64_t n_elem, const double *u,
                                  double *strains) {
  if (!ga || n_elem < 0 || !u || (n_elem > 0 && (!conn || !strains))) return FEMB200_ERR_BAD_ARG;
  femb200_ctx *ctx = ga->ctx;
  cudaStream_t st = ctx->stream;
  FEMB_CUDA_OK(ctx, cudaSetDevice(ctx->device));
  ga->err_kind = 0; ga->err_elem = -1; ga->err_qp = -1;
  ElemShape sh;
  std::vector<double> blob;
  int rc = prepare(ga, desc, &sh, &blob);
  if (rc) { ga->err_kind = rc; return rc; }
  *ctx->h_err = kNoError;
  if ((rc = reset_call_scalars(ctx))) { ga->err_kind = rc; return rc; }
  int32_t *conn32 = nullptr, *d_dofmap = nullptr;
  void *staged = nullptr;
  double *d_tab = nullptr, *d_u = nullptr, *d_s = nullptr;
  const int64_t ndof = ga->K.n_rows;
  const size_t ns = (size_t)n_elem * sh.nqp * sh.dimC;
  ScratchScope scratch_scope(ctx); // conn32 + staged connectivity
  rc = stage_conn(ga, conn, conn_is_int64 != 0, conn_on_device != 0, n_elem * sh.nn, sh.nn, &conn32, &staged);
  if (!rc) rc = dev_alloc(ctx, &d_tab, blob.size());
  if (!rc) rc = dev_alloc(ctx, &d_u, (size_t)ndof);
  if (!rc) rc = dev_alloc(ctx, &d_s, ns);
  if (!rc) rc = dev_alloc(ctx, &d_dofmap, kMaxNdn);
  if (!rc) {
    cudaError_t e = cudaMemcpyAsync(d_tab, blob.data(), sizeof(double) * blob.size(), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_u, u, sizeof(double) * ndof, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_dofmap, desc->dofmap, sizeof(int32_t) * sh.ndn, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(ctx->h_err, ctx->d_err, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) { ctx->last_cuda_error = cudaGetErrorString(e); rc = FEMB200_ERR_CUDA; }
  }
  if (!rc) rc = decode_error(ga); // out-of-range node ids
  if (!rc) rc = strains_phase(ga, sh, d_tab, conn32, n_elem, d_dofmap, d_u, d_s);
  if (!rc) {
    cudaError_t e = cudaMemcpyAsync(ctx->h_err, ctx->d_err, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess && ns) e = cudaMemcpyAsync(strains, d_s, sizeof(double) * ns, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) { ctx->last_cuda_error = cudaGetErrorString(e); rc = FEMB200_ERR_CUDA; }
  }
  if (!rc) rc = decode_error(ga);
  dev_free(ctx, conn32); dev_free(ctx, d_tab); dev_free(ctx, d_u); dev_free(ctx, d_s); dev_free(ctx, d_dofmap);
  cudaStreamSynchronize(st);
  if (rc) ga->err_kind = rc;
  return rc;
}

int femb200_probe_peaks(femb200_ctx *ctx, int32_t reps, double out[2]) {
  if (!ctx || !out) return FEMB200_ERR_BAD_ARG;
  FEMB_CUDA_OK(ctx, cudaSetDevice(ctx->device));
  return probe_peaks(ctx, reps, out);
}

int femb200_phase_ms(const femb200_assembler *ga, float out[FEMB200_N_TIMERS]) {
  if (!ga || !out) return FEMB200_ERR_BAD_ARG;
  for (int i = 0; i < FEMB200_N_TIMERS; ++i) out[i] = ga->phase_ms[i];
  return 0;
}

} // extern "C"
