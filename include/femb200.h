/*
 * femb200.h -- C ABI of libfemb200.so: B200 (sm_100a) isoparametric stiffness assembly.
 *
 * This is the drop-in boundary for ONE path of soypat/go-fem:
 *     fem.NewGeneralAssembler -> (*GeneralAssembler).AddIsoparametric -> Ksolid
 * Each entry point names the reference interface it replaces (paths relative to the go-fem tree).
 * The Go side binds these with cgo; only plain pointers and sizes cross.
 *
 * Ownership: the caller owns every input buffer and it is only read during the call (cgo rule:
 * Go memory is never retained).  The library owns all device memory and the CSR until *_destroy.
 * Threading: one context per OS thread / assembler; calls on one assembler are serialised by the
 * caller, exactly like the reference (GeneralAssembler has no locks, assembler.go:13-18).
 * There is NO CPU fallback: every call fails with FEMB200_ERR_CUDA if no sm_100 device is usable.
 */
#ifndef FEMB200_H
#define FEMB200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define FEMB200_ABI_VERSION 1

/* Strain-displacement layouts = the fillers of constitution/solids/straindisplacement.go:5-58 and
 * thermal.go:16-58.  The Go `solids` package tags its constituter with one of these
 * (solids.isoconstituter is unexported, solids.go:8-12, so the kind cannot be discovered otherwise). */
enum femb200_bkind {
  FEMB200_B_XYZ = 0,            /* SetStrainDisplacementMatrixXYZ          d=3 ndn=3 dimC=6 */
  FEMB200_B_PLANE = 1,          /* SetStrainDisplacementMatrixPlane        d=2 ndn=2 dimC=3 */
  FEMB200_B_AXISYM = 2,         /* SetStrainDisplacementMatrixAxisymmetric d=2 ndn=2 dimC=4, scale=r */
  FEMB200_B_THERMAL = 3,        /* thermal.go:18-21,31-34  B=dN            ndn=1 dimC=d */
  FEMB200_B_THERMAL_AXISYM = 4  /* thermal.go:43-56                        d=2 ndn=1 dimC=2, scale=r */
};

/* Status codes.  1..4 are the per-element errors of assembler_isoparametric.go:95-107; the Go shim
 * formats the reference's exact message from (kind, elem, qp) -- see femb200_last_error. */
enum femb200_status {
  FEMB200_OK = 0,
  FEMB200_ERR_NEG_DET = 1,      /* :96  negative determinant of jacobian of element #%d */
  FEMB200_ERR_ZERO_DET = 2,     /* :98  zero determinant (dJac < 1e-12, absolute) */
  FEMB200_ERR_SOLVE = 3,        /* :102 form factor solve: cond_1(J) > 1e16 (mat.ConditionTolerance) */
  FEMB200_ERR_NAN_SCALE = 4,    /* :106 NaN scale returned by SetStrainDisplacementMatrix */
  FEMB200_ERR_NODE_OOB = 8,     /* node index outside [0, n_nodes): Go slice-bounds panic, assembler.go:137 */
  FEMB200_ERR_BAD_ARG = 20,     /* malformed descriptor (nil pointers, sizes out of range) */
  FEMB200_ERR_UNSUPPORTED = 21, /* (d, nn, ndn, dimC, bkind) combination with no compiled kernel */
  FEMB200_ERR_CUDA = 22,        /* CUDA runtime failure; text via femb200_last_cuda_error */
  FEMB200_ERR_TOO_LARGE = 23,   /* index width exceeded (row with > 65535 node blocks, nnz/row > 2^31) */
  FEMB200_ERR_PATTERN = 24      /* femb200_csr_add_entries_device: an entry is not in the pattern of K */
};

/* Everything AddIsoparametric evaluates ONCE per call on the host (assembler_isoparametric.go:29-71)
 * crosses the boundary as tables, so user-defined fem.Isoparametric elements keep working:
 *   d     = len(elemT.BasisDiff(0)) / elemT.LenNodes()                      (:40)
 *   nn    = elemT.LenNodes(), ndn = elemT.Dofs().Count()                    (:44-46)
 *   Npg   [nqp][nn]     = elemT.Basis(upg[q])                               (:69)
 *   dNpg  [nqp][d][nn]  = elemT.BasisDiff(upg[q])   row-major               (:70)
 *   wpg   [nqp]         = quadrature weights                                (:56)
 *   C     [dimC][dimC]  = c.Constitutive() copied dense, row-major          (:29,:63)
 *   dofmap[ndn]         = mapdofs(modelDofs, elemT.Dofs())                  (fem.go:155-172)
 * model_dofs_per_node = modelDofs.Count(); the reference requires it to equal ndn
 * (fem.go:138 + assembler.go:148-150 panic "bad length"), the shim checks that before calling. */
typedef struct femb200_elem_desc {
  int32_t d, nn, ndn, nqp, dimC, bkind;
  int32_t model_dofs_per_node;
  int32_t reserved;
  const double *Npg;
  const double *dNpg;
  const double *wpg;
  const double *C;
  const int32_t *dofmap;
} femb200_elem_desc;

typedef struct femb200_ctx femb200_ctx;             /* one CUDA device + stream + memory pool */
typedef struct femb200_assembler femb200_assembler; /* twin of fem.GeneralAssembler (assembler.go:13-18) */

int femb200_abi_version(void);

/* Context on CUDA device `device` (ordinal).  Fails with FEMB200_ERR_CUDA when there is no GPU. */
int femb200_ctx_create(int device, femb200_ctx **out);
void femb200_ctx_destroy(femb200_ctx *ctx);
/* Runs on an externally owned stream (cudaStream_t as void*), e.g. torch's current stream. */
int femb200_ctx_set_stream(femb200_ctx *ctx, void *cuda_stream);
/* Several contexts of ONE device may run add calls at the same time (one host thread and one stream each) once
 * every one of them has been switched to concurrent mode: the per-call element tables then live in global memory
 * instead of the device-wide __constant__ bank.  (A rank of the multi-GPU path assembles its owned row block and
 * its outgoing interface rows this way, the small second call filling the gaps of the first.) */
int femb200_ctx_set_concurrent(femb200_ctx *ctx, int32_t on);
/* Returns the context's per-call scratch arena to the device (it is kept between calls so that a repeated assembly
 * allocates nothing; after a large one-off assembly -- the element-matrix staging of a Hexa20 mesh is 28.8 KB per
 * element -- a caller that goes on to solve on the same GPU wants the memory back).  K and the cached plans are untouched. */
int femb200_ctx_trim(femb200_ctx *ctx);
const char *femb200_last_cuda_error(const femb200_ctx *ctx);

/* fem.NewGeneralAssembler(nodes []r3.Vec, modelDofs) (assembler.go:21-28).
 * nodes: n_nodes x {X,Y,Z} float64, 24-byte stride (the memory of a Go []r3.Vec).
 * nodes_on_device != 0: `nodes` is a device pointer that stays valid for the assembler's life. */
int femb200_assembler_create(femb200_ctx *ctx, const double *nodes, int64_t n_nodes,
                             int32_t model_dofs_per_node, int32_t nodes_on_device,
                             femb200_assembler **out);
void femb200_assembler_destroy(femb200_assembler *ga);

/* (*GeneralAssembler).AddIsoparametric (assembler_isoparametric.go:28-123) for n_elem elements whose
 * connectivity the shim flattened from getElement(i) (fem.go:141): conn[n_elem][nn], 32- or 64-bit
 * (Go int), on the host or already on the device.  Symbolic phase, numeric phase and the deterministic
 * row reduction all run on the GPU; the result is accumulated into K (repeated calls add up, :121).
 * On an element error K is left untouched (:118-120) and the lowest failing element is reported. */
int femb200_add_isoparametric(femb200_assembler *ga, const femb200_elem_desc *desc,
                              const void *conn, int32_t conn_is_int64, int32_t conn_on_device,
                              int64_t n_elem);

/* The same with a halo: the LAST n_pattern_only elements of conn take part in the symbolic phase only (their
 * node pairs enter the CSR pattern with value 0) and contribute no values; they are not validated geometrically.
 * Multi-GPU: a rank lists the neighbouring ranks' elements that touch its rows, so that the rows it owns have
 * their final pattern before the interface values arrive. */
int femb200_add_isoparametric_halo(femb200_assembler *ga, const femb200_elem_desc *desc,
                                   const void *conn, int32_t conn_is_int64, int32_t conn_on_device,
                                   int64_t n_elem, int64_t n_pattern_only);

/* After a non-zero status: which error, the lowest failing element index and its first failing
 * quadrature point (-1 when not applicable). */
int femb200_last_error(const femb200_assembler *ga, int32_t *kind, int64_t *elem, int32_t *qp);

/* Numeric-only re-assembly on the cached pattern (the reference rebuilds everything on every call):
 * recomputes the values of K for the connectivity of the ONE AddIsoparametric call this assembler has committed, with a
 * new descriptor (material, quadrature order, tables; same element shape) and optionally new node coordinates
 * (n_nodes x 3, NULL = keep).  The symbolic phase is skipped: geometry kernel + node-row kernel only.  Same error
 * semantics as femb200_add_isoparametric (K untouched on error).  FEMB200_ERR_BAD_ARG if more than one call was merged. */
int femb200_reassemble(femb200_assembler *ga, const femb200_elem_desc *desc, const double *nodes, int32_t nodes_on_device);

/* (*GeneralAssembler).TotalDofs (assembler.go:34-37) and the size of Ksolid() (assembler.go:31). */
int64_t femb200_total_dofs(const femb200_assembler *ga);
int64_t femb200_nnz(const femb200_assembler *ga);
/* K as canonical CSR (rows and columns ascending, every emitted (row,col) pair kept, duplicates
 * summed in ascending element index then call order): indptr[ndof+1] int64, indices[nnz] int32,
 * values[nnz] float64, copied into caller memory.  NULL pointers are skipped. */
int femb200_csr_copy(const femb200_assembler *ga, int64_t *indptr, int32_t *indices, double *values);
/* K at node-block granularity (additive API: the 4-byte scalar column indices of femb200_csr_copy are a pure function of
 * the block pattern, 120 of the 366 MB a config-1 result sends over PCIe).  Valid while K holds exactly ONE
 * AddIsoparametric call with the identity dof map (else FEMB200_ERR_UNSUPPORTED): n_nodes block rows,
 * blkptr[n_nodes+1] int64, blkcols[nblk] int32 = column NODE of every block (ascending within a row),
 * values[nnz] exactly as in femb200_csr_copy.  Scalar row (I, i) of K has the columns mdpn*blkcols[p] + j,
 * p in [blkptr[I], blkptr[I+1]), j in [0, mdpn), and starts at values[mdpn*mdpn*blkptr[I] + i*mdpn*(blkptr[I+1]-blkptr[I])].
 * NULL pointers are skipped; femb200_bsr_blocks returns nblk.  On an assembler with a column map (the rank-local assembler of
 * femb200_dist_*) blkcols are rank-LOCAL node ids: global column node = l2g[blkcols[p]] with l2g from femb200_dist_local_nodes. */
int64_t femb200_bsr_blocks(const femb200_assembler *ga);
int femb200_bsr_copy(const femb200_assembler *ga, int64_t *blkptr, int32_t *blkcols, double *values);
/* The same arrays left on the device (valid until the next add/destroy). */
int femb200_csr_device(const femb200_assembler *ga, const int64_t **indptr, const int32_t **indices,
                       const double **values);
/* lap.Sparse.At(i,j) (call sites example_test.go:38, patch_test.go:80): one entry, 0 if absent. */
int femb200_csr_at(const femb200_assembler *ga, int64_t i, int64_t j, double *out);
/* K*x on the device, x and y on the host (forces.MulVec(K,u), patch_test.go:92).  (Assemblers with a column map hold
 * global column ids over local rows: femb200_csr_at / _matvec / _slice address K by its own row numbering only.) */
int femb200_csr_matvec(const femb200_assembler *ga, const double *x, double *y);

/* K += triplets (assembler.go:225-231: AddElement3 writes its 12x12 beam blocks into the same ksolid with
 * Set(At+v), so models mix isoparametric solids and beams in one K).  rows / cols / vals are n host triplets in any
 * order, duplicates allowed (summed in input order); the pattern of K grows by the new (row, col) pairs. */
int femb200_add_coo(femb200_assembler *ga, int64_t n, const int64_t *rows, const int64_t *cols, const double *vals);

/* ---- multi-GPU row-block exchange (one process per GPU; the collective itself is NCCL all-to-all driven by
 * the host layer, these are the device-side pack/unpack steps either side of it) -------------------------
 * K += A, A given on the device as n_rows scattered rows: rows[n_rows] ascending global row ids,
 * indptr[n_rows+1] with indptr[0] == 0, indices / values packed row after row.  Pattern union, values
 * added after the existing ones (fixed order: the caller feeds pieces in ascending source rank). */
int femb200_csr_accumulate_rows_device(femb200_assembler *ga, int64_t n_rows, const int64_t *rows,
                                       const int64_t *indptr, const int32_t *indices, const double *values);
/* Row-block restriction: from now on only the rows of nodes with node_mask[node] != 0 (device bytes, one per node,
 * NULL = all) are assembled; the other rows stay empty.  A rank assembles its owned row block with one assembler
 * and the interface rows it must send with another one, so no row is computed and then thrown away. */
int femb200_assembler_set_node_mask_device(femb200_assembler *ga, const uint8_t *node_mask);
/* Local node numbering of a rank (multi-GPU): the assembler is created over the rank's LOCAL nodes (own + ghost, ascending
 * global id) and node_l2g[local node] = global node (device int32, strictly increasing; NULL = identity).  Rows of K stay
 * local (row r = local node r / mdpn), K.indices become GLOBAL dof ids mdpn * node_l2g[node] + dof: every per-call array is
 * sized by the rank's share of the mesh, not by the global mesh. */
int femb200_assembler_set_column_map_device(femb200_assembler *ga, const int32_t *node_l2g, int64_t n_global_nodes);
/* K[rows] += A in place: like femb200_csr_accumulate_rows_device but every (row, column) of A must already be
 * in the pattern of K (true for owned rows assembled with their halo).  No allocation, no pattern change;
 * FEMB200_ERR_PATTERN if an entry is missing (checked in a first pass: K is untouched on failure). */
int femb200_csr_add_entries_device(femb200_assembler *ga, int64_t n_rows, const int64_t *rows,
                                   const int64_t *indptr, const int32_t *indices, const double *values);
/* Planned interface exchange (no reference counterpart; go-fem_b200/distributed.py): the partition plan fixes, once
 * per mesh, which rows every rank sends to every owner and how long they are, so the per-assembly exchange is
 * pack -> ONE collective -> add with no size round trip and no host synchronisation.  Both calls only enqueue
 * work on the context stream; *mismatch (device uint64, may be NULL) counts rows / entries that contradict the plan.
 *  pack: entries of row rows[w] of K -> out_indices / out_values [out_indptr[w], out_indptr[w+1])
 *  add : K[rows[w], indices[k]] += values[k], k in [indptr[w], indptr[w+1]); one call per source rank, ascending. */
int femb200_csr_pack_rows_async(femb200_assembler *ga, int64_t n_rows, const int64_t *rows, const int64_t *out_indptr,
                                int32_t *out_indices, double *out_values, uint64_t *mismatch);
int femb200_csr_add_entries_async(femb200_assembler *ga, int64_t n_rows, const int64_t *rows, const int64_t *indptr,
                                  const int32_t *indices, const double *values, uint64_t *mismatch);
/* Drops every row whose keep[row] == 0 (device byte mask, one per scalar row): after sending the rows it
 * does not own, a rank keeps exactly its row block of the distributed K. */
int femb200_csr_keep_rows_device(femb200_assembler *ga, const uint8_t *keep);

/* ---- multi-GPU behind the C ABI: device-side partition, owner-computes row blocks --------------------------
 * The reference has ONE assembler and ONE call (assembler.go:21-31, assembler_isoparametric.go:28).  A femb200_dist is rank
 * `rank` of `world` (one process per GPU; or one host thread per GPU inside a femb200_group): it receives the WHOLE mesh,
 * partitions it on its device -- element order (0 = as given, 1 = Morton order of the element centroids, 21 bits per axis),
 * contiguous chunks of equal element count, a node (row block) belongs to the lowest rank touching it -- and assembles the
 * rows it owns from every element that touches them (its chunk's plus the halo of the other chunks), in ascending global
 * element index: the row block is bit-identical to the same rows of a single-GPU K, and no data-path collective exists.
 * K of femb200_dist_assembler(): rows = local nodes (femb200_dist_local_nodes gives their global ids and which are owned,
 * rows of non-owned nodes are empty), column indices = GLOBAL dof ids. */
typedef struct femb200_dist femb200_dist;
int femb200_dist_create(femb200_ctx *ctx, int32_t rank, int32_t world, const double *nodes, int32_t nodes_on_device, int64_t n_nodes,
                        int32_t model_dofs_per_node, const void *conn, int32_t conn_is_int64, int32_t conn_on_device, int64_t n_elem,
                        int32_t nn, int32_t order_mode, femb200_dist **out);
void femb200_dist_destroy(femb200_dist *d);
/* (*GeneralAssembler).AddIsoparametric for this rank's row block; same error semantics (femb200_dist_last_error reports the
 * lowest failing element of THIS rank in global numbering; the lowest over all ranks is what the reference would report). */
int femb200_dist_add_isoparametric(femb200_dist *d, const femb200_elem_desc *desc);
int femb200_dist_last_error(const femb200_dist *d, int32_t *kind, int64_t *elem, int32_t *qp);
/* Empties the rank's K (a fresh local assembler over the same partition): the next femb200_dist_add_isoparametric starts
 * from NewGeneralAssembler again instead of accumulating. */
int femb200_dist_reset(femb200_dist *d);
femb200_assembler *femb200_dist_assembler(femb200_dist *d);
/* out = { local nodes, owned nodes, local elements (own + halo), elements of the rank's chunk, halo elements, nnz of the row
 * block, global nodes, global elements }; *partition_ms = device time of the partition (CUDA events). */
int femb200_dist_info(const femb200_dist *d, int64_t out[8], float *partition_ms);
int femb200_dist_local_nodes(const femb200_dist *d, int32_t *node_l2g /* [local nodes] */, uint8_t *owned /* [local nodes] */);
/* One process, several devices -- what a cgo caller uses: a context, a femb200_dist and a host thread per
 * device; nodes / conn are host arrays (a Go []r3.Vec and flattened []int). */
typedef struct femb200_group femb200_group;
int femb200_group_create(const int32_t *devices, int32_t ndev, const double *nodes, int64_t n_nodes, int32_t model_dofs_per_node,
                         const void *conn, int32_t conn_is_int64, int64_t n_elem, int32_t nn, int32_t order_mode, femb200_group **out);
void femb200_group_destroy(femb200_group *g);
int32_t femb200_group_size(const femb200_group *g);
femb200_dist *femb200_group_rank(femb200_group *g, int32_t rank);
/* AddIsoparametric on every device at once; on an element error returns the status of the rank holding the lowest failing
 * element (what the sequential reference loop reports, fem.go:140) and stores that rank in *rank_out (-1 when all succeeded). */
int femb200_group_add_isoparametric(femb200_group *g, const femb200_elem_desc *desc, int32_t *rank_out);

/* Synthetic meshes generated on the device (benchmarks and tests only; no reference counterpart): the BCC tetrahedral mesh
 * of an nx x ny x nz cell box, node for node and element for element what go-fem_b200/meshes.py bcc_tetra4_box builds on the
 * host (benchmark_test.go:34-35 uses manigold's BCC mesher: only its counts are pinned).  Arrays are allocated by the library
 * (device memory) and released with femb200_mesh_free. */
int femb200_mesh_bcc_tetra4(femb200_ctx *ctx, int32_t nx, int32_t ny, int32_t nz, double h, int32_t jittered, double **d_nodes,
                            int64_t *n_nodes, int32_t **d_conn, int64_t *n_elem);
void femb200_mesh_free(femb200_ctx *ctx, void *device_ptr);
/* Copies `bytes` of a device array of this library (a generated mesh, a CSR array of femb200_csr_device) into host memory. */
int femb200_copy_to_host(femb200_ctx *ctx, void *host_dst, const void *device_src, int64_t bytes);

/* Element -> CSR slot map of the last add call (north_star subsystem 1):
 * slot[e][a][b] = index into the node-block CSR of block (node a of e, node b of e); the scalar entry
 * of dof pair (i,j) sits at values[indptr[mdpn*node_a+dofmap[i]] + pos*nu + j'].  Copied to the host. */
int femb200_last_slot_map(const femb200_assembler *ga, int64_t *slot /* [n_elem][nn][nn] */);

/* (*GeneralAssembler).IsoparametricStrains (assembler_isoparametric.go:126-213): strains[n_elem][nqp][dimC]
 * = B(q) * u[elemDofs], host in / host out. */
int femb200_isoparametric_strains(femb200_assembler *ga, const femb200_elem_desc *desc,
                                  const void *conn, int32_t conn_is_int64, int32_t conn_on_device,
                                  int64_t n_elem, const double *u, double *strains);

/* lap.Slice(K, rows, cols) (constitution/solids/patch_test.go:80,87: K(free,free), K(free,fixed) from fem.Fixity's
 * FreeDofs / FixedDofs, fem.go:175-232), extracted on the device: the selected rows of K restricted to the selected
 * columns, as a CSR whose column indices are positions in `cols` (cols ascending and unique; rows in any order).
 * Two-call pattern: with indptr = indices = values = NULL only *nnz is written; then indptr[n_rows+1], indices[nnz],
 * values[nnz] (host arrays) are filled. */
int femb200_csr_slice(const femb200_assembler *ga, int64_t n_rows, const int64_t *rows, int64_t n_cols, const int64_t *cols,
                      int64_t *indptr, int32_t *indices, double *values, int64_t *nnz);

/* Device-side phase timers of the last add call, milliseconds (CUDA events on the context stream):
 * [0] upload+validate [1] node->element adjacency [2] row pattern [3] CSR expand
 * [4] numeric + row reduction [5] merge into K [6] total device [7] numeric kernel launches in total
 * [8] of [4]: the batched per-element geometry kernel  [9] of [4]: the row kernel (B^T C B + reduction). */
#define FEMB200_N_TIMERS 10
int femb200_phase_ms(const femb200_assembler *ga, float out[FEMB200_N_TIMERS]);
/* Number of kernels this library launched since the context was created (bench.py "gpu_launches"). */
int64_t femb200_kernel_launches(const femb200_ctx *ctx);
/* Blocks until all work queued on the context stream is done. */
int femb200_sync(femb200_ctx *ctx);
/* Roofline denominators measured on this device (no reference counterpart; used by bench.py only):
 * out[0] = FP64 vector peak in TFLOP/s (independent DFMA chains, all SMs), out[1] = HBM copy bandwidth in GB/s
 * (read + write bytes of a 1 GiB device-to-device copy kernel), both best of `reps` CUDA-event timings. */
int femb200_probe_peaks(femb200_ctx *ctx, int32_t reps, double out[2]);

#ifdef __cplusplus
}
#endif
#endif
