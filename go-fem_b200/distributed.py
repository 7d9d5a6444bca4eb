"""Round-1 row-exchange variant of the multi-GPU assembly (kept for its gloo tests; the product path is the owner-computes
scheme behind the C ABI, csrc/dist.cu: femb200_dist_* / femb200_group_*, DESIGN.md section 7, which needs no collective).

Element partition in space-filling-curve order, one process per GPU, row-block ownership, and the exchange of interface
rows (the NCCL reduce-scatter style scheme BASELINE.json's north_star names; measured at 60-68 % weak-scaling efficiency):

    rank p:  K_p  = assemble(elements of chunk p)                   local symbolic + numeric on its GPU
             for q != p: send the rows of K_p owned by q            pack (device gathers) -> all-to-all (NCCL / NVLink)
             K_p  <- rows owned by p                                femb200_csr_keep_rows_device
             K_p += piece from q, q ascending                       femb200_csr_add_entries_device (in place: the local symbolic
                                                                    phase listed the halo elements, so owned rows already have
                                                                    their final pattern) or ..._accumulate_rows_device (union);
                                                                    fixed order => bit-reproducible run to run
    => every rank owns a row block of the distributed CSR K, in the caller's global DOF numbering.

Row ownership: a node belongs to the lowest rank that has an element touching it (unreferenced nodes to rank
0), so with a space-filling-curve element order only the partition surfaces are exchanged.

The host-side logic here (partition, ownership, send lists, the all-to-all of ragged CSR pieces) is backend
agnostic: tests run it on CPU with gloo and world_size 2; on the B200 box the same code moves device tensors
over NCCL.  Local assembly and the two device-side pack/unpack steps are injected (`LocalBackend`).
"""
import numpy as np


# ----------------------------------------------------------------------------- partition (host, once per mesh)
def morton_order(nodes, conn, bits=21):
    """Element permutation by the Morton (Z-order) key of the element centroid, 21 bits per axis."""
    c = nodes[conn].mean(axis=1)
    lo, hi = c.min(axis=0), c.max(axis=0)
    span = np.where(hi > lo, hi - lo, 1.0)
    q = np.minimum(((c - lo) / span * (2 ** bits - 1)).astype(np.uint64), np.uint64(2 ** bits - 1))

    def spread(v):
        v = v & np.uint64(0x1FFFFF)
        v = (v | (v << np.uint64(32))) & np.uint64(0x1F00000000FFFF)
        v = (v | (v << np.uint64(16))) & np.uint64(0x1F0000FF0000FF)
        v = (v | (v << np.uint64(8))) & np.uint64(0x100F00F00F00F00F)
        v = (v | (v << np.uint64(4))) & np.uint64(0x10C30C30C30C30C3)
        v = (v | (v << np.uint64(2))) & np.uint64(0x1249249249249249)
        return v

    key = spread(q[:, 0]) | (spread(q[:, 1]) << np.uint64(1)) | (spread(q[:, 2]) << np.uint64(2))
    return np.argsort(key, kind="stable")


def chunk_bounds(n_elem, world):
    """Contiguous chunks of (almost) equal element count."""
    return [(n_elem * r) // world for r in range(world + 1)]


def node_owner(n_nodes, conn, bounds):
    """owner[node] = lowest rank with an element touching the node (rank 0 for unreferenced nodes)."""
    world = len(bounds) - 1
    owner = np.full(n_nodes, world, dtype=np.int32)
    for r in range(world - 1, -1, -1):          # descending: the lowest rank writes last
        touched = np.unique(conn[bounds[r]:bounds[r + 1]])
        owner[touched] = r
    owner[owner == world] = 0
    return owner


class ExchangePlan:
    """What rank `rank` sends: for every other rank q the ascending scalar rows (global ids) of nodes that
    this rank's elements touch and q owns; plus the keep mask of its own row block."""

    def __init__(self, n_nodes, conn, bounds, rank, mdpn, owner=None):
        world = len(bounds) - 1
        self.world, self.rank, self.mdpn = world, rank, mdpn
        self.owner = node_owner(n_nodes, conn, bounds) if owner is None else owner
        touched = np.unique(conn[bounds[rank]:bounds[rank + 1]])
        self.send_rows = []
        for q in range(world):
            if q == rank:
                self.send_rows.append(np.zeros(0, dtype=np.int64))
                continue
            nd = touched[self.owner[touched] == q].astype(np.int64)
            rows = (nd[:, None] * mdpn + np.arange(mdpn, dtype=np.int64)[None, :]).ravel()
            self.send_rows.append(rows)
        self.keep_mask = np.repeat(self.owner == rank, mdpn).astype(np.uint8)
        self.owned_rows = np.nonzero(self.keep_mask)[0].astype(np.int64)
        # halo: elements of other chunks that touch a node this rank owns.  Listed as pattern-only elements in the
        # local symbolic phase, they give the owned rows their FINAL pattern, so incoming interface values are
        # added in place (no pattern union, no reallocation).
        mine = self.owner == rank
        self.own_node_mask = mine.astype(np.uint8)
        own_conn = conn[bounds[rank]:bounds[rank + 1]]
        foreign = ~mine
        self.send_node_mask = np.zeros(n_nodes, dtype=np.uint8)
        self.send_node_mask[touched[foreign[touched]]] = 1
        # own elements that touch a node owned by someone else: the only ones that produce interface rows
        self.send_elems = np.nonzero(foreign[own_conn].any(axis=1))[0]
        other = np.ones(conn.shape[0], dtype=bool)
        other[bounds[rank]:bounds[rank + 1]] = False
        self.halo_elems = np.nonzero(other & mine[conn].any(axis=1))[0]

    def local_connectivity(self, conn, bounds):
        """(connectivity to hand to the local assembler, number of trailing pattern-only elements)"""
        own = conn[bounds[self.rank]:bounds[self.rank + 1]]
        if self.halo_elems.size == 0:
            return own, 0
        return np.concatenate([own, conn[self.halo_elems]], axis=0), int(self.halo_elems.size)


class LocalNumbering:
    """Rank-local node numbering: the ascending list of the global nodes the rank's elements (own + halo) touch.  The
    device assembler is created over these nodes only, so every per-call array is sized by the rank's share of the mesh;
    rows of the local K are local, its column indices are global (femb200_assembler_set_column_map_device)."""

    def __init__(self, nodes, conn_global_rows):
        self.l2g = np.unique(conn_global_rows)
        self.nodes = np.ascontiguousarray(nodes[self.l2g])

    def conn(self, conn_global_rows):
        return np.ascontiguousarray(np.searchsorted(self.l2g, conn_global_rows).astype(np.int32))

    def node_mask(self, global_mask):
        return np.ascontiguousarray(global_mask[self.l2g].astype(np.uint8))

    def rows(self, rows_global, mdpn):
        node, dof = rows_global // mdpn, rows_global % mdpn
        ln = np.searchsorted(self.l2g, node)
        assert np.array_equal(self.l2g[np.minimum(ln, self.l2g.size - 1)], node), "row of a node outside the local numbering"
        return (ln * mdpn + dof).astype(np.int64)


# ----------------------------------------------------------------------------- planned exchange (host, once per mesh)
def interface_block_rows(n_nodes, conn, bounds, owner, p):
    """Interface pattern of sender p, exact: for every node J that p's elements touch and another rank owns, the
    number of distinct column nodes p's elements couple it to.  Returns (nodes ascending, counts)."""
    own = conn[bounds[p]:bounds[p + 1]]
    foreign = owner[own] != p
    ei, ai = np.nonzero(foreign)
    if ei.size == 0:
        return np.zeros(0, dtype=np.int64), np.zeros(0, dtype=np.int64)
    rown = own[ei, ai].astype(np.int64)                                    # foreign row node of the pair
    keys = np.unique((rown[:, None] * n_nodes + own[ei].astype(np.int64)).ravel())
    rn, cnt = np.unique(keys // n_nodes, return_counts=True)
    return rn, cnt.astype(np.int64)


class InterfacePlan:
    """Everything the per-assembly exchange needs, fixed by (mesh, partition): for every peer q the scalar rows this
    rank sends to q with their exact lengths, and for every source p the rows it receives, as packed-CSR offsets.
    One payload per peer: [values f64 x nnz | indices int32 x nnz]; the split sizes are known on both sides, so the
    exchange is ONE all-to-all with no size round trip.  `nu` = scalar dofs per node block (identity dof map)."""

    def __init__(self, xplan, n_nodes, conn, bounds, nu, dist=None):
        self.world, self.rank, self.nu = xplan.world, xplan.rank, nu
        rn, cnt = interface_block_rows(n_nodes, conn, bounds, xplan.owner, self.rank)
        own_of = xplan.owner[rn] if rn.size else np.zeros(0, dtype=np.int32)
        self.send = []                                                     # per peer: (rows int64, indptr int64)
        for q in range(self.world):
            sel = own_of == q
            self.send.append(self._rows_ptr(rn[sel], cnt[sel], nu))
        mine = [(r, p) for (r, p) in self.send]
        if dist is not None and self.world > 1:
            allp = [None] * self.world
            dist.all_gather_object(allp, mine)
            self.recv = [allp[p][self.rank] for p in range(self.world)]
        else:                                                              # single process (tests): derive the peers' plans locally
            self.recv = []
            for p in range(self.world):
                rnp, cntp = interface_block_rows(n_nodes, conn, bounds, xplan.owner, p)
                sel = xplan.owner[rnp] == self.rank if rnp.size else np.zeros(0, dtype=bool)
                self.recv.append(self._rows_ptr(rnp[sel], cntp[sel], nu) if p != self.rank else self._rows_ptr(rnp[:0], cntp[:0], nu))
        self.send_nnz = [int(p[-1]) for _, p in self.send]
        self.recv_nnz = [int(p[-1]) for _, p in self.recv]
        self.send_bytes = [(12 * n + 15) // 16 * 16 for n in self.send_nnz]   # 16-byte aligned payloads
        self.recv_bytes = [(12 * n + 15) // 16 * 16 for n in self.recv_nnz]

    @staticmethod
    def _rows_ptr(nodes_, cnt, nu):
        rows = (nodes_[:, None] * nu + np.arange(nu, dtype=np.int64)[None, :]).ravel()
        lens = np.repeat(cnt * nu, nu)
        ptr = np.zeros(rows.size + 1, dtype=np.int64)
        np.cumsum(lens, out=ptr[1:])
        return rows.astype(np.int64), ptr


class PlannedExchange:
    """Per-rank state of the planned exchange (torch tensors on `device`); pack / add are delegated to the local
    backends (the C ABI kernels on the GPU, numpy in the gloo tests)."""

    def __init__(self, iplan, device, local=None):
        """local: LocalNumbering of the rank -- the plan's global row ids are translated to the local rows of its K."""
        import torch
        self.p, self.device = iplan, device
        t = lambda a: torch.as_tensor(a, device=device)
        loc = (lambda r: local.rows(r, iplan.nu)) if local is not None else (lambda r: r)
        self.send_rows = [t(loc(r)) for r, _ in iplan.send]
        self.send_ptr = [t(p) for _, p in iplan.send]
        self.recv_rows = [t(loc(r)) for r, _ in iplan.recv]
        self.recv_ptr = [t(p) for _, p in iplan.recv]
        self.sendbuf = torch.zeros(max(sum(iplan.send_bytes), 16), dtype=torch.uint8, device=device)
        self.recvbuf = torch.zeros(max(sum(iplan.recv_bytes), 16), dtype=torch.uint8, device=device)
        self.mismatch = torch.zeros(1, dtype=torch.int64, device=device)
        self.soff = np.concatenate([[0], np.cumsum(iplan.send_bytes)]).astype(np.int64)
        self.roff = np.concatenate([[0], np.cumsum(iplan.recv_bytes)]).astype(np.int64)

    @staticmethod
    def _views(buf, off, n):
        import torch
        val = buf[off:off + 8 * n].view(torch.float64)
        idx = buf[off + 8 * n:off + 12 * n].view(torch.int32)
        return idx, val

    def run(self, owned, outgoing, dist, after_pack=None, after_collective=None):
        """outgoing (backend holding the interface rows) -> peers -> added into owned, ascending source rank.
        On the GPU everything is only enqueued (pack on outgoing's stream, the collective on torch's current stream,
        the adds on owned's stream); the two hooks let the caller order those streams (stream.wait_stream)."""
        p = self.p
        for q in range(p.world):
            if q == p.rank or self.send_rows[q].numel() == 0:
                continue
            idx, val = self._views(self.sendbuf, int(self.soff[q]), p.send_nnz[q])
            outgoing.pack_rows_into(self.send_rows[q], self.send_ptr[q], idx, val, self.mismatch)
        if after_pack is not None:
            after_pack()
        dist.all_to_all_single(self.recvbuf[:int(self.roff[-1])], self.sendbuf[:int(self.soff[-1])],
                               output_split_sizes=[int(b) for b in p.recv_bytes], input_split_sizes=[int(b) for b in p.send_bytes])
        if after_collective is not None:
            after_collective()
        for src in range(p.world):                   # fixed order: ascending source rank => bit-reproducible
            if src == p.rank or self.recv_rows[src].numel() == 0:
                continue
            idx, val = self._views(self.recvbuf, int(self.roff[src]), p.recv_nnz[src])
            owned.add_entries_from(self.recv_rows[src], self.recv_ptr[src], idx, val, self.mismatch)
        return int(sum(12 * n for n in p.send_nnz))

    def mismatches(self):
        return int(self.mismatch.item())


# ----------------------------------------------------------------------------- ragged CSR pieces (torch tensors)
def gather_rows(indptr, indices, values, rows):
    """Packs the given (ascending) rows of a CSR held in torch tensors: returns (lens, indices, values)."""
    import torch
    if rows.numel() == 0:
        return rows.new_zeros(0), indices.new_zeros(0), values.new_zeros(0)
    start = indptr[rows]
    lens = indptr[rows + 1] - start
    total = int(lens.sum().item())
    if total == 0:
        return lens, indices.new_zeros(0), values.new_zeros(0)
    off = torch.cumsum(lens, 0) - lens
    gidx = torch.repeat_interleave(start - off, lens) + torch.arange(total, device=rows.device, dtype=rows.dtype)
    return lens, indices[gidx], values[gidx]


def all_to_all_ragged(pieces, world, rank, device, dist):
    """pieces[q] = (rows int64, lens int64, indices int32, values f64) to send to q.  Returns the pieces
    received from every rank, same tuple layout.  Two collectives: the sizes, then ONE byte payload per peer
    ([rows | lens | values | indices], 8-byte fields first) -- the interface is small, so latency dominates."""
    import torch
    sizes = torch.tensor([[p[0].numel(), p[2].numel()] for p in pieces], dtype=torch.int64, device=device)
    rsizes = torch.empty_like(sizes)
    dist.all_to_all_single(rsizes, sizes)
    rs = rsizes.cpu().numpy()
    ss = sizes.cpu().numpy()

    def nbytes(a):
        return int(a[0]) * 16 + int(a[1]) * 12

    def as_bytes(t):
        if t.numel() == 0:
            return torch.empty(0, dtype=torch.uint8, device=device)
        return t.contiguous().view(torch.uint8)

    send = torch.cat([torch.cat([as_bytes(p[0].to(torch.int64)), as_bytes(p[1].to(torch.int64)), as_bytes(p[3].to(torch.float64)),
                                 as_bytes(p[2].to(torch.int32))]) for p in pieces])
    recv = torch.empty(sum(nbytes(r) for r in rs), dtype=torch.uint8, device=device)
    dist.all_to_all_single(recv, send, output_split_sizes=[nbytes(r) for r in rs], input_split_sizes=[nbytes(x) for x in ss])
    out, off = [], 0
    for q in range(world):
        nr, nz = int(rs[q, 0]), int(rs[q, 1])
        rows = recv[off:off + 8 * nr].view(torch.int64); off += 8 * nr
        lens = recv[off:off + 8 * nr].view(torch.int64); off += 8 * nr
        val = recv[off:off + 8 * nz].view(torch.float64); off += 8 * nz
        idx = recv[off:off + 4 * nz].view(torch.int32); off += 4 * nz
        out.append((rows, lens, idx, val))
    return out


class LocalBackend:
    """Injected per-rank operations.  csr_tensors() -> (indptr, indices, values) torch views of the local K;
    keep_rows(mask_tensor); accumulate_rows(rows, indptr0, indices, values)."""

    def csr_tensors(self):
        raise NotImplementedError

    def keep_rows(self, mask):
        raise NotImplementedError

    def accumulate_rows(self, rows, indptr0, indices, values):
        raise NotImplementedError

    def add_entries(self, rows, indptr0, indices, values):
        """In-place add when the pattern already holds every entry; returns False when it does not."""
        return False


def exchange_and_own(backend, plan, device, dist, timings=None):
    """The collective step: after local assembly, leaves `backend` holding exactly the rank's row block.
    timings: optional dict filled with host-clock milliseconds per sub-step (synchronising; diagnostics only)."""
    import time
    import torch

    def mark(name, t0):
        if timings is not None:
            if device.type == "cuda":
                torch.cuda.synchronize()
            timings[name] = timings.get(name, 0.0) + 1e3 * (time.perf_counter() - t0)
        return time.perf_counter()

    t0 = time.perf_counter()
    indptr, indices, values = backend.csr_tensors()
    pieces = []
    for q in range(plan.world):
        rows = torch.as_tensor(plan.send_rows[q], device=device)
        lens, idx, val = gather_rows(indptr, indices, values, rows)
        pieces.append((rows, lens, idx, val))
    t0 = mark("pack", t0)
    received = all_to_all_ragged(pieces, plan.world, plan.rank, device, dist)
    t0 = mark("all_to_all", t0)
    backend.keep_rows(torch.as_tensor(plan.keep_mask, device=device))
    t0 = mark("keep_rows", t0)
    for q in range(plan.world):                      # fixed order: ascending source rank
        if q == plan.rank:
            continue
        rows, lens, idx, val = received[q]
        if rows.numel() == 0:
            continue
        ptr0 = torch.zeros(rows.numel() + 1, dtype=torch.int64, device=device)
        torch.cumsum(lens, 0, out=ptr0[1:])
        if not backend.add_entries(rows, ptr0, idx, val):     # owned rows carry their halo pattern: in place
            backend.accumulate_rows(rows, ptr0, idx, val)     # generic path: pattern union
    mark("accumulate", t0)
    return sum(int(p[3].numel()) * 12 + int(p[0].numel()) * 16 for p in pieces)   # bytes sent (values+indices, rows+lens)


class DeviceView:
    """Zero-copy torch view of library-owned device memory (via __cuda_array_interface__)."""

    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": typestr, "data": (int(ptr), False), "version": 2}


def device_tensor(ptr, n, typestr, device):
    import torch
    if n == 0 or not ptr:
        dt = {"<i8": torch.int64, "<i4": torch.int32, "<f8": torch.float64}[typestr]
        return torch.empty(0, dtype=dt, device=device)
    return torch.as_tensor(DeviceView(ptr, n, typestr), device=device)


class Femb200Backend(LocalBackend):
    """The GPU library as the local backend (RawAssembler from the package)."""

    def __init__(self, pkg, raw, device):
        import ctypes
        self.pkg, self.raw, self.device, self.C = pkg, raw, device, ctypes

    def csr_tensors(self):
        ip, ix, va = self.raw.csr_device()
        n, nnz = self.raw.total_dofs(), self.raw.nnz()
        return (device_tensor(ip, n + 1, "<i8", self.device), device_tensor(ix, nnz, "<i4", self.device),
                device_tensor(va, nnz, "<f8", self.device))

    def keep_rows(self, mask):
        rc = self.pkg.lib().femb200_csr_keep_rows_device(self.raw.h, self.C.c_void_p(mask.data_ptr()))
        if rc:
            raise RuntimeError(f"femb200_csr_keep_rows_device -> {rc}")

    def accumulate_rows(self, rows, indptr0, indices, values):
        c = self.C.c_void_p
        rc = self.pkg.lib().femb200_csr_accumulate_rows_device(self.raw.h, rows.numel(), c(rows.data_ptr()), c(indptr0.data_ptr()),
                                                                c(indices.data_ptr()), c(values.data_ptr()))
        if rc:
            raise RuntimeError(f"femb200_csr_accumulate_rows_device -> {rc}")


    def add_entries(self, rows, indptr0, indices, values):
        c = self.C.c_void_p
        rc = self.pkg.lib().femb200_csr_add_entries_device(self.raw.h, rows.numel(), c(rows.data_ptr()), c(indptr0.data_ptr()),
                                                            c(indices.data_ptr()), c(values.data_ptr()))
        if rc not in (0, 24):
            raise RuntimeError(f"femb200_csr_add_entries_device -> {rc}")
        return rc == 0

    def _h(self):
        return self.raw.h

    def pack_rows_into(self, rows, ptr, idx, val, mismatch):
        c = self.C.c_void_p
        rc = self.pkg.lib().femb200_csr_pack_rows_async(self._h(), rows.numel(), c(rows.data_ptr()), c(ptr.data_ptr()), c(idx.data_ptr()),
                                                        c(val.data_ptr()), c(mismatch.data_ptr()))
        if rc:
            raise RuntimeError(f"femb200_csr_pack_rows_async -> {rc}")

    def add_entries_from(self, rows, ptr, idx, val, mismatch):
        c = self.C.c_void_p
        rc = self.pkg.lib().femb200_csr_add_entries_async(self._h(), rows.numel(), c(rows.data_ptr()), c(ptr.data_ptr()), c(idx.data_ptr()),
                                                          c(val.data_ptr()), c(mismatch.data_ptr()))
        if rc:
            raise RuntimeError(f"femb200_csr_add_entries_async -> {rc}")


class NumpyBackend(LocalBackend):
    """CPU stand-in used by the gloo tests: holds a canonical CSR in numpy, merges like the device kernels
    (pattern union, explicit zeros kept, existing value + piece)."""

    def __init__(self, indptr, indices, values):
        self.indptr, self.indices, self.values = indptr.copy(), indices.copy(), values.copy()

    def csr_tensors(self):
        import torch
        return torch.from_numpy(self.indptr), torch.from_numpy(self.indices), torch.from_numpy(self.values)

    def keep_rows(self, mask):
        keep = mask.numpy().astype(bool)
        lens = np.diff(self.indptr) * keep
        rowid = np.repeat(np.arange(lens.size), np.diff(self.indptr))
        sel = keep[rowid]
        self.indices, self.values = self.indices[sel], self.values[sel]
        self.indptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)

    def add_entries(self, rows, indptr0, indices, values):
        rows, ptr, idx, val = rows.numpy(), indptr0.numpy(), indices.numpy(), values.numpy()
        slots = np.empty(idx.size, dtype=np.int64)
        for w, r in enumerate(rows):
            lo, hi = self.indptr[r], self.indptr[r + 1]
            seg = idx[ptr[w]:ptr[w + 1]]
            pos = lo + np.searchsorted(self.indices[lo:hi], seg)
            if np.any(pos >= hi) or np.any(self.indices[np.minimum(pos, hi - 1)] != seg):
                return False
            slots[ptr[w]:ptr[w + 1]] = pos
        self.values[slots] += val
        return True

    def pack_rows_into(self, rows, ptr, idx, val, mismatch):
        rows, ptr, oi, ov = rows.numpy(), ptr.numpy(), idx.numpy(), val.numpy()
        for w, r in enumerate(rows):
            lo, hi = self.indptr[r], self.indptr[r + 1]
            if hi - lo != ptr[w + 1] - ptr[w]:
                mismatch += 1
                continue
            oi[ptr[w]:ptr[w + 1]] = self.indices[lo:hi]
            ov[ptr[w]:ptr[w + 1]] = self.values[lo:hi]

    def add_entries_from(self, rows, ptr, idx, val, mismatch):
        if not self.add_entries(rows, ptr, idx, val):
            mismatch += 1

    def restrict_rows(self, node_mask, mdpn):
        """Keeps only the rows of the masked nodes (what the device assembler does with a node mask)."""
        import torch
        self.keep_rows(torch.from_numpy(np.repeat(node_mask.astype(np.uint8), mdpn)))

    def accumulate_rows(self, rows, indptr0, indices, values):
        rows, ptr, idx, val = rows.numpy(), indptr0.numpy(), indices.numpy(), values.numpy()
        n = self.indptr.size - 1
        a_row = np.repeat(np.arange(n), np.diff(self.indptr))
        b_row = np.repeat(rows, np.diff(ptr))
        r = np.concatenate([a_row, b_row])
        c = np.concatenate([self.indices, idx]).astype(np.int64)
        v = np.concatenate([self.values, val])
        order = np.lexsort((np.arange(r.size), c, r))            # stable: existing entry before the piece
        r, c, v = r[order], c[order], v[order]
        first = np.ones(r.size, dtype=bool)
        first[1:] = (r[1:] != r[:-1]) | (c[1:] != c[:-1])
        starts = np.nonzero(first)[0]
        self.values = np.add.reduceat(v, starts) if r.size else v
        self.indices = c[starts].astype(np.int32)
        self.indptr = np.concatenate([[0], np.cumsum(np.bincount(r[starts], minlength=n))]).astype(np.int64)


def gather_distributed_csr_local(backend, local, own_mask_local, mdpn, n_global_nodes, dist):
    """Test helper for ranks with a LocalNumbering: every rank contributes the rows of its OWN local nodes (global column
    ids already) -> the full canonical CSR in global numbering (host numpy) on every rank."""
    indptr, indices, values = backend.csr_tensors()
    ip, ix, va = indptr.cpu().numpy(), indices.cpu().numpy(), values.cpu().numpy()
    own_rows = np.nonzero(np.repeat(own_mask_local.astype(bool), mdpn))[0]
    grow = (local.l2g[own_rows // mdpn].astype(np.int64) * mdpn + own_rows % mdpn)
    lens = ip[own_rows + 1] - ip[own_rows]
    sel = np.repeat(ip[own_rows], lens) + (np.arange(int(lens.sum())) - np.repeat(np.cumsum(lens) - lens, lens))
    objs = [None] * dist.get_world_size()
    dist.all_gather_object(objs, (grow, lens, ix[sel], va[sel]))
    n = n_global_nodes * mdpn
    total = np.zeros(n, dtype=np.int64)
    for g, ln, _, _ in objs:
        assert not np.any(total[g] > 0), "two ranks own the same row"
        total[g] = ln
    ptr = np.concatenate([[0], np.cumsum(total)]).astype(np.int64)
    idx = np.zeros(ptr[-1], dtype=np.int32)
    val = np.zeros(ptr[-1])
    for g, ln, ixp, vap in objs:
        dst = np.repeat(ptr[g], ln) + (np.arange(int(ln.sum())) - np.repeat(np.cumsum(ln) - ln, ln))
        idx[dst], val[dst] = ixp, vap
    return ptr, idx, val


def gather_distributed_csr(backend, plan, dist, device):
    """Test helper: collects every rank's row block on every rank -> the full canonical CSR (host numpy)."""
    import torch
    indptr, indices, values = backend.csr_tensors()
    lens = (indptr[1:] - indptr[:-1]).cpu()
    world = plan.world
    objs = [None] * world
    dist.all_gather_object(objs, (lens.numpy(), indices.cpu().numpy(), values.cpu().numpy()))
    n = lens.numel()
    total_lens = np.zeros(n, dtype=np.int64)
    for ln, _, _ in objs:
        assert not np.any((total_lens > 0) & (ln > 0)), "two ranks own the same row"
        total_lens += ln
    ptr = np.concatenate([[0], np.cumsum(total_lens)]).astype(np.int64)
    idx = np.zeros(ptr[-1], dtype=np.int32)
    val = np.zeros(ptr[-1])
    for ln, ix, va in objs:
        rows = np.nonzero(ln)[0]
        src = np.concatenate([[0], np.cumsum(ln[rows])])
        for k, r in enumerate(rows):
            idx[ptr[r]:ptr[r + 1]] = ix[src[k]:src[k + 1]]
            val[ptr[r]:ptr[r + 1]] = va[src[k]:src[k + 1]]
    return ptr, idx, val
