"""Synthetic mesh generators for the benchmark / parity configurations.

All generators return (nodes float64 [n,3] in r3.Vec layout, conn int64 [nelem, nn]) with connectivity in
the node order of the reference element types (elements/<type>.go IsoparametricNodes) so that
every Jacobian determinant is positive.  Interior nodes are jittered by 0.15*h*(u-0.5) per axis with
u = splitmix64(0x9E3779B97F4A7C15 ^ (3*node_id + axis)) / 2^64 so index bugs cannot hide behind symmetry.

The BCC tetra mesh reproduces the (dofs, elems) pairs of the reference benchmark
(benchmark_test.go:28-51, README.md:82-86): N=2 -> (105, 48) ... N=48 -> (684723, 1299456).
"""
import numpy as np

_SEED = np.uint64(0x9E3779B97F4A7C15)


def splitmix64(x):
    x = np.asarray(x, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = x + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return z


def jitter(nodes, interior_mask, h, amp=0.15, axes=3):
    """Moves interior nodes by amp*h*(u-0.5) per axis (deterministic in the node id)."""
    n = nodes.shape[0]
    ids = np.arange(n, dtype=np.uint64)
    out = nodes.copy()
    for ax in range(axes):
        with np.errstate(over="ignore"):
            u = splitmix64(_SEED ^ (ids * np.uint64(3) + np.uint64(ax))).astype(np.float64) / 2.0 ** 64
        out[:, ax] += np.where(interior_mask, amp * h * (u - 0.5), 0.0)
    return out


# ----------------------------------------------------------------------------- Tetra4 (BCC cube)
def bcc_tetra4(N, box=1.0, jittered=True):
    """BCC tetra mesh of [0,box]^3 with N cells per side: 12*N^2*(N-1) tets, (N+1)^3+N^3 nodes."""
    return bcc_tetra4_box(N, N, N, box / N, jittered)


def bcc_tetra4_box(nx, ny, nz, h=1.0, jittered=True):
    """BCC tetra mesh of an nx x ny x nz cell box (cell size h): corner nodes (x-major) then cell centres;
    4 tets (centre-, centre+, face edge k) per interior cell face, elements in cell-major order."""
    Mx, My, Mz = nx + 1, ny + 1, nz + 1
    cx, cy, cz = np.meshgrid(np.arange(Mx) * h, np.arange(My) * h, np.arange(Mz) * h, indexing="ij")
    corners = np.stack([cx.ravel(), cy.ravel(), cz.ravel()], axis=1).astype(np.float64)
    mx, my, mz = np.meshgrid((np.arange(nx) + 0.5) * h, (np.arange(ny) + 0.5) * h, (np.arange(nz) + 0.5) * h, indexing="ij")
    centres = np.stack([mx.ravel(), my.ravel(), mz.ravel()], axis=1).astype(np.float64)
    nodes = np.concatenate([corners, centres], axis=0)
    ncorner = Mx * My * Mz

    def cid(i, j, k):
        return (i * My + j) * Mz + k

    def mid(i, j, k):
        return ncorner + (i * ny + j) * nz + k

    tets, keys = [], []
    dims = (nx, ny, nz)
    for axis in range(3):
        # interior faces orthogonal to `axis`, between cell a-1 and cell a along it, a = 1..n_axis-1
        o1, o2 = [d for d in range(3) if d != axis]
        a = np.arange(1, dims[axis])
        A, B, Cc = np.meshgrid(a, np.arange(dims[o1]), np.arange(dims[o2]), indexing="ij")
        A, B, Cc = A.ravel(), B.ravel(), Cc.ravel()
        if axis == 0:
            lo, hi = mid(A - 1, B, Cc), mid(A, B, Cc)
            f = [cid(A, B, Cc), cid(A, B + 1, Cc), cid(A, B + 1, Cc + 1), cid(A, B, Cc + 1)]
            cell = ((A - 1) * ny + B) * nz + Cc
        elif axis == 1:
            lo, hi = mid(B, A - 1, Cc), mid(B, A, Cc)
            f = [cid(B, A, Cc), cid(B, A, Cc + 1), cid(B + 1, A, Cc + 1), cid(B + 1, A, Cc)]
            cell = (B * ny + (A - 1)) * nz + Cc
        else:
            lo, hi = mid(B, Cc, A - 1), mid(B, Cc, A)
            f = [cid(B, Cc, A), cid(B + 1, Cc, A), cid(B + 1, Cc + 1, A), cid(B, Cc + 1, A)]
            cell = (B * ny + Cc) * nz + (A - 1)
        for k in range(4):
            tets.append(np.stack([lo, hi, f[k], f[(k + 1) % 4]], axis=1))
            keys.append(cell * 12 + axis * 4 + k)
    conn = np.concatenate(tets, axis=0).astype(np.int64)
    order = np.argsort(np.concatenate(keys), kind="stable")       # cell-major element order (locality)
    conn = conn[order]
    if jittered:
        ext = np.array([nx, ny, nz]) * h
        interior = np.all((nodes > 1e-12) & (nodes < ext - 1e-12), axis=1)
        nodes = jitter(nodes, interior, h)
    conn = orient_tets(nodes, conn)
    return nodes, conn


def orient_tets(nodes, conn):
    """Swaps the last two vertices wherever det[x1-x0; x2-x0; x3-x0] < 0 (Tetra4 Jacobian, tetra4.go:40-46)."""
    p = nodes[conn[:, :4]]
    d = np.linalg.det(p[:, 1:4, :] - p[:, 0:1, :])
    neg = d < 0
    conn = conn.copy()
    conn[neg, 2], conn[neg, 3] = conn[neg, 3].copy(), conn[neg, 2].copy()
    return conn


# ----------------------------------------------------------------------------- Tetra10 from Tetra4
def tetra10_from_tetra4(nodes, conn4):
    """Adds one mid-edge node per unique edge, ordering 4:(0-1) 5:(0-2) 6:(0-3) 7:(1-2) 8:(2-3) 9:(1-3)
    (tetra10.go:29-42)."""
    pairs = [(0, 1), (0, 2), (0, 3), (1, 2), (2, 3), (1, 3)]
    n0 = nodes.shape[0]
    e = np.stack([np.stack([conn4[:, a], conn4[:, b]], axis=1) for a, b in pairs], axis=1)  # [ne,6,2]
    es = np.sort(e, axis=2).reshape(-1, 2)
    key = es[:, 0].astype(np.int64) * np.int64(n0) + es[:, 1]
    uniq, inv = np.unique(key, return_inverse=True)
    ua, ub = uniq // n0, uniq % n0
    mids = 0.5 * (nodes[ua] + nodes[ub])
    nodes10 = np.concatenate([nodes, mids], axis=0)
    conn10 = np.concatenate([conn4, (n0 + inv.reshape(-1, 6)).astype(np.int64)], axis=1)
    return nodes10, conn10


def bcc_tetra10(N, box=1.0, jittered=True):
    n4, c4 = bcc_tetra4(N, box, jittered)
    return tetra10_from_tetra4(n4, c4)


# ----------------------------------------------------------------------------- lattice-based structured meshes
_H8 = [(-1, -1, -1), (1, -1, -1), (1, 1, -1), (-1, 1, -1), (-1, -1, 1), (1, -1, 1), (1, 1, 1), (-1, 1, 1)]   # hexa8.go:32-43
_H20 = [(-1, -1, -1), (-1, 1, -1), (1, 1, -1), (1, -1, -1), (-1, -1, 1), (-1, 1, 1), (1, 1, 1), (1, -1, 1),
        (-1, 0, -1), (0, 1, -1), (1, 0, -1), (0, -1, -1), (-1, 0, 1), (0, 1, 1), (1, 0, 1), (0, -1, 1),
        (-1, -1, 0), (-1, 1, 0), (1, 1, 0), (1, -1, 0)]                                                      # hexa20.go:33-56
_Q4 = [(-1, -1), (1, -1), (1, 1), (-1, 1)]                                                                  # quad4.go:35-42
_Q8 = _Q4 + [(0, -1), (1, 0), (0, 1), (-1, 0)]                                                              # quad8.go:36-47


def _lattice_mesh(shape, ref_nodes, h=1.0, origin=(0.0, 0.0, 0.0), jittered=True, quadratic=False):
    """Elements on an nx x ny (x nz) grid; local node with reference coordinate xi in {-1,0,1}^dim sits at
    lattice point 2*cell + 1 + xi of the (2n+1)^dim half-step lattice.  Unused lattice points are dropped."""
    dim = len(shape)
    n = list(shape) + [0] * (3 - dim)
    L = [2 * n[0] + 1, 2 * n[1] + 1, (2 * n[2] + 1) if dim == 3 else 1]
    idx = np.meshgrid(*[np.arange(s) for s in shape], indexing="ij")
    cells = [a.ravel() for a in idx] + [None] * (3 - dim)
    cols = []
    for xi in ref_nodes:
        p = [2 * cells[a] + 1 + xi[a] for a in range(dim)]
        if dim == 2:
            p.append(np.zeros_like(p[0]))
        cols.append((p[0] * L[1] + p[1]) * L[2] + p[2])
    lat = np.stack(cols, axis=1)
    used, inv = np.unique(lat.ravel(), return_inverse=True)
    conn = inv.reshape(lat.shape).astype(np.int64)
    pk = used % L[2]
    pj = (used // L[2]) % L[1]
    pi = used // (L[2] * L[1])
    nodes = np.stack([origin[0] + 0.5 * h * pi, origin[1] + 0.5 * h * pj, origin[2] + 0.5 * h * pk], axis=1).astype(np.float64)
    if jittered:
        if quadratic:
            # jitter corner (even-even-even) lattice nodes only, then put mid-edge nodes at edge midpoints
            corner = (pi % 2 == 0) & (pj % 2 == 0) & (pk % 2 == 0)
            interior = corner & (pi > 0) & (pi < L[0] - 1) & (pj > 0) & (pj < L[1] - 1)
            if dim == 3:
                interior &= (pk > 0) & (pk < L[2] - 1)
            jn = jitter(nodes, interior, h, axes=dim)
            # midpoint of the two neighbouring corner lattice points along the odd axis
            lut = -np.ones(L[0] * L[1] * L[2], dtype=np.int64)
            lut[used] = np.arange(used.size)
            out = jn.copy()
            for ax, pc in enumerate((pi, pj, pk)[:dim]):
                odd = (pc % 2 == 1)
                step = [L[1] * L[2], L[2], 1][ax]
                a = lut[used[odd] - step]
                b = lut[used[odd] + step]
                out[odd] = 0.5 * (jn[a] + jn[b])
            nodes = out
        else:
            interior = (pi > 0) & (pi < L[0] - 1) & (pj > 0) & (pj < L[1] - 1)
            if dim == 3:
                interior &= (pk > 0) & (pk < L[2] - 1)
            nodes = jitter(nodes, interior, h, axes=dim)
    return nodes, conn


def hexa8_grid(nx, ny, nz, h=1.0, jittered=True):
    return _lattice_mesh((nx, ny, nz), _H8, h, jittered=jittered)


def hexa20_grid(nx, ny, nz, h=1.0, jittered=True):
    return _lattice_mesh((nx, ny, nz), _H20, h, jittered=jittered, quadratic=True)


def quad4_grid(nx, ny, h=1.0, r0=1.0, jittered=True):
    """Quad4 grid on r in [r0, r0+nx*h], z in [0, ny*h] (config 2: axisymmetric, r0=1)."""
    return _lattice_mesh((nx, ny), _Q4, h, origin=(r0, 0.0, 0.0), jittered=jittered)


def quad8_grid(nx, ny, h=1.0, r0=1.0, jittered=True):
    return _lattice_mesh((nx, ny), _Q8, h, origin=(r0, 0.0, 0.0), jittered=jittered, quadratic=True)


def triangle3_grid(nx, ny, h=1.0, r0=1.0, jittered=True):
    nodes, q = quad4_grid(nx, ny, h, r0, jittered)
    t = np.concatenate([q[:, [0, 1, 2]], q[:, [0, 2, 3]]], axis=0)
    order = np.argsort(np.concatenate([np.arange(q.shape[0]) * 2, np.arange(q.shape[0]) * 2 + 1]), kind="stable")
    return nodes, t[order].astype(np.int64)


def mesh_for(kind_name, size, jittered=True):
    """Convenience dispatch used by tests and bench: kind_name in
    {tetra4, tetra10, hexa8, hexa20, quad4, quad8, triangle3}; size = cells per side (int or tuple)."""
    s = size if isinstance(size, (tuple, list)) else None
    if kind_name == "tetra4":
        return bcc_tetra4(size, jittered=jittered)
    if kind_name == "tetra10":
        return bcc_tetra10(size, jittered=jittered)
    if kind_name == "hexa8":
        return hexa8_grid(*(s or (size,) * 3), jittered=jittered)
    if kind_name == "hexa20":
        return hexa20_grid(*(s or (size,) * 3), jittered=jittered)
    if kind_name == "quad4":
        return quad4_grid(*(s or (size,) * 2), jittered=jittered)
    if kind_name == "quad8":
        return quad8_grid(*(s or (size,) * 2), jittered=jittered)
    if kind_name == "triangle3":
        return triangle3_grid(*(s or (size,) * 2), jittered=jittered)
    raise ValueError(kind_name)
