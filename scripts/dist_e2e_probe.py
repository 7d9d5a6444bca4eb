"""Host-clock breakdown of one multi-GPU end-to-end step (bench.py run_multi_gpu: e2e_step) for ONE rank of `world`, emulated on
a single device: femb200_dist_create (whole mesh in from pinned host memory + partition) / add / csr_copy / destroy."""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import __graft_entry__ as entry  # noqa: E402
import torch  # noqa: E402

world = int(sys.argv[1]) if len(sys.argv) > 1 else 2
rank = int(sys.argv[2]) if len(sys.argv) > 2 else 0
pkg = entry.load_package()
L = pkg.lib()
N = int(round(48 * world ** (1.0 / 3.0)))
holder = pkg.RawAssembler(np.zeros((1, 3)), 3, device=0)
ctx = holder.ctx
dn, n_nodes, dc, n_elem = pkg.device_mesh_bcc_tetra4(ctx, N, N, N, 1.0 / N)
desc = holder.make_desc(pkg.elements.Tetra4(), pkg.solids.Isotropic(200e9, 0.3).Solid3D())
h_nodes = torch.empty(n_nodes * 3, dtype=torch.float64).pin_memory()
h_conn = torch.empty(n_elem * 4, dtype=torch.int32).pin_memory()
L.femb200_copy_to_host(ctx, h_nodes.data_ptr(), dn, h_nodes.numel() * 8)
L.femb200_copy_to_host(ctx, h_conn.data_ptr(), dc, h_conn.numel() * 4)
d = pkg.DistRank(ctx, rank, world, 3, 4, nodes_ptr=dn, conn_ptr=dc, n_nodes=n_nodes, n_elem=n_elem, order_mode=1)
info = d.info()
assert d.add(desc) == 0
nnz = d.info()["nnz"]
d.close()
h_indptr = torch.empty(info["local_nodes"] * 3 + 1, dtype=torch.int64).pin_memory()
h_indices = torch.empty(nnz, dtype=torch.int32).pin_memory()
h_values = torch.empty(nnz, dtype=torch.float64).pin_memory()
acc = np.zeros(4)
reps = 8
for it in range(reps + 2):
    t = [time.perf_counter()]
    h = C.c_void_p()
    assert L.femb200_dist_create(ctx, rank, world, h_nodes.data_ptr(), 0, n_nodes, 3, h_conn.data_ptr(), 0, 0, n_elem, 4, 1, C.byref(h)) == 0
    torch.cuda.synchronize(); t.append(time.perf_counter())
    _o = (C.c_int64 * 8)(); _pm = C.c_float(); L.femb200_dist_info(h, _o, C.byref(_pm)); part_ms = _pm.value
    assert L.femb200_dist_add_isoparametric(h, C.byref(desc)) == 0
    torch.cuda.synchronize(); t.append(time.perf_counter())
    assert L.femb200_csr_copy(L.femb200_dist_assembler(h), h_indptr.data_ptr(), h_indices.data_ptr(), h_values.data_ptr()) == 0
    torch.cuda.synchronize(); t.append(time.perf_counter())
    L.femb200_dist_destroy(h)
    torch.cuda.synchronize(); t.append(time.perf_counter())
    if it >= 2:
        acc += np.diff(t)
    if os.environ.get('PROBE_VERBOSE'):
        print('   iter', it, ' '.join(f'{1e3*x:.2f}' for x in np.diff(t)), f'partition(device events) {part_ms:.2f}', flush=True)
acc *= 1e3 / reps
print(f"world={world} rank={rank} N={N} elements={n_elem} local={info['local_elems']} nnz_rank={nnz}: create {acc[0]:.2f} ms (H2D {(h_nodes.numel()*8+h_conn.numel()*4)/1e6:.0f} MB, partition on device {info['partition_ms']:.2f} ms)  add {acc[1]:.2f}  csr_copy {acc[2]:.2f} ({(h_indptr.numel()*8+h_indices.numel()*4+h_values.numel()*8)/1e6:.0f} MB)  destroy {acc[3]:.2f}  total {acc.sum():.2f} ms")
