"""Prints the device phase timers of config 1 (Tetra4 N=48) for quick kernel experiments."""
import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, torch
w = bench.make_workload(sys.argv[1] if len(sys.argv) > 1 else "tetra4", int(sys.argv[2]) if len(sys.argv) > 2 else 0)
pkg = w["pkg"]
ra = pkg.RawAssembler(w["nodes"], w["ndn"])
desc = ra.make_desc(w["elemT"], w["c"])
conn32 = w["conn"].astype(np.int32)
import ctypes as C
L = pkg.lib()
res = []
for it in range(6):
    h = C.c_void_p()
    L.femb200_assembler_create(ra.ctx, C.c_void_p(w["nodes"].ctypes.data), w["nodes"].shape[0], w["ndn"], 0, C.byref(h))
    rc = L.femb200_add_isoparametric(h, C.byref(desc), C.c_void_p(conn32.ctypes.data), 0, 0, conn32.shape[0])
    ph = (C.c_float * pkg.N_TIMERS)(); L.femb200_phase_ms(h, ph)
    L.femb200_assembler_destroy(h)
    if it >= 2: res.append([float(x) for x in ph])
m = np.mean(res, axis=0)
print(os.environ.get("FEMB200_G", "-"), w["desc"][:40], "rc", rc, " ".join(f"{n}={v:.3f}" for n, v in zip(pkg.TIMER_NAMES, m)))
