import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def O():
    """CPU oracle (C restatement of the reference path)."""
    o = entry.load_oracle()
    o.build()
    return o


@pytest.fixture(scope="session")
def pkg():
    """The product package; the shared library must already be built (no fallback)."""
    p = entry.load_package()
    if not os.path.exists(p.LIB_PATH):
        entry.build()
    return p


@pytest.fixture(scope="session")
def meshes(pkg):
    from importlib import import_module
    return import_module("gofem_b200.meshes")


def go_rand_float64(n):
    """First n values of Go's math/rand.New(rand.NewSource(1)).Float64() are not reproducible without Go's
    ALFG table; the patch tests only need *some* interior jitter, so a fixed LCG stands in."""
    out, s = [], 0x2545F4914F6CDD1D
    for _ in range(n):
        s = (s * 6364136223846793005 + 1442695040888963407) & 0xFFFFFFFFFFFFFFFF
        out.append((s >> 11) / float(1 << 53))
    return out


# element kind table shared by CPU and GPU tests:
# name -> (oracle kind id, constituter mode id, model dofs flag, default mesh size, golden dims)
KINDS = {
    "tetra4": dict(kind=0, mode=0, flags=7),
    "tetra10": dict(kind=1, mode=0, flags=7),
    "hexa8": dict(kind=2, mode=0, flags=7),
    "hexa20": dict(kind=3, mode=0, flags=7),
    "quad4": dict(kind=4, mode=1, flags=3),
    "quad8": dict(kind=5, mode=1, flags=3),
    "triangle3": dict(kind=6, mode=1, flags=3),
}


def compare_csr(got, ref, absum, tol=1e-12):
    """Pattern bit-exact; values |a-b| <= tol*max(|b|, s_ij), s_ij = sum_e |Ke_e[i][j]|.
    Returns (max normalised error, max plain relative error on entries with |b| > 1e-6*s)."""
    gp, gi, gv = got
    rp, ri, rv = ref
    assert gp.dtype == np.int64 and gi.dtype == np.int32 and gv.dtype == np.float64
    assert np.array_equal(gp, rp), "indptr differs"
    assert np.array_equal(gi, ri), "indices differ"
    scale = np.maximum(np.abs(rv), absum)
    err = np.abs(gv - rv)
    norm = err / np.where(scale > 0, scale, 1.0)
    assert np.all(err <= tol * scale), f"values differ: max normalised error {norm.max():.3e}"
    big = np.abs(rv) > 1e-6 * absum
    rel = (err[big] / np.abs(rv[big])).max() if big.any() else 0.0
    return float(norm.max() if norm.size else 0.0), float(rel)
