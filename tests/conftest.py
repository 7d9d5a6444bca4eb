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


GO_RAND_SEED1 = (0.6046602879796196, 0.9405090880450124, 0.6645600532184904, 0.4377141871869802,
                 0.4246374970712657, 0.6868230728671094, 0.0656370192174762, 0.1565192547327912)


def go_rand_float64(n):
    """The first values of Go's math/rand.New(rand.NewSource(1)).Float64(): the stream patch_test.go:28-40 jitters its
    interior nodes with.  The eight values the patch tests consume are the ones SURVEY.md section 4 records (Go's additive
    lagged Fibonacci generator cannot be re-derived without its 607-entry seed table, which is not in the reference tree);
    asking for more than eight is an error rather than a silent stand-in."""
    if n > len(GO_RAND_SEED1):
        raise ValueError("only the first 8 values of Go's seed-1 stream are recorded (SURVEY.md section 4)")
    return list(GO_RAND_SEED1[:n])


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
