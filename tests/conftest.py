import json
import os
import sys

import numpy as np
import pytest

# rank-threads of one process share one GPU in the single-GPU tests: give every stream its own hardware queue, so that a
# halo stream waiting for a peer's flag can never sit in front of that peer's own work (set before CUDA initialises)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
# ... and load every kernel up front: a lazy load (first launch) synchronises the context, which between rank-threads of
# one process can wait for a halo stream that waits for this very thread (direct halo; not an issue between processes)
os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _dec(v, dtype):
    if isinstance(v, dict):
        return np.array(v["re"], dtype=np.float64) + 1j * np.array(v["im"], dtype=np.float64)
    return np.array(v, dtype=dtype)


def load_fixtures():
    with open(os.path.join(ROOT, "tests", "golden", "reference_fixtures.json")) as f:
        raw = json.load(f)
    out = []
    for fx in raw["fixtures"]:
        dt = np.complex128 if fx["dtype"] == "c128" else np.float64
        c = dict(fx)
        for k in ("V", "x", "y", "yT", "y_adj", "xT", "AT_dense", "dot_with", "dot_xy", "dot_xx"):
            if k in c:
                c[k] = _dec(c[k], dt).astype(dt)
        out.append(c)
    return out, raw["plans"]


def load_product_fixtures():
    """Literal inputs of the reference's product tests (sparse*sparse, sparse*dense) with dense-numpy expectations."""
    with open(os.path.join(ROOT, "tests", "golden", "reference_fixtures.json")) as f:
        raw = json.load(f)
    out = []
    for fx in raw.get("products", []):
        dt = np.complex128 if fx["dtype"] == "c128" else np.float64
        c = dict(fx)
        for side in ("A", "B"):
            if side in c:
                c[side] = dict(c[side], V=_dec(c[side]["V"], dt).astype(dt))
        for k in ("Bdense", "C", "CT"):
            if k in c:
                c[k] = _dec(c[k], dt).astype(dt)
        out.append(c)
    return out


FIXTURES, PLAN_TABLES = load_fixtures()
PRODUCT_FIXTURES = load_product_fixtures()


def fixture_matrix(fx):
    """scipy CSR of a reference fixture, canonicalised like Julia's sparse(I,J,V,m,n)."""
    import scipy.sparse as sp

    A = sp.coo_matrix((fx["V"], (np.array(fx["I"]) - 1, np.array(fx["J"]) - 1)), shape=(fx["m"], fx["n"]))
    A = sp.csr_matrix(A)
    A.sum_duplicates()
    A.sort_indices()
    return A


@pytest.fixture(scope="session")
def fixtures():
    return FIXTURES


@pytest.fixture(scope="session")
def plan_tables():
    return PLAN_TABLES
