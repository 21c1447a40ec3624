"""oracle/rowcheck.py — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Oracle values of selected rows of y = A*x (or transpose(A)*x) for the SYNTHETIC workloads, computed from the generators
alone: the row block is regenerated (hpcla_synth), the slice of x it touches is regenerated from the seed, and the rows
are summed by the oracle's row loop (orc_spmv_*: src/sparse.jl:2055-2066, left to right, no FMA).  Nothing global is
materialised, so the check scales to the full BASELINE sizes and to every rank of a multi-GPU run — including the rows
whose columns are ghosts, whose x values are all distinct (a mis-routed halo cannot reproduce them).

Used by tests/ and by bench.py's correctness guard (the oracle as the checker, never as the thing measured).
"""
from __future__ import annotations

import numpy as np

import hpcla_synth as S

from . import oracle as orc

TOL = {np.dtype(np.float32): 1e-5, np.dtype(np.float64): 1e-12, np.dtype(np.complex128): 1e-12}


def n_rows(kind: int, grid) -> int:
    return int(grid[0]) if kind == 3 else S.stencil_rows(kind, grid)


def _reach(kind: int, grid) -> int:
    """How far (in rows) a stencil row's columns lie from the row itself."""
    nx, ny, _ = S.grid3(grid)
    if kind == 0:
        return nx
    if kind == 1:
        return nx * ny
    return nx * ny + nx + 1


def _rows(kind: int, grid, g0: int, g1: int, T, Ti):
    if kind == 3:
        return S.powerlaw_local(int(grid[0]), S.POWERLAW_SEED, S.POWERLAW_MAX_LEN, g0, g1, T, Ti)
    return S.stencil_local(kind, grid, g0, g1, T, Ti)


class XSource:
    """x[g] = 2u(g) - 1 (hpcla_synth_vector), regenerated on demand; the whole vector is cached once it has been asked for."""

    def __init__(self, T, seed: int, n: int):
        self.T, self.seed, self.n, self._full = np.dtype(T), seed, n, None

    def hull(self, lo: int, hi: int) -> np.ndarray:
        if self._full is not None:
            return self._full[lo:hi]
        if hi - lo > self.n // 2:
            self._full = S.vector_local(self.T, self.seed, 0, self.n)
            return self._full[lo:hi]
        return S.vector_local(self.T, self.seed, lo, hi)


def expected_rows(kind: int, grid, g0: int, g1: int, T, Ti, x: XSource, transpose: bool = False) -> np.ndarray:
    """Oracle y[g0:g1] (0-based global rows) of A*x, or of transpose(A)*x (never conjugated, src/sparse.jl:2375-2379)."""
    n = n_rows(kind, grid)
    if not transpose:
        rp, c, v = _rows(kind, grid, g0, g1, T, Ti)
        if len(c) == 0:
            return np.zeros(g1 - g0, dtype=np.dtype(T))
        lo, hi = int(c.min()) - 1, int(c.max())
        return orc.spmv_csr(rp, (c.astype(np.int64) - lo).astype(rp.dtype), v, x.hull(lo, hi))
    if kind == 3:
        raise ValueError("transpose rows of the power-law matrix would need every row of A")
    import scipy.sparse as sp

    # rows of A^T = columns of A: every row j of A with an entry in columns [g0, g1) lies within `reach` of them
    r = _reach(kind, grid)
    j0, j1 = max(0, g0 - r), min(n, g1 + r)
    rp, c, v = _rows(kind, grid, j0, j1, T, Ti)
    blk = sp.csr_matrix((v, c.astype(np.int64) - 1, rp.astype(np.int64) - 1), shape=(j1 - j0, n))
    bt = sp.csr_matrix(blk[:, g0:g1].T)  # (g1-g0) x (j1-j0): row g of A^T restricted to the block, ascending j
    bt.sort_indices()
    return orc.spmv_csr((bt.indptr + 1).astype(np.int64), (bt.indices + 1).astype(np.int64), bt.data.astype(np.dtype(T)), x.hull(j0, j1))


def sample_ranges(kind: int, grid, b: int, e: int, seed: int, n_random: int = 8, run: int = 512):
    """Row ranges (0-based global, inside this rank's block [b, e)) to check: the first and the last boundary plane of
    the block — the rows that read ghosts — plus a few random interior runs."""
    if e <= b:
        return []
    nx, ny, _ = S.grid3(grid)
    plane = run if kind == 3 else (nx if kind == 0 else nx * ny)
    plane = min(plane, e - b)
    out = [(b, b + plane), (e - plane, e)]
    rng = np.random.default_rng(seed)
    if e - b > run:
        for s in rng.integers(b, e - run, size=n_random):
            out.append((int(s), int(s) + run))
    return out


def check_rows(kind: int, grid, b: int, e: int, y_local: np.ndarray, T, Ti, x_seed: int = S.X_SEED, transpose: bool = False, seed: int = 0,
               n_random: int = 8, run: int = 512):
    """Compare y_local (this rank's block, rows [b, e)) with the oracle on sample_ranges.  Returns (rows checked, max
    normwise relative error over the ranges, number of ranges that are bit-identical)."""
    n = n_rows(kind, grid)
    x = XSource(T, x_seed, n)
    rows, worst, exact = 0, 0.0, 0
    ranges = sample_ranges(kind, grid, b, e, seed, n_random, run)
    for g0, g1 in ranges:
        ref = expected_rows(kind, grid, g0, g1, T, Ti, x, transpose)
        got = np.asarray(y_local[g0 - b : g1 - b])
        d = np.linalg.norm(got.astype(np.complex128) - ref.astype(np.complex128))
        worst = max(worst, float(d / max(np.linalg.norm(ref.astype(np.complex128)), 1e-300)))
        exact += int(np.array_equal(got, ref))
        rows += g1 - g0
    return rows, worst, exact, len(ranges)
