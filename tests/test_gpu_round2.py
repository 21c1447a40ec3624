"""GPU parity tests added in round 2: the boundary the way the Julia binding calls it (imported plans), the two new
multiply kernels (row walk on compact tiles, nnz-split kernel for irregular matrices), CUDA-graph replay, the full-size
BASELINE configs 3 and 4, and the irregular-tile / ghost-column case of ADVICE.md.  Same tolerances as
test_gpu_parity.py (1e-12 Float64 / ComplexF64, 1e-5 Float32, normwise); integer arrays bit-exact."""
import numpy as np
import pytest
import scipy.sparse as sp
import torch

import hpcla_b200 as la
from oracle import oracle as orc
from oracle import rowcheck
from test_gpu_parity import ALL_TYPES, TOL, backends, relerr, spmd

pytestmark = pytest.mark.gpu


def _itype(Ti):
    return "i32" if np.dtype(Ti) == np.int32 else "i64"


# ---------------------------------------------------------------------------------------------------------------
# hpcla_plan_import: the plan the reference built (here: the oracle's restatement of src/sparse.jl:1875-1984, arrays in
# Ti width) handed to the library, as julia/HPCLinearAlgebraB200Ext.jl does
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("Ti", [np.int32, np.int64])
@pytest.mark.parametrize("P", [1, 2, 3])
def test_plan_import_route(Ti, P):
    rng = np.random.default_rng(11)
    m, n = 900, 700
    G = sp.random(m, n, density=0.02, random_state=rng, format="csr").astype(np.float64)
    G.data = rng.uniform(-1, 1, G.nnz)
    xh = rng.uniform(-1, 1, n)
    xp = np.concatenate([[1], np.sort(rng.integers(1, n + 1, size=P - 1)), [n + 1]]).astype(np.int64)
    olocs = orc.distribute(G, P, itype=_itype(Ti))
    oplans = orc.vector_plans(olocs, xp)

    def body(rank, bs):
        b = bs[rank]
        torch.cuda.set_device(b.torch_device())
        A = la.HPCSparseMatrix.from_global(G, b)
        x = la.HPCVector.from_global(xh, b, partition=xp)
        o = oplans[rank]
        n0 = la.sparse.plan_build_count
        plan = la.sparse.import_vector_plan(A, x, o.send_rank_ids, o.send_indices, o.recv_rank_ids, o.recv_perm, o.local_src_indices,
                                            o.local_dst_indices, o.n_gathered)
        assert la.get_vector_plan(A, x) is plan and la.sparse.plan_build_count == n0  # imported, not rebuilt
        # every array comes back out exactly as it went in
        assert plan.send_rank_ids.tolist() == o.send_rank_ids.tolist() and plan.recv_rank_ids.tolist() == o.recv_rank_ids.tolist()
        assert np.array_equal(plan.local_src_indices, o.local_src_indices) and np.array_equal(plan.local_dst_indices, o.local_dst_indices)
        assert all(np.array_equal(a, c) for a, c in zip(plan.send_indices, o.send_indices))
        assert all(np.array_equal(a, c) for a, c in zip(plan.recv_perm, o.recv_perm))
        assert plan.local_src_indices.dtype == np.dtype(Ti) and plan.n_gathered == o.n_gathered
        g = la.execute_plan(plan, A, x)
        torch.cuda.synchronize()
        y = A * x
        return y.to_global(), g.cpu().numpy().copy()

    la.clear_plan_cache()
    res = spmd(backends(P, np.float64, Ti), body)
    y_ref = orc.matvec(olocs, xh, xp)
    W = orc.PlanWorld(olocs, xp)
    g_ref = W.execute(orc.split_vector(xh, xp))
    W.close()
    for r in range(P):
        assert relerr(res[r][0], y_ref) <= 1e-12
        assert np.array_equal(res[r][1], g_ref[r])


def test_plan_import_rejects_a_wrong_plan():
    b = la.backend_cuda_serial(np.float64, np.int32)
    A = la.synth.stencil_matrix(1, 8, b)
    x = la.synth.vector(512, b)
    good = orc.vector_plans(orc.distribute(sp.identity(512, format="csr"), 1, itype="i32"), orc.uniform_partition(512, 1))[0]
    with pytest.raises(la.HPCLAError):  # covers 511 of 512 gathered positions
        p = la.sparse.import_vector_plan(A, x, [], [], [], [], good.local_src_indices[:-1], good.local_dst_indices[:-1], 512)
        la.sparse._bound_op(A, p, x)
    la.clear_plan_cache()
    with pytest.raises(la.HPCLAError):  # a source index outside x.v
        bad = good.local_src_indices.copy()
        bad[3] = 513
        p = la.sparse.import_vector_plan(A, x, [], [], [], [], bad, good.local_dst_indices, 512)
        la.sparse._bound_op(A, p, x)
    la.clear_plan_cache()


def test_csr_create_validates_the_borrowed_arrays():
    b = la.backend_cuda_serial(np.float64, np.int32)
    A = la.synth.stencil_matrix(1, 8, b)
    x = la.synth.vector(512, b)
    A.colval_target[5] = 513  # a column beyond ncols_compressed
    with pytest.raises(la.HPCLAError, match="valid 1-based CSR"):
        A * x
    A.colval_target[5] = int(A.colval[5])
    A.rowptr_target[7] = int(A.rowptr[9])  # not monotone
    with pytest.raises(la.HPCLAError, match="valid 1-based CSR"):
        A * x
    A.rowptr_target[7] = int(A.rowptr[7])
    rp, c, v = la.synth.stencil_local(1, 8, 0, 512, np.float64, np.int32)
    ref = orc.spmv_csr(rp, c, v, x.local_values())
    assert np.array_equal((A * x).local_values(), ref)


# ---------------------------------------------------------------------------------------------------------------
# row walk on compact tiles
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kind,N,T,Ti", [(1, 48, np.float64, np.int32), (1, 41, np.float64, np.int64), (1, 37, np.float32, np.int32),
                                         (2, 24, np.complex128, np.int32), (2, 21, np.float64, np.int64), (0, 301, np.float64, np.int64)])
@pytest.mark.parametrize("P", [1, 2, 3])
@pytest.mark.parametrize("ring", [0, 1])
def test_compact_row_walk(kind, N, T, Ti, P, ring, monkeypatch):
    """Interior tiles of the stencils go through the compact row walk (16-bit positions into bulk-copied runs of x);
    results equal the plain row walk bit for bit (same order of additions), and the oracle within tolerance."""
    S = la.synth
    monkeypatch.setenv("HPCLA_COMPACT", "1")  # (by default only matrices larger than L2 take the compact walk)
    monkeypatch.setenv("HPCLA_RING", str(ring))
    grid = (N, N) if kind == 0 else N
    n = S.stencil_rows(kind, grid)
    rp, c, v = S.stencil_local(kind, grid, 0, n, T, Ti)
    G = sp.csr_matrix((v, c.astype(np.int64) - 1, rp.astype(np.int64) - 1), shape=(n, n))
    xh = S.vector_local(T, S.X_SEED, 0, n)

    def body(rank, bs):
        b = bs[rank]
        torch.cuda.set_device(b.torch_device())
        A = S.stencil_matrix(kind, grid, b)
        xv = S.vector(n, b)
        y = (A * xv).to_global()
        info = la.spmv_info(A, xv)
        # x.v at an address that is not 16-byte aligned: the bulk copies of x cannot be used, the plain walk takes over
        buf = torch.zeros(xv.local_size + 1, dtype=xv.v.dtype, device=xv.v.device)
        x_off = la.HPCVector(xv.structural_hash, xv.partition, buf[1:], b)
        x_off.v.copy_(xv.v)
        y_off = (A * x_off).to_global() if xv.v.element_size() < 16 else y
        return y, info, y_off

    la.clear_plan_cache()
    res = spmd(backends(P, T, Ti), body)
    monkeypatch.setenv("HPCLA_COMPACT", "0")
    la.clear_plan_cache()
    plain = spmd(backends(P, T, Ti), body)
    y_ref = orc.matvec(orc.distribute(G, P, itype=_itype(Ti)), xh)
    for (y, info, y_off), (yp, infop, _) in zip(res, plain):
        assert info["compact_tiles"] > 0 and infop["compact_tiles"] == 0, (info, infop)
        assert info["compact_tiles"] + info["plain_interior_rowwalk_tiles"] == infop["plain_interior_rowwalk_tiles"]
        assert np.array_equal(y, yp) and np.array_equal(y_off, yp)
        assert relerr(y, y_ref) <= TOL[np.dtype(T)]
        if kind in (0, 1):
            assert np.array_equal(y, y_ref)


def test_compact_row_walk_banded_with_odd_sizes_and_cg(monkeypatch):
    """A banded matrix whose x runs reach the very end of an odd-length x.v (tail elements no 16-byte copy may fetch),
    and CG with the fused p.q partials coming from compact, plain and boundary tiles."""
    monkeypatch.setenv("HPCLA_COMPACT", "1")
    rng = np.random.default_rng(3)
    n = 30011
    diags = [rng.uniform(-1, 1, n) for _ in range(7)]
    offs = [-700, -33, -1, 0, 1, 33, 700]
    G = sp.diags(diags, offs, shape=(n, n), format="csr")
    G = (G + G.T + sp.identity(n) * 20).tocsr()
    xh = rng.uniform(-1, 1, n)
    for P in (1, 2):
        def body(rank, bs):
            b = bs[rank]
            torch.cuda.set_device(b.torch_device())
            A = la.HPCSparseMatrix.from_global(G, b)
            x = la.HPCVector.from_global(xh, b)
            return (A * x).to_global(), la.spmv_info(A, x)

        la.clear_plan_cache()
        res = spmd(backends(P, np.float64, np.int32), body)
        y_ref = orc.matvec(orc.distribute(G, P, itype="i32"), xh)
        for y, info in res:
            assert info["compact_tiles"] > 0
            assert np.array_equal(y, y_ref)  # 7 entries per row, one lane per row: the reference's order
    b = la.backend_cuda_serial(np.float64, np.int32)
    A = la.HPCSparseMatrix.from_global(G, b)
    bh = G @ np.ones(n)
    sol, hist = la.cg(A, la.HPCVector.from_global(bh, b), 12)
    xo, ho = orc.cg(orc.distribute(G, 1, itype="i32"), bh, 12)
    assert relerr(sol.to_global(), xo) <= 1e-9 and np.allclose(hist, ho, rtol=1e-8)


# ---------------------------------------------------------------------------------------------------------------
# nnz-split kernel (irregular matrices)
# ---------------------------------------------------------------------------------------------------------------
def _irregular(rng, m, n, T, long_rows=(), empty_rows=(), base_len=9):
    rows, cols = [], []
    for r in range(m):
        if r in empty_rows:
            continue
        L = long_rows.get(r, int(rng.pareto(1.5) * base_len) + 1) if isinstance(long_rows, dict) else base_len
        L = max(1, min(L, n))
        cs = np.sort(rng.choice(n, size=L, replace=False))
        rows.append(np.full(L, r))
        cols.append(cs)
    rows, cols = np.concatenate(rows), np.concatenate(cols)
    vals = rng.uniform(-1, 1, len(rows))
    if np.dtype(T) == np.complex128:
        vals = vals + 1j * rng.uniform(-1, 1, len(rows))
    return sp.csr_matrix((vals.astype(T), (rows, cols)), shape=(m, n))


@pytest.mark.parametrize("T,Ti", ALL_TYPES, ids=[f"{np.dtype(t).name}-{np.dtype(i).name}" for t, i in ALL_TYPES])
@pytest.mark.parametrize("P", [1, 2, 3])
def test_flat_kernel_irregular_rows(T, Ti, P, monkeypatch):
    """Power-law row lengths, empty rows (first, last and in between), rows spanning many warp chunks, rows above the
    split threshold, a stored-entry count that is not a multiple of 4."""
    monkeypatch.setenv("HPCLA_SPMV_KIND", "flat")  # (every rank's block, whatever its tile statistics)
    rng = np.random.default_rng(5)
    m, n = 6000, 5000
    long_rows = {17: 3000, 18: 4999, 2500: 1700, 2501: 700, 5998: 2100}
    empty = {0, 1, 40, 41, 42, 3000, 5999}
    G = _irregular(rng, m, n, T, long_rows=long_rows, empty_rows=empty)
    if G.nnz % 4 == 0:
        G = sp.csr_matrix(G + sp.csr_matrix(([1.0], ([5], [4])), shape=(m, n)).astype(T))
    xh = rng.uniform(-1, 1, n).astype(T)

    def body(rank, bs):
        b = bs[rank]
        torch.cuda.set_device(b.torch_device())
        A = la.HPCSparseMatrix.from_global(G, b)
        x = la.HPCVector.from_global(xh, b)
        y = la.HPCVector.zeros(b, m, partition=A.row_partition)
        y.v.fill_(float("nan"))
        la.mul(y, A, x)
        y2 = A * x
        assert torch.equal(y.v, y2.v)  # deterministic (no atomics)
        return y.to_global(), la.spmv_info(A, x)

    la.clear_plan_cache()
    res = spmd(backends(P, T, Ti), body)
    y_ref = orc.matvec(orc.distribute(G, P, itype=_itype(Ti)), xh)
    for y, info in res:
        assert info["flat_chunks"] > 0, info
        assert relerr(y, y_ref) <= TOL[np.dtype(T)], relerr(y, y_ref)
        assert not np.isnan(y).any()


def test_flat_kernel_very_long_row_and_tiny_matrices(monkeypatch):
    rng = np.random.default_rng(9)
    monkeypatch.setenv("HPCLA_SPMV_KIND", "flat")
    b = la.backend_cuda_serial(np.float32, np.int32)
    # one row far above the split threshold (16 384) in the middle of short ones
    m, n = 3000, 60000
    G = _irregular(rng, m, n, np.float32, long_rows={1000: 40000, 1001: 17000, 1002: 16384, 1003: 16385})
    xh = rng.uniform(-1, 1, n).astype(np.float32)
    A = la.HPCSparseMatrix.from_global(G, b)
    x = la.HPCVector.from_global(xh, b)
    info = la.spmv_info(A, x)
    assert info["flat_chunks"] > 0 and info["long_rows"] == 3
    assert relerr((A * x).to_global(), orc.matvec(orc.distribute(G, 1, itype="i32"), xh)) <= 1e-5
    # regular matrices too (a tridiagonal matrix through the nnz-split kernel), and matrices smaller than one warp chunk
    for G in (sp.csr_matrix(np.array([[2.0, 0, 1], [0, 0, 0], [1, 1, 1]], dtype=np.float32)),
              sp.random(50, 40, density=0.2, random_state=rng, format="csr").astype(np.float32),
              sp.csr_matrix(sp.diags([1.0, 2.0, 3.0], [-1, 0, 1], shape=(5000, 5000)).astype(np.float32))):
        la.clear_plan_cache()
        A = la.HPCSparseMatrix.from_global(G, b)
        xh = rng.uniform(-1, 1, G.shape[1]).astype(np.float32)
        x = la.HPCVector.from_global(xh, b)
        assert la.spmv_info(A, x)["flat_chunks"] > 0
        assert relerr((A * x).to_global(), orc.matvec(orc.distribute(G, 1, itype="i32"), xh)) <= 1e-5


def test_general_tiles_next_to_ghost_columns():
    """an interior tile of the general kernel whose 16-byte aligned hull reaches into a neighbouring
    tile that holds ghost columns.  Regular banded matrix (row-walk shape) with unbalanced stretches (general tiles)
    placed right behind rows whose last entries are ghosts, many junctions so that some fall on tile boundaries."""
    rng = np.random.default_rng(21)
    n = 24000
    P = 2
    half = n // 2
    rows, cols = [], []
    for r in range(n):
        own_lo = 0 if r < half else half
        base = [c for c in (r - 40, r - 1, r, r + 1, r + 40) if own_lo <= c < own_lo + half]
        k = r % 600
        if 100 <= k < 110:  # ghost columns as the LAST (rank 0) / FIRST (rank 1) entries of the row
            ghost = list(rng.choice(np.arange(half, n) if r < half else np.arange(0, half), size=3, replace=False))
            cs = sorted(set(base + ghost))
        elif 110 <= k < 140:  # an unbalanced stretch without ghosts: one long row among short ones -> general tiles
            extra = list(own_lo + rng.choice(half, size=(400 if k == 120 else 1), replace=False))
            cs = sorted(set(base + extra))
        else:
            cs = base
        rows += [r] * len(cs)
        cols += cs
    G = sp.csr_matrix((rng.uniform(-1, 1, len(rows)), (rows, cols)), shape=(n, n))
    xh = rng.uniform(-1, 1, n)

    def body(rank, bs):
        b = bs[rank]
        torch.cuda.set_device(b.torch_device())
        A = la.HPCSparseMatrix.from_global(G, b)
        x = la.HPCVector.from_global(xh, b)
        return (A * x).to_global(), la.spmv_info(A, x)

    la.clear_plan_cache()
    res = spmd(backends(P, np.float64, np.int32), body)
    y_ref = orc.matvec(orc.distribute(G, P, itype="i32"), xh)
    for y, info in res:
        assert info["general_tiles"] > 0 and info["rowwalk_tiles"] > 0 and info["boundary_tiles"] > 0, info
        assert relerr(y, y_ref) <= 1e-12


# ---------------------------------------------------------------------------------------------------------------
# CUDA graph replay, host buffers, timeline
# ---------------------------------------------------------------------------------------------------------------
def test_graph_replay_and_host_buffers(monkeypatch):
    monkeypatch.setenv("HPCLA_TIMELINE", "1")
    la.clear_plan_cache()
    b = la.backend_cuda_serial(np.float64, np.int32)
    n = 40**3
    A = la.synth.stencil_matrix(1, 40, b)
    x = la.synth.vector(n, b)
    y = A * x
    yg = la.HPCVector.zeros(b, n)
    for _ in range(3):
        yg.v.zero_()
        la.mul_graph(yg, A, x)
        torch.cuda.synchronize()
        assert torch.equal(yg.v, y.v)
    x.v.mul_(2.0)  # same buffers, new values: the replay reads them
    la.mul_graph(yg, A, x)
    torch.cuda.synchronize()
    assert torch.equal(yg.v, 2.0 * y.v)
    la.mul(yg, A, x)
    tl = la.spmv_timeline(A, x)
    assert tl["exchange_ms"] == -1 and tl["boundary_ms"] == -1 and 0 < tl["interior_ms"] <= tl["end_ms"]
    # NUMA-local pinned host buffers through the staged multiply
    hx, hy = la.host_buffer(b, n), la.host_buffer(b, n)
    hx.array[:] = la.synth.vector_local(np.float64, la.synth.X_SEED, 0, n)
    x2, y2 = la.HPCVector.zeros(b, n), la.HPCVector.zeros(b, n)
    la.mul_staged(y2, A, x2, hx.array, hy.array)
    torch.cuda.synchronize()
    assert np.array_equal(hy.array, y.local_values())
    hx.close(), hy.close()


def test_cg_graph_matches_plain_loop():
    import os
    import subprocess
    import sys

    from conftest import ROOT

    code = """
import numpy as np, torch
import hpcla_b200 as la
b = la.backend_cuda_serial(np.float64, np.int32)
n = 32**3
A = la.synth.stencil_matrix(1, 32, b)
rhs = A * la.HPCVector.from_global(np.ones(n), b)
xs, work = la.HPCVector.zeros(b, n), torch.empty(3 * n, dtype=torch.float64, device='cuda')
h = [la.cg(A, rhs, 15, x=xs, work=work)[1].copy() for _ in range(3)]
print('HIST', ' '.join(repr(float(v)) for v in h[2]), 'SAME', int(np.array_equal(h[0], h[1]) and np.array_equal(h[1], h[2])))
"""
    outs = []
    for graph in ("0", "1"):
        env = dict(os.environ, HPCLA_CG_GRAPH=graph, PYTHONPATH=ROOT)
        r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout + r.stderr
        line = [l for l in r.stdout.splitlines() if l.startswith("HIST")][0]
        assert line.endswith("SAME 1")
        outs.append(line)
    assert outs[0] == outs[1]  # the graph replays exactly the plain loop


# ---------------------------------------------------------------------------------------------------------------
# full-size BASELINE configs 3 and 4 (config 2 is in test_gpu_parity.py)
# ---------------------------------------------------------------------------------------------------------------
def test_full_size_stencil27_192_complex_and_transpose():
    """BASELINE config 3: 3-D 27-point stencil 192^3, ComplexF64 / Int32, A*x and transpose(A)*x at full size.
    Oracle rows regenerated from the generator (first / last planes + random runs), the bilinear identity
    <A^T x, z> = <x, A z> (no conjugation), and linearity."""
    S = la.synth
    N, T, Ti = 192, np.complex128, np.int32
    n = N**3
    b = la.backend_cuda_serial(T, Ti)
    A = S.stencil_matrix(2, N, b)
    assert A.nnz_local == (3 * N - 2) ** 3
    x = S.vector(n, b)
    y = A * x
    rows, worst, exact, nr = rowcheck.check_rows(2, N, 0, n, y.local_values(), T, Ti, n_random=12, run=1024)
    assert rows >= 2 * N * N and worst <= 1e-12, worst
    yT = la.transpose(A) * x
    rows, worstT, _, _ = rowcheck.check_rows(2, N, 0, n, yT.local_values(), T, Ti, transpose=True, n_random=6, run=512)
    assert worstT <= 1e-12, worstT
    assert relerr(yT.local_values(), y.local_values()) > 1e-3  # the generator's perturbation makes A^T differ from A
    z = S.vector(n, b, seed=77)
    lhs = torch.sum(yT.v * z.v).item()  # bilinear, no conjugation
    rhs = torch.sum(x.v * (A * z).v).item()
    assert abs(lhs - rhs) <= 1e-11 * abs(rhs)
    info = la.spmv_info(A, x)
    assert info["rowwalk_tiles"] > 0.99 * info["tiles"], info


def test_full_size_powerlaw_20m():
    """BASELINE config 4: power-law rows, 20 M x 20 M, Float32 / Int32, against the oracle's row loop on every row."""
    S = la.synth
    n, T, Ti = 20_000_000, np.float32, np.int32
    b = la.backend_cuda_serial(T, Ti)
    A = S.powerlaw_matrix(n, b)
    x = S.vector(n, b)
    y = (A * x).local_values()
    info = la.spmv_info(A, x)
    assert info["flat_chunks"] > 0 and info["long_rows"] > 0, info
    del A
    rp, c, v = S.powerlaw_local(n, S.POWERLAW_SEED, S.POWERLAW_MAX_LEN, 0, n, T, Ti)
    assert len(c) > 4.0e8
    y_ref = orc.spmv_csr(rp, c, v, S.vector_local(T, S.X_SEED, 0, n))
    assert relerr(y, y_ref) <= 1e-5, relerr(y, y_ref)
    # per-row: Float32 sums of up to 10^6 terms differ by summation order only
    err = np.abs(y - y_ref)
    scale = np.maximum(np.abs(y_ref), 1.0)
    assert float(np.max(err / scale)) <= 1e-2


# ---------------------------------------------------------------------------------------------------------------
# sparse x dense on compact tiles
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kind,N,T,Ti", [(1, 40, np.float64, np.int32), (1, 33, np.float32, np.int64), (2, 20, np.float64, np.int32), (2, 16, np.complex128, np.int64)])
@pytest.mark.parametrize("P", [1, 2])
@pytest.mark.parametrize("ring", [0, 3])
def test_sparse_times_dense_on_compact_tiles(kind, N, T, Ti, P, ring, monkeypatch):
    """A * B::HPCMatrix with the interior tiles on the compact kernel (x runs of all columns staged by bulk copies):
    every column equals A * B[:, k] bit for bit, 13 columns = one pass of 8, one of 4 and one single column.
    ring = 3: the same tiles behind rings of persistent CTAs (producer warp + consumer warps), multiply and product."""
    monkeypatch.setenv("HPCLA_COMPACT", "1")
    monkeypatch.setenv("HPCLA_RING", str(ring))
    S = la.synth
    n = S.stencil_rows(kind, N)
    ncols = 13
    rp, c, v = S.stencil_local(kind, N, 0, n, T, Ti)
    G = sp.csr_matrix((v, c.astype(np.int64) - 1, rp.astype(np.int64) - 1), shape=(n, n))
    Bg = np.stack([S.vector_local(T, S.X_SEED + j, 0, n) for j in range(ncols)], axis=1)

    def body(rank, bs):
        b = bs[rank]
        torch.cuda.set_device(b.torch_device())
        A = S.stencil_matrix(kind, N, b)
        B = la.HPCMatrix.from_global(Bg, b)
        C = A * B
        cols = [(A * B.column(j)).to_global() for j in (0, 5, 12)]
        return C.to_global(), cols, la.spmv_info(A, B.column(0))

    la.clear_plan_cache()
    res = spmd(backends(P, T, Ti), body)
    ref = orc.matmat(orc.distribute(G, P, itype=_itype(Ti)), Bg)
    for Cg, cols, info in res:
        assert info["compact_tiles"] > 0
        assert relerr(Cg, ref) <= TOL[np.dtype(T)]
        for j, col in zip((0, 5, 12), cols):
            assert np.array_equal(Cg[:, j], col), j


# ---------------------------------------------------------------------------------------------------------------
# direct halo (peer push + flags instead of ncclSend/ncclRecv); here between the rank-threads of one process
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("P", [2, 3])
def test_direct_halo_between_rank_threads(P):
    """The push / flag protocol of the direct halo between rank-threads on one GPU, in a FRESH process (see the worker's
    docstring: a stream in a flag wait must not share a hardware queue with its peer's streams, which a long-lived pytest
    process cannot guarantee) and under a hard time limit: a protocol bug shows as a hang, and a hang must fail the test, not
    stall the suite.  The multi-process form (CUDA IPC) is section 6 of tests/_nccl_worker.py."""
    import os
    import subprocess
    import sys

    from conftest import ROOT

    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.path.join(ROOT, "tests"))
    try:
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "_direct_halo_worker.py"), str(P)], env=env, capture_output=True, text=True, timeout=150)
    except subprocess.TimeoutExpired as e:
        pytest.fail("the direct halo between rank-threads did not finish within 150 s: " + ((e.stdout or b"").decode(errors="replace")[-2000:] if isinstance(e.stdout, bytes) else str(e.stdout)[-2000:]))
    assert r.returncode == 0 and "DIRECT_HALO_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
