"""GPU parity tests: the CUDA path, called through the C ABI of libhpcla_b200.so (via the python mirror of the
reference API), against the CPU oracle on the same inputs.

Tolerances (BASELINE.json north_star): ||y - y_ref||_2 / ||y_ref||_2 <= 1e-12 (Float64, ComplexF64), <= 1e-5 (Float32);
plan arrays, ghost maps and the gathered vector bit-exact.  Multi-rank cases on a single GPU run as a single-process
world of rank-threads (device-to-device copies stand in for ncclSend/ncclRecv; kernels, pack, interior/boundary split and
ghost addressing are the same code)."""
import numpy as np
import pytest
import scipy.sparse as sp
import torch

import hpcla_b200 as la
from conftest import FIXTURES, fixture_matrix
from oracle import oracle as orc

pytestmark = pytest.mark.gpu

TOL = {np.dtype(np.float32): 1e-5, np.dtype(np.float64): 1e-12, np.dtype(np.complex128): 1e-12}
ALL_TYPES = [(np.float32, np.int32), (np.float32, np.int64), (np.float64, np.int32), (np.float64, np.int64),
             (np.complex128, np.int32), (np.complex128, np.int64)]


def relerr(y, ref):
    d = np.linalg.norm(np.asarray(y, dtype=np.complex128) - np.asarray(ref, dtype=np.complex128))
    return d / max(np.linalg.norm(np.asarray(ref, dtype=np.complex128)), 1e-300)


def backends(P, T, Ti):
    return [la.backend_cuda_serial(T, Ti)] if P == 1 else la.backends_threads(P, T, Ti, cuda=True)


def spmd(bs, fn, *args):
    if len(bs) == 1:
        return [fn(0, bs, *args)]
    return bs[0].comm.world.run(fn, bs, *args)


def _matvec_body(rank, bs, A_global, x_global, row_partition, x_partition, want_gathered, use_mul):
    b = bs[rank]
    torch.cuda.set_device(b.torch_device())
    A = la.HPCSparseMatrix.from_global(A_global, b, row_partition=row_partition)
    x = la.HPCVector.from_global(x_global, b, partition=x_partition)
    out = {}
    if want_gathered:
        g = la.execute_plan(la.get_vector_plan(A, x), A, x)
        torch.cuda.synchronize()
        out["gathered"] = g.cpu().numpy().copy()
    if use_mul:
        y = la.HPCVector.zeros(b, A_global.shape[0], partition=A.row_partition)
        y.v.fill_(float("nan"))
        assert la.mul(y, A, x) is y
    else:
        y = A * x
    assert y.partition.tolist() == A.row_partition.tolist() and y.backend is b and y.v.is_cuda
    out["y"] = y.to_global()
    out["info"] = la.spmv_info(A, x)
    At = la.materialize_transpose(A)
    xT = la.HPCVector.from_global(np.resize(x_global, A_global.shape[0]), b, partition=A.row_partition)
    yT = la.transpose(A) @ xT
    assert yT.partition.tolist() == A.col_partition.tolist()
    out["yT"] = yT.to_global()
    out["xT"] = xT.to_global()
    del At
    return out


def run_case(A_global, x_global, P, T, Ti, row_partition=None, x_partition=None, want_gathered=True, use_mul=False):
    A_global = sp.csr_matrix(A_global).astype(T)
    x_global = np.asarray(x_global).astype(T)
    la.clear_plan_cache()
    res = spmd(backends(P, T, Ti), _matvec_body, A_global, x_global, row_partition, x_partition, want_gathered, use_mul)
    itype = "i32" if Ti == np.int32 else "i64"
    olocs = orc.distribute(A_global, P, row_partition=row_partition, itype=itype)
    xp = orc.uniform_partition(A_global.shape[1], P) if x_partition is None else np.asarray(x_partition, dtype=np.int64)
    y_ref = orc.matvec(olocs, x_global, xp)
    for r in range(P):
        assert relerr(res[r]["y"], y_ref) <= TOL[np.dtype(T)], (r, relerr(res[r]["y"], y_ref))
    if want_gathered:
        W = orc.PlanWorld(olocs, xp)
        g_ref = W.execute(orc.split_vector(x_global, xp))
        W.close()
        for r in range(P):
            assert np.array_equal(res[r]["gathered"], g_ref[r]), f"gathered differs on rank {r}"
    yT_ref = orc.matvec(orc.transpose(olocs), res[0]["xT"])
    assert relerr(res[0]["yT"], yT_ref) <= TOL[np.dtype(T)]
    return res, y_ref


@pytest.mark.parametrize("fx", FIXTURES, ids=[f["name"] for f in FIXTURES])
def test_reference_fixtures_on_device(fx):
    """The reference's own test cases (test/test_vector_multiplication.jl etc.), A*x and mul!, 1/2/3 ranks."""
    T = np.complex128 if fx["dtype"] == "c128" else np.float64
    A = fixture_matrix(fx)
    for P, Ti, use_mul in [(1, np.int64, False), (2, np.int64, True), (2, np.int32, False), (3, np.int32, True)]:
        rp = fx.get("row_partition") if P == 2 else None
        res, _ = run_case(A, fx["x"], P, T, Ti, row_partition=rp, use_mul=use_mul)
        assert np.max(np.abs(res[0]["y"] - fx["y"])) < fx["tol"]  # the test's own tolerance (test_utils.jl:154-157)


@pytest.mark.parametrize("fx", [f for f in FIXTURES if "yT" in f], ids=[f["name"] for f in FIXTURES if "yT" in f])
def test_reference_fixtures_transpose_paths(fx):
    """transpose(A)*x, transpose(x)*A and x'*A (test_vector_multiplication.jl:141-159, test_new_operations.jl:73-76)."""
    T = np.complex128 if fx["dtype"] == "c128" else np.float64

    def body(rank, bs):
        b = bs[rank]
        torch.cuda.set_device(b.torch_device())
        A = la.HPCSparseMatrix.from_global(fixture_matrix(fx), b)
        x = la.HPCVector.from_global(fx.get("xT", fx["x"]), b, partition=A.row_partition)
        out = {"yT": (la.transpose(A) * x).to_global(), "vtA": la.vec_transpose_mul(x, A).to_global()}
        if "y_adj" in fx:
            out["adj"] = la.vec_adjoint_mul(x, A).to_global()
        if "dot_xy" in fx:
            y0 = la.HPCVector.from_global(fx["dot_with"], b)
            out["dot_xy"], out["dot_xx"], out["nrm"] = la.dot(x, y0), la.dot(x, x), la.norm(x)
        return out

    for P in (1, 2):
        la.clear_plan_cache()
        for r in spmd(backends(P, T, np.int64), body):
            assert np.max(np.abs(r["yT"] - fx["yT"])) < fx["tol"] and np.max(np.abs(r["vtA"] - fx["yT"])) < fx["tol"]
            if "adj" in r:
                assert np.max(np.abs(r["adj"] - fx["y_adj"])) < fx["tol"]
            if "dot_xy" in r:
                assert abs(r["dot_xy"] - fx["dot_xy"]) < fx["tol"] and abs(r["dot_xx"] - fx["dot_xx"]) < fx["tol"]
                assert abs(r["nrm"] - np.linalg.norm(fx["x"])) < fx["tol"]


def _ragged(rng, m, n, dens, T, long_rows=()):
    A = sp.random(m, n, density=dens, random_state=rng, format="lil")
    for r, L in long_rows:
        cols = rng.choice(n, size=min(L, n), replace=False)
        A[r, cols] = 1.0
    A = sp.csr_matrix(A)
    A.data = rng.uniform(-1, 1, A.nnz)
    if np.dtype(T).kind == "c":
        A = A.astype(np.complex128)
        A.data = A.data + 1j * rng.uniform(-1, 1, A.nnz)
    A = sp.csr_matrix(sp.diags(np.r_[0.0, np.ones(m - 2), 0.0]) @ A)  # first/last rows empty
    A.eliminate_zeros()
    A.sort_indices()
    return A.astype(T)


def _vec(rng, n, T):
    x = rng.uniform(-1, 1, n)
    if np.dtype(T).kind == "c":
        x = x + 1j * rng.uniform(-1, 1, n)
    return x.astype(T)


@pytest.mark.parametrize("T,Ti", ALL_TYPES, ids=[f"{np.dtype(t).name}-{np.dtype(i).name}" for t, i in ALL_TYPES])
@pytest.mark.parametrize("P", [1, 2, 4])
def test_random_ragged_all_types(T, Ti, P):
    rng = np.random.default_rng(42 + P)
    for (m, n, dens) in [(257, 300, 0.05), (3000, 3000, 0.004), (40, 5000, 0.2)]:
        run_case(_ragged(rng, m, n, dens, T), _vec(rng, n, T), P, T, Ti)


@pytest.mark.parametrize("P", [1, 3])
def test_every_row_length_regime(P):
    """Tiles reduced by 1..32 lanes per row, tiles that do not fit shared memory (warp-per-row), split long rows."""
    rng = np.random.default_rng(7)
    for T, Ti in [(np.float64, np.int32), (np.float32, np.int64), (np.complex128, np.int32)]:
        for mean_len in (3, 20, 40, 80, 150, 400):
            m, n = 600, 4000
            A = _ragged(rng, m, n, mean_len / n, T)
            run_case(A, _vec(rng, n, T), P, T, Ti, want_gathered=False)
        # rows of 3k (does not fit a tile -> warp per row) and 40k / 70k (split across CTAs), among short rows
        m, n = 500, 80000
        A = _ragged(rng, m, n, 5 / n, T, long_rows=[(10, 3000), (200, 40000), (201, 70000), (430, 17000)])
        res, _ = run_case(A, _vec(rng, n, T), P, T, Ti, want_gathered=False)
        if P == 1:
            assert res[0]["info"]["long_rows"] == 3


def test_partitions_with_empty_ranks_and_foreign_x_partition():
    rng = np.random.default_rng(3)
    A = _ragged(rng, 90, 70, 0.1, np.float64)
    x = _vec(rng, 70, np.float64)
    run_case(A, x, 3, np.float64, np.int32, row_partition=np.array([1, 46, 46, 91]), x_partition=np.array([1, 11, 61, 71]))
    run_case(A, x, 4, np.float64, np.int64, row_partition=np.array([1, 1, 31, 91, 91]), x_partition=np.array([1, 71, 71, 71, 71]))
    # a matrix whose own column block has holes: x.v cannot be read in place, the local-copy path runs
    B = sp.lil_matrix((64, 64))
    for i in range(64):
        B[i, (i * 2) % 64] = 1.0 + i
        B[i, (i * 2 + 6) % 64] = -2.0
    res, _ = run_case(sp.csr_matrix(B), _vec(rng, 64, np.float64), 2, np.float64, np.int32)
    assert res[0]["info"]["x_in_place"] == 0


@pytest.mark.parametrize("kind,N,T", [(0, 97, np.float64), (1, 40, np.float64), (1, 33, np.float32), (2, 20, np.complex128)])
@pytest.mark.parametrize("P", [1, 2, 4])
def test_synthetic_stencils(kind, N, T, P):
    """The BASELINE generators at test size; Float64 stencil rows (<= 12 entries, one lane per row, products rounded
    separately) reproduce the reference's summation order exactly => bit-identical y."""
    S = la.synth
    n = S.stencil_rows(kind, N)
    rp, c, v = S.stencil_local(kind, N, 0, n, T, np.int32)
    G = sp.csr_matrix((v, c - 1, rp - 1), shape=(n, n))
    x = S.vector_local(T, S.X_SEED, 0, n)

    def body(rank, bs):
        b = bs[rank]
        torch.cuda.set_device(b.torch_device())
        A = S.stencil_matrix(kind, N, b)
        xv = S.vector(n, b)
        y = A * xv
        yT = la.transpose(A) * xv
        return y.to_global(), yT.to_global(), la.spmv_info(A, xv)

    la.clear_plan_cache()
    res = spmd(backends(P, T, np.int32), body)
    olocs = orc.distribute(G, P, itype="i32")
    y_ref = orc.matvec(olocs, x)
    yT_ref = orc.matvec(orc.transpose(olocs), x)
    for y, yT, info in res:
        assert relerr(y, y_ref) <= TOL[np.dtype(T)] and relerr(yT, yT_ref) <= TOL[np.dtype(T)]
        if kind in (0, 1):
            assert info["lanes_per_row"] == 1, info
            assert np.array_equal(y, y_ref), "one lane per row must be bit-identical to the reference's summation order"
            assert info["rowwalk_tiles"] > 0
        assert info["x_in_place"] == 1 and info["sends_contiguous"] == 1
        if P > 1:
            assert info["boundary_tiles"] > 0 and info["interior_tiles"] > 0


def test_powerlaw_load_balance_case_small():
    S = la.synth
    n = 60000
    rp, c, v = S.powerlaw_local(n, 0xC4, 30000, 0, n, np.float32, np.int32)
    G = sp.csr_matrix((v, c - 1, rp - 1), shape=(n, n))
    x = S.vector_local(np.float32, S.X_SEED, 0, n)
    for P in (1, 2):
        def body(rank, bs):
            torch.cuda.set_device(bs[rank].torch_device())
            A = S.powerlaw_matrix(n, bs[rank], max_len=30000)
            return (A * S.vector(n, bs[rank])).to_global()

        la.clear_plan_cache()
        res = spmd(backends(P, np.float32, np.int32), body)
        y_ref = orc.matvec(orc.distribute(G, P, itype="i32"), x)
        assert relerr(res[0], y_ref) <= 1e-5


def test_memoised_plan_and_in_place_value_updates():
    b = la.backend_cuda_serial(np.float64, np.int32)
    A = la.synth.stencil_matrix(1, 16, b)
    x = la.synth.vector(16**3, b)
    la.clear_plan_cache()
    n0 = la.sparse.plan_build_count
    y1 = (A * x).to_global()
    y2 = (A * x).to_global()
    assert la.sparse.plan_build_count == n0 + 1 and np.array_equal(y1, y2)
    A.nzval.mul_(2.0)  # in-place value write keeps the structure hash and the plan (src/indexing.jl:932-982)
    A.values_changed()
    assert np.array_equal((A * x).to_global(), 2.0 * y1) and la.sparse.plan_build_count == n0 + 1
    info = la.spmv_info(A, x)
    assert 3 <= info["launches"] <= 6 and info["boundary_tiles"] == 0  # one or two kernels per multiply on a single rank (compact + plain row walk)


def test_vector_ops_and_cg():
    for T in (np.float64, np.float32):
        b = la.backend_cuda_serial(T, np.int32)
        N = 12
        n = N**3
        A = la.synth.stencil_matrix(1, N, b)
        rng = np.random.default_rng(0)
        xh, yh = rng.uniform(-1, 1, n).astype(T), rng.uniform(-1, 1, n).astype(T)
        x, y = la.HPCVector.from_global(xh, b), la.HPCVector.from_global(yh, b)
        tol = 1e-5 if T == np.float32 else 1e-12
        assert abs(la.dot(x, y) - np.dot(xh.astype(np.float64), yh.astype(np.float64))) <= tol * n
        assert abs(la.norm(x) - np.linalg.norm(xh.astype(np.float64))) <= tol * n
        la.axpby(0.5, x, -2.0, y)
        assert relerr(y.to_global(), 0.5 * xh - 2.0 * yh) <= tol
        # CG against the oracle's textbook CG on the same matrix
        rp, c, v = la.synth.stencil_local(1, N, 0, n, T, np.int32)
        G = sp.csr_matrix((v, c - 1, rp - 1), shape=(n, n))
        bvec = (G @ np.ones(n)).astype(T)
        sol, hist = la.cg(A, la.HPCVector.from_global(bvec, b), 25)
        xo, ho = orc.cg(orc.distribute(G, 1, itype="i32"), bvec.astype(np.float64), 25)
        assert relerr(sol.to_global(), xo) <= (1e-3 if T == np.float32 else 1e-9)
        assert np.allclose(hist, np.array(ho, dtype=np.float64), rtol=1e-2 if T == np.float32 else 1e-8)
    bc = la.backend_cuda_serial(np.complex128, np.int64)
    xh = rng.uniform(-1, 1, 1000) + 1j * rng.uniform(-1, 1, 1000)
    yh = rng.uniform(-1, 1, 1000) + 1j * rng.uniform(-1, 1, 1000)
    x, y = la.HPCVector.from_global(xh, bc), la.HPCVector.from_global(yh, bc)
    assert abs(la.dot(x, y) - np.vdot(xh, yh)) <= 1e-10  # Julia's dot conjugates its first argument


def test_distributed_dot_norm_threads():
    def body(rank, bs):
        torch.cuda.set_device(bs[rank].torch_device())
        x = la.synth.vector(10007, bs[rank])
        y = la.synth.vector(10007, bs[rank], seed=99)
        return la.dot(x, y), la.norm(x)

    xh = la.synth.vector_local(np.float64, la.synth.X_SEED, 0, 10007)
    yh = la.synth.vector_local(np.float64, 99, 0, 10007)
    for d, nr in spmd(backends(3, np.float64, np.int64), body):
        assert abs(d - np.dot(xh, yh)) <= 1e-10 and abs(nr - np.linalg.norm(xh)) <= 1e-10


def test_full_size_poisson_256_properties():
    """BASELINE config 2 (3-D 7-point Poisson 256^3, Float64/Int32) at full size: size-independent properties plus the
    oracle's row-serial result on the whole vector (bit-identical: 7 entries per row, one lane per row)."""
    S = la.synth
    N = 256
    n = N**3
    b = la.backend_cuda_serial(np.float64, np.int32)
    A = S.stencil_matrix(1, N, b)
    assert A.nnz_local == 117047296 and A.shape == (n, n)
    ones = la.HPCVector.from_global(np.ones(n), b)
    y1 = (A * ones).local_values()
    g = np.arange(n)
    ix, iy, iz = g % N, (g // N) % N, g // (N * N)
    expect = ((ix == 0).astype(float) + (ix == N - 1) + (iy == 0) + (iy == N - 1) + (iz == 0) + (iz == N - 1))
    assert np.array_equal(y1, expect)  # row sums: 6 - (#neighbours)
    x = S.vector(n, b)
    y = A * x
    # linearity: A(2x + 1) == 2Ax + A1
    z = x.copy()
    la.axpby(1.0, ones, 2.0, z)
    lin = (A * z).local_values() - (2.0 * y.local_values() + y1)
    assert np.linalg.norm(lin) <= 1e-12 * np.linalg.norm(y.local_values())
    rp, c, v = S.stencil_local(1, N, 0, n, np.float64, np.int32)
    L = orc.LocalMatrix(0, A.row_partition, A.col_partition, A.col_indices, rp, c, v, n, n)  # all columns present: colval == global col
    y_ref = orc.spmv_local(L, x.local_values())
    assert np.array_equal(y.local_values(), y_ref)
    # symmetric operator: transpose(A)*x == A*x
    assert relerr((la.transpose(A) * x).local_values(), y_ref) <= 1e-12


@pytest.mark.parametrize("case", ["poisson", "stencil27c", "ragged", "powerlaw"])
def test_staged_multiply_matches_device_multiply(case, monkeypatch):
    """hpcla_spmv_run_staged (host x in, host y out, pipelined over row blocks) == copy; mul!; copy — bit for bit."""
    monkeypatch.setenv("HPCLA_STAGE_CHUNK_KB", "64")  # many small blocks at test size
    S = la.synth
    if case == "poisson":
        T, Ti = np.float64, np.int32
        b = la.backend_cuda_serial(T, Ti)
        A = S.stencil_matrix(1, 48, b)
    elif case == "stencil27c":
        T, Ti = np.complex128, np.int64
        b = la.backend_cuda_serial(T, Ti)
        A = S.stencil_matrix(2, 24, b)
    elif case == "ragged":
        T, Ti = np.float64, np.int32
        b = la.backend_cuda_serial(T, Ti)
        rng = np.random.default_rng(7)
        A = la.HPCSparseMatrix.from_global(_ragged(rng, 40000, 40000, 0.0005, T), b)
    else:
        T, Ti = np.float32, np.int32
        b = la.backend_cuda_serial(T, Ti)
        A = S.powerlaw_matrix(60000, b, max_len=30000)
    n = A.shape[1]
    x = S.vector(n, b)
    y_ref = (A * x).local_values()
    tdt = x.v.dtype
    xh = torch.empty(n, dtype=tdt, pin_memory=True).copy_(x.v)
    yh = torch.zeros(A.shape[0], dtype=tdt, pin_memory=True)
    x2 = la.HPCVector.zeros(b, n)
    y2 = la.HPCVector.zeros(b, A.shape[0])
    for _ in range(3):  # repeated calls reuse the pipeline state
        yh.zero_()
        la.mul_staged(y2, A, x2, xh, yh)
        torch.cuda.synchronize()
        assert np.array_equal(yh.numpy(), y_ref)
        assert np.array_equal(y2.local_values(), y_ref) and torch.equal(x2.v.cpu(), xh)


# ---------------------------------------------------------------------------------------------------------------
# sparse x dense: A * B::HPCMatrix in one library call vs the reference's column-by-column loop
# ---------------------------------------------------------------------------------------------------------------
def _spmm_body(rank, bs, A_global, B_global, row_partition, b_row_partition):
    b = bs[rank]
    torch.cuda.set_device(b.torch_device())
    A = la.HPCSparseMatrix.from_global(A_global, b, row_partition=row_partition)
    B = la.HPCMatrix.from_global(B_global, b, row_partition=b_row_partition)
    C = A * B
    assert isinstance(C, la.HPCMatrix) and C.row_partition.tolist() == A.row_partition.tolist()
    assert C.col_partition.tolist() == la.uniform_partition(B_global.shape[1], len(bs)).tolist()
    assert C.A.shape == (A.nrows_local, B_global.shape[1]) and (C.A.shape[0] <= 1 or C.A.stride(0) == 1)
    cols = [(A * B.column(k)).to_global() for k in range(B_global.shape[1])]  # the reference's loop, on the device
    CT = la.transpose(A) * la.HPCMatrix.from_global(np.resize(B_global, (A_global.shape[0], B_global.shape[1])), b, row_partition=A.row_partition)
    return C.to_global(), np.stack(cols, axis=1), CT.to_global(), la.spmv_info(A, B.column(0))


@pytest.mark.parametrize("case", ["poisson", "stencil27c", "ragged", "powerlaw", "laplace2d_f32"])
@pytest.mark.parametrize("P", [1, 2, 4])
def test_sparse_times_dense(case, P):
    rng = np.random.default_rng(100 + P)
    S = la.synth
    rowp = bpart = None
    if case == "poisson":
        T, Ti, ncols = np.float64, np.int32, 6
        rp, c, v = S.stencil_local(1, 24, 0, 24**3, T, Ti)
        A = sp.csr_matrix((v, c - 1, rp - 1), shape=(24**3, 24**3))
    elif case == "stencil27c":
        T, Ti, ncols = np.complex128, np.int64, 5
        rp, c, v = S.stencil_local(2, 14, 0, 14**3, T, Ti)
        A = sp.csr_matrix((v, c - 1, rp - 1), shape=(14**3, 14**3))
    elif case == "laplace2d_f32":
        T, Ti, ncols = np.float32, np.int32, 9
        rp, c, v = S.stencil_local(0, (90, 70), 0, 6300, T, Ti)
        A = sp.csr_matrix((v, c - 1, rp - 1), shape=(6300, 6300))
    elif case == "ragged":
        T, Ti, ncols = np.float64, np.int32, 7
        A = _ragged(rng, 3000, 2500, 0.004, T, long_rows=((5, 1700),))
        cuts = np.sort(rng.integers(1, 2501, size=P - 1))
        bpart = np.concatenate([[1], cuts, [2501]]).astype(np.int64)  # B's rows partitioned unlike A's columns
    else:
        T, Ti, ncols = np.float32, np.int32, 4
        n = 40000
        rp, c, v = S.powerlaw_local(n, 0xC4, 30000, 0, n, T, Ti)
        A = sp.csr_matrix((v, c - 1, rp - 1), shape=(n, n))
    A = sp.csr_matrix(A).astype(T)
    Bg = rng.uniform(-1, 1, (A.shape[1], ncols)).astype(T)
    if np.dtype(T) == np.complex128:
        Bg = Bg + 1j * rng.uniform(-1, 1, Bg.shape)
    la.clear_plan_cache()
    res = spmd(backends(P, T, Ti), _spmm_body, A, Bg, rowp, bpart)
    itype = "i32" if Ti == np.int32 else "i64"
    olocs = orc.distribute(A, P, itype=itype)
    C_ref = orc.matmat(olocs, Bg, bpart)
    BT = np.resize(Bg, (A.shape[0], ncols))
    CT_ref = orc.matmat(orc.transpose(olocs), BT)
    for C, cols, CT, info in res:
        assert relerr(C, C_ref) <= TOL[np.dtype(T)] and relerr(CT, CT_ref) <= TOL[np.dtype(T)]
        assert relerr(C, cols) <= TOL[np.dtype(T)]
        if case in ("poisson", "laplace2d_f32"):
            assert info["lanes_per_row"] == 1 and np.array_equal(C, C_ref), "one lane per row: the reference's summation order, every column"


# ---------------------------------------------------------------------------------------------------------------
# repartition on the device and the partition-mismatch fallbacks of dot / axpby
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("T", [np.float64, np.complex128, np.float32])
@pytest.mark.parametrize("P", [1, 2, 4])
def test_repartition_and_mismatched_partitions(T, P):
    rng = np.random.default_rng(7 + P)
    n = 1000
    v = rng.uniform(-1, 1, n).astype(T)
    w = rng.uniform(-1, 1, n).astype(T)
    if np.dtype(T).kind == "c":
        v = v + 1j * rng.uniform(-1, 1, n)
    old = np.concatenate([[1], np.sort(rng.integers(1, n + 2, size=P - 1)), [n + 1]]).astype(np.int64)
    new = np.concatenate([[1], np.sort(rng.integers(1, n + 2, size=P - 1)), [n + 1]]).astype(np.int64)
    if P > 1:
        new[1] = new[0]  # an empty rank in the target partition

    def body(rank, bs):
        b = bs[rank]
        torch.cuda.set_device(b.torch_device())
        x = la.HPCVector.from_global(v, b, partition=old)
        y = la.repartition(x, new)
        assert y.partition.tolist() == new.tolist() and y.v.is_cuda and y.local_size == int(new[rank + 1] - new[rank])
        assert la.repartition(x, old) is x  # the reference's fast path (test/test_repartition.jl:67-69)
        plan = la.get_repartition_plan(x, new)
        y2 = la.repartition(x, new)
        assert la.get_repartition_plan(x, new) is plan and y2.structural_hash == y.structural_hash  # memoised (:684-693)
        z = la.HPCVector.from_global(w, b, partition=new)
        d = la.dot(x, z)  # z is repartitioned to x's partition (src/vectors.jl:806-811)
        la.axpby(2.0, x, 1.0, z)  # x is repartitioned to z's
        return y.to_global(), d, z.to_global()

    la.clear_plan_cache()
    for yg, d, zg in spmd(backends(P, T, np.int64), body):
        assert np.array_equal(yg, v)
        assert abs(d - np.vdot(v, w)) <= (1e-3 if T == np.float32 else 1e-10) * n
        assert relerr(zg, 2 * v + w) <= TOL[np.dtype(T)]


# ---------------------------------------------------------------------------------------------------------------
# device-side transpose: same arrays as the host builder and the oracle, bit for bit
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("T,Ti", [(np.float64, np.int32), (np.complex128, np.int64), (np.float32, np.int32)])
def test_device_transpose_matches_host_builder(T, Ti, monkeypatch):
    rng = np.random.default_rng(77)
    S = la.synth
    b = la.backend_cuda_serial(T, Ti)
    mats = [_ragged(rng, 400, 300, 0.03, T, long_rows=((7, 250),)), _ragged(rng, 50, 2000, 0.01, T), sp.csr_matrix((5, 9), dtype=T)]
    rp, c, v = S.stencil_local(2, 12, 0, 12**3, T, Ti)
    mats.append(sp.csr_matrix((v, c - 1, rp - 1), shape=(12**3, 12**3)))
    for M in mats:
        monkeypatch.setenv("HPCLA_TRANSPOSE", "device")
        A = la.HPCSparseMatrix.from_global(M, b)
        Yd = la.materialize_transpose(A)
        monkeypatch.setenv("HPCLA_TRANSPOSE", "host")
        A2 = la.HPCSparseMatrix.from_global(M, b)
        Yh = la.materialize_transpose(A2)
        assert Yd is not Yh and Yd.nrows_local == Yh.nrows_local and Yd.ncols_compressed == Yh.ncols_compressed
        for name in ("rowptr", "colval", "col_indices"):
            assert np.array_equal(getattr(Yd, name), getattr(Yh, name)), name
        assert np.array_equal(Yd.nzval_host(), Yh.nzval_host())
        assert np.array_equal(Yd.rowptr_target.cpu().numpy(), Yh.rowptr) and np.array_equal(Yd.colval_target.cpu().numpy(), Yh.colval)
        ref = sp.csr_matrix(M.T)
        ref.sort_indices()
        assert np.array_equal(Yd.rowptr - 1, ref.indptr) and np.array_equal(Yd.col_indices[Yd.colval - 1] - 1, ref.indices)
        assert la.materialize_transpose(A) is Yd and la.materialize_transpose(Yd) is A  # cached both ways (src/sparse.jl:1858-1859)


@pytest.mark.parametrize("Ti", [np.int32, np.int64])
def test_cg_fused_dot_matches_unfused(Ti, monkeypatch):
    """hpcla_cg lets p.q ride on the multiply (one partial per row-walk CTA) for stencil-like matrices; the residual
    history agrees with the separate dot kernel and with the oracle's textbook CG.  Int64 indices exercise the switch
    away from the direct row walk while the dot is requested; a ragged matrix (general tiles) keeps the separate dot."""
    T = np.float64
    b = la.backend_cuda_serial(T, Ti)
    N = 20
    n = N**3
    A = la.synth.stencil_matrix(1, N, b)
    rp, c, v = la.synth.stencil_local(1, N, 0, n, T, Ti)
    G = sp.csr_matrix((v, c - 1, rp - 1), shape=(n, n))
    bvec = (G @ np.ones(n)).astype(T)
    bv = la.HPCVector.from_global(bvec, b)
    sol_f, hist_f = la.cg(A, bv, 30)
    monkeypatch.setenv("HPCLA_CG_UNFUSED", "1")
    sol_u, hist_u = la.cg(A, bv, 30)
    monkeypatch.delenv("HPCLA_CG_UNFUSED")
    xo, ho = orc.cg(orc.distribute(G, 1, itype="i32" if Ti == np.int32 else "i64"), bvec, 30)
    # the two dots round differently and CG amplifies it: agreement to 1e-7 over 30 iterations, not to the last bits
    assert np.allclose(hist_f, hist_u, rtol=1e-7, atol=1e-12 * ho[0]) and np.allclose(hist_f, ho, rtol=1e-7, atol=1e-12 * ho[0])
    assert relerr(sol_f.to_global(), xo) <= 1e-9 and relerr(sol_u.to_global(), xo) <= 1e-9
    # a matrix with general tiles: SPD by construction, the separate dot kernel runs
    rng = np.random.default_rng(3)
    R = _ragged(rng, 600, 600, 0.02, T, long_rows=((11, 400),))
    S = sp.csr_matrix(R @ R.T + sp.identity(600) * 5.0)
    S.sort_indices()
    As = la.HPCSparseMatrix.from_global(S, b)
    bs = rng.uniform(-1, 1, 600)
    sol, hist = la.cg(As, la.HPCVector.from_global(bs, b), 20)
    xo, ho = orc.cg(orc.distribute(S, 1, itype="i32" if Ti == np.int32 else "i64"), bs, 20)
    assert relerr(sol.to_global(), xo) <= 1e-6 and np.allclose(hist, ho, rtol=1e-5, atol=1e-8 * ho[0])  # below that the history is rounding noise (the oracle itself is not monotone there)


# ---------------------------------------------------------------------------------------------------------------
# sparse x sparse: memoised symbolic product + numeric kernel vs the restatement of Base.:*(A, B)
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("T,Ti", [(np.float64, np.int32), (np.complex128, np.int64), (np.float32, np.int32)])
@pytest.mark.parametrize("P", [1, 2, 4])
def test_sparse_times_sparse(T, Ti, P):
    rng = np.random.default_rng(300 + P)
    S = la.synth
    rp, c, v = S.stencil_local(1, 10, 0, 1000, T, Ti)
    G = sp.csr_matrix((v, c - 1, rp - 1), shape=(1000, 1000))
    cases = [(G, G), (_ragged(rng, 300, 200, 0.03, T), _ragged(rng, 200, 260, 0.05, T))]
    itype = "i32" if Ti == np.int32 else "i64"
    for A, B in cases:
        A, B = sp.csr_matrix(A).astype(T), sp.csr_matrix(B).astype(T)

        def body(rank, bs):
            b = bs[rank]
            torch.cuda.set_device(b.torch_device())
            Am, Bm = la.HPCSparseMatrix.from_global(A, b), la.HPCSparseMatrix.from_global(B, b)
            C = Am * Bm
            assert isinstance(C, la.HPCSparseMatrix) and C.row_partition.tolist() == Am.row_partition.tolist() and C.col_partition.tolist() == Bm.col_partition.tolist()
            n0 = len(la.spgemm_module._matrix_plan_cache)
            # new values, same structure: the numeric phase alone (in-place writes are seen, the plan is reused)
            Bm.nzval.mul_(2.0)
            C2 = Am * Bm
            assert len(la.spgemm_module._matrix_plan_cache) == n0
            x = S.vector(B.shape[1], b)
            y = C * x  # the product is a first-class operand of the hot path
            return (C.rowptr, C.colval, C.col_indices, C.nzval_host()), C2.nzval_host(), y.to_global()

        la.clear_plan_cache()
        res = spmd(backends(P, T, Ti), body)
        lC = orc.spgemm(orc.distribute(A, P, itype=itype), orc.distribute(B, P, itype=itype), itype=itype)
        xh = S.vector_local(T, S.X_SEED, 0, B.shape[1])
        y_ref = (A @ (B @ xh))
        for r, ((rowptr, colval, ci, vals), vals2, y) in enumerate(res):
            o = lC[r]
            assert np.array_equal(rowptr, o.rowptr) and np.array_equal(colval, o.colval) and np.array_equal(ci, o.col_indices)
            assert np.array_equal(vals, o.nzval), "same terms, same order, products rounded separately: the reference's values bit for bit"
            assert np.array_equal(vals2, 2 * o.nzval)
            assert relerr(y, y_ref) <= (1e-4 if T == np.float32 else 1e-12)


def test_reference_product_fixtures_on_device():
    """The reference's own product tests (sparse*sparse, sparse*dense, transpose(sparse)*dense: literal inputs,
    tests/golden) through the library at 1 and 2 ranks, at the tests' tolerance."""
    from conftest import PRODUCT_FIXTURES

    for fx in PRODUCT_FIXTURES:
        T = np.complex128 if fx["dtype"] == "c128" else np.float64
        A = fixture_matrix(fx["A"])
        for P, Ti in [(1, np.int64), (2, np.int32)]:
            def body(rank, bs):
                b = bs[rank]
                torch.cuda.set_device(b.torch_device())
                Am = la.HPCSparseMatrix.from_global(A, b)
                if fx["kind"] == "sparse*sparse":
                    C = Am * la.HPCSparseMatrix.from_global(fixture_matrix(fx["B"]), b)
                    rows = sp.csr_matrix((C.nzval_host(), C.col_indices[C.colval - 1] - 1, C.rowptr - 1), shape=(C.nrows_local, fx["C"].shape[1])).toarray()
                    return np.concatenate(la.comm_allgather(b.comm, rows), axis=0), None
                Bm = la.HPCMatrix.from_global(fx["Bdense"], b)
                return (Am * Bm).to_global(), (la.transpose(Am) * Bm).to_global()

            la.clear_plan_cache()
            for C, CT in spmd(backends(P, T, Ti), body):
                assert np.max(np.abs(C - fx["C"])) < fx["tol"], fx["name"]
                if CT is not None:
                    assert np.max(np.abs(CT - fx["CT"])) < fx["tol"], fx["name"]
