"""CPU tests pinning the oracle (oracle/) against the reference's own test fixtures (tests/golden/) and against its
independent numpy twin.  No GPU, no product code."""
import numpy as np
import pytest
import scipy.sparse as sp

from conftest import FIXTURES, PLAN_TABLES, fixture_matrix
from oracle import oracle as orc


def test_uniform_partition_matches_reference_docstring():
    # src/HPCLinearAlgebra.jl:270-277: uniform_partition(10, 4) == [1, 4, 7, 9, 11]
    assert orc.uniform_partition(10, 4).tolist() == [1, 4, 7, 9, 11]
    assert orc.np_uniform_partition(10, 4).tolist() == [1, 4, 7, 9, 11]
    for n, P in [(8, 2), (7, 3), (3, 5), (0, 2), (16777216, 8), (1000000, 4)]:
        a, b = orc.uniform_partition(n, P), orc.np_uniform_partition(n, P)
        assert a.tolist() == b.tolist()
        assert a[0] == 1 and a[-1] == n + 1 and np.all(np.diff(a) >= 0)
    assert orc.uniform_partition(16777216, 8).tolist() == [1 + 2097152 * r for r in range(9)]


def test_owner_clamp_and_empty_ranks():
    p = np.array([1, 4, 4, 7, 9], dtype=np.int64)  # rank 1 is empty
    assert [orc.owner(p, g) for g in range(1, 10)] == [0, 0, 0, 2, 2, 2, 3, 3, 3]  # g=9 == last boundary -> clamp
    assert orc.np_owner(p, np.arange(1, 10)).tolist() == [0, 0, 0, 2, 2, 2, 3, 3, 3]


@pytest.mark.parametrize("fx", FIXTURES, ids=[f["name"] for f in FIXTURES])
@pytest.mark.parametrize("itype", ["i32", "i64"])
def test_reference_fixture_matvec(fx, itype):
    """A*x / mul! known answers of the reference's tests (2 ranks, as test/runtests.jl always uses)."""
    A = fixture_matrix(fx)
    locs = orc.distribute(A, fx["nranks"], row_partition=fx.get("row_partition"), itype=itype)
    y = orc.matvec(locs, fx["x"])
    assert np.max(np.abs(y - fx["y"])) < fx["tol"]
    # also at 1, 3 and 4 ranks: the answer does not depend on the partition
    for P in (1, 3, 4):
        y = orc.matvec(orc.distribute(A, P, itype=itype), fx["x"])
        assert np.max(np.abs(y - fx["y"])) < fx["tol"]


@pytest.mark.parametrize("fx", [f for f in FIXTURES if "yT" in f], ids=[f["name"] for f in FIXTURES if "yT" in f])
def test_reference_fixture_transpose_matvec(fx):
    """transpose(A)*x, transpose(x)*A and x'*A known answers."""
    A = fixture_matrix(fx)
    for P in (1, 2, 3):
        locs = orc.distribute(A, P)
        At = orc.transpose(locs)
        xT = fx.get("xT", fx["x"])
        yT = orc.matvec(At, xT)
        assert np.max(np.abs(yT - fx["yT"])) < fx["tol"]
        if "y_adj" in fx:  # x'*A = transpose(transpose(A)*conj(x)), src/vectors.jl:746
            assert np.max(np.abs(orc.matvec(At, np.conj(fx["x"])) - fx["y_adj"])) < fx["tol"]
        if "AT_dense" in fx:
            dense = np.zeros_like(fx["AT_dense"])
            for L in At:
                r0 = int(L.row_partition[L.rank]) - 1
                for i in range(L.nrows_local):
                    for k in range(int(L.rowptr[i]) - 1, int(L.rowptr[i + 1]) - 1):
                        dense[r0 + i, L.col_indices[int(L.colval[k]) - 1] - 1] = L.nzval[k]
            assert np.max(np.abs(dense - fx["AT_dense"])) < fx["tol"]  # test/test_transpose.jl:52-54


@pytest.mark.parametrize("fx", [f for f in FIXTURES if "dot_xy" in f], ids=[f["name"] for f in FIXTURES if "dot_xy" in f])
def test_reference_fixture_dot(fx):
    p = orc.uniform_partition(8, 2)
    xs, ys = orc.split_vector(fx["x"], p), orc.split_vector(fx["dot_with"], p)
    assert abs(orc.dot(xs, ys) - fx["dot_xy"]) < fx["tol"]
    assert abs(orc.dot(xs, xs) - fx["dot_xx"]) < fx["tol"]
    assert abs(orc.norm2(xs) - np.linalg.norm(fx["x"])) < fx["tol"]


def _check_plan_table(locs, plans, table):
    for r, (L, p) in enumerate(zip(locs, plans)):
        t = table[f"rank{r}"]
        assert L.rowptr.tolist() == t["rowptr"]
        assert L.colval.tolist() == t["colval"]
        assert L.col_indices.tolist() == t["col_indices"]
        if "recv_rank_ids" in t:
            assert p.recv_rank_ids.tolist() == t["recv_rank_ids"]
            assert [a.tolist() for a in p.recv_perm] == t["recv_perm"]
            assert p.send_rank_ids.tolist() == t["send_rank_ids"]
            assert [a.tolist() for a in p.send_indices] == t["send_indices"]
            assert p.local_src_indices.tolist() == t["local_src"]
            assert p.local_dst_indices.tolist() == t["local_dst"]
        if "nzval" in t:
            assert np.real(L.nzval).tolist() == t["nzval"]


@pytest.mark.parametrize("name", ["tridiag8", "nonsquare6x8", "repart8x6"])
def test_worked_plan_tables(name):
    """SURVEY App. C worked plans (C.1, C.3, C.5): structure and VectorPlan arrays with ==."""
    fx = next(f for f in FIXTURES if f["name"] == name + "_f64")
    locs = orc.distribute(fixture_matrix(fx), 2, row_partition=fx.get("row_partition"))
    xp = orc.uniform_partition(fx["n"], 2)
    _check_plan_table(locs, orc.vector_plans(locs, xp), PLAN_TABLES[name])
    _check_plan_table(locs, orc.np_vector_plans([L.col_indices for L in locs], xp), PLAN_TABLES[name])


def test_worked_transpose_table():
    """materialised transpose of the 10x8 fixture of test/test_transpose.jl:39-54."""
    fx = next(f for f in FIXTURES if f["name"] == "transpose10x8_f64")
    locs = orc.distribute(fixture_matrix(fx), 2)
    assert locs[0].row_partition.tolist() == [1, 6, 11] and locs[0].col_partition.tolist() == [1, 5, 9]
    At = orc.transpose(locs)
    assert At[0].row_partition.tolist() == [1, 5, 9] and At[0].col_partition.tolist() == [1, 6, 11]
    _check_plan_table(At, [None, None], PLAN_TABLES["transpose10x8_AT"])


def _random_matrix(rng, m, n, density, dtype, empty_rows=True):
    A = sp.random(m, n, density=density, random_state=rng, format="csr", dtype=np.float64)
    A.data = rng.uniform(-1, 1, size=A.nnz)
    if np.dtype(dtype).kind == "c":
        A = A.astype(np.complex128)
        A.data = A.data + 1j * rng.uniform(-1, 1, size=A.nnz)
    if empty_rows and m > 4:
        A = sp.csr_matrix(sp.diags(np.r_[0.0, np.ones(m - 2), 0.0]) @ A)  # first and last rows empty
        A.eliminate_zeros()
    A.sort_indices()
    return A.astype(dtype)


@pytest.mark.parametrize("dtype", ["f32", "f64", "c128"])
@pytest.mark.parametrize("P", [1, 2, 3, 5])
def test_cpp_vs_numpy_twin_random(dtype, P):
    """C++ restatement == numpy twin on random ragged matrices: plans with ==, y bit-for-bit (same order, no FMA)."""
    rng = np.random.default_rng(1234 + P)
    dt = orc.DTYPES[dtype]
    for (m, n) in [(37, 41), (64, 64), (5, 90), (90, 5)]:
        A = _random_matrix(rng, m, n, 0.15, dt)
        rp = None
        if P == 3:  # a non-uniform partition with an EMPTY rank
            rp = np.array([1, 1 + m // 2, 1 + m // 2, m + 1], dtype=np.int64)
        locs = orc.distribute(A, P, row_partition=rp, itype="i32")
        xp = orc.uniform_partition(n, P)
        p_cpp = orc.vector_plans(locs, xp)
        p_np = orc.np_vector_plans([L.col_indices for L in locs], xp)
        for a, b in zip(p_cpp, p_np):
            assert a.send_rank_ids.tolist() == b.send_rank_ids.tolist()
            assert a.recv_rank_ids.tolist() == b.recv_rank_ids.tolist()
            assert [s.tolist() for s in a.send_indices] == [s.tolist() for s in b.send_indices]
            assert [s.tolist() for s in a.recv_perm] == [s.tolist() for s in b.recv_perm]
            assert a.local_src_indices.tolist() == b.local_src_indices.tolist()
            assert a.local_dst_indices.tolist() == b.local_dst_indices.tolist()
        for L in locs:
            ci, cv = orc.np_compress(L.col_indices[L.colval.astype(np.int64) - 1])
            assert ci.tolist() == L.col_indices.tolist() and cv.tolist() == L.colval.tolist()
        x = rng.uniform(-1, 1, n).astype(dt)
        if dt.kind == "c":
            x = x + 1j * rng.uniform(-1, 1, n)
        xs = orc.split_vector(x, xp)
        W = orc.PlanWorld(locs, xp)
        g_cpp = W.execute(xs)
        W.close()
        g_np = orc.np_execute(p_np, xs)
        for L, a, b in zip(locs, g_cpp, g_np):
            assert np.array_equal(a, b)
            assert np.array_equal(a, x[L.col_indices - 1])  # postcondition, SURVEY App. B.6
            y_cpp = orc.spmv_local(L, a)
            y_np = orc.np_spmv_local(L.rowptr, L.colval, L.nzval, b)
            assert np.array_equal(y_cpp, y_np)
        y = orc.matvec(locs, x)
        ref = A @ x
        tol = 1e-5 if dtype == "f32" else 1e-13
        assert np.linalg.norm(y - ref) <= tol * max(np.linalg.norm(ref), 1e-30)


@pytest.mark.parametrize("P", [1, 2, 4])
def test_transpose_cpp_vs_scipy(P):
    rng = np.random.default_rng(7)
    for dtype in ("f64", "c128"):
        A = _random_matrix(rng, 53, 31, 0.2, orc.DTYPES[dtype])
        locs = orc.distribute(A, P)
        a, b = orc.transpose(locs), orc.np_transpose(locs)
        for La, Lb in zip(a, b):
            assert La.rowptr.tolist() == Lb.rowptr.tolist()
            assert La.colval.tolist() == Lb.colval.tolist()
            assert La.col_indices.tolist() == Lb.col_indices.tolist()
            assert np.array_equal(La.nzval, Lb.nzval)
        # transposing twice gives back A (cached_transpose is bidirectional, src/sparse.jl:1858-1859)
        back = orc.transpose(a)
        for L0, L2 in zip(locs, back):
            assert L0.rowptr.tolist() == L2.rowptr.tolist() and L0.colval.tolist() == L2.colval.tolist()
            assert np.array_equal(L0.nzval, L2.nzval)


def test_segments_are_contiguous_and_ordered_by_owner():
    """gathered = [from rank 0 | ... | own | ... | from rank P-1], every recv_perm a consecutive range."""
    rng = np.random.default_rng(3)
    A = _random_matrix(rng, 200, 200, 0.05, np.float64, empty_rows=False)
    for P in (2, 4, 7):
        locs = orc.distribute(A, P)
        for r, p in enumerate(orc.vector_plans(locs, orc.uniform_partition(200, P))):
            segs = [(int(q), perm) for q, perm in zip(p.recv_rank_ids, p.recv_perm)] + [(r, p.local_dst_indices)]
            segs.sort(key=lambda s: s[0])
            cat = np.concatenate([s[1] for s in segs]) if segs else np.array([])
            assert cat.tolist() == list(range(1, p.n_gathered + 1))
            for s in p.send_indices:
                assert np.all(np.diff(s) > 0)


def test_cpu_baseline_runner_matches_matvec():
    rng = np.random.default_rng(11)
    A = _random_matrix(rng, 300, 300, 0.03, np.float64)
    for P in (1, 3):
        locs = orc.distribute(A, P, itype="i32")
        xp = orc.uniform_partition(300, P)
        x = rng.uniform(-1, 1, 300)
        times, ys = orc.bench_spmv(locs, orc.split_vector(x, xp), xp, warmup=1, reps=2)
        assert len(times) == 2 and np.all(times > 0)
        assert np.array_equal(np.concatenate(ys), orc.matvec(locs, x))


def test_cg_restatement_converges_on_laplacian():
    n = 64
    A = sp.diags([-1.0, 2.0, -1.0], [-1, 0, 1], shape=(n, n), format="csr")
    locs = orc.distribute(A, 2)
    b = A @ np.ones(n)
    x, hist = orc.cg(locs, b, 70)
    assert np.linalg.norm(x - 1.0) < 1e-8


def test_reference_product_fixtures():
    """The oracle's restatements of the hot path's callers against the literal inputs of the reference's own product
    tests (test/test_matrix_multiplication.jl:38-92, test/test_new_operations.jl:46-88) at the tests' tolerance."""
    from conftest import PRODUCT_FIXTURES, fixture_matrix

    assert len(PRODUCT_FIXTURES) == 6
    for fx in PRODUCT_FIXTURES:
        A = fixture_matrix(fx["A"])
        for P in (1, 2, 3):
            lA = orc.distribute(A, P)
            if fx["kind"] == "sparse*sparse":
                B = fixture_matrix(fx["B"])
                C = orc.to_global(orc.spgemm(lA, orc.distribute(B, P)), fx["C"].shape).toarray()
                assert np.max(np.abs(C - fx["C"])) < fx["tol"], fx["name"]
            else:
                assert np.max(np.abs(orc.matmat(lA, fx["Bdense"]) - fx["C"])) < fx["tol"], fx["name"]
                assert np.max(np.abs(orc.matmat(orc.transpose(lA), fx["Bdense"]) - fx["CT"])) < fx["tol"], fx["name"]


def test_rowcheck_agrees_with_the_full_oracle():
    """oracle/rowcheck.py (rows regenerated from the generators) == the distributed oracle, bit for bit, on every rank's
    boundary planes and random runs; and it notices a single wrong ghost value."""
    import hpcla_synth as S
    from oracle import rowcheck as rc

    for kind, grid, T, Ti in [(1, (14, 12, 10), np.float64, np.int32), (2, (9, 9, 9), np.complex128, np.int32), (0, (40, 30, 1), np.float64, np.int64),
                              (3, (20000, 1, 1), np.float32, np.int32)]:
        n = rc.n_rows(kind, grid)
        rp, c, v = rc._rows(kind, grid, 0, n, T, Ti)
        A = sp.csr_matrix((v, c.astype(np.int64) - 1, rp.astype(np.int64) - 1), shape=(n, n))
        x = S.vector_local(T, S.X_SEED, 0, n)
        locs = orc.distribute(A, 1, itype="i32" if Ti == np.int32 else "i64")
        y = orc.matvec(locs, x)
        for P in (1, 3):
            part = orc.uniform_partition(n, P)
            for r in range(P):
                b, e = int(part[r]) - 1, int(part[r + 1]) - 1
                rows, worst, exact, nr = rc.check_rows(kind, grid, b, e, y[b:e], T, Ti, seed=r, run=64)
                assert worst == 0.0 and exact == nr and rows > 0
        if kind != 3:
            yT = orc.matvec(orc.transpose(locs), x)
            rows, worst, exact, nr = rc.check_rows(kind, grid, 0, n, yT, T, Ti, transpose=True, run=64)
            assert worst == 0.0 and exact == nr
        # a multiply that used x[j] + 1 for one ghost of rank 1 (a mis-routed halo element) is caught
        if kind == 1:
            part = orc.uniform_partition(n, 2)
            b, e = int(part[1]) - 1, int(part[2]) - 1
            xb = x.copy()
            xb[b - 3] += 1.0  # an element owned by rank 0 that rank 1 reads as a ghost
            yb = A @ xb
            _, worst, exact, nr = rc.check_rows(kind, grid, b, e, yb[b:e], T, Ti, seed=1, run=64)
            assert worst > 1e-6 and exact < nr


def test_reference_arm_on_synthetic_workloads():
    """orc_bench_synth (the CPU reference arm: every worker generates and first-touches its own block) computes the same
    y as the oracle on a matrix built outside it, for every workload kind; and its CG leg runs."""
    import hpcla_synth as S

    for kind, grid, T, Ti in [(1, (12, 10, 8), np.float64, np.int32), (2, (9, 9, 9), np.complex128, np.int64), (3, (5000, 1, 1), np.float32, np.int32),
                              (0, (30, 20, 1), np.float64, np.int64)]:
        n = grid[0] if kind == 3 else S.stencil_rows(kind, grid)
        t, nnz, yn = orc.bench_synth(kind, grid, T, Ti, workers=3, warmup=1, reps=2)
        if kind == 3:
            rp, c, v = S.powerlaw_local(n, S.POWERLAW_SEED, S.POWERLAW_MAX_LEN, 0, n, T, Ti)
        else:
            rp, c, v = S.stencil_local(kind, grid, 0, n, T, Ti)
        A = sp.csr_matrix((v, c.astype(np.int64) - 1, rp.astype(np.int64) - 1), shape=(n, n))
        y = orc.matvec(orc.distribute(A, 1, itype="i32" if Ti == np.int32 else "i64"), S.vector_local(T, S.X_SEED, 0, n))
        assert nnz == A.nnz and len(t) == 2 and np.all(t > 0)
        assert abs(yn - float(np.sum(np.abs(y.astype(np.complex128)) ** 2))) <= 1e-9 * yn
    t, nnz, _ = orc.bench_synth(1, (12, 10, 8), np.float64, np.int32, workers=4, reps=3, op="cg")
    assert len(t) == 3 and np.all(t > 0)
    t, nnz, _ = orc.bench_synth(1, (12, 10, 8), np.float64, np.int32, workers=2, reps=2, inner=4)
    assert len(t) == 2
