"""CPU tests of the product's host-side logic (libhpcla_b200.so host functions + the python mirror of the reference
API) against the oracle.  No GPU: DeviceCPU backends carry structure only; multi-rank runs use a single-process world
of rank-threads and a 2-process gloo world."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import scipy.sparse as sp

import hpcla_b200 as la
from conftest import FIXTURES, PLAN_TABLES, ROOT, fixture_matrix
from oracle import oracle as orc


def _declared_symbols(header="hpcla_b200.h"):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return set(re.findall(r"\b(hpcla_[a-z0-9_]+)\s*\(", src))


def test_library_loads_and_exports_every_declared_symbol():
    L = la._lib.lib()
    declared = _declared_symbols()
    assert len(declared) >= 45
    out = subprocess.check_output(["nm", "-D", "--defined-only", la.LIB_PATH], text=True)
    exported = set(re.findall(r" T (hpcla_[a-z0-9_]+)", out))
    assert declared <= exported, f"missing: {sorted(declared - exported)}"
    assert declared == set(la._lib.SIGNATURES), f"binding/header mismatch: {sorted(declared ^ set(la._lib.SIGNATURES))}"
    assert L.hpcla_abi_version() == 1
    # the synthetic generators are test infrastructure in a library of their own: the product exports none of them
    assert not any(n.startswith("hpcla_synth_") for n in exported)
    import hpcla_synth

    hpcla_synth.lib()
    out = subprocess.check_output(["nm", "-D", "--defined-only", hpcla_synth.LIB_PATH], text=True)
    exported = set(re.findall(r" T (hpcla_[a-z0-9_]+)", out))
    assert _declared_symbols("hpcla_synth.h") == exported == set(hpcla_synth.SIGNATURES)


def test_no_device_means_loud_failure_not_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    b = la.backend_cuda_serial(np.float64, np.int32)
    with pytest.raises(la.HPCLAError):
        b.ctx()
    bc = la.backend_cpu_serial()
    A = la.HPCSparseMatrix.from_global(sp.identity(4, format="csr"), bc)
    x = la.HPCVector.from_global(np.ones(4), bc)
    with pytest.raises(la.HPCLAError):
        la.matvec(A, x)
    with pytest.raises(la.HPCLAError):
        la.mul(x.similar(), A, x)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "linearalgebrampi.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+\.*oracle|libhpcla_oracle|orc_[a-z_]+\(", src, flags=re.M), f


def test_uniform_partition_and_compress_match_oracle():
    for n, P in [(10, 4), (8, 2), (3, 5), (0, 3), (1000000, 4), (16777216, 8)]:
        assert la.uniform_partition(n, P).tolist() == orc.uniform_partition(n, P).tolist()
    rng = np.random.default_rng(0)
    for Ti in (np.int32, np.int64):
        for ncols, nnz in [(50, 200), (1000, 10), (7, 0), (100000, 300000)]:
            g = rng.integers(1, ncols + 1, size=nnz).astype(Ti)
            import ctypes

            colval = np.empty(nnz, dtype=Ti)
            ci = np.empty(min(nnz, ncols), dtype=np.int64)
            ncc = ctypes.c_int64()
            la._lib.check(la._lib.lib().hpcla_compress_columns(la._lib.itype_code(Ti), nnz, la._lib.ptr(g), ncols, la._lib.ptr(colval), la._lib.ptr(ci), ctypes.byref(ncc)))
            oci, ocv = orc.compress(g.astype(np.int64))
            assert ci[: ncc.value].tolist() == oci.tolist()
            assert colval.astype(np.int64).tolist() == ocv.tolist()


def _compare_matrix(A, L):
    assert A.row_partition.tolist() == L.row_partition.tolist()
    assert A.col_partition.tolist() == L.col_partition.tolist()
    assert A.col_indices.tolist() == L.col_indices.tolist()
    assert A.rowptr.dtype == L.rowptr.dtype and A.rowptr.tolist() == L.rowptr.tolist()
    assert A.colval.dtype == L.colval.dtype and A.colval.tolist() == L.colval.tolist()
    assert np.array_equal(A.nzval_host(), L.nzval)
    assert A.nrows_local == L.nrows_local and A.ncols_compressed == L.ncols_compressed


def _compare_plan(p, o):
    assert p.send_rank_ids.tolist() == o.send_rank_ids.tolist()
    assert p.recv_rank_ids.tolist() == o.recv_rank_ids.tolist()
    assert [a.tolist() for a in p.send_indices] == [a.tolist() for a in o.send_indices]
    assert [a.tolist() for a in p.recv_perm] == [a.tolist() for a in o.recv_perm]
    assert p.local_src_indices.tolist() == o.local_src_indices.tolist()
    assert p.local_dst_indices.tolist() == o.local_dst_indices.tolist()
    assert p.n_gathered == o.n_gathered


def _spmd_structure_check(rank, backends, A_global, row_partition, x_partition, olocs, oplans, oT):
    b = backends[rank]
    A = la.HPCSparseMatrix.from_global(A_global, b, row_partition=row_partition)
    _compare_matrix(A, olocs[rank])
    n = A_global.shape[1]
    x = la.HPCVector.from_global(np.arange(1, n + 1, dtype=np.float64).astype(b.T), b, partition=x_partition)
    plan = la.get_vector_plan(A, x)
    assert plan.local_src_indices.dtype == b.Ti
    _compare_plan(plan, oplans[rank])
    At = la.materialize_transpose(A)
    _compare_matrix(At, oT[rank])
    assert la.materialize_transpose(A) is At and At.cached_transpose is A  # src/sparse.jl:1849-1859
    return True


def _run_structure_case(A_global, P, T, Ti, row_partition=None, x_partition=None):
    A_global = sp.csr_matrix(A_global).astype(T)
    itype = "i32" if Ti == np.int32 else "i64"
    olocs = orc.distribute(A_global, P, row_partition=row_partition, itype=itype)
    xp = orc.uniform_partition(A_global.shape[1], P) if x_partition is None else np.asarray(x_partition, dtype=np.int64)
    oplans = orc.vector_plans(olocs, xp)
    oT = orc.transpose(olocs)
    la.clear_plan_cache()
    if P == 1:
        assert _spmd_structure_check(0, [la.backend_cpu_serial(T, Ti)], A_global, row_partition, xp, olocs, oplans, oT)
    else:
        bs = la.backends_threads(P, T, Ti, cuda=False)
        assert all(bs[0].comm.world.run(_spmd_structure_check, bs, A_global, row_partition, xp, olocs, oplans, oT))


@pytest.mark.parametrize("fx", FIXTURES, ids=[f["name"] for f in FIXTURES])
def test_reference_fixtures_structure_plan_transpose(fx):
    T = np.complex128 if fx["dtype"] == "c128" else np.float64
    for Ti in (np.int32, np.int64):
        _run_structure_case(fixture_matrix(fx), 2, T, Ti, row_partition=fx.get("row_partition"))
    _run_structure_case(fixture_matrix(fx), 1, T, np.int64)
    _run_structure_case(fixture_matrix(fx), 3, T, np.int32)


@pytest.mark.parametrize("name", ["tridiag8", "nonsquare6x8", "repart8x6"])
def test_worked_plan_tables_through_the_library(name):
    fx = next(f for f in FIXTURES if f["name"] == name + "_f64")

    def body(rank, bs):
        A = la.HPCSparseMatrix.from_global(fixture_matrix(fx), bs[rank], row_partition=fx.get("row_partition"))
        x = la.HPCVector.from_global(fx["x"], bs[rank])
        p = la.get_vector_plan(A, x)
        t = PLAN_TABLES[name][f"rank{rank}"]
        assert A.rowptr.tolist() == t["rowptr"] and A.colval.tolist() == t["colval"] and A.col_indices.tolist() == t["col_indices"]
        assert p.recv_rank_ids.tolist() == t["recv_rank_ids"] and [a.tolist() for a in p.recv_perm] == t["recv_perm"]
        assert p.send_rank_ids.tolist() == t["send_rank_ids"] and [a.tolist() for a in p.send_indices] == t["send_indices"]
        assert p.local_src_indices.tolist() == t["local_src"] and p.local_dst_indices.tolist() == t["local_dst"]
        return True

    la.clear_plan_cache()
    bs = la.backends_threads(2, np.float64, np.int64, cuda=False)
    assert all(bs[0].comm.world.run(body, bs))


@pytest.mark.parametrize("P", [1, 2, 3, 5, 8])
def test_random_ragged_matrices(P):
    rng = np.random.default_rng(100 + P)
    for (m, n, dens) in [(40, 40, 0.2), (64, 23, 0.1), (9, 120, 0.3), (130, 130, 0.02)]:
        A = sp.random(m, n, density=dens, random_state=rng, format="csr")
        A.data[:] = rng.uniform(-1, 1, A.nnz)
        rp = None
        if P == 3:
            rp = np.array([1, 1 + m // 3, 1 + m // 3, m + 1], dtype=np.int64)  # an empty rank
        _run_structure_case(A, P, np.float64, np.int32, row_partition=rp)
        _run_structure_case(A.astype(np.complex128) * (1 + 0.5j), P, np.complex128, np.int64, row_partition=rp)
    # x partitioned differently from the columns' uniform partition
    A = sp.random(30, 50, density=0.2, random_state=rng, format="csr")
    xp = np.array([1] + sorted(rng.integers(1, 51, size=P - 1).tolist()) + [51], dtype=np.int64)
    _run_structure_case(A, P, np.float32, np.int32, x_partition=xp)


def test_memoisation_and_cache_behaviour():
    """same structure + same partition + same types -> no second plan construction."""
    b = la.backend_cpu_serial(np.float64, np.int64)
    A1 = la.HPCSparseMatrix.from_global(fixture_matrix(FIXTURES[0]), b)
    A2 = la.HPCSparseMatrix.from_global(fixture_matrix(FIXTURES[0]) * 3.0, b)  # same structure, other values
    x = la.HPCVector.from_global(np.ones(8), b)
    la.clear_plan_cache()
    n0 = la.sparse.plan_build_count
    p1 = la.get_vector_plan(A1, x)
    assert la.sparse.plan_build_count == n0 + 1 and la.cache_sizes()["vector_plan"] == 1
    assert la.get_vector_plan(A1, x) is p1 and la.get_vector_plan(A2, x) is p1
    assert la.sparse.plan_build_count == n0 + 1
    b32 = la.backend_cpu_serial(np.float64, np.int32)  # another Ti -> another key (src/sparse.jl:1994)
    A3 = la.HPCSparseMatrix.from_global(fixture_matrix(FIXTURES[0]), b32)
    assert la.get_vector_plan(A3, la.HPCVector.from_global(np.ones(8), b32)) is not p1
    la.clear_plan_cache()
    assert la.cache_sizes()["vector_plan"] == 0
    assert la.get_vector_plan(A1, x) is not p1
    A1.invalidate_structure()
    assert A1.structural_hash is None and A1.cached_transpose is None


def test_errors_mirror_the_reference():
    b = la.backend_cpu_serial()
    with pytest.raises(ValueError):  # src/backends.jl:460-464
        la.assert_backends_compatible(b, la.backend_cuda_serial())
    assert la.backends_compatible(b, la.backend_cpu_serial(np.float32, np.int32))
    assert la.retype_backend(b, np.float32).T == np.float32 and la.retype_backend(b, np.float64) is b
    bs = la.backends_threads(2, cuda=False)

    def body(rank, bs):
        rowptr = np.array([1, 2], dtype=np.int64)
        with pytest.raises(ValueError, match="same number of columns"):  # src/sparse.jl:487-490
            la.HPCSparseMatrix.from_local(rowptr, np.array([1]), np.array([1.0]), 4 + rank, bs[rank])
        return True

    assert all(bs[0].comm.world.run(body, bs))
    with pytest.raises(la.HPCLAError):
        la.backends.HPCBackend(np.float16, np.int64, la.DeviceCPU(), la.CommSerial(), la.SolverMUMPS())


def test_synthetic_generators():
    S = la.synth
    # 2-D 5-point Laplacian == kron(I, L1) + kron(L1, I) (tools/benchmark_vs_petsc.jl:42-49)
    N = 7
    L1 = sp.diags([-1.0, 2.0, -1.0], [-1, 0, 1], shape=(N, N))
    ref = sp.csr_matrix(sp.kron(sp.identity(N), L1) + sp.kron(L1, sp.identity(N)))
    rowptr, cols, vals = S.stencil_local(S.LAPLACE2D_5PT, N, 0, N * N, np.float64, np.int32)
    got = sp.csr_matrix((vals, cols - 1, rowptr - 1), shape=(N * N, N * N))
    assert abs(got - ref).max() == 0 and np.all(np.diff(cols[rowptr[3] - 1 : rowptr[4] - 1]) > 0)
    # 3-D 7-point
    N = 5
    ref = sp.csr_matrix(sp.kron(sp.kron(sp.identity(N), sp.identity(N)), L1.tocsr()[:N, :N]) * 0)
    L1 = sp.diags([-1.0, 2.0, -1.0], [-1, 0, 1], shape=(N, N))
    I = sp.identity(N)
    ref = sp.csr_matrix(sp.kron(sp.kron(I, I), L1) + sp.kron(sp.kron(I, L1), I) + sp.kron(sp.kron(L1, I), I))
    rowptr, cols, vals = S.stencil_local(S.POISSON3D_7PT, N, 0, N**3, np.float64, np.int64)
    got = sp.csr_matrix((vals, cols - 1, rowptr - 1), shape=(N**3, N**3))
    assert abs(got - ref).max() == 0
    # counts of SURVEY App. A: nnz = 5n-4N, 7n-6N^2, (3N-2)^3
    import hpcla_synth

    L = hpcla_synth.lib()
    assert L.hpcla_synth_stencil_nnz(0, 1000, 1000, 1, 0, 10**6) == 4996000
    assert L.hpcla_synth_stencil_nnz(1, 64, 64, 64, 0, 64**3) == 7 * 64**3 - 6 * 64**2
    assert L.hpcla_synth_stencil_nnz(2, 12, 12, 12, 0, 12**3) == (3 * 12 - 2) ** 3
    # non-cubic grid (weak-scaling slabs): same operator as the kron construction
    nx, ny, nz = 4, 3, 5
    Lx, Ly, Lz = [sp.diags([-1.0, 2.0, -1.0], [-1, 0, 1], shape=(k, k)) for k in (nx, ny, nz)]
    ref = sp.csr_matrix(sp.kron(sp.kron(sp.identity(nz), sp.identity(ny)), Lx) + sp.kron(sp.kron(sp.identity(nz), Ly), sp.identity(nx))
                        + sp.kron(sp.kron(Lz, sp.identity(ny)), sp.identity(nx)))
    rp3, c3, v3 = S.stencil_local(S.POISSON3D_7PT, (nx, ny, nz), 0, nx * ny * nz, np.float64, np.int32)
    assert abs(sp.csr_matrix((v3, c3 - 1, rp3 - 1), shape=(60, 60)) - ref).max() == 0
    # row slices agree with the whole, for the 27-point complex stencil
    N = 6
    rp, c, v = S.stencil_local(S.STENCIL3D_27PT, N, 0, N**3, np.complex128, np.int32)
    rp2, c2, v2 = S.stencil_local(S.STENCIL3D_27PT, N, 50, 120, np.complex128, np.int32)
    lo, hi = rp[50] - 1, rp[120] - 1
    assert np.array_equal(c[lo:hi], c2) and np.array_equal(v[lo:hi], v2) and np.array_equal(rp[50:121] - rp[50] + 1, rp2)
    A = sp.csr_matrix((v, c - 1, rp - 1), shape=(N**3, N**3))
    assert abs(A - A.T).max() > 1e-3  # non-symmetric by construction
    # power law: ascending distinct columns, heavy tail
    n = 20000
    rp, c, v = S.powerlaw_local(n, 0xC4, 5000, 0, n, np.float32, np.int32)
    lens = np.diff(rp)
    assert lens.min() >= 7 and lens.max() > 500 and np.median(lens) <= 12
    for r in np.argsort(lens)[-3:]:
        seg = c[rp[r] - 1 : rp[r + 1] - 1]
        assert np.all(np.diff(seg) > 0) and seg[0] >= 1 and seg[-1] <= n
    x = S.vector_local(np.complex128, S.X_SEED, 10, 20)
    assert np.array_equal(x, S.vector_local(np.complex128, S.X_SEED, 0, 30)[10:20]) and np.all(np.abs(x.real) <= 1)


def test_c1_structure_known_answer():
    """2-D 5-point N=1000 at P=4."""
    N, P = 1000, 4
    bs = la.backends_threads(P, np.float64, np.int64, cuda=False)

    def body(rank, bs):
        A = la.synth.stencil_matrix(la.synth.LAPLACE2D_5PT, N, bs[rank])
        x = la.synth.vector(N * N, bs[rank])
        p = la.get_vector_plan(A, x)
        return (A.row_partition.tolist(), A.nnz_local, A.ncols_compressed, p.recv_rank_ids.tolist(),
                [(int(a[0]), int(a[-1])) for a in p.recv_perm], [(int(a[0]), int(a[-1])) for a in p.send_indices],
                (int(p.local_dst_indices[0]), int(p.local_dst_indices[-1])))

    la.clear_plan_cache()
    res = bs[0].comm.world.run(body, bs)
    assert res[0][0] == [1, 250001, 500001, 750001, 1000001]
    assert [r[1] for r in res] == [1248500, 1249500, 1249500, 1248500]
    assert [r[2] for r in res] == [251000, 252000, 252000, 251000]
    assert res[1][3] == [0, 2] and res[1][4] == [(1, 1000), (251001, 252000)] and res[1][5] == [(1, 1000), (249001, 250000)]
    assert res[1][6] == (1001, 251000)
    assert res[0][4] == [(250001, 251000)] and res[0][5] == [(249001, 250000)]
    assert res[3][4] == [(1, 1000)] and res[3][5] == [(1, 1000)] and res[3][6] == (1001, 251000)


def test_gloo_world_size_2():
    """The N>1 host path over a real process group (gloo, 2 processes): plans and transposes == oracle."""
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.path.join(ROOT, "tests"))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tests", "_dist_worker.py")]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("DIST_OK") == 2


def test_hpcmatrix_structure_and_no_cpu_product():
    """HPCMatrix constructors (src/dense.jl:125-202) on a structure-only backend; A*B refuses to compute on the CPU."""
    P = 3
    bs = la.backends_threads(P, np.float64, np.int64, cuda=False)
    M = np.arange(7 * 4, dtype=np.float64).reshape(7, 4)
    R = sp.random(5, 7, density=0.5, random_state=np.random.default_rng(3), format="csr")

    def body(rank, bs):
        b = bs[rank]
        B = la.HPCMatrix.from_global(M, b)
        assert B.row_partition.tolist() == orc.uniform_partition(7, P).tolist() and B.col_partition.tolist() == orc.uniform_partition(4, P).tolist()
        lo, hi = int(B.row_partition[rank]) - 1, int(B.row_partition[rank + 1]) - 1
        assert np.array_equal(B.local_values(), M[lo:hi]) and B.shape == (7, 4)
        col = B.column(2)
        assert col.partition.tolist() == B.row_partition.tolist() and np.array_equal(col.local_values(), M[lo:hi, 2])
        B2 = la.HPCMatrix_local(M[lo:hi], b)
        assert B2.row_partition.tolist() == B.row_partition.tolist() and np.array_equal(B2.to_global(), M)
        with pytest.raises(IndexError):
            B.column(4)
        A = la.HPCSparseMatrix.from_global(R, b)
        with pytest.raises(la.HPCLAError):
            A * B
        return True

    assert all(bs[0].comm.world.run(body, bs))


def test_oracle_matmat_is_the_column_loop():
    rng = np.random.default_rng(11)
    R = sp.random(40, 30, density=0.2, random_state=rng, format="csr")
    Bg = rng.uniform(-1, 1, (30, 5))
    for P in (1, 3):
        C = orc.matmat(orc.distribute(R, P), Bg)
        assert np.linalg.norm(C - R @ Bg) <= 1e-13 * np.linalg.norm(R @ Bg)


def _reference_test_partition(n, nranks):
    """The non-uniform target partition test/test_repartition.jl:44-57 builds (1-based starts)."""
    if nranks < 2:
        return orc.uniform_partition(n, nranks)
    p, total = [1], 0
    for r in range(nranks):
        count = n // nranks - 1 + (1 if r < n % nranks else 0) if r < nranks - 1 else n - total
        total += count
        p.append(total + 1)
    return np.asarray(p, dtype=np.int64)


@pytest.mark.parametrize("P", [1, 2, 3, 4, 7])
def test_repartition_plan_matches_the_restatement(P):
    """hpcla_repartition_plan (bisection over the partitions) == the loop-for-loop restatement of
    VectorRepartitionPlan (src/vectors.jl:519-616), every field, including empty ranks; and the restated execute
    moves the data where the reference's own test expects it (test/test_repartition.jl:38-62)."""
    rng = np.random.default_rng(50 + P)
    cases = [(12, orc.uniform_partition(12, P), _reference_test_partition(12, P))]
    for _ in range(30):
        n = int(rng.integers(0, 40))
        old = np.concatenate([[1], np.sort(rng.integers(1, n + 2, size=P - 1)), [n + 1]]).astype(np.int64)
        new = np.concatenate([[1], np.sort(rng.integers(1, n + 2, size=P - 1)), [n + 1]]).astype(np.int64)
        cases.append((n, old, new))
    for n, old, new in cases:
        for r in range(P):
            got = la.VectorRepartitionPlan(r, P, old, new)
            ref = orc.repartition_plan(r, old, new)
            assert got.send_rank_ids == ref["send_rank_ids"] and got.send_ranges == ref["send_ranges"]
            assert got.recv_rank_ids == ref["recv_rank_ids"] and got.recv_counts == ref["recv_counts"] and got.recv_offsets == ref["recv_offsets"]
            assert got.result_local_size == ref["result_local_size"] and got.local_dst_offset == ref["local_dst_offset"]
            a, b = ref["local_src_range"]
            assert got.local_src_range == ((a, b) if b >= a else (1, 0))
        v = np.arange(1.0, n + 1)
        out = orc.repartition(orc.split_vector(v, old), old, new)
        assert np.array_equal(np.concatenate(out) if out else v, v)
        assert [len(o) for o in out] == np.diff(new).tolist()
    with pytest.raises(la.HPCLAError):
        la.VectorRepartitionPlan(0, P, orc.uniform_partition(5, P), orc.uniform_partition(6, P))


@pytest.mark.parametrize("P", [1, 2, 4])
def test_matrix_plan_and_symbolic_product_match_the_restatement(P):
    """MatrixPlan(A, B) (structure of B[A.col_indices, :], src/sparse.jl:579-897) and the memoised symbolic product
    (hpcla_spgemm_symbolic) against the oracle's restatement of Base.:*(A, B) (src/sparse.jl:991-1059): gathered
    pattern, rowptr, compressed colval and col_indices of every rank's block, bit for bit; rectangular operands, empty
    rows, non-uniform partitions.  No device needed: the symbolic phase is host code."""
    rng = np.random.default_rng(20 + P)
    cases = []
    for (m, k, n, da, db) in [(60, 45, 70, 0.15, 0.1), (33, 80, 20, 0.05, 0.3), (10, 10, 10, 0.0, 0.5)]:
        A = sp.random(m, k, density=da, random_state=rng, format="lil")
        B = sp.random(k, n, density=db, random_state=rng, format="lil")
        if m > 3:
            A[1, :] = 0  # an empty row
        cases.append((sp.csr_matrix(A), sp.csr_matrix(B)))
    for A, B in cases:
        m, k = A.shape
        rpA = np.concatenate([[1], np.sort(rng.integers(1, m + 2, size=P - 1)), [m + 1]]).astype(np.int64)
        rpB = np.concatenate([[1], np.sort(rng.integers(1, k + 2, size=P - 1)), [k + 1]]).astype(np.int64)
        lA = orc.distribute(A, P, row_partition=rpA, itype="i32")
        lB = orc.distribute(B, P, row_partition=rpB, itype="i32")
        lC = orc.spgemm(lA, lB, itype="i32")
        ref = sp.csr_matrix(A @ B)
        assert abs(orc.to_global(lC, ref.shape) - ref).max() <= 1e-13
        bs = la.backends_threads(P, np.float64, np.int32, cuda=False)

        def body(rank, bs):
            b = bs[rank]
            Am = la.HPCSparseMatrix.from_global(A, b, row_partition=rpA)
            Bm = la.HPCSparseMatrix.from_global(B, b, row_partition=rpB)
            plan = la.get_matrix_plan(Am, Bm)
            assert la.get_matrix_plan(Am, Bm) is plan  # memoised (src/sparse.jl:900-916)
            bg = orc.gather_rows(lB, lA[rank].col_indices)
            o = lC[rank]
            assert np.array_equal(plan.bg_rowptr, bg[0]) and np.array_equal(plan.bg_cols, bg[1])
            assert np.array_equal(plan.rowptr, o.rowptr) and np.array_equal(plan.colval, o.colval) and np.array_equal(plan.col_indices, o.col_indices)
            with pytest.raises(la.HPCLAError):
                Am * Bm  # structure-only backend: no CPU arithmetic
            return True

        la.clear_plan_cache()
        assert all(bs[0].comm.world.run(body, bs))


def test_next_row_entry_points_reject_bad_arguments():
    """Argument errors of the host-side entry points added for SURVEY §8(f): status code + message, never an abort."""
    L = la._lib.lib()
    import ctypes

    h = ctypes.c_void_p()
    rowptr = np.array([1, 3], dtype=np.int32)
    colval = np.array([1, 5], dtype=np.int32)  # gathered row 5 of 2
    bg_rowptr = np.array([1, 2, 3], dtype=np.int64)
    bg_cols = np.array([1, 2], dtype=np.int64)
    rc = L.hpcla_spgemm_symbolic(la._lib.I32, 1, la._lib.ptr(rowptr), la._lib.ptr(colval), 2, la._lib.ptr(bg_rowptr), la._lib.ptr(bg_cols), ctypes.byref(h))
    assert rc == 1 and b"gathered row 5 of 2" in L.hpcla_last_error()
    assert L.hpcla_spgemm_symbolic(7, 1, la._lib.ptr(rowptr), la._lib.ptr(colval), 2, la._lib.ptr(bg_rowptr), la._lib.ptr(bg_cols), ctypes.byref(h)) == 1
    assert L.hpcla_spgemm_sizes(None, None, None, None) == 1 and L.hpcla_dtb_sizes(None, None, None, None) == 1
    a = np.zeros(4, dtype=np.int64)
    out = [np.zeros(2, dtype=np.int64) for _ in range(9)]
    bad = np.array([1, 4, 3], dtype=np.int64)  # decreasing
    ok = np.array([1, 2, 3], dtype=np.int64)
    rc = L.hpcla_repartition_plan(0, 2, la._lib.ptr(bad), la._lib.ptr(ok), *[la._lib.ptr(o) for o in out[:8]], la._lib.ptr(a), la._lib.ptr(out[8]))
    assert rc == 1 and b"non-decreasing" in L.hpcla_last_error()
    # a well-formed tiny product, through the plain C ABI
    colval[1] = 2
    assert L.hpcla_spgemm_symbolic(la._lib.I32, 1, la._lib.ptr(rowptr), la._lib.ptr(colval), 2, la._lib.ptr(bg_rowptr), la._lib.ptr(bg_cols), ctypes.byref(h)) == 0
    nnz, ncc, nt = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int64()
    assert L.hpcla_spgemm_sizes(h, ctypes.byref(nnz), ctypes.byref(ncc), ctypes.byref(nt)) == 0 and (nnz.value, ncc.value, nt.value) == (2, 2, 2)
    L.hpcla_spgemm_destroy(h)


def test_plan_import_round_trip_on_the_host():
    """hpcla_plan_import is a host function: a plan built by the reference's algorithm (the oracle), handed over in Ti width
    (Int32 and Int64), comes back out of hpcla_plan_get field by field exactly as it went in — the route of the Julia binding,
    checked here without a device (the device half is tests/test_gpu_round2.py::test_plan_import_route)."""
    import ctypes

    L = la._lib.lib()
    rng = np.random.default_rng(4)
    G = sp.random(500, 420, density=0.03, random_state=rng, format="csr")
    for P in (1, 2, 4):
        xp = np.concatenate([[1], np.sort(rng.integers(1, 421, size=P - 1)), [421]]).astype(np.int64)
        for itype, Ti in (("i32", np.int32), ("i64", np.int64)):
            plans = orc.vector_plans(orc.distribute(G, P, itype=itype), xp)
            for rank, o in enumerate(plans):
                sidx = [np.ascontiguousarray(a, dtype=Ti) for a in o.send_indices]
                rprm = [np.ascontiguousarray(a, dtype=Ti) for a in o.recv_perm]
                slen = np.array([len(a) for a in sidx], dtype=np.int64)
                rlen = np.array([len(a) for a in rprm], dtype=np.int64)
                sids = np.ascontiguousarray(o.send_rank_ids, dtype=np.int64)
                rids = np.ascontiguousarray(o.recv_rank_ids, dtype=np.int64)
                lsrc = np.ascontiguousarray(o.local_src_indices, dtype=Ti)
                ldst = np.ascontiguousarray(o.local_dst_indices, dtype=Ti)
                n_x = int(xp[rank + 1] - xp[rank])
                ph = ctypes.c_void_p()
                la._lib.check(L.hpcla_plan_import(rank, P, la._lib.itype_code(Ti), o.n_gathered, n_x, len(sids), la._lib.ptr(sids), la._lib.ptr(slen),
                                                  la._lib.ptr_array(sidx), len(rids), la._lib.ptr(rids), la._lib.ptr(rlen), la._lib.ptr_array(rprm), len(lsrc),
                                                  la._lib.ptr(lsrc), la._lib.ptr(ldst), ctypes.byref(ph)))
                plan = la.VectorPlan(ph.value, Ti, n_x)
                assert plan.send_rank_ids.tolist() == sids.tolist() and plan.recv_rank_ids.tolist() == rids.tolist()
                assert np.array_equal(plan.local_src_indices, lsrc) and np.array_equal(plan.local_dst_indices, ldst)
                assert all(np.array_equal(a, b) for a, b in zip(plan.send_indices, sidx)) and len(plan.send_indices) == len(sidx)
                assert all(np.array_equal(a, b) for a, b in zip(plan.recv_perm, rprm)) and len(plan.recv_perm) == len(rprm)
                assert plan.n_gathered == o.n_gathered and plan.local_src_indices.dtype == np.dtype(Ti)


def test_bench_describes_the_workloads_consistently():
    """bench.py's closed forms (rows, nnz per workload) equal what the generators produce, both arms print the same `config`,
    and the algorithmic bytes follow SURVEY §8(d)."""
    import importlib.util

    import hpcla_synth

    spec = importlib.util.spec_from_file_location("_bench", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    for name, n_gpus in [("poisson256", 1), ("poisson256", 8), ("laplace2d-1000", 1), ("stencil27-192", 1), ("poisson512-strong", 8), ("cg-512", 8)]:
        w = bench.workload_spec(name, n_gpus)
        n, nnz = bench.workload_counts(w)
        assert n == hpcla_synth.stencil_rows(w["kind"], w["grid"])
        small = tuple(min(g, 12) for g in w["grid"])
        ws = dict(w, grid=small)
        ns, nnzs = bench.workload_counts(ws)
        assert nnzs == int(hpcla_synth.lib().hpcla_synth_stencil_nnz(w["kind"], *small, 0, ns))
        cfg = bench.shared_config(w, n, nnz, n_gpus)
        assert set(cfg) == {"workload", "index_type", "op", "rows", "nnz", "l2"} and cfg["rows"] == n and cfg["nnz"] == nnz
    b, f = bench.algorithmic_bytes_flops(16777216, 16777216, 117047296, "f64", "i32", "mul")
    assert b == 1740111876 and f == 234094592
    assert bench.workload_counts(bench.workload_spec("poisson256", 8)) == (134217728, 937951232)  # config 5
