"""Multi-process NCCL worker for tests/test_gpu_nccl.py (one process per GPU, run under torch.distributed.run).
Checks the real grouped ncclSend/ncclRecv halo exchange against the oracle on every rank."""
import os
import sys

import numpy as np
import scipy.sparse as sp
import torch
import torch.distributed as dist

import hpcla_b200 as la
from oracle import oracle as orc

TOL = {np.dtype(np.float32): 1e-5, np.dtype(np.float64): 1e-12, np.dtype(np.complex128): 1e-12}


def relerr(y, ref):
    return np.linalg.norm(np.asarray(y) - np.asarray(ref)) / max(np.linalg.norm(np.asarray(ref)), 1e-300)


def main():
    os.environ["HPCLA_STAGE_CHUNK_KB"] = "64"  # many small blocks at test size
    local_rank = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    rank, P = dist.get_rank(), dist.get_world_size()
    comm = la.CommMPI()
    rng = np.random.default_rng(2024)  # identical on all ranks ("test matrices must be deterministic")
    S = la.synth
    # 1. stencils, all element types: A*x, mul!, transpose(A)*x, gathered
    for kind, N, T, Ti in [(1, 40, np.float64, np.int32), (1, 31, np.float32, np.int64), (2, 18, np.complex128, np.int32), (0, 300, np.float64, np.int64)]:
        b = la.backend_cuda_mpi(T, Ti, comm=comm, device=local_rank)
        n = S.stencil_rows(kind, N if kind else (N, N))
        grid = N if kind else (N, N)
        A = S.stencil_matrix(kind, grid, b)
        x = S.vector(n, b)
        y = A * x
        y2 = la.mul(la.HPCVector.zeros(b, n), A, x)
        yT = la.transpose(A) * x
        g = la.execute_plan(la.get_vector_plan(A, x), A, x)
        torch.cuda.synchronize()
        # staged multiply (host x in, host y out, pipelined): the halo waits for the last upload chunk
        xh = torch.empty(x.local_size, dtype=x.v.dtype, pin_memory=True).copy_(x.v)
        yh = torch.zeros(y.local_size, dtype=y.v.dtype, pin_memory=True)
        x3, y3 = la.HPCVector.zeros(b, n), la.HPCVector.zeros(b, n)
        for _ in range(2):
            la.mul_staged(y3, A, x3, xh, yh)
            torch.cuda.synchronize()
            assert torch.equal(yh, y.v.cpu()) and torch.equal(y3.v, y.v), ("staged", kind)
        rp, c, v = S.stencil_local(kind, grid, 0, n, T, Ti)
        G = sp.csr_matrix((v, c - 1, rp - 1), shape=(n, n))
        olocs = orc.distribute(G, P, itype="i32" if Ti == np.int32 else "i64")
        xh = S.vector_local(T, S.X_SEED, 0, n)
        ref = orc.matvec(olocs, xh)
        refT = orc.matvec(orc.transpose(olocs), xh)
        assert relerr(y.to_global(), ref) <= TOL[np.dtype(T)], ("A*x", kind, relerr(y.to_global(), ref))
        assert relerr(y2.to_global(), ref) <= TOL[np.dtype(T)]
        assert relerr(yT.to_global(), refT) <= TOL[np.dtype(T)]
        # sparse x dense: one exchange for all columns (row-major ghost rows straight from NCCL), 4 columns per pass
        Bg = np.stack([S.vector_local(T, S.X_SEED + j, 0, n) for j in range(6)], axis=1)
        C = A * la.HPCMatrix.from_global(Bg, b)
        Cg = C.to_global()
        assert relerr(Cg, orc.matmat(olocs, Bg)) <= TOL[np.dtype(T)], ("A*B", kind)
        assert np.array_equal(Cg[:, 0], y.to_global()), ("A*B column 0 vs A*x", kind)
        # device-side transpose over NCCL: this rank's block of A^T, array by array, against the oracle's TransposePlan
        At, ot = la.materialize_transpose(A), orc.transpose(olocs)[rank]
        assert np.array_equal(At.rowptr, ot.rowptr) and np.array_equal(At.colval, ot.colval) and np.array_equal(At.col_indices, ot.col_indices)
        assert np.array_equal(At.nzval_host(), ot.nzval), ("transpose values", kind)
        W = orc.PlanWorld(olocs, orc.uniform_partition(n, P))
        assert np.array_equal(g.cpu().numpy(), W.execute(orc.split_vector(xh, orc.uniform_partition(n, P)))[rank])
        W.close()
        info = la.spmv_info(A, x)
        assert info["x_in_place"] == 1 and info["sends_contiguous"] == 1 and info["boundary_tiles"] > 0
        if kind in (0, 1) and T == np.float64:
            assert np.array_equal(y.to_global(), ref)
    # 2. random ragged matrix: non-contiguous sends (pack kernel), foreign x partition, empty rows
    for T, Ti in [(np.float64, np.int32), (np.complex128, np.int64)]:
        b = la.backend_cuda_mpi(T, Ti, comm=comm, device=local_rank)
        m, n = 2000, 1500
        R = sp.random(m, n, density=0.01, random_state=np.random.default_rng(5), format="csr")
        R.data = np.random.default_rng(6).uniform(-1, 1, R.nnz)
        R = (R @ sp.diags((np.arange(n) % 3 != 1).astype(np.float64))).tocsr()  # every third column empty: the send runs have holes
        R.eliminate_zeros()
        R = R.astype(T)
        xh = np.random.default_rng(7).uniform(-1, 1, n).astype(T)
        xp = np.concatenate([[1], np.sort(np.random.default_rng(8).integers(1, n + 1, size=P - 1)) , [n + 1]]).astype(np.int64)
        A = la.HPCSparseMatrix.from_global(R, b)
        x = la.HPCVector.from_global(xh, b, partition=xp)
        y = A * x
        olocs = orc.distribute(R, P, itype="i32" if Ti == np.int32 else "i64")
        assert relerr(y.to_global(), orc.matvec(olocs, xh, xp)) <= TOL[np.dtype(T)]
        assert la.spmv_info(A, x)["sends_contiguous"] == 0
    # 2b. repartition over NCCL: contiguous ranges straight between device buffers
    for T in (np.float64, np.complex128):
        b = la.backend_cuda_mpi(T, np.int64, comm=comm, device=local_rank)
        n = 5000
        vh = np.random.default_rng(9).uniform(-1, 1, n).astype(T)
        old = np.concatenate([[1], np.sort(np.random.default_rng(10).integers(1, n + 2, size=P - 1)), [n + 1]]).astype(np.int64)
        new = np.concatenate([[1], np.sort(np.random.default_rng(11).integers(1, n + 2, size=P - 1)), [n + 1]]).astype(np.int64)
        xv = la.HPCVector.from_global(vh, b, partition=old)
        yv = la.repartition(xv, new)
        assert yv.partition.tolist() == new.tolist() and np.array_equal(yv.to_global(), vh)
        assert np.array_equal(yv.local_values(), orc.repartition(orc.split_vector(vh, old), old, new)[rank])
        zv = la.HPCVector.from_global(vh, b, partition=new)
        assert abs(la.dot(xv, zv) - np.vdot(vh, vh)) <= 1e-9 * n
    # 2c. sparse x sparse over NCCL: structure gathered on the host communicator, values exchanged between devices
    b = la.backend_cuda_mpi(np.float64, np.int32, comm=comm, device=local_rank)
    Ag = sp.random(400, 300, density=0.03, random_state=np.random.default_rng(12), format="csr")
    Bg2 = sp.random(300, 350, density=0.04, random_state=np.random.default_rng(13), format="csr")
    Cm = la.HPCSparseMatrix.from_global(Ag, b) * la.HPCSparseMatrix.from_global(Bg2, b)
    oC = orc.spgemm(orc.distribute(Ag, P, itype="i32"), orc.distribute(Bg2, P, itype="i32"), itype="i32")[rank]
    assert np.array_equal(Cm.rowptr, oC.rowptr) and np.array_equal(Cm.colval, oC.colval) and np.array_equal(Cm.col_indices, oC.col_indices)
    assert np.array_equal(Cm.nzval_host(), oC.nzval), "A*B values"
    # 3. reductions + CG over NCCL
    b = la.backend_cuda_mpi(np.float64, np.int32, comm=comm, device=local_rank)
    N = 20
    n = N**3
    A = S.stencil_matrix(1, N, b)
    x = S.vector(n, b)
    xh = S.vector_local(np.float64, S.X_SEED, 0, n)
    assert abs(la.dot(x, x) - np.dot(xh, xh)) <= 1e-10 * n and abs(la.norm(x) - np.linalg.norm(xh)) <= 1e-10
    rp, c, v = S.stencil_local(1, N, 0, n, np.float64, np.int32)
    G = sp.csr_matrix((v, c - 1, rp - 1), shape=(n, n))
    bh = G @ np.ones(n)
    sol, hist = la.cg(A, la.HPCVector.from_global(bh, b), 30)
    xo, ho = orc.cg(orc.distribute(G, 1, itype="i32"), bh, 30)
    assert relerr(sol.to_global(), xo) <= 1e-9 and np.allclose(hist, ho, rtol=1e-8)
    # 4. the route the Julia binding takes: a communicator created OUTSIDE the library (ncclCommInitRank through ctypes,
    #    as ext/HPCLinearAlgebraCUDAExt.jl:411-443 does) handed to hpcla_ctx_adopt_nccl, and a plan built by the reference's
    #    algorithm (the oracle, in Ti width) handed to hpcla_plan_import
    import ctypes

    nccl = ctypes.CDLL("libnccl.so.2")

    class NcclUniqueId(ctypes.Structure):
        _fields_ = [("internal", ctypes.c_byte * 128)]

    uid = NcclUniqueId()
    if rank == 0:
        assert nccl.ncclGetUniqueId(ctypes.byref(uid)) == 0
    raw = la.comm_bcast(comm, bytes(uid.internal), 0)
    ctypes.memmove(ctypes.byref(uid), raw, 128)
    ext_comm = ctypes.c_void_p()
    nccl.ncclCommInitRank.argtypes = [ctypes.c_void_p, ctypes.c_int, NcclUniqueId, ctypes.c_int]
    assert nccl.ncclCommInitRank(ctypes.byref(ext_comm), P, uid, rank) == 0
    comm2 = la.CommMPI(nccl_comm=ext_comm.value)
    for kind, N, T, Ti in [(1, 24, np.float64, np.int32), (2, 14, np.complex128, np.int64)]:
        b2 = la.backend_cuda_mpi(T, Ti, comm=comm2, device=local_rank)
        assert b2.ctx().world == "nccl"
        n = S.stencil_rows(kind, N)
        A = S.stencil_matrix(kind, N, b2)
        x = S.vector(n, b2)
        rp, c, v = S.stencil_local(kind, N, 0, n, T, Ti)
        G = sp.csr_matrix((v, c - 1, rp - 1), shape=(n, n))
        olocs = orc.distribute(G, P, itype="i32" if Ti == np.int32 else "i64")
        xp = orc.uniform_partition(n, P)
        oplan = orc.vector_plans(olocs, xp)[rank]
        plan = la.sparse.import_vector_plan(A, x, oplan.send_rank_ids, oplan.send_indices, oplan.recv_rank_ids, oplan.recv_perm,
                                            oplan.local_src_indices, oplan.local_dst_indices, oplan.n_gathered)
        assert la.get_vector_plan(A, x) is plan  # memoised under the reference's key
        xh = S.vector_local(T, S.X_SEED, 0, n)
        ref = orc.matvec(olocs, xh)
        y = A * x
        assert relerr(y.to_global(), ref) <= TOL[np.dtype(T)], ("adopted comm + imported plan", kind)
        g = la.execute_plan(plan, A, x)
        torch.cuda.synchronize()
        W = orc.PlanWorld(olocs, xp)
        assert np.array_equal(g.cpu().numpy(), W.execute(orc.split_vector(xh, xp))[rank])
        W.close()
        # the same multiply replayed from a CUDA graph (captured grouped ncclSend/ncclRecv)
        yg = la.HPCVector.zeros(b2, n)
        for _ in range(3):
            yg.v.zero_()
            la.mul_graph(yg, A, x)
            torch.cuda.synchronize()
            assert torch.equal(yg.v, y.v), ("graph replay", kind)
    # 5. irregular matrix (nnz-split kernel) with ghosts, and a stencil-like matrix whose general tiles border ghost columns
    b = la.backend_cuda_mpi(np.float32, np.int32, comm=comm, device=local_rank)
    n = 40000
    A = S.powerlaw_matrix(n, b, max_len=20000)
    x = S.vector(n, b)
    rp, c, v = S.powerlaw_local(n, S.POWERLAW_SEED, 20000, 0, n, np.float32, np.int32)
    G = sp.csr_matrix((v, c - 1, rp - 1), shape=(n, n))
    ref = orc.matvec(orc.distribute(G, P, itype="i32"), S.vector_local(np.float32, S.X_SEED, 0, n))
    assert relerr((A * x).to_global(), ref) <= 1e-5, "power-law rows over NCCL"
    # 6. direct halo between processes: `gathered` and the flags mapped with CUDA IPC, ghosts pushed by the copy engine
    b = la.backend_cuda_mpi(np.float64, np.int32, comm=comm, device=local_rank)
    N = 36
    n = N**3
    A = S.stencil_matrix(1, N, b)
    x = S.vector(n, b)
    y_nccl = (A * x).to_global()
    la.enable_direct_halo(A, x)
    for k in range(5):
        x.v.mul_(-1.0 if k else 1.0)
        yk = (A * x).to_global()
        assert np.array_equal(yk, (y_nccl if k % 2 == 0 else -y_nccl)), ("direct halo step", k)
    sol_d, hist_d = la.cg(A, A * la.HPCVector.from_global(np.ones(n), b), 10)
    A3 = S.stencil_matrix(1, N, b)  # the same system over NCCL
    sol_n, hist_n = la.cg(A3, A3 * la.HPCVector.from_global(np.ones(n), b), 10)
    assert np.array_equal(hist_d, hist_n), "CG over the direct halo == CG over NCCL"
    xh2, yh2 = la.host_buffer(b, x.local_size), la.host_buffer(b, x.local_size)
    xh2.array[:] = x.local_values()
    x4, y4 = la.HPCVector.zeros(b, n), la.HPCVector.zeros(b, n)
    for _ in range(2):
        la.mul_staged(y4, A, x4, xh2.array, yh2.array)
        torch.cuda.synchronize()
    assert np.array_equal(y4.to_global(), y_nccl), "staged multiply over the direct halo"
    dist.barrier()
    print("NCCL_OK", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
