"""2-process gloo worker for tests/test_host_logic.py::test_gloo_world_size_2 (run under torch.distributed.run)."""
import numpy as np
import scipy.sparse as sp
import torch.distributed as dist

import hpcla_b200 as la
from conftest import FIXTURES, fixture_matrix
from oracle import oracle as orc
from test_host_logic import _compare_matrix, _compare_plan


def main():
    dist.init_process_group("gloo")
    rank, P = dist.get_rank(), dist.get_world_size()
    comm = la.CommMPI()
    rng = np.random.default_rng(5)
    cases = [(fixture_matrix(f), f.get("row_partition"), np.complex128 if f["dtype"] == "c128" else np.float64) for f in FIXTURES]
    R = sp.random(300, 280, density=0.03, random_state=rng, format="csr")
    cases.append((R, None, np.float64))
    cases.append((la.synth.stencil_matrix(la.synth.POISSON3D_7PT, 6, la.backend_cpu_serial()) and None, None, None))
    for A_global, rp, T in cases:
        if A_global is None:
            continue
        for Ti in (np.int32, np.int64):
            b = la.backend_cpu_mpi(T, Ti, comm=comm)
            A_global = sp.csr_matrix(A_global).astype(T)
            A = la.HPCSparseMatrix.from_global(A_global, b, row_partition=rp)
            olocs = orc.distribute(A_global, P, row_partition=rp, itype="i32" if Ti == np.int32 else "i64")
            _compare_matrix(A, olocs[rank])
            x = la.HPCVector.from_global(np.ones(A_global.shape[1], dtype=T), b)
            plan = la.get_vector_plan(A, x)
            _compare_plan(plan, orc.vector_plans(olocs, orc.uniform_partition(A_global.shape[1], P))[rank])
            At = la.materialize_transpose(A)
            _compare_matrix(At, orc.transpose(olocs)[rank])
            assert np.array_equal(x.to_global(), np.ones(A_global.shape[1], dtype=T))
    # locally generated rows (HPCSparseMatrix_local route) agree with distributing the global matrix
    b = la.backend_cpu_mpi(np.float64, np.int32, comm=comm)
    A = la.synth.stencil_matrix(la.synth.POISSON3D_7PT, 6, b)
    rp_, c_, v_ = la.synth.stencil_local(la.synth.POISSON3D_7PT, 6, 0, 216, np.float64, np.int32)
    G = sp.csr_matrix((v_, c_ - 1, rp_ - 1), shape=(216, 216))
    _compare_matrix(A, orc.distribute(G, P, itype="i32")[rank])
    # the next rows over the real process group: MatrixPlan structure exchange (tags 1, 2) + symbolic product, repartition
    # plans, HPCMatrix_local's Allgather of row counts
    Bs = sp.random(280, 150, density=0.05, random_state=np.random.default_rng(6), format="csr")
    b = la.backend_cpu_mpi(np.float64, np.int64, comm=comm)
    Am, Bm = la.HPCSparseMatrix.from_global(R, b), la.HPCSparseMatrix.from_global(Bs, b)
    lA, lB = orc.distribute(R, P, itype="i64"), orc.distribute(Bs, P, itype="i64")
    mp = la.get_matrix_plan(Am, Bm)
    oC = orc.spgemm(lA, lB, itype="i64")[rank]
    bg = orc.gather_rows(lB, lA[rank].col_indices)
    assert np.array_equal(mp.bg_rowptr, bg[0]) and np.array_equal(mp.bg_cols, bg[1])
    assert np.array_equal(mp.rowptr, oC.rowptr) and np.array_equal(mp.colval, oC.colval) and np.array_equal(mp.col_indices, oC.col_indices)
    old, new = orc.uniform_partition(37, P), np.array([1, 30, 38], dtype=np.int64)
    rpl, ref = la.VectorRepartitionPlan(rank, P, old, new), orc.repartition_plan(rank, old, new)
    assert rpl.send_rank_ids == ref["send_rank_ids"] and rpl.recv_offsets == ref["recv_offsets"] and rpl.result_local_size == ref["result_local_size"]
    M = np.arange(20.0).reshape(10, 2)
    lo, hi = int(orc.uniform_partition(10, P)[rank]) - 1, int(orc.uniform_partition(10, P)[rank + 1]) - 1
    assert np.array_equal(la.HPCMatrix_local(M[lo:hi], b).to_global(), M)
    assert la.comm_allreduce(comm, rank + 1) == P * (P + 1) // 2
    dist.barrier()
    print("DIST_OK", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
