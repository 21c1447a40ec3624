"""Measurements of the SURVEY §8(f) rows built so far (development tool; prints one JSON line per row on rank 0).

    python tools/bench_next_rows.py                       # 1 GPU
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/bench_next_rows.py

  f.1 sparse x dense: hpcla_spmm_run against the reference's column loop (ncols SpMVs), Poisson 256^3 per GPU
  f.2 repartition:    uniform -> shifted partition of a 256^3-per-GPU Float64 vector (bytes that change owner / time)
  f.3 transpose:      hpcla_transpose_device against the host builder, 27-point 128^3 ComplexF64
  f.4 sparse x sparse: A*A for the 7-point Poisson matrix, 96^3 rows per GPU: first call (plan + symbolic) and numeric phase
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hpcla_b200 as la  # noqa: E402


def timed(fn, reps, sync):
    fn()
    sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    sync()
    return e0.elapsed_time(e1) / reps


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        mk = lambda T, Ti: la.backend_cuda_mpi(T, Ti, comm=la.CommMPI(), device=local_rank)  # noqa: E731
    else:
        mk = lambda T, Ti: la.backend_cuda_serial(T, Ti, device=local_rank)  # noqa: E731

    def sync():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    def emit(**kw):
        if rank == 0:
            print(json.dumps(dict(n_gpus=world, **kw)), flush=True)

    S = la.synth
    # ---- f.1 -----------------------------------------------------------------------------------------------------
    b = mk(np.float64, np.int32)
    grid = {1: (256, 256, 256), 2: (512, 256, 256), 4: (512, 512, 256), 8: (512, 512, 512)}.get(world, (256, 256, 256 * world))
    n = grid[0] * grid[1] * grid[2]
    A = S.stencil_matrix(1, grid, b)
    for k in (4, 8, 16):
        B = la.HPCMatrix.from_local(torch.stack([S.vector(n, b, seed=S.X_SEED + j).v for j in range(k)]).T, b)
        cols = [B.column(j) for j in range(k)]
        t_mm = timed(lambda: la.spmm(A, B), 10, sync)
        t_loop = timed(lambda: [la.matvec(A, c) for c in cols], 5, sync)
        emit(row="f.1 sparse x dense", grid=grid, ncols=k, spmm_ms=t_mm, column_loop_ms=t_loop, speedup=t_loop / t_mm)
    del A, B, cols
    # ---- f.2 -----------------------------------------------------------------------------------------------------
    x = S.vector(n, b)
    old = x.partition
    shift = n // (4 * world)
    new = old.copy()
    new[1:-1] += shift  # every interior boundary moves by a quarter of a slice
    moved = int(sum(max(0, min(old[r + 1], new[r + 1]) - max(old[r], new[r])) for r in range(world)))
    moved = n - moved  # elements that change owner
    la.repartition(x, new)
    t_rep = timed(lambda: la.repartition(x, new), 20, sync) if world > 1 else timed(lambda: la.repartition(x, new), 20, sync)
    emit(row="f.2 repartition", n=n, elements_changing_owner=moved, ms=t_rep, moved_gbs=(moved * 8 / 1e9) / (t_rep * 1e-3) if t_rep > 0 else None,
         note="uniform -> every interior boundary shifted by a quarter slice; Float64")
    # ---- f.3 -----------------------------------------------------------------------------------------------------
    bc = mk(np.complex128, np.int32)
    N = 128
    Ac = S.stencil_matrix(2, N, bc)
    out = {}
    for mode in ("device", "host", "device"):  # the first device build also pays NCCL's lazy connection set-up
        os.environ["HPCLA_TRANSPOSE"] = mode
        Ac.cached_transpose = None
        sync()
        t0 = time.time()
        Y = la.materialize_transpose(Ac)
        sync()
        out[mode if mode not in out else mode + "_again"] = time.time() - t0
        del Y
    os.environ.pop("HPCLA_TRANSPOSE", None)
    emit(row="f.3 transpose materialisation", grid=(N, N, N), nnz_local=Ac.nnz_local, device_first_s=out["device"], device_s=out["device_again"], host_s=out["host"], speedup=out["host"] / out["device_again"],
         note="wall clock including the read-back of rowptr/colval that the Python mirror keeps on the host")
    # ---- f.4 -----------------------------------------------------------------------------------------------------
    Ns = 96
    As = S.stencil_matrix(1, (Ns, Ns, Ns * world), b)
    la.clear_plan_cache()
    sync()
    t0 = time.time()
    C = As * As  # MatrixPlan + symbolic product (host, memoised) + first numeric pass
    sync()
    t_first = time.time() - t0
    plan = la.get_matrix_plan(As, As)
    t_num = timed(lambda: la.spgemm(As, As), 10, sync)
    cpu_s = None
    if world == 1:
        import scipy.sparse as sp

        rp, c, v = S.stencil_local(1, (Ns, Ns, Ns), 0, Ns**3, np.float64, np.int32)
        G = sp.csr_matrix((v, c - 1, rp - 1), shape=(Ns**3, Ns**3))
        t0 = time.time()
        G @ G
        cpu_s = time.time() - t0
    emit(row="f.4 sparse x sparse (A*A, 7-point Poisson)", grid=(Ns, Ns, Ns * world), nnz_c_local=plan.nnz, terms_local=plan.nterms, first_call_s=t_first,
         numeric_ms=t_num, host_scipy_csr_matmul_s=cpu_s,
         note="first call = structure gather + symbolic product on the host (memoised); later calls = value exchange + one kernel; scipy = one-core host product, for scale")
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
