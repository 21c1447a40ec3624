"""Kernel tuning harness (development tool): times the SpMV tile kernel variants / tile windows on one GPU.

    python tools/tune_spmv.py [--workload poisson256|stencil27|powerlaw|laplace2d] [--reps 100]
"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hpcla_b200 as la  # noqa: E402
from bench import algorithmic_bytes_flops  # noqa: E402


def clone_matrix(A):
    """A second HPCSparseMatrix over the SAME device arrays (fresh tile table / operator)."""
    return la.HPCSparseMatrix(A.structural_hash, A.row_partition, A.col_partition, A.col_indices, A.rowptr, A.colval, A.nzval, A.nrows_local,
                              A.ncols_compressed, A.rowptr_target, A.colval_target, A.backend)


def time_config(A, x, y, env, reps):
    for k, v in env.items():
        os.environ[k] = str(v)
    B = clone_matrix(A)
    la.mul(y, B, x)
    torch.cuda.synchronize()
    for _ in range(5):
        la.mul(y, B, x)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(3):
        e0.record()
        for _ in range(reps):
            la.mul(y, B, x)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / reps)
    info = la.spmv_info(B, x)
    res = y.v.clone()
    for k in env:
        os.environ.pop(k, None)
    return best, info, res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="poisson256")
    ap.add_argument("--reps", type=int, default=100)
    ap.add_argument("--windows", default="")
    ap.add_argument("--kinds", default="", help="comma list of HPCLA_SPMV_KIND values (general, rowwalk)")
    ap.add_argument("--lanes", default="", help="comma list of HPCLA_LANES values")
    ap.add_argument("--sweep", default="", help="comma list of lanes:window pairs")
    ap.add_argument("--no-direct", action="store_true", help="also time HPCLA_DIRECT=0 (the record-driven row walk)")
    ap.add_argument("--cusparse", action="store_true")
    ap.add_argument("--n", type=int, default=20_000_000, help="rows of the power-law matrix")
    args = ap.parse_args()
    S = la.synth
    if args.workload == "poisson256":
        T, Ti, tn, tin = np.float64, np.int32, "f64", "i32"
        b = la.backend_cuda_serial(T, Ti)
        A = S.stencil_matrix(1, 256, b)
    elif args.workload == "poisson256-i64":
        T, Ti, tn, tin = np.float64, np.int64, "f64", "i64"
        b = la.backend_cuda_serial(T, Ti)
        A = S.stencil_matrix(1, 256, b)
    elif args.workload == "stencil27":
        T, Ti, tn, tin = np.complex128, np.int32, "c128", "i32"
        b = la.backend_cuda_serial(T, Ti)
        A = S.stencil_matrix(2, 192, b)
    elif args.workload == "stencil27-f64":
        T, Ti, tn, tin = np.float64, np.int32, "f64", "i32"
        b = la.backend_cuda_serial(T, Ti)
        A = S.stencil_matrix(2, 192, b)
    elif args.workload == "laplace2d":
        T, Ti, tn, tin = np.float64, np.int64, "f64", "i64"
        b = la.backend_cuda_serial(T, Ti)
        A = S.stencil_matrix(0, (1000, 1000), b)
    elif args.workload == "powerlaw":
        T, Ti, tn, tin = np.float32, np.int32, "f32", "i32"
        b = la.backend_cuda_serial(T, Ti)
        A = S.powerlaw_matrix(args.n, b)
    else:
        raise SystemExit("unknown workload")
    n = A.shape[0]
    x = S.vector(n, b)
    y = la.HPCVector.zeros(b, n)
    bts, fl = algorithmic_bytes_flops(n, n, A.nnz_local, tn, tin, "mul")
    print(f"workload {args.workload}: n={n} nnz={A.nnz_local} bytes={bts/1e9:.3f} GB")
    configs = [{}]
    for k in [v for v in args.kinds.split(",") if v]:
        configs.append({"HPCLA_SPMV_KIND": k})
    for g in [int(v) for v in args.lanes.split(",") if v]:
        configs.append({"HPCLA_LANES": g})
    for spec in [v for v in args.sweep.split(",") if v]:  # lanes:window pairs
        g, w = spec.split(":")
        configs.append({"HPCLA_LANES": int(g), "HPCLA_TILE_WINDOW": int(w)})
    for w in [int(v) for v in args.windows.split(",") if v]:
        configs.append({"HPCLA_TILE_WINDOW": w})
    if args.no_direct:
        configs.append({"HPCLA_DIRECT": 0})

    ref = None
    for env in configs:
        ms, info, res = time_config(A, x, y, env, args.reps)
        if ref is None:
            ref = res
        err = float((res - ref).abs().max() / ref.abs().max())
        print(f"{str(env):50s} {ms*1e3:9.1f} us  {bts/ms/1e6:8.1f} GB/s  {fl/ms/1e6:8.1f} GFLOP/s  tiles={info['tiles']} G={info['lanes_per_row']} "
              f"win={info['tile_window']} rowwalk/general={info['rowwalk_tiles']}/{info['general_tiles']} maxrelerr_vs_first={err:.2e}", flush=True)
    if args.cusparse:
        # same-box comparator (not part of the reference): cuSPARSE CSR SpMV through torch
        crow = (A.rowptr_target - 1)
        col = (A.colval_target - 1)
        M = torch.sparse_csr_tensor(crow, col, A.nzval, size=(n, n))
        xv = x.v
        for _ in range(3):
            yy = M @ xv
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = 1e9
        for _ in range(3):
            e0.record()
            for _ in range(max(args.reps // 4, 5)):
                yy = M @ xv
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / max(args.reps // 4, 5))
        err = float((yy - ref).abs().max() / ref.abs().max())
        print(f"{'cuSPARSE (torch.sparse_csr @ x, allocates y)':50s} {best*1e3:9.1f} us  {bts/best/1e6:8.1f} GB/s  {fl/best/1e6:8.1f} GFLOP/s  maxrelerr_vs_first={err:.2e}", flush=True)


if __name__ == "__main__":
    main()
