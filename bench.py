#!/usr/bin/env python
"""bench.py — distributed SpMV throughput on B200 (BASELINE.json metric: GFLOP/s and achieved HBM GB/s vs roofline).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload NAME]

One "step" = one `mul!(y, A, x)` (src/sparse.jl:2019-2037) on the whole distributed matrix.

Workload (BASELINE.json configs[1]): 3-D 7-point Poisson, Float64 values, Int32 indices.  N=1: the 256^3 grid
(16.7 M rows, 117 M nnz).  N>1: the rows per GPU stay at 256^3 (weak scaling): 512x256x256, 512x512x256 and at N=8 the
512^3 grid on which BASELINE.json states the 8-GPU efficiency target.  `--workload` selects the other configs.

Output: ONE JSON line on rank 0 (see the keys at the bottom).  `value` is device-timed with inputs resident in HBM
(`ms_per_step` = the K timed steps / K, max over ranks; `median_ms_per_step` from per-step events); `e2e` is the same
multiply with HOST (pinned, NUMA-local) x and y, copies inside the timed region, next to what the two copies alone cost
on this box; `roofline` relates the kernel to the measured HBM copy peak; `cpu_baseline` is the CPU restatement of the
reference's path (oracle/) on this box's cores (the reference itself is Julia+MPI, neither of which exists in this
image).  `check` says how the result was verified before anything was timed: rows of y on every rank — the first and the
last boundary plane, which read ghosts, plus random interior runs — against the oracle's row loop on regenerated inputs.
`--impl reference` times the CPU restatement alone on the FULL workload (every worker thread generates and first-touches
its own row block), with identical `config`.  It never loads the product library.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

NOMINAL_HBM_GBS = 8000.0  # the north-star's "~8 TB/s"; the measured copy peak comes from MEASURED_PEAKS.json
FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback


# ---------------------------------------------------------------------------------------------------------------
# workloads
# ---------------------------------------------------------------------------------------------------------------
def weak_grid(n_gpus: int):
    base = {1: (256, 256, 256), 2: (512, 256, 256), 4: (512, 512, 256), 8: (512, 512, 512)}
    return base.get(n_gpus, (256, 256, 256 * n_gpus))


def workload_spec(name: str, n_gpus: int) -> dict:
    if name == "poisson256":
        g = weak_grid(n_gpus)
        return dict(name=f"poisson3d_7pt {g[0]}x{g[1]}x{g[2]} (256^3 rows per GPU)", kind=1, grid=g, T="f64", Ti="i32", op="mul", scaling="weak")
    if name == "poisson256-strong":
        return dict(name="poisson3d_7pt 256x256x256", kind=1, grid=(256, 256, 256), T="f64", Ti="i32", op="mul", scaling="strong")
    if name == "poisson256-i64":
        return dict(name="poisson3d_7pt 256x256x256 Int64 indices", kind=1, grid=(256, 256, 256), T="f64", Ti="i64", op="mul", scaling="strong")
    if name == "poisson512-strong":
        return dict(name="poisson3d_7pt 512x512x512", kind=1, grid=(512, 512, 512), T="f64", Ti="i32", op="mul", scaling="strong")
    if name == "laplace2d-1000":
        return dict(name="laplace2d_5pt 1000x1000", kind=0, grid=(1000, 1000, 1), T="f64", Ti="i64", op="mul", scaling="strong")
    if name == "stencil27-192":
        return dict(name="stencil3d_27pt 192^3 ComplexF64 A*x", kind=2, grid=(192, 192, 192), T="c128", Ti="i32", op="mul", scaling="strong")
    if name == "stencil27-192-T":
        return dict(name="stencil3d_27pt 192^3 ComplexF64 transpose(A)*x", kind=2, grid=(192, 192, 192), T="c128", Ti="i32", op="transpose", scaling="strong")
    if name == "powerlaw-20m":
        return dict(name="powerlaw 20M x 20M Float32 Int32", kind=3, grid=(20_000_000, 1, 1), T="f32", Ti="i32", op="mul", scaling="strong")
    if name == "powerlaw-2m":
        return dict(name="powerlaw 2M x 2M Float32 Int32", kind=3, grid=(2_000_000, 1, 1), T="f32", Ti="i32", op="mul", scaling="strong")
    if name.startswith("poisson256-spmm"):  # A * B::HPCMatrix with k columns, e.g. poisson256-spmm8
        k = int(name[len("poisson256-spmm"):] or 8)
        g = weak_grid(n_gpus)
        return dict(name=f"poisson3d_7pt {g[0]}x{g[1]}x{g[2]} times a dense {k}-column HPCMatrix", kind=1, grid=g, T="f64", Ti="i32", op="spmm", scaling="weak", ncols=k)
    if name == "cg-512":
        g = (512, 512, 512) if n_gpus == 8 else weak_grid(n_gpus)
        if os.environ.get("HPCLA_BENCH_CG_GRID"):  # e.g. 512: the 8-GPU problem on fewer GPUs ("vs the 1-GPU run")
            g = (int(os.environ["HPCLA_BENCH_CG_GRID"]),) * 3
        return dict(name=f"CG on poisson3d_7pt {g[0]}x{g[1]}x{g[2]}", kind=1, grid=g, T="f64", Ti="i32", op="cg", scaling="weak")
    raise SystemExit(f"unknown workload {name!r}")


NP_T = {"f32": np.float32, "f64": np.float64, "c128": np.complex128}
NP_TI = {"i32": np.int32, "i64": np.int64}


def algorithmic_bytes_flops(n_rows, n_cols, nnz, T, Ti, op):
    """bytes = nnz*(sizeof T + sizeof Ti) + (n+1)*sizeof Ti + n_cols*sizeof T [x] + n_rows*sizeof T [y]."""
    sT, sI = np.dtype(NP_T[T]).itemsize, np.dtype(NP_TI[Ti]).itemsize
    b = nnz * (sT + sI) + (n_rows + 1) * sI + n_cols * sT + n_rows * sT
    f = (8 if T == "c128" else 2) * nnz
    if op == "cg":  # one CG iteration: SpMV + 12 n sizeof(T) of vector traffic, 2 nnz + 10 n flops
        b += 12 * n_rows * sT
        f += 10 * n_rows
    return b, f


def spmm_bytes_flops(n_rows, n_cols, nnz, T, Ti, k):
    """A read once, k columns of B read and k columns of C written."""
    sT, sI = np.dtype(NP_T[T]).itemsize, np.dtype(NP_TI[Ti]).itemsize
    return nnz * (sT + sI) + (n_rows + 1) * sI + k * (n_cols + n_rows) * sT, (8 if T == "c128" else 2) * nnz * k


# ---------------------------------------------------------------------------------------------------------------
# clocks (sampled DURING the timed region)
# ---------------------------------------------------------------------------------------------------------------
REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown",
           0x2: "applications_clocks_setting", 0x10: "sync_boost"}


class ClockSampler:
    def __init__(self, device_index: int, interval_s: float = 0.002):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        self.ok = False
        try:
            import pynvml
            import torch

            pynvml.nvmlInit()
            self.nv = pynvml
            try:
                uuid = str(torch.cuda.get_device_properties(device_index).uuid)
                self.h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
            except Exception:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.interval = interval_s
            self.ok = True
        except Exception as e:  # noqa: BLE001
            self.err = repr(e)

    def _loop(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.interval)

    def start(self):
        if self.ok:
            self._stop.clear()
            self._thread = threading.Thread(target=self._loop, daemon=True)
            self._thread.start()

    def stop(self):
        if self._thread is not None:
            self._stop.set()
            self._thread.join()
            self._thread = None

    def summary(self, region: str) -> dict:
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0, "region": region, "note": "nvml unavailable"}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples), "region": region}


# ---------------------------------------------------------------------------------------------------------------
# shared by both arms: what the workload is (identical `config` in both JSON lines)
# ---------------------------------------------------------------------------------------------------------------
def workload_counts(spec: dict):
    """(rows, nnz) of the workload from the generators' closed forms; None for nnz where only the
    generator knows (power-law)."""
    kind, (nx, ny, nz) = spec["kind"], spec["grid"]
    if kind == 0:
        return nx * ny, 5 * nx * ny - 2 * nx - 2 * ny
    if kind == 1:
        return nx * ny * nz, 7 * nx * ny * nz - 2 * (nx * ny + ny * nz + nx * nz)
    if kind == 2:
        return nx * ny * nz, (3 * nx - 2) * (3 * ny - 2) * (3 * nz - 2)
    return nx, None


def shared_config(spec: dict, n: int, nnz: int, n_gpus: int) -> dict:
    bytes_step, _ = algorithmic_bytes_flops(n, n, nnz, spec["T"], spec["Ti"], spec["op"])
    return {"workload": spec["name"], "index_type": spec["Ti"], "op": spec["op"], "rows": int(n), "nnz": int(nnz),
            "l2": "inputs exceed L2 (no flush needed)" if bytes_step / n_gpus > 2 * 126e6 else "inputs do NOT exceed L2"}


# ---------------------------------------------------------------------------------------------------------------
# CPU baseline: the oracle's restatement of the reference's distributed A*x, one worker thread per "MPI rank"
# ---------------------------------------------------------------------------------------------------------------
def cpu_baseline_run(spec: dict, reps: int, warmup: int, workers: int = 0, max_rows: int = 150_000_000):
    """The reference's CPU path (execute_plan! + the row loop, src/vectors.jl:394-463, src/sparse.jl:2055-2066) on the FULL
    workload: one worker thread per host core stands for one MPI rank; every worker generates, first-touches and owns
    its row block (orc_bench_synth), pinned to one core.  Grids above `max_rows` rows are cut to a slab of whole planes
    (said in the sample description).  Returns (median seconds per step, flops, bytes, workers, description, nnz)."""
    from oracle import oracle as orc

    # one worker per host core; BASELINE.json quotes config 1 (2-D Laplacian) on `mpiexec -n 4`: 4 workers there
    workers = workers or (4 if spec["kind"] == 0 else (os.cpu_count() or 1))
    kind, (nx, ny, nz), T, Ti, op = spec["kind"], spec["grid"], spec["T"], spec["Ti"], spec["op"]
    grid, desc = (nx, ny, nz), "full workload"
    n_rows = nx if kind == 3 else (nx * ny if kind == 0 else nx * ny * nz)
    if n_rows > max_rows and kind in (1, 2):
        planes = max(workers, max_rows // (nx * ny))
        grid, n_rows = (nx, ny, planes), nx * ny * planes
        desc = f"{nx}x{ny}x{planes} slab ({planes}/{nz} of the planes of the workload)"
    workers = max(1, min(workers, n_rows))
    inner = spec.get("ncols", 1) if op == "spmm" else 1  # the reference multiplies a dense B column by column (src/sparse.jl:2398-2403)
    times, nnz, _ = orc.bench_synth(kind, grid, NP_T[T], NP_TI[Ti], workers, warmup=warmup, reps=reps, op="cg" if op == "cg" else "mul", inner=inner)
    bts, fl = algorithmic_bytes_flops(n_rows, n_rows, nnz, T, Ti, op)
    if op == "spmm":
        bts, fl = spmm_bytes_flops(n_rows, n_rows, nnz, T, Ti, inner)
    desc += f"; {workers} worker threads, one row block each, generated and first-touched by its worker, pinned to one core; median of {reps} steps"
    if op == "transpose":
        desc += "; transpose(A)*x is A^T*x on the cached materialisation (src/sparse.jl:2375-2379): same stored entries, timed as a multiply with A"
    return float(np.median(times)), fl, bts, workers, desc, nnz


def run_reference_arm(args, spec):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t, fl, bts, workers, desc, nnz = cpu_baseline_run(spec, reps=max(args.steps, 1), warmup=max(args.warmup, 1), workers=args.cpu_workers)
    n, nnz_full = workload_counts(spec)
    val = fl / t / 1e9
    line = {
        "impl": "reference", "metric": "spmv_gflops", "value": val, "unit": "GFLOP/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": spec["scaling"], "vs_baseline": None,
        "dtype": spec["T"], "data": "synthetic",
        "config": shared_config(spec, n, nnz_full if nnz_full is not None else nnz, args.gpus),
        "cpu_baseline": {"value": val, "unit": "GFLOP/s", "cores": workers, "kind": "port", "sample": desc,
                         "achieved_gbs": bts / t / 1e9,
                         "note": "CPU restatement (oracle/) of the reference's execute_plan! + row-serial CSR loop, one worker thread per core "
                                 "standing for one MPI rank; the reference itself (Julia + MPI) cannot run in this image"},
        "e2e": {"value": val, "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------
# the B200 arm
# ---------------------------------------------------------------------------------------------------------------
def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def oracle_check(spec, rank, row_partition, y_local, x_seed, transpose=False, n_random=8):
    """Rows of this rank's y against the oracle (oracle/rowcheck.py): the first and the last plane of the rank's block —
    the rows that read ghosts, every ghost value distinct — plus random interior runs.  Returns (rows, max relerr)."""
    from oracle import rowcheck

    b, e = int(row_partition[rank]) - 1, int(row_partition[rank + 1]) - 1
    rows, worst, _, _ = rowcheck.check_rows(spec["kind"], spec["grid"], b, e, y_local, NP_T[spec["T"]], NP_TI[spec["Ti"]], x_seed=x_seed,
                                            transpose=transpose, seed=1000 + rank, n_random=n_random)
    return rows, worst


def run_b200_arm(args, spec):
    if args.timeline:
        os.environ["HPCLA_TIMELINE"] = "1"
        args.graph = False  # (timing events are not part of a captured graph)
    if args.halo == "direct":
        args.graph = False  # (the direct halo counts steps in its flags: no replay)
    if args.graph:
        os.environ["HPCLA_CG_GRAPH"] = "1"
    import torch

    import hpcla_b200 as la

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit(f"--gpus {args.gpus} needs `python -m torch.distributed.run --nproc-per-node {args.gpus} bench.py ...`")
        args.gpus = world
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the B200 backend has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    T, Ti = NP_T[spec["T"]], NP_TI[spec["Ti"]]
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        backend = la.backend_cuda_mpi(T, Ti, comm=la.CommMPI(), device=local_rank)
    else:
        dist = None
        backend = la.backend_cuda_serial(T, Ti, device=local_rank)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(v: float) -> float:
        if dist is None:
            return float(v)
        t = torch.tensor([v], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- inputs (synthetic, generated per rank: nothing global is materialised) --------------------------------
    t_setup = time.time()
    kind, grid = spec["kind"], spec["grid"]
    if kind == 3:
        n = grid[0]
        A = la.synth.powerlaw_matrix(n, backend)
    else:
        n = la.synth.stencil_rows(kind, grid)
        A = la.synth.stencil_matrix(kind, grid, backend)
    x = la.synth.vector(n, backend)
    op = spec["op"]
    Aop = la.materialize_transpose(A) if op == "transpose" else A
    y = la.HPCVector.zeros(backend, n, partition=Aop.row_partition)
    nnz_local = Aop.nnz_local
    nnz = int(la.comm_allreduce(backend.comm, nnz_local, "+"))
    bytes_step, flops_step = algorithmic_bytes_flops(n, n, nnz, spec["T"], spec["Ti"], op)
    if op == "spmm":
        bytes_step, flops_step = spmm_bytes_flops(n, n, nnz, spec["T"], spec["Ti"], spec["ncols"])
    plan = la.get_vector_plan(Aop, x)
    L = la._lib.lib()
    opnd = la.sparse._bound_op(Aop, plan, x)
    if args.halo == "direct" and world > 1:
        la.enable_direct_halo(Aop, x)  # ghosts pushed into the neighbours' buffers by the copy engine + flags (NCCL is the default)
    info = la.spmv_info(Aop, x)
    setup_s = time.time() - t_setup

    warmup = max(args.warmup, 3)
    steps = max(args.steps, 1)
    part = Aop.row_partition
    tol = {"f32": 1e-5, "f64": 1e-12, "c128": 1e-12}[spec["T"]]

    # ---- correctness guard, before anything is timed: a number from a wrong kernel is worthless -------------------
    # every rank checks rows of its own block against the oracle on the REAL synthetic x (all values distinct)
    checks = []

    def guard(name, y_local, x_seed, transpose=False, n_random=8):
        rows, worst = oracle_check(spec, rank, part, y_local, x_seed, transpose=transpose, n_random=n_random)
        worst = allmax(worst)
        if not worst <= tol:
            raise SystemExit(f"bench.py: {name} disagrees with the oracle (max normwise relative error {worst:.3e} > {tol:g}); refusing to time a wrong result")
        checks.append(f"{name}: {rows} rows per rank (first + last plane of the rank's block + random runs) vs oracle, max relerr {worst:.2e}")

    cg_x = cg_work = None
    if op == "cg":
        ones = la.HPCVector.from_local(np.ones(x.local_size, dtype=T), backend)
        bvec = la.matvec(A, ones)  # b = A*1
        # q = A*b on the real path (ghosts = row sums of the neighbours' planes, regenerated by the checker from the
        # generator: b is not the synthetic x, so the check is on A*x with the synthetic x, plus the CG recurrence below)
        guard("A*x", la.matvec(A, x).local_values(), la.synth.X_SEED)
        q = la.matvec(A, bvec)
        bb, bq = la.dot(bvec, bvec), la.dot(bvec, q)
        alpha = bb / bq
        r1 = bvec.copy()
        la.axpby(-alpha, q, 1.0, r1)
        rr1 = la.dot(r1, r1)
        cg_x = la.HPCVector.zeros(backend, n, partition=A.row_partition)
        cg_work = torch.empty(3 * x.local_size, dtype=x.v.dtype, device=x.v.device)
        _, hist = la.cg(A, bvec, 2, x=cg_x, work=cg_work)
        if not abs(hist[0] - rr1) <= 1e-9 * abs(rr1):
            raise SystemExit(f"bench.py: CG's first residual {hist[0]!r} differs from the composition of A*b, dot and axpby {rr1!r}")
        checks.append(f"cg: rr after iteration 1 == ||b - (b.b / b.Ab) A b||^2 composed from the checked multiply, dot and axpby (rel {abs(hist[0] - rr1) / abs(rr1):.1e})")
        del ones, q, r1

        def step():  # warm-up only; the timed region is ONE hpcla_cg call of `steps` iterations (a step = an iteration)
            la.cg(A, bvec, steps, x=cg_x, work=cg_work)
    elif op == "spmm":
        k = spec["ncols"]
        Bm = la.HPCMatrix.from_local(torch.stack([la.synth.vector(n, backend, seed=la.synth.X_SEED + j).v for j in range(k)]).T, backend)

        def step():
            step.C = la.spmm(A, Bm)

        step()
        for j in sorted({0, k - 1}):
            guard(f"(A*B)[:, {j}]", step.C.A[:, j].contiguous().cpu().numpy(), la.synth.X_SEED + j, n_random=4)
    else:
        if args.graph:
            def step():
                la.mul_graph(y, Aop, x)
        else:
            def step():
                la.mul(y, Aop, x)

        y.v.fill_(float("nan"))
        step()
        guard("transpose(A)*x" if op == "transpose" else "A*x", y.local_values(), la.synth.X_SEED, transpose=(op == "transpose"))
    check = "; ".join(checks)

    sampler = ClockSampler(local_rank)
    for _ in range(warmup if op != "cg" else 1):
        step()
    barrier()

    # ---- timed region: exactly `steps` steps, device-timed, max over ranks ---------------------------------------
    launches0 = int(L.hpcla_spmv_launch_count(opnd))
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    sampler.start()
    barrier()
    evs[0].record()
    if op == "cg":
        la.cg(A, bvec, steps, x=cg_x, work=cg_work)  # x0 = 0, r = p = b (one pass), then `steps` iterations, all scalars on the device
        evs[-1].record()
    else:
        for i in range(steps):
            step()
            evs[i + 1].record()
    barrier()
    sampler.stop()
    elapsed_ms = allmax(evs[0].elapsed_time(evs[-1]))
    median_ms = None
    if op != "cg":
        per_step = np.array([evs[i].elapsed_time(evs[i + 1]) for i in range(steps)])
        median_ms = allmax(float(np.median(per_step)))
    launches = int(L.hpcla_spmv_launch_count(opnd)) - launches0
    region = "timed"
    n_samples = len(sampler.samples)
    if dist is not None:
        lt = torch.tensor([launches], device="cuda", dtype=torch.int64)
        dist.all_reduce(lt, op=dist.ReduceOp.SUM)
        launches = int(lt.item())
        ns = torch.tensor([n_samples], device="cuda", dtype=torch.int64)
        dist.all_reduce(ns, op=dist.ReduceOp.MIN)
        n_samples = int(ns.item())
    if n_samples < 8 and op != "cg":
        # short timed region: keep the same step running (the SAME number of steps on every rank: the multiply is
        # collective) so that the clocks are sampled under load
        probe_steps = int(min(4000, max(20, 400.0 / max(elapsed_ms / steps, 1e-3))))
        sampler.start()
        for i in range(probe_steps):
            step()
            if i % 20 == 19:
                torch.cuda.synchronize()
        torch.cuda.synchronize()
        sampler.stop()
        region = f"timed + {probe_steps} more of the same step"
    ms_per_step = elapsed_ms / steps
    gflops = flops_step / (ms_per_step * 1e-3) / 1e9
    gbs = bytes_step / (ms_per_step * 1e-3) / 1e9
    timeline = None
    if args.timeline and op in ("mul", "transpose") and not args.graph:
        for _ in range(3):  # steady state, ranks aligned: the timeline is of the last of three back-to-back multiplies
            barrier()
            la.mul(y, Aop, x)
            la.mul(y, Aop, x)
            la.mul(y, Aop, x)
        tl = la.spmv_timeline(Aop, x)
        rows = [tl] if dist is None else [None] * world
        if dist is not None:
            dist.all_gather_object(rows, tl)
        timeline = {"per_rank_ms_from_x_ready": rows}

    # ---- end to end: host (pinned) x in, host y out, every step ---------------------------------------------------
    e2e = None
    if op not in ("cg", "spmm") and not args.no_e2e:
        # pinned host buffers allocated and first-touched next to this rank's GPU (hpcla_host_alloc)
        hx, hy = la.host_buffer(backend, x.local_size), la.host_buffer(backend, y.local_size)
        hx.array[:] = x.local_values()
        xh, yh = torch.from_numpy(hx.array), torch.from_numpy(hy.array)
        if T == np.complex128:
            xh, yh = xh.view(torch.complex128), yh.view(torch.complex128)

        def e2e_serial_step():  # the three calls a user of the device API would make, back to back on one stream
            x.v.copy_(xh, non_blocking=True)
            la.mul(y, Aop, x)
            yh.copy_(y.v, non_blocking=True)

        def e2e_step():  # the staged entry point: upload, multiply and download pipelined block by block
            la.mul_staged(y, Aop, x, hx.array, hy.array)

        s_up, s_down = torch.cuda.Stream(), torch.cuda.Stream()

        def copies_only_step():  # what the two PCIe directions alone cost on this box: the ceiling of any staged multiply
            cur = torch.cuda.current_stream()
            s_up.wait_stream(cur)
            s_down.wait_stream(cur)
            with torch.cuda.stream(s_up):
                x.v.copy_(xh, non_blocking=True)
            with torch.cuda.stream(s_down):
                yh.copy_(y.v, non_blocking=True)
            cur.wait_stream(s_up)
            cur.wait_stream(s_down)

        e2e_steps = max(1, min(steps, 50))

        def time_e2e(fn):
            for _ in range(3):
                fn()
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(e2e_steps):
                fn()
            e1.record()
            barrier()
            return allmax(e0.elapsed_time(e1)) / e2e_steps

        hy.array[:] = 0
        e2e_step()
        torch.cuda.synchronize()
        if not np.array_equal(hy.array, y.local_values()):
            raise SystemExit("bench.py: the staged multiply did not deliver y to the host buffer")
        rows, worst = oracle_check(spec, rank, part, hy.array, la.synth.X_SEED, transpose=(op == "transpose"), n_random=2)
        if not allmax(worst) <= tol:
            raise SystemExit("bench.py: the staged multiply's host result disagrees with the oracle")
        copies_ms = time_e2e(copies_only_step)
        serial_ms = time_e2e(e2e_serial_step)
        staged_ms = time_e2e(e2e_step)
        hb, db = int(x.local_size * x.v.element_size()), int(y.local_size * y.v.element_size())
        e2e = {"value": flops_step / (staged_ms * 1e-3) / 1e9, "unit": "GFLOP/s", "h2d_bytes_per_step": hb * world,
               "d2h_bytes_per_step": db * world, "steps": e2e_steps, "ms_per_step": staged_ms,
               "path": "hpcla_b200.mul_staged(y, A, x, x_host, y_host) = hpcla_spmv_run_staged: pinned host x -> x.v, y.v = A*x, y.v -> pinned host y, "
                       "pipelined over row blocks inside the timed region; A resident; host buffers from hpcla_host_alloc (first-touched on the GPU's NUMA node)",
               "serial_copies_ms_per_step": serial_ms,
               "serial_copies_value": flops_step / (serial_ms * 1e-3) / 1e9,
               "copies_only_ms_per_step": copies_ms,
               "copies_only_gbs_per_direction_per_gpu": hb / (copies_ms * 1e-3) / 1e9,
               "frac_of_copy_ceiling": copies_ms / staged_ms,
               "host_numa_node": hx.numa_node,
               "note": "copies_only = the same H2D and D2H copies issued concurrently with no multiply, all ranks at once: the ceiling the link/host allows; "
                       "frac_of_copy_ceiling = copies_only / staged"}
        del xh, yh
        hx.close(), hy.close()

    peak, peak_src = measured_peak()
    per_gpu_gbs = gbs / world
    traffic = None  # ncu DRAM bytes per launch of the dominant kernel (one --set full capture, profiles/traffic.json)
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            entry = json.load(f).get(args.workload)
        if entry and world == 1:
            traffic = int(entry["bytes"])
    except Exception:
        traffic = None
    if info["flat_chunks"] > 0:
        kernel = "spmv_flat_kernel"
    elif info["compact_tiles"] * 2 >= info["tiles"]:
        kernel = "spmv_cwalk_kernel"
    elif info["rowwalk_tiles"] >= info["general_tiles"]:
        kernel = "spmv_rowwalk_kernel"
    else:
        kernel = "spmv_tile_kernel"
    roofline = {"bound": "hbm", "achieved": per_gpu_gbs, "peak": peak, "unit": "GB/s", "frac": per_gpu_gbs / peak, "traffic": traffic,
                "peak_source": peak_src, "frac_of_nominal_8TBs": per_gpu_gbs / NOMINAL_HBM_GBS,
                "kernel": kernel + ("" if world == 1 else " (interior + boundary launches; step time includes the NCCL halo)"),
                "algorithmic_bytes_per_step": bytes_step, "flops_per_step": flops_step,
                "note": "achieved = algorithmic bytes (SURVEY §8d formula: values + indices in the reference's width + row pointers + x + y) / time; "
                        "the compact row walk moves 16-bit positions instead of the column indices, so its DRAM traffic is below the algorithmic bytes"}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        t, fl, bts, workers, desc, _ = cpu_baseline_run(spec, reps=7, warmup=1, workers=args.cpu_workers)
        cpu = {"value": fl / t / 1e9, "unit": "GFLOP/s", "cores": workers, "kind": "port", "sample": desc, "achieved_gbs": bts / t / 1e9}

    if rank == 0:
        line = {
            "metric": "spmv_gflops", "value": gflops, "unit": "GFLOP/s", "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": ms_per_step, "median_ms_per_step": median_ms,
            "value_at_median": (flops_step / (median_ms * 1e-3) / 1e9) if median_ms else None,
            "higher_is_better": True, "scaling": spec["scaling"], "vs_baseline": None, "dtype": spec["T"],
            "data": "synthetic",
            "config": shared_config(spec, n, nnz, world),
            "check": check,
            "detail": {"tiles": info["tiles"], "interior_tiles": info["interior_tiles"], "boundary_tiles": info["boundary_tiles"],
                       "compact_tiles": info["compact_tiles"], "flat_chunks": info["flat_chunks"], "x_in_place": info["x_in_place"],
                       "lanes_per_row": info["lanes_per_row"], "tile_window": info["tile_window"], "rowwalk_tiles": info["rowwalk_tiles"],
                       "general_tiles": info["general_tiles"], "long_rows": info["long_rows"], "setup_s": round(setup_s, 2),
                       "cuda_graph": bool(args.graph), "halo": args.halo if world > 1 else "none", "timeline": timeline},
            "achieved_gbs": gbs, "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches,
            "clocks": sampler.summary(region),
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="poisson256")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-in / host-out leg (experiments only: a line without e2e is not a bench line)")
    ap.add_argument("--halo", default="nccl", choices=["nccl", "direct"], help="halo exchange: grouped ncclSend/ncclRecv (default) or the direct peer push")
    ap.add_argument("--graph", dest="graph", action="store_true", default=False,
                    help="replay each multiply (or the whole CG loop) from a CUDA graph: hpcla_b200.mul_graph / HPCLA_CG_GRAPH=1")
    ap.add_argument("--no-graph", dest="graph", action="store_false", help="plain calls: hpcla_b200.mul, one driver call per launch / event (default)")
    ap.add_argument("--timeline", action="store_true", help="record the per-rank timeline of one multiply (HPCLA_TIMELINE=1) into detail.timeline")
    ap.add_argument("--cpu-workers", type=int, default=0, help="worker threads of the CPU arm (0 = one per host core; 4 for the 2-D Laplacian, as BASELINE.json)")
    args = ap.parse_args()
    spec = workload_spec(args.workload, args.gpus)
    if args.impl == "reference":
        run_reference_arm(args, spec)
    else:
        run_b200_arm(args, spec)


if __name__ == "__main__":
    main()
