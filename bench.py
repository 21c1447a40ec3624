#!/usr/bin/env python
"""bench.py — distributed SpMV throughput on B200 (BASELINE.json metric: GFLOP/s and achieved HBM GB/s vs roofline).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload NAME]

One "step" = one `mul!(y, A, x)` (src/sparse.jl:2019-2037) on the whole distributed matrix.

Workload (BASELINE.json configs[1]): 3-D 7-point Poisson, Float64 values, Int32 indices.  N=1: the 256^3 grid
(16.7 M rows, 117 M nnz).  N>1: the rows per GPU stay at 256^3 (weak scaling): 512x256x256, 512x512x256 and at N=8 the
512^3 grid on which BASELINE.json states the 8-GPU efficiency target.  `--workload` selects the other configs.

Output: ONE JSON line on rank 0 (see the keys at the bottom).  `value` is device-timed with inputs resident in HBM;
`e2e` is the same multiply with HOST (pinned) x and y, copies inside the timed region; `roofline` relates the kernel to
the measured HBM copy peak; `cpu_baseline` is the CPU restatement of the reference's path (oracle/) on this box's
cores (the reference itself is Julia+MPI, neither of which exists in this image).
`--impl reference` times that CPU restatement alone, with the same metric and config keys.
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
# CPU baseline: the oracle's restatement of the reference's distributed A*x, one worker thread per "MPI rank"
# ---------------------------------------------------------------------------------------------------------------
def cpu_baseline_run(spec: dict, reps: int, warmup: int, budget_rows: int = 2_200_000, workers: int = 0):
    """Bounded sample: the same stencil on a slab with the workload's plane size but fewer planes (about `budget_rows`
    rows), one worker per host core.  Returns (median seconds, flops, bytes, workers, sample description)."""
    import hpcla_b200 as la
    from oracle import oracle as orc

    # one worker per host core; BASELINE.json quotes config 1 (2-D Laplacian) on `mpiexec -n 4`: 4 workers there
    workers = workers or (4 if spec["kind"] == 0 else (os.cpu_count() or 1))
    kind, (nx, ny, nz), T, Ti = spec["kind"], spec["grid"], spec["T"], spec["Ti"]
    if kind == 3:
        n = min(nx, budget_rows)
        grid, n_rows = (n, 1, 1), n
        desc = f"powerlaw {n} x {n} (same generator, {n}/{nx} of the rows)"
    elif kind == 0:
        grid, n_rows = (nx, ny, 1), nx * ny
        desc = f"laplace2d_5pt {nx}x{ny} (full)"
    else:
        planes = max(workers, min(nz, max(1, budget_rows // (nx * ny))))
        planes = min(planes, nz)
        grid, n_rows = (nx, ny, planes), nx * ny * planes
        desc = f"{'poisson3d_7pt' if kind == 1 else 'stencil3d_27pt'} {nx}x{ny}x{planes} slab ({planes}/{nz} of the planes of the workload)"
    workers = max(1, min(workers, n_rows))
    part = orc.uniform_partition(n_rows, workers)
    locs, xs = [], []
    for r in range(workers):
        b, e = int(part[r]) - 1, int(part[r + 1]) - 1
        if kind == 3:
            rowptr, cols, vals = la.synth.powerlaw_local(grid[0], 0xC4, 1_000_000, b, e, NP_T[T], NP_TI[Ti])
        else:
            rowptr, cols, vals = la.synth.stencil_local(kind, grid, b, e, NP_T[T], NP_TI[Ti])
        locs.append(orc.local_matrix(r, rowptr.astype(np.int64), cols.astype(np.int64), vals, part, part, itype=Ti))
        xs.append(la.synth.vector_local(NP_T[T], la.synth.X_SEED, b, e))
    nnz = sum(int(m.rowptr[-1]) - 1 for m in locs)
    times, _ = orc.bench_spmv(locs, xs, part, warmup=warmup, reps=reps)
    bts, fl = algorithmic_bytes_flops(n_rows, n_rows, nnz, T, Ti, "mul")
    return float(np.median(times)), fl, bts, workers, desc


def run_reference_arm(args, spec):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t, fl, bts, workers, desc = cpu_baseline_run(spec, reps=max(args.steps, 1), warmup=max(args.warmup, 1), workers=args.cpu_workers)
    val = fl / t / 1e9
    line = {
        "impl": "reference", "metric": "spmv_gflops", "value": val, "unit": "GFLOP/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": spec["scaling"], "vs_baseline": None,
        "dtype": spec["T"], "data": "synthetic",
        "config": {"workload": spec["name"], "index_type": spec["Ti"], "op": spec["op"]},
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


def run_b200_arm(args, spec):
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
    info = la.spmv_info(Aop, x)
    setup_s = time.time() - t_setup

    warmup = max(args.warmup, 3)
    steps = max(args.steps, 1)
    cg_iters_per_step = 1

    if op == "cg":
        bvec = la.matvec(A, la.HPCVector.from_local(np.ones(x.local_size, dtype=T), backend))  # b = A*1

        def step():  # warm-up only; the timed region is ONE hpcla_cg call of `steps` iterations (a step = an iteration)
            la.cg(A, bvec, cg_iters_per_step)
    elif op == "spmm":
        k = spec["ncols"]
        Bm = la.HPCMatrix.from_local(torch.stack([la.synth.vector(n, backend, seed=la.synth.X_SEED + j).v for j in range(k)]).T, backend)
        cols_ref = (A * Bm.column(k - 1)).v.clone()

        def step():
            step.C = la.spmm(A, Bm)

        step()
        if not torch.equal(step.C.A[:, k - 1], cols_ref):
            raise SystemExit("bench.py: A*B disagrees with A*B[:, k]; refusing to time a wrong result")
    else:
        def step():
            la.mul(y, Aop, x)

    # ---- correctness guard: a number from a wrong kernel is worthless ------------------------------------------
    check = "none"
    if kind == 1 and op != "cg":
        ones = la.HPCVector.from_local(np.ones(x.local_size, dtype=T), backend)
        y1 = la.matvec(Aop, ones).local_values()
        part = Aop.row_partition
        g = np.arange(int(part[rank]) - 1, int(part[rank + 1]) - 1)
        nx, ny, nz = grid
        ix, iy, iz = g % nx, (g // nx) % ny, g // (nx * ny)
        expect = (ix == 0).astype(np.float64) + (ix == nx - 1) + (iy == 0) + (iy == ny - 1) + (iz == 0) + (iz == nz - 1)
        if not np.array_equal(y1, expect):
            raise SystemExit("bench.py: A*ones does not reproduce the exact row sums of the Poisson stencil; refusing to time a wrong result")
        check = "A*ones == exact row sums (bitwise)"
        del ones, y1, g, ix, iy, iz, expect

    sampler = ClockSampler(local_rank)
    for _ in range(warmup):
        step()
    barrier()

    # ---- timed region: exactly `steps` steps, device-timed, max over ranks ---------------------------------------
    launches0 = int(L.hpcla_spmv_launch_count(opnd))
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.start()
    barrier()
    ev0.record()
    if op == "cg":
        la.cg(A, bvec, steps)  # x0 = 0, r = p = b (one pass), then `steps` iterations, all scalars on the device
    else:
        for _ in range(steps):
            step()
    ev1.record()
    barrier()
    sampler.stop()
    elapsed_ms = ev0.elapsed_time(ev1)
    launches = int(L.hpcla_spmv_launch_count(opnd)) - launches0
    region = "timed"
    n_samples = len(sampler.samples)
    if dist is not None:
        tmax = torch.tensor([elapsed_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        elapsed_ms = float(tmax.item())
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

    # ---- end to end: host (pinned) x in, host y out, every step ---------------------------------------------------
    e2e = None
    if op not in ("cg", "spmm"):
        xh = torch.empty(x.local_size, dtype=x.v.dtype, pin_memory=True).copy_(x.v)
        yh = torch.empty(y.local_size, dtype=y.v.dtype, pin_memory=True)

        def e2e_serial_step():  # the three calls a user of the device API would make, back to back on one stream
            x.v.copy_(xh, non_blocking=True)
            la.mul(y, Aop, x)
            yh.copy_(y.v, non_blocking=True)

        def e2e_step():  # the staged entry point: upload, multiply and download pipelined block by block
            la.mul_staged(y, Aop, x, xh, yh)

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
            ms = e0.elapsed_time(e1)
            if dist is not None:
                tmax = torch.tensor([ms], device="cuda", dtype=torch.float64)
                dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
                ms = float(tmax.item())
            return ms / e2e_steps

        yh.zero_()
        e2e_step()
        torch.cuda.synchronize()
        if not torch.equal(yh, y.v.cpu()):
            raise SystemExit("bench.py: the staged multiply did not deliver y to the host buffer")
        serial_ms = time_e2e(e2e_serial_step)
        staged_ms = time_e2e(e2e_step)
        e2e = {"value": flops_step / (staged_ms * 1e-3) / 1e9, "unit": "GFLOP/s", "h2d_bytes_per_step": int(n * x.v.element_size()),
               "d2h_bytes_per_step": int(n * y.v.element_size()), "steps": e2e_steps, "ms_per_step": staged_ms,
               "path": "hpcla_b200.mul_staged(y, A, x, x_host, y_host) = hpcla_spmv_run_staged: pinned host x -> x.v, y.v = A*x, y.v -> pinned host y, "
                       "pipelined over row blocks inside the timed region; A resident",
               "serial_copies_ms_per_step": serial_ms,
               "serial_copies_value": flops_step / (serial_ms * 1e-3) / 1e9}

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
    roofline = {"bound": "hbm", "achieved": per_gpu_gbs, "peak": peak, "unit": "GB/s", "frac": per_gpu_gbs / peak, "traffic": traffic,
                "peak_source": peak_src, "frac_of_nominal_8TBs": per_gpu_gbs / NOMINAL_HBM_GBS,
                "kernel": ("spmv_rowwalk_kernel" if info["rowwalk_tiles"] >= info["general_tiles"] else "spmv_tile_kernel") + ("" if world == 1 else " (interior + boundary launches; step time includes the NCCL halo)"),
                "algorithmic_bytes_per_step": bytes_step, "flops_per_step": flops_step}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        t, fl, bts, workers, desc = cpu_baseline_run(spec, reps=7, warmup=1, workers=args.cpu_workers)
        cpu = {"value": fl / t / 1e9, "unit": "GFLOP/s", "cores": workers, "kind": "port", "sample": desc, "achieved_gbs": bts / t / 1e9}

    if rank == 0:
        line = {
            "metric": "spmv_gflops", "value": gflops, "unit": "GFLOP/s", "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": spec["scaling"], "vs_baseline": None, "dtype": spec["T"],
            "data": "synthetic",
            "config": {"workload": spec["name"], "index_type": spec["Ti"], "op": op, "rows": n, "nnz": nnz, "l2": "inputs exceed L2 (no flush needed)" if bytes_step / world > 2 * 126e6 else "inputs do NOT exceed L2",
                       "tiles": info["tiles"], "interior_tiles": info["interior_tiles"], "boundary_tiles": info["boundary_tiles"],
                       "x_in_place": info["x_in_place"], "lanes_per_row": info["lanes_per_row"], "tile_window": info["tile_window"], "rowwalk_tiles": info["rowwalk_tiles"], "general_tiles": info["general_tiles"], "check": check, "setup_s": round(setup_s, 2)},
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
    ap.add_argument("--cpu-workers", type=int, default=0, help="worker threads of the CPU arm (0 = one per host core; 4 for the 2-D Laplacian, as BASELINE.json)")
    args = ap.parse_args()
    spec = workload_spec(args.workload, args.gpus)
    if args.impl == "reference":
        run_reference_arm(args, spec)
    else:
        run_b200_arm(args, spec)


if __name__ == "__main__":
    main()
