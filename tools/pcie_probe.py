#!/usr/bin/env python
"""pcie_probe.py — what the host link allows for the staged multiply (e2e): pinned H2D alone, D2H alone and both at once,
134 MB each (x and y of BASELINE config 2), for torch's pinned allocator and for hpcla_host_alloc (first-touched on the
GPU's NUMA node), on one GPU or — under torch.distributed.run — on all ranks at the same time.

    python tools/pcie_probe.py                                   # 1 GPU
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/pcie_probe.py

One JSON line per case on rank 0: GB/s per direction per GPU (max time over ranks).
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hpcla_b200 as la  # noqa: E402


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = 256**3
    dx, dy = torch.zeros(n, dtype=torch.float64, device="cuda"), torch.zeros(n, dtype=torch.float64, device="cuda")
    b = la.backend_cuda_serial(np.float64, np.int32, device=local)
    hb = [la.host_buffer(b, n), la.host_buffer(b, n)]
    bufs = {
        "torch pin_memory": (torch.empty(n, dtype=torch.float64, pin_memory=True), torch.empty(n, dtype=torch.float64, pin_memory=True)),
        "hpcla_host_alloc (GPU's NUMA node)": (torch.from_numpy(hb[0].array), torch.from_numpy(hb[1].array)),
    }
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, reps=10):
        for _ in range(2):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1) / reps
        if dist is not None:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    def chunks(t, k):
        m = t.numel()
        step = (m + k - 1) // k
        return [(i, min(m, i + step)) for i in range(0, m, step)]

    for name, (hx, hy) in bufs.items():
        for nchunk in (1, 16):
            def h2d():
                for lo, hi in chunks(dx, nchunk):
                    dx[lo:hi].copy_(hx[lo:hi], non_blocking=True)

            def d2h():
                for lo, hi in chunks(dy, nchunk):
                    hy[lo:hi].copy_(dy[lo:hi], non_blocking=True)

            def both():
                cur = torch.cuda.current_stream()
                s1.wait_stream(cur)
                s2.wait_stream(cur)
                with torch.cuda.stream(s1):
                    h2d()
                with torch.cuda.stream(s2):
                    d2h()
                cur.wait_stream(s1)
                cur.wait_stream(s2)

            res = {"host_memory": name, "chunks": nchunk, "n_gpus": world, "bytes_each_way": n * 8, "numa_node": hb[0].numa_node}
            for label, fn in (("h2d_alone", h2d), ("d2h_alone", d2h), ("both_at_once", both)):
                ms = timed(fn)
                res[label + "_ms"] = round(ms, 4)
                res[label + "_gbs_per_direction"] = round(n * 8 / (ms * 1e-3) / 1e9, 2)
            if rank == 0:
                print(json.dumps(res), flush=True)
    if rank == 0:
        try:
            print(json.dumps({"nproc": os.cpu_count(), "affinity": len(os.sched_getaffinity(0)),
                              "numa_nodes": sorted(d for d in os.listdir("/sys/devices/system/node") if d.startswith("node"))}), flush=True)
        except Exception as e:  # noqa: BLE001
            print(json.dumps({"topology": repr(e)}), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
