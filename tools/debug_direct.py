"""Step-by-step run of the direct halo between two rank-threads on one GPU, with flag dumps (debugging aid)."""
import ctypes, os, sys, threading, time
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
if os.environ.get("EAGER"): os.environ["CUDA_MODULE_LOADING"] = "EAGER"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import hpcla_b200 as la

P = 2
bs = la.backends_threads(P, np.float64, np.int32, cuda=True)
N = 30; n = N**3
ops = {}
def dump(tag):
    L = la._lib.lib()
    for r, op in sorted(ops.items()):
        out = (ctypes.c_uint * (2 * P + 1))()
        L.hpcla_spmv_halo_debug(op, out)
        print(f"   [{tag}] rank {r}: arrival {list(out[:P])} consumed {list(out[P:2*P])} step {out[2*P]}", flush=True)

def body(rank, bs):
    b = bs[rank]
    torch.cuda.set_device(b.torch_device())
    A = la.synth.stencil_matrix(1, N, b)
    x = la.synth.vector(n, b)
    y = A * x
    la.enable_direct_halo(A, x)
    ops[rank] = la.sparse._bound_op(A, la.get_vector_plan(A, x), x)
    for k in range(3):
        y = A * x
        y.v.cpu()
        if rank == 0: print("multiply", k, "done (stream sync)", flush=True)
    b.comm.world.barrier.wait()
    if rank == 0: print("device-wide synchronize after multiplies ...", flush=True)
    torch.cuda.synchronize()
    if rank == 0: print("... ok", flush=True)
    b.comm.world.barrier.wait()
    g = la.execute_plan(la.get_vector_plan(A, x), A, x)
    if rank == 0: print("gather enqueued", flush=True)
    time.sleep(2)
    b.comm.world.barrier.wait()
    if rank == 0:
        dump("after gather, before any sync")
        for name in ("ev_x", "ev_packed", "ev_halo"):
            pass
    b.comm.world.barrier.wait()
    torch.cuda.current_stream().synchronize()
    if rank == 0: print("gather: stream sync ok", flush=True)
    torch.cuda.synchronize()
    if rank == 0: print("gather: device sync ok", flush=True)
    y = A * x
    y.v.cpu()
    if rank == 0: print("multiply after gather ok", flush=True)
    return 0

def watchdog():
    for i in range(8):
        time.sleep(5)
        try:
            dump(f"t+{5*(i+1)}s")
        except Exception as e:
            print("dump failed", e, flush=True)
bs[0].comm.world.run(body, bs)
print("ALL DONE", flush=True)
