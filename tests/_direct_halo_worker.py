"""Direct halo (copy-engine peer push + stream-wait flags) between the rank-threads of ONE process on one GPU, checked
against the oracle.  Run in a FRESH process by tests/test_gpu_round2.py::test_direct_halo_between_rank_threads: a stream that
sits in a flag wait must never share a hardware queue with the peer's streams, and a pytest process that has already created
hundreds of streams cannot guarantee that (CUDA_DEVICE_MAX_CONNECTIONS queues are shared by all streams of the context).
Between processes — the real multi-GPU case, tests/_nccl_worker.py section 6 — every rank has its own context."""
import os
import sys

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import scipy.sparse as sp  # noqa: E402
import torch  # noqa: E402

import hpcla_b200 as la  # noqa: E402
from oracle import oracle as orc  # noqa: E402


def relerr(y, ref):
    return np.linalg.norm(np.asarray(y) - np.asarray(ref)) / max(np.linalg.norm(np.asarray(ref)), 1e-300)


def main(P):
    S = la.synth
    N = 30
    n = N**3
    rp, c, v = S.stencil_local(1, N, 0, n, np.float64, np.int32)
    G = sp.csr_matrix((v, c - 1, rp - 1), shape=(n, n))
    xh = S.vector_local(np.float64, S.X_SEED, 0, n)
    rng = np.random.default_rng(31)
    R = sp.random(1800, 1500, density=0.01, random_state=rng, format="csr")
    R.data = rng.uniform(-1, 1, R.nnz)
    R = (R @ sp.diags((np.arange(1500) % 3 != 1).astype(np.float64))).tocsr()  # send runs with holes -> pack kernel
    R.eliminate_zeros()
    xr = rng.uniform(-1, 1, 1500)
    xp = np.concatenate([[1], np.sort(rng.integers(1, 1501, size=P - 1)), [1501]]).astype(np.int64)

    def body(rank, bs):
        b = bs[rank]
        torch.cuda.set_device(b.torch_device())
        A = S.stencil_matrix(1, N, b)
        x = S.vector(n, b)
        y_pull = (A * x).to_global()  # rank-thread exchange (device-to-device copies pulled by the receiver)
        la.enable_direct_halo(A, x)
        ys = []
        for k in range(4):  # back-to-back steps: the flags count steps, the consumed flags hold back the next push
            x.v.mul_(-1.0 if k else 1.0)
            ys.append((A * x).to_global())
        g = la.execute_plan(la.get_vector_plan(A, x), A, x)
        torch.cuda.synchronize()
        gath = g.cpu().numpy().copy()
        y_after = (A * x).to_global()
        A2 = la.HPCSparseMatrix.from_global(R, b)
        x2 = la.HPCVector.from_global(xr, b, partition=xp)
        la.enable_direct_halo(A2, x2)
        y2 = [(A2 * x2).to_global() for _ in range(2)]
        return y_pull, ys, gath, y_after, y2, la.spmv_info(A2, x2)

    bs = la.backends_threads(P, np.float64, np.int32, cuda=True)
    res = bs[0].comm.world.run(body, bs)
    olocs = orc.distribute(G, P, itype="i32")
    y_ref = orc.matvec(olocs, xh)
    part = orc.uniform_partition(n, P)
    W = orc.PlanWorld(olocs, part)
    g_ref = W.execute(orc.split_vector(-xh, part))  # x was negated an odd number of times before the gather
    W.close()
    y2_ref = orc.matvec(orc.distribute(R, P, itype="i32"), xr, xp)
    for r, (y0, ys, gath, y_after, y2, info2) in enumerate(res):
        assert np.array_equal(y0, y_ref)
        sign = 1.0
        for k, y in enumerate(ys):
            sign *= -1.0 if k else 1.0
            assert np.array_equal(y, sign * y_ref), k
        assert np.array_equal(gath, g_ref[r])
        assert np.array_equal(y_after, -y_ref)
        assert info2["sends_contiguous"] == 0
        for y in y2:
            assert relerr(y, y2_ref) <= 1e-12
    print("DIRECT_HALO_OK", flush=True)


if __name__ == "__main__":
    main(int(sys.argv[1]))
