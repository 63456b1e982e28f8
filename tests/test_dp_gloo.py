"""CPU, world_size 2 over gloo: the data-parallel decomposition the device layers use —
batch rows sharded by `shard_range`, per-layer all-reduce of the fp64 batch-norm sums, of the BN
backward reductions and of dW, loss gradient divided by the GLOBAL batch — reproduces the
single-process reference step.  The local arithmetic here is NumPy (test infrastructure standing in
for the kernels); the collectives go through mlpcode.parallel.Comm exactly as on the GPUs."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest

REPO = Path(__file__).resolve().parent.parent
EPS = 1e-4


def dp_train_step(p, X, y, lr, comm, Bg):
    """One step on this rank's shard; mirrors mlpcode.layers.BinaryLayer forward/backwards."""
    import torch
    import bnn_oracle as O

    def allreduce(a):
        t = torch.from_numpy(np.ascontiguousarray(a, np.float64))
        comm.allreduce_sum(t)
        return t.numpy()

    a, caches = X, []
    n = len(p["W"])
    for i in range(n):
        Wb = O.binarize(p["W"][i])
        z = a.dot(Wb)
        sums = allreduce(np.stack([z.astype(np.float64).sum(0), (z.astype(np.float64) ** 2).sum(0)]))
        mu = (sums[0] / Bg).astype(np.float32)
        var = np.maximum(sums[1] / Bg - (sums[0] / Bg) ** 2, 0).astype(np.float32)
        bn = p["bn"][i]
        out = bn["gamma"] * ((z - mu) / np.sqrt(var + np.float32(EPS))) + bn["beta"]
        caches.append(dict(x=a, z=z, mu=mu, var=var, sums=sums))
        a = O.unitstep(out) if i < n - 1 else O.softmax(out)
    loss_local = float(O.cross_entropy_loss(a, y).astype(np.float64).sum())
    delta = ((a - y) / np.float32(Bg)).astype(np.float32)
    for i in reversed(range(n)):
        c, bn = caches[i], p["bn"][i]
        xm = (c["z"] - c["mu"]).astype(np.float64)
        red = allreduce(np.stack([delta.astype(np.float64).sum(0), (delta.astype(np.float64) * xm).sum(0)]))
        sinv = 1.0 / np.sqrt((c["var"] + np.float32(EPS)).astype(np.float64))
        g = bn["gamma"].astype(np.float64)
        dvar = -0.5 * g * red[1] * sinv ** 3
        resid = c["sums"][0] / Bg - c["mu"].astype(np.float64)
        dmu = -g * sinv * red[0] + dvar * (-2.0 * resid)
        nd = ((g * sinv) * delta + (2.0 * dvar / Bg) * xm + dmu / Bg).astype(np.float32)
        bn["gamma"] -= np.float32(lr) * (red[1] * sinv).astype(np.float32)
        bn["beta"] -= np.float32(lr) * red[0].astype(np.float32)
        bn["mu"] = np.float32(0.9) * bn["mu"] + np.float32(1.0 - 0.9) * c["mu"]
        bn["sigma"] = np.float32(0.9) * bn["sigma"] + np.float32(1.0 - 0.9) * c["var"]
        dW = allreduce(c["x"].T.astype(np.float64).dot(nd.astype(np.float64))).astype(np.float32)
        delta = nd.dot(p["W"][i].T)
        p["W"][i] -= np.float32(lr) * dW
    return allreduce(np.array([loss_local]))[0] / Bg


def _worker(rank, world, port, name, q):
    sys.path[:0] = [str(REPO / "deep-neural-net_b200"), str(REPO / "oracle"), str(REPO / "tests")]
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import torch.distributed as dist
    from mlpcode.parallel import init_from_env, shard_range
    from conftest import params_from_golden
    comm = init_from_env("gloo")
    assert comm.world == world and comm.rank == rank
    with np.load(REPO / "tests" / "golden" / f"{name}.npz") as f:
        g = {k: f[k] for k in f.files}
    units = [int(u) for u in g["units"]]
    p = params_from_golden(g, "init_", len(units) - 1)
    worst = 0.0
    for s in range(g["X"].shape[0]):
        Bg = g["X"].shape[1]
        lo, hi = shard_range(Bg, rank, world)
        cost = dp_train_step(p, g["X"][s][lo:hi].copy(), g["Y"][s][lo:hi], float(g["lr"]), comm, Bg)
        worst = max(worst, abs(cost - g["costs"][s]) / abs(g["costs"][s]))
        for i in range(len(units) - 1):
            ref = g[f"after{s}_W{i}"].astype(np.float64)
            worst = max(worst, np.linalg.norm(p["W"][i] - ref) / np.linalg.norm(ref))
            for k in ("gamma", "beta", "mu", "sigma"):
                ref = g[f"after{s}_{k}{i}"].astype(np.float64)
                worst = max(worst, np.linalg.norm(p["bn"][i][k] - ref) / max(np.linalg.norm(ref), 1e-3))
    gathered = comm.allgather(__import__("torch").tensor([worst]))
    comm.barrier()
    if rank == 0:
        q.put((float(gathered.max()), comm.collectives))
    dist.destroy_process_group()


@pytest.mark.parametrize("name", ["train_tiny", "train_ragged"])
def test_two_rank_step_equals_reference(name):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, name, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    worst, collectives = q.get(timeout=120)
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    assert worst < 1e-5, worst          # N ranks == 1 rank == reference, 1e-5 relative
    assert collectives > 0
