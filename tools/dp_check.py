"""torchrun --nproc-per-node N tools/dp_check.py — N-GPU data-parallel training steps on the golden
fixtures must equal the reference's full-batch steps (state within 1e-5): sync-BN statistics, global
loss gradient and all-reduced dW over NCCL."""
import sys
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(REPO / "deep-neural-net_b200"), str(REPO / "tests")]


def main():
    import torch
    from conftest import params_from_golden, rel_err
    from mlpcode import parallel
    from mlpcode.activation import ActivationFuncs as af
    from mlpcode.loss import LossFuncs as lf
    from mlpcode.network import Network
    comm = parallel.init_from_env("nccl")
    assert comm is not None and comm.world >= 2
    worst = 0.0
    for name in ("train_tiny", "train_small", "train_ragged"):
        with np.load(REPO / "tests" / "golden" / f"{name}.npz") as f:
            g = {k: f[k] for k in f.files}
        units = [int(u) for u in g["units"]]
        n = len(units) - 1
        p = params_from_golden(g, "init_", n)
        nn = Network(units, useGpu=True, binarized=True, useBatchNorm=True)
        nn.compile(lr=float(g["lr"]), hiddenAf=af.sign, outAf=af.softmax, lossF=lf.cross_entropy)
        nn.load_weights(p["W"])
        nn.loadBatchNormParameters(*[[b[k] for b in p["bn"]] for k in ("gamma", "beta", "mu", "sigma")])
        nn.set_data_parallel(comm)
        for s in range(g["X"].shape[0]):
            Bg = g["X"].shape[1]
            lo, hi = parallel.shard_range(Bg, comm.rank, comm.world)
            nn.train_step(g["X"][s][lo:hi], g["Y"][s][lo:hi], float(g["lr"]), global_batch=Bg)
            gam, bet, mus, sig = nn.batchNormParams
            for i in range(n):
                errs = {"W": rel_err(nn.weights[i].get(), g[f"after{s}_W{i}"])}
                for key, arr in (("gamma", gam), ("beta", bet), ("mu", mus), ("sigma", sig)):
                    errs[key] = rel_err(arr[i].get(), g[f"after{s}_{key}{i}"])
                if comm.rank == 0:
                    print(name, "step", s, "layer", i, {k: f"{v:.2e}" for k, v in errs.items()})
                worst = max(worst, max(errs.values()))
    # ---- widths that take the fused epilogues and the int8 dW product (the fixtures are narrower):
    # one full-batch step of the oracle vs the same step sharded over the ranks
    sys.path.insert(0, str(REPO / "oracle"))
    import bnn_oracle as O
    units, Bg, lr = [64, 512, 256, 10], 512, 0.05
    rng = np.random.default_rng(21)
    params = O.new_params(units, rng)
    X = rng.random((Bg, units[0]), dtype=np.float32) * 2 - 1
    Y = np.eye(10, dtype=np.int8)[rng.integers(0, 10, Bg)]
    ref = O.copy_params(params)
    ref_cost = O.train_step(ref, X, Y, lr)
    nn = Network(units, useGpu=True, binarized=True, useBatchNorm=True)
    nn.compile(lr=lr, hiddenAf=af.sign, outAf=af.softmax, lossF=lf.cross_entropy)
    nn.load_weights([w.copy() for w in params["W"]])
    nn.set_data_parallel(comm)
    lo, hi = parallel.shard_range(Bg, comm.rank, comm.world)
    nn.train_step(X[lo:hi], Y[lo:hi], lr, global_batch=Bg)
    gam, bet, mus, sig = nn.batchNormParams
    n = len(units) - 1
    for i in range(n):
        errs = {"W": rel_err(nn.weights[i].get(), ref["W"][i]),
                "dW": rel_err(nn.weights[i].get().astype(np.float64) - params["W"][i],
                              ref["W"][i].astype(np.float64) - params["W"][i])}
        for key, arr in (("gamma", gam), ("mu", mus), ("sigma", sig)) + ((("beta", bet),) if i == n - 1 else ()):
            errs[key] = rel_err(arr[i].get(), ref["bn"][i][key])
        if comm.rank == 0:
            print("fused_widths layer", i, {k: f"{v:.2e}" for k, v in errs.items()})
        worst = max(worst, max(v for k, v in errs.items() if k != "dW"))
        assert errs["dW"] < 1e-3, errs   # the update itself (fp32 cancellation in W - W0 limits this)
    t = torch.tensor([worst], device="cuda")
    torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    if comm.rank == 0:
        print(f"worst relative state error over ranks: {t.item():.3e}")
        assert t.item() < 1e-5
        print("dp_check ok")
    torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
