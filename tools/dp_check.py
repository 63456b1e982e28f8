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
    t = torch.tensor([worst], device="cuda")
    torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    if comm.rank == 0:
        print(f"worst relative state error over ranks: {t.item():.3e}")
        assert t.item() < 1e-5
        print("dp_check ok")
    torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
