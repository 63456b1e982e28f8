"""Epoch loop on the device: the permutation gather kernel and Network.train against
tests/golden/train_epochs.npz — the history and the state the REFERENCE's own Network.train (network.py:264-366)
produced on these inputs with these permutations (oracle/gen_golden.py gen_epochs; the oracle's train_epochs equals
it exactly).  tests/test_oracle.py re-checks the oracle against the file on the CPU."""
import numpy as np
import pytest

from conftest import GOLDEN, rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def T():
    import torch
    from mlpcode import _native
    _native.load()
    return torch


@pytest.mark.parametrize("n,F,dt", [(1000, 784, "float32"), (77, 10, "int8"), (513, 33, "float32"),
                                    (64, 3072, "uint8"), (5, 1, "int8")])
def test_gather_rows(T, n, F, dt):
    """dst[i] = src[idx[i]]: vector path (rows of whole 16-byte words) and byte path, pitched source and
    destination, repeated indices; indices outside the set are zero-filled and counted; guard rows untouched."""
    from mlpcode import kernels as K
    rng = np.random.default_rng(n + F)
    tdt = getattr(T, dt)
    pitch = F + (4 if F % 4 == 0 else 3)
    src_full = T.from_numpy(rng.integers(-100, 100, (n, pitch)).astype(dt)).cuda()
    src = src_full[:, :F]
    m = 2 * n + 3
    idx_h = rng.integers(0, n, m)
    idx_h[::7] = [-1, n, n + 5, -100][0:1] * len(idx_h[::7])          # out of range
    idx_h[3::11] = n
    idx = T.from_numpy(idx_h.astype(np.int64)).cuda()
    dst_full = T.full((m + 2, pitch), 55, dtype=tdt, device="cuda")
    dst = dst_full[1:m + 1, :F]
    bad = T.zeros(1, dtype=T.int32, device="cuda")
    K.gather_rows(src, idx, dst, bad)
    ok = (idx_h >= 0) & (idx_h < n)
    want = np.where(ok[:, None], src.cpu().numpy()[np.clip(idx_h, 0, n - 1)], 0)
    assert np.array_equal(dst.cpu().numpy(), want)
    assert int(bad.item()) == int((~ok).sum())
    g = dst_full.cpu().numpy()
    assert np.all(g[0] == 55) and np.all(g[-1] == 55) and np.all(g[1:m + 1, F:] == 55)


def _net(g, prefix):
    from mlpcode.activation import ActivationFuncs as af
    from mlpcode.loss import LossFuncs as lf
    from mlpcode.network import Network
    units = [int(u) for u in g["units"]]
    n = len(units) - 1
    nn = Network(units, useGpu=True, binarized=True, useBatchNorm=True)
    nn.compile(lr=float(g["lr"]), hiddenAf=af.sign, outAf=af.softmax, lossF=lf.cross_entropy)
    nn.load_weights([g[f"{prefix}W{i}"].copy() for i in range(n)])
    nn.loadBatchNormParameters(*[[g[f"{prefix}{k}{i}"].copy() for i in range(n)]
                                 for k in ("gamma", "beta", "mu", "sigma")])
    return nn, n


def _state_ok(nn, g, prefix, n):
    gm, bt, mu, sg = nn.batchNormParams
    for i in range(n):
        assert rel_err(nn.weights[i].get(), g[f"{prefix}W{i}"]) < 1e-5, f"W{i}"
        for name, dev in (("gamma", gm[i]), ("mu", mu[i]), ("sigma", sg[i])):
            assert rel_err(dev.get(), g[f"{prefix}{name}{i}"]) < 1e-5, f"{name}{i}"
        got, want = bt[i].get(), g[f"{prefix}beta{i}"]
        if i < n - 1:      # hidden-layer beta holds rounding noise only (sum_b delta vanishes after a BN backward)
            assert np.abs(got - want).max() < 1e-6
        else:
            assert rel_err(got, want) < 1e-5, "beta (output layer)"


@pytest.mark.parametrize("graph", [False, True])
@pytest.mark.parametrize("best", [False, True])
def test_train_epochs_match_the_reference(T, golden, graph, best, capsys):
    """Three shuffled epochs of two mini-batches: epoch losses within 1e-5, train / validation accuracies equal,
    final state (or the restored best-validation snapshot) within 1e-5 — eager launches and CUDA-graph replays,
    whose static batch buffers the gather kernel fills directly."""
    g = golden("train_epochs")
    nn, n = _net(g, "init_")
    nn.enable_step_graph(graph)
    perms = iter(g["perms"])
    nn.shuffle_source = lambda count: next(perms)
    hist = nn.train(g["X"].copy(), g["Y"].copy(), len(g["loss"]), batch_size=int(g["batch_size"]), shuffle=True,
                    valX=g["Xv"].copy(), valY=g["yv"].copy(), save_best_params=best)
    assert np.allclose(hist["loss"], g["loss"], rtol=1e-5, atol=0)
    assert list(hist["accuracy"]) == list(g["accuracy"])
    assert list(hist["val_accuracy"]) == list(g["val_accuracy"])
    assert next(perms, None) is None                        # one permutation per epoch was consumed
    _state_ok(nn, g, "best_" if best else "final_", n)
    out = capsys.readouterr().out
    assert "Epoch 3 / 3" in out and ("Switching to best params" in out) == best


def test_train_without_shuffle_and_default_permutation(T, golden):
    """shuffle=False takes plain slices (the oracle on the unshuffled batches); the default shuffle source is a
    device permutation: every row is visited exactly once per epoch (the epoch loss is permutation dependent,
    the set of rows is not — checked through the gather's own bad-index counter and a run that completes)."""
    import bnn_oracle as O
    from conftest import params_from_golden
    g = golden("train_epochs")
    nn, n = _net(g, "init_")
    ref = params_from_golden(g, "init_", n)
    hist = nn.train(g["X"].copy(), g["Y"].copy(), 1, batch_size=int(g["batch_size"]), shuffle=False,
                    valX=g["Xv"].copy(), valY=g["yv"].copy(), save_best_params=False)
    want = O.train_epochs(ref, g["X"].copy(), g["Y"].copy(), float(g["lr"]), 1, int(g["batch_size"]))
    assert np.allclose(hist["loss"], want["loss"], rtol=1e-5) and hist["accuracy"] == want["accuracy"]
    nn2, _ = _net(g, "init_")
    T.manual_seed(3)
    h2 = nn2.train(g["X"].copy(), g["Y"].copy(), 2, batch_size=128, shuffle=True, valX=g["Xv"].copy(),
                   valY=g["yv"].copy(), save_best_params=False)          # ragged last batch: 400 = 3 * 128 + 16
    assert len(h2["loss"]) == 2 and all(np.isfinite(h2["loss"]))
