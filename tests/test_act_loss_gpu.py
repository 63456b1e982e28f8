"""Output-layer activations and losses beyond softmax + cross-entropy against
tests/golden/act_loss.npz — the outputs of the reference's own functions (activation.py:8-168, loss.py:10-156) and
of its Network for one training step + a prediction per activation / loss pairing (oracle/gen_golden.py
gen_act_loss).  Element-wise maps: 2 ulp of the exp / tanh implementations (2e-6 relative); network state: 1e-5."""
import numpy as np
import pytest

from conftest import rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def g(golden):
    import torch  # noqa: F401
    from mlpcode import _native
    _native.load()
    return golden("act_loss")


def close(got, want, tol=2e-6, atol=1e-37):
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    fin = np.isfinite(want)
    assert np.array_equal(np.isnan(got), np.isnan(want)) and np.array_equal(np.isinf(got), np.isinf(want))
    assert np.array_equal(np.sign(got[~fin & ~np.isnan(want)]), np.sign(want[~fin & ~np.isnan(want)]))
    return bool(np.all(np.abs(got[fin] - want[fin]) <= tol * np.maximum(np.abs(want[fin]), 1e-30) + atol))


@pytest.mark.parametrize("name", ["sigmoid", "softmax", "relu", "tanh", "sign", "leaky_relu", "identity"])
def test_activation_and_derivative(g, name):
    from mlpcode.activation import ACTIVATION_DERIVATIVES, ACTIVATION_FUNCTIONS, ActivationFuncs as af
    act = af[name]
    a = ACTIVATION_FUNCTIONS[act](g["x"].copy())
    a = a.get() if hasattr(a, "get") else np.asarray(a)
    # the device softmax hands out probabilities already clipped to [1e-15, 1] (the in-place clip the loss applies
    # to them anyway, loss.py:32): entries below 1e-15 differ from the reference's by less than that
    assert close(a, g[f"act_{name}"], atol=1.01e-15 if name == "softmax" else 1e-37), name
    if name in ("relu", "leaky_relu", "sign", "identity"):          # no transcendental: bit-exact
        assert np.array_equal(a.view(np.uint32), g[f"act_{name}"].view(np.uint32))
    d = ACTIVATION_DERIVATIVES[act](g["dA"].copy(), g[f"grad_in_{name}"].copy()).get()
    if name == "softmax":   # s_j * (dA_j - sum_k s_k dA_k): the difference cancels, so the bar is norm-wise
        assert rel_err(d, g["grad_softmax"]) < 1e-6
    else:
        assert close(d, g[f"grad_{name}"]), name


def test_hard_sigmoid_and_hard_tanh(g):
    from mlpcode import activation as A
    for name in ("hard_sigmoid", "hard_tanh"):
        got = getattr(A, name)(g["x"].copy()).get()
        assert np.array_equal(got.view(np.uint32), g[f"act_{name}"].view(np.uint32)), name


def test_losses_and_derivatives(g):
    import torch
    from mlpcode import loss as L
    from mlpcode.device import DeviceArray
    p, y, ypm, t = g["loss_p"], g["loss_y"], g["loss_ypm"], g["loss_t"]

    def dev(a):
        return DeviceArray(torch.from_numpy(a.copy()).cuda())
    pd = dev(p)
    rows = L.cross_entropy_loss(pd, y)
    assert close(rows.get(), g["ce_rows"], 2e-6)
    assert abs(float(rows.mean()) - float(g["ce_rows"].mean())) <= 1e-6 * abs(float(g["ce_rows"].mean()))
    assert pd.get().min() >= 1e-15 and pd.get()[0, 0] == np.float32(1e-15)       # clipped IN PLACE (loss.py:32)
    assert close(L.cross_entropy_derivative(dev(p), y, with_softmax=True).get(), g["ce_grad_fused"])
    assert close(L.cross_entropy_derivative(dev(g["ce_general_p"]), y, with_softmax=False).get(),
                 g["ce_grad_general"], 4e-6)
    assert close(L.mse(dev(p), y).get(), g["mse_rows"]) and close(L.mse_derivative(dev(p), y).get(), g["mse_grad"])
    assert close(L.hinge_loss(dev(t), ypm).get(), g["hinge_rows"])
    assert close(L.hinge_derivative(dev(t), ypm).get(), g["hinge_grad"].astype(np.float32))
    # the reference's own 0/0 and 1/0 at p == 1 are reproduced, not repaired
    one = np.array([[1.0, 0.0]], np.float32)
    got = L.cross_entropy_derivative(dev(one), np.array([[1, 0]], np.int8), with_softmax=False).get()
    assert np.isnan(got[0, 0]) or np.isinf(got[0, 0])


def _state_close(nn, g, tag, n):
    gm, bt, mu, sg = nn.batchNormParams
    for i in range(n):
        assert rel_err(nn.weights[i].get(), g[f"{tag}W{i}"]) < 1e-5, (tag, f"W{i}")
        for name, dev in (("gamma", gm[i]), ("mu", mu[i]), ("sigma", sg[i])):
            assert rel_err(dev.get(), g[f"{tag}{name}{i}"]) < 1e-5, (tag, name, i)
        got, want = bt[i].get(), g[f"{tag}beta{i}"]
        if i < n - 1:
            assert np.abs(got - want).max() < 1e-6, (tag, "beta", i)
        else:
            assert rel_err(got, want) < 1e-5, (tag, "beta out")


@pytest.mark.parametrize("pair", ["sigmoid+mse", "sigmoid+cross_entropy", "tanh+hinge", "softmax+mse", "relu+mse",
                                  "leaky_relu+mse", "identity+mse", "softmax+hinge"])
def test_network_step_with_other_outputs_and_losses(g, pair):
    """A binarized network ending in each activation, trained one step against each loss (network.py:377-387:
    fused gradient for softmax / sigmoid + cross-entropy, activation derivative otherwise), then predict()."""
    from mlpcode.activation import ActivationFuncs as af
    from mlpcode.loss import LossFuncs as lf
    from mlpcode.network import Network
    assert pair in list(g["net_pairs"])
    out_af, loss = pair.split("+")
    units = [int(u) for u in g["net_units"]]
    n = len(units) - 1
    nn = Network(units, useGpu=True, binarized=True, useBatchNorm=True)
    nn.compile(lr=0.01, hiddenAf=af.sign, outAf=af[out_af], lossF=lf[loss])
    nn.load_weights([g[f"net_init_W{i}"].copy() for i in range(n)])
    Y = g["net_Y"].copy()
    if loss == "hinge":
        Y = np.where(Y == 0, -1, Y).astype(np.int8)                               # network.py:301-303
    cost = nn.train_step(g["net_X"].copy(), Y, 0.01)
    tag = f"net_{out_af}_{loss}_"
    assert abs(cost - float(g[tag + "cost"])) <= 1e-5 * abs(float(g[tag + "cost"])), (cost, float(g[tag + "cost"]))
    _state_close(nn, g, tag, n)
    pred = nn.predict(g["net_X2"].copy()).get()
    assert rel_err(pred, g[tag + "pred"]) < 1e-5


def test_train_converts_labels_for_hinge(g, capsys):
    """Network.train with the hinge loss maps one-hot {0, 1} to {-1, +1} (network.py:301-303) without touching the
    caller's array, and runs."""
    from mlpcode.activation import ActivationFuncs as af
    from mlpcode.loss import LossFuncs as lf
    from mlpcode.network import Network
    units = [int(u) for u in g["net_units"]]
    nn = Network(units, useGpu=True, binarized=True, useBatchNorm=True)
    nn.compile(lr=0.01, hiddenAf=af.sign, outAf=af.tanh, lossF=lf.hinge)
    nn.load_weights([g[f"net_init_W{i}"].copy() for i in range(len(units) - 1)])
    Y = g["net_Y"].copy()
    hist = nn.train(g["net_X"].copy(), Y, 1, batch_size=-1, shuffle=False, valX=g["net_X2"].copy(),
                    valY=g["net_Y"].copy(), save_best_params=False)
    assert np.array_equal(Y, g["net_Y"])
    assert abs(hist["loss"][0] - float(g["net_tanh_hinge_cost"])) <= 1e-5 * float(g["net_tanh_hinge_cost"])
