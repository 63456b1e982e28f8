"""Real-valued LinearLayer networks (binarized=False; layers.py:130-271 with batch-norm, no bias) against
tests/golden/linear_net.npz — three training steps and a prediction of the reference's own Network for each smooth
hidden activation (oracle/gen_golden.py gen_linear) — and the eval-mode weight hook that carries the non-BNN fault
path (ErrorCallback(p, mode, bnn=False): IEEE-754 bit flips of the real weights, callbacks.py:85-123).
Forward and backward are three fp16-split tensor-core passes per product; bar 1e-5."""
import numpy as np
import pytest

from conftest import rel_err

pytestmark = pytest.mark.gpu


def _net(g, hidden, lr=0.05):
    from mlpcode.activation import ActivationFuncs as af
    from mlpcode.loss import LossFuncs as lf
    from mlpcode.layers import BinaryLayer, LinearLayer
    from mlpcode.network import Network
    units = [int(u) for u in g["units"]]
    n = len(units) - 1
    nn = Network(units, useGpu=True, binarized=False, useBatchNorm=True)
    assert all(type(layer) is LinearLayer and not isinstance(layer, BinaryLayer) for layer in nn._layers)
    nn.compile(lr=lr, hiddenAf=af[hidden], outAf=af.softmax, lossF=lf.cross_entropy)
    nn.load_weights([g[f"init_W{i}"].copy() for i in range(n)])
    return nn, n


@pytest.mark.parametrize("hidden", ["sigmoid", "relu", "tanh", "leaky_relu"])
def test_linear_network_steps_match_reference(golden, hidden):
    g = golden("linear_net")
    assert hidden in list(g["hidden"])
    nn, n = _net(g, hidden)
    for i in range(3):
        cost = nn.train_step(g[f"X{i}"].copy(), g[f"Y{i}"].copy(), 0.05)
        want = float(g[f"{hidden}_cost{i}"])
        assert abs(cost - want) <= 1e-5 * abs(want), (i, cost, want)
        gm, bt, mu, sg = nn.batchNormParams
        for j in range(n):
            tag = f"{hidden}_s{i}_"
            assert rel_err(nn.weights[j].get(), g[f"{tag}W{j}"]) < 1e-5, (i, j)
            for name, dev in (("gamma", gm[j]), ("mu", mu[j]), ("sigma", sg[j])):
                assert rel_err(dev.get(), g[f"{tag}{name}{j}"]) < 1e-5, (i, j, name)
            got, wantb = bt[j].get(), g[f"{tag}beta{j}"]
            if j < n - 1:      # dbeta = sum_b delta vanishes after a batch-norm backward: rounding noise on both sides
                assert np.abs(got - wantb).max() < 1e-6
            else:
                assert rel_err(got, wantb) < 1e-5
    assert rel_err(nn.predict(g["X_eval"].copy()).get(), g[f"{hidden}_pred"]) < 1e-5


def test_float_fault_hook_on_linear_layers(golden):
    """Eval mode: every LinearLayer hands its dense fp32 weights to the callback (layers.py:215-218) and multiplies
    with what comes back.  With ErrorCallback(p, mode, bnn=False) that is the IEEE-754 bit-flip path; the same flips
    applied by hand (same seed and call counter) and loaded as plain weights give the identical prediction, and the
    latent weights stay untouched."""
    from mlpcode.callbacks import ErrorCallback
    g = golden("linear_net")
    nn, n = _net(g, "relu")
    X = g["X_eval"].copy()
    clean = nn.predict(X).get()
    before = [w.get().copy() for w in nn.weights]
    nn.addCallbacks(ErrorCallback(0.05, mode=2, bnn=False, seed=5))
    faulty = nn.predict(X).get()
    assert not np.array_equal(faulty, clean)
    assert all(np.array_equal(w.get(), b) for w, b in zip(nn.weights, before))          # latent weights untouched
    nn.clearCallbacks()
    hand = ErrorCallback(0.05, mode=2, bnn=False, seed=5)      # one object: its call counter numbers the trials
    by_hand = [hand(w).get() for w in nn.weights]
    assert any(not np.array_equal(a, b) for a, b in zip(by_hand, before))
    nn2, _ = _net(g, "relu")
    nn2.load_weights(by_hand)
    same = nn2.predict(X).get()
    assert np.array_equal(np.isnan(same), np.isnan(faulty))
    ok = ~np.isnan(same)
    assert np.array_equal(same[ok], faulty[ok])
