"""GPU parity of the drop-in classes (mlpcode.layers / network / callbacks) against the fixtures
generated from the reference and against the oracle on fresh seeded inputs.

Tolerance (north star): bit-exact for +-1 GEMM outputs, packed weights and flip masks; 1e-5
relative (norm-wise) for BN outputs, loss and optimiser state."""
import os

import numpy as np
import pytest

from conftest import params_from_golden, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-5


def build_net(units, params, out="softmax", lr=0.01):
    from mlpcode.activation import ActivationFuncs as af
    from mlpcode.loss import LossFuncs as lf
    from mlpcode.network import Network
    nn = Network(units, useGpu=True, binarized=True, useBatchNorm=True)
    nn.compile(lr=lr, hiddenAf=af.sign, outAf=af.softmax if out == "softmax" else af.identity,
               lossF=lf.cross_entropy)
    nn.load_weights([w.copy() for w in params["W"]])
    nn.loadBatchNormParameters(*[[b[k].copy() for b in params["bn"]]
                                 for k in ("gamma", "beta", "mu", "sigma")])
    return nn


def state_of(nn):
    g, b, m, s = nn.batchNormParams
    return dict(W=[w.get() for w in nn.weights],
                bn=[dict(gamma=g[i].get(), beta=b[i].get(), mu=m[i].get(), sigma=s[i].get())
                    for i in range(len(g))])


@pytest.fixture(params=["tcgen05", "simt"])
def gemm_impl(request, monkeypatch):
    monkeypatch.setenv("BNN_GEMM_IMPL", request.param)
    return request.param


@pytest.mark.parametrize("name", ["train_tiny", "train_small", "train_ragged"])
def test_training_steps_match_reference(golden, gemm_impl, name):
    g = golden(name)
    units = [int(u) for u in g["units"]]
    n = len(units) - 1
    nn = build_net(units, params_from_golden(g, "init_", n), lr=float(g["lr"]))
    for s in range(g["X"].shape[0]):
        cost = nn.train_step(g["X"][s], g["Y"][s], float(g["lr"]))
        assert abs(cost - g["costs"][s]) <= TOL * abs(g["costs"][s]), (s, cost, g["costs"][s])
        st = state_of(nn)
        for i in range(n):
            assert rel_err(st["W"][i], g[f"after{s}_W{i}"]) < TOL, (s, i)
            # the weight UPDATE itself (lr*dW), not just the weights it is a small part of
            prev = g[f"after{s - 1}_W{i}"] if s else g[f"init_W{i}"]
            upd_ref = g[f"after{s}_W{i}"].astype(np.float64) - prev
            upd = st["W"][i].astype(np.float64) - prev
            assert rel_err(upd, upd_ref) < 2e-2, (s, i)   # fp32 cancellation in W - W_prev limits this
            for k in ("gamma", "beta", "mu", "sigma"):
                assert rel_err(st["bn"][i][k], g[f"after{s}_{k}{i}"]) < TOL, (s, i, k)
    probs = nn.predict(g["X_eval"]).get()
    assert rel_err(probs, g["probs_eval"]) < TOL
    assert nn.get_accuracy(g["X_eval"], g["labels_eval"]) == float(g["acc_eval"])


@pytest.mark.parametrize("name", ["train_tiny", "train_small", "train_ragged"])
def test_stage_parity_teacher_forced(golden, name):
    """Each stage of the first step on the reference's own stage inputs: z bit-exact for +-1
    layers, BN output within 1e-5, sign mismatches only where |BN output| is ~0, dW / dX within 1e-5."""
    import torch
    from mlpcode import kernels as K
    from mlpcode.layers import SignActivation
    g = golden(name)
    units = [int(u) for u in g["units"]]
    n = len(units) - 1
    nn = build_net(units, params_from_golden(g, "init_", n), lr=float(g["lr"]))
    B = g["X"].shape[1]
    acts = []
    for i, layer in enumerate(nn._layers):
        if i == 0:
            x_in = g["X"][0]
        else:                                            # teacher forcing: the reference's activations
            a = g[f"s0_act{i - 1}"]
            i8 = torch.zeros((B, (a.shape[1] + 15) // 16 * 16), dtype=torch.int8, device="cuda")
            i8[:, :a.shape[1]] = torch.from_numpy(a.astype(np.int8)).cuda()
            t16 = torch.zeros((a.shape[1], (B + 7) // 8 * 8), dtype=torch.float16, device="cuda")
            t16[:, :B] = torch.from_numpy(a.T.astype(np.float16).copy()).cuda()
            x_in = SignActivation(i8, a.shape[1], t16=t16)
        out = layer.forward(x_in, isTrain=True)
        z = layer.cache["z"][:, :units[i + 1]].cpu().numpy()
        if i == 0:
            assert rel_err(z, g["s0_z0"]) < 5e-6          # real x +-1 through the fp16 split GEMM
        else:
            assert np.array_equal(z.astype(np.float32), g[f"s0_z{i}"])   # bit-exact integers
        bn_ref = g[f"s0_bn_out{i}"]
        if i < n - 1:
            got = out.t.cpu().numpy()
            bad = got != g[f"s0_act{i}"]
            assert np.all(np.abs(bn_ref[bad]) < 1e-4 * np.abs(bn_ref).max()), f"layer {i}"
            assert bad.mean() < 1e-3
        else:
            assert rel_err(out.logits.cpu().numpy(), bn_ref) < TOL
            assert rel_err(out.get(), g["s0_act" + str(i)]) < TOL
        acts.append(out)
    # backward, teacher forced with the reference's incoming deltas
    lr = float(g["lr"])
    for i in reversed(range(n)):
        layer = nn._layers[i]
        d_in = g["s0_dLdA"] if i == n - 1 else g[f"s0_dX{i + 1}"]
        w_before = layer.weights.get().astype(np.float64)
        dX = layer.backwards(torch.from_numpy(d_in).cuda(), lr, activDeriv=False)
        dW = (w_before - layer.weights.get().astype(np.float64)) / lr
        assert rel_err(dW, g[f"s0_dW{i}"]) < 5e-3, i     # recovered through fp32 W: cancellation
        if i > 0:
            assert rel_err(dX.get(), g[f"s0_dX{i}"]) < TOL, i
        else:
            assert dX is None                            # layer-0 dX is dead work in the reference


def test_popc_and_i8_variants_are_bit_identical(golden, monkeypatch):
    g = golden("train_small")
    units = [int(u) for u in g["units"]]
    outs = {}
    for variant in ("i8", "popc"):
        monkeypatch.setenv("BNN_BINARY_GEMM", variant)
        nn = build_net(units, params_from_golden(g, "after1_", len(units) - 1))
        a = g["X_eval"]
        zs = []
        for layer in nn._layers:
            a = layer.forward(a, isTrain=True)
            zs.append(layer.cache["z"][:, :layer.layerUnits].clone())  # pad columns are scratch
        outs[variant] = zs
    import torch
    for z1, z2 in zip(outs["i8"][1:], outs["popc"][1:]):
        assert torch.equal(z1, z2)


def test_weight_fault_callbacks(golden, oracle):
    from mlpcode.callbacks import ErrorCallback
    g = golden("weight_faults")
    units = [int(u) for u in g["units"]]
    n = len(units) - 1
    params = params_from_golden(g, "", n)
    nn = build_net(units, params)
    assert rel_err(nn.predict(g["X"]).get(), g["probs_clean"]) < TOL
    # (i) external-mask mode: the reference's own seeded draws, one mask per layer
    cbs = [ErrorCallback(0.1, mode=2, bnn=True, mask_source=(lambda m: (lambda shape: m))(g[f"mask{i}"]))
           for i in range(n)]
    nn.addCallbacks(cbs)
    assert rel_err(nn.predict(g["X"]).get(), g["probs_fault"]) < TOL
    assert nn.get_accuracy(g["X"], g["labels"]) == float(g["acc_fault"])
    nn.clearCallbacks()
    assert rel_err(nn.predict(g["X"]).get(), g["probs_clean"]) < TOL   # latent weights untouched
    # (ii) in-kernel RNG mode: equals the oracle network run with the oracle's restatement of the mask
    cb = ErrorCallback(0.2, mode=2, bnn=True, seed=4242)
    nn.addCallbacks(cb)
    got = nn.predict(g["X"]).get()
    masks = [oracle.c_bernoulli_mask(units[i], units[i + 1], 0.2, 4242, i, i) for i in range(n)]
    # one shared callback object: its call counter advances layer by layer (trial = call index)
    hooks = [(lambda m: (lambda W: oracle.flip_bnn_mask(W, m)))(m) for m in masks]
    want = oracle.predict(params, g["X"].copy(), weight_hooks=hooks)
    assert rel_err(got, want) < TOL
    # a second forward draws fresh masks (callbacks.py:81 draws per call)
    assert not np.array_equal(nn.predict(g["X"]).get(), got)
    # reference call signature on a dense +-1 array
    cb2 = ErrorCallback(0.2, mode=2, bnn=True, seed=7)
    flipped = cb2(g["cb_Wb"], gpu=True).get()
    m = oracle.c_bernoulli_mask(*g["cb_Wb"].shape, 0.2, 7, 0, 0)
    assert np.array_equal(flipped, oracle.flip_bnn_mask(g["cb_Wb"], m))


def test_image_error_callback_and_normalize(golden, oracle):
    import torch
    from mlpcode import kernels as K
    from mlpcode.callbacks import ImageErrorCallback
    g = golden("image_faults")
    imgs = g["images"]
    for mode in (0, 1, 2):
        icb = ImageErrorCallback(0.14, mode=mode, seed=31)
        assert icb.flips_per_row(imgs.shape[1]) == int(g["k"])
        out = icb(imgs, gpu=True).get()
        assert np.array_equal(out, oracle.c_flip_image_rows(imgs, int(g["k"]), mode, 31, 0))
        out2 = icb(imgs, gpu=True).get()                      # next call = next trial
        assert np.array_equal(out2, oracle.c_flip_image_rows(imgs, int(g["k"]), mode, 31, 1))
    assert ImageErrorCallback(0, mode=2)(imgs) is imgs         # p == 0 returns the input (callbacks.py:150)
    with pytest.raises(ValueError):
        ImageErrorCallback(0.1)(imgs.astype(np.float32))


def test_batchnorm_layer_api(golden):
    from mlpcode.layers import BatchNormLayer
    g = golden("batchnorm")
    N = g["Z"].shape[1]
    bn = BatchNormLayer(N, gpu=True)
    bn.loadBatchNormParams(g["gamma"].copy(), g["beta"].copy(), g["mu_run"].copy(), g["sigma_run"].copy())
    bn.build()
    assert rel_err(bn.forward(g["Z"], isTrain=False).get(), g["out_eval"]) < TOL
    assert rel_err(bn.forward(g["Z"], isTrain=True).get(), g["out_train"]) < TOL
    nd = bn.backwards(g["delta"], float(g["lr"]))
    assert rel_err(nd.get(), g["new_delta"]) < TOL
    for k, v in zip(("gamma", "beta", "mu", "sigma"), bn.batchNormParams):
        assert rel_err(v.get(), g[f"{k}_after"]) < TOL, k
    assert bn.cache == {}


def test_network_api_surface(golden, tmp_path, capsys):
    from mlpcode.activation import ActivationFuncs as af
    from mlpcode.network import Network
    g = golden("train_tiny")
    units = [int(u) for u in g["units"]]
    n = len(units) - 1
    nn = build_net(units, params_from_golden(g, "init_", n))
    path = nn.save_weights("tiny", directory=tmp_path)
    nn2 = Network.fromModel(path, useGpu=True, binarized=True)
    nn2.compile(hiddenAf=af.sign, outAf=af.softmax)
    assert np.array_equal(nn2.predict(g["X_eval"]).get(), nn.predict(g["X_eval"]).get())
    assert nn2.unitList == units and nn2.useBatchNorm and not nn2.useBias
    X = g["X"].reshape(-1, units[0])
    Y = g["Y"].reshape(-1, 10)
    hist = nn.train(X, Y, epochs=2, batch_size=16, valX=g["X_eval"],
                    valY=np.eye(10, dtype=np.int8)[g["labels_eval"]])
    assert len(hist["loss"]) == 2 and len(hist["val_accuracy"]) == 2 and np.isfinite(hist["loss"]).all()
    assert "Best Validation Accuracy" in capsys.readouterr().out
    classes = nn.predict_classes(g["X_eval"])
    assert classes.shape == (g["X_eval"].shape[0],)
    # identity output (the sweep scripts compile with outAf=identity, imageBitflipTrainModelSel.py:21)
    nn3 = build_net(units, params_from_golden(g, "after2_", n), out="identity")
    logits = nn3.predict(g["X_eval"]).get()
    assert np.array_equal(logits.argmax(1), g["probs_eval"].argmax(1))
    # what stays outside this build raises, naming the reason: no batch-norm / bias (every script of the reference
    # runs batch-norm networks, which drop the bias, network.py:37-41) and the CPU path
    with pytest.raises(NotImplementedError):
        Network(units, useGpu=True, binarized=False, useBatchNorm=False).compile()
    with pytest.raises(NotImplementedError):
        Network(units, useGpu=False, binarized=True, useBatchNorm=True).compile()


def test_config2_step_properties():
    """Full-size (784-4096-4096-4096-10, B=8192) step: finite loss, deterministic, the tcgen05 and
    SIMT paths agree, and a linearity property of dW: lr -> 2*lr doubles the update."""
    import torch
    units, B = [784, 4096, 4096, 4096, 10], 8192
    rng = np.random.default_rng(0)
    import bnn_oracle as O
    params = O.new_params(units, rng)
    X = torch.from_numpy(rng.random((B, 784), dtype=np.float32) * 2 - 1).cuda()
    Y = torch.from_numpy(np.eye(10, dtype=np.int8)[rng.integers(0, 10, B)]).cuda()
    results = {}
    for tag, lr, env in (("a", 0.01, "tcgen05"), ("b", 0.01, "tcgen05"), ("c", 0.02, "tcgen05")):
        os.environ["BNN_GEMM_IMPL"] = env
        nn = build_net(units, params, lr=lr)
        cost = nn.train_step(X, Y, lr)
        results[tag] = (cost, [w.t.clone() for w in nn.weights], nn.batchNormParams[0][0].t.clone())
    os.environ.pop("BNN_GEMM_IMPL", None)
    assert np.isfinite(results["a"][0]) and abs(results["a"][0] - np.log(10)) < 1.5
    assert abs(results["a"][0] - results["b"][0]) < 1e-6
    for wa, wb in zip(results["a"][1], results["b"][1]):
        assert bool(torch.isfinite(wa).all())
        assert float((wa.double() - wb.double()).norm() / wa.double().norm()) < 1e-6   # repeatable
    # dW does not depend on lr inside one step, so the last layer's update doubles with lr (lower
    # layers' updates sit at the fp32 resolution of W and cannot be recovered by subtraction)
    i = len(units) - 2
    w0 = torch.from_numpy(params["W"][i]).cuda().double()
    da, dc = results["a"][1][i].double() - w0, results["c"][1][i].double() - w0
    assert float(dc.norm()) > 0 and float((dc - 2 * da).norm() / dc.norm()) < 1e-3


# ------------------------------------------------------------------------------------- sweeps
def _tiled_eval_set(g, reps=2):
    return np.tile(g["X"], (reps, 1)), np.tile(g["labels"], reps)


def test_weight_fault_sweep_bernoulli_matches_oracle(golden, oracle):
    from mlpcode.sweep import WeightFaultSweep
    g = golden("weight_faults")
    units = [int(u) for u in g["units"]]
    n = len(units) - 1
    params = params_from_golden(g, "", n)
    nn = build_net(units, params)
    X, labels = _tiled_eval_set(g)
    sweep = WeightFaultSweep(nn, X, labels, trials_per_launch=2)
    assert sweep.clean_accuracy() == oracle.accuracy(params, X.copy(), labels)
    p, seed, n_trials = 0.15, 20251, 5                     # 5 trials, 2 per launch: ragged last group
    got = sweep.run_bernoulli(p, n_trials, seed)
    want = []
    for t in range(n_trials):
        masks = [oracle.c_bernoulli_mask(units[i], units[i + 1], p, seed, i, t) for i in range(n)]
        hooks = [(lambda m: (lambda W: oracle.flip_bnn_mask(W, m)))(m) for m in masks]
        want.append(oracle.accuracy(params, X.copy(), labels, weight_hooks=hooks))
    assert np.array_equal(got, np.array(want))
    assert len(set(got)) > 1                                # the faults actually change the outcome
    assert np.array_equal(sweep.run_bernoulli(p, n_trials, seed), got)   # reproducible


def test_weight_fault_sweep_exactk_matches_oracle(golden, oracle):
    from mlpcode.sweep import WeightFaultSweep
    g = golden("weight_faults")
    units = [int(u) for u in g["units"]]
    n = len(units) - 1
    params = params_from_golden(g, "", n)
    nn = build_net(units, params)
    X, labels = _tiled_eval_set(g)
    sweep = WeightFaultSweep(nn, X, labels, trials_per_launch=3)
    ks = [0, 1, 40, 400, 1200]
    total = sum(units[i] * units[i + 1] for i in range(n))
    assert sweep.total_weights == total == 1200
    seed = 77
    got = sweep.run_exactk(lambda t: ks[t], len(ks), seed)
    offs = np.cumsum([0] + [units[i] * units[i + 1] for i in range(n)])
    want = []
    for t, k in enumerate(ks):
        pos = oracle.c_exactk_positions(total, seed, t, k).astype(np.int64)
        hooks = []
        for i in range(n):
            Kd, N = units[i], units[i + 1]
            mine = pos[(pos >= offs[i]) & (pos < offs[i + 1])] - offs[i]
            mask = np.zeros((Kd, N), bool)
            mask[mine % Kd, mine // Kd] = True             # e = n*K + k  ->  W[k, n]
            hooks.append((lambda m: (lambda W: oracle.flip_bnn_mask(W, m)))(mask))
        want.append(oracle.accuracy(params, X.copy(), labels, weight_hooks=hooks))
    assert np.array_equal(got, np.array(want))
    assert got[0] == sweep.clean_accuracy()                 # k = 0 is the clean network


def test_image_fault_sweep_matches_oracle(golden, oracle):
    from mlpcode.sweep import ImageFaultSweep
    g = golden("train_tiny")
    units = [int(u) for u in g["units"]]
    n = len(units) - 1
    params = params_from_golden(g, "after2_", n)
    nn = build_net(units, params, out="identity")
    rng = np.random.default_rng(8)
    imgs = rng.integers(0, 256, (40, units[0]), dtype=np.uint8)
    labels = rng.integers(0, 10, 40).astype(np.uint8)
    sweep = ImageFaultSweep(nn, imgs, labels, trials_per_launch=2)
    p, seed = 0.14, 99
    got = sweep.run(p, 3, seed)
    k = int(round(p * units[0] * 8))
    want = []
    for t in range(3):
        x = oracle.normalize(oracle.c_flip_image_rows(imgs, k, 2, seed, t), -1, 1)
        want.append(oracle.accuracy(params, x, labels, out_af="identity"))
    assert np.array_equal(got, np.array(want))


def test_data_parallel_two_gpus_equal_reference():
    """N-GPU == 1-GPU == reference (needs 2 GPUs: `gpurun --gpus 2`)."""
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from conftest import REPO
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
                          "--nproc-per-node=2", "--master-addr", "127.0.0.1", "--master-port",
                          "29611", str(REPO / "tools" / "dp_check.py")], capture_output=True, text=True,
                         timeout=300)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "dp_check ok" in out.stdout


def test_narrow_output_layer_split_k_matches_oracle(oracle):
    """A net whose 10-unit output layer takes the split-K dW route (K >= 1024, batch % 256 == 0)."""
    units, B = [64, 1024, 10], 256
    rng = np.random.default_rng(5)
    params = oracle.new_params(units, rng)
    X = rng.random((B, units[0]), dtype=np.float32) * 2 - 1
    Y = np.eye(10, dtype=np.int8)[rng.integers(0, 10, B)]
    nn = build_net(units, params)
    ref = oracle.copy_params(params)
    for _ in range(2):
        cost = nn.train_step(X, Y, 0.01)
        cost_ref = oracle.train_step(ref, X.copy(), Y, 0.01)
        assert abs(cost - cost_ref) <= TOL * abs(cost_ref)
    st = state_of(nn)
    for i in range(2):
        assert rel_err(st["W"][i], ref["W"][i]) < TOL
        upd, upd_ref = st["W"][i] - params["W"][i], ref["W"][i] - params["W"][i]
        assert rel_err(upd, upd_ref) < 2e-2, i
        for k in ("gamma", "mu", "sigma"):
            assert rel_err(st["bn"][i][k], ref["bn"][i][k]) < TOL, (i, k)


def test_train_stream_equals_stepwise_training(golden):
    """The pipelined host-batch API (copies overlapped with compute) gives the reference's losses and state."""
    g = golden("train_tiny")
    units = [int(u) for u in g["units"]]
    n = len(units) - 1
    nn = build_net(units, params_from_golden(g, "init_", n), lr=float(g["lr"]))
    steps = g["X"].shape[0]
    losses = nn.train_stream([(g["X"][s], g["Y"][s]) for s in range(steps)], float(g["lr"]))
    assert len(losses) == steps
    for s in range(steps):
        assert abs(losses[s] - g["costs"][s]) <= TOL * abs(g["costs"][s])
    st = state_of(nn)
    for i in range(n):
        assert rel_err(st["W"][i], g[f"after{steps - 1}_W{i}"]) < TOL
        for k in ("gamma", "beta", "mu", "sigma"):
            assert rel_err(st["bn"][i][k], g[f"after{steps - 1}_{k}{i}"]) < TOL
    assert nn.train_stream([], 0.01) == []


def test_int32_preactivations_for_very_wide_layers(oracle):
    """K >= 32768 inputs: z no longer fits int16, the layer switches to int32 Z."""
    units, B = [64, 33000, 10], 64
    rng = np.random.default_rng(9)
    params = oracle.new_params(units, rng)
    X = rng.random((B, units[0]), dtype=np.float32) * 2 - 1
    nn = build_net(units, params)
    got = nn.predict(X).get()
    want = oracle.predict(params, X.copy())
    assert rel_err(got, want) < TOL
    Y = np.eye(10, dtype=np.int8)[rng.integers(0, 10, B)]
    ref = oracle.copy_params(params)
    cost, cost_ref = nn.train_step(X, Y, 0.01), oracle.train_step(ref, X.copy(), Y, 0.01)
    assert abs(cost - cost_ref) <= TOL * abs(cost_ref)
    assert rel_err(nn.weights[1].get(), ref["W"][1]) < TOL


def test_step_graph_replay_equals_eager_training(oracle):
    """Network.enable_step_graph: steps 3.. are replays of one captured CUDA graph and must leave
    exactly the state of eager launches (same kernels, same arguments) — bit for bit, loss included —
    also after lr or the parameters change (re-capture)."""
    import torch
    units, B = [96, 256, 256, 10], 384       # 256-wide: the GEMM epilogues carry the BN reductions
    rng = np.random.default_rng(11)
    params = oracle.new_params(units, rng)
    batches = [(torch.from_numpy(rng.random((B, units[0]), dtype=np.float32) * 2 - 1).cuda(),
                torch.from_numpy(np.eye(10, dtype=np.int8)[rng.integers(0, 10, B)]).cuda())
               for _ in range(7)]
    runs = {}
    for mode in ("eager", "graph"):
        nn = build_net(units, params, lr=0.05)
        nn.enable_step_graph(mode == "graph")
        costs = []
        for i, (X, Y) in enumerate(batches):
            lr = 0.05 if i < 5 else 0.02          # lr change -> new capture
            costs.append(float(nn.train_step_async(X, Y, lr).mean()))
        runs[mode] = (costs, state_of(nn))
        if mode == "graph":
            assert any("graph" in st for st in nn._graphs.values())
    assert runs["eager"][0] == runs["graph"][0]
    for i in range(len(units) - 1):
        assert np.array_equal(runs["eager"][1]["W"][i], runs["graph"][1]["W"][i])
        for k in ("gamma", "beta", "mu", "sigma"):
            assert np.array_equal(runs["eager"][1]["bn"][i][k], runs["graph"][1]["bn"][i][k])
    # the pipelined host-batch API on top of graph replays (batch buffers released right after the
    # graph's input copies): same losses from pinned host batches
    nn = build_net(units, params, lr=0.05)
    nn.enable_step_graph(True)
    host = [(X.cpu().pin_memory(), Y.cpu().pin_memory()) for X, Y in batches[:5]]
    losses = nn.train_stream(host, 0.05)
    assert losses == pytest.approx(runs["eager"][0][:5], rel=1e-6)


def test_fused_epilogue_reductions_match_oracle_training(oracle, monkeypatch):
    """Widths that are multiples of the GEMM tile take the fused path (column sums in the forward
    epilogue, batch-norm backward reductions in the dX epilogue).  One full training step from seeded
    parameters: loss and state within 1e-5 of the oracle, and the fused and stand-alone reductions
    agree with each other to the same tolerance.  (One step only: binarized training is chaotic, a
    single borderline sign decision changes later steps discretely; multi-step parity is pinned on the
    reference's own fixtures in test_training_steps_match_reference.)"""
    from mlpcode import kernels as K
    units, B = [64, 512, 256, 10], 300       # ragged batch, fusable widths
    assert K.fused_stats_ok(512) and K.fused_stats_ok(256)
    rng = np.random.default_rng(5)
    params = oracle.new_params(units, rng)
    X = rng.random((B, units[0]), dtype=np.float32) * 2 - 1
    Y = np.eye(10, dtype=np.int8)[rng.integers(0, 10, B)]
    ref = oracle.copy_params(params)
    ref_cost = oracle.train_step(ref, X, Y, 0.05)
    states = {}
    errs = {}
    for fused in ("1", "0"):
        monkeypatch.setenv("BNN_FUSE_STATS", fused)
        nn = build_net(units, params, lr=0.05)
        cost = nn.train_step(X, Y, 0.05)
        states[fused] = state_of(nn)
        assert abs(cost - ref_cost) <= TOL * abs(ref_cost), (fused, cost, ref_cost)
        for i in range(len(units) - 1):
            errs[(fused, i, "W")] = rel_err(states[fused]["W"][i], ref["W"][i])
            for k in ("gamma", "beta", "mu", "sigma"):
                errs[(fused, i, k)] = rel_err(states[fused]["bn"][i][k], ref["bn"][i][k])
            if i < len(units) - 2:
                # hidden layers: dbeta = sum_b delta = (sum_b delta'_above) . W^T, and batch-norm
                # backward makes sum_b delta' vanish identically — both sides hold rounding noise
                # only, so it is compared on an absolute floor (|beta| stays below 1e-6 here)
                diff = np.abs(states[fused]["bn"][i]["beta"] - ref["bn"][i]["beta"]).max()
                assert np.abs(ref["bn"][i]["beta"]).max() < 1e-6 and diff < 1e-6, (fused, i, diff)
                errs.pop((fused, i, "beta"))
    bad = {k: v for k, v in errs.items() if not v < TOL}
    assert not bad, bad
    for i in range(len(units) - 1):
        assert rel_err(states["1"]["W"][i], states["0"]["W"][i]) < TOL
        for k in ("gamma", "mu", "sigma") + (("beta",) if i == len(units) - 2 else ()):
            assert rel_err(states["1"]["bn"][i][k], states["0"]["bn"][i][k]) < TOL


def test_dw_exchange_kernels_on_one_device():
    """The data-parallel dW path — reduce-scatter in the GEMM epilogue (BNN_EPI_SCATTER_F32), owner
    update, all-gather (kernel stores and the copy-engine form) — with both 'ranks' living on one GPU:
    the kernels only see pointers, so two sets of buffers on the same device exercise the protocol
    (flags, epochs, row ownership, tile rotation) without a second GPU.
    Expected: both copies of W equal W0 - lr * (dW_rank0 + dW_rank1) (layers.py:260,268)."""
    import torch
    from mlpcode import _native as nat, kernels as K
    T = torch
    world, Kd, N, Bt, lr = 2, 300, 200, 256, 0.05       # Kd: 3 row tiles -> rows_per_owner 256
    rpo = (((Kd + 127) // 128 + world - 1) // world) * 128
    assert rpo == 256
    rng = np.random.default_rng(3)
    W0 = rng.standard_normal((Kd, N)).astype(np.float32)
    Wc = [T.from_numpy(W0).cuda().clone() for _ in range(world)]
    stage = [T.zeros((world, rpo, N), dtype=T.float32, device="cuda") for _ in range(world)]
    flags = [T.zeros((4, 16), dtype=T.int32, device="cuda") for _ in range(world)]
    counters = [T.zeros(1, dtype=T.int32, device="cuda") for _ in range(world)]
    step = T.zeros(1, dtype=T.int32, device="cuda")
    one = T.ones(1, device="cuda")
    dWs = []
    for epoch in (1, 2):                                  # two steps: epochs advance, buffers are reused
        K.step_tick(step)
        parts = []
        for r in range(world):
            a = (rng.standard_normal((Kd, Bt)) / 8).astype(np.float16)
            b = (rng.standard_normal((N, Bt)) / 8).astype(np.float16)
            A, Bm = T.from_numpy(a).cuda(), T.from_numpy(b).cuda()
            parts.append(a.astype(np.float64) @ b.astype(np.float64).T)
            K.gemm(M=Kd, N=N, K=Bt, A=(A,), B=(Bm,), lda=Bt, ldb=Bt, D=None, ldd=N,
                   in_dtype=nat.DT_F16, epilogue=nat.EPI_SCATTER_F32, sa=one, sb=one,
                   scatter=dict(world=world, rank=r, rows_per_owner=rpo,
                                stage=[s.data_ptr() for s in stage]))
            K.peer_signal([f.data_ptr() for f in flags], 0, r, world, step)
        dWs.append(sum(parts))
        for r in range(world):
            if epoch == 1:   # all-gather by stores from the owner-update kernel
                K.dw_reduce_sgd_p2p(stage[r].data_ptr(), [w.data_ptr() for w in Wc],
                                    [f.data_ptr() for f in flags], 0, 1, r, world, step, Kd, N, rpo,
                                    lr, True, counters[r])
            else:            # owner update locally, then the copy-engine all-gather
                K.dw_reduce_sgd_p2p(stage[r].data_ptr(), [w.data_ptr() for w in Wc],
                                    [f.data_ptr() for f in flags], 0, 1, r, world, step, Kd, N, rpo,
                                    lr, False, counters[r])
                K.peer_broadcast_rows([w.data_ptr() for w in Wc], [f.data_ptr() for f in flags], 1, r,
                                      world, Kd, N, rpo, step)
        for r in range(world):
            K.peer_wait(flags[r].data_ptr(), 1, 2, 1, r, world, step)
        T.cuda.synchronize()
        want = W0.astype(np.float64) - lr * sum(dWs)
        for r in range(world):
            got = Wc[r].cpu().numpy()
            assert rel_err(got - W0, want - W0) < 1e-5, (epoch, r)
        assert T.equal(Wc[0], Wc[1])
        assert [int(f[0, 0]) for f in flags] == [epoch] * world and int(flags[0][1, 1]) == epoch
