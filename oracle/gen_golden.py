"""Generate tests/golden/*.npz by running the UNMODIFIED reference (volf52/deep-neural-net).

Run in the build container only (`python oracle/gen_golden.py`): it needs /root/reference, which does
not exist on the GPU box.  The reference is imported from a throw-away copy (importing
mlpcode.network creates a models/ directory, mlpcode/utils.py:19-21) with oracle/refshim first on
sys.path (cupy -> numpy stand-in, empty h5py).  For every fixture the NumPy oracle
(oracle/bnn_oracle.py) is run on the same inputs and must agree EXACTLY with the reference before
the reference's outputs are written — that is the pin of the oracle.
"""
from __future__ import annotations

import shutil
import sys
import tempfile
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
REPO = HERE.parent
GOLDEN = REPO / "tests" / "golden"
REFERENCE = Path("/root/reference")


def import_reference():
    tmp = Path(tempfile.mkdtemp(prefix="bnn_ref_"))
    shutil.copytree(REFERENCE, tmp / "ref")
    sys.path.insert(0, str(tmp / "ref"))
    sys.path.insert(0, str(HERE / "refshim"))
    import mlpcode.layers as L  # noqa: E402
    import mlpcode.activation as A  # noqa: E402
    import mlpcode.loss as Lo  # noqa: E402
    import mlpcode.callbacks as Cb  # noqa: E402
    import mlpcode.network as Nw  # noqa: E402
    import mlpcode.utils as U  # noqa: E402
    return L, A, Lo, Cb, Nw, U


def exact(a, b, what):
    a, b = np.asarray(a), np.asarray(b)
    if a.shape != b.shape or not np.array_equal(a, b, equal_nan=True):
        raise SystemExit(f"oracle != reference for {what}: max|diff| = "
                         f"{np.max(np.abs(a.astype(np.float64) - b.astype(np.float64)))}")


def make_net(Nw, A, Lo, units, params, out_af):
    nn = Nw.Network(units, useGpu=False, binarized=True, useBatchNorm=True)
    nn.compile(lr=0.01, hiddenAf=A.ActivationFuncs.sign, outAf=out_af,
               lossF=Lo.LossFuncs.cross_entropy)
    nn.load_weights([w.copy() for w in params["W"]])
    nn.loadBatchNormParameters([b["gamma"].copy() for b in params["bn"]],
                               [b["beta"].copy() for b in params["bn"]],
                               [b["mu"].copy() for b in params["bn"]],
                               [b["sigma"].copy() for b in params["bn"]])
    return nn


def net_state(nn):
    g, b, m, s = nn.batchNormParams
    return dict(W=[w.copy() for w in nn.weights],
                bn=[dict(gamma=g[i].copy(), beta=b[i].copy(), mu=m[i].copy(), sigma=s[i].copy())
                    for i in range(len(g))])


def flat_state(prefix, st):
    out = {}
    for i, w in enumerate(st["W"]):
        out[f"{prefix}W{i}"] = w
        for k, v in st["bn"][i].items():
            out[f"{prefix}{k}{i}"] = v
    return out


def gen_checkpoint(A, Lo, Nw, O):
    """tests/golden/checkpoint_ref.npz is WRITTEN BY THE REFERENCE: its unmodified Network.save_weights
    (network.py:428-472) runs against the npz-backed h5py stand-in (oracle/refshim/h5py.py) after two training
    steps on a small network.  The reference's fromModel (:157-228) then has to read back both that file and the
    one this package's writer (mlpcode/checkpoint.py, loaded by path: both packages are called `mlpcode`)
    produces from the same state — predictions identical in both cases.  tests/test_checkpoint.py holds this
    package's reader and writer against the reference-written file."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("b200_checkpoint",
                                                  REPO / "deep-neural-net_b200" / "mlpcode" / "checkpoint.py")
    ck = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ck)
    af = A.ActivationFuncs
    rng = np.random.default_rng(63)
    units = [40, 24, 12, 10]
    params = O.new_params(units, rng)
    X = rng.random((32, 40), dtype=np.float32) * 2 - 1
    Y = np.eye(10, dtype=np.int8)[rng.integers(0, 10, 32)]
    nn = make_net(Nw, A, Lo, units, params, af.softmax)
    for _ in range(2):
        nn._Network__updateBatch((X.copy(), Y.copy()), 0.01)
    probs = nn.predict(X.copy(), isTraining=False)
    before = set(Nw.MODELDIR.glob("*.hdf5"))
    nn.save_weights("golden_ckpt")
    written, = set(Nw.MODELDIR.glob("*.hdf5")) - before
    # the reference reads its own file ...
    back = Nw.Network.fromModel(written, useGpu=False, binarized=True)
    back.compile(hiddenAf=af.sign, outAf=af.softmax)
    exact(back.predict(X.copy(), isTraining=False), probs, "reference fromModel(reference save_weights)")
    # ... and this package's writer produces a file the reference reads to the same effect
    st = net_state(nn)
    bn = [[b[k] for b in st["bn"]] for k in ("gamma", "beta", "mu", "sigma")]
    ours = ck.write_checkpoint(Nw.MODELDIR / "ours.hdf5", units, False, True, st["W"], None, bn, packed=True)
    back2 = Nw.Network.fromModel(ours, useGpu=False, binarized=True)
    back2.compile(hiddenAf=af.sign, outAf=af.softmax)
    exact(back2.predict(X.copy(), isTraining=False), probs, "reference fromModel(this package's writer)")
    # packed-only file: latent weights of +-1 give the same eval-mode result (sign(W) is all that is read)
    only = ck.read_checkpoint(ck.write_checkpoint(Nw.MODELDIR / "packed.npz", units, False, True, st["W"], None, bn,
                                                  packed=True, latent=False))
    assert only["packed_only"] and all(np.array_equal(w, O.binarize(v)) for w, v in zip(only["weights"], st["W"]))
    shutil.copyfile(written, GOLDEN / "checkpoint_ref.npz")
    np.savez_compressed(GOLDEN / "checkpoint_ref_io.npz", X=X, probs=probs, **flat_state("", st))


def gen_epochs(A, Lo, Nw, O):
    """tests/golden/train_epochs.npz: the reference's own Network.train (network.py:264-366) — three epochs of two
    mini-batches with the cumulative shuffle of :312-315 drawn from np.random.seed(s), end-of-epoch accuracies,
    best-parameter snapshot — and the oracle's train_epochs replaying the same permutations, which must agree
    EXACTLY (history and final state).  Inputs sit on the k/16 grid and the seed is the first one for which no
    sign decision of any step lies within 1e-6 of zero (tests/test_config_parity_gpu.py explains why a
    free-running comparison needs that), so the CUDA path can be held to the file step for step."""
    af = A.ActivationFuncs
    units, n, bs, epochs, lr = [30, 16, 12, 10], 400, 200, 3, 0.01
    for seed in range(200):
        rng = np.random.default_rng(1000 + seed)
        params = O.new_params(units, rng)
        X = (rng.integers(-16, 17, (n, units[0])) / 16.0).astype(np.float32)
        Y = np.eye(10, dtype=np.int8)[rng.integers(0, 10, n)]
        Xv = (rng.integers(-16, 17, (100, units[0])) / 16.0).astype(np.float32)
        yv = rng.integers(0, 10, 100).astype(np.uint8)
        np.random.seed(seed)
        perms = [np.random.permutation(n) for _ in range(epochs)]
        amb, first = [0], [True]

        def on_step(rec):
            for cache in rec["caches"][:-1]:
                bn = cache["z"]
                a = np.abs(bn) < 1e-6
                if first[0]:
                    a &= bn != 0
                amb[0] += int(a.sum())
            first[0] = False
        probe = O.copy_params(params)
        hist_o = O.train_epochs(probe, X, Y, lr, epochs, bs, perms=perms, on_step=on_step)
        if amb[0] != 0:
            continue
        nn = make_net(Nw, A, Lo, units, params, af.softmax)
        np.random.seed(seed)
        hist = nn.train(X.copy(), Y.copy(), epochs, batch_size=bs, shuffle=True, valX=Xv.copy(), valY=yv.copy(),
                        save_best_params=False)
        if int(np.argmax(hist["val_accuracy"])) < epochs - 1:   # the restored snapshot differs from the final state
            break
    else:
        raise SystemExit("no well-conditioned seed for the epoch fixture")
    exact(np.array(hist["loss"]), np.array(hist_o["loss"]), "epoch losses")
    exact(np.array(hist["accuracy"]), np.array(hist_o["accuracy"]), "epoch train accuracies")
    st = net_state(nn)
    for i in range(len(units) - 1):
        exact(st["W"][i], probe["W"][i], f"W{i} after train()")
        for k in ("gamma", "beta", "mu", "sigma"):
            exact(st["bn"][i][k], probe["bn"][i][k], f"{k}{i} after train()")
    val_acc = np.array(hist["val_accuracy"])
    # default flow (save_best_params=True): the best-validation snapshot is restored at the end (:344-366)
    nn2 = make_net(Nw, A, Lo, units, params, af.softmax)
    np.random.seed(seed)
    nn2.train(X.copy(), Y.copy(), epochs, batch_size=bs, shuffle=True, valX=Xv.copy(), valY=yv.copy())
    blob = dict(units=np.array(units), X=X, Y=Y, Xv=Xv, yv=yv, perms=np.stack(perms), lr=np.float64(lr),
                batch_size=np.int64(bs), loss=np.array(hist["loss"]), accuracy=np.array(hist["accuracy"]),
                val_accuracy=val_acc, seed=np.int64(seed))
    blob.update(flat_state("init_", params))
    blob.update(flat_state("final_", st))
    blob.update(flat_state("best_", net_state(nn2)))
    np.savez_compressed(GOLDEN / "train_epochs.npz", **blob)
    print(f"train_epochs.npz: seed {seed}, loss {hist['loss']}, acc {hist['accuracy']}, val {list(val_acc)}")


def gen_act_loss(A, Lo, Nw, O):
    """tests/golden/act_loss.npz: every activation, activation derivative, loss and loss derivative of the
    reference (activation.py:8-168, loss.py:10-156) on seeded arrays with edge values, and one training step plus a
    prediction of a small binarized network for each output-activation / loss pairing the reference treats
    differently (network.py:377-387: fused `ypred - y` for softmax / sigmoid + cross-entropy, activation derivative
    otherwise).  Outputs are the reference's own; there is no oracle in between."""
    af, lf = A.ActivationFuncs, Lo.LossFuncs
    rng = np.random.default_rng(71)
    x = (rng.standard_normal((37, 10)) * np.exp(rng.standard_normal((37, 10)))).astype(np.float32)
    x[0, :8] = [0.0, -0.0, 1.0, -1.0, 1.5, -1.5, 30.0, -30.0]
    dA = rng.standard_normal((37, 10)).astype(np.float32)
    blob = dict(x=x, dA=dA)
    fwd = dict(A.ACTIVATION_FUNCTIONS)
    fwd_extra = {"hard_sigmoid": A.hard_sigmoid, "hard_tanh": A.hard_tanh}
    for k, f in list(fwd.items()) + list(fwd_extra.items()):
        name = k.name if hasattr(k, "name") else k
        a = f(x.copy())
        blob[f"act_{name}"] = np.asarray(a, np.float32)
    for k, f in A.ACTIVATION_DERIVATIVES.items():
        a = fwd[k](x.copy())
        if k == af.sign:            # hard_tanh_derivative is also meaningful off the +-1 lattice
            a = x.copy()
        blob[f"grad_in_{k.name}"] = np.asarray(a, np.float32)
        blob[f"grad_{k.name}"] = np.asarray(f(dA.copy(), a.copy()), np.float32)
    R, C = 29, 10
    p = rng.random((R, C)).astype(np.float32)
    p /= p.sum(axis=1, keepdims=True)
    p[0, 0], p[1, 1] = 0.0, 1e-20                       # exercised by the lower clip
    y = np.eye(C, dtype=np.int8)[rng.integers(0, C, R)]
    ypm = np.where(y == 0, -1, 1).astype(np.int8)
    t = (rng.standard_normal((R, C)) * 1.5).astype(np.float32)      # tanh-like scores for hinge
    blob.update(loss_p=p, loss_y=y, loss_ypm=ypm, loss_t=t)
    blob["ce_rows"] = Lo.cross_entropy_loss(p.copy(), y).astype(np.float32)
    q = p.copy()
    blob["ce_grad_fused"] = Lo.cross_entropy_derivative(q, y, with_softmax=True).astype(np.float32)
    q = np.clip(p, 1e-6, 1 - 1e-6).astype(np.float32)               # away from 1: the general form divides by 1 - p
    blob["ce_general_p"] = q.copy()
    blob["ce_grad_general"] = Lo.cross_entropy_derivative(q, y, with_softmax=False).astype(np.float32)
    blob["mse_rows"] = Lo.mse(p.copy(), y).astype(np.float32)
    blob["mse_grad"] = Lo.mse_derivative(p.copy(), y).astype(np.float32)
    blob["hinge_rows"] = Lo.hinge_loss(t.copy(), ypm).astype(np.float32)
    blob["hinge_grad"] = np.asarray(Lo.hinge_derivative(t.copy(), ypm), np.float64)
    # one training step + a prediction per pairing, on grid inputs (layer-0 sums exact, beta = 0: every sign
    # decision of the step is determined; tests/test_config_parity_gpu.py explains)
    units, B = [30, 16, 12, 10], 64
    params = O.new_params(units, rng)
    X = (rng.integers(-16, 17, (B, units[0])) / 16.0).astype(np.float32)
    X2 = (rng.integers(-16, 17, (B, units[0])) / 16.0).astype(np.float32)
    Y = np.eye(10, dtype=np.int8)[rng.integers(0, 10, B)]
    blob.update(net_units=np.array(units), net_X=X, net_X2=X2, net_Y=Y)
    blob.update(flat_state("net_init_", params))
    pairs = [("sigmoid", "mse"), ("sigmoid", "cross_entropy"), ("tanh", "hinge"), ("softmax", "mse"),
             ("relu", "mse"), ("leaky_relu", "mse"), ("identity", "mse"), ("softmax", "hinge")]
    blob["net_pairs"] = np.array(["+".join(pr) for pr in pairs])
    for out_af, loss in pairs:
        nn = Nw.Network(units, useGpu=False, binarized=True, useBatchNorm=True)
        nn.compile(lr=0.01, hiddenAf=af.sign, outAf=af[out_af], lossF=lf[loss])
        nn.load_weights([w.copy() for w in params["W"]])
        Yt = np.where(Y == 0, -1, Y).astype(np.int8) if loss == "hinge" else Y.copy()   # network.py:301-303
        cost = nn._Network__updateBatch((X.copy(), Yt), 0.01)
        tag = f"net_{out_af}_{loss}_"
        blob[tag + "cost"] = np.float64(cost)
        blob.update(flat_state(tag, net_state(nn)))
        blob[tag + "pred"] = np.asarray(nn.predict(X2.copy(), isTraining=False), np.float32)
    np.savez_compressed(GOLDEN / "act_loss.npz", **blob)
    print("act_loss.npz:", {k: float(blob[f"net_{a}_{l}_cost"]) for a, l in pairs for k in [f"{a}+{l}"]})


def gen_linear(A, Lo, Nw, O):
    """tests/golden/linear_net.npz: the reference's real-valued network (binarized=False: LinearLayer,
    layers.py:130-271, batch-norm, no bias) trained three steps on three batches for each smooth hidden activation,
    plus its eval-mode prediction.  No sign() anywhere, so the comparison is well-conditioned over several steps."""
    af, lf = A.ActivationFuncs, Lo.LossFuncs
    rng = np.random.default_rng(83)
    units, B = [30, 24, 16, 10], 48
    params = O.new_params(units, rng)
    batches = [((rng.random((B, units[0]), dtype=np.float32) * 2 - 1),
                np.eye(10, dtype=np.int8)[rng.integers(0, 10, B)]) for _ in range(3)]
    X2 = rng.random((B, units[0]), dtype=np.float32) * 2 - 1
    blob = dict(units=np.array(units), X_eval=X2)
    for i, (X, Y) in enumerate(batches):
        blob[f"X{i}"], blob[f"Y{i}"] = X, Y
    blob.update(flat_state("init_", params))
    hidden = ["sigmoid", "relu", "tanh", "leaky_relu"]
    blob["hidden"] = np.array(hidden)
    for h in hidden:
        nn = Nw.Network(units, useGpu=False, binarized=False, useBatchNorm=True)
        nn.compile(lr=0.05, hiddenAf=af[h], outAf=af.softmax, lossF=lf.cross_entropy)
        nn.load_weights([w.copy() for w in params["W"]])
        for i, (X, Y) in enumerate(batches):
            blob[f"{h}_cost{i}"] = np.float64(nn._Network__updateBatch((X.copy(), Y.copy()), 0.05))
            blob.update(flat_state(f"{h}_s{i}_", net_state(nn)))
        blob[f"{h}_pred"] = np.asarray(nn.predict(X2.copy(), isTraining=False), np.float32)
    np.savez_compressed(GOLDEN / "linear_net.npz", **blob)
    print("linear_net.npz:", {h: [float(blob[f"{h}_cost{i}"]) for i in range(3)] for h in hidden})


def gen_float_faults(Cb, O):
    """IEEE-754 bit flips of real-valued weights (ErrorCallback(bnn=False), callbacks.py:35-59, 85-123):
    the reference draws from the global NumPy stream; replaying the same stream through the oracle's
    explicit-draw restatement must give identical bits."""
    rng = np.random.default_rng(57)
    W = (rng.standard_normal((12, 9)) * np.exp(rng.standard_normal((12, 9)) * 3)).astype(np.float32)
    W[0, :4] = [1.0, -1.0, 0.0, 3.0]                 # +-1 (a binarized weight), zero, exponent MSB set
    W[1, 0], W[1, 1] = np.float32(2.0 ** 64), np.float32(2.0 ** -100)
    blob = dict(W=W)
    for mode in (0, 1, 2):
        for p in (0.1, 0.5):
            cb = Cb.ErrorCallback(p, mode=mode, bnn=False)
            np.random.seed(300 + mode)
            out = cb(W.copy())
            np.random.seed(300 + mode)
            elem = np.random.choice(W.size, round(W.size * p), replace=False)
            mine = O.flip_float_bits(W, elem, p, mode, lambda c, k: np.random.choice(c, k, replace=False))
            exact(mine.view(np.uint32), np.asarray(out, np.float32).view(np.uint32),
                  f"ErrorCallback float flips mode {mode} p {p}")
            tag = f"m{mode}_p{int(p * 10)}"
            blob[f"out_{tag}"] = np.asarray(out, np.float32)
            blob[f"elem_{tag}"] = elem.astype(np.int64)
    np.savez_compressed(GOLDEN / "float_faults.npz", **blob)


def main():
    if sys.argv[1:] in (["checkpoint"], ["epochs"], ["act_loss"], ["linear"]):      # only this fixture (the others stay byte-identical)
        L, A, Lo, Cb, Nw, U = import_reference()
        sys.path.insert(0, str(HERE))
        import bnn_oracle as O
        {"checkpoint": gen_checkpoint, "epochs": gen_epochs, "act_loss": gen_act_loss, "linear": gen_linear}[sys.argv[1]](A, Lo, Nw, O)
        print(f"{sys.argv[1]} fixtures written")
        return
    _main_all()


def _main_all():
    sys.path.insert(0, str(HERE))
    if "--only-float-faults" in sys.argv:            # leaves the other fixtures byte-identical
        import bnn_oracle as O
        _L, _A, _Lo, Cb, _Nw, _U = import_reference()
        gen_float_faults(Cb, O)
        print("float_faults.npz written")
        return
    L, A, Lo, Cb, Nw, U = import_reference()
    import bnn_oracle as O

    GOLDEN.mkdir(parents=True, exist_ok=True)
    af = A.ActivationFuncs

    # ------------------------------------------------------------------ elementwise ops (A1, A6, A7)
    rng = np.random.default_rng(11)
    x = rng.standard_normal((9, 37), dtype=np.float32)
    x[0, :4] = [0.0, -0.0, 1e-30, -1e-30]
    ref_bin = L.BinaryLayer.binarize(x)
    exact(O.binarize(x), ref_bin, "binarize")
    ref_step = A.unitstep(x)
    exact(O.unitstep(x), ref_step, "unitstep")
    dA = rng.standard_normal((9, 37), dtype=np.float32)
    ste = A.hard_tanh_derivative(dA, ref_step)
    exact(dA, ste, "STE identity on +-1 activations")
    logits = rng.standard_normal((13, 10), dtype=np.float32) * 3
    ref_sm = A.softmax(logits)
    exact(O.softmax(logits), ref_sm, "softmax")
    y = np.eye(10, dtype=np.int8)[rng.integers(0, 10, 13)]
    sm1, sm2 = ref_sm.copy(), ref_sm.copy()
    sm1[0, 0] = 0.0  # exercise the lower clip
    sm2[0, 0] = 0.0
    ref_loss = Lo.cross_entropy_loss(sm1, y)
    exact(O.cross_entropy_loss(sm2, y), ref_loss, "cross_entropy_loss")
    exact(sm1, sm2, "in-place clip")
    ref_grad = Lo.cross_entropy_derivative(sm1.copy(), y, with_softmax=True)
    exact(O.cross_entropy_derivative_softmax(sm2.copy(), y), ref_grad, "cross_entropy_derivative")
    np.savez(GOLDEN / "elementwise.npz", x=x, binarized=ref_bin, unitstep=ref_step, dA=dA, ste=ste,
             logits=logits, softmax=ref_sm, onehot=y, clipped=sm1, loss=ref_loss, grad=ref_grad)

    # ------------------------------------------------------------------ batch norm (A4, A5, A8)
    rng = np.random.default_rng(12)
    B, N = 48, 21
    Zi = rng.integers(-30, 31, (B, N)).astype(np.float32) * 2  # integer valued like a +-1 GEMM
    gamma = (rng.random(N, dtype=np.float32) + 0.5)
    gamma[3] = -gamma[3]
    beta = rng.standard_normal(N, dtype=np.float32) * 0.1
    mu_run = rng.standard_normal(N, dtype=np.float32)
    sig_run = rng.random(N, dtype=np.float32) + 0.5
    bn = L.BatchNormLayer(N)
    bn.loadBatchNormParams(gamma.copy(), beta.copy(), mu_run.copy(), sig_run.copy())
    bn.build()
    out_tr = bn.forward(Zi.copy(), isTrain=True)
    o_out, o_cache = O.bn_forward_train(Zi.copy(), gamma.copy(), beta.copy())
    exact(o_out, out_tr, "bn train forward")
    delta = rng.standard_normal((B, N), dtype=np.float32) * 1e-3
    batch_mu, batch_var = bn.cache["mu"].copy(), bn.cache["var"].copy()
    new_delta = bn.backwards(delta.copy(), lr=0.05)
    op = dict(gamma=gamma.copy(), beta=beta.copy(), mu=mu_run.copy(), sigma=sig_run.copy())
    o_nd, o_dg, o_db = O.bn_backward(delta.copy(), o_cache, op, 0.05)
    exact(o_nd, new_delta, "bn backward delta")
    for k, v in zip(("gamma", "beta", "mu", "sigma"), bn.batchNormParams):
        exact(op[k], v, f"bn backward {k}")
    bn2 = L.BatchNormLayer(N)
    bn2.loadBatchNormParams(gamma.copy(), beta.copy(), mu_run.copy(), sig_run.copy())
    bn2.build()
    out_ev = bn2.forward(Zi.copy(), isTrain=False)
    exact(O.bn_forward_eval(Zi.copy(), gamma, beta, mu_run, sig_run), out_ev, "bn eval forward")
    np.savez(GOLDEN / "batchnorm.npz", Z=Zi, gamma=gamma, beta=beta, mu_run=mu_run, sigma_run=sig_run,
             out_train=out_tr, batch_mu=batch_mu, batch_var=batch_var, delta=delta, lr=np.float32(0.05),
             new_delta=new_delta, dgamma=o_dg, dbeta=o_db,
             gamma_after=bn.gamma, beta_after=bn.beta, mu_after=bn.mu, sigma_after=bn.sigma,
             out_eval=out_ev)

    # ------------------------------------------------------------------ training steps (A2, A9, A10, A13)
    for name, units, B, steps, seed in (("train_tiny", [20, 16, 12, 10], 16, 3, 21),
                                        ("train_small", [784, 64, 32, 10], 32, 2, 22),
                                        ("train_ragged", [50, 33, 10], 19, 2, 23)):
        rng = np.random.default_rng(seed)
        params = O.new_params(units, rng)
        for b in params["bn"]:  # non-trivial BN state
            b["gamma"] += rng.standard_normal(b["gamma"].shape, dtype=np.float32) * 0.1
            b["beta"] += rng.standard_normal(b["beta"].shape, dtype=np.float32) * 0.1
        X = rng.random((steps, B, units[0]), dtype=np.float32) * 2 - 1
        Y = np.eye(10, dtype=np.int8)[rng.integers(0, 10, (steps, B))]
        nn = make_net(Nw, A, Lo, units, params, af.softmax)
        op = O.copy_params(params)
        blob = dict(units=np.array(units), X=X, Y=Y, lr=np.float32(0.01))
        blob.update(flat_state("init_", params))
        costs = []
        for s in range(steps):
            # stage tensors of the first step for teacher-forced parity
            if s == 0:
                acts, a = [], X[0]
                for layer in nn._layers:
                    a = layer.forward(a, isTrain=True)
                    acts.append((layer.cache["z"].copy(), np.asarray(a).copy()))
                for i, (z, a_) in enumerate(acts):
                    blob[f"s0_bn_out{i}"], blob[f"s0_act{i}"] = z, a_
                rec = {}
                o_tmp = O.copy_params(op)
                O.train_step(o_tmp, X[0].copy(), Y[0], 0.01, record=rec)
                for i in range(len(units) - 1):
                    blob[f"s0_dW{i}"] = rec["dW"][i]
                    blob[f"s0_dX{i}"] = rec["dX"][i]
                    blob[f"s0_delta_bn{i}"] = rec["delta_bn"][i]
                    blob[f"s0_z{i}"] = rec["caches"][i]["z_pre"]
                blob["s0_dLdA"] = rec["dLdA"]
            c_ref = nn._Network__updateBatch((X[s].copy(), Y[s]), 0.01)
            c_or = O.train_step(op, X[s].copy(), Y[s], 0.01)
            exact(c_or, c_ref, f"{name} cost step {s}")
            st = net_state(nn)
            for i in range(len(units) - 1):
                exact(op["W"][i], st["W"][i], f"{name} W{i} step {s}")
                for k in ("gamma", "beta", "mu", "sigma"):
                    exact(op["bn"][i][k], st["bn"][i][k], f"{name} {k}{i} step {s}")
            costs.append(c_ref)
            blob.update(flat_state(f"after{s}_", st))
        blob["costs"] = np.array(costs, np.float64)
        # eval-mode predictions with the trained state (running statistics)
        Xe = rng.random((B + 5, units[0]), dtype=np.float32) * 2 - 1
        pe = nn.predict(Xe.copy(), isTraining=False)
        exact(O.predict(op, Xe.copy()), pe, f"{name} eval predict")
        blob["X_eval"], blob["probs_eval"] = Xe, pe
        labels = rng.integers(0, 10, B + 5).astype(np.uint8)
        blob["labels_eval"] = labels
        blob["acc_eval"] = np.float64(nn.get_accuracy(Xe.copy(), labels))
        assert O.accuracy(op, Xe.copy(), labels) == float(blob["acc_eval"])
        np.savez_compressed(GOLDEN / f"{name}.npz", **blob)

    # ------------------------------------------------------------------ weight faults (A3, A11)
    rng = np.random.default_rng(31)
    units = [40, 24, 10]
    params = O.new_params(units, rng)
    Xf = rng.random((30, 40), dtype=np.float32) * 2 - 1
    labels = rng.integers(0, 10, 30).astype(np.uint8)
    nn = make_net(Nw, A, Lo, units, params, af.softmax)
    Wb = L.BinaryLayer.binarize(params["W"][0])
    cb = Cb.ErrorCallback(0.2, mode=2, bnn=True)
    np.random.seed(77)
    flipped = cb(Wb)
    np.random.seed(77)
    mask0 = np.random.uniform(size=Wb.shape) < 0.2
    exact(O.flip_bnn_mask(Wb, mask0), flipped, "ErrorCallback._flipBNNMask")
    # full eval with a callback on every layer: one fresh mask per layer per forward, drawn in order
    nn.addCallbacks(Cb.ErrorCallback(0.1, mode=2, bnn=True))
    np.random.seed(78)
    probs_fault = nn.predict(Xf.copy(), isTraining=False)
    np.random.seed(78)
    masks = [np.random.uniform(size=w.shape) < 0.1 for w in params["W"]]
    hooks = [(lambda m: (lambda Wq: O.flip_bnn_mask(Wq, m)))(m) for m in masks]
    exact(O.predict(params, Xf.copy(), weight_hooks=hooks), probs_fault, "faulty predict")
    np.random.seed(78)
    acc_fault = nn.get_accuracy(Xf.copy(), labels)
    nn.clearCallbacks()
    blob = dict(units=np.array(units), X=Xf, labels=labels, cb_Wb=Wb, cb_mask=mask0, cb_flipped=flipped,
                probs_fault=probs_fault, acc_fault=np.float64(acc_fault),
                probs_clean=nn.predict(Xf.copy(), isTraining=False))
    for i, m in enumerate(masks):
        blob[f"mask{i}"] = m
    blob.update(flat_state("", params))
    np.savez_compressed(GOLDEN / "weight_faults.npz", **blob)

    # ------------------------------------------------------------------ image faults (A12) + normalize
    rng = np.random.default_rng(41)
    imgs = rng.integers(0, 256, (6, 49), dtype=np.uint8)
    imgs[0] = 0
    imgs[1] = 255
    blob = dict(images=imgs)
    for mode in (0, 1, 2):
        icb = Cb.ImageErrorCallback(0.14, mode=mode)
        np.random.seed(90 + mode)
        out = icb(imgs)
        # replay the reference's draws through the oracle's explicit-index flip
        np.random.seed(90 + mode)
        mine = np.empty_like(imgs)
        for r in range(imgs.shape[0]):
            unpacked = np.unpackbits(imgs[r])
            nbits = round(0.14 * unpacked.size)
            if mode == 2:
                idx = np.random.choice(unpacked.size, nbits, replace=False)
            else:
                chosen, = np.where(unpacked == mode)
                idx = np.random.choice(chosen, min(chosen.size, nbits), replace=False)
            mine[r] = O.flip_image_row(imgs[r], idx)
        exact(mine, out, f"ImageErrorCallback mode {mode}")
        blob[f"out_mode{mode}"] = out
    blob["k"] = np.int64(round(0.14 * 49 * 8))
    norm = U.normalize(imgs, newMin=-1, newMax=1)
    exact(O.normalize(imgs, -1, 1), norm.astype(np.float32), "normalize")
    blob["normalized"] = norm.astype(np.float32)
    np.savez_compressed(GOLDEN / "image_faults.npz", **blob)

    gen_float_faults(Cb, O)
    gen_checkpoint(A, Lo, Nw, O)
    gen_epochs(A, Lo, Nw, O)
    gen_act_loss(A, Lo, Nw, O)
    gen_linear(A, Lo, Nw, O)

    sizes = {p.name: p.stat().st_size for p in sorted(GOLDEN.glob("*.npz"))}
    print("golden fixtures written:", sizes)


if __name__ == "__main__":
    main()
