"""Parity on BASELINE.json's own configurations, where the benchmark runs.

  config 1  784-256-256-10,          batch 256   (/root/reference/test.py:9-47 shape; SURVEY 8d recipe)
  config 2  784-4096-4096-4096-10,   batch 8192  (the bench.py workload)

against the NumPy oracle (a bit-exact restatement of the reference, pinned by oracle/gen_golden.py) on the
same seeded inputs.  Bars (north star): +-1 GEMM outputs bit-exact; loss, BN outputs, dX, dW and the
optimiser state within 1e-5 relative (norm-wise).  The weight GRADIENT is compared directly — the state
test is blind to it, because lr * dW sits at or below the fp32 resolution of W (VERDICT r1 "weak" #2):
`BinaryLayer.capture_dw` receives X^T . delta' formed from the very operands and passes of the update
GEMM (int8 digit product for the 4096-wide layers, fp16 split for layer 0 and the 10-wide output layer).
Reference lines: mlpcode/layers.py:257-268 (dW, dX, update), :74-127 (batch-norm), network.py:372-392.

Free-running steps and ill-conditioned sign decisions.  sign(BN(z)) is discontinuous, and the reference
takes two kinds of decision that its OWN arithmetic does not determine (tools/debug_parity.py prints them):
  (1) layer 0, real-valued input: z = X . sign(W) is an fp32 SGEMM whose rounding depends on the BLAS
      summation order; of the 33.5 M decisions of config 2 about three have |BN out| < 1e-7 and fall either
      way, and each flip cascades (569 flips in layer 1, 31 916 in layer 2, loss moves by 3e-4);
  (2) hidden layers after the first update: a column whose batch mean is an exact integer has rows with
      z == mu, whose BN output is 0 * gamma + beta = beta — and a hidden layer's beta holds nothing but the
      rounding noise of sum_b delta (1e-13), whose sign is arbitrary.
Neither is an error of either side, so the FREE-RUNNING tests run where the reference is well-conditioned:
inputs on the dyadic grid k/16 (every partial sum of layer 0 is exact in fp32 in any order, so z0 is
bit-identical on both sides) and, for multi-step runs, a seed for which the oracle itself reports no
decision within 1e-6 of zero (the precondition is asserted, not assumed).  The TEACHER-FORCED test keeps
continuous inputs and checks every stage on the oracle's own stage inputs, counting the decisions of
kind (1).
"""
import numpy as np
import pytest

from conftest import rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-5


def synthetic(units, B, seed, n_batches=1, grid=False):
    """X uniform [-1, 1] fp32 (grid: uniform on k/16), labels uniform, W = randn * sqrt(1/K),
    BN at its initial state."""
    import bnn_oracle as O
    rng = np.random.default_rng(seed)
    params = O.new_params(units, rng)
    batches = []
    for _ in range(n_batches):
        if grid:
            X = (rng.integers(-16, 17, (B, units[0])) / 16.0).astype(np.float32)
        else:
            X = rng.random((B, units[0]), dtype=np.float32) * 2 - 1
        Y = np.eye(units[-1], dtype=np.int8)[rng.integers(0, units[-1], B)]
        batches.append((X, Y))
    return params, batches


def build_net(units, params, lr):
    from mlpcode.activation import ActivationFuncs as af
    from mlpcode.loss import LossFuncs as lf
    from mlpcode.network import Network
    nn = Network(units, useGpu=True, binarized=True, useBatchNorm=True)
    nn.compile(lr=lr, hiddenAf=af.sign, outAf=af.softmax, lossF=lf.cross_entropy)
    nn.load_weights([w.copy() for w in params["W"]])
    nn.loadBatchNormParameters(*[[b[k].copy() for b in params["bn"]]
                                 for k in ("gamma", "beta", "mu", "sigma")])
    return nn


def state_errors(nn, ref, n, beta_out_exact=None):
    """Norm-wise relative error of every state tensor; hidden-layer beta on an absolute floor (dbeta =
    sum_b delta vanishes identically below a batch-norm backward: both sides hold rounding noise).
    beta_out_exact: the output layer's beta evaluated in fp64 from the (bit-exact) integer pre-activations of
    the last layer, see `output_beta_fp64`; when given, this build's beta is held to 1e-5 of THAT value and the
    oracle's fp32 value only to the amplified bound derived there."""
    g, b, m, s = nn.batchNormParams
    errs = {}
    for i in range(n):
        errs[f"W{i}"] = rel_err(nn.weights[i].get(), ref["W"][i])
        for name, dev in (("gamma", g[i]), ("mu", m[i]), ("sigma", s[i])):
            errs[f"{name}{i}"] = rel_err(dev.get(), ref["bn"][i][name])
        got, want = b[i].get(), ref["bn"][i]["beta"]
        if i < n - 1:
            assert np.abs(want).max() < 1e-5, "hidden-layer beta should only hold rounding noise"
            errs[f"beta{i}"] = float(np.abs(got - want).max()) * 10   # < 1e-6 absolute
        elif beta_out_exact is None:
            errs[f"beta{i}"] = rel_err(got, want)
        else:
            errs[f"beta{i}"] = rel_err(got, beta_out_exact)
            errs[f"beta{i}_oracle_fp32_vs_fp64"] = rel_err(want, beta_out_exact) / 20   # see output_beta_fp64
    return errs


def output_beta_fp64(z_last, Y, lr, eps):
    """beta of the output layer after the first step, in fp64 from the last layer's integer pre-activations.

    beta = -lr * sum_b (p - y) / B is a 60-fold cancellation at B = 8192 (sum |delta| = 0.18 per class against
    |sum delta| = 0.003), and it inherits every relative error of the logits' scale 17-fold: NumPy's fp32
    Z.var(axis=0) adds 8192 rows one after the other and lands 3e-6 from the exact variance (this build sums
    the integers exactly), which moves the reference's own beta by 2.5e-5.  So the 1e-5 bar is applied against
    the fp64 evaluation, and the fp32 oracle has to sit within 20x of it (its amplified summation error)."""
    z = z_last.astype(np.float64)
    logits = (z - z.mean(axis=0)) / np.sqrt(z.var(axis=0) + eps)          # gamma = 1, beta = 0 at the first step
    e = np.exp(logits - logits.max(axis=1, keepdims=True))
    p = e / e.sum(axis=1, keepdims=True)
    return -lr * ((p - Y) / z.shape[0]).sum(axis=0)


def ambiguous_decisions(rec, first_step):
    """Hidden-layer sign decisions of one recorded oracle step that the reference's own arithmetic does not
    determine: |BN output| < 1e-6.  An exact zero in the very first step is NOT ambiguous (beta is exactly 0,
    0 * gamma + 0 = +0 on both sides, sign(+0) = +1)."""
    n = 0
    for cache in rec["caches"][:-1]:
        bn = cache["z"]
        amb = np.abs(bn) < 1e-6
        if first_step:
            amb &= bn != 0
        n += int(amb.sum())
    return n


# ---------------------------------------------------------------------------------------- config 1
def test_config1_first_step_continuous_input(oracle):
    """BASELINE configs[0] with the SURVEY 8(d) inputs as they are (continuous X): one step."""
    units, B, lr = [784, 256, 256, 10], 256, 0.01
    params, batches = synthetic(units, B, seed=0)
    X, Y = batches[0]
    nn = build_net(units, params, lr)
    ref, rec = oracle.copy_params(params), {}
    cost_ref = oracle.train_step(ref, X.copy(), Y, lr, record=rec)
    assert ambiguous_decisions(rec, True) == 0
    cost = nn.train_step(X, Y, lr)
    assert abs(cost - cost_ref) <= TOL * abs(cost_ref), (cost, cost_ref)
    bad = {k: v for k, v in state_errors(nn, ref, len(units) - 1).items() if not v < TOL}
    assert not bad, bad


def test_config1_three_steps_match_oracle(oracle):
    """BASELINE configs[0]: three SGD steps on three batches, loss and full state after every step, on the
    first seed whose oracle run is well-conditioned (module docstring)."""
    units, B, lr = [784, 256, 256, 10], 256, 0.01
    n = len(units) - 1
    for seed in range(64):
        params, batches = synthetic(units, B, seed=seed, n_batches=3, grid=True)
        probe = oracle.copy_params(params)
        amb = 0
        for s, (X, Y) in enumerate(batches):
            rec = {}
            oracle.train_step(probe, X.copy(), Y, lr, record=rec)
            amb += ambiguous_decisions(rec, s == 0)
        if amb == 0:
            break
    else:
        pytest.fail("no well-conditioned seed among 64")
    nn = build_net(units, params, lr)
    ref = oracle.copy_params(params)
    for s, (X, Y) in enumerate(batches):
        cost = nn.train_step(X, Y, lr)
        cost_ref = oracle.train_step(ref, X.copy(), Y, lr)
        assert abs(cost - cost_ref) <= TOL * abs(cost_ref), (s, cost, cost_ref)
        errs = state_errors(nn, ref, n)
        bad = {k: v for k, v in errs.items() if not v < TOL}
        assert not bad, (s, bad)
    print(f"config 1 (seed {seed}): 3 steps, worst state error {max(errs.values()):.2e}")
    probs = nn.predict(batches[0][0]).get()
    assert rel_err(probs, oracle.predict(ref, batches[0][0].copy())) < TOL


# ---------------------------------------------------------------------------------------- config 2
UNITS2, B2, LR2 = [784, 4096, 4096, 4096, 10], 8192, 0.01


@pytest.fixture(scope="module")
def config2(oracle):
    """One oracle step at full size (5-10 s of NumPy on the host), with every intermediate recorded."""
    params, batches = synthetic(UNITS2, B2, seed=0)
    X, Y = batches[0]
    ref = oracle.copy_params(params)
    rec = {}
    cost = oracle.train_step(ref, X.copy(), Y, LR2, record=rec)
    return dict(params=params, X=X, Y=Y, ref=ref, rec=rec, cost=cost)


def test_config2_full_step_matches_oracle(oracle):
    """The benchmarked path (fused statistics, int8 digit dW, one step at B = 8192), free running on
    grid inputs: every sign decision (100 M of them), the loss and the whole optimiser state."""
    params, batches = synthetic(UNITS2, B2, seed=0, grid=True)
    X, Y = batches[0]
    ref, rec = oracle.copy_params(params), {}
    cost_ref = oracle.train_step(ref, X.copy(), Y, LR2, record=rec)
    assert ambiguous_decisions(rec, True) == 0
    nn = build_net(UNITS2, params, LR2)
    cost = nn.train_step(X, Y, LR2)
    assert abs(cost - cost_ref) <= TOL * abs(cost_ref), (cost, cost_ref)
    beta_exact = output_beta_fp64(rec["caches"][-1]["z_pre"], Y, LR2, oracle.EPSILON)
    errs = state_errors(nn, ref, len(UNITS2) - 1, beta_out_exact=beta_exact)
    print("config 2 state errors:", {k: f"{v:.1e}" for k, v in errs.items()})
    bad = {k: v for k, v in errs.items() if not v < TOL}
    assert not bad, bad


def _sign_activation(a_ref, consumer_units):
    """The oracle's +-1 activations in the layer-to-layer operand form a producing layer would emit for
    a consumer of that width: transposed int8 +-1 / +-127 pieces when the consumer forms dW on the int8
    pipe, the transposed fp16 copy otherwise (layers.py of this package, BinaryLayer.forward)."""
    import torch
    from mlpcode import kernels as K
    from mlpcode.layers import SignActivation
    B, n = a_ref.shape
    a = torch.from_numpy(a_ref.astype(np.int8)).cuda()
    i8 = torch.zeros((B, (n + 15) // 16 * 16), dtype=torch.int8, device="cuda")
    i8[:, :n] = a
    if K.int8_dw_ok(consumer_units, B):
        t8 = torch.zeros((n, (B + 15) // 16 * 16), dtype=torch.int8, device="cuda")
        t8[:, :B] = a.t()
        return SignActivation(i8, n, t8=t8, t8x127=(t8.to(torch.int16) * 127).to(torch.int8))
    t16 = torch.zeros((n, (B + 7) // 8 * 8), dtype=torch.float16, device="cuda")
    t16[:, :B] = a.t().to(torch.float16)
    return SignActivation(i8, n, t16=t16)


def test_config2_stages_teacher_forced(config2):
    """Every stage of the step on the ORACLE's stage inputs: z of the +-1 layers bit-exact, layer-0 z and
    BN outputs within tolerance, sign mismatches only where |BN output| ~ 0; backward: dX within 1e-5 and
    dW — captured from the update GEMM's own operands — within 1e-5 of an fp64 product and of the oracle."""
    import torch
    c = config2
    rec, n = c["rec"], len(UNITS2) - 1
    nn = build_net(UNITS2, c["params"], LR2)
    caches = rec["caches"]
    # ---- forward
    for i, layer in enumerate(nn._layers):
        x_in = c["X"] if i == 0 else _sign_activation(caches[i]["input"], UNITS2[i + 1])
        out = layer.forward(x_in, isTrain=True)
        z = layer.cache["z"][:, :UNITS2[i + 1]]
        z_ref = caches[i]["z_pre"]
        if i == 0:
            assert rel_err(z.cpu().numpy(), z_ref) < 5e-6
        else:
            assert torch.equal(z.to(torch.float32), torch.from_numpy(z_ref).cuda()), f"z{i} not bit-exact"
        bn_ref = caches[i]["z"]
        if i < n - 1:
            got = out.t.cpu().numpy()
            bad = got != caches[i + 1]["input"]
            frac = float(bad.mean())
            assert np.all(np.abs(bn_ref[bad]) < 1e-4 * np.abs(bn_ref).max()), f"layer {i}"
            assert frac < 1e-4, (i, frac)
            print(f"layer {i}: {int(bad.sum())} of {bad.size} sign decisions differ (all at |BN out| ~ 0)")
        else:
            assert rel_err(out.logits.cpu().numpy(), bn_ref) < TOL
            assert rel_err(out.get(), rec["probs"]) < TOL
    # ---- backward, each layer fed the oracle's incoming delta; the forward caches above belong to the
    # oracle's inputs, so every operand of dW / dX is the reference's
    for i in reversed(range(n)):
        layer = nn._layers[i]
        Kd, N = layer.weights_shape
        d_in = rec["dLdA"] if i == n - 1 else rec["dX"][i + 1]
        cap = torch.zeros((Kd, N), dtype=torch.float32, device="cuda")
        layer.capture_dw = cap
        dX = layer.backwards(torch.from_numpy(d_in).cuda(), LR2, activDeriv=False)
        layer.capture_dw = None
        dprime = rec["delta_bn"][i]                       # delta' = BN backward output (layers.py:257)
        col_rms = np.sqrt((dprime.astype(np.float64) ** 2).mean(axis=0))
        ratio = np.abs(dprime).max(axis=0) / np.maximum(col_rms, 1e-300)
        x_ref = torch.from_numpy(caches[i]["input"]).cuda().double()
        dW64 = (x_ref.t() @ torch.from_numpy(dprime).cuda().double())
        e64 = float((cap.double() - dW64).norm() / dW64.norm())
        e_or = rel_err(cap.cpu().numpy(), rec["dW"][i])
        e_np = float((torch.from_numpy(rec["dW"][i]).cuda().double() - dW64).norm() / dW64.norm())
        print(f"layer {i}: dW vs fp64 {e64:.2e}, vs oracle fp32 {e_or:.2e} (NumPy SGEMM vs fp64 {e_np:.2e}); "
              f"max/rms of delta' per column: median {np.median(ratio):.1f}, worst {ratio.max():.1f}")
        assert e64 < TOL / 2 and e_or < TOL, (i, e64, e_or)     # 2x margin against the exact product
        if i > 0:
            assert rel_err(dX.get(), rec["dX"][i]) < TOL, i
        else:
            assert dX is None                               # layer-0 dX is dead work in the reference


# ---------------------------------------------------------------- the int8 dW quantiser on its own
@pytest.mark.parametrize("dist", ["gaussian", "row_sparse", "lognormal2"])
def test_int8_dw_quantiser_full_size(dist):
    """bnn_bn_bwd_finalize (column scales) -> bnn_bn_bwd_apply_q (digits) -> int8 digit GEMM at config-2
    size (B = 8192, 4096 x 4096) against an fp64 product, for a benign delta', a late-training one (a few
    samples carry the gradient, the rest are tiny but non-zero) and the heavy-tailed one of
    test_split_gemm_precision_at_full_size.  Error model: 7e-8 * bound / rms per column."""
    import torch as T
    from mlpcode import _native as nat, kernels as K
    g = T.Generator(device="cuda").manual_seed(5)
    B, N, Kd = 8192, 4096, 4096
    base = T.randn((B, N), generator=g, device="cuda") * 1e-6
    if dist == "row_sparse":
        rows = (T.rand((B, 1), generator=g, device="cuda") < 0.002).float()
        delta = base * (rows + 1e-4)
    elif dist == "lognormal2":
        delta = base * T.exp(2 * T.randn((B, N), generator=g, device="cuda"))
    else:
        delta = base
    delta = delta.contiguous()
    Z = T.zeros((B, N), dtype=T.int16, device="cuda")
    # gamma = 1, var + eps = 1, zero column sums  =>  delta' = c0 * delta with c0 = 1 (to fp32 rounding)
    f32 = dict(dtype=T.float32, device="cuda")
    mu, var = T.zeros(N, **f32), T.full((N,), 1.0 - 1e-4, **f32)
    gamma, beta, mu_run, sig_run = T.ones(N, **f32), T.zeros(N, **f32), T.zeros(N, **f32), T.ones(N, **f32)
    red = T.zeros((2, N), dtype=T.float64, device="cuda")
    red_max = T.zeros((2, N), dtype=T.int32, device="cuda")
    red_max[0] = delta.abs().amax(dim=0).view(T.int32)
    coef, bound, colq = T.zeros((3, N), **f32), T.zeros(1, **f32), T.zeros((2, N), **f32)
    K.bn_bwd_finalize(red, red_max, N, B, mu, var, 1e-4, gamma, beta, mu_run, sig_run, 0.0, 0.9, None,
                      coef, bound, col_quant=colq)
    q1, q3, q2 = (T.zeros((N, B), dtype=T.int8, device="cuda") for _ in range(3))
    K.bn_bwd_apply_q(delta, Z, B, N, coef, bound, colq, q1, q3, q2)
    dprime = (coef[0].double() * delta.double())            # what the kernel quantises (c1 = c2 = 0)
    assert float(coef[1].abs().max()) == 0 and float(coef[2].abs().max()) == 0
    Xt = (T.randint(0, 2, (Kd, B), generator=g, device="cuda") * 2 - 1).to(T.int8)
    Xt127 = (Xt.to(T.int16) * 127).to(T.int8)
    dW = T.zeros((Kd, N), **f32)
    K.gemm(M=Kd, N=N, K=B, A=(Xt, Xt127), B=(q1, q3, q2), lda=B, ldb=B, D=dW, ldd=N, in_dtype=nat.DT_I8,
           epilogue=nat.EPI_STORE_F32, passes=((0, 0), (1, 1), (0, 2)), hi_from_pass=1, col_scale=colq[1])
    rows = slice(0, 512)
    ref = Xt[rows].double() @ dprime
    err = float((dW[rows].double() - ref).norm() / ref.norm())
    rms = dprime.pow(2).mean(dim=0).sqrt()
    ratio = (delta.abs().amax(dim=0).double() / rms)
    model = 7e-8 * float((ratio.pow(2).mean()).sqrt())
    print(f"{dist}: int8 digit dW vs fp64 {err:.2e}; max/rms per column median {float(ratio.median()):.1f}; "
          f"error model {model:.2e}")
    assert err < (TOL / 2 if dist != "lognormal2" else TOL), (dist, err)
    assert err < 2.5 * model + 2e-7
