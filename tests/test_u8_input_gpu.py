"""GPU parity of the rows SURVEY 8(f) marks "next": the IEEE-754 weight bit flips of the non-BNN
ErrorCallback branch (row 4, at the end of this file) and the uint8-input layer 0 (row 1): the image fault sweep of
imageBitflipTrainModelSel.py:19-32 with ImageErrorCallback output (callbacks.py:125-172) and
utils.normalize (utils.py:178-187) folded into an exact u8 x s8 tensor-core GEMM.

Bar: bit-exact for the integer product, the flipped bytes and the per-trial min / max; sign decisions
equal the reference's except where its batch-norm output is within fp32 rounding of 0 (the reference
forms z with an fp32 SGEMM, the device with exact integer sums)."""
import numpy as np
import pytest

from conftest import params_from_golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def T():
    import torch
    from mlpcode import _native
    _native.load()
    return torch


def dev(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def pm1(rng, shape):
    return np.where(rng.random(shape) < 0.5, -1, 1).astype(np.int8)


def padded(a, mult):
    import torch
    r, c = a.shape
    cp = (c + mult - 1) // mult * mult
    t = torch.zeros((r, cp), dtype=dev(a[:1, :1]).dtype, device="cuda")
    t[:, :c] = dev(a)
    return t


@pytest.mark.parametrize("impl", ["tcgen05", "simt"])
@pytest.mark.parametrize("M,N,K_", [(128, 256, 128), (130, 40, 200), (300, 300, 784), (64, 16, 32),
                                    (257, 513, 3072), (1000, 4096, 784)])
def test_gemm_u8_s8_exact(T, impl, M, N, K_):
    """a_unsigned: D = U . B^T with U uint8 (0..255, including bytes >= 128) and B = +-1, exact."""
    from mlpcode import _native as nat, kernels as K
    rng = np.random.default_rng(M + 3 * N + K_)
    u = rng.integers(0, 256, (M, K_), dtype=np.uint8)
    u[0, :] = 255
    b = pm1(rng, (N, K_))
    ref = u.astype(np.int32) @ b.astype(np.int32).T
    A, Bm = padded(u, 16), padded(b, 16)
    ldd = (N + 3) // 4 * 4
    D = T.full((M, ldd), 12345, dtype=T.int32, device="cuda")
    K.gemm(M=M, N=N, K=K_, A=(A,), B=(Bm,), lda=A.stride(0), ldb=Bm.stride(0), D=D, ldd=ldd,
           in_dtype=nat.DT_I8, epilogue=nat.EPI_STORE_S32, a_unsigned=True,
           impl=nat.GEMM_TCGEN05 if impl == "tcgen05" else nat.GEMM_SIMT)
    got = D.cpu().numpy()
    assert np.array_equal(got[:, :N], ref)
    assert np.all(got[:, N:] == 12345)
    # the same bytes read as int8 (a_unsigned off) give the two's-complement product: the flag matters
    D2 = T.zeros((M, ldd), dtype=T.int32, device="cuda")
    K.gemm(M=M, N=N, K=K_, A=(A,), B=(Bm,), lda=A.stride(0), ldb=Bm.stride(0), D=D2, ldd=ldd,
           in_dtype=nat.DT_I8, epilogue=nat.EPI_STORE_S32,
           impl=nat.GEMM_TCGEN05 if impl == "tcgen05" else nat.GEMM_SIMT)
    ref2 = u.view(np.int8).astype(np.int32) @ b.astype(np.int32).T
    assert np.array_equal(D2.cpu().numpy()[:, :N], ref2)


def test_gemm_rejects_bad_u8_arguments(T):
    from mlpcode import _native as nat, kernels as K
    A = T.zeros((128, 64), dtype=T.float16, device="cuda")
    D = T.zeros((128, 128), dtype=T.float32, device="cuda")
    with pytest.raises(nat.BnnError, match="a_unsigned"):
        K.gemm(M=128, N=128, K=64, A=(A,), B=(A,), lda=64, ldb=64, D=D, ldd=128, in_dtype=nat.DT_F16,
               epilogue=nat.EPI_STORE_F32, a_unsigned=True)
    A8 = T.zeros((128, 64), dtype=T.int8, device="cuda")
    with pytest.raises(nat.BnnError, match="epi_thr_stride"):
        K.gemm(M=128, N=128, K=64, A=(A8,), B=(A8,), lda=64, ldb=64, D=D, ldd=128, in_dtype=nat.DT_I8,
               epilogue=nat.EPI_STORE_S32, epi_thr_stride=128)


@pytest.mark.parametrize("mode", [0, 1, 2])
@pytest.mark.parametrize("xor_out", [0, 0x80])
def test_flip_images_ex_byproducts(T, oracle, mode, xor_out):
    """Flipped bytes (XOR-ed into operand form) and per-trial min / max of the corrupted set, for word
    (F % 4 == 0) and byte (F % 4 != 0) row layouts."""
    from mlpcode import kernels as K
    rng = np.random.default_rng(10 * mode + (xor_out > 0))
    for R, F, kk, lo, hi in ((40, 784, 878, 0, 256), (33, 64, 2, 64, 128), (7, 30, 5, 16, 32), (5, 18, 1, 0, 2)):
        images = rng.integers(lo, hi, (R, F), dtype=np.uint8)
        nt = 3
        out = T.zeros((nt, R, F), dtype=T.uint8, device="cuda")
        mm = T.full((nt, 2), -7, dtype=T.int32, device="cuda")
        K.flip_images(dev(images), kk, mode, 0xC0FFEE, 5, nt, out, xor_out=xor_out, minmax=mm)
        got, gmm = out.cpu().numpy(), mm.cpu().numpy()
        for t in range(nt):
            want = oracle.c_flip_image_rows(images, kk, mode, 0xC0FFEE, 5 + t)
            assert np.array_equal(got[t] ^ np.uint8(xor_out), want), (R, F, t)
            assert gmm[t].tolist() == [int(want.min()), int(want.max())], (R, F, t)


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_flip_walk_schedules_agree(T, oracle, monkeypatch, mode):
    """The lock-step walk (default) and round 1's one-walk-per-draw loop are two schedules of the same
    draws: identical bytes, both equal to the oracle."""
    from mlpcode import kernels as K
    rng = np.random.default_rng(70 + mode)
    images = rng.integers(0, 256, (50, 784), dtype=np.uint8)
    images[rng.random(images.shape) < 0.8] = 0
    outs = []
    for walk in ("lockstep", "perdraw"):
        monkeypatch.setenv("BNN_FLIP_WALK", walk)
        out = T.zeros((2, 50, 784), dtype=T.uint8, device="cuda")
        K.flip_images(dev(images), 878, mode, 0xFEED, 3, 2, out)
        outs.append(out.cpu().numpy())
    assert np.array_equal(outs[0], outs[1])
    for t in range(2):
        assert np.array_equal(outs[0][t], oracle.c_flip_image_rows(images, 878, mode, 0xFEED, 3 + t))


@pytest.mark.parametrize("variant", [pytest.param("u8", id="native"), pytest.param("s8", id="biased")])
def test_u8_layer0_sign_decisions_match_reference(T, oracle, variant):
    """Per-trial thresholds + batched u8 x s8 GEMM with the sign epilogue == the reference's
    normalize -> dot -> eval BN -> unitstep on each trial's images (different min / max per trial)."""
    from mlpcode import _native as nat, kernels as K
    rng = np.random.default_rng(21)
    nt, R, F, N = 3, 200, 784, 136
    ranges = [(0, 255), (5, 250), (30, 31)]
    U = np.stack([rng.integers(lo, hi + 1, (R, F), dtype=np.uint8) for lo, hi in ranges])
    for t, (lo, hi) in enumerate(ranges):
        U[t, 0, 0], U[t, 1, 1] = lo, hi
    W = rng.standard_normal((F, N)).astype(np.float32)
    bnp = dict(gamma=(rng.random(N) + 0.5).astype(np.float32), beta=(rng.standard_normal(N) * 0.3).astype(np.float32),
               mu=(rng.standard_normal(N) * 3).astype(np.float32), sigma=(rng.random(N) * 20 + 1).astype(np.float32))
    bnp["gamma"][::7] *= -1
    bias = 0 if variant == "u8" else 128
    wt = T.zeros((N, F), dtype=T.int8, device="cuda")
    bits = T.zeros((N, (F + 31) // 32), dtype=T.int32, device="cuda")
    K.binarize_weights_t(dev(W), wt_i8=wt, wt_bits=bits)
    thr, sgn = (T.zeros(N, dtype=T.float32, device="cuda") for _ in range(2))
    K.bn_sign_thresholds(dev(bnp["mu"]), dev(bnp["sigma"]), dev(bnp["gamma"]), dev(bnp["beta"]), 1e-4, N, thr, sgn)
    mm = dev(np.array(ranges, dtype=np.int32))
    thr_t = T.zeros((nt, N), dtype=T.float32, device="cuda")
    K.u8_input_thresholds(bits, N, F, thr, sgn, mm, nt, -1.0, 1.0, bias, thr_t)
    # the kernel's thresholds are the fp64 algebra's, up to the rounding of the fp32 BN threshold
    for t, (lo, hi) in enumerate(ranges):
        tau, s = oracle.u8_input_thresholds(W, bnp, lo, hi, bias=bias)
        assert np.array_equal(sgn.cpu().numpy(), s.astype(np.float32))
        assert np.max(np.abs(thr_t[t].cpu().numpy() - tau)) <= 1.0, t
    A = dev(U ^ np.uint8(0x80 if bias else 0))
    Np = (N + 15) // 16 * 16
    out = T.full((nt, R, Np), 7, dtype=T.int8, device="cuda")
    K.gemm(M=R, N=N, K=F, A=(A,), B=(wt,), lda=F, ldb=F, D=out, ldd=Np, in_dtype=nat.DT_I8,
           epilogue=nat.EPI_SIGN_I8, batch=nt, stride_a=R * F, stride_b=0, stride_d=R * Np, epi_thr=thr_t,
           epi_sgn=sgn, epi_thr_stride=N, a_unsigned=bias == 0)
    got = out.cpu().numpy()
    assert np.all(got[:, :, N:] == 7)
    for t in range(nt):
        want, zt = oracle.u8_layer0_reference(U[t], W, bnp)
        diff = got[t, :, :N] != want.astype(np.int8)
        assert np.all(np.abs(zt[diff]) < 1e-4), (t, np.abs(zt[diff]).max())
        assert diff.mean() < 0.01, t


def _small_net(golden):
    from test_layers_gpu import build_net
    g = golden("train_small")                          # 784-64-32-10
    units = [int(u) for u in g["units"]]
    params = params_from_golden(g, "after1_", len(units) - 1)
    return units, params, build_net(units, params, out="identity")


@pytest.mark.parametrize("variant", [pytest.param("u8", id="native"), pytest.param("s8", id="biased")])
def test_batched_image_sweep_matches_oracle_and_per_trial_path(golden, oracle, monkeypatch, variant):
    from mlpcode.sweep import ImageFaultSweep
    monkeypatch.setenv("BNN_U8_GEMM", variant)
    units, params, nn = _small_net(golden)
    rng = np.random.default_rng(8)
    R = 96
    imgs = rng.integers(0, 256, (R, units[0]), dtype=np.uint8)
    labels = rng.integers(0, 10, R).astype(np.uint8)
    p, seed, n_trials = 0.14, 99, 5
    k = int(round(p * units[0] * 8))
    fast = ImageFaultSweep(nn, imgs, labels, trials_per_launch=2)
    assert fast.batched
    slow = ImageFaultSweep(nn, imgs, labels, trials_per_launch=2, batched=False)
    for mode in (2, 0):
        got = fast.run(p, n_trials, seed, mode=mode)
        ref = slow.run(p, n_trials, seed, mode=mode)
        want = np.array([oracle.accuracy(params, oracle.normalize(
            oracle.c_flip_image_rows(imgs, k, mode, seed, t), -1, 1), labels, out_af="identity")
            for t in range(n_trials)])
        assert np.array_equal(ref, want)
        # exact sums vs the reference's fp32 SGEMM: at most one row may sit on a sign boundary
        assert np.max(np.abs(got - want)) <= 100.0 / R + 1e-9, (got, want)


def test_batched_image_sweep_at_full_size_equals_per_trial_path():
    """BASELINE-sized property: 10 000 MNIST-shaped images through the 784-4096-4096-4096-10 network.
    The exact u8 x s8 layer 0 and the fp32-normalize + fp16-split layer 0 are independent
    implementations of the same map; their layer-0 activations may differ only where z sits within
    rounding of its threshold (batch-norm parameters are generic reals here, as after training —
    untrained mu = beta = 0 would put the threshold ON the lattice of exact sums, where thousands of
    units per trial tie), and a random 4-layer network amplifies each such unit into a changed
    prediction, so accuracies agree to a few rows, not exactly."""
    import torch
    import bnn_oracle as O
    from mlpcode import kernels as K
    from mlpcode.sweep import ImageFaultSweep
    from test_layers_gpu import build_net
    units = [784, 4096, 4096, 4096, 10]
    rng = np.random.default_rng(3)
    params = O.new_params(units, rng)
    for i, b in enumerate(params["bn"][:-1]):
        n = b["mu"].size
        b["sigma"][:] = 1000.0                       # pre-activations of +-1 layers have variance ~ K
        b["mu"][:] = (rng.standard_normal(n) * 3).astype(np.float32)
        b["beta"][:] = (rng.standard_normal(n) * 0.1).astype(np.float32)
        b["gamma"][:] = (rng.random(n) + 0.5).astype(np.float32) * np.where(rng.random(n) < 0.2, -1, 1)
    nn = build_net(units, params, out="identity")
    R, T = 10000, 4
    imgs = rng.integers(0, 256, (R, 784), dtype=np.uint8)
    imgs[rng.random((R, 784)) < 0.8] = 0             # MNIST-like: mostly background
    labels = rng.integers(0, 10, R).astype(np.uint8)
    fast = ImageFaultSweep(nn, imgs, labels, trials_per_launch=T)
    slow = ImageFaultSweep(nn, imgs, labels, trials_per_launch=T, batched=False)
    assert fast.batched and not slow.batched
    got, ref = fast.run(0.14, T, 77), slow.run(0.14, T, 77)
    assert len(set(np.round(got, 6))) > 1                    # the trials really differ
    # both sweeps hold the same corrupted bytes, and the fused min / max equals a separate pass
    assert torch.equal(fast.flipped, slow.flipped)
    mm = torch.zeros(2, dtype=torch.int32, device="cuda")
    for t in range(T):
        K.minmax_u8(slow.flipped[t], mm)
        assert fast.mm[t].tolist() == mm.tolist()
    # layer 0, trial by trial: at most 1e-5 of the R x 4096 sign decisions may sit on a rounding boundary
    N0, xf, mmf = units[1], torch.empty((R, 784), device="cuda"), torch.zeros(2, device="cuda")
    for t in range(T):
        K.minmax_u8(slow.flipped[t], mm, mmf)
        K.normalize_u8(slow.flipped[t], mmf[0:1], mmf[1:2], -1.0, 1.0, xf)
        a_ref = nn._layers[0].forward(xf, isTrain=False).t                    # +-1 as fp32, whatever the operand form
        a_got = fast.act[0][t * R:(t + 1) * R]
        if fast.f4[1]:   # E2M1 nibbles (the layer above runs on the FP4 pipe): 0x2 -> +1, 0xA -> -1
            nib = torch.stack((a_got & 0xF, a_got >> 4), dim=-1).reshape(R, -1)[:, :N0]
            assert bool(((nib == 0x2) | (nib == 0xA)).all())
            a_got = torch.where(nib == 0x2, 1.0, -1.0)
        else:
            a_got = a_got[:, :N0].float()
        assert int((a_ref != a_got).sum()) <= 1e-5 * R * N0, t
    assert np.max(np.abs(got - ref)) <= 0.3, (got, ref)      # <= 30 of 10 000 rows


def test_unchanged_script_flow_on_bytes_matches_oracle(golden, oracle):
    """imageBitflipTrainModelSel.py:24-28 verbatim: icb(testX) -> normalize(newX, -1, 1) -> get_accuracy.
    `normalize` hands layer 0 the bytes (NormalizedU8); results equal the reference semantics, also with
    weight faults attached and with a batched evaluation."""
    from mlpcode.callbacks import ErrorCallback, ImageErrorCallback
    from mlpcode.utils import NormalizedU8, normalize
    units, params, nn = _small_net(golden)
    rng = np.random.default_rng(12)
    R = 96
    imgs = rng.integers(0, 256, (R, units[0]), dtype=np.uint8)
    labels = rng.integers(0, 10, R).astype(np.uint8)
    k = int(round(0.14 * units[0] * 8))
    icb = ImageErrorCallback(0.14, seed=31)
    for trial in range(3):
        newX = icb(imgs, gpu=True)
        nx = normalize(newX, newMin=-1, newMax=1)
        assert isinstance(nx, NormalizedU8)
        acc = nn.get_accuracy(nx, labels)
        want_x = oracle.normalize(oracle.c_flip_image_rows(imgs, k, 2, 31, trial), -1, 1)
        want = oracle.accuracy(params, want_x, labels, out_af="identity")
        assert abs(acc - want) <= 100.0 / R + 1e-9, (trial, acc, want)
        assert abs(nn.get_accuracy(nx, labels, batch_size=40) - acc) < 1e-9      # row slices keep min / max
        assert np.array_equal(nx.get(), want_x)              # materialised values = the reference's bits
    # weight faults on top: the callbacks flip the packed copies the u8 GEMM reads, and S_n follows them
    nn.addCallbacks(ErrorCallback(0.1, mode=2, bnn=True, seed=1234))
    got = nn.predict(nx).get()
    n = len(units) - 1
    masks = [oracle.c_bernoulli_mask(units[i], units[i + 1], 0.1, 1234, i, i) for i in range(n)]
    hooks = [(lambda mk: (lambda W: oracle.flip_bnn_mask(W, mk)))(mk) for mk in masks]
    ref = oracle.predict(params, want_x.copy(), out_af="identity", weight_hooks=hooks)
    assert np.mean(got.argmax(1) == ref.argmax(1)) >= 1 - 1.5 / R
    nn.clearCallbacks()


def test_weight_fault_sweep_on_uint8_images_matches_oracle(golden, oracle):
    """WeightFaultSweep fed `normalize(uint8 images)` (weightsErrorInjectionv2.py:15-25): layer 0 runs on
    the bytes against each trial's corrupted weights, thresholds carry each trial's own column sums.
    Accuracies equal the reference semantics for Bernoulli and exact-k faults."""
    from mlpcode.sweep import WeightFaultSweep
    from mlpcode.utils import normalize
    units, params, nn = _small_net(golden)                    # 784-64-32-10, identity output
    n = len(units) - 1
    rng = np.random.default_rng(15)
    R = 96
    imgs = rng.integers(0, 256, (R, units[0]), dtype=np.uint8)
    imgs[rng.random(imgs.shape) < 0.5] = 0
    labels = rng.integers(0, 10, R).astype(np.uint8)
    X = oracle.normalize(imgs, -1, 1)
    sweep = WeightFaultSweep(nn, normalize(imgs, newMin=-1., newMax=1.), labels, trials_per_launch=2)
    assert sweep.u8 is not None
    tol = 100.0 / R + 1e-9                                    # one row may sit on a rounding boundary
    assert abs(sweep.clean_accuracy() - oracle.accuracy(params, X.copy(), labels, out_af="identity")) <= tol
    p, seed, n_trials = 0.15, 20251, 5
    got = sweep.run_bernoulli(p, n_trials, seed)
    want = []
    for t in range(n_trials):
        masks = [oracle.c_bernoulli_mask(units[i], units[i + 1], p, seed, i, t) for i in range(n)]
        hooks = [(lambda m: (lambda W: oracle.flip_bnn_mask(W, m)))(m) for m in masks]
        want.append(oracle.accuracy(params, X.copy(), labels, out_af="identity", weight_hooks=hooks))
    assert np.max(np.abs(got - np.array(want))) <= tol, (got, want)
    assert len(set(got)) > 1
    # the same sweep on the materialised fp32 images (fp16-split layer 0) agrees
    ref = WeightFaultSweep(nn, X, labels, trials_per_launch=2)
    assert ref.u8 is None
    assert np.max(np.abs(ref.run_bernoulli(p, n_trials, seed) - got)) <= tol
    # exact-k over the whole network
    ks = [0, 1, 500, 20000]
    total = sum(units[i] * units[i + 1] for i in range(n))
    got = sweep.run_exactk(lambda t: ks[t], len(ks), 77)
    offs = np.cumsum([0] + [units[i] * units[i + 1] for i in range(n)])
    want = []
    for t, k in enumerate(ks):
        pos = oracle.c_exactk_positions(total, 77, t, k).astype(np.int64)
        hooks = []
        for i in range(n):
            Kd, N = units[i], units[i + 1]
            mine = pos[(pos >= offs[i]) & (pos < offs[i + 1])] - offs[i]
            mask = np.zeros((Kd, N), bool)
            mask[mine % Kd, mine // Kd] = True                # e = n*K + k  ->  W[k, n]
            hooks.append((lambda m: (lambda W: oracle.flip_bnn_mask(W, m)))(mask))
        want.append(oracle.accuracy(params, X.copy(), labels, out_af="identity", weight_hooks=hooks))
    assert np.max(np.abs(got - np.array(want))) <= tol, (got, want)


# ------------------------------------------------------------------------- IEEE-754 weight bit flips
@pytest.mark.parametrize("mode", [0, 1, 2])
def test_f32_bit_flips_match_oracle(T, oracle, golden, mode):
    """bnn_flip_f32_bits (ErrorCallback non-BNN branch, callbacks.py:35-59, 100-123) == the C restatement
    of the same draw, bit for bit; the restatement's semantics are pinned against the reference's own
    outputs in tests/test_oracle.py."""
    from mlpcode import kernels as K
    rng = np.random.default_rng(40 + mode)
    big = (rng.standard_normal((300, 257)) * np.exp(rng.standard_normal((300, 257)) * 4)).astype(np.float32)
    pm = np.where(rng.random((784, 64)) < 0.5, -1.0, 1.0).astype(np.float32)
    for arr in (golden("float_faults")["W"], big, pm):
        for p in (0.1, 0.5, 1.0):
            d = dev(arr.copy())
            K.flip_f32_bits(d, p, mode, 0xABCD1234, 2)
            want = oracle.c_flip_f32_bits(arr, p, mode, seed=0xABCD1234, trial=2)
            assert np.array_equal(d.cpu().numpy().view(np.uint32), want.view(np.uint32)), (arr.shape, p)


def test_error_callback_float_branch(T, oracle, golden):
    """ErrorCallback(p, mode, bnn=False)(array): reference call signature on a dense real-valued array."""
    from mlpcode.callbacks import ErrorCallback
    W = golden("float_faults")["W"]
    cb = ErrorCallback(0.5, mode=2, bnn=False, seed=77)
    out0 = cb(W, gpu=True).get()
    out1 = cb(W, gpu=True).get()                               # next call = next trial
    assert np.array_equal(out0.view(np.uint32), oracle.c_flip_f32_bits(W, 0.5, 2, 77, 0).view(np.uint32))
    assert np.array_equal(out1.view(np.uint32), oracle.c_flip_f32_bits(W, 0.5, 2, 77, 1).view(np.uint32))
    changed = (out0.view(np.uint32) != W.view(np.uint32)).sum()
    assert 0 < changed <= round(W.size * 0.5)
    small = ErrorCallback(0.001, mode=2, bnn=False, seed=1)    # round(108 * 0.001) == 0 elements
    assert small(W, gpu=True) is W                             # callbacks.py:111-112 returns the input
