"""The +-1 x +-1 product on the FP4 tensor pipe (BNN_DT_F4: tcgen05 kind::mxf4.block_scale, unit scale factors).

+1 / -1 / 0 are exact in E2M1 (0x2 / 0xA / 0x0) and the fp32 accumulators hold exact integers for K < 2^24, so the
bar is the one of the int8 path: BIT-EXACT against an integer product (north star: "bit-exact for binarized GEMM
outputs").  Covered: the bits -> nibbles expansion (bnn_bits_to_f4, incl. ragged K padding), every epilogue the
kind has (STORE_S32 / STORE_S16 / SIGN_I8 / SIGN_F4), ragged M / N, all three tile widths (32 / 128 / 224), batched
products with shared and per-batch operands, the SIGN_F4 epilogue of the int8 and fp16 kinds (layer 0 feeding an
FP4 layer), guard bands around every output, and the evaluation paths end to end with BNN_FP4 on / off
(reference: X.dot(weight) mlpcode/layers.py:220 on binarized operands, layers.py:283-292)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GUARD = 0x5C


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


def to_f4(x, ld=None):
    """int8 matrix of +-1 / 0 -> E2M1 nibbles, two per byte (low nibble = even column), pitch ld elements."""
    r, c = x.shape
    ld = ld or (c + 31) // 32 * 32
    nib = np.zeros((r, ld), np.uint8)
    nib[:, :c] = np.where(x > 0, 0x2, np.where(x < 0, 0xA, 0x0))
    return (nib[:, 0::2] | (nib[:, 1::2] << 4)).astype(np.uint8)


def from_f4(b, n):
    """E2M1 nibbles (only 0x2 / 0xA / 0x0 occur) -> int8 +-1 / 0, first n columns."""
    b = np.asarray(b, np.uint8)
    nib = np.empty((b.shape[0], b.shape[1] * 2), np.uint8)
    nib[:, 0::2], nib[:, 1::2] = b & 0xF, b >> 4
    assert np.isin(nib, (0x0, 0x2, 0xA)).all()
    return np.where(nib == 0x2, 1, np.where(nib == 0xA, -1, 0)).astype(np.int8)[:, :n]


def guarded(T, rows, row_bytes, fill=GUARD):
    """A [rows, row_bytes] uint8 view inside a larger buffer whose other bytes must stay untouched."""
    buf = T.full(((rows + 2) * row_bytes + 512,), fill, dtype=T.uint8, device="cuda")
    view = buf[256 + row_bytes: 256 + row_bytes + rows * row_bytes].view(rows, row_bytes)
    return buf, view


def guards_intact(buf, rows, row_bytes):
    b = buf.cpu().numpy()
    return bool(np.all(b[:256 + row_bytes] == GUARD) and np.all(b[256 + row_bytes + rows * row_bytes:] == GUARD))


@pytest.mark.parametrize("R,C", [(1, 32), (5, 100), (64, 784), (300, 4096), (7, 33)])
def test_bits_to_f4(T, oracle, R, C):
    from mlpcode import kernels as K
    rng = np.random.default_rng(R + C)
    x = pm1(rng, (R, C))
    bits = dev(oracle.pack_sign_rows(x.astype(np.float32)).view(np.int32))
    ld = (C + 31) // 32 * 32
    buf, out = guarded(T, R, ld // 2)
    K.bits_to_f4(bits, C, out)
    got = out.cpu().numpy()
    assert np.array_equal(got, to_f4(x, ld))                   # incl. zero nibbles in the K padding
    assert np.array_equal(from_f4(got, C), x)
    assert guards_intact(buf, R, ld // 2)


@pytest.mark.parametrize("K_,N", [(32, 8), (784, 100), (1000, 33), (4096, 512), (40, 5)])
def test_binarize_weights_emits_nibbles(T, K_, N):
    """bnn_binarize_weights_t(wt_f4): the nibbles bnn_bits_to_f4 makes of the packed words, straight from W."""
    from mlpcode import kernels as K
    rng = np.random.default_rng(K_ + N)
    W = rng.standard_normal((K_, N)).astype(np.float32)
    W[rng.random((K_, N)) < 0.05] = 0.0                      # sign(0) = +1 (layers.py:278-279)
    ld = (K_ + 31) // 32 * 32
    buf, out = guarded(T, N, ld // 2, fill=0)                # the layer's buffer: zero-initialised, padding stays zero
    bits = T.zeros((N, ld // 32), dtype=T.int32, device="cuda")
    K.binarize_weights_t(dev(W), wt_bits=bits, wt_f4=out)
    want = to_f4(np.where(W.T >= 0, 1, -1).astype(np.int8), ld)
    assert np.array_equal(out.cpu().numpy(), want)
    ref = T.zeros((N, ld // 2), dtype=T.uint8, device="cuda")
    K.bits_to_f4(bits, K_, ref)
    assert T.equal(ref, out) and not bool(buf[:256 + ld // 2].any()) and not bool(buf[256 + (N + 1) * (ld // 2):].any())


@pytest.mark.parametrize("B,N", [(8, 8), (130, 100), (257, 513), (1024, 4096), (5, 33)])
def test_bn_sign_emits_nibbles(T, B, N):
    """bnn_bn_sign(act_f4): the nibble form of the int8 activations it writes in the same launch."""
    from mlpcode import kernels as K
    rng = np.random.default_rng(B + N)
    Z = dev(rng.integers(-300, 301, (B, (N + 7) // 8 * 8)).astype(np.int16))
    thr = dev((rng.standard_normal(N) * 100).astype(np.float32))
    sgn = dev(np.where(rng.random(N) < 0.3, -1.0, 1.0).astype(np.float32))
    ld8, ld4 = (N + 15) // 16 * 16, (N + 31) // 32 * 32
    a8 = T.zeros((B, ld8), dtype=T.int8, device="cuda")
    buf, a4 = guarded(T, B, ld4 // 2, fill=0)
    K.bn_sign(Z, B, N, thr, sgn, act_i8=a8, act_f4=a4)
    want = T.where(Z[:, :N].float() * sgn >= thr, 1, -1).to(T.int8)
    assert T.equal(a8[:, :N], want)
    assert np.array_equal(a4.cpu().numpy(), to_f4(want.cpu().numpy(), ld4))
    assert not bool(buf[:256 + ld4 // 2].any()) and not bool(buf[256 + (B + 1) * (ld4 // 2):].any())


F4_SHAPES = [(128, 224, 256), (130, 40, 64), (256, 10, 512), (300, 300, 800), (64, 16, 32), (1, 1, 32),
             (257, 513, 4096), (1000, 4096, 4096), (128, 128, 160), (129, 225, 8192)]


@pytest.mark.parametrize("M,N,K_", F4_SHAPES)
def test_gemm_f4_exact(T, M, N, K_):
    """All four epilogues of the kind against the integer product; the int8 kind on the same operands agrees."""
    from mlpcode import _native as nat, kernels as K
    rng = np.random.default_rng(M * 7 + N * 3 + K_)
    a, b = pm1(rng, (M, K_)), pm1(rng, (N, K_))
    ref = (dev(a).float() @ dev(b).float().T).to(T.int32)       # exact: |sum| <= K < 2^24
    A4, B4 = dev(to_f4(a)), dev(to_f4(b))
    for epi, dt, mult in ((nat.EPI_STORE_S32, T.int32, 4), (nat.EPI_STORE_S16, T.int16, 8)):
        ldd = (N + mult - 1) // mult * mult
        D = T.full((M + 2, ldd), 12345, dtype=dt, device="cuda")
        K.gemm(M=M, N=N, K=K_, A=(A4,), B=(B4,), lda=2 * A4.stride(0), ldb=2 * B4.stride(0), D=D[1:M + 1], ldd=ldd,
               in_dtype=nat.DT_F4, epilogue=epi)
        assert T.equal(D[1:M + 1, :N].to(T.int32), ref)
        assert bool((D[1:M + 1, N:] == 12345).all()) and bool((D[0] == 12345).all()) and bool((D[M + 1] == 12345).all())
    # sign epilogues: thresholds that cut through the distribution, both orientations
    thr = dev((rng.standard_normal(N) * np.sqrt(K_) * 0.5).astype(np.float32))
    sgn = dev(np.where(rng.random(N) < 0.3, -1.0, 1.0).astype(np.float32))
    want = T.where(ref.float() * sgn >= thr, 1, -1).to(T.int8)
    ld8 = (N + 15) // 16 * 16
    bufI, SI = guarded(T, M, ld8)
    K.gemm(M=M, N=N, K=K_, A=(A4,), B=(B4,), lda=2 * A4.stride(0), ldb=2 * B4.stride(0), D=SI.view(T.int8), ldd=ld8,
           in_dtype=nat.DT_F4, epilogue=nat.EPI_SIGN_I8, epi_thr=thr, epi_sgn=sgn)
    assert T.equal(SI.view(T.int8)[:, :N], want) and guards_intact(bufI, M, ld8)
    assert bool((SI[:, N:] == GUARD).all())
    ld4 = (N + 31) // 32 * 32
    buf4, S4 = guarded(T, M, ld4 // 2)
    K.gemm(M=M, N=N, K=K_, A=(A4,), B=(B4,), lda=2 * A4.stride(0), ldb=2 * B4.stride(0), D=S4, ldd=ld4,
           in_dtype=nat.DT_F4, epilogue=nat.EPI_SIGN_F4, epi_thr=thr, epi_sgn=sgn)
    got4 = S4.cpu().numpy()
    n16 = (N + 15) // 16 * 16
    assert np.array_equal(from_f4(got4[:, :n16 // 2], N), want.cpu().numpy()) and guards_intact(buf4, M, ld4 // 2)
    # nibbles in [N, 16 * ceil(N / 16)) are zero (the consumer's K padding); bytes beyond are never written
    assert not from_f4(got4[:, :n16 // 2], n16)[:, N:].any() and np.all(got4[:, n16 // 2:] == GUARD)
    # the int8 kind with the nibble epilogue writes the same bytes
    if K_ % 16 == 0:
        buf8, S8 = guarded(T, M, ld4 // 2)
        K.gemm(M=M, N=N, K=K_, A=(dev(a),), B=(dev(b),), lda=K_, ldb=K_, D=S8, ldd=ld4, in_dtype=nat.DT_I8,
               epilogue=nat.EPI_SIGN_F4, epi_thr=thr, epi_sgn=sgn)
        assert T.equal(S8, S4) and guards_intact(buf8, M, ld4 // 2)


def test_gemm_f4_zero_padding_in_k(T):
    """Zero nibbles inside the K extent contribute nothing (K padding of a 784- or 1000-wide layer)."""
    from mlpcode import _native as nat, kernels as K
    rng = np.random.default_rng(3)
    M, N, Kl, Kp = 200, 300, 1000, 1024
    a, b = pm1(rng, (M, Kl)), pm1(rng, (N, Kl))
    ref = a.astype(np.int32) @ b.astype(np.int32).T
    A4, B4 = dev(to_f4(a, Kp)), dev(to_f4(b, Kp))
    D = T.zeros((M, 304), dtype=T.int32, device="cuda")
    K.gemm(M=M, N=N, K=Kp, A=(A4,), B=(B4,), lda=Kp, ldb=Kp, D=D, ldd=304, in_dtype=nat.DT_F4,
           epilogue=nat.EPI_STORE_S32)
    assert np.array_equal(D.cpu().numpy()[:, :N], ref)


@pytest.mark.parametrize("shared", ["a", "b", "none"])
def test_gemm_f4_batched(T, shared):
    """batch = T trials: shared input with per-trial weights (weight sweep), per-trial input with shared weights
    (image sweep), both per trial; per-batch thresholds (epi_thr_stride)."""
    from mlpcode import _native as nat, kernels as K
    rng = np.random.default_rng(11)
    nb, M, N, K_ = 3, 150, 260, 512
    a = pm1(rng, (1 if shared == "a" else nb, M, K_))
    b = pm1(rng, (1 if shared == "b" else nb, N, K_))
    A4 = dev(np.stack([to_f4(x) for x in a]))
    B4 = dev(np.stack([to_f4(x) for x in b]))
    thr = dev((rng.standard_normal((nb, 264)) * 10).astype(np.float32))
    sgn = dev(np.ones(N, np.float32))
    ld4 = 288
    out = T.zeros((nb, M, ld4 // 2), dtype=T.uint8, device="cuda")
    K.gemm(M=M, N=N, K=K_, A=(A4,), B=(B4,), lda=K_, ldb=K_, D=out, ldd=ld4, in_dtype=nat.DT_F4,
           epilogue=nat.EPI_SIGN_F4, batch=nb, stride_a=0 if shared == "a" else M * K_,
           stride_b=0 if shared == "b" else N * K_, stride_d=M * ld4, epi_thr=thr, epi_sgn=sgn, epi_thr_stride=264)
    for t in range(nb):
        ref = a[0 if shared == "a" else t].astype(np.int32) @ b[0 if shared == "b" else t].astype(np.int32).T
        want = np.where(ref >= thr[t, :N].cpu().numpy(), 1, -1)
        assert np.array_equal(from_f4(out[t].cpu().numpy(), N), want), t


@pytest.mark.parametrize("M,N,K_", [(128, 128, 64), (300, 512, 800), (1000, 384, 256), (77, 32, 64), (200, 10, 128),
                                    (257, 513, 64), (300, 300, 800), (130, 224, 96), (8192, 4096, 512)])
def test_gemm_f4_fused_colstats(T, M, N, K_):
    """BNN_STATS_COLSUM on the FP4 kind (training forward of +-1 layers; all three tile widths, ragged M and N):
    (sum z, sum z^2) per column are exact integers and must equal the int64 sums of the stored result
    (layers.py:81-82); nothing is written outside the result."""
    from mlpcode import _native as nat, kernels as K
    rng = np.random.default_rng(M + N + K_)
    a, b = pm1(rng, (M, K_)), pm1(rng, (N, K_))
    ref = (dev(a).float() @ dev(b).float().T).to(T.int64)
    A4, B4 = dev(to_f4(a)), dev(to_f4(b))
    for epi, dt, mult in ((nat.EPI_STORE_S16, T.int16, 8), (nat.EPI_STORE_S32, T.int32, 4)):
        ldd = (N + mult - 1) // mult * mult
        D = T.full((M + 2, ldd), 12345, dtype=dt, device="cuda")
        sums = T.zeros((2, N), dtype=T.float64, device="cuda")
        K.gemm(M=M, N=N, K=K_, A=(A4,), B=(B4,), lda=2 * A4.stride(0), ldb=2 * B4.stride(0), D=D[1:M + 1], ldd=ldd,
               in_dtype=nat.DT_F4, epilogue=epi, stats=dict(mode=nat.STATS_COLSUM, sums=sums))
        assert T.equal(D[1:M + 1, :N].to(T.int64), ref)
        assert bool((D[1:M + 1, N:] == 12345).all()) and bool((D[0] == 12345).all()) and bool((D[M + 1] == 12345).all())
        assert T.equal(sums[0].to(T.int64), ref.sum(0))
        assert T.equal(sums[1].to(T.int64), (ref * ref).sum(0))


def test_gemm_f4_rejects_bad_shapes(T):
    from mlpcode import _native as nat, kernels as K
    A = T.zeros((128, 64), dtype=T.uint8, device="cuda")
    D = T.zeros((128, 128), dtype=T.int32, device="cuda")
    with pytest.raises(RuntimeError):          # K not a multiple of 32
        K.gemm(M=128, N=128, K=112, A=(A,), B=(A,), lda=128, ldb=128, D=D, ldd=128, in_dtype=nat.DT_F4,
               epilogue=nat.EPI_STORE_S32)
    with pytest.raises(RuntimeError):          # fp32 epilogues belong to the other kinds
        K.gemm(M=128, N=128, K=128, A=(A,), B=(A,), lda=128, ldb=128, D=D.view(T.float32), ldd=128,
               in_dtype=nat.DT_F4, epilogue=nat.EPI_STORE_F32)


def _net(units, seed=0):
    import bnn_oracle as O
    from mlpcode.activation import ActivationFuncs as af
    from mlpcode.network import Network
    params = O.new_params(units, np.random.default_rng(seed))
    rng = np.random.default_rng(seed + 1)
    for bn in params["bn"]:                       # non-trivial running statistics
        n = bn["gamma"].size
        bn["gamma"][:] = 1 + 0.1 * rng.standard_normal(n)
        bn["beta"][:] = 0.1 * rng.standard_normal(n)
        bn["mu"][:] = rng.standard_normal(n)
        bn["sigma"][:] = 1 + rng.random(n)
    nn = Network(units, useGpu=True, binarized=True, useBatchNorm=True)
    nn.compile(lr=0.01, hiddenAf=af.sign, outAf=af.softmax)
    nn.load_weights([w.copy() for w in params["W"]])
    nn.loadBatchNormParameters(*[[b[k].copy() for b in params["bn"]] for k in ("gamma", "beta", "mu", "sigma")])
    return nn, params


def test_predict_identical_with_and_without_fp4(T, oracle, monkeypatch):
    """Network.predict (eval mode): hidden +-1 layers on the FP4 pipe give the bytes the int8 pipe gives, and
    both match the oracle's probabilities."""
    units = [784, 512, 288, 256, 10]
    nn, params = _net(units)
    X = np.random.default_rng(5).random((333, 784), dtype=np.float32) * 2 - 1
    outs = {}
    for flag in ("0", "1"):
        monkeypatch.setenv("BNN_FP4", flag)
        outs[flag] = nn.predict(X).get()
    assert np.array_equal(outs["0"], outs["1"])
    from conftest import rel_err
    assert rel_err(outs["1"], oracle.predict(params, X.copy())) < 1e-5


def test_sweeps_identical_with_and_without_fp4(T, monkeypatch):
    """Weight and image fault sweeps: per-trial correct-counts are integers and must not depend on the pipe."""
    from mlpcode.sweep import ImageFaultSweep, WeightFaultSweep
    from mlpcode.utils import normalize
    units = [784, 512, 256, 10]
    nn, _ = _net(units, seed=4)
    rng = np.random.default_rng(6)
    imgs = rng.integers(0, 256, (500, 784), dtype=np.uint8)
    imgs[rng.random((500, 784)) < 0.7] = 0
    labels = rng.integers(0, 10, 500).astype(np.uint8)
    Xf = rng.random((500, 784), dtype=np.float32) * 2 - 1
    res = {}
    for flag in ("0", "1"):
        monkeypatch.setenv("BNN_FP4", flag)
        w8 = WeightFaultSweep(nn, normalize(imgs, newMin=-1., newMax=1.), labels, trials_per_launch=4)
        wf = WeightFaultSweep(nn, Xf, labels, trials_per_launch=4)
        im = ImageFaultSweep(nn, imgs, labels, trials_per_launch=4, batched=True)
        res[flag] = (w8.clean_accuracy(), w8.run_bernoulli(0.05, 10, 3), w8.run_exactk(lambda t: 1 + 37 * t, 10, 3),
                     wf.run_bernoulli(0.05, 6, 3), im.run(0.05, 10, 9))
    for a, b in zip(res["0"], res["1"]):
        assert np.array_equal(np.asarray(a), np.asarray(b))


def test_training_identical_with_and_without_fp4(T, monkeypatch):
    """Training steps with the wide hidden layers' forward on the FP4 pipe (kernels.f4_train_ok) leave exactly the
    weights, batch-norm parameters and loss the int8 pipe leaves: z is the same integer on both, everything after
    it is shared."""
    units = [784, 512, 384, 256, 10]
    rng = np.random.default_rng(21)
    X = (rng.integers(-16, 17, (4, 256, 784)) / 16).astype(np.float32)
    y = np.eye(10, dtype=np.float32)[rng.integers(0, 10, (4, 256))]
    res = {}
    for flag in ("0", "1"):
        monkeypatch.setenv("BNN_FP4_TRAIN", flag)
        nn, _ = _net(units, seed=8)
        losses = [float(nn.train_step(X[i], y[i])) for i in range(4)]
        nn.predict(X[0], isTraining=True)   # one more training-mode forward (both settings) to look at the operands
        acts = [l.cache["a"] for l in nn._layers[:-1]]
        if flag == "1":   # the path under test is the one that ran
            assert acts[0].f4 is not None and acts[0].i8 is None and acts[1].f4 is not None
            assert acts[2].f4 is not None          # the 10-wide output layer reads nibbles too
        else:
            assert all(a.f4 is None and a.i8 is not None for a in acts)
        res[flag] = (losses, [l.weights.get().copy() for l in nn._layers],
                     [l.batchNormLayer.gamma.get().copy() for l in nn._layers],
                     [l.batchNormLayer.mu.get().copy() for l in nn._layers])
    assert res["0"][0] == res["1"][0]
    for grp in (1, 2, 3):
        for a, b in zip(res["0"][grp], res["1"][grp]):
            assert np.array_equal(a, b)
