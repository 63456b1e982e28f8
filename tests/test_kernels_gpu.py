"""GPU parity of every libbnn_b200 kernel against the oracle, called through the C ABI.

Bar: bit-exact for integer / byte / index work (packing, +-1 GEMMs, masks, flips, batch-norm sign
decisions on identical inputs); for floating point the tolerance is written in each test (north
star: 1e-5 relative)."""
import numpy as np
import pytest

from conftest import rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def T():
    import torch
    from mlpcode import _native, kernels
    _native.load()
    return torch


def dev(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def pm1(rng, shape):
    return np.where(rng.random(shape) < 0.5, -1, 1).astype(np.int8)


def padded(a, mult, dtype=None):
    """Copy a 2-D array into a zero buffer whose pitch is a multiple of `mult` elements."""
    import torch
    r, c = a.shape
    cp = (c + mult - 1) // mult * mult
    t = torch.zeros((r, cp), dtype=dev(a[:1, :1]).dtype if dtype is None else dtype, device="cuda")
    t[:, :c] = dev(a)
    return t


# ------------------------------------------------------------------------------------- A1 packing
@pytest.mark.parametrize("R,C", [(1, 1), (3, 31), (5, 32), (4, 33), (64, 784), (17, 4097)])
def test_pack_sign_rows(T, oracle, R, C):
    from mlpcode import kernels as K
    rng = np.random.default_rng(R * 1000 + C)
    x = rng.standard_normal((R, C)).astype(np.float32)
    x[0, 0] = -0.0
    x[-1, -1] = 0.0
    words = (C + 31) // 32
    bits = T.full((R, words + 1), -1, dtype=T.int32, device="cuda")
    K.pack_sign_rows(dev(x), bits)
    got = bits.cpu().numpy().view(np.uint32)
    assert np.array_equal(got[:, :words], oracle.pack_sign_rows(x))
    assert np.all(got[:, words] == 0xFFFFFFFF)  # pad word beyond ceil(C/32) untouched


@pytest.mark.parametrize("K_,N", [(32, 32), (784, 256), (100, 10), (4096, 70), (33, 129), (130, 260),
                                  (1024, 1028), (257, 4)])
def test_binarize_weights_t(T, oracle, K_, N):
    from mlpcode import kernels as K
    rng = np.random.default_rng(K_ + N)
    W = rng.standard_normal((K_, N)).astype(np.float32)
    W[0, 0], W[1 % K_, 0] = 0.0, -0.0
    words = (K_ + 31) // 32
    i8 = T.zeros((N, (K_ + 15) // 16 * 16), dtype=T.int8, device="cuda")
    f16 = T.zeros((N, (K_ + 7) // 8 * 8), dtype=T.float16, device="cuda")
    bits = T.zeros((N, words), dtype=T.int32, device="cuda")
    K.binarize_weights_t(dev(W), wt_i8=i8, wt_f16=f16, wt_bits=bits)
    ref = oracle.binarize(W).T  # [N, K]
    assert np.array_equal(i8.cpu().numpy()[:, :K_], ref.astype(np.int8))
    assert np.array_equal(f16.cpu().numpy()[:, :K_].astype(np.float32), ref)
    assert np.array_equal(bits.cpu().numpy().view(np.uint32), oracle.c_pack_weights_t(W))
    assert not i8[:, K_:].any() and not f16[:, K_:].any()   # pitch padding past K stays zero
    # round trip through unpack
    back = T.zeros_like(i8)
    K.unpack_bits(bits, K_, out_i8=back)
    assert T.equal(back[:, :K_], i8[:, :K_])


@pytest.mark.parametrize("R,C", [(7, 10), (64, 64), (100, 784), (130, 33)])
def test_split_f16_reconstructs_fp32(T, R, C):
    from mlpcode import kernels as K
    rng = np.random.default_rng(R + C)
    x = (rng.standard_normal((R, C)) * np.exp(rng.standard_normal((R, C)) * 3) * 1e-3).astype(np.float32)
    xd = dev(x)
    amax = T.zeros(1, dtype=T.float32, device="cuda")
    scale = T.zeros(1, dtype=T.float32, device="cuda")
    Cp, Rp = (C + 7) // 8 * 8, (R + 7) // 8 * 8
    hi, lo = (T.zeros((R, Cp), dtype=T.float16, device="cuda") for _ in range(2))
    hit, lot = (T.zeros((C, Rp), dtype=T.float16, device="cuda") for _ in range(2))
    K.absmax_f32(xd, amax)
    assert amax.item() == np.abs(x).max()
    K.split_f16(xd, amax, hi=hi, lo=lo, hi_t=hit, lo_t=lot, scale_out=scale)
    s = scale.item()
    assert s > 0 and np.log2(s) == int(np.log2(s))          # power of two
    assert np.abs(x).max() * s <= 2.0 ** 14                  # head-room below fp16 max
    rec = (hi.double() + lo.double()).cpu().numpy()[:, :C] / s
    # two fp16 pieces carry >= 21 mantissa bits of the scaled value
    tol = 2.0 ** -21 * np.abs(x) + 2.0 ** -24 / s
    assert np.all(np.abs(rec - x.astype(np.float64)) <= tol)
    assert T.equal(hit[:, :R], hi[:, :C].t()) and T.equal(lot[:, :R], lo[:, :C].t())


# ------------------------------------------------------------------------------------- A2 GEMMs
I8_SHAPES = [(128, 256, 128), (130, 40, 200), (256, 10, 512), (300, 300, 784), (64, 16, 32),
             (1, 1, 16), (257, 513, 4096), (512, 4096, 256)]


@pytest.mark.parametrize("impl", ["tcgen05", "simt"])
@pytest.mark.parametrize("M,N,K_", I8_SHAPES)
def test_gemm_i8_exact(T, oracle, impl, M, N, K_):
    from mlpcode import _native as nat, kernels as K
    rng = np.random.default_rng(M * 7 + N * 3 + K_)
    a, b = pm1(rng, (M, K_)), pm1(rng, (N, K_))
    ref = a.astype(np.int32) @ b.astype(np.int32).T
    if M * N * K_ < 2e7:
        assert np.array_equal(oracle.c_gemm_pm1(a, b), ref)
    A, Bm = padded(a, 16), padded(b, 16)
    which = nat.GEMM_TCGEN05 if impl == "tcgen05" else nat.GEMM_SIMT
    for epi, dt, mult in ((nat.EPI_STORE_S32, T.int32, 4), (nat.EPI_STORE_S16, T.int16, 8)):
        ldd = (N + mult - 1) // mult * mult
        D = T.full((M, ldd), 12345, dtype=dt, device="cuda")
        K.gemm(M=M, N=N, K=K_, A=(A,), B=(Bm,), lda=A.stride(0), ldb=Bm.stride(0), D=D, ldd=ldd,
               in_dtype=nat.DT_I8, epilogue=epi, impl=which)
        got = D.cpu().numpy()
        assert np.array_equal(got[:, :N].astype(np.int32), ref)
        assert np.all(got[:, N:] == 12345)               # nothing written past column N


def test_gemm_i8_unaligned_output_and_batch(T):
    from mlpcode import _native as nat, kernels as K
    rng = np.random.default_rng(5)
    nb, M, N, K_ = 3, 200, 70, 160
    a, b = pm1(rng, (nb, M, K_)), pm1(rng, (nb, N, K_))
    ref = np.einsum("bmk,bnk->bmn", a.astype(np.int32), b.astype(np.int32))
    A, Bm = dev(a), dev(b)
    D = T.zeros((nb, M, N + 1), dtype=T.int32, device="cuda")   # pitch 71: scalar store path
    K.gemm(M=M, N=N, K=K_, A=(A,), B=(Bm,), lda=K_, ldb=K_, D=D, ldd=N + 1, in_dtype=nat.DT_I8,
           epilogue=nat.EPI_STORE_S32, batch=nb, stride_a=M * K_, stride_b=N * K_,
           stride_d=M * (N + 1), impl=nat.GEMM_TCGEN05)
    assert np.array_equal(D.cpu().numpy()[:, :, :N], ref)
    # operand A shared by every batch (stride 0): the layout of a fault sweep's first layer
    D2 = T.zeros((nb, M, N + 1), dtype=T.int32, device="cuda")
    K.gemm(M=M, N=N, K=K_, A=(A[0],), B=(Bm,), lda=K_, ldb=K_, D=D2, ldd=N + 1, in_dtype=nat.DT_I8,
           epilogue=nat.EPI_STORE_S32, batch=nb, stride_a=0, stride_b=N * K_,
           stride_d=M * (N + 1), impl=nat.GEMM_TCGEN05)
    ref2 = np.einsum("mk,bnk->bmn", a[0].astype(np.int32), b.astype(np.int32))
    assert np.array_equal(D2.cpu().numpy()[:, :, :N], ref2)


@pytest.mark.parametrize("M,N,K_", [(5, 7, 1), (128, 128, 512), (130, 200, 784), (64, 10, 4096),
                                    (300, 129, 33)])
def test_gemm_popc_exact(T, oracle, M, N, K_):
    from mlpcode import kernels as K
    rng = np.random.default_rng(M + N + K_)
    a, b = pm1(rng, (M, K_)), pm1(rng, (N, K_))
    ref = a.astype(np.int32) @ b.astype(np.int32).T
    ab = oracle.pack_sign_rows(a.astype(np.float32))
    bb = oracle.pack_sign_rows(b.astype(np.float32))
    if M * N * K_ < 2e7:
        assert np.array_equal(oracle.c_gemm_xnor(ab, bb, K_), ref)
    A, Bm = dev(ab.view(np.int32)), dev(bb.view(np.int32))
    D = T.zeros((M, N), dtype=T.int32, device="cuda")
    K.gemm_popc(A, Bm, M, N, K_, D, out_s16=False)
    assert np.array_equal(D.cpu().numpy(), ref)
    D16 = T.zeros((M, N), dtype=T.int16, device="cuda")
    K.gemm_popc(A, Bm, M, N, K_, D16, out_s16=True)
    assert np.array_equal(D16.cpu().numpy().astype(np.int32), ref)
    # the pack kernel feeds the same words
    bits = T.zeros((M, (K_ + 31) // 32), dtype=T.int32, device="cuda")
    K.pack_i8_rows(dev(a), K_, bits)
    assert np.array_equal(bits.cpu().numpy().view(np.uint32), ab)


def test_binary_gemm_variants_agree_at_full_size(T):
    """Config-2 layer shape (8192 x 4096 x 4096): tcgen05 kind::i8 and XOR+popc must agree bit for
    bit, and satisfy the parity invariant z == K (mod 2), |z| <= K."""
    from mlpcode import _native as nat, kernels as K
    M, N, K_ = 8192, 4096, 4096
    g = T.Generator(device="cuda").manual_seed(1)
    A = (T.randint(0, 2, (M, K_), generator=g, device="cuda", dtype=T.int8) * 2 - 1)
    Bm = (T.randint(0, 2, (N, K_), generator=g, device="cuda", dtype=T.int8) * 2 - 1)
    Z1 = T.zeros((M, N), dtype=T.int16, device="cuda")
    K.gemm(M=M, N=N, K=K_, A=(A,), B=(Bm,), lda=K_, ldb=K_, D=Z1, ldd=N, in_dtype=nat.DT_I8,
           epilogue=nat.EPI_STORE_S16, impl=nat.GEMM_TCGEN05)
    ab = T.zeros((M, K_ // 32), dtype=T.int32, device="cuda")
    bb = T.zeros((N, K_ // 32), dtype=T.int32, device="cuda")
    K.pack_i8_rows(A, K_, ab)
    K.pack_i8_rows(Bm, K_, bb)
    Z2 = T.zeros((M, N), dtype=T.int16, device="cuda")
    K.gemm_popc(ab, bb, M, N, K_, Z2, out_s16=True)
    assert T.equal(Z1, Z2)
    assert int((Z1.int() & 1).sum()) == 0 and int(Z1.abs().max()) <= K_
    # linearity spot check against a direct fp64 product of a few rows/columns
    rows = T.tensor([0, 1, 4095, 8191], device="cuda")
    ref = A[rows].double() @ Bm[:64].double().t()
    assert T.equal(Z1[rows][:, :64].double(), ref)


F16_SHAPES = [(128, 256, 64), (200, 130, 784), (256, 10, 300), (784, 64, 32), (300, 4096, 10),
              (16, 16, 16), (1000, 520, 2048)]


@pytest.mark.parametrize("impl", ["tcgen05", "simt"])
@pytest.mark.parametrize("M,N,K_", F16_SHAPES)
def test_gemm_f16_split_matches_fp64(T, impl, M, N, K_):
    """Real x real product through 3 fp16 split passes: error must be at fp32-SGEMM level
    (<= 1e-6 norm-wise, 10x below the 1e-5 parity tolerance)."""
    from mlpcode import _native as nat, kernels as K
    rng = np.random.default_rng(M + 2 * N + 3 * K_)
    a = (rng.standard_normal((M, K_)) * 1e-3 * np.exp(rng.standard_normal((M, K_)))).astype(np.float32)
    b = (rng.standard_normal((N, K_)) / 32).astype(np.float32)
    ref = a.astype(np.float64) @ b.astype(np.float64).T
    Kp = (K_ + 7) // 8 * 8
    ops = {}
    for name, x in (("a", a), ("b", b)):
        xd = dev(x)
        amax = T.zeros(1, dtype=T.float32, device="cuda")
        sc = T.zeros(1, dtype=T.float32, device="cuda")
        hi, lo = (T.zeros((x.shape[0], Kp), dtype=T.float16, device="cuda") for _ in range(2))
        K.absmax_f32(xd, amax)
        K.split_f16(xd, amax, hi=hi, lo=lo, scale_out=sc)
        ops[name] = (hi, lo, sc)
    which = nat.GEMM_TCGEN05 if impl == "tcgen05" else nat.GEMM_SIMT
    ldd = (N + 3) // 4 * 4
    D = T.zeros((M, ldd), dtype=T.float32, device="cuda")
    K.gemm(M=M, N=N, K=K_, A=ops["a"][:2], B=ops["b"][:2], lda=Kp, ldb=Kp, D=D, ldd=ldd,
           in_dtype=nat.DT_F16, epilogue=nat.EPI_STORE_F32, passes=((0, 0), (0, 1), (1, 0)),
           sa=ops["a"][2], sb=ops["b"][2], impl=which)
    got = D.cpu().numpy()[:, :N]
    # 3e-6: the fp32 accumulation floor of the tensor pipe (and of the sequential SIMT sum); the fp16
    # split itself contributes < 1e-7.  Parity tolerance is 1e-5.
    assert rel_err(got, ref) < 3e-6
    # fused SGD epilogue: W -= lr * (A B^T)   (layers.py:268)
    for W0 in (np.zeros((M, N), np.float32), rng.standard_normal((M, N)).astype(np.float32)):
        Wd = dev(W0).clone()
        K.gemm(M=M, N=N, K=K_, A=ops["a"][:2], B=ops["b"][:2], lda=Kp, ldb=Kp, D=Wd, ldd=N,
               in_dtype=nat.DT_F16, epilogue=nat.EPI_SGD_F32, passes=((0, 0), (0, 1), (1, 0)),
               sa=ops["a"][2], sb=ops["b"][2], host_scale=0.01, impl=which)
        want = W0.astype(np.float64) - 0.01 * ref
        assert rel_err(Wd.cpu().numpy(), want) < 3e-6


@pytest.mark.parametrize("M,N,K_", [(128, 256, 128), (300, 512, 784), (1000, 128, 256), (77, 32, 64),
                                    (8192, 4096, 512)])
def test_gemm_fused_colstats_epilogue(T, M, N, K_):
    """BNN_STATS_COLSUM: the +-1 GEMM's epilogue accumulates (sum z, sum z^2) per column — exact
    integers, so it must equal the int64 column sums of the result bit for bit (layers.py:81-82)."""
    from mlpcode import _native as nat, kernels as K
    assert K.fused_stats_ok(N)
    rng = np.random.default_rng(M + N + K_)
    a, b = pm1(rng, (M, K_)), pm1(rng, (N, K_))
    A, Bm = padded(a, 16), padded(b, 16)
    ref = (dev(a).float() @ dev(b).float().T).to(T.int64)        # exact: |z| <= K < 2^24
    for epi, dt in ((nat.EPI_STORE_S16, T.int16), (nat.EPI_STORE_S32, T.int32)):
        D = T.zeros((M, N), dtype=dt, device="cuda")
        sums = T.zeros((2, N), dtype=T.float64, device="cuda")
        K.gemm(M=M, N=N, K=K_, A=(A,), B=(Bm,), lda=A.stride(0), ldb=Bm.stride(0), D=D, ldd=N,
               in_dtype=nat.DT_I8, epilogue=epi, stats=dict(mode=nat.STATS_COLSUM, sums=sums))
        assert T.equal(D.to(T.int64), ref)
        assert T.equal(sums[0].to(T.int64), ref.sum(0))
        assert T.equal(sums[1].to(T.int64), (ref * ref).sum(0))
    # and it agrees with the stand-alone kernel it replaces
    sums2 = T.zeros((2, N), dtype=T.float64, device="cuda")
    K.colstats(D, M, N, sums2)
    assert T.equal(sums, sums2)


@pytest.mark.parametrize("M,N,K_,zdt", [(128, 256, 64, "s16"), (300, 512, 200, "f32"), (1000, 128, 96, "s16"),
                                        (4096, 4096, 256, "s16")])
def test_gemm_fused_bn_bwd_reduce_epilogue(T, M, N, K_, zdt):
    """BNN_STATS_BN_BWD: the dX GEMM's epilogue accumulates sum(d), sum(d*(z-mu)), max|d|, max|z-mu| of
    its own output d against the layer below's Z — the reductions of layers.py:107-115 that
    bnn_bn_bwd_reduce would otherwise re-read d and Z for.  Tolerance 1e-6 (fp32 32-row partials,
    fp64 across; the stand-alone kernel uses 16-row partials)."""
    from mlpcode import _native as nat, kernels as K
    rng = np.random.default_rng(M * 3 + N + K_)
    a = (rng.standard_normal((M, K_)) * 1e-3).astype(np.float32)
    b = (rng.standard_normal((N, K_)) / 16).astype(np.float32)
    Kp = (K_ + 7) // 8 * 8
    ops = {}
    for name, x in (("a", a), ("b", b)):
        xd = dev(x)
        amax = T.zeros(1, dtype=T.float32, device="cuda")
        sc = T.zeros(1, dtype=T.float32, device="cuda")
        hi, lo = (T.zeros((x.shape[0], Kp), dtype=T.float16, device="cuda") for _ in range(2))
        K.absmax_f32(xd, amax)
        K.split_f16(xd, amax, hi=hi, lo=lo, scale_out=sc)
        ops[name] = (hi, lo, sc)
    if zdt == "s16":
        Z = dev(rng.integers(-K_, K_ + 1, (M, N)).astype(np.int16))
    else:
        Z = dev(rng.standard_normal((M, N)).astype(np.float32) * 20)
    mu = dev(rng.standard_normal(N).astype(np.float32))
    D = T.zeros((M, N), dtype=T.float32, device="cuda")
    red = T.zeros((2, N), dtype=T.float64, device="cuda")
    red_max = T.zeros((2, N), dtype=T.int32, device="cuda")
    K.gemm(M=M, N=N, K=K_, A=ops["a"][:2], B=ops["b"][:2], lda=Kp, ldb=Kp, D=D, ldd=N,
           in_dtype=nat.DT_F16, epilogue=nat.EPI_STORE_F32, passes=((0, 0), (0, 1), (1, 0)),
           sa=ops["a"][2], sb=ops["b"][2],
           stats=dict(mode=nat.STATS_BN_BWD, sums=red, max=red_max, Z=Z, mu=mu))
    # same product without the fused reductions: the stored result must not change
    D0 = T.zeros_like(D)
    K.gemm(M=M, N=N, K=K_, A=ops["a"][:2], B=ops["b"][:2], lda=Kp, ldb=Kp, D=D0, ldd=N,
           in_dtype=nat.DT_F16, epilogue=nat.EPI_STORE_F32, passes=((0, 0), (0, 1), (1, 0)),
           sa=ops["a"][2], sb=ops["b"][2])
    assert T.equal(D, D0)
    d64, xm = D.double(), Z.double() - mu.double()
    want = T.stack([d64.sum(0), (d64 * (Z.float() - mu).double()).sum(0)])
    scale = T.stack([d64.abs().sum(0), (d64 * xm).abs().sum(0)])
    assert float(((red - want).abs() / scale).max()) < 1e-6
    assert T.equal(red_max[0].view(T.float32), D.abs().amax(0))
    assert T.equal(red_max[1].view(T.float32), (Z.float() - mu).abs().amax(0))
    # the stand-alone kernel on the same delta agrees
    red2 = T.zeros_like(red)
    red_max2 = T.zeros_like(red_max)
    K.bn_bwd_reduce(D, Z, M, N, mu, red2, red_max2)
    assert float(((red - red2).abs() / scale).max()) < 1e-6
    assert T.equal(red_max, red_max2)


@pytest.mark.parametrize("M,N,K_", [(128, 256, 128), (300, 512, 1000), (4096, 256, 8192)])
def test_gemm_i8_digit_split(T, M, N, K_):
    """Fixed-point product on the int8 tensor pipe (the dW path): A = +-1 (and 64 * A as a second piece),
    B = three signed-digit matrices of q = q1 + 256 * (q2 + 127 * q3).  Passes (A, q1) -> LOW accumulator,
    (127A, q3) + (A, q2) -> HIGH accumulator, epilogue low + 256 * high, times the column's power-of-two
    scale, times host_scale.  The integer sums are exact; only the fp32 combine rounds (2^-24)."""
    from mlpcode import _native as nat, kernels as K
    rng = np.random.default_rng(M + N + K_)
    x = pm1(rng, (M, K_))
    q1 = rng.integers(-128, 128, (N, K_)).astype(np.int8)
    q2 = rng.integers(-63, 64, (N, K_)).astype(np.int8)
    q3 = rng.integers(-128, 128, (N, K_)).astype(np.int8)
    q = q1.astype(np.int64) + 256 * (q2.astype(np.int64) + 127 * q3.astype(np.int64))
    ref = (dev(x).double() @ dev(q).double().T).cpu().numpy()          # exact in fp64 (|sum| < 2^53)
    cs = (2.0 ** -rng.integers(0, 12, N)).astype(np.float32)
    Kp = (K_ + 15) // 16 * 16
    A1, A64 = padded(x, 16), padded((127 * x.astype(np.int16)).astype(np.int8), 16)
    Bs = [padded(b, 16) for b in (q1, q3, q2)]
    lr = 0.05
    for epi in (nat.EPI_STORE_F32, nat.EPI_SGD_F32):
        W0 = rng.standard_normal((M, N)).astype(np.float32) if epi == nat.EPI_SGD_F32 else np.zeros((M, N), np.float32)
        D = dev(W0).clone()
        K.gemm(M=M, N=N, K=K_, A=(A1, A64), B=Bs, lda=Kp, ldb=Kp, D=D, ldd=N, in_dtype=nat.DT_I8,
               epilogue=epi, passes=((0, 0), (1, 1), (0, 2)), hi_from_pass=1, col_scale=dev(cs),
               host_scale=lr)
        upd = lr * ref * cs[None, :].astype(np.float64)
        want = W0 - upd if epi == nat.EPI_SGD_F32 else upd
        got = D.cpu().numpy().astype(np.float64)
        assert rel_err(got - (W0 if epi == nat.EPI_SGD_F32 else 0), want - (W0 if epi == nat.EPI_SGD_F32 else 0)) < (
            2e-7 if epi == nat.EPI_STORE_F32 else 1e-5)   # SGD: cancellation against W0's ulp
    # a ragged width is refused, not mis-computed
    with pytest.raises(nat.BnnError):
        K.gemm(M=M, N=N - 8, K=K_, A=(A1, A64), B=Bs, lda=Kp, ldb=Kp, D=D, ldd=N, in_dtype=nat.DT_I8,
               epilogue=nat.EPI_STORE_F32, passes=((0, 0), (1, 1), (0, 2)), hi_from_pass=1)


def test_gemm_fused_stats_refuses_ragged_widths(T):
    from mlpcode import _native as nat, kernels as K
    assert not K.fused_stats_ok(130) and not K.fused_stats_ok(10) and K.fused_stats_ok(4096)
    a = T.ones((64, 64), dtype=T.int8, device="cuda")
    D = T.zeros((64, 136), dtype=T.int16, device="cuda")
    sums = T.zeros((2, 130), dtype=T.float64, device="cuda")
    b = T.ones((130, 64), dtype=T.int8, device="cuda")
    with pytest.raises(nat.BnnError) as e:
        K.gemm(M=64, N=130, K=64, A=(a,), B=(b,), lda=64, ldb=64, D=D, ldd=136, in_dtype=nat.DT_I8,
               epilogue=nat.EPI_STORE_S16, stats=dict(mode=nat.STATS_COLSUM, sums=sums))
    assert e.value.code == nat.BNN_ERR_UNSUPPORTED


def test_split_gemm_precision_at_full_size(T):
    """dW- and dX-shaped products of config 2 (contraction over the 8192-sample batch / over 4096
    units) against an fp64 product on the device: the error must stay at fp32-SGEMM level."""
    from mlpcode import _native as nat, kernels as K
    g = T.Generator(device="cuda").manual_seed(3)
    Bsz, Kd, N = 8192, 4096, 4096

    def split(x):
        amax = T.zeros(1, dtype=T.float32, device="cuda")
        sc = T.zeros(1, dtype=T.float32, device="cuda")
        hi, lo = (T.zeros(x.shape, dtype=T.float16, device="cuda") for _ in range(2))
        K.absmax_f32(x, amax)
        K.split_f16(x, amax, hi=hi, lo=lo, scale_out=sc)
        return hi, lo, sc

    # dW = X^T delta: X +-1 (exact in fp16), delta heavy tailed like a back-propagated gradient
    Xt = (T.randint(0, 2, (Kd, Bsz), generator=g, device="cuda") * 2 - 1).to(T.float16)
    dT = T.randn((N, Bsz), generator=g, device="cuda") * 1e-6 * T.exp(
        2 * T.randn((N, Bsz), generator=g, device="cuda"))
    d_hi, d_lo, d_s = split(dT)
    dW = T.zeros((Kd, N), dtype=T.float32, device="cuda")
    K.gemm(M=Kd, N=N, K=Bsz, A=(Xt,), B=(d_hi, d_lo), lda=Bsz, ldb=Bsz, D=dW, ldd=N,
           in_dtype=nat.DT_F16, epilogue=nat.EPI_STORE_F32, passes=((0, 0), (0, 1)), sb=d_s,
           impl=nat.GEMM_TCGEN05)
    rows = slice(0, 512)
    ref = Xt[rows].double() @ dT.double().t()
    err_dw = float((dW[rows].double() - ref).norm() / ref.norm())
    # dX = delta W^T: both real
    dl = dT[:, :Bsz].t().contiguous()                      # [B, N]
    W = T.randn((Kd, N), generator=g, device="cuda") / 64
    a_hi, a_lo, a_s = split(dl)
    w_hi, w_lo, w_s = split(W)
    dX = T.zeros((Bsz, Kd), dtype=T.float32, device="cuda")
    K.gemm(M=Bsz, N=Kd, K=N, A=(a_hi, a_lo), B=(w_hi, w_lo), lda=N, ldb=N, D=dX, ldd=Kd,
           in_dtype=nat.DT_F16, epilogue=nat.EPI_STORE_F32, passes=((0, 0), (0, 1), (1, 0)),
           sa=a_s, sb=w_s, impl=nat.GEMM_TCGEN05)
    ref = dl[rows].double() @ W.double().t()
    err_dx = float((dX[rows].double() - ref).norm() / ref.norm())
    sg = dl[rows] @ W.t()                                 # cuBLAS fp32 SGEMM of the same product
    err_sgemm = float((sg.double() - ref).norm() / ref.norm())
    print(f"split-GEMM error vs fp64: dW {err_dw:.3e}  dX {err_dx:.3e}  (cuBLAS SGEMM {err_sgemm:.3e})")
    assert err_dw < 5e-6 and err_dx < 5e-6               # tolerance 1e-5 with 2x margin


def test_gemm_rejects_bad_arguments(T):
    from mlpcode import _native as nat, kernels as K
    A = T.zeros((128, 100), dtype=T.int8, device="cuda")        # pitch 100 bytes: not 16-aligned
    D = T.zeros((128, 128), dtype=T.int32, device="cuda")
    with pytest.raises(nat.BnnError) as e:
        K.gemm(M=128, N=128, K=100, A=(A,), B=(A,), lda=100, ldb=100, D=D, ldd=128,
               in_dtype=nat.DT_I8, epilogue=nat.EPI_STORE_S32, impl=nat.GEMM_TCGEN05)
    assert e.value.code == nat.BNN_ERR_ARG and "16 bytes" in str(e.value)
    with pytest.raises(nat.BnnError):
        K.gemm(M=0, N=128, K=128, A=(A,), B=(A,), lda=128, ldb=128, D=D, ldd=128,
               in_dtype=nat.DT_I8, epilogue=nat.EPI_STORE_S32)


# ------------------------------------------------------------------------------------- A4-A6 BN forward
@pytest.mark.parametrize("B,N,zdt", [(48, 21, "s16"), (300, 130, "s16"), (64, 64, "f32"),
                                     (257, 10, "f32"), (1, 33, "s32")])
def test_colstats_and_bn_apply(T, oracle, B, N, zdt):
    from mlpcode import kernels as K
    rng = np.random.default_rng(B + N)
    if zdt == "f32":
        Z = (rng.standard_normal((B, N)) * 20).astype(np.float32)
        Zd = dev(Z)
    else:
        Z = (rng.integers(-200, 201, (B, N)) * 2).astype(np.float32)
        Zd = dev(Z.astype(np.int16 if zdt == "s16" else np.int32))
    sums = T.zeros((2, N), dtype=T.float64, device="cuda")
    mu, var = (T.zeros(N, dtype=T.float32, device="cuda") for _ in range(2))
    K.colstats(Zd, B, N, sums)
    z64 = Z.astype(np.float64)
    s = sums.cpu().numpy()
    if zdt == "f32":   # fp32 partial sums over <= 16 rows per thread, fp64 across threads
        assert rel_err(s[0], z64.sum(0)) < 1e-6 and rel_err(s[1], (z64 ** 2).sum(0)) < 1e-6
    else:
        assert np.array_equal(s[0], z64.sum(0)) and np.array_equal(s[1], (z64 ** 2).sum(0))
    K.bn_stats_finalize(sums, N, B, mu, var)
    assert np.allclose(mu.cpu().numpy(), z64.mean(0), rtol=1e-6, atol=1e-6)
    assert np.allclose(var.cpu().numpy(), z64.var(0), rtol=1e-6, atol=1e-6)
    # parity with NumPy's own fp32 statistics (what the reference computes) within 1e-5
    assert rel_err(mu.cpu().numpy(), Z.mean(axis=0)) < 1e-5 or np.abs(Z.mean(axis=0)).max() < 1e-3
    assert rel_err(var.cpu().numpy(), Z.var(axis=0)) < 1e-5

    gamma = (rng.random(N) + 0.5).astype(np.float32)
    gamma[0] = -gamma[0]
    beta = (rng.standard_normal(N) * 0.1).astype(np.float32)
    mu_h, var_h = mu.cpu().numpy(), var.cpu().numpy()
    # identical inputs + identical fp32 op order => identical bits (layers.py:84-86 / 94-97)
    ref = oracle.bn_forward_eval(Z.copy(), gamma, beta, mu_h, var_h)
    out = T.zeros((B, N), dtype=T.float32, device="cuda")
    i8 = T.zeros((B, (N + 15) // 16 * 16), dtype=T.int8, device="cuda")
    t16 = T.zeros((N, (B + 7) // 8 * 8), dtype=T.float16, device="cuda")
    bits = T.zeros((B, (N + 31) // 32), dtype=T.int32, device="cuda")
    K.bn_apply(Zd, B, N, mu, var, dev(gamma), dev(beta), 1e-4, out_f32=out, act_i8=i8, act_t16=t16,
               act_bits=bits)
    assert np.array_equal(out.cpu().numpy(), ref)
    sgn = oracle.unitstep(ref)
    assert np.array_equal(i8.cpu().numpy()[:, :N], sgn.astype(np.int8))
    assert np.array_equal(t16.cpu().numpy()[:, :B].astype(np.float32), sgn.T)
    assert np.array_equal(bits.cpu().numpy().view(np.uint32), oracle.pack_sign_rows(ref))
    # the threshold form of batch-norm + sign (one compare per element) decides identically, also
    # for gamma < 0 (column 0), gamma == 0 and a beta that pins the sign
    for variant in range(2):
        gam = gamma.copy()
        if variant == 1 and N > 2:
            gam[1], beta[1] = 0.0, -0.25
            gam[2], beta[2] = 0.0, 0.0
        ref_v = oracle.bn_forward_eval(Z.copy(), gam, beta, mu_h, var_h)
        thr, sg = (T.zeros(N, dtype=T.float32, device="cuda") for _ in range(2))
        K.bn_sign_thresholds(mu, var, dev(gam), dev(beta), 1e-4, N, thr, sg)
        i8b, t16b, bitsb = T.zeros_like(i8), T.zeros_like(t16), T.zeros_like(bits)
        t8, t8x127 = (T.zeros((N, (B + 15) // 16 * 16), dtype=T.int8, device="cuda") for _ in range(2))
        K.bn_sign(Zd, B, N, thr, sg, act_i8=i8b, act_t16=t16b, act_bits=bitsb, act_t8=t8, act_t8x127=t8x127)
        want = oracle.unitstep(ref_v)
        assert np.array_equal(i8b.cpu().numpy()[:, :N], want.astype(np.int8))
        assert np.array_equal(t16b.cpu().numpy()[:, :B].astype(np.float32), want.T)
        assert np.array_equal(bitsb.cpu().numpy().view(np.uint32), oracle.pack_sign_rows(ref_v))
        # transposed int8 copies (+-1 and +-127): the A pieces of the int8 dW product; padding untouched
        assert np.array_equal(t8.cpu().numpy()[:, :B], want.T.astype(np.int8))
        assert np.array_equal(t8x127.cpu().numpy()[:, :B], (127 * want.T).astype(np.int8))
        assert not t8[:, B:].any() and not t8x127[:, B:].any()


@pytest.mark.parametrize("impl", ["tcgen05", "simt"])
@pytest.mark.parametrize("M,N,K_", [(128, 256, 128), (300, 130, 784), (64, 40, 4096), (1000, 520, 256)])
def test_gemm_fused_bn_sign_epilogue(T, oracle, impl, M, N, K_):
    """BNN_EPI_SIGN_I8: the +-1 GEMM writing sign(BN(z)) directly equals GEMM -> Z -> BN -> sign."""
    from mlpcode import _native as nat, kernels as K
    rng = np.random.default_rng(M + N + K_)
    a, b = pm1(rng, (M, K_)), pm1(rng, (N, K_))
    z = (a.astype(np.int32) @ b.astype(np.int32).T).astype(np.float32)
    mu = (rng.standard_normal(N) * 3).astype(np.float32)
    var = (rng.random(N) * K_ + 1).astype(np.float32)
    gamma = (rng.random(N) + 0.5).astype(np.float32)
    gamma[::7] *= -1
    beta = (rng.standard_normal(N) * 0.3).astype(np.float32)
    want = oracle.unitstep(oracle.bn_forward_eval(z.copy(), gamma, beta, mu, var)).astype(np.int8)
    thr, sg = (T.zeros(N, dtype=T.float32, device="cuda") for _ in range(2))
    K.bn_sign_thresholds(dev(mu), dev(var), dev(gamma), dev(beta), 1e-4, N, thr, sg)
    A, Bm = padded(a, 16), padded(b, 16)
    Np = (N + 15) // 16 * 16
    out = T.full((M, Np), 7, dtype=T.int8, device="cuda")
    K.gemm(M=M, N=N, K=K_, A=(A,), B=(Bm,), lda=A.stride(0), ldb=Bm.stride(0), D=out, ldd=Np,
           in_dtype=nat.DT_I8, epilogue=nat.EPI_SIGN_I8, epi_thr=thr, epi_sgn=sg,
           impl=nat.GEMM_TCGEN05 if impl == "tcgen05" else nat.GEMM_SIMT)
    got = out.cpu().numpy()
    assert np.array_equal(got[:, :N], want)
    assert np.all(got[:, N:] == 7)


def test_softmax_xent(T, golden, oracle):
    from mlpcode import kernels as K
    g = golden("elementwise")
    logits, y = g["logits"], g["onehot"]
    B, C = logits.shape
    probs, grad = (T.zeros((B, C), dtype=T.float32, device="cuda") for _ in range(2))
    rows = T.zeros(B, dtype=T.float32, device="cuda")
    tot = T.zeros(1, dtype=T.float64, device="cuda")
    K.softmax_xent(dev(logits), B, C, onehot=dev(y), probs=probs, loss_rows=rows, grad=grad,
                   count=B, loss_sum=tot)
    sm = oracle.softmax(logits.copy())
    loss = oracle.cross_entropy_loss(sm, y)
    gr = oracle.cross_entropy_derivative_softmax(sm.copy(), y)
    assert np.allclose(probs.cpu().numpy(), sm, rtol=1e-5, atol=1e-9)       # tolerance: 1e-5 rel
    assert np.allclose(rows.cpu().numpy(), loss, rtol=1e-5, atol=1e-7)
    assert np.allclose(grad.cpu().numpy(), gr, rtol=1e-5, atol=1e-9)
    assert abs(tot.item() - float(loss.astype(np.float64).sum())) < 1e-5 * abs(loss.sum())
    # the lower clip of loss.py:32: a logit so small its probability underflows to 0
    lg = np.zeros((2, 10), np.float32)
    lg[0, 3] = -200.0
    p2 = T.zeros((2, 10), dtype=T.float32, device="cuda")
    K.softmax_xent(dev(lg), 2, 10, probs=p2)
    assert p2[0, 3].item() == np.float32(1e-15)


# ------------------------------------------------------------------------------------- A8 BN backward
def test_bn_backward_against_reference_fixture(T, golden):
    from mlpcode import kernels as K
    g = golden("batchnorm")
    Z, delta = g["Z"], g["delta"]
    B, N = Z.shape
    Zd, dd = dev(Z.astype(np.int16)), dev(delta)
    sums = T.zeros((2, N), dtype=T.float64, device="cuda")
    mu, var = (T.zeros(N, dtype=T.float32, device="cuda") for _ in range(2))
    K.colstats(Zd, B, N, sums)
    K.bn_stats_finalize(sums, N, B, mu, var)
    assert rel_err(mu.cpu().numpy(), g["batch_mu"]) < 1e-6
    assert rel_err(var.cpu().numpy(), g["batch_var"]) < 1e-6
    gamma, beta = dev(g["gamma"]).clone(), dev(g["beta"]).clone()
    mu_run, sig_run = dev(g["mu_run"]).clone(), dev(g["sigma_run"]).clone()
    red = T.zeros((2, N), dtype=T.float64, device="cuda")
    red_max = T.zeros((2, N), dtype=T.int32, device="cuda")
    coef = T.zeros((3, N), dtype=T.float32, device="cuda")
    bound = T.zeros(1, dtype=T.float32, device="cuda")
    dgam, dbet = (T.zeros(N, dtype=T.float32, device="cuda") for _ in range(2))
    K.bn_bwd_reduce(dd, Zd, B, N, mu, red, red_max)
    colq = T.zeros((2, N), dtype=T.float32, device="cuda")
    K.bn_bwd_finalize(red, red_max, N, B, mu, var, 1e-4, gamma, beta, mu_run, sig_run,
                      float(g["lr"]), 0.9, sums, coef, bound, dgam, dbet, col_quant=colq)
    Np, Bp = (N + 7) // 8 * 8, (B + 7) // 8 * 8
    out = T.zeros((B, N), dtype=T.float32, device="cuda")
    dhi, dlo = (T.zeros((B, Np), dtype=T.float16, device="cuda") for _ in range(2))
    thi, tlo = (T.zeros((N, Bp), dtype=T.float16, device="cuda") for _ in range(2))
    sc = T.zeros(1, dtype=T.float32, device="cuda")
    K.bn_bwd_apply(dd, Zd, B, N, mu, coef, bound, out_f32=out, d_hi=dhi, d_lo=dlo, dt_hi=thi,
                   dt_lo=tlo, scale_out=sc)
    nd = out.cpu().numpy()
    assert rel_err(nd, g["new_delta"]) < 1e-5                              # tolerance: 1e-5 rel
    assert rel_err(dgam.cpu().numpy(), g["dgamma"]) < 1e-5
    assert rel_err(dbet.cpu().numpy(), g["dbeta"]) < 1e-5
    for name, t in (("gamma", gamma), ("beta", beta), ("mu", mu_run), ("sigma", sig_run)):
        assert rel_err(t.cpu().numpy(), g[f"{name}_after"]) < 1e-6, name
    # the split pieces carry delta' at fp32 precision and never overflow fp16
    s = sc.item()
    assert np.abs(nd).max() * s <= 2.0 ** 15 and bound.item() >= np.abs(nd).max()
    rec = (dhi.double() + dlo.double()).cpu().numpy()[:, :N] / s
    assert rel_err(rec, nd) < 1e-6
    assert T.equal(thi[:, :B], dhi[:, :N].t()) and T.equal(tlo[:, :B], dlo[:, :N].t())
    # fixed-point form (int8 dW): per-column scales s with |delta' * s| <= 4.1e6 (22.97 bits of the
    # column's bound, which sits within a few per cent of max |delta'|), digits
    # q1 + 256 * (q2 + 127 * q3) = rn(delta' * s) (+-1: the kernel's FMA chain vs this product), and
    # the same row-major fp16 pieces as the fp16 form
    cq = colq.cpu().numpy()
    assert np.abs(cq[0].astype(np.float64) * cq[1] - 1).max() < 2e-7
    assert np.all(np.abs(nd).max(0) * cq[0] <= 4.1e6 + 1) and np.all(np.abs(nd).max(0) * cq[0] > 4.1e6 / 2)
    Bq = (B + 15) // 16 * 16
    q1, q3, q2 = (T.zeros((N, Bq), dtype=T.int8, device="cuda") for _ in range(3))
    dhi2, dlo2 = T.zeros_like(dhi), T.zeros_like(dlo)
    K.bn_bwd_apply_q(dd, Zd, B, N, coef, bound, colq, q1, q3, q2, d_hi=dhi2, d_lo=dlo2, scale_out=sc)
    assert T.equal(dhi2, dhi) and T.equal(dlo2, dlo)
    d1, d2, d3 = (t.cpu().numpy()[:, :B].astype(np.int64) for t in (q1, q2, q3))
    assert np.abs(d2).max() <= 63 and np.abs(d3).max() <= 127
    q = (d1 + 256 * (d2 + 127 * d3)).T
    assert np.abs(q - np.rint(nd.astype(np.float64) * cq[0])).max() <= 1
    assert not q1[:, B:].any() and not q2[:, B:].any() and not q3[:, B:].any()


def test_sgd_and_count_correct(T):
    from mlpcode import kernels as K
    rng = np.random.default_rng(3)
    p, g = rng.standard_normal(10007).astype(np.float32), rng.standard_normal(10007).astype(np.float32)
    pd = dev(p).clone()
    K.sgd_update(pd, dev(g), np.float32(0.01))
    assert np.array_equal(pd.cpu().numpy(), p - np.float32(0.01) * g)     # weights -= lr*dw, exact
    g3 = rng.standard_normal((3, 10007)).astype(np.float32)               # split-K partial sums
    pd = dev(p).clone()
    K.sgd_update(pd, dev(g3), np.float32(0.01), parts=3)
    assert np.array_equal(pd.cpu().numpy(), p - np.float32(0.01) * ((g3[0] + g3[1]) + g3[2]))
    scores = rng.standard_normal((1000, 10)).astype(np.float32)
    scores[5, 2] = scores[5, 7] = 99.0                                     # tie: first maximum wins
    labels = rng.integers(0, 10, 1000).astype(np.int32)
    cnt = T.zeros(1, dtype=T.int32, device="cuda")
    K.count_correct(dev(scores), 1000, 10, dev(labels), cnt)
    assert cnt.item() == int((scores.argmax(1) == labels).sum())


# ------------------------------------------------------------------------------------- A11 weight faults
@pytest.mark.parametrize("K_,N", [(64, 8), (784, 33), (100, 10), (4096, 16)])
def test_bernoulli_flips_match_oracle_mask(T, oracle, K_, N):
    from mlpcode import kernels as K
    rng = np.random.default_rng(K_ * N)
    W = rng.standard_normal((K_, N)).astype(np.float32)
    Wb = oracle.binarize(W)
    words = (K_ + 31) // 32
    Kp = (K_ + 15) // 16 * 16
    i8 = T.zeros((N, Kp), dtype=T.int8, device="cuda")
    bits = T.zeros((N, words), dtype=T.int32, device="cuda")
    K.binarize_weights_t(dev(W), wt_i8=i8, wt_bits=bits)
    nt, p, seed, stream = 3, 0.1, 0xC0FFEE1234, 7
    o_i8 = T.zeros((nt, N, Kp), dtype=T.int8, device="cuda")
    o_f16 = T.zeros((nt, N, Kp), dtype=T.float16, device="cuda")
    o_bits = T.zeros((nt, N, words), dtype=T.int32, device="cuda")
    m_bits = T.zeros((nt, N, words), dtype=T.int32, device="cuda")
    K.flip_weights_bernoulli(N=N, K=K_, p=p, seed=seed, stream_id=stream, trial0=5, n_trials=nt,
                             wt_bits=bits, out_i8=o_i8, out_f16=o_f16, out_bits=o_bits,
                             mask_bits=m_bits)
    for t in range(nt):
        mask = oracle.c_bernoulli_mask(K_, N, p, seed, stream, 5 + t)       # [K, N] bool
        want = oracle.flip_bnn_mask(Wb, mask)                               # callbacks.py:79-83
        assert np.array_equal(o_i8[t].cpu().numpy()[:, :K_], want.T.astype(np.int8))
        assert np.array_equal(o_f16[t].cpu().numpy()[:, :K_].astype(np.float32), want.T)
        assert np.array_equal(o_bits[t].cpu().numpy().view(np.uint32), oracle.c_pack_weights_t(want))
        assert np.array_equal(m_bits[t].cpu().numpy().view(np.uint32),
                              oracle.pack_sign_rows(np.where(mask.T, 1.0, -1.0).astype(np.float32)))
    # same result when the clean weights are supplied as int8 instead of bits
    o2 = T.zeros((1, N, Kp), dtype=T.int8, device="cuda")
    K.flip_weights_bernoulli(N=N, K=K_, p=p, seed=seed, stream_id=stream, trial0=5, n_trials=1,
                             wt_i8=i8, out_i8=o2)
    assert T.equal(o2[0], o_i8[0])
    # p = 0 leaves the weights alone, p = 1 negates all of them
    for pp, sign in ((0.0, 1), (1.0, -1)):
        K.flip_weights_bernoulli(N=N, K=K_, p=pp, seed=1, stream_id=0, trial0=0, n_trials=1,
                                 wt_bits=bits, out_i8=o2)
        assert np.array_equal(o2[0].cpu().numpy()[:, :K_], sign * Wb.T.astype(np.int8))


def test_exactk_flips_match_oracle_positions(T, oracle):
    from mlpcode import kernels as K
    dims = [(40, 24), (24, 10)]                       # (K, N) per layer
    total = sum(k * n for k, n in dims)
    ks = np.array([0, 1, 17, 300], np.int32)
    seed, trial0 = 99, 4
    offset = 0
    flipped_total = np.zeros(len(ks), int)
    for K_, N in dims:
        Kp, words = (K_ + 15) // 16 * 16, (K_ + 31) // 32
        rng = np.random.default_rng(K_)
        base = pm1(rng, (N, K_))
        io = T.zeros((len(ks), N, Kp), dtype=T.int8, device="cuda")
        io[:, :, :K_] = dev(base)
        iob = T.zeros((len(ks), N, words), dtype=T.int32, device="cuda")
        iob[:] = dev(oracle.pack_sign_rows(base.astype(np.float32)).view(np.int32))
        K.flip_weights_exactk(N=N, K=K_, total=total, offset=offset, seed=seed, trial0=trial0,
                              k_per_trial=dev(ks), k_max=int(ks.max()), io_i8=io, io_bits=iob)
        for t, k in enumerate(ks):
            pos = oracle.c_exactk_positions(total, seed, trial0 + t, int(k)).astype(np.int64)
            mine = pos[(pos >= offset) & (pos < offset + K_ * N)] - offset
            want = base.copy().reshape(-1)
            want[mine] = -want[mine]
            want = want.reshape(N, K_)
            assert np.array_equal(io[t].cpu().numpy()[:, :K_], want)
            assert np.array_equal(iob[t].cpu().numpy().view(np.uint32),
                                  oracle.pack_sign_rows(want.astype(np.float32)))
            flipped_total[t] += len(mine)
        offset += K_ * N
    assert np.array_equal(flipped_total, ks)          # exactly k_t flips over the whole network


def test_external_mask_flips(T, golden, oracle):
    from mlpcode import kernels as K
    g = golden("weight_faults")
    Wb, mask = g["cb_Wb"], g["cb_mask"]                # the reference's own seeded draw
    K_, N = Wb.shape
    Kp, words = (K_ + 15) // 16 * 16, (K_ + 31) // 32
    i8 = T.zeros((N, Kp), dtype=T.int8, device="cuda")
    bits = T.zeros((N, words), dtype=T.int32, device="cuda")
    K.binarize_weights_t(dev(Wb), wt_i8=i8, wt_bits=bits)
    K.flip_weights_mask(dev(mask.astype(np.uint8)), K_, N, io_i8=i8, io_bits=bits)
    assert np.array_equal(i8.cpu().numpy()[:, :K_], g["cb_flipped"].T.astype(np.int8))
    assert np.array_equal(bits.cpu().numpy().view(np.uint32), oracle.c_pack_weights_t(g["cb_flipped"]))


# ------------------------------------------------------------------------------------- A12 image faults
@pytest.mark.parametrize("mode", [0, 1, 2])
def test_image_flips_match_oracle(T, golden, oracle, mode):
    from mlpcode import kernels as K
    g = golden("image_faults")
    imgs, k = g["images"], int(g["k"])
    rng = np.random.default_rng(mode)
    big = rng.integers(0, 256, (40, 784), dtype=np.uint8)                  # MNIST-shaped rows
    for images, kk in ((imgs, k), (big, round(0.14 * 784 * 8)), (big[:3, :5], 1000)):
        R, F = images.shape
        nt = 3
        out = T.zeros((nt, R, F), dtype=T.uint8, device="cuda")
        K.flip_images(dev(images), kk, mode, 0xABCDEF, 2, nt, out)
        for t in range(nt):
            want = oracle.c_flip_image_rows(images, kk, mode, 0xABCDEF, 2 + t)
            assert np.array_equal(out[t].cpu().numpy(), want)


def test_normalize_u8(T, golden, oracle):
    from mlpcode import kernels as K
    g = golden("image_faults")
    imgs = g["images"]
    d = dev(imgs)
    mm = T.zeros(2, dtype=T.int32, device="cuda")
    mmf = T.zeros(2, dtype=T.float32, device="cuda")
    K.minmax_u8(d, mm, mmf)
    assert mm.tolist() == [int(imgs.min()), int(imgs.max())]
    out = T.zeros(imgs.shape, dtype=T.float32, device="cuda")
    K.normalize_u8(d, mmf[0:1], mmf[1:2], -1.0, 1.0, out)
    assert np.array_equal(out.cpu().numpy(), g["normalized"])              # utils.py:178-187, exact
