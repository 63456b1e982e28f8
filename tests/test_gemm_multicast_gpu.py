"""Cluster-multicast form of the tcgen05 GEMM (BNN_GEMM_MC=1: two CTAs on vertically adjacent tiles share the B tile
through TMA multicast, single-CTA MMAs) against the plain single-CTA form.  It performs the same UMMAs on the same
operands in the same K order, so every epilogue's output equals the single-CTA form's bit for bit (fp64 column-sum
accumulators: to rounding of the atomics' order).  Opt-in (BNN_TEST_MC=1) like the pair-form tests: a protocol bug
traps the kernel and poisons the CUDA context of the whole pytest process."""
import os

import numpy as np
import pytest

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("BNN_TEST_MC", "0") != "1",
                                 reason="cluster-multicast GEMM is opt-in: set BNN_TEST_MC=1")]

SHAPES = [(256, 256, 128), (300, 512, 784), (1000, 256, 256), (129, 300, 100), (8192, 4096, 4096)]


@pytest.fixture(scope="module")
def T():
    import torch
    from mlpcode import _native
    _native.load()
    return torch


def dev(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def padded(a, mult):
    import torch
    r, c = a.shape
    cp = (c + mult - 1) // mult * mult
    t = torch.zeros((r, cp), dtype=dev(a[:1, :1]).dtype, device="cuda")
    t[:, :c] = dev(a)
    return t


def both_forms(monkeypatch, run):
    outs = []
    for flag in ("0", "1"):
        monkeypatch.setenv("BNN_GEMM_MC", flag)
        outs.append(run())
    return outs


@pytest.mark.parametrize("M,N,K_", SHAPES)
def test_multicast_int8_products_equal_single_cta(T, monkeypatch, M, N, K_):
    from mlpcode import _native as nat, kernels as K
    rng = np.random.default_rng(M + N + K_)
    a = np.where(rng.random((M, K_)) < 0.5, -1, 1).astype(np.int8)
    b = np.where(rng.random((N, K_)) < 0.5, -1, 1).astype(np.int8)
    A, Bm = padded(a, 16), padded(b, 16)
    ldd = (N + 7) // 8 * 8

    def run():
        D = T.full((M, ldd), 77, dtype=T.int32, device="cuda")
        K.gemm(M=M, N=N, K=K_, A=(A,), B=(Bm,), lda=A.stride(0), ldb=Bm.stride(0), D=D, ldd=ldd,
               in_dtype=nat.DT_I8, epilogue=nat.EPI_STORE_S32, impl=nat.GEMM_TCGEN05)
        T.cuda.synchronize()
        return D.cpu().numpy()
    single, pair = both_forms(monkeypatch, run)
    assert np.array_equal(single, pair)
    if M * N * K_ < 3e9:
        assert np.array_equal(pair[:, :N], a.astype(np.int32) @ b.astype(np.int32).T)


@pytest.mark.parametrize("M,N,K_", [(256, 256, 128), (300, 512, 784), (4096, 4096, 784)])
def test_multicast_fp16_split_and_fused_stats_equal_single_cta(T, monkeypatch, M, N, K_):
    from mlpcode import _native as nat, kernels as K
    rng = np.random.default_rng(M * 3 + N + K_)
    x = (rng.standard_normal((M, K_))).astype(np.float32)
    w = np.where(rng.random((N, K_)) < 0.5, -1.0, 1.0).astype(np.float16)
    Xd = dev(x)
    amax, scale = T.zeros(1, device="cuda"), T.zeros(1, device="cuda")
    Kp = (K_ + 7) // 8 * 8
    hi, lo = T.zeros((M, Kp), dtype=T.float16, device="cuda"), T.zeros((M, Kp), dtype=T.float16, device="cuda")
    K.absmax_f32(Xd, amax)
    K.split_f16(Xd, amax, hi=hi, lo=lo, scale_out=scale)
    Wd = padded(w, 8)

    def run():
        D = T.zeros((M, N), dtype=T.float32, device="cuda")
        sums = T.zeros((2, N), dtype=T.float64, device="cuda")
        K.gemm(M=M, N=N, K=K_, A=(hi, lo), B=(Wd,), lda=Kp, ldb=Wd.stride(0), D=D, ldd=N,
               in_dtype=nat.DT_F16, epilogue=nat.EPI_STORE_F32, passes=((0, 0), (1, 0)), sa=scale,
               stats=dict(mode=nat.STATS_COLSUM, sums=sums), impl=nat.GEMM_TCGEN05)
        T.cuda.synchronize()
        return D.cpu().numpy(), sums.cpu().numpy()
    (d0, s0), (d1, s1) = both_forms(monkeypatch, run)
    assert np.array_equal(d0, d1)
    assert np.allclose(s0, s1, rtol=1e-12, atol=1e-9)


def test_multicast_training_step_equals_single_cta(monkeypatch):
    """Two training steps with every 256-wide GEMM in pair form: the same state as the single-CTA form."""
    import bnn_oracle as O
    from test_layers_gpu import build_net, state_of
    units, B = [784, 512, 512, 10], 1024
    rng = np.random.default_rng(5)
    params = O.new_params(units, rng)
    X = rng.random((B, units[0]), dtype=np.float32) * 2 - 1
    Y = np.eye(10, dtype=np.int8)[rng.integers(0, 10, B)]
    states = []
    for flag in ("0", "1"):
        monkeypatch.setenv("BNN_GEMM_MC", flag)
        nn = build_net(units, O.copy_params(params))
        for _ in range(2):
            nn.train_step(X, Y, 0.01)
        states.append(state_of(nn))
    for w0, w1 in zip(states[0]["W"], states[1]["W"]):
        assert np.array_equal(w0, w1)
    for b0, b1 in zip(states[0]["bn"], states[1]["bn"]):
        for k in b0:
            assert np.allclose(b0[k], b1[k], rtol=1e-6, atol=1e-7), k


@pytest.mark.parametrize("M,N,K_", [(256, 224, 256), (1000, 4096, 1024), (257, 500, 512)])
def test_multicast_fp4_products_equal_single_cta(T, monkeypatch, M, N, K_):
    from mlpcode import _native as nat, kernels as K
    rng = np.random.default_rng(M + N + K_)
    def f4(x):
        nib = np.where(x > 0, 0x2, 0xA).astype(np.uint8)
        return nib[:, 0::2] | (nib[:, 1::2] << 4)
    a = np.where(rng.random((M, K_)) < 0.5, -1, 1).astype(np.int8)
    b = np.where(rng.random((N, K_)) < 0.5, -1, 1).astype(np.int8)
    A4, B4 = dev(f4(a)), dev(f4(b))
    ldd = (N + 7) // 8 * 8

    def run():
        D = T.full((M, ldd), 77, dtype=T.int32, device="cuda")
        K.gemm(M=M, N=N, K=K_, A=(A4,), B=(B4,), lda=K_, ldb=K_, D=D, ldd=ldd, in_dtype=nat.DT_F4,
               epilogue=nat.EPI_STORE_S32)
        T.cuda.synchronize()
        return D.cpu().numpy()
    single, multi = both_forms(monkeypatch, run)
    assert np.array_equal(single, multi)
    assert np.array_equal(multi[:, :N], a.astype(np.int32) @ b.astype(np.int32).T)
