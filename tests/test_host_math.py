"""The pure per-element arithmetic of the fault kernels (deep-neural-net_b200/csrc/fault_math.cuh and the
RNG / keyed bijection of rng.cuh) is written as __host__ __device__ functions.  Here the SAME source is
compiled for the host with g++ and compared with the oracle — so the arithmetic the kernels execute is
checked on machines without a GPU; the `-m gpu` tests then check the kernels around it."""
import ctypes as C
import subprocess

import numpy as np
import pytest

from conftest import REPO

SHIM = r"""
#define __host__
#define __device__
#include "fault_math.cuh"
using namespace bnn;
extern "C" {
void host_flip_f32(uint32_t* w, uint64_t n, uint64_t n_sel, double p, int mode, uint64_t seed, uint32_t trial) {
    const int ebits = rng::feistel_bits(n);
    for (uint64_t i = 0; i < n_sel; ++i) {
        const uint64_t e = fault::f32_flip_element(i, n, ebits, seed, trial);
        w[e] ^= fault::f32_flip_mask(w[e], e, p, mode, seed, trial);
    }
}
float host_u8_threshold(float tz, double s, double S, double mn, double mx, double lo, double hi, double bias) {
    return fault::u8_input_threshold(tz, s, S, mn, mx, lo, hi, bias);
}
void host_exactk(uint64_t total, uint64_t seed, uint32_t trial, int k, uint64_t* out) {
    const rng::FeistelKey key = rng::feistel_key(seed, rng::kStreamExactK, trial, 0u);
    const int bits = rng::feistel_bits(total);
    for (int i = 0; i < k; ++i) out[i] = rng::feistel_perm((uint64_t)i, total, bits, key);
}
void host_philox(const uint32_t* c, const uint32_t* k, uint32_t* out) {
    const rng::U4 r = rng::philox4x32_10(c[0], c[1], c[2], c[3], k[0], k[1]);
    out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
}
}
"""


@pytest.fixture(scope="module")
def shim(tmp_path_factory):
    d = tmp_path_factory.mktemp("hostmath")
    (d / "shim.cpp").write_text(SHIM)
    so = d / "libhostmath.so"
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-Wno-unknown-pragmas", "-I",
                           str(REPO / "deep-neural-net_b200" / "csrc"), str(d / "shim.cpp"), "-o", str(so)])
    lib = C.CDLL(str(so))
    lib.host_u8_threshold.restype = C.c_float
    lib.host_u8_threshold.argtypes = [C.c_float] + [C.c_double] * 7
    return lib


def test_device_rng_source_matches_oracle_restatement(shim, oracle):
    """rng.cuh (device source, host-compiled) == bnn_oracle.c (independent restatement): Philox known
    answers and exact-k positions."""
    out = (C.c_uint32 * 4)()
    ctr = (C.c_uint32 * 4)(0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344)
    key = (C.c_uint32 * 2)(0xA4093822, 0x299F31D0)
    shim.host_philox(ctr, key, out)
    assert list(out) == [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]      # Random123 kat_vectors
    assert list(out) == list(oracle.c_philox(list(ctr), list(key)))
    for total, k in ((1000, 1000), (36806656, 500), (5, 5), (2 ** 33 + 7, 64), (6272, 6272), (100000, 777),
                     (2 ** 31 + 5, 100), (2 ** 32 + 1, 100), (3, 3), (2, 2), (1, 1)):
        pos = (C.c_uint64 * k)()
        shim.host_exactk(C.c_uint64(total), C.c_uint64(0xBEEF), 9, k, pos)
        want = oracle.c_exactk_positions(total, 0xBEEF, 9, k)
        assert np.array_equal(np.array(pos[:], dtype=np.uint64), np.asarray(want, dtype=np.uint64))
        got = np.array(pos[:], dtype=np.uint64)
        assert got.max() < total and np.unique(got).size == k          # distinct, in range
        if k == total:
            assert np.array_equal(np.sort(got), np.arange(total, dtype=np.uint64))   # a permutation of [0, n)


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_device_f32_flip_source_matches_oracle(shim, oracle, golden, mode):
    rng = np.random.default_rng(mode)
    big = (rng.standard_normal(5000) * np.exp(rng.standard_normal(5000) * 4)).astype(np.float32)
    pm = np.where(rng.random(4096) < 0.5, -1.0, 1.0).astype(np.float32)       # binarized weights
    for arr in (golden("float_faults")["W"].ravel(), big, pm):
        for p in (0.0, 0.1, 0.14, 0.5, 1.0):
            want = oracle.c_flip_f32_bits(arr, p, mode, seed=0xABCD1234, trial=2)
            got = arr.copy()
            shim.host_flip_f32(got.ctypes.data_as(C.c_void_p), C.c_uint64(got.size),
                               C.c_uint64(round(got.size * p)), C.c_double(p), mode, C.c_uint64(0xABCD1234), 2)
            assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), (arr.size, p)


@pytest.mark.parametrize("bias", [0, 128])
def test_device_u8_threshold_source_matches_oracle_algebra(shim, oracle, bias):
    rng = np.random.default_rng(bias)
    F, N = 784, 64
    W = rng.standard_normal((F, N)).astype(np.float32)
    bnp = dict(gamma=(rng.random(N) + 0.5).astype(np.float32), beta=(rng.standard_normal(N) * 0.3).astype(np.float32),
               mu=(rng.standard_normal(N) * 3).astype(np.float32), sigma=(rng.random(N) * 20 + 1).astype(np.float32))
    bnp["gamma"][::7] *= -1
    S = oracle.binarize(W).sum(axis=0).astype(np.float64)
    g, be = bnp["gamma"].astype(np.float64), bnp["beta"].astype(np.float64)
    d = np.sqrt(bnp["sigma"].astype(np.float64) + oracle.EPSILON)
    for lo, hi in ((0, 255), (7, 200), (30, 31)):
        tau, sgn = oracle.u8_input_thresholds(W, bnp, lo, hi, bias=bias)
        tz = (sgn * (bnp["mu"].astype(np.float64) - be * d / g)).astype(np.float32)
        got = np.array([shim.host_u8_threshold(float(tz[n]), float(sgn[n]), float(S[n]), lo, hi, -1.0, 1.0,
                                               float(bias)) for n in range(N)])
        assert np.max(np.abs(got - tau)) <= 1.0          # tz was rounded to fp32 on the way in
        assert np.mean(got == tau) > 0.9
    inf = float("inf")
    assert shim.host_u8_threshold(1.0, 1.0, 10.0, 5, 5, -1.0, 1.0, 0.0) == inf         # max == min
    assert shim.host_u8_threshold(inf, 1.0, 10.0, 0, 255, -1.0, 1.0, 0.0) == inf
    assert shim.host_u8_threshold(-inf, 1.0, 10.0, 0, 255, -1.0, 1.0, 0.0) == -inf
    assert shim.host_u8_threshold(float("nan"), 1.0, 10.0, 0, 255, -1.0, 1.0, 0.0) == inf


def test_device_u8_threshold_is_the_exact_decision_boundary(shim):
    """For random (range, column sum, BN threshold, orientation, operand bias): the integer compare
    s*g >= thr decides exactly like s*z >= tz evaluated in rational arithmetic, for every integer g around
    the boundary — the ceiling is on the right side and the fp64 evaluation does not lose the boundary."""
    from fractions import Fraction
    rng = np.random.default_rng(99)
    checked = 0
    for _ in range(400):
        lo = int(rng.integers(0, 200))
        hi = int(rng.integers(lo + 1, 256))
        K = int(rng.choice([16, 784, 3072, 4096]))
        S = int(rng.integers(-K // 4, K // 4 + 1))
        s = float(rng.choice([-1.0, 1.0]))
        bias = int(rng.choice([0, 128]))
        new_min, new_max = (-1.0, 1.0) if rng.random() < 0.7 else (0.0, 1.0)
        tz = float(np.float32(rng.standard_normal() * 20))
        thr = shim.host_u8_threshold(tz, s, float(S), float(lo), float(hi), new_min, new_max, float(bias))
        if not np.isfinite(thr):
            continue
        r = Fraction(new_max - new_min) / (hi - lo)
        for g in range(int(thr) - 3, int(thr) + 4):
            for sg in (g, -g):
                z = r * (sg + (bias - lo) * S) + Fraction(new_min) * S
                want = Fraction(s) * z >= Fraction(tz)
                got = s * sg >= thr
                assert want == got, (lo, hi, K, S, s, bias, tz, thr, sg)
                checked += 1
    assert checked > 3000
