"""bnn_oracle.py — TEST INFRASTRUCTURE ONLY.

CPU (NumPy) restatement of the reference algorithm for the binarized-layer hot path of
volf52/deep-neural-net, written as pure functions over an explicit parameter dict.  Only tests/,
`__graft_entry__.smoke()` and bench.py's `cpu_baseline` / `--impl reference` legs may import it, and
only as the checker / the timed CPU baseline — never as part of the product path
(deep-neural-net_b200/ never imports anything under oracle/).

Every function cites the reference lines it follows (paths relative to the reference root).  The
arithmetic deliberately uses the same NumPy calls in the same order as the reference (`ndarray.dot`,
`mean`/`var` over axis 0, boolean-mask assignment ...), so on the same NumPy build its results are
bit-identical to the reference's — which is how it is PINNED: the reference has no tests or golden
vectors of its own, so oracle/gen_golden.py imports the unmodified reference (through the
cupy->numpy shim in oracle/refshim) in the build container, runs both on seeded inputs, asserts exact
equality, and commits the reference's outputs as fixtures under tests/golden/.  tests/test_oracle.py
re-checks this oracle against those fixtures wherever the tests run.

The integer / RNG restatements live in bnn_oracle.c (ctypes wrappers at the bottom of this file).
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

EPSILON = 1e-4   # BatchNormLayer.EPSILON, mlpcode/layers.py:27
MA_BETA = 0.9    # BatchNormLayer.MA_BETA, mlpcode/layers.py:28


# ----------------------------------------------------------------------------------------------
# A1 / A6: sign()                                             layers.py:275-281, activation.py:106-113
# ----------------------------------------------------------------------------------------------
def binarize(x: np.ndarray) -> np.ndarray:
    out = x.copy()
    out[x >= 0] = 1.0
    out[x < 0] = -1.0
    return out


unitstep = binarize  # activation.py:106-113 is the same three lines on activations


def pack_sign_rows(x: np.ndarray) -> np.ndarray:
    """bits[r, w] bit j = (x[r, 32w+j] >= 0), zero padded — the packing the CUDA kernels use."""
    b = np.packbits(x >= 0, axis=1, bitorder="little")
    pad = (-b.shape[1]) % 4
    if pad:
        b = np.concatenate([b, np.zeros((b.shape[0], pad), np.uint8)], axis=1)
    return np.ascontiguousarray(b).view("<u4")


def unpack_sign_rows(bits: np.ndarray, n: int) -> np.ndarray:
    b = np.unpackbits(np.ascontiguousarray(bits).view(np.uint8), axis=1, bitorder="little")[:, :n]
    return np.where(b == 1, np.float32(1), np.float32(-1))


# ----------------------------------------------------------------------------------------------
# A4 / A5: batch-norm forward                                                  layers.py:74-99
# ----------------------------------------------------------------------------------------------
def bn_forward_train(Z, gamma, beta):
    mu = Z.mean(axis=0)
    var = Z.var(axis=0)
    znorm = (Z - mu) / np.sqrt(var + EPSILON)
    out = gamma * znorm + beta
    return out, dict(input=Z, znorm=znorm, mu=mu, var=var)


def bn_forward_eval(Z, gamma, beta, mu_run, sigma_run):
    out = Z - mu_run
    out /= np.sqrt(sigma_run + EPSILON)
    out *= gamma
    out += beta
    return out


# ----------------------------------------------------------------------------------------------
# A8: batch-norm backward + in-place SGD + running statistics                 layers.py:101-127
# ----------------------------------------------------------------------------------------------
def bn_backward(delta, cache, p, lr):
    """p: dict with gamma, beta, mu, sigma (updated in place / rebound like the reference)."""
    n = cache["input"].shape[0]
    x_mu = cache["input"] - cache["mu"]
    std_inv = 1.0 / np.sqrt(cache["var"] + EPSILON)
    dznorm = delta * p["gamma"]
    dvar = -0.5 * np.sum(dznorm * x_mu, axis=0) * std_inv ** 3
    dmu = np.sum(dznorm * -std_inv, axis=0) + dvar * np.mean(-2.0 * x_mu, axis=0)
    new_delta = (dznorm * std_inv) + (dvar * 2 * x_mu / n) + (dmu / n)
    dgamma = np.sum(delta * cache["znorm"], axis=0)
    dbeta = np.sum(delta, axis=0)
    p["gamma"] -= lr * dgamma
    p["beta"] -= lr * dbeta
    p["mu"] = (MA_BETA * p["mu"]) + ((1 - MA_BETA) * cache["mu"])
    p["sigma"] = (MA_BETA * p["sigma"]) + ((1 - MA_BETA) * cache["var"])
    return new_delta, dgamma, dbeta


# ----------------------------------------------------------------------------------------------
# A6 / A10: softmax, cross-entropy and its softmax-fused derivative
#                                                   activation.py:44-53, loss.py:10-36, loss.py:39-68
# ----------------------------------------------------------------------------------------------
def softmax(x):
    a = x - x.max(axis=1)[:, np.newaxis]
    np.exp(a, out=a)
    a /= a.sum(axis=1)[:, np.newaxis]
    return a


def cross_entropy_loss(ypred, y, eps=1e-15):
    np.clip(ypred, eps, 1 - eps, out=ypred)  # in place, like loss.py:32
    return -(y * np.log(ypred)).sum(axis=1)


def cross_entropy_derivative_softmax(ypred, y):
    dA = ypred - y
    dA /= ypred.shape[0]
    return dA


# ----------------------------------------------------------------------------------------------
# A2 / A3 / A9 / A13: the binarized layer and the per-batch loop
#                           layers.py:209-238, 247-271, 283-292; network.py:372-392, 394-399
# ----------------------------------------------------------------------------------------------
def new_params(units, rng):
    """Synthetic parameters with the reference's init distribution (layers.py:191-194) from a
    seeded generator (the reference's own init is unseeded, so parity tests always inject)."""
    W = [(rng.standard_normal((k, n), dtype=np.float32) * np.float32(np.sqrt(1.0 / k)))
         for k, n in zip(units[:-1], units[1:])]
    bn = [dict(gamma=np.ones(n, np.float32), beta=np.zeros(n, np.float32),
               mu=np.zeros(n, np.float32), sigma=np.ones(n, np.float32)) for n in units[1:]]
    return dict(W=W, bn=bn)


def copy_params(p):
    return dict(W=[w.copy() for w in p["W"]], bn=[{k: v.copy() for k, v in b.items()} for b in p["bn"]])


def layer_forward(X, W, bnp, is_last, train, out_af="softmax", weight_hook=None):
    Wb = binarize(W)                                   # layers.py:286
    if not train and weight_hook is not None:          # layers.py:215-218
        Wb = weight_hook(Wb)
    z = X.dot(Wb)                                      # layers.py:220
    if train:
        zt, cache = bn_forward_train(z, bnp["gamma"], bnp["beta"])
    else:
        zt, cache = bn_forward_eval(z, bnp["gamma"], bnp["beta"], bnp["mu"], bnp["sigma"]), None
    if not is_last:
        a = unitstep(zt)                               # activation.py:106-113
    elif out_af == "softmax":
        a = softmax(zt)
    else:
        a = zt                                         # identity, activation.py:8-9
    return a, dict(input=X, z_pre=z, z=zt, bn=cache)


def layer_backward(dA, cache, W, bnp, lr):
    """STE: hard_tanh_derivative on post-sign activations is the identity (activation.py:122-131,
    layers.py:252-254), so delta = dA.  Returns dX and updates W, bnp in place."""
    delta, dgamma, dbeta = bn_backward(dA, cache["bn"], bnp, lr)   # layers.py:257-258
    dw = cache["input"].T.dot(delta)                               # layers.py:260
    dx = delta.dot(W.T)                                            # layers.py:266 (latent W)
    W -= lr * dw                                                   # layers.py:268
    return dx, dw, delta


def predict(p, X, train=False, out_af="softmax", weight_hooks=None, keep=False):
    a, caches = X, []
    n = len(p["W"])
    for i in range(n):
        hook = weight_hooks[i] if weight_hooks is not None else None
        a, c = layer_forward(a, p["W"][i], p["bn"][i], i == n - 1, train, out_af, hook)
        caches.append(c)
    return (a, caches) if keep else a


def train_step(p, X, y, lr, record=None):
    """One `Network.__updateBatch` (network.py:372-392) with softmax + cross-entropy."""
    out, caches = predict(p, X, train=True, keep=True)
    cost = float(cross_entropy_loss(out, y).mean())
    delta = cross_entropy_derivative_softmax(out, y)
    if record is not None:
        record["probs"] = out.copy()
        record["dLdA"] = delta.copy()
        record["caches"] = caches
        record["dW"], record["dX"], record["delta_bn"] = [], [], []
    for i in reversed(range(len(p["W"]))):
        delta_in = delta
        delta, dw, dbn = layer_backward(delta_in, caches[i], p["W"][i], p["bn"][i], lr)
        if record is not None:
            record["dW"].insert(0, dw)
            record["dX"].insert(0, delta)
            record["delta_bn"].insert(0, dbn)
    return cost


def train_epochs(p, X, Y, lr, epochs, batch_size, perms=None, on_step=None):
    """The epoch loop of Network.train (network.py:305-337) with a constant learning rate: per epoch an optional
    cumulative shuffle `X = X[perm, :]` (:312-315; perms[e] is the index vector the reference would draw),
    mini-batches X[k:k+bs] (:480), epoch loss = mean of the batch losses (:325), accuracy of the WHOLE training set
    in eval mode (:326-328).  on_step(record) receives every step's intermediates.  Returns the history dict."""
    n = len(X)
    bs = n if batch_size == -1 else batch_size
    hist = dict(loss=[], accuracy=[])
    for e in range(epochs):
        if perms is not None:
            X, Y = X[perms[e], :], Y[perms[e], :]
        costs = []
        for k in range(0, n, bs):
            rec = {} if on_step is not None else None
            costs.append(train_step(p, X[k:k + bs].copy(), Y[k:k + bs], lr, record=rec))
            if on_step is not None:
                on_step(rec)
        hist["loss"].append(sum(costs) / len(costs))
        hist["accuracy"].append(accuracy(p, X.copy(), Y.argmax(axis=1)))
    return hist


def accuracy(p, X, labels, out_af="softmax", weight_hooks=None):
    """network.py:401-426 with batch_size=-1 (one batch = the whole set)."""
    preds = predict(p, X, False, out_af, weight_hooks).argmax(axis=1)
    return int((labels == preds).sum()) * 100.0 / X.shape[0]


# ----------------------------------------------------------------------------------------------
# A11 / A12: fault injection with an explicit mask / index choice
#                                                            callbacks.py:79-83, 130-147
# ----------------------------------------------------------------------------------------------
def flip_bnn_mask(W, mask):
    out = W.copy()
    out[mask] = -out[mask]
    return out


def flip_image_row(row_u8, idx):
    unpacked = np.unpackbits(row_u8)
    unpacked[idx] = np.logical_not(unpacked[idx])
    return np.packbits(unpacked)


def normalize(x, new_min=0.0, new_max=1.0):
    """utils.normalize, mlpcode/utils.py:178-187: global min / max; (x - min) stays in the input
    dtype, is scaled by the new range in fp32, divided by the (float64) old range, then shifted."""
    lo = x.min()
    span = 1.0 * (x.max() - lo)
    out = (x - lo).astype(np.float32) * (new_max - new_min)
    out /= span
    out += new_min
    return out


# ----------------------------------------------------------------------------------------------
# IEEE-754 bit flips of real-valued weights (the non-BNN branch)       callbacks.py:35-59, 85-123
# ----------------------------------------------------------------------------------------------
def float_flip_candidates(bits32, mode):
    """ErrorCallback._flipFunc, callbacks.py:35-52: bits32 = np.unpackbits of the value's 4 bytes in
    memory order (little endian: index 24 = sign, 25 = exponent MSB, 26..31 + 16 = the other exponent
    bits).  With the exponent MSB set, the other exponent bits are protected, otherwise the MSB itself;
    mode 0 keeps only 0-bits, mode 1 only 1-bits.  Returns the candidate indices in ascending order."""
    idx = np.arange(0, 32, dtype=np.int8)
    if bits32[25]:
        idx[26:32] = -1
        idx[16] = -1
    else:
        idx[25] = -1
    if mode == 0:
        idx[bits32[idx]] = -1
    elif mode == 1:
        idx[~bits32[idx]] = -1
    return idx[idx != -1]


def flip_float_bits(values, elem_idx, p, mode, choose):
    """ErrorCallback.__call__ non-BNN branch (callbacks.py:100-123) with explicit draws: `elem_idx` are
    the chosen elements (round(size*p) of them, callbacks.py:109); for each, `choose(cand, k)` picks
    k = min(len(cand), round(len(cand)*p)) candidate bit indices (callbacks.py:54-56), which are inverted."""
    shape = values.shape
    flat = np.array(values.flatten())
    picked = flat[elem_idx]
    unpacked = np.unpackbits(np.frombuffer(picked, dtype=np.uint8)).astype(bool).reshape(-1, 32)
    for row in unpacked:
        cand = float_flip_candidates(row, mode)
        k = min(cand.size, round(cand.size * p))
        sel = choose(cand, k)
        row[sel] = ~row[sel]
    flat[elem_idx] = np.frombuffer(np.packbits(unpacked), dtype=np.float32)
    return flat.reshape(shape)


def u8_layer0_reference(U, W, bnp):
    """What the reference computes for layer 0 of the image sweep (imageBitflipTrainModelSel.py:26-28):
    normalize(U, -1, 1) (utils.py:178-187) -> X.dot(sign(W)) (layers.py:286, 220) -> eval batch-norm
    (layers.py:93-97) -> unitstep (activation.py:106-113).  Returns (activations, BN output)."""
    x = normalize(U, new_min=-1, new_max=1)
    zt = bn_forward_eval(x.dot(binarize(W)), bnp["gamma"], bnp["beta"], bnp["mu"], bnp["sigma"])
    return unitstep(zt), zt


def u8_input_thresholds(W, bnp, lo, hi, new_min=-1.0, new_max=1.0, bias=0):
    """fp64 restatement of the algebra behind bnn_u8_input_thresholds: with acc_n = sum_k u_k*Wb[k,n],
    S_n = sum_k Wb[k,n] and r = (new_max-new_min)/(hi-lo), the reference's pre-activation is
    z_n = r*(acc_n - lo*S_n) + new_min*S_n, and sign(BN(z_n)) = +1 <=> (acc_n - bias*S_n)*sgn_n >= tau_n.
    Returns (ceil(tau), sgn).  The real-valued BN threshold is used (the device kernel takes the
    bit-exact fp32 one from bnn_bn_sign_thresholds; they differ by rounding only)."""
    Wb = binarize(W).astype(np.float64)
    S = Wb.sum(axis=0)
    g, be = bnp["gamma"].astype(np.float64), bnp["beta"].astype(np.float64)
    mu, d = bnp["mu"].astype(np.float64), np.sqrt(bnp["sigma"].astype(np.float64) + EPSILON)
    sgn = np.where(g < 0, -1.0, 1.0)
    with np.errstate(divide="ignore", invalid="ignore"):
        tz = sgn * (mu - be * d / g)                   # s*z >= tz  <=>  gamma*(z-mu)/d + beta >= 0
    r = (new_max - new_min) / (float(hi) - float(lo))
    tau = (tz - sgn * new_min * S) / r - sgn * (bias - float(lo)) * S
    return np.ceil(tau), sgn


# ----------------------------------------------------------------------------------------------
# ctypes front-end of bnn_oracle.c
# ----------------------------------------------------------------------------------------------
_HERE = Path(__file__).resolve().parent
C_LIB = _HERE / "_build" / "libbnn_oracle.so"
_clib = None


def build_c(force=False) -> Path:
    src = _HERE / "bnn_oracle.c"
    if force or not C_LIB.exists() or C_LIB.stat().st_mtime < src.stat().st_mtime:
        C_LIB.parent.mkdir(exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-std=c11", "-shared", "-fPIC", "-o", str(C_LIB), str(src), "-lm"])
    return C_LIB


def clib():
    global _clib
    if _clib is None:
        _clib = C.CDLL(str(build_c()))
    return _clib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def c_pack_sign_rows(x):
    x = np.ascontiguousarray(x, np.float32)
    R, Cn = x.shape
    words = (Cn + 31) // 32
    bits = np.zeros((R, words), np.uint32)
    clib().oracle_pack_sign_rows(_p(x), R, Cn, C.c_int64(Cn), _p(bits), C.c_int64(words))
    return bits


def c_pack_weights_t(W):
    W = np.ascontiguousarray(W, np.float32)
    K, N = W.shape
    words = (K + 31) // 32
    bits = np.zeros((N, words), np.uint32)
    clib().oracle_pack_weights_t(_p(W), K, N, _p(bits), C.c_int64(words))
    return bits


def c_gemm_xnor(a_bits, b_bits, K):
    a_bits, b_bits = np.ascontiguousarray(a_bits, np.uint32), np.ascontiguousarray(b_bits, np.uint32)
    M, N = a_bits.shape[0], b_bits.shape[0]
    D = np.zeros((M, N), np.int32)
    clib().oracle_gemm_xnor(_p(a_bits), C.c_int64(a_bits.shape[1]), _p(b_bits),
                            C.c_int64(b_bits.shape[1]), M, N, K, _p(D), C.c_int64(N))
    return D


def c_gemm_pm1(A, B):
    A, B = np.ascontiguousarray(A, np.int8), np.ascontiguousarray(B, np.int8)
    M, K = A.shape
    N = B.shape[0]
    D = np.zeros((M, N), np.int32)
    clib().oracle_gemm_pm1(_p(A), C.c_int64(K), _p(B), C.c_int64(K), M, N, K, _p(D), C.c_int64(N))
    return D


def c_philox(ctr, key):
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    clib().oracle_philox4x32_10(c, k, o)
    return tuple(o)


def bernoulli_threshold(p: float) -> int:
    return 0 if p <= 0 else min(1 << 32, int(p * 4294967296.0))


def c_bernoulli_mask(K, N, p, seed, stream, trial):
    mask = np.zeros((K, N), np.uint8)
    clib().oracle_bernoulli_mask(K, N, C.c_uint64(bernoulli_threshold(p)), C.c_uint64(seed),
                                 C.c_uint32(stream), C.c_uint32(trial), _p(mask))
    return mask.astype(bool)


def c_exactk_positions(total, seed, trial, k):
    pos = np.zeros(k, np.uint64)
    clib().oracle_exactk_positions(C.c_uint64(total), C.c_uint64(seed), C.c_uint32(trial), k, _p(pos))
    return pos


def c_flip_image_rows(images, k, mode, seed, trial):
    images = np.ascontiguousarray(images, np.uint8)
    R, F = images.shape
    out = np.zeros_like(images)
    clib().oracle_flip_image_rows(_p(images), R, F, k, mode, C.c_uint64(seed), C.c_uint32(trial), _p(out))
    return out


def c_flip_f32_bits(values, p, mode, seed, trial):
    """In-kernel draw of bnn_flip_f32_bits restated: elements P(0..n_sel-1), bits by per-element Feistel."""
    v = np.ascontiguousarray(values, dtype=np.float32).copy()
    n_sel = round(v.size * p)
    clib().oracle_flip_f32_bits(_p(v), C.c_uint64(v.size), C.c_uint64(n_sel), C.c_double(p),
                                C.c_int(mode), C.c_uint64(seed), C.c_uint32(trial))
    return v
