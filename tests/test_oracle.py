"""CPU: the oracle against the fixtures generated from the reference, and the C restatement against
NumPy / published known-answer vectors.  These pin the checker the GPU tests rely on."""
import numpy as np
import pytest

from conftest import params_from_golden


def test_elementwise_against_reference(golden, oracle):
    g = golden("elementwise")
    assert np.array_equal(oracle.binarize(g["x"]), g["binarized"])
    assert np.array_equal(oracle.unitstep(g["x"]), g["unitstep"])
    # sign(0) = sign(-0) = +1
    assert g["binarized"][0, 0] == 1 and g["binarized"][0, 1] == 1 and g["binarized"][0, 3] == -1
    # STE is numerically the identity on post-sign activations
    assert np.array_equal(g["ste"], g["dA"])
    assert np.array_equal(oracle.softmax(g["logits"]), g["softmax"])
    sm = g["softmax"].copy()
    sm[0, 0] = 0.0
    assert np.array_equal(oracle.cross_entropy_loss(sm, g["onehot"]), g["loss"])
    assert np.array_equal(sm, g["clipped"])
    assert np.array_equal(oracle.cross_entropy_derivative_softmax(sm.copy(), g["onehot"]), g["grad"])


def test_batchnorm_against_reference(golden, oracle):
    g = golden("batchnorm")
    out, cache = oracle.bn_forward_train(g["Z"].copy(), g["gamma"].copy(), g["beta"].copy())
    assert np.array_equal(out, g["out_train"])
    assert np.array_equal(cache["mu"], g["batch_mu"]) and np.array_equal(cache["var"], g["batch_var"])
    p = dict(gamma=g["gamma"].copy(), beta=g["beta"].copy(), mu=g["mu_run"].copy(),
             sigma=g["sigma_run"].copy())
    nd, dg, db = oracle.bn_backward(g["delta"].copy(), cache, p, float(g["lr"]))
    assert np.array_equal(nd, g["new_delta"])
    for k in ("gamma", "beta", "mu", "sigma"):
        assert np.array_equal(p[k], g[f"{k}_after"]), k
    ev = oracle.bn_forward_eval(g["Z"].copy(), g["gamma"], g["beta"], g["mu_run"], g["sigma_run"])
    assert np.array_equal(ev, g["out_eval"])


@pytest.mark.parametrize("name", ["train_tiny", "train_small", "train_ragged"])
def test_train_steps_against_reference(golden, oracle, name):
    g = golden(name)
    units = [int(u) for u in g["units"]]
    n = len(units) - 1
    p = params_from_golden(g, "init_", n)
    for s in range(g["X"].shape[0]):
        cost = oracle.train_step(p, g["X"][s].copy(), g["Y"][s], float(g["lr"]))
        assert cost == g["costs"][s]
        for i in range(n):
            assert np.array_equal(p["W"][i], g[f"after{s}_W{i}"])
            for k in ("gamma", "beta", "mu", "sigma"):
                assert np.array_equal(p["bn"][i][k], g[f"after{s}_{k}{i}"]), (s, i, k)
    assert np.array_equal(oracle.predict(p, g["X_eval"].copy()), g["probs_eval"])
    assert oracle.accuracy(p, g["X_eval"].copy(), g["labels_eval"]) == float(g["acc_eval"])


def test_weight_faults_against_reference(golden, oracle):
    g = golden("weight_faults")
    assert np.array_equal(oracle.flip_bnn_mask(g["cb_Wb"], g["cb_mask"]), g["cb_flipped"])
    units = [int(u) for u in g["units"]]
    p = params_from_golden(g, "", len(units) - 1)
    hooks = [(lambda m: (lambda W: oracle.flip_bnn_mask(W, m)))(g[f"mask{i}"])
             for i in range(len(units) - 1)]
    assert np.array_equal(oracle.predict(p, g["X"].copy(), weight_hooks=hooks), g["probs_fault"])
    assert np.array_equal(oracle.predict(p, g["X"].copy()), g["probs_clean"])


def test_image_faults_against_reference(golden, oracle):
    g = golden("image_faults")
    imgs, k = g["images"], int(g["k"])
    for mode in (0, 1, 2):
        out = g[f"out_mode{mode}"]
        diff = np.unpackbits(imgs ^ out, axis=1)
        before = np.unpackbits(imgs, axis=1)
        for r in range(imgs.shape[0]):
            cand = before.shape[1] if mode == 2 else int((before[r] == mode).sum())
            assert diff[r].sum() == min(k, cand)          # exactly k distinct flips
            if mode != 2:                                  # only candidate bits were touched
                assert np.all(before[r][diff[r] == 1] == mode)
    assert np.array_equal(oracle.normalize(imgs, -1, 1), g["normalized"])


# ------------------------------------------------------------------------- C restatement
def test_c_pack_matches_numpy(oracle):
    rng = np.random.default_rng(0)
    for R, C in ((1, 1), (3, 31), (5, 32), (4, 33), (7, 100)):
        x = rng.standard_normal((R, C)).astype(np.float32)
        x[0, 0] = -0.0
        assert np.array_equal(oracle.c_pack_sign_rows(x), oracle.pack_sign_rows(x))
        assert np.array_equal(oracle.unpack_sign_rows(oracle.pack_sign_rows(x), C), oracle.binarize(x))
    W = rng.standard_normal((45, 7)).astype(np.float32)
    assert np.array_equal(oracle.c_pack_weights_t(W), oracle.pack_sign_rows(np.ascontiguousarray(W.T)))


def test_c_xnor_gemm_is_the_pm1_product(oracle):
    rng = np.random.default_rng(1)
    for M, N, K in ((3, 5, 1), (4, 4, 31), (9, 6, 64), (5, 11, 200)):
        a = np.sign(rng.standard_normal((M, K))).astype(np.float32)
        b = np.sign(rng.standard_normal((N, K))).astype(np.float32)
        a[a == 0] = 1
        b[b == 0] = 1
        ref = (a @ b.T).astype(np.int32)                 # X.dot(W) on +-1 operands, layers.py:220
        assert np.array_equal(oracle.c_gemm_pm1(a.astype(np.int8), b.astype(np.int8)), ref)
        got = oracle.c_gemm_xnor(oracle.pack_sign_rows(a), oracle.pack_sign_rows(b), K)
        assert np.array_equal(got, ref)


def test_philox_known_answers(oracle):
    """Random123 v1.09 kat_vectors, philox4x32-10."""
    assert oracle.c_philox((0, 0, 0, 0), (0, 0)) == (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)
    f = 0xFFFFFFFF
    assert oracle.c_philox((f, f, f, f), (f, f)) == (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)
    assert oracle.c_philox((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344),
                           (0xA4093822, 0x299F31D0)) == (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)


def test_bernoulli_mask_statistics_and_determinism(oracle):
    K, N = 96, 50
    m1 = oracle.c_bernoulli_mask(K, N, 0.2, seed=5, stream=1, trial=0)
    m2 = oracle.c_bernoulli_mask(K, N, 0.2, seed=5, stream=1, trial=0)
    m3 = oracle.c_bernoulli_mask(K, N, 0.2, seed=5, stream=1, trial=1)
    assert np.array_equal(m1, m2) and not np.array_equal(m1, m3)
    assert abs(m1.mean() - 0.2) < 4 * np.sqrt(0.2 * 0.8 / (K * N))
    assert not oracle.c_bernoulli_mask(K, N, 0.0, 5, 1, 0).any()
    assert oracle.c_bernoulli_mask(K, N, 1.0, 5, 1, 0).all()


def test_exactk_positions_are_distinct_and_cover(oracle):
    for total in (1, 2, 7, 64, 1000, 36_806_656):
        k = min(total, 1000)
        pos = oracle.c_exactk_positions(total, seed=9, trial=3, k=k)
        assert pos.max() < total and len(np.unique(pos)) == k
    # a bijection: asking for every index yields a permutation
    pos = oracle.c_exactk_positions(777, seed=1, trial=0, k=777)
    assert np.array_equal(np.sort(pos), np.arange(777, dtype=np.uint64))
    assert not np.array_equal(pos, oracle.c_exactk_positions(777, seed=1, trial=1, k=777))


def test_c_image_flips_follow_reference_semantics(golden, oracle):
    g = golden("image_faults")
    imgs, k = g["images"], int(g["k"])
    before = np.unpackbits(imgs, axis=1)
    for mode in (0, 1, 2):
        out = oracle.c_flip_image_rows(imgs, k, mode, seed=3, trial=0)
        diff = np.unpackbits(imgs ^ out, axis=1)
        for r in range(imgs.shape[0]):
            cand = before.shape[1] if mode == 2 else int((before[r] == mode).sum())
            assert diff[r].sum() == min(k, cand)
            if mode != 2:
                assert np.all(before[r][diff[r] == 1] == mode)
        assert not np.array_equal(out, oracle.c_flip_image_rows(imgs, k, mode, seed=3, trial=1))
    # uniformity of the chosen positions over many trials (mode 2)
    row = np.zeros((1, 16), np.uint8)
    hits = np.zeros(128)
    for t in range(400):
        hits += np.unpackbits(oracle.c_flip_image_rows(row, 16, 2, seed=11, trial=t), axis=1)[0]
    assert hits.sum() == 400 * 16
    assert hits.min() > 20 and hits.max() < 85        # mean 50, sd ~6.6


@pytest.mark.parametrize("bias", [0, 128])
@pytest.mark.parametrize("lo_hi", [(0, 255), (3, 250), (17, 18)])
def test_u8_input_threshold_algebra_reproduces_reference_layer0(oracle, bias, lo_hi):
    """normalize + X.dot(sign(W)) + eval BN + sign on a uint8 batch (utils.py:178-187, layers.py:220,
    93-97) == integer compare of acc = U . Wb against per-unit thresholds; the only units allowed to
    differ are those whose BN output is within fp32 rounding of 0."""
    rng = np.random.default_rng(bias + lo_hi[0])
    R, F, N = 96, 112, 40
    lo, hi = lo_hi
    U = rng.integers(lo, hi + 1, (R, F), dtype=np.uint8)
    U[0, 0], U[1, 1] = lo, hi                       # the set's min / max are exactly lo / hi
    W = rng.standard_normal((F, N)).astype(np.float32)
    bnp = dict(gamma=(rng.random(N) + 0.5).astype(np.float32), beta=(rng.standard_normal(N) * 0.3).astype(np.float32),
               mu=(rng.standard_normal(N) * 2).astype(np.float32), sigma=(rng.random(N) * 5 + 0.5).astype(np.float32))
    bnp["gamma"][::5] *= -1
    want, zt = oracle.u8_layer0_reference(U, W, bnp)
    tau, sgn = oracle.u8_input_thresholds(W, bnp, lo, hi, bias=bias)
    Wb = oracle.binarize(W).astype(np.int64)
    g = (U.astype(np.int64) - bias) @ Wb              # what the int8 tensor pipe sums (u8, or u-128 as s8)
    got = np.where(g * sgn >= tau, 1.0, -1.0)
    diff = got != want
    assert np.all(np.abs(zt[diff]) < 1e-4), np.abs(zt[diff]).max()
    assert diff.mean() < 0.01


def test_float_bit_flips_against_reference(golden, oracle):
    """ErrorCallback(bnn=False) (callbacks.py:35-59, 85-123): the fixture holds the reference's outputs
    and element draws; the per-element bit draws are re-derived from the output itself."""
    g = golden("float_faults")
    W = g["W"]
    for mode in (0, 1, 2):
        for p in (0.1, 0.5):
            tag = f"m{mode}_p{int(p * 10)}"
            out, elem = g[f"out_{tag}"], g[f"elem_{tag}"]
            assert elem.size == round(W.size * p) and np.unique(elem).size == elem.size
            xor = W.view(np.uint32) ^ out.view(np.uint32)

            def choose(cand, k, _state={"i": 0}):
                e = elem[_state["i"]]
                _state["i"] += 1
                x = np.array([xor.flat[e]], dtype=np.uint32)
                flipped = np.flatnonzero(np.unpackbits(x.view(np.uint8)))
                assert flipped.size == k and np.all(np.isin(flipped, cand))
                return flipped
            mine = oracle.flip_float_bits(W, elem, p, mode, choose)
            assert np.array_equal(mine.view(np.uint32), out.view(np.uint32))
            untouched = np.ones(W.size, bool)
            untouched[elem] = False
            assert not xor.flat[untouched].any()


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_c_float_bit_flips_follow_reference_semantics(golden, oracle, mode):
    """The device's draw (restated in C): exactly round(size*p) distinct elements; in each, exactly
    round(cnt*p) candidate bits per _flipFunc (callbacks.py:35-59) are inverted, nothing else."""
    W = golden("float_faults")["W"]
    for p in (0.1, 0.5, 1.0):
        out = oracle.c_flip_f32_bits(W, p, mode, seed=0xF00D, trial=3)
        xor = (W.view(np.uint32) ^ out.view(np.uint32)).ravel()
        bits_all = np.unpackbits(W.ravel().view(np.uint8)).astype(bool).reshape(-1, 32)
        # which elements were selected cannot be seen when round(cnt*p) == 0; count those that must change
        changed = np.flatnonzero(xor)
        for e in changed:
            cand = oracle.float_flip_candidates(bits_all[e], mode)
            flipped = np.flatnonzero(np.unpackbits(np.array([xor[e]], np.uint32).view(np.uint8)))
            assert flipped.size == min(cand.size, round(cand.size * p)) and np.all(np.isin(flipped, cand))
        assert changed.size <= round(W.size * p)
        if p == 1.0 and mode == 2:
            assert changed.size == W.size
        again = oracle.c_flip_f32_bits(W, p, mode, seed=0xF00D, trial=3)
        assert np.array_equal(again.view(np.uint32), out.view(np.uint32))
        if p < 1.0:                                   # p = 1 flips every candidate of every element
            other = oracle.c_flip_f32_bits(W, p, mode, seed=0xF00D, trial=4)
            assert not np.array_equal(other.view(np.uint32), out.view(np.uint32))


def test_train_epochs_equals_reference_train(golden, oracle):
    """tests/golden/train_epochs.npz holds what the reference's own Network.train produced (history, final state);
    the oracle's epoch loop on the recorded permutations reproduces it exactly."""
    from conftest import params_from_golden
    g = golden("train_epochs")
    n = len(g["units"]) - 1
    p = params_from_golden(g, "init_", n)
    hist = oracle.train_epochs(p, g["X"].copy(), g["Y"].copy(), float(g["lr"]), len(g["loss"]),
                               int(g["batch_size"]), perms=list(g["perms"]))
    assert np.array_equal(hist["loss"], g["loss"]) and np.array_equal(hist["accuracy"], g["accuracy"])
    for i in range(n):
        assert np.array_equal(p["W"][i], g[f"final_W{i}"])
        for k in ("gamma", "beta", "mu", "sigma"):
            assert np.array_equal(p["bn"][i][k], g[f"final_{k}{i}"])
