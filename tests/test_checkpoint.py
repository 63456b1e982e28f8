"""Checkpoint bridge: the `.npz` files with the reference's HDF5 key layout, and the packed-bit
export.  tests/golden/checkpoint_ref.npz was WRITTEN BY THE UNMODIFIED REFERENCE (Network.save_weights,
network.py:428-472, through oracle/refshim/h5py.py) after two training steps; checkpoint_ref_io.npz holds the
state it saved and the reference's eval-mode probabilities on a fixed input (oracle/gen_golden.py gen_checkpoint,
which also makes the reference's fromModel read this package's writer output)."""
import numpy as np
import pytest

from conftest import GOLDEN, rel_err

REF = GOLDEN / "checkpoint_ref.npz"
LAYOUT = ({"units", "useBias", "useBatchNorm"}
          | {f"{g}/{g}_{i}" for g in ("weights", "gammas", "betas", "mus", "sigmas") for i in range(3)})


@pytest.fixture(scope="module")
def io():
    with np.load(GOLDEN / "checkpoint_ref_io.npz") as f:
        return {k: f[k] for k in f.files}


def test_reader_takes_the_reference_written_file(io):
    from mlpcode import checkpoint as ck
    with np.load(REF) as f:
        assert set(f.files) == LAYOUT                      # exactly the members network.py:438-470 creates
        assert f["units"].dtype == np.uint32 and f["useBias"].dtype == np.bool_
    got = ck.read_checkpoint(REF)
    assert got["units"] == [40, 24, 12, 10] and got["useBatchNorm"] and not got["useBias"] and not got["packed_only"]
    for i in range(3):
        assert np.array_equal(got["weights"][i], io[f"W{i}"]) and got["weights"][i].dtype == np.float32
        for g, k in zip(ck.BN_GROUPS, ("gamma", "beta", "mu", "sigma")):
            assert np.array_equal(got[g][i], io[f"{k}{i}"])


def test_writer_reproduces_the_reference_file(tmp_path):
    """read -> write gives the same members with the same dtypes, shapes and bytes as the reference's file."""
    from mlpcode import checkpoint as ck
    got = ck.read_checkpoint(REF)
    out = ck.write_checkpoint(tmp_path / "again.hdf5", got["units"], got["useBias"], got["useBatchNorm"],
                              got["weights"], None, [got[g] for g in ck.BN_GROUPS])
    assert out.name == "again.hdf5"                        # np.savez must not append ".npz"
    with np.load(REF) as a, np.load(out) as b:
        assert set(a.files) == set(b.files)
        for k in a.files:
            assert a[k].dtype == b[k].dtype and a[k].shape == b[k].shape and np.array_equal(a[k], b[k]), k


@pytest.mark.parametrize("K,N", [(40, 24), (33, 5), (64, 64), (1, 1), (784, 100)])
def test_packed_export_is_the_library_bit_layout(oracle, K, N):
    """packed_weights_i[n, w] bit j = (W[32 w + j, n] >= 0): the oracle's row packing of W^T (= bnn_binarize_
    weights_t `wt_bits`, checked against the kernel in test_kernels_gpu.py), sign(0) = sign(-0.0) = +1."""
    from mlpcode import checkpoint as ck
    rng = np.random.default_rng(K * N)
    W = rng.standard_normal((K, N)).astype(np.float32)
    W.flat[:: 7] = 0.0
    W.flat[3:: 11] = -0.0
    bits = ck.pack_weights_t(W)
    assert bits.dtype == np.dtype("<u4") and bits.shape == (N, (K + 31) // 32)
    assert np.array_equal(bits, oracle.pack_sign_rows(np.ascontiguousarray(W.T)))
    assert np.array_equal(ck.unpack_weights_t(bits, K), oracle.binarize(W))
    if K % 32:                                              # pad bits are zero
        assert not (bits[:, -1] >> (K % 32)).any()


def test_packed_only_file_round_trip(tmp_path, oracle):
    from mlpcode import checkpoint as ck
    got = ck.read_checkpoint(REF)
    bn = [got[g] for g in ck.BN_GROUPS]
    both = ck.read_checkpoint(ck.write_checkpoint(tmp_path / "both.npz", got["units"], False, True, got["weights"],
                                                  None, bn, packed=True))
    assert not both["packed_only"] and all(np.array_equal(a, b) for a, b in zip(both["weights"], got["weights"]))
    p = ck.write_checkpoint(tmp_path / "bits.npz", got["units"], False, True, got["weights"], None, bn,
                            packed=True, latent=False)
    only = ck.read_checkpoint(p)
    assert only["packed_only"]
    for w, v in zip(only["weights"], got["weights"]):
        assert np.array_equal(w, oracle.binarize(v))
    with np.load(p) as f:
        assert not any(k.startswith("weights/") for k in f.files)
        assert sum(f[k].nbytes for k in f.files if k.startswith("packed")) * 8 <= 32 * sum(
            n * ((k + 31) // 32) for k, n in zip(got["units"][:-1], got["units"][1:]))
    with pytest.raises(AssertionError):
        ck.write_checkpoint(tmp_path / "none.npz", got["units"], False, True, got["weights"], None, bn,
                            packed=False, latent=False)


@pytest.mark.gpu
def test_from_model_reproduces_the_reference_predictions(io, tmp_path):
    """Network.fromModel on the reference-written file: eval-mode probabilities equal the reference's; saving
    with binarized=True and loading the packed-only file gives bit-identical probabilities."""
    from mlpcode.activation import ActivationFuncs as af
    from mlpcode.network import Network
    nn = Network.fromModel(REF, useGpu=True, binarized=True)
    nn.compile(hiddenAf=af.sign, outAf=af.softmax)
    probs = nn.predict(io["X"]).get()
    assert rel_err(probs, io["probs"]) < 1e-5
    path = nn.save_weights("deploy", binarized=True, directory=tmp_path, latent=False)
    assert path.name == "deploy_binarized.npz"
    nn2 = Network.fromModel(path, useGpu=True, binarized=True)
    nn2.compile(hiddenAf=af.sign, outAf=af.softmax)
    assert np.array_equal(nn2.predict(io["X"]).get(), probs)
    full = nn.save_weights("full", directory=tmp_path)
    with np.load(full) as f:
        assert set(f.files) == LAYOUT
