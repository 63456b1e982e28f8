"""Host logic of the fault sweeps, exercised WITHOUT a GPU: device buffers are replaced by CPU tensors
and every C-ABI call goes to the real libbnn_b200, which validates its arguments (shapes, pitches,
alignment, enum combinations) and then fails at its first CUDA call with BNN_ERR_CUDA.  A
BNN_ERR_ARG / BNN_ERR_UNSUPPORTED from any call is a bug in the Python side of the boundary.  Nothing is
computed here; the numbers are checked by the `-m gpu` tests."""
import numpy as np
import pytest


@pytest.fixture
def dry(monkeypatch):
    import torch
    if torch.cuda.is_available():   # with a device the calls would really launch — on host pointers
        pytest.skip("argument-validation dry run is for machines without a GPU")
    from mlpcode import _native as nat, device, kernels, layers, network, sweep
    nat.load()
    cpu = torch.device("cpu")

    def to_tensor(a, dtype=None):
        t = a.t if isinstance(a, device.DeviceArray) else (
            a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a)))
        if dtype is not None and t.dtype != dtype:
            t = t.to(dtype)
        return t.contiguous()

    for mod in (device, kernels, layers, network, sweep):
        for name, fn in (("require_cuda", lambda: cpu), ("stream_ptr", lambda: 0), ("to_tensor", to_tensor)):
            if hasattr(mod, name):
                monkeypatch.setattr(mod, name, fn)
    calls = []

    def call(name, *args):
        lib = nat.load()
        rc = getattr(lib, name)(*args)
        calls.append((name, rc))
        if rc not in (nat.BNN_OK, nat.BNN_ERR_CUDA):
            raise nat.BnnError(name, rc, (lib.bnn_last_error() or b"").decode(errors="replace"))

    monkeypatch.setattr(nat, "call", call)
    monkeypatch.setenv("BNN_GEMM_IMPL", "tcgen05")
    return calls


def _net(units):
    import bnn_oracle as O
    from mlpcode.activation import ActivationFuncs as af
    from mlpcode.network import Network
    params = O.new_params(units, np.random.default_rng(0))
    nn = Network(units, useGpu=True, binarized=True, useBatchNorm=True)
    nn.compile(lr=0.01, hiddenAf=af.sign, outAf=af.identity)
    nn.load_weights(params["W"])
    return nn


@pytest.mark.parametrize("variant", ["u8", "s8"])
def test_batched_image_sweep_emits_wellformed_abi_calls(dry, monkeypatch, variant):
    from mlpcode.sweep import ImageFaultSweep
    monkeypatch.setenv("BNN_U8_GEMM", variant)
    units = [784, 64, 32, 10]
    nn = _net(units)
    imgs = np.random.default_rng(1).integers(0, 256, (96, 784), dtype=np.uint8)
    sw = ImageFaultSweep(nn, imgs, np.zeros(96, np.uint8), trials_per_launch=2)
    assert sw.batched
    del dry[:]
    acc = sw.run(0.14, 3, seed=5)                       # groups of 2 + 1 trials
    assert acc.shape == (3,)
    names = [n for n, _ in dry]
    per_group = ["bnn_flip_images_ex", "bnn_u8_input_thresholds", "bnn_gemm", "bnn_gemm", "bnn_gemm",
                 "bnn_bn_apply", "bnn_count_correct"]
    assert names == per_group * 2                       # no normalize, no fp16 split, no per-trial loop
    # shapes the batched path cannot take fall back to the per-trial device path (never to the CPU)
    assert not ImageFaultSweep.batch_ok(nn, 96, 780) and not ImageFaultSweep.batch_ok(nn, 16, 784)
    with pytest.raises(ValueError):
        ImageFaultSweep(_net([780, 64, 10]), imgs[:, :780], np.zeros(96, np.uint8), batched=True)


def test_weight_sweep_emits_wellformed_abi_calls(dry):
    from mlpcode.sweep import WeightFaultSweep
    units = [784, 256, 128, 10]
    nn = _net(units)
    X = np.random.default_rng(2).random((64, 784), dtype=np.float32) * 2 - 1
    sw = WeightFaultSweep(nn, X, np.zeros(64, np.uint8), trials_per_launch=4)
    del dry[:]
    assert sw.run_bernoulli(0.1, 6, seed=3).shape == (6,)
    assert sw.run_exactk(lambda t: 1 + t, 6, seed=3).shape == (6,)
    names = [n for n, _ in dry]
    # the +-1 x +-1 layers run on the FP4 pipe: their flips act on packed words, expanded by bnn_bits_to_f4
    assert sw.f4 == [False, True, True]
    # per layer and launch group (2 groups); exact-k replicates the clean fp16 copy of layer 0 with a p = 0 call
    assert names.count("bnn_flip_weights_bernoulli") == 2 * 3 + 2 * 1
    assert names.count("bnn_flip_weights_exactk") == 2 * 3
    assert names.count("bnn_bits_to_f4") == 2 * 2 + 2 * 2
    assert names.count("bnn_gemm") == 4 * 3
    assert all(rc in (0, -2) for _, rc in dry)


def test_weight_sweep_without_fp4_keeps_int8_copies(dry, monkeypatch):
    from mlpcode.sweep import WeightFaultSweep
    monkeypatch.setenv("BNN_FP4", "0")
    nn = _net([784, 256, 128, 10])
    X = np.random.default_rng(2).random((64, 784), dtype=np.float32) * 2 - 1
    sw = WeightFaultSweep(nn, X, np.zeros(64, np.uint8), trials_per_launch=4)
    del dry[:]
    assert sw.run_bernoulli(0.1, 6, seed=3).shape == (6,)
    assert sw.run_exactk(lambda t: 1 + t, 6, seed=3).shape == (6,)
    names = [n for n, _ in dry]
    assert sw.f4 == [False, False, False] and "bnn_bits_to_f4" not in names
    assert names.count("bnn_flip_weights_bernoulli") == 2 * 3 + 2 * 3
    assert names.count("bnn_gemm") == 4 * 3


def test_predict_routes_wide_hidden_layers_through_fp4(dry, monkeypatch):
    """Eval-mode forward: a layer whose consumer takes nibbles ends in BNN_EPI_SIGN_F4, the consumer's weights come
    out of bnn_binarize_weights_t as nibbles — or, with a fault hook registered, out of bnn_bits_to_f4 after the
    hook has acted on the packed words; every call passes the library's argument validation."""
    from mlpcode.callbacks import ErrorCallback
    nn = _net([784, 256, 128, 10])
    X = np.random.default_rng(3).random((40, 784), dtype=np.float32) * 2 - 1
    del dry[:]
    nn.predict(X)
    names = [n for n, _ in dry]
    assert "bnn_bits_to_f4" not in names and names.count("bnn_gemm") == 3
    nn._layers[1].callbacks = [ErrorCallback(0.1, bnn=True)]
    del dry[:]
    nn.predict(X)
    assert [n for n, _ in dry].count("bnn_bits_to_f4") == 1
    nn._layers[1].callbacks = []
    monkeypatch.setenv("BNN_FP4", "0")
    del dry[:]
    nn.predict(X)
    assert "bnn_bits_to_f4" not in [n for n, _ in dry]


def test_training_step_routes_the_forward_through_fp4(dry, monkeypatch):
    """Training forward: +-1 layers read nibbles written by the producer's bnn_bn_sign (no conversion kernels) and
    fuse their column sums into the FP4 product (no bnn_colstats for them); BNN_FP4_TRAIN=0 is the int8 form."""
    nn = _net([784, 256, 128, 10])
    X = np.random.default_rng(3).random((64, 784), dtype=np.float32) * 2 - 1
    y = np.eye(10, dtype=np.float32)[np.arange(64) % 10]
    monkeypatch.setenv("BNN_GRAPH", "0")
    monkeypatch.setenv("BNN_OVERLAP_DW", "0")      # no side stream on the CPU-only host
    del dry[:]
    nn.train_step(X, y, 0.01)
    names = [n for n, _ in dry]
    assert "bnn_bits_to_f4" not in names and "bnn_colstats" not in names
    monkeypatch.setenv("BNN_FP4_TRAIN", "0")
    del dry[:]
    nn.train_step(X, y, 0.01)
    assert [n for n, _ in dry].count("bnn_colstats") == 1      # the 10-wide int8 output layer


def test_unchanged_script_flow_stays_on_bytes(dry):
    """imageBitflipTrainModelSel.py:24-28 verbatim — icb(testX); normalize(newX, -1, 1); get_accuracy —
    reaches layer 0 as uint8: no normalize kernel, no fp16 split, one u8 x s8 GEMM."""
    from mlpcode.callbacks import ImageErrorCallback
    from mlpcode.utils import NormalizedU8, normalize
    nn = _net([784, 64, 32, 10])
    testX = np.random.default_rng(1).integers(0, 256, (96, 784), dtype=np.uint8)
    testY = np.zeros(96, np.uint8)
    icb = ImageErrorCallback(0.14)
    del dry[:]
    newX = icb(testX, gpu=True)
    newX = normalize(newX, newMin=-1, newMax=1)
    assert isinstance(newX, NormalizedU8) and newX.shape == (96, 784) and newX.dtype == np.float32
    nn.get_accuracy(newX, testY)
    names = [n for n, _ in dry]
    assert "bnn_normalize_u8" not in names and "bnn_split_f16" not in names
    assert names.count("bnn_u8_input_thresholds") == 1 and names.count("bnn_gemm") == 3
    # a batched evaluation slices rows and keeps the parent's min / max
    part = newX[10:50]
    assert isinstance(part, NormalizedU8) and part.shape == (40, 784) and part.mm_u32 is newX.mm_u32
    # generic consumers still get the fp32 values (materialised on demand)
    del dry[:]
    assert tuple(newX.t.shape) == (96, 784)
    assert [n for n, _ in dry] == ["bnn_normalize_u8"]


def test_weight_sweep_on_uint8_images_emits_wellformed_abi_calls(dry):
    """weightsErrorInjectionv2.py:15 normalizes the uint8 test set; the sweep keeps it as bytes."""
    from mlpcode.sweep import WeightFaultSweep
    from mlpcode.utils import normalize
    nn = _net([784, 256, 128, 10])
    imgs = np.random.default_rng(4).integers(0, 256, (64, 784), dtype=np.uint8)
    sw = WeightFaultSweep(nn, normalize(imgs, newMin=-1., newMax=1.), np.zeros(64, np.uint8), trials_per_launch=4)
    assert sw.u8 is not None
    del dry[:]
    assert sw.run_bernoulli(0.1, 6, seed=3).shape == (6,)
    assert sw.run_exactk(lambda t: 1 + t, 6, seed=3).shape == (6,)
    names = [n for n, _ in dry]
    assert "bnn_split_f16" not in names and "bnn_normalize_u8" not in names
    assert names.count("bnn_u8_input_thresholds") == 4          # once per launch group
    assert names.count("bnn_gemm") == 4 * 3
    # a real-valued evaluation set keeps the fp16-split layer 0
    X = np.random.default_rng(2).random((64, 784), dtype=np.float32) * 2 - 1
    assert WeightFaultSweep(nn, X, np.zeros(64, np.uint8)).u8 is None


def test_dataset_readers_parse_idx_and_npy(dry, tmp_path, monkeypatch):
    """mlpcode.utils.loadDataset (utils.py:191-661): IDX files for mnist / fashion-mnist, NPY for mnist-c;
    images stay uint8, training labels one-hot int8, test labels uint8."""
    import struct
    from mlpcode import utils
    rng = np.random.default_rng(0)
    d = tmp_path / "mnist"
    d.mkdir()
    for prefix, n in (("train", 50), ("t10k", 20)):
        imgs = rng.integers(0, 256, (n, 28, 28), dtype=np.uint8)
        labs = rng.integers(0, 10, n, dtype=np.uint8)
        (d / f"{prefix}-images-idx3-ubyte").write_bytes(struct.pack(">IIII", 2051, n, 28, 28) + imgs.tobytes())
        (d / f"{prefix}-labels-idx1-ubyte").write_bytes(struct.pack(">II", 2049, n) + labs.tobytes())
        if prefix == "t10k":
            want_test = (imgs.reshape(n, 784), labs)
    monkeypatch.setattr(utils, "DATADIR", tmp_path)
    trainX, trainY, testX, testY = utils.loadDataset(utils.DATASETS.mnist, useGpu=True)
    assert trainX.shape == (50, 784) and trainX.dtype == np.uint8
    assert trainY.shape == (50, 10) and trainY.dtype == np.int8 and int(trainY.t.sum()) == 50
    assert np.array_equal(testX.get(), want_test[0]) and np.array_equal(testY.get(), want_test[1])
    nx = utils.normalize(testX, newMin=-1, newMax=1)
    assert isinstance(nx, utils.NormalizedU8) and nx.mm_u32.shape == (2,)
    (d / "t10k-labels-idx1-ubyte").write_bytes(struct.pack(">II", 1234, 20) + bytes(20))
    with pytest.raises(ValueError, match="Magic number"):
        utils.loadDataset(utils.DATASETS.mnist, useGpu=True)
    with pytest.raises(FileNotFoundError):
        utils.loadDataset(utils.DATASETS.fashion, useGpu=True)
    c = tmp_path / "mnist-c" / "fog"
    c.mkdir(parents=True)
    for name, arr in (("train_images", rng.integers(0, 256, (30, 28, 28, 1), dtype=np.uint8)),
                      ("train_labels", rng.integers(0, 10, 30)), ("test_images", rng.integers(0, 256, (10, 28, 28, 1), dtype=np.uint8)),
                      ("test_labels", rng.integers(0, 10, 10))):
        np.save(c / f"{name}.npy", arr)
    tX, tY, eX, eY = utils.loadDataset(utils.DATASETS.mnistc_fog, useGpu=True, encoded=False)
    assert tX.shape == (30, 784) and tY.shape == (30,) and eX.shape == (10, 784) and eY.dtype == np.uint8
    assert str(utils.DATASETS.mnistc_fog) == "mnist_c-fog" and len(list(utils.DATASETS)) == 34
