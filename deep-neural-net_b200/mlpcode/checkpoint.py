"""Checkpoint files with the key layout of the reference's HDF5 models, stored as `.npz` (h5py is not in the image).

Reference: Network.save_weights (mlpcode/network.py:428-472) writes, and Network.fromModel (:157-228) reads,
    units            uint32 [L+1]
    useBias          bool scalar
    useBatchNorm     bool scalar
    weights/weights_i   float32 [K_i, N_i]            (latent weights, i = 0..L-1)
    biases/biases_i     float32 [N_i]                 (only when useBias)
    gammas/gammas_i, betas/betas_i, mus/mus_i, sigmas/sigmas_i   float32 [N_i]   (only when useBatchNorm)
An HDF5 dataset `group/name` becomes the npz member of that name, so the reference's own functions produce and
consume exactly these files when its `h5py` import is pointed at an npz-backed stand-in (the test infrastructure
has one — that is how tests/golden/checkpoint_ref.npz was written by the UNMODIFIED reference).

Packed-bit export, an addition of this package for deployment of a trained binarized network:
    packed_weights/packed_weights_i   uint32 [N_i, ceil(K_i / 32)]
bit j of word w of row n = (W[32 w + j, n] >= 0) — sign(0) = +1 (layers.py:278-279), pad bits 0 — the K-major
operand the library's kernels read (bnn_binarize_weights_t `wt_bits`) and the fault kernels mutate: 1 bit per
weight, 4.6 MB for the 784-4096-4096-4096-10 network instead of 147 MB.  A file may carry both forms; one that
carries only the packed form loads as latent weights of exactly +-1, which every eval-mode result depends on
through sign(W) alone (layers.py:283-292).

This module is plain NumPy on host arrays (no torch, no package-relative imports): the golden generator loads it
next to the reference, whose package is also called `mlpcode`."""
from __future__ import annotations

from pathlib import Path

import numpy as np

BN_GROUPS = ("gammas", "betas", "mus", "sigmas")


def pack_weights_t(W: np.ndarray) -> np.ndarray:
    """[K, N] real -> [N, ceil(K/32)] uint32 sign words, K-major (see module docstring)."""
    W = np.asarray(W)
    K, N = W.shape
    words = (K + 31) // 32
    bits = np.zeros((N, words * 32), np.uint8)
    bits[:, :K] = (W >= 0).T
    return np.packbits(bits, axis=1, bitorder="little").view("<u4").reshape(N, words)


def unpack_weights_t(bits: np.ndarray, K: int) -> np.ndarray:
    """Inverse of pack_weights_t up to the magnitude: [N, words] uint32 -> [K, N] float32 of +-1."""
    bits = np.ascontiguousarray(bits, dtype="<u4")
    b = np.unpackbits(bits.view(np.uint8), axis=1, bitorder="little")[:, :K]
    return np.where(b.T == 1, np.float32(1), np.float32(-1))


def write_checkpoint(path, units, use_bias, use_bn, weights, biases=None, bn=None, packed=False,
                     latent=True) -> Path:
    """weights: list of [K, N] float32; bn: (gammas, betas, mus, sigmas) lists; packed: add the sign words;
    latent=False drops the float32 weights (packed-only deployment file)."""
    assert latent or packed, "a checkpoint needs the latent or the packed weights"
    blob = {"units": np.array(units, dtype=np.uint32), "useBias": np.array(bool(use_bias)),
            "useBatchNorm": np.array(bool(use_bn))}
    for i, w in enumerate(weights):
        w = np.asarray(w, dtype=np.float32)
        assert w.shape == (units[i], units[i + 1]), (i, w.shape)
        if latent:
            blob[f"weights/weights_{i}"] = w
        if packed:
            blob[f"packed_weights/packed_weights_{i}"] = pack_weights_t(w)
    if use_bias:
        for i, b in enumerate(biases):
            blob[f"biases/biases_{i}"] = np.asarray(b, dtype=np.float32)
    if use_bn:
        for name, group in zip(BN_GROUPS, bn):
            for i, v in enumerate(group):
                blob[f"{name}/{name}_{i}"] = np.asarray(v, dtype=np.float32)
    path = Path(path)
    path.parent.mkdir(parents=True, exist_ok=True)
    with open(path, "wb") as fh:      # a file object: np.savez must not append ".npz" to e.g. "model.hdf5"
        np.savez(fh, **blob)
    return path


def read_checkpoint(path) -> dict:
    """-> dict(units, useBias, useBatchNorm, weights, biases, gammas, betas, mus, sigmas, packed_only).
    Members are looked up BY INDEX (weights_0, weights_1, ...), not in h5py's alphabetical iteration order that
    network.py:189-190 relies on (which misorders layers 10+)."""
    with np.load(Path(path)) as fp:
        names = set(fp.files)
        for key in ("units", "useBias", "useBatchNorm"):
            assert key in names, f"{path}: missing '{key}' (network.py:163-167)"
        units = [int(u) for u in fp["units"]]
        n = len(units) - 1
        use_bias, use_bn = bool(fp["useBias"]), bool(fp["useBatchNorm"])
        out = dict(units=units, useBias=use_bias, useBatchNorm=use_bn, biases=None, packed_only=False)
        if f"weights/weights_{0}" in names:
            out["weights"] = [fp[f"weights/weights_{i}"].astype(np.float32) for i in range(n)]
        else:
            assert "packed_weights/packed_weights_0" in names, f"{path}: neither 'weights' nor 'packed_weights'"
            out["weights"] = [unpack_weights_t(fp[f"packed_weights/packed_weights_{i}"], units[i]) for i in range(n)]
            out["packed_only"] = True
        for i, w in enumerate(out["weights"]):
            assert w.shape == (units[i], units[i + 1]), f"{path}: weights_{i} has shape {w.shape}"
        if use_bias:
            out["biases"] = [fp[f"biases/biases_{i}"].astype(np.float32) for i in range(n)]
        for g in BN_GROUPS:
            out[g] = [fp[f"{g}/{g}_{i}"].astype(np.float32) for i in range(n)] if use_bn else None
    return out
