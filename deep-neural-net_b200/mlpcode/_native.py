"""ctypes binding of libbnn_b200.so (the C ABI declared in include/bnn_b200.h).

There is deliberately no fallback: if the shared library is missing, or a call fails, an
exception is raised.  The library is built in-tree by `__graft_entry__.build()` /
`make -C deep-neural-net_b200/csrc` into deep-neural-net_b200/lib/.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_PKG = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ.get("BNN_B200_LIB", _PKG.parent / "lib" / "libbnn_b200.so"))

BNN_OK, BNN_ERR_ARG, BNN_ERR_CUDA, BNN_ERR_UNSUPPORTED = 0, -1, -2, -3
DT_I8, DT_F16, DT_F4 = 0, 1, 2
EPI_STORE_F32, EPI_STORE_S32, EPI_STORE_S16, EPI_SGD_F32, EPI_SIGN_I8, EPI_SCATTER_F32, EPI_SIGN_F4 = 0, 1, 2, 3, 4, 5, 6
MAX_RANKS = 16
STATS_NONE, STATS_COLSUM, STATS_BN_BWD = 0, 1, 2
GEMM_TCGEN05, GEMM_SIMT = 0, 1
ACT_IDENTITY, ACT_SIGMOID, ACT_TANH, ACT_RELU, ACT_LEAKY_RELU, ACT_HARD_SIGMOID, ACT_HARD_TANH = range(7)
LOSS_CROSS_ENTROPY, LOSS_MSE, LOSS_HINGE = range(3)
Z_F32, Z_S16, Z_S32 = 0, 1, 2
GEMM_MAX_PASS = 4


class BnnError(RuntimeError):
    """A libbnn_b200 call returned a non-zero status."""

    def __init__(self, fn: str, code: int, text: str):
        super().__init__(f"{fn} failed ({code}): {text}")
        self.code = code


class PeerScatter(C.Structure):
    """Mirror of `struct bnn_peer_scatter`."""

    _fields_ = [("world", C.c_int32), ("rank", C.c_int32), ("rows_per_owner", C.c_int32),
                ("reserved", C.c_int32), ("stage", C.c_uint64 * MAX_RANKS)]


class GemmDesc(C.Structure):
    """Mirror of `struct bnn_gemm_desc`."""

    _fields_ = [
        ("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32),
        ("batch", C.c_int32), ("in_dtype", C.c_int32), ("epilogue", C.c_int32),
        ("n_pass", C.c_int32),
        ("pa", C.c_int32 * GEMM_MAX_PASS), ("pb", C.c_int32 * GEMM_MAX_PASS),
        ("A", C.c_void_p * 2), ("B", C.c_void_p * 2),
        ("lda", C.c_int64), ("ldb", C.c_int64),
        ("stride_a", C.c_int64), ("stride_b", C.c_int64),
        ("D", C.c_void_p), ("ldd", C.c_int64), ("stride_d", C.c_int64),
        ("sa", C.c_void_p), ("sb", C.c_void_p),
        ("host_scale", C.c_float), ("a_unsigned", C.c_int32),
        ("epi_thr", C.c_void_p), ("epi_sgn", C.c_void_p),
        ("scatter", C.POINTER(PeerScatter)),
        ("stats", C.c_int32), ("stats_z_dtype", C.c_int32),
        ("stats_sums", C.c_void_p), ("stats_max", C.c_void_p), ("stats_z", C.c_void_p),
        ("stats_ldz", C.c_int64), ("stats_mu", C.c_void_p),
        ("B2", C.c_void_p), ("hi_from_pass", C.c_int32), ("epi_thr_stride", C.c_int32),
        ("col_scale", C.c_void_p),
    ]


_p, _i, _i64, _u32, _u64, _f, _d = (C.c_void_p, C.c_int, C.c_int64, C.c_uint32, C.c_uint64,
                                    C.c_float, C.c_double)

# name -> argtypes (return type is always int unless listed in _RESTYPE)
PROTOTYPES = {
    "bnn_version": [],
    "bnn_device_query": [C.POINTER(C.c_int)] * 3,
    "bnn_pack_sign_rows": [_p, _i, _i, _i64, _p, _i64, _p],
    "bnn_pack_i8_rows": [_p, _i, _i, _i64, _p, _i64, _p],
    "bnn_binarize_weights_t": [_p, _i, _i, _p, _i64, _p, _i64, _p, _i64, _p, _i64, _p, _p],
    "bnn_unpack_bits": [_p, _i, _i, _i64, _p, _i64, _p, _i64, _p],
    "bnn_bits_to_f4": [_p, _i64, _i, _i64, _p, _i64, _p],
    "bnn_gather_rows": [_p, _i64, _i64, _p, _i, _i64, _p, _i64, _p, _p],
    "bnn_activation": [_p, _i64, _i, _p, _p],
    "bnn_activation_grad": [_p, _p, _i64, _i, _p, _p],
    "bnn_softmax_grad": [_p, _p, _i, _i, _p, _p],
    "bnn_loss_rows": [_p, _p, _i, _i, _i, _p, _p, _p],
    "bnn_loss_grad": [_p, _p, _i, _i, _i, _d, _p, _p],
    "bnn_absmax_f32": [_p, _i64, _p, _i, _p],
    "bnn_split_f16": [_p, _i, _i, _i64, _p, _p, _p, _i64, _p, _p, _i64, _p, _p],
    "bnn_gemm": [C.POINTER(GemmDesc), _i, _p],
    "bnn_zero_async": [_p, _i64, _p],
    "bnn_gemm_popc": [_p, _i64, _p, _i64, _i, _i, _i, _p, _i64, _i, _p],
    "bnn_colstats": [_p, _i, _i, _i, _i64, _p, _i, _p],
    "bnn_bn_stats_finalize": [_p, _i, _d, _p, _p, _p],
    "bnn_bn_apply": [_p, _i, _i, _i, _i64, _p, _p, _p, _p, _f, _p, _i64, _p, _i64, _p, _i64, _p,
                     _i64, _p],
    "bnn_bn_sign_thresholds": [_p, _p, _p, _p, _f, _i, _p, _p, _p],
    "bnn_bn_sign": [_p, _i, _i, _i, _i64, _p, _p, _p, _i64, _p, _i64, _p, _i64, _p, _p, _i64, _p, _i64, _p],
    "bnn_softmax_xent": [_p, _i, _i, _i64, _p, _p, _p, _p, _d, _p, _p],
    "bnn_bn_bwd_reduce": [_p, _i64, _p, _i, _i64, _i, _i, _p, _p, _p, _i, _p],
    "bnn_bn_bwd_finalize": [_p, _p, _i, _d, _p, _p, _f, _p, _p, _p, _p, _f, _d, _p, _p, _p, _p, _p,
                            _p, _p],
    "bnn_bn_bwd_apply_q": [_p, _i64, _p, _i, _i64, _i, _i, _p, _p, _p, _p, _p, _i64, _p, _p, _p, _i64,
                           _p, _p],
    "bnn_bn_bwd_apply": [_p, _i64, _p, _i, _i64, _i, _i, _p, _p, _p, _p, _i64, _p, _p, _i64, _p, _p,
                         _i64, _p, _p],
    "bnn_step_tick": [_p, _p],
    "bnn_bn_stats_finalize_p2p": [_p, _p, _i, _i, _u32, _p, _i, _d, _p, _p, _p, _p],
    "bnn_allreduce_f64_p2p": [_p, _p, _i, _i, _u32, _p, _i, _p, _p],
    "bnn_sgd_update": [_p, _p, _i64, _f, _i, _i64, _p],
    "bnn_peer_signal": [_p, _i, _i, _i, _u32, _p, _p],
    "bnn_peer_wait": [_p, _i, _i, _i, _i, _i, _u32, _p, _p],
    "bnn_dw_reduce_sgd_p2p": [_p, _p, _p, _i, _i, _i, _i, _u32, _p, _i, _i, _i, _f, _i, _p, _p, _p],
    "bnn_peer_broadcast_rows": [_p, _p, _i, _i, _i, _i, _i, _i, _p, _p],
    "bnn_flip_weights_bernoulli": [_p, _i64, _p, _i64, _i, _i, _u64, _u64, _u32, _u32, _i, _p, _i64,
                                   _p, _i64, _i64, _p, _i64, _p, _i64, _p],
    "bnn_flip_weights_exactk": [_i, _i, _u64, _u64, _u64, _u32, _i, _p, _i, _p, _i64, _i64, _p, _i64,
                                _i64, _p, _i64, _i64, _p],
    "bnn_flip_weights_mask": [_p, _i, _i, _p, _i64, _p, _i64, _p, _i64, _p],
    "bnn_flip_f32_bits": [_p, _i64, _i64, _d, _i, _u64, _u32, _p],
    "bnn_flip_images": [_p, _i, _i, _i, _i, _u64, _u32, _i, _p, _p],
    "bnn_flip_images_ex": [_p, _i, _i, _i, _i, _u64, _u32, _i, _p, _i, _p, _p],
    "bnn_u8_input_thresholds": [_p, _i64, _i64, _i, _i, _p, _p, _p, _i, _i, _f, _f, _i, _p, _i64, _p],
    "bnn_normalize_u8": [_p, _i64, _p, _p, _f, _f, _p, _p],
    "bnn_minmax_u8": [_p, _i64, _p, _p, _p],
    "bnn_count_correct": [_p, _i, _i, _i64, _p, _p, _i, _p],
}
_RESTYPE = {"bnn_last_error": C.c_char_p}

_lib = None


def load() -> C.CDLL:
    """Load the shared library (once) and attach prototypes.  Raises if it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; "
            "g.build()'` or `make -C deep-neural-net_b200/csrc` — there is no CPU fallback")
    lib = C.CDLL(str(LIB_PATH))
    lib.bnn_last_error.restype = C.c_char_p
    lib.bnn_last_error.argtypes = []
    for name, argtypes in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError here = header/library mismatch
        fn.argtypes = argtypes
        fn.restype = C.c_int
    _lib = lib
    return lib


def exported_symbols():
    return ["bnn_last_error", *PROTOTYPES.keys()]


def call(name: str, *args) -> None:
    """Invoke `name`, raising BnnError on a non-zero status."""
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        raise BnnError(name, rc, (lib.bnn_last_error() or b"").decode(errors="replace"))


def version() -> int:
    return load().bnn_version()


def device_query():
    sm, maj, mnr = C.c_int(), C.c_int(), C.c_int()
    call("bnn_device_query", C.byref(sm), C.byref(maj), C.byref(mnr))
    return sm.value, maj.value, mnr.value
