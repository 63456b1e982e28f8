"""CPU: the C-ABI library loads and exports every symbol include/bnn_b200.h declares; host-side logic
that needs no device (LR schedule, shard ranges, thresholds, argument validation, lazy construction)."""
import ctypes
import re
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

REPO = Path(__file__).resolve().parent.parent
HEADER = REPO / "include" / "bnn_b200.h"


def declared_functions():
    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    return sorted(set(re.findall(r"\b(bnn_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from mlpcode import _native
    lib = _native.load()
    names = declared_functions()
    assert len(names) >= 25
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/bnn_b200.h but not exported"
    # and the ctypes table covers exactly the header
    assert sorted(_native.exported_symbols()) == names
    assert _native.version() == 100


def test_header_compiles_as_plain_c(tmp_path):
    src = tmp_path / "t.c"
    src.write_text('#include "bnn_b200.h"\nint main(void){bnn_gemm_desc d; (void)d; return BNN_OK;}\n')
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", str(REPO / "include"), "-c",
                           str(src), "-o", str(tmp_path / "t.o")])


def test_gemm_desc_layout_matches_c(tmp_path):
    """ctypes mirror of bnn_gemm_desc has the C compiler's size and field offsets."""
    from mlpcode._native import GemmDesc
    fields = ["M", "n_pass", "pa", "pb", "A", "B", "lda", "stride_a", "D", "ldd", "sa", "host_scale",
              "a_unsigned", "epi_thr_stride", "epi_thr", "epi_sgn", "scatter", "stats", "stats_z_dtype", "stats_sums", "stats_max",
              "stats_z", "stats_ldz", "stats_mu", "B2", "hi_from_pass", "col_scale"]
    prog = ['#include <stdio.h>', '#include <stddef.h>', '#include "bnn_b200.h"', 'int main(void){',
            'printf("%zu\\n", sizeof(bnn_gemm_desc));']
    prog += [f'printf("%zu\\n", offsetof(bnn_gemm_desc, {f}));' for f in fields]
    prog += ['return 0;}']
    (tmp_path / "o.c").write_text("\n".join(prog))
    subprocess.check_call(["gcc", "-I", str(REPO / "include"), str(tmp_path / "o.c"), "-o",
                           str(tmp_path / "o")])
    out = [int(x) for x in subprocess.check_output([str(tmp_path / "o")]).split()]
    assert out[0] == ctypes.sizeof(GemmDesc)
    for f, off in zip(fields, out[1:]):
        assert getattr(GemmDesc, f).offset == off, f


def test_argument_errors_do_not_need_a_device():
    from mlpcode import _native as nat
    lib = nat.load()
    d = nat.GemmDesc()
    assert lib.bnn_gemm(ctypes.byref(d), 0, None) == nat.BNN_ERR_ARG
    assert b"empty shape" in lib.bnn_last_error()
    assert lib.bnn_gemm(None, 0, None) == nat.BNN_ERR_ARG
    with pytest.raises(nat.BnnError) as e:
        nat.call("bnn_flip_images", None, 1, 1, 1, 7, 0, 0, 1, None, None)
    assert e.value.code == nat.BNN_ERR_ARG


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: no product source may import, include, load or run it."""
    pkg = REPO / "deep-neural-net_b200"
    for path in pkg.rglob("*.py"):
        text = path.read_text()
        assert not re.search(r"bnn_oracle|gen_golden|refshim|[\"']oracle[\"'/]", text), path
    for path in list(pkg.rglob("*.cu")) + list(pkg.rglob("*.cuh")) + [pkg / "csrc" / "Makefile"]:
        for line in path.read_text().splitlines():
            if line.lstrip().startswith(("#include", "SRCS", "-l", "LIB")):
                assert "oracle" not in line, (path, line)


def test_lr_scheduler_matches_reference_formulae():
    from mlpcode.optim import LRScheduler, LRSchedulerStrat as S
    s = LRScheduler(alpha=0.1)
    assert s.value == 0.1
    s.step()
    assert s.value == 0.1
    t = LRScheduler(alpha=0.5, decay_rate=0.1, strategy=S.time)
    vals = [t.value]
    for _ in range(3):
        t.step()
        vals.append(t.value)
    assert np.allclose(vals, [0.5, 0.5 / 1.1, 0.5 / 1.2, 0.5 / 1.3])
    e = LRScheduler(alpha=1.0, decay_rate=0.5, strategy=S.exp)
    e.step()
    e.step()
    assert e.value == 0.25
    d = LRScheduler(alpha=1.0, decay_rate=2.0, strategy=S.drop, drop_every_n=3)
    seen = [d.value]
    for _ in range(4):
        d.step()
        seen.append(d.value)
    assert seen == [1.0, 1.0, 0.5, 0.5, 0.5]
    with pytest.raises(ValueError):
        LRScheduler(decay_rate=2.0, strategy=S.exp)
    with pytest.raises(ValueError):
        LRScheduler(decay_rate=0.5, strategy=S.drop)


def test_shard_range_partitions_exactly():
    from mlpcode.parallel import shard_range
    for n in (0, 1, 7, 8, 10_000, 10_001):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_dw_exchange_layout_covers_every_row_once():
    """Row ownership of the fused dW reduce-scatter: whole 128-row tiles, every row owned by exactly
    one rank, regions aligned and disjoint."""
    from mlpcode.parallel import dw_exchange_layout
    shapes = [(784, 4096), (4096, 4096), (4096, 10), (100, 7), (128, 128)]
    for world in (2, 3, 4, 8, 16):
        flag_bytes, layout, total = dw_exchange_layout(shapes, world)
        assert flag_bytes >= 2 * len(shapes) * 16 * 4 and flag_bytes % 1024 == 0
        end = flag_bytes
        for (k, n), L in zip(shapes, layout):
            rpo = L["rows_per_owner"]
            assert rpo % 128 == 0 and rpo * world >= k and rpo * world - k < 128 * world
            owned = [max(0, min(rpo, k - r * rpo)) for r in range(world)]
            assert sum(owned) == k
            assert L["w_off"] == end and L["w_off"] % 1024 == 0 and L["stage_off"] % 1024 == 0
            assert L["stage_off"] >= L["w_off"] + k * n * 4
            end = L["stage_off"] + -(-(world * rpo * n * 4) // 1024) * 1024
        assert end == total


def test_thresholds_and_flip_counts():
    from mlpcode.callbacks import ImageErrorCallback
    from mlpcode.kernels import bernoulli_threshold
    assert bernoulli_threshold(0.0) == 0 and bernoulli_threshold(1.0) == 1 << 32
    assert bernoulli_threshold(0.5) == 1 << 31 and bernoulli_threshold(-1) == 0
    assert ImageErrorCallback(0.14).flips_per_row(784) == 878
    assert ImageErrorCallback(0.5).flips_per_row(1) == 4
    assert ImageErrorCallback(0.0625).flips_per_row(1) == 0          # round(0.5) -> 0 (half to even)


def test_network_constructs_without_a_device():
    from mlpcode.network import Network
    nn = Network([784, 256, 256, 10], useGpu=True, binarized=True, useBias=True, useBatchNorm=True)
    assert not nn.useBias and len(nn._layers) == 3 and not nn.isCompiled
    assert nn._layers[0].skip_input_grad and not nn._layers[1].skip_input_grad
    assert nn._layers[1].weights_shape == (256, 256) and nn.biases == [None, None, None]
    import torch
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU fallback|no CUDA"):
            nn.compile(lr=0.01)
