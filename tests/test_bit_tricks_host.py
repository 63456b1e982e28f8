"""Host-side check of the bit expansions the E2M1 (FP4) operand writers share (csrc/common.cuh:
spread_bits_to_nibbles, bits_to_e2m1x8 — used by bnn_bits_to_f4, bnn_binarize_weights_t(wt_f4), bnn_bn_sign(act_f4)
and the BNN_EPI_SIGN_F4 epilogue).  They are __host__ __device__, so nvcc builds them into a host program and all 256
inputs are compared with the definition: sign bit j (1 = +1) -> nibble j = 0x2 (+1.0) or 0xA (-1.0), low nibble =
even element.  No GPU needed."""
import shutil
import subprocess
from pathlib import Path

import pytest

REPO = Path(__file__).resolve().parent.parent
SRC = r'''
#include <cstdio>
#include "common.cuh"
int main() {
    for (unsigned b = 0; b < 256; ++b)
        std::printf("%u %08x %08x\n", b, bnn::spread_bits_to_nibbles(b), bnn::bits_to_e2m1x8(b));
    // bits above the low byte must be ignored
    std::printf("hi %08x %08x\n", bnn::spread_bits_to_nibbles(0xFFFFFF00u), bnn::bits_to_e2m1x8(0xABCDEF00u));
    return 0;
}
'''


@pytest.mark.skipif(shutil.which("nvcc") is None, reason="nvcc not on PATH")
def test_bits_to_e2m1_nibbles_on_the_host(tmp_path):
    cu = tmp_path / "tricks.cu"
    cu.write_text(SRC)
    exe = tmp_path / "tricks"
    subprocess.run(["nvcc", "-std=c++17", "-O1", "-I", str(REPO / "deep-neural-net_b200" / "csrc"), "-o", str(exe),
                    str(cu)], check=True, capture_output=True, timeout=300)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True, timeout=60).stdout.splitlines()
    assert len(out) == 257
    for line in out[:256]:
        b, spread, e2m1 = line.split()
        b, spread, e2m1 = int(b), int(spread, 16), int(e2m1, 16)
        want_spread = sum(((b >> j) & 1) << (4 * j) for j in range(8))
        want_e2m1 = sum((0x2 if (b >> j) & 1 else 0xA) << (4 * j) for j in range(8))
        assert spread == want_spread and e2m1 == want_e2m1, (b, hex(spread), hex(e2m1))
    _, spread_hi, e2m1_hi = out[256].split()
    assert int(spread_hi, 16) == 0 and int(e2m1_hi, 16) == 0xAAAAAAAA
