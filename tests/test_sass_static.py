"""Static evidence from the compiled object (cuobjdump -sass, no GPU): which tensor-core / TMA / cluster
instructions each GEMM kernel instantiation contains.  The single-CTA kernels (the product path) must hold
tcgen05 MMAs and TMA loads and NO CTA-pair instruction; the opt-in pair-form instantiations must hold
the 2-CTA forms of all of them; the FP4 instantiations hold the block-scaled MMA (UTCOMMA) and nothing else.
The last two template arguments of gemm_tc_kernel are <kCta2, kF4>."""
import re
import shutil
import subprocess

import pytest

from conftest import PKG_ROOT

OBJ = PKG_ROOT / "build" / "gemm_tc.o"
PAT = re.compile(r"\b(UTC[A-Z0-9_.]*MMA[A-Z0-9_.]*|UTMALDG[A-Z0-9_.]*|UTCBAR[A-Z0-9_.]*|UTCATOMSWS[A-Z0-9_.]*|UCGABAR_[A-Z]+)\b")


@pytest.fixture(scope="module")
def kernels():
    if shutil.which("cuobjdump") is None or not OBJ.exists():
        pytest.skip("needs cuobjdump and the built gemm_tc.o")
    out = subprocess.run(["cuobjdump", "-sass", str(OBJ)], capture_output=True, text=True, check=True).stdout
    found, name = {}, None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = m.group(1)
            found[name] = set()
        elif name:
            m = PAT.search(line)
            if m:
                found[name].add(m.group(1))
    return {k: v for k, v in found.items() if "gemm_tc_kernel" in k}


def test_single_cta_kernels_use_tcgen05_and_tma_only(kernels):
    single = {k: v for k, v in kernels.items() if k.endswith("ELb0ELb0EEEv14CUtensorMap_stS2_S2_S2_S2_NS0_8TcParamsE")}
    assert len(single) >= 40
    for name, ops in single.items():
        assert any(o.startswith(("UTCIMMA", "UTCHMMA")) for o in ops), name        # tcgen05.mma kind::i8 / kind::f16
        assert "UTMALDG.3D" in ops, name                                            # cp.async.bulk.tensor.3d
        # no 2-CTA instruction; cluster barriers / multicast commits belong to the run-time multicast form
        assert not any("2CTA" in o for o in ops), (name, ops)


def test_pair_kernels_use_the_two_cta_forms(kernels):
    pair = {k: v for k, v in kernels.items() if "ELb1ELb0EEEv" in k}
    assert len(pair) >= 8
    for name, ops in pair.items():
        assert any(o in ("UTCIMMA.2CTA", "UTCHMMA.2CTA") for o in ops), (name, ops)
        assert "UTMALDG.3D.2CTA" in ops and "UTCBAR.2CTA.MULTICAST" in ops, (name, ops)
        assert "UCGABAR_ARV" in ops and "UCGABAR_WAIT" in ops, (name, ops)
        assert not any(o in ("UTCIMMA", "UTCHMMA", "UTMALDG.3D", "UTCBAR") for o in ops), (name, ops)


def test_fp4_kernels_use_the_block_scaled_mma(kernels):
    f4 = {k: v for k, v in kernels.items() if "ELb0ELb1EEEv" in k}
    assert len(f4) == 18            # 3 tile widths x (4 integer epilogues + 2 stores with column sums)
    for name, ops in f4.items():
        assert any(o.startswith("UTCOMMA") for o in ops), (name, ops)             # tcgen05.mma kind::mxf4.block_scale
        assert "UTMALDG.3D" in ops, name
        assert not any(o.startswith(("UTCIMMA", "UTCHMMA")) or "2CTA" in o for o in ops), (name, ops)
