"""CPU-side checks on the machine code of the built library (cuobjdump -sass; no GPU needed): the hot kernels are the
Blackwell-native ones (tcgen05 UMMA + TMA in the conv kernels, TMA store in the tap-stacked kernel), and the codegen
pathology of DESIGN.md section 4.1 -- a single-lane TMA producer loop compiles every UTMALDG into a lane-serialisation
loop ending in BRA.U.ANY -- stays out."""
import os
import re
import shutil
import subprocess

import pytest

from dualdiffusion_b200 import build

LIB = os.path.join(build.ROOT, "dualdiffusion_b200", "lib", "libdualdiffusion_b200.so")


@pytest.fixture(scope="module")
def sass():
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not available")
    build.build()
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels, name = {}, None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            kernels[name] = []
        elif name is not None:
            kernels[name].append(line)
    return kernels


def count(lines, mnemonic):
    pat = re.compile(r"\b" + re.escape(mnemonic) + r"\b")
    return sum(1 for l in lines if pat.search(l))


def pick(kernels, key):
    hit = {n: l for n, l in kernels.items() if key in n}
    assert hit, f"no kernel named *{key}* in the library"
    return hit


def test_conv_kernels_are_tcgen05_and_tma(sass):
    for key in ("conv3x3_dx_kernel", "conv_igemm_kernel", "conv3x3_halo_kernel", "conv_wgrad_kernel"):
        for name, lines in pick(sass, key).items():
            assert count(lines, "UTCHMMA") > 0, f"{name}: no tcgen05.mma"
            assert count(lines, "UTMALDG") > 0, f"{name}: no TMA load"
            assert count(lines, "HMMA") == 0, f"{name}: legacy mma.sync in a tcgen05 kernel"
    for name, lines in pick(sass, "conv3x3_dx_kernel").items():
        assert count(lines, "UTMASTG") > 0, f"{name}: the tap-stacked kernel stores its tiles with TMA"


def test_no_lane_serialised_tma_issue(sass):
    for key in ("conv3x3_dx_kernel", "conv_igemm_kernel", "conv3x3_halo_kernel", "conv_wgrad_kernel"):
        for name, lines in pick(sass, key).items():
            assert count(lines, "BRA.U.ANY") == 0, f"{name}: TMA issue wrapped in a lane-serialisation loop (DESIGN.md 4.1)"


def test_glue_kernels_use_one_mufu_transcendentals(sass):
    """the streaming glue kernels were bound by their instruction streams.  Their SiLU is one
    MUFU.TANH per element -- no exp + reciprocal sequence -- and the attention softmax one MUFU.EX2 per score without the
    denormal range test exp2f adds (an FSETP against -126 next to every MUFU)."""
    for key in ("cat_silu_kernel", "pixnorm_silu_kernel", "up2_silu_pad_kernel", "silu_scale_bwd_kernel"):
        for name, lines in pick(sass, key).items():
            assert count(lines, "MUFU.TANH") > 0, f"{name}: SiLU not on MUFU.TANH"
            assert count(lines, "MUFU.EX2") == 0, f"{name}: the exp + division SiLU is back"
    for name, lines in pick(sass, "attention_kernel").items():
        assert count(lines, "MUFU.EX2") > 0
        assert not any("-126" in l and "FSETP" in l for l in lines), f"{name}: range-tested exp2f in the softmax"
