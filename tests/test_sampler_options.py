"""CPU tests of the sampler-loop options (`seamless_loop`, `stereo_fix`; pipelines/dual_diffusion_pipeline.py:638-656,
:729-732): the oracle against a golden from the unmodified reference, and the index maps / arithmetic of csrc/sampler.cu
(compiled for the host from csrc/sampler_math.cuh) bit for bit against torch.roll / torch.cat / the reference's mp_sum."""
import ctypes
import os
import subprocess

import pytest
import torch

from oracle import sampler_oracle, unet_oracle as uo

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "sampler_options_small.pt")


def rel_err(a, b):
    return float((a - b).norm() / b.norm())


def test_sampler_oracle_options_vs_reference_golden():
    g = torch.load(GOLD, weights_only=False)
    spec = uo.small_spec()
    sd = uo.synth_state_dict(spec, seed=0)
    for name, case in g["cases"].items():
        rec = {}
        out = sampler_oracle.diffusion_decode(sd, spec, g["clap"], (1, 4, 32, 48), seed=case["seed"],
                                              x_ref=g["x_ref"] if case["use_ref"] else None,
                                              stereo_noise=case["stereo_noise"], record=rec, **case["kwargs"])
        assert rel_err(out, case["sample"]) < 1e-4, name
        if case["kwargs"].get("seamless_loop"):
            assert len(rec["loop_shifts"]) == case["kwargs"]["num_steps"] and all(0 <= s < 48 for s in rec["loop_shifts"])


@pytest.fixture(scope="module")
def host(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("sampler_host") / "libsampler_host_check.so")
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off",
                    "-I", os.path.join(ROOT, "dualdiffusion_b200", "csrc"), "-x", "c++",
                    os.path.join(ROOT, "tests", "csrc", "sampler_host_check.cpp"), "-o", out], check=True)
    lib = ctypes.CDLL(out)
    vp, ci, cl, cf = ctypes.c_void_p, ctypes.c_int, ctypes.c_long, ctypes.c_float
    lib.host_roll_pad_w.argtypes = [vp, vp, cl, ci, ci, ci, ci]
    lib.host_crop_unroll_w.argtypes = [vp, vp, cl, ci, ci, ci]
    lib.host_stereo_fix_noise.argtypes = [vp, vp, cf, vp, ci, ci, cl]
    for f in (lib.host_roll_pad_w, lib.host_crop_unroll_w, lib.host_stereo_fix_noise):
        f.restype = None
    return lib


@pytest.mark.parametrize("W,pad,shift", [(48, 32, 0), (48, 32, 17), (48, 32, 47), (688, 32, 300), (33, 33, 5), (40, 0, 39)])
def test_roll_pad_and_crop_unroll_index_maps_are_bit_exact(host, W, pad, shift):
    gen = torch.Generator().manual_seed(W + shift)
    x = torch.randn(2, 3, 4, W, generator=gen)
    r = torch.roll(x, shifts=shift, dims=-1)
    ref = torch.cat((r[..., W - pad:], r, r[..., :pad]), dim=-1).repeat(2, 1, 1, 1)
    out = torch.full((4, 3, 4, W + 2 * pad), float("nan"))
    host.host_roll_pad_w(x.data_ptr(), out.data_ptr(), 2 * 3 * 4, W, shift, pad, 2)
    assert torch.equal(out, ref)
    back = torch.full_like(x, float("nan"))
    host.host_crop_unroll_w(out.data_ptr(), back.data_ptr(), 2 * 3 * 4, W, shift, pad)
    assert torch.equal(back, x)                                            # round trip
    y = torch.randn(2, 3, 4, W + 2 * pad, generator=gen)                   # independent of the forward map
    ref_back = torch.roll(y[..., pad:pad + W], shifts=-shift, dims=-1)
    host.host_crop_unroll_w(y.data_ptr(), back.data_ptr(), 2 * 3 * 4, W, shift, pad)
    assert torch.equal(back, ref_back)


@pytest.mark.parametrize("t", [0.3, 0.5, 0.8])
def test_stereo_fix_noise_matches_reference_formula(host, t):
    gen = torch.Generator().manual_seed(11)
    noise, fresh = torch.randn(2, 4, 6, 10, generator=gen), torch.randn(2, 4, 6, 10, generator=gen)
    ref = noise.clone()
    ref[:, ::2] = ref[:, 1::2]
    ref = uo.mp_sum(fresh, ref, t)
    out = torch.empty_like(noise)
    host.host_stereo_fix_noise(noise.data_ptr(), fresh.data_ptr(), t, out.data_ptr(), 2, 4, 60)
    torch.testing.assert_close(out, ref, rtol=2e-7, atol=1e-7)
    torch.testing.assert_close(out[:, 0], uo.mp_sum(fresh[:, 0], noise[:, 1], t), rtol=2e-7, atol=1e-7)
