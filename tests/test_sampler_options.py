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
        out = sampler_oracle.diffusion_decode(sd, spec, g["clap"], tuple(g["shape"]), seed=case["seed"],
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


# ---------------------------------------------------------------------------------------------------------------------
# host logic of DualDiffusionPipeline.diffusion_decode (shift sequence, padded frame, noise handling, CFG / Heun order)
# with the C-ABI calls replaced by CPU stand-ins that follow include/dualdiffusion_b200.h and the oracle as the UNet
# ---------------------------------------------------------------------------------------------------------------------
class _OracleUNet(torch.nn.Module):
    device = torch.device("cpu")

    def __init__(self, sd, spec):
        super().__init__()
        self.sd, self.spec = sd, spec
        self.config = type("C", (), dict(sigma_max=spec.sigma_max, sigma_min=spec.sigma_min, sigma_data=spec.sigma_data))()

    def get_embeddings(self, emb_in, mask):
        return uo.get_embeddings(self.sd, emb_in, mask)

    def forward(self, x, sigma, fmt, emb, x_ref=None):
        return uo.unet_forward(self.sd, self.spec, x, sigma, emb, x_ref)


def _cpu_ops(monkeypatch):
    from dualdiffusion_b200 import ops

    def cfg_lerp(d, sample, cfg_scale, t_hat, cfg_out, x_hat_out, dup=False):
        B = cfg_out.shape[0]
        cfg = d if dup & 2 else d[B:].lerp(d[:B], cfg_scale)
        cfg_out.copy_(cfg)
        if x_hat_out is not None:
            xh = torch.lerp(cfg, sample[:B], t_hat)
            x_hat_out.copy_(xh.repeat(x_hat_out.shape[0] // B, 1, 1, 1))

    def update(cfg1, d2, cfg_scale, use_heun, t, p, noise, sample, cfg_out, dup=False):
        B = cfg1.shape[0]
        cfg = cfg1
        if use_heun:
            cfg2 = d2 if dup & 2 else d2[B:].lerp(d2[:B], cfg_scale)
            cfg = torch.lerp(cfg1, cfg2, 0.5)
        new = torch.lerp(cfg, sample[:B], t)
        if noise is not None:
            new = new + p * noise
        sample.copy_(new.repeat(sample.shape[0] // B, 1, 1, 1))
        if cfg_out is not None:
            cfg_out.copy_(cfg)

    def roll_pad(x, shift, pad, copies=1, out=None):
        r = torch.roll(x, shifts=shift, dims=-1)
        W = x.shape[-1]
        res = torch.cat((r[..., W - pad:], r, r[..., :pad]), dim=-1).repeat(copies, *([1] * (x.ndim - 1)))
        if out is None:
            return res
        out.copy_(res)
        return out

    def crop_unroll(xp, shift, pad, out=None):
        W = xp.shape[-1] - 2 * pad
        res = torch.roll(xp[..., pad:pad + W], shifts=-shift, dims=-1)
        if out is None:
            return res
        out.copy_(res)
        return out

    def stereo(noise, fresh, t):
        n = noise.clone()
        n[:, ::2] = n[:, 1::2]
        return uo.mp_sum(fresh, n, t)

    monkeypatch.setattr(ops, "sampler_cfg_lerp", cfg_lerp)
    monkeypatch.setattr(ops, "sampler_update", update)
    monkeypatch.setattr(ops, "roll_pad_w", roll_pad)
    monkeypatch.setattr(ops, "crop_unroll_w", crop_unroll)
    monkeypatch.setattr(ops, "stereo_fix_noise", stereo)


@pytest.mark.parametrize("golden", ["sampler_small.pt", "sampler_options_small.pt"])
def test_diffusion_decode_host_logic_replays_reference_goldens(monkeypatch, golden):
    from dualdiffusion_b200.pipelines.dual_diffusion_pipeline import DualDiffusionPipeline, SampleParams
    _cpu_ops(monkeypatch)
    g = torch.load(os.path.join(ROOT, "tests", "golden", golden), weights_only=False)
    spec = uo.small_spec()
    sd = uo.synth_state_dict(spec, seed=0)
    pipe = DualDiffusionPipeline({"unet": _OracleUNet(sd, spec)})
    pipe.collect_debug_info = True
    for name, case in g["cases"].items():
        x_ref = g["x_ref"] if case.get("use_ref") else None
        params = SampleParams(seed=case["seed"], batch_size=1, **case["kwargs"])
        # a CPU generator seeded like the reference's: the default noise path (no injection) must reproduce the golden
        out = pipe.diffusion_decode(params, quiet=True, audio_embedding=g["clap"],
                                    sample_shape=tuple(g.get("shape", (1, 4, 32, 48))),
                                    x_ref=x_ref, stereo_noise=case.get("stereo_noise"))
        assert rel_err(out, case["sample"]) < 1e-4, name
        assert len(pipe.last_debug_info["sample_std"]) == case["kwargs"]["num_steps"]
