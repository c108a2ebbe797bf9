"""CPU tests of the dataset pre-encode compute (dualdiffusion_b200/dataset/encode.py) against a golden
produced by the unmodified reference `EncodeProcess.process` with its own format and a reduced DAE_D3: the host logic
(augmentation order, the reference's batching quirks, embedding handling) runs here with the two GPU operators replaced by
their CPU oracles; the reference computes the DAE in bf16, the oracle in fp32 on the bf16-rounded mel: 3e-2 relative L2."""
import os

import pytest
import torch

from dualdiffusion_b200.dataset import encode as enc
from oracle import dae_oracle as do, format_oracle as fo, unet_oracle as uo

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "encode_small.pt")


class _OracleFormat:
    """Shape helpers from the product format class (host-only code), the transform from the CPU oracle."""

    def __init__(self):
        from dualdiffusion_b200.modules.formats.ms_mdct_dual import MS_MDCT_DualFormat, MS_MDCT_DualFormatConfig
        self._fmt = MS_MDCT_DualFormat(MS_MDCT_DualFormatConfig())
        self.config = self._fmt.config

    def get_raw_crop_width(self, n):
        return self._fmt.get_raw_crop_width(n)

    def raw_to_mel_spec(self, raw):
        return fo.raw_to_mel_spec(raw, fo.MSDualSpec())


class _OracleDAE:
    dtype = torch.float32

    def __init__(self):
        self.spec = do.small_dae_spec()
        self.sd = do.synth_dae_state_dict(self.spec, seed=0)
        self.calls = []

    def get_embeddings(self, emb):
        return do.dae_get_embeddings(self.sd, emb)

    def encode(self, x, emb):
        self.calls.append(("encode", tuple(x.shape)))
        return do.dae_encode(self.sd, self.spec, x.float())

    def tiled_encode(self, x, emb, max_chunk=6144, overlap=256):
        self.calls.append(("tiled", tuple(x.shape), max_chunk, overlap))
        assert x.shape[-1] <= max_chunk
        return do.dae_encode(self.sd, self.spec, x.float())


def rel_err(a, b):
    return float((a.float() - b.float()).norm() / b.float().norm())


def test_encode_latents_replays_reference_encode_process(monkeypatch):
    g = torch.load(GOLD, weights_only=False)
    monkeypatch.setattr(enc, "normalize", uo.normalize)
    fmt, dae = _OracleFormat(), _OracleDAE()
    assert abs(float(sum(v.double().abs().sum() for v in dae.sd.values())) - g["weight_checksum"]) < 1e-6 * g["weight_checksum"]
    for name, case in g["cases"].items():
        cfg = enc.EncodeLatentsConfig(**case["config"])
        dae.calls.clear()
        out = enc.encode_latents(g["audio"], g["clap"], fmt, dae, cfg)
        ref = case["latents"]
        assert out.dtype == torch.bfloat16 and out.shape == ref.shape, name
        assert rel_err(out, ref) < 3e-2, (name, rel_err(out, ref))
        kinds = {c[0] for c in dae.calls}
        assert kinds == ({"tiled"} if cfg.latents_tiled_encode else {"encode"})
        assert all(c[1][0] == cfg.latents_batch_size for c in dae.calls)


def test_augmented_audio_order_and_reference_batching_quirk():
    fmt = _OracleFormat()
    hop = fmt.config.ms_frame_hop_length
    audio = torch.arange(2 * 45000, dtype=torch.float32).view(2, 45000)
    cfg = enc.EncodeLatentsConfig(latents_num_time_offset_augmentations=3, latents_stereo_mirroring_augmentation=True)
    v = enc.augmented_audio(audio, fmt, cfg)
    w = fmt.get_raw_crop_width(45000 - 3 * hop)
    assert v.shape == (6, 2, w)
    for i in range(3):
        assert torch.equal(v[2 * i], audio[:, i * hop:i * hop + w])
        assert torch.equal(v[2 * i + 1], audio[:, i * hop:i * hop + w].flip(0))
    cfg2 = enc.EncodeLatentsConfig(latents_num_time_offset_augmentations=3, latents_stereo_mirroring_augmentation=False)
    assert enc.augmented_audio(audio, fmt, cfg2).shape == (3, 2, w)
