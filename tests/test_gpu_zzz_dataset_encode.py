"""GPU parity test of the dataset pre-encode compute: `encode_latents` with the product format and DAE
through the C ABI against the golden produced by the unmodified reference `EncodeProcess.process` (bf16 DAE body: 3e-2
relative L2, as the other whole-network comparisons).

The file name sorts last on purpose: written after the round's GPU budget was spent; not yet run on a GPU (the operators
it composes -- raw_to_mel_spec, DAE_D3.encode / tiled_encode -- have their own green GPU tests)."""
import os

import pytest
import torch

from oracle import dae_oracle as do

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "encode_small.pt")


def test_encode_latents_vs_reference_encode_process():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from dualdiffusion_b200.dataset.encode import EncodeLatentsConfig, encode_latents
    from dualdiffusion_b200.modules.daes.dae_edm2_d3 import DAE_D3, DAE_D3_Config
    from dualdiffusion_b200.modules.formats.ms_mdct_dual import MS_MDCT_DualFormat, MS_MDCT_DualFormatConfig
    dev = torch.device("cuda:0")
    g = torch.load(GOLD, weights_only=False)
    spec = do.small_dae_spec()
    sd = do.synth_dae_state_dict(spec, seed=0)
    cfg = DAE_D3_Config(in_channels_emb=spec.in_channels_emb, model_channels=spec.model_channels,
                        channel_mult_enc=spec.channel_mult_enc, channel_mult_dec=tuple(spec.channel_mult_dec),
                        channel_mult_emb=spec.channel_mult_emb, num_enc_layers=spec.num_enc_layers,
                        num_dec_layers_per_block=spec.num_dec_layers_per_block, mlp_multiplier=spec.mlp_multiplier)
    dae = DAE_D3(cfg)
    dae.load_state_dict(sd, strict=True)
    dae = dae.requires_grad_(False).train(False).to(dev)
    fmt = MS_MDCT_DualFormat(MS_MDCT_DualFormatConfig())
    for name, case in g["cases"].items():
        out = encode_latents(g["audio"].to(dev), g["clap"].to(dev), fmt, dae, EncodeLatentsConfig(**case["config"]))
        ref = case["latents"]
        assert out.dtype == torch.bfloat16 and tuple(out.shape) == tuple(ref.shape), name
        err = float((out.float().cpu() - ref.float()).norm() / ref.float().norm())
        assert err < 3e-2, (name, err)
