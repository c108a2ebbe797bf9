"""Drop-in for the reference's newer latent UNet lineage, `modules.unets.unet_edm2_b4_2.UNet` (SURVEY.md 8(f) N4;
/root/reference/src/modules/unets/unet_edm2_b4_2.py): same constructor, config fields, state_dict keys / shapes and call
signatures.  It is the b4 launch schedule (unet_edm2_b4.py here) with the lineage's differences switched on:

  * attention: ONE fused q|k|v projection whose input carries a single embedding gain (`attn_qkv`, `emb_linear_qkv`,
    `emb_gain_qkv`, :112-117, :146-157), no gain / activation on the attention output -- `dd_attention_qkv` reads the
    thirds of the projection in place, the weight rows are de-interleaved once by `dd_weight_prep(DD_WPERM_QKV)`;
  * noise level shifted before the Fourier embeddings, c_noise = (ln sigma - offset) / 4, and a bandwidth factor in the
    embedding frequencies (:181, :237-238, :258-259) -- the kernels take sigma, so sigma * exp(-offset) is passed;
  * 8-channel latents: the stem is a K = 128 patch GEMM (9 * (8 + 2) = 90 columns), `dd_stem_patches_cols`;
  * `emb_linear` has its own group count (`emb_linear_groups`), 3 layers per level, mlp_multiplier 1 (config :44-69).

Eval-mode forward (sampler) only: the train step of this lineage is not built (raises NotImplementedError)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Sequence

from .unet import DualDiffusionUNetConfig
from .unet_edm2_b4 import UNet as _UNetB4


@dataclass
class UNetConfig(DualDiffusionUNetConfig):
    """unet_edm2_b4_2.py:44-69 (field for field, same defaults)."""
    in_channels: int = 8
    out_channels: int = 8
    in_channels_emb: int = 1024
    sigma_max: float = 400.0
    sigma_min: float = 0.004
    sigma_data: float = 1.0
    mp_fourier_ln_sigma_offset: float = 0.5
    mp_fourier_bandwidth: float = 1.4
    model_channels: int = 256
    logvar_channels: int = 192
    channel_mult: Sequence[int] = (2, 2, 3, 4, 5)
    channel_mult_noise: Optional[int] = None
    channel_mult_emb: Optional[int] = None
    channels_per_head: int = 64
    num_layers_per_block: int = 3
    label_balance: float = 0.5
    concat_balance: float = 0.5
    res_balance: float = 0.3
    attn_balance: float = 0.3
    attn_levels: Sequence[int] = (2, 3, 4)
    mlp_multiplier: int = 1
    mlp_groups: int = 8
    emb_linear_groups: int = 1


class UNet(_UNetB4):
    config_class = UNetConfig
    fused_qkv = True
    stem_cols = 128

    def __init__(self, config: UNetConfig) -> None:
        super().__init__(config)
