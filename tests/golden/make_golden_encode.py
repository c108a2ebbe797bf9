"""Regenerates tests/golden/encode_small.pt from the UNMODIFIED reference (CPU, build container only): the CUDA stage of the
dataset pre-encode process, `EncodeProcess.process` (src/dataset/processes/encode.py:278-360), driven with the reference's
own `MS_MDCT_DualFormat` and a reduced `DAE_D3` (synthetic weights) and with the audio embeddings supplied (the CLAP model is
not in this image; `has_audio_embeddings=True` takes the reference's own branch for that).

    python tests/golden/make_golden_encode.py
"""
import logging
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_shim, dae_oracle as do  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
CASES = {"offsets4_mirror_b2": dict(latents_batch_size=2, latents_num_time_offset_augmentations=4,
                                    latents_stereo_mirroring_augmentation=True),
         "offsets3_b1_plain_encode": dict(latents_batch_size=1, latents_num_time_offset_augmentations=3,
                                          latents_stereo_mirroring_augmentation=False, latents_tiled_encode=False)}


def main():
    ref_shim.install()
    from dataset.processes.encode import EncodeProcess, EncodeProcessConfig
    from modules.daes.dae_edm2_d3 import DAE_D3, DAE_D3_Config
    from modules.formats.ms_mdct_dual import MS_MDCT_DualFormat, MS_MDCT_DualFormatConfig
    torch.set_grad_enabled(False)
    fmt = MS_MDCT_DualFormat(MS_MDCT_DualFormatConfig())
    dspec = do.small_dae_spec()
    dsd = do.synth_dae_state_dict(dspec, seed=0)
    dae = DAE_D3(DAE_D3_Config(in_channels_emb=dspec.in_channels_emb, model_channels=dspec.model_channels,
                               channel_mult_enc=dspec.channel_mult_enc, channel_mult_dec=tuple(dspec.channel_mult_dec),
                               channel_mult_emb=dspec.channel_mult_emb, num_enc_layers=dspec.num_enc_layers,
                               num_dec_layers_per_block=dspec.num_dec_layers_per_block,
                               mlp_multiplier=dspec.mlp_multiplier)).eval()
    dae.load_state_dict(dsd, strict=True)
    g = torch.Generator().manual_seed(17)
    audio = 0.1 * torch.randn(2, 45000, generator=g)
    clap = torch.randn(3, dspec.in_channels_emb, generator=g)
    cases = {}
    for name, kw in CASES.items():
        pc = EncodeProcessConfig(**kw)
        proc = EncodeProcess(pc)
        # what start_process (:231-268) derives from the loaded pipeline, set directly
        proc.logger = logging.getLogger("golden")
        proc.device = torch.device("cpu")
        proc.format, proc.dae = fmt, dae
        proc.embedding = type("E", (), {"config": type("C", (), {"sample_crop_width": 1000})()})()
        proc.pipeline = type("P", (), {"get_latent_shape": staticmethod(lambda s: s),
                                       "get_mel_spec_shape": staticmethod(lambda raw_length=None: (1, 2, 0, 0))})()
        proc.format_config = fmt.config
        n = pc.latents_num_time_offset_augmentations
        proc.dae_encode_offset_padding = fmt.config.ms_frame_hop_length * n if n > 0 else 0
        proc.dae_encode_offsets = [i * fmt.config.ms_frame_hop_length for i in range(n)]
        proc.dae_batch_size = pc.latents_batch_size
        proc.dae_num_batches_per_sample = (n + proc.dae_batch_size - 1) // proc.dae_batch_size
        proc.use_tiled_encode = pc.latents_tiled_encode
        proc.tiled_max_chunk_size = pc.latents_tiled_max_chunk_size
        proc.tiled_overlap = pc.latents_tiled_overlap
        proc.dae_encode_formats = [fmt]
        res = proc.process(dict(safetensors_file_path="x.safetensors", file_path="x.flac", audio=audio.clone(),
                                sample_rate=fmt.config.sample_rate, prompt=None,
                                latents={"clap_audio_embeddings": clap.clone()}, latents_metadata={},
                                has_latents=False, has_audio_embeddings=True, has_text_embeddings=True))
        lat = res["latents"]["latents"]
        cases[name] = dict(config=kw, latents=lat.clone())
        print(name, tuple(lat.shape), lat.dtype, float(lat.float().std()))
    torch.save(dict(audio=audio, clap=clap, cases=cases,
                    weight_checksum=float(sum(v.double().abs().sum() for v in dsd.values()))),
               os.path.join(OUT, "encode_small.pt"))


if __name__ == "__main__":
    main()
