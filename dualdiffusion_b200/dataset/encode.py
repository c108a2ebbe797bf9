"""Dataset pre-encode compute: raw audio -> augmented variants -> live mel-STFT
(`MS_MDCT_DualFormat.raw_to_mel_spec`, row A15) -> `DAE_D3.tiled_encode` (row A16) -> bf16 latents, i.e. the CUDA stage of
the reference's `EncodeProcess.process` (src/dataset/processes/encode.py:303-352) without its file / queue / CLAP plumbing
(EncodeLoad, EncodeSave and the embedding model are control plane and stay the reference's).

Every item is independent: across GPUs the file list is sharded with `dualdiffusion_b200.replicas` (replicas only, no
collective -- SURVEY 8(e)).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Sequence

import torch

from ..modules.mp_tools import normalize

Tensor = torch.Tensor


@dataclass
class EncodeLatentsConfig:
    """The fields of `EncodeProcessConfig` (encode.py:48-63) that shape the compute, same names and defaults."""
    latents_batch_size: int = 1
    latents_num_time_offset_augmentations: int = 8
    latents_stereo_mirroring_augmentation: bool = True
    latents_tiled_encode: bool = True
    latents_tiled_max_chunk_size: int = 6144
    latents_tiled_overlap: int = 256


def augmented_audio(audio: Tensor, fmt, cfg: EncodeLatentsConfig) -> Tensor:
    """encode.py:303-315: sub-latent-pixel time offsets (multiples of the mel hop) cropped to a common width, each optionally
    followed by its stereo-swapped copy.  audio (channels, length) -> (variants, channels, crop_width).  Index-only."""
    n_off = cfg.latents_num_time_offset_augmentations
    hop = fmt.config.ms_frame_hop_length
    padding = hop * n_off if n_off > 0 else 0                                   # :258
    offsets = [i * hop for i in range(n_off)]                                   # :259
    crop_width = fmt.get_raw_crop_width(audio.shape[-1] - padding)              # :304
    variants: List[Tensor] = []
    for off in offsets:
        v = audio[:, off:off + crop_width].unsqueeze(0)
        variants.append(v)
        if cfg.latents_stereo_mirroring_augmentation:
            variants.append(torch.flip(v, dims=(1,)))
    return torch.cat(variants, dim=0)


@torch.inference_mode()
def encode_latents(audio: Tensor, clap_audio_embeddings: Tensor, fmt, dae, cfg: EncodeLatentsConfig,
                   extra_formats: Sequence = ()) -> Tensor:
    """encode.py:303-352 for one file: audio (channels, length) at the format's sample rate and the file's CLAP audio
    embeddings (chunks, emb) -> latents (variations, 2*latent_channels, H/r, W/r) bf16.  `extra_formats` are the pitch-shifted
    formats of :265-268 (empty by default, as in the reference config).

    Two quirks of the reference are kept because the stored datasets depend on them: the spectrogram loop takes
    ceil(num_offsets / batch) batches of the variant list (:318-322), so with stereo mirroring only the first num_offsets
    of the 2*num_offsets variants are encoded; and the latent loop drops a trailing partial batch (:343)."""
    bsz = cfg.latents_batch_size
    n_off = cfg.latents_num_time_offset_augmentations
    batches_per_sample = (n_off + bsz - 1) // bsz                               # :261
    variants = augmented_audio(audio, fmt, cfg)
    mels: List[Tensor] = []
    for f in (fmt, *extra_formats):                                             # :318-323
        for b in range(batches_per_sample):
            batch = variants[b * bsz:(b + 1) * bsz]
            mels.append(f.raw_to_mel_spec(batch).to(torch.bfloat16))
    mel = torch.cat(mels, dim=0)
    emb = clap_audio_embeddings.mean(dim=0, keepdim=True)                       # :337-339
    emb = normalize(emb.to(device=mel.device, dtype=torch.float32))
    dae_emb = dae.get_embeddings(emb.to(dtype=getattr(dae, "dtype", torch.float32)))
    out: List[Tensor] = []
    for b in range(mel.shape[0] // bsz):                                        # :343-352
        batch = mel[b * bsz:(b + 1) * bsz]
        if cfg.latents_tiled_encode:
            lat = dae.tiled_encode(batch, dae_emb, max_chunk=cfg.latents_tiled_max_chunk_size,
                                   overlap=cfg.latents_tiled_overlap)
        else:
            lat = dae.encode(batch, dae_emb)
        out.append(lat)
    latents = torch.cat(out, dim=0).to(dtype=torch.bfloat16)
    assert latents.ndim == 4
    return latents
