"""Golden vector for SURVEY row A9 (axis attention of the legacy ddec UNets): the UNMODIFIED reference Block of
src/modules/unets/old/unet_edm2_ddec_mdct_b3.py runs forward on CPU; a forward hook on `attn_qkv` captures the qkv tensor
and a forward pre-hook on `attn_proj` captures mp_silu(y) -- i.e. exactly what lines 144-163 compute between the two
convolutions (permute to (b, z, w, c, h), fold (b, z, w) into the batch, cosine-normalise, SDPA over h, reshape / permute
back).  Writes tests/golden/axis_attention_b3.pt.      python tests/golden/make_golden_axis_attention.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402

ref_shim.install()
from modules.unets.old.unet_edm2_ddec_mdct_b3 import Block  # noqa: E402

torch.manual_seed(0)
out = {}
for tag, (b, c, z, h, w, cph) in {"a": (2, 64, 2, 12, 5, 32), "b": (1, 128, 2, 7, 3, 64)}.items():
    blk = Block(level=0, in_channels=c, out_channels=c, emb_channels=16, num_freqs=h, flavor="dec", use_attention=True,
                channels_per_head=cph).eval()
    with torch.no_grad():
        blk.emb_gain.fill_(0.5)
    cap = {}
    blk.attn_qkv.register_forward_hook(lambda m, i, o: cap.__setitem__("qkv", o.detach().clone()))
    blk.attn_proj.register_forward_pre_hook(lambda m, i: cap.__setitem__("y_silu", i[0].detach().clone()))
    x = torch.randn(b, c, z, h, w)
    emb = torch.randn(b, 16, 1, 1, 1)
    with torch.no_grad():
        blk(x, emb)
    out[tag] = dict(qkv=cap["qkv"].contiguous(), y_silu=cap["y_silu"].contiguous(), heads=blk.num_heads)
    print(tag, tuple(cap["qkv"].shape), "heads", blk.num_heads, "std", float(cap["y_silu"].std()))
torch.save(out, os.path.join(ROOT, "tests", "golden", "axis_attention_b3.pt"))
