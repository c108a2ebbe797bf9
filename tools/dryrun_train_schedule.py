"""CPU dry run of the train-step launch schedule (no GPU, no kernels): every ops.* wrapper is replaced by a
shape-checking stub so the control flow, tensor shapes and descriptor tables of
dualdiffusion_b200/modules/unets/unet_train.py can be exercised in the build container."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dualdiffusion_b200 import ops, _lib as L  # noqa: E402
from oracle import unet_oracle as uo  # noqa: E402

bf = torch.bfloat16
calls = []


def E(*shape, dtype=bf):
    return torch.zeros(shape, dtype=dtype)


def weight_prep(w, gain=None, gain_host=1.0, normalize=False, fmt=0, qk_head_dim=0, out=None, pad_rows=0, row_stride=0,
                qkv_head_dim=0):
    O, I = w.shape[0], w.shape[1]
    taps = w[0, 0].numel()
    if pad_rows or row_stride:
        return E(max(O, pad_rows), row_stride or taps * I)
    return E(O, taps, I)


def mpconv(x, w, k, groups=1, *, epi=0, epi2=0, alpha=1.0, beta=0.0, clip=0.0, scale=None, scale2=None, residual=None,
           out=None, out2=None):
    B, H, W, Cin = x.shape
    Cout = w.shape[0]
    wk = w.numel() // Cout
    assert wk == k * k * (Cin // groups) or (k == 1 and wk == 64 and Cin == 64), (x.shape, w.shape, k, groups)
    assert (Cin // groups) % 32 == 0 and (Cout // groups) % 16 == 0, (Cin, Cout, groups)
    if scale is not None:
        assert scale.shape == (B, Cout)
    if residual is not None:
        assert residual.shape == (B, H, W, Cout)
    calls.append(("mpconv", tuple(x.shape), Cout, k, groups))
    o = E(B, H, W, Cout)
    return (o, E(B, H, W, Cout)) if epi2 else o


def wgrad(x, dy, k, groups=1, scale=1.0, out=None, accumulate=False):
    B, H, W, Cin = x.shape
    Cout = dy.shape[-1]
    assert dy.shape[:3] == x.shape[:3]
    assert (Cin // groups) % 32 == 0 and Cin % 8 == 0 and Cout % 8 == 0
    n = Cout * k * k * (Cin // groups)
    assert out is not None and out.numel() == n, (out.shape, Cout, k, Cin, groups)
    calls.append(("wgrad", tuple(x.shape), Cout, k, groups))
    return out


def weight_transpose(wp, cout, cin_g, taps, groups, out=None):
    assert wp.numel() == cout * cin_g * taps, (wp.shape, cout, cin_g, taps)
    return E(groups * cin_g, taps, cout // groups)


def same(*a, **k):
    return torch.zeros_like(a[0])


stubs = dict(
    weight_prep=weight_prep, mpconv=mpconv, mpconv_wgrad=wgrad, weight_transpose=weight_transpose,
    noise_embedding=lambda sigma, fr, ph, w, lab, lb, normalize=False, out=None: E(sigma.numel(), w.shape[0], dtype=torch.float32),
    emb_affine=lambda *a: None,
    stem_patches=lambda x, s, sd, lf, out=None: E(x.shape[0], x.shape[2], x.shape[3], 64),
    pixnorm_silu=lambda t, x_out=None, s_out=None: (torch.zeros_like(t), torch.zeros_like(t)),
    avgpool2=lambda x: E(x.shape[0], x.shape[1] // 2, x.shape[2] // 2, x.shape[3]),
    attention_train=lambda qk, v, sv, heads, hd=64: (torch.zeros_like(v), torch.zeros_like(v)),
    conv_out=lambda x, w, x_in, s, sd, x_ref=None, out=None: torch.zeros_like(x_in),
    weight_prep_bwd=lambda buf, n, rows: calls.append(("wbwd", n, rows)),
    silu_scale_bwd=lambda dy, coef, pre, scale, dscale, out=None: (_ for _ in ()).throw(AssertionError) if dscale.shape != scale.shape or dy.shape != pre.shape else torch.zeros_like(dy),
    pixnorm_silu_bwd=lambda g, ca, ds, t0: torch.zeros_like(t0) if g.shape == ds.shape == t0.shape else 1 / 0,
    enc_grad_combine=lambda dx0, down, dskip, x_prev, clip, shape: E(*shape) if (dskip is None or tuple(dskip.shape) == tuple(shape)) and tuple(dx0.shape) == ((shape[0], shape[1] // 2, shape[2] // 2, shape[3]) if down else tuple(shape)) else 1 / 0,
    attn_in_bwd=lambda g3, ca, dxv, dxs, x2, cqk, dc: torch.zeros_like(x2) if g3.shape == dxv.shape == dxs.shape == x2.shape else 1 / 0,
    attention_bwd=lambda qk, v, a, da, heads, hd=64: (torch.zeros_like(qk), torch.zeros_like(v)),
    emb_affine_bwd=lambda *a, **k: None,
    head_grad=lambda dD, s, sd, xr, cpad=32: E(dD.shape[0], dD.shape[2], dD.shape[3], cpad),
)


def cat_silu(a, b, wa, wb, up, need_cat=True):
    B, Ha, Wa, Ca = a.shape
    H, W = (Ha * 2, Wa * 2) if up else (Ha, Wa)
    Cb = 0 if b is None else b.shape[-1]
    return (E(B, H, W, Ca + Cb) if need_cat else None), E(B, H, W, Ca + Cb)


def cat_silu_bwd(d_xc, c1, d_s, xc, a_prev, clip, wa, wb, up, Ca, Cb):
    B, H, W, Ct = xc.shape
    assert Ct == Ca + Cb and d_xc.shape == xc.shape == d_s.shape, (d_xc.shape, xc.shape, d_s.shape)
    Ha, Wa = (H // 2, W // 2) if up else (H, W)
    assert tuple(a_prev.shape) == (B, Ha, Wa, Ca)
    return E(B, Ha, Wa, Ca), (E(B, H, W, Cb) if Cb else None)


def noise_embedding_bwd(sigma, fr, ph, w, lab, lb, demb, normalize, dweff=None):
    assert dweff is not None and dweff.numel() == w.numel()
    return dweff, torch.zeros_like(lab)


stubs.update(weight_prep_batched=lambda b, n, r: None, weight_transpose_batched=lambda b, n, t: None, cat_silu=cat_silu, cat_silu_bwd=cat_silu_bwd, noise_embedding_bwd=noise_embedding_bwd)
for k, v in stubs.items():
    assert hasattr(ops, k), k
    setattr(ops, k, v)
L.require_cuda = lambda *a: None

from dualdiffusion_b200.modules.unets import unet_edm2_b4 as U  # noqa: E402
from dualdiffusion_b200.modules.unets import unet_train as T  # noqa: E402

U._Plan.__init__.__globals__["torch"].cuda.Stream = lambda device=None: None


def run(spec, shape):
    cfg = U.UNetConfig(**{k: getattr(spec, k) for k in U.UNetConfig.__dataclass_fields__ if hasattr(spec, k)})
    net = U.UNet(cfg).train()
    plan = U._Plan(net)
    plan.device = torch.device("cpu")
    net._plan = plan
    net._get_plan = lambda: plan
    ts = T.get_train_state(net, plan)
    ts.refresh()
    B = shape[0]
    x = torch.randn(shape)
    sg = torch.ones(B)
    em = torch.randn(B, net.cemb)
    lf = torch.zeros(shape[2])
    calls.clear()
    d, saved = T.train_forward(net, plan, x, x, sg, em, lf, None)
    nf = len(calls)
    dl = T.train_backward(net, plan, saved, torch.randn(shape))
    print(f"{len(net.enc) + len(net.dec)} blocks: forward {nf} conv launches, backward {len(calls) - nf} conv/wgrad/wbwd launches;"
          f" buckets {[f'{(hi - lo) * 4 / 2**20:.0f}MB' for lo, hi in ts.bucket_ranges]}")
    n_params = sum(p.numel() for p in net.parameters())
    covered = sum(s.param.numel() for s in ts.slots.values()) + ts.n_gains
    extra = sum(p.numel() for n, p in net.named_parameters() if n.startswith(("emb_label", "logvar_linear")))
    assert covered + extra == n_params, (covered, extra, n_params)
    assert all(c[0] != "wbwd" or c[1] > 0 for c in calls)
    assert sum(1 for c in calls if c[0] == "wbwd") == len(ts.buckets)
    assert dl.shape == em.shape


if __name__ == "__main__":
    run(uo.small_spec(), (2, 4, 32, 48))
    run(uo.default_spec(), (1, 4, 32, 64))
    print("dry run ok")
