"""GPU parity tests (-m gpu) of the UNet train step: every backward kernel against
PyTorch autograd of the CPU oracle's op (oracle/unet_oracle.py, fp32) on the same bf16-rounded inputs, and the
assembled forward+backward against (a) autograd through the oracle and (b) gradient statistics of the unmodified
reference (tests/golden/unet_small_train.pt).  Everything goes through the C ABI.

Tolerances: gradients are products of bf16-stored activations and bf16-stored upstream gradients accumulated in fp32;
one op is held to BF16_GRAD_OP = 6e-3 relative L2 (two bf16 roundings), fp32 reductions to 2e-5, the whole-network
parameter gradients to BF16_GRAD_NET = 6e-2 relative L2 per tensor (measured 1.2e-2 typical, 2.9e-2 worst; printed by
tools/dev_check_backward.py), with a cosine similarity above 0.998 (measured >= 0.9996).
"""
import math

import pytest
import torch
import torch.nn.functional as F

from conftest import load_golden, rel_err
from oracle import unet_oracle as uo

pytestmark = pytest.mark.gpu

BF16_GRAD_OP = 6e-3
BF16_GRAD_NET = 6e-2
FP32 = 2e-5


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


def nhwc_bf16(x, dev):
    return x.permute(0, 2, 3, 1).contiguous().to(device=dev, dtype=torch.bfloat16)


def to_nchw(y):
    return y.float().permute(0, 3, 1, 2).cpu()


def bf16_round(x):
    return x.to(torch.bfloat16).float()


def coeffs(t):
    n = math.hypot(1 - t, t)
    return (1 - t) / n, t / n


# ------------------------------------------------------------------------------------------
# MPConv backward: wgrad (tcgen05, pixels as K) and dgrad (forward kernel on transposed weights)
# ------------------------------------------------------------------------------------------
WGRAD_CASES = [
    # B, H, W, Cin, Cout, k, groups
    (1, 8, 16, 64, 64, 1, 1),        # one pixel tile
    (2, 7, 13, 96, 64, 1, 1),        # ragged pixel count, Cin not a multiple of 64
    (2, 32, 48, 64, 256, 1, 1),      # stem: patches (K padded to 64) x dY
    (1, 16, 24, 256, 512, 1, 1),     # 4 input-channel chunks per CTA
    (2, 4, 43, 1280, 1280, 1, 1),    # attn_proj at the coarsest attention level
    (1, 4, 10, 1280, 2560, 1, 1),    # attn_qk
    (1, 16, 24, 256, 512, 3, 8),     # grouped 3x3, cin_g 32 -> cout_g 64
    (2, 5, 43, 512, 256, 3, 8),      # cin_g 64 -> cout_g 32, odd sizes (ht = 6)
    (2, 2, 43, 1280, 2560, 3, 8),    # coarsest level, cin_g 160, cout_g 320
    (1, 32, 40, 256, 32, 3, 1),      # head: dF padded to 32 channels
    (2, 17, 20, 768, 512, 3, 8),     # cin_g 96
    (1, 1, 1, 64, 32, 3, 1),         # single pixel
]


@pytest.mark.parametrize("B,H,W,Cin,Cout,k,g", WGRAD_CASES)
def test_wgrad_vs_autograd(dev, B, H, W, Cin, Cout, k, g):
    from dualdiffusion_b200 import ops
    gen = torch.Generator().manual_seed(B * 1000 + H * 100 + Cin + Cout + k)
    x = bf16_round(torch.randn(B, Cin, H, W, generator=gen))
    dy = bf16_round(torch.randn(B, Cout, H, W, generator=gen))
    w = torch.zeros(Cout, Cin // g, k, k, requires_grad=True)
    F.conv2d(x, w, padding=k // 2, groups=g).backward(dy)
    ref = w.grad.permute(0, 2, 3, 1).reshape(Cout, k * k, Cin // g)          # [co][tap][ci]
    got = ops.mpconv_wgrad(nhwc_bf16(x, dev), nhwc_bf16(dy, dev), k, g, scale=0.5)
    assert rel_err(got, 0.5 * ref) < FP32 * 5
    # accumulate adds into the buffer
    ops.mpconv_wgrad(nhwc_bf16(x, dev), nhwc_bf16(dy, dev), k, g, scale=0.5, out=got, accumulate=True)
    assert rel_err(got, ref) < FP32 * 5


@pytest.mark.parametrize("B,H,W,Cin,Cout,k,g", [(2, 8, 20, 256, 512, 3, 8), (1, 16, 24, 512, 256, 3, 8),
                                                 (2, 4, 11, 768, 768, 1, 1), (1, 32, 24, 256, 32, 3, 1),
                                                 (2, 2, 43, 1280, 2560, 3, 8)])
def test_dgrad_vs_autograd(dev, B, H, W, Cin, Cout, k, g):
    from dualdiffusion_b200 import ops
    gen = torch.Generator().manual_seed(Cin + Cout + k)
    w = torch.randn(Cout, Cin // g, k, k, generator=gen)
    dy = bf16_round(torch.randn(B, Cout, H, W, generator=gen))
    wp = ops.weight_prep(w.to(dev))
    wt = ops.weight_transpose(wp, Cout, Cin // g, k * k, g)
    dx = ops.mpconv(nhwc_bf16(dy, dev), wt, k, g)
    x = torch.zeros(B, Cin, H, W, requires_grad=True)
    F.conv2d(x, bf16_round(uo.mp_weight(w)), padding=k // 2, groups=g).backward(dy)
    assert rel_err(to_nchw(dx), x.grad) < 4e-3


def test_weight_prep_bwd_vs_autograd(dev):
    from dualdiffusion_b200 import ops, _lib as L
    gen = torch.Generator().manual_seed(31)
    cases = [dict(shape=(512, 32, 3, 3), gain=None, normalize=True, perm=0),
             dict(shape=(256, 6, 3, 3), gain=None, normalize=True, perm=0, row_stride=64),
             dict(shape=(384, 96, 1, 1), gain=0.7, normalize=True, perm=0),
             dict(shape=(256, 128, 1, 1), gain=None, normalize=True, perm=L.WPERM_QK, head_dim=64),
             dict(shape=(4, 256, 3, 3), gain=0.5, normalize=False, perm=0)]
    entries, refs = [], []
    dgains = torch.zeros(len(cases), device=dev)
    for i, c in enumerate(cases):
        O, I, kh, kw = c["shape"]
        taps = kh * kw
        w = torch.randn(c["shape"], generator=gen, requires_grad=True)
        gain = None if c["gain"] is None else torch.tensor(c["gain"], requires_grad=True)
        rs = c.get("row_stride", I * taps)
        G = torch.randn(O, rs, generator=gen)                               # dL/dW_eff in [o][tap][i] (+ padding) layout
        w_eff = uo.mp_weight(w, 1.0 if gain is None else gain, c["normalize"])            # OIHW
        eff_rows = w_eff.permute(0, 2, 3, 1).reshape(O, taps * I)
        if c["perm"] == L.WPERM_QK:      # prepared rows are (j, head, d); parameter rows are (head, d, j)
            heads = O // (2 * c["head_dim"])
            eff_rows = eff_rows.view(heads, c["head_dim"], 2, taps * I).permute(2, 0, 1, 3).reshape(O, taps * I)
        (eff_rows * G[:, :taps * I]).sum().backward()
        dw = torch.empty(c["shape"], device=dev)
        entries.append(dict(w=w.detach().to(dev), dweff=G.to(dev), dw=dw,
                            gain=None if gain is None else gain.detach().to(dev).view(1),
                            dgain=None if gain is None else dgains[i:i + 1], O=O, I_g=I, taps=taps,
                            normalize=c["normalize"], perm=c["perm"], head_dim=c.get("head_dim", 0), row_stride=rs))
        refs.append((w.grad, None if gain is None else gain.grad, dw))
    buf, rows = ops.make_wbwd_descs(entries, dev)
    ops.weight_prep_bwd(buf, len(entries), rows)
    for i, (gw, gg, dw) in enumerate(refs):
        assert rel_err(dw, gw) < FP32 * 5, i
        if gg is not None:
            assert abs(float(dgains[i]) - float(gg)) < 1e-4 * (1 + abs(float(gg))), i
    # weight-norm makes the gradient orthogonal to the weight, per output channel
    w0, dw0 = entries[0]["w"].flatten(1), refs[0][2].flatten(1)
    cos = (w0 * dw0).sum(1) / (w0.norm(dim=1) * dw0.norm(dim=1))
    assert cos.abs().max().item() < 1e-3


# ------------------------------------------------------------------------------------------
# block glue
# ------------------------------------------------------------------------------------------
def test_silu_scale_bwd_vs_autograd(dev):
    from dualdiffusion_b200 import ops
    gen = torch.Generator().manual_seed(37)
    B, H, W, C = 2, 9, 15, 512
    pre = bf16_round(torch.randn(B, C, H, W, generator=gen) * 2).requires_grad_(True)
    sc = (torch.randn(B, C, generator=gen) * 0.3 + 1).requires_grad_(True)
    dy = bf16_round(torch.randn(B, C, H, W, generator=gen))
    (uo.mp_silu(pre * sc[:, :, None, None]) * dy * 0.4).sum().backward()
    dsc = torch.zeros(B, C, device=dev)
    dpre = ops.silu_scale_bwd(nhwc_bf16(dy, dev), 0.4, nhwc_bf16(pre.detach(), dev), sc.detach().to(dev), dsc)
    assert rel_err(to_nchw(dpre), pre.grad) < BF16_GRAD_OP
    assert rel_err(dsc, sc.grad) < 1e-4


def test_pixnorm_silu_bwd_vs_autograd(dev):
    from dualdiffusion_b200 import ops
    gen = torch.Generator().manual_seed(41)
    B, H, W, C = 2, 6, 10, 768
    t0 = bf16_round(torch.randn(B, C, H, W, generator=gen) * 3).requires_grad_(True)
    g = bf16_round(torch.randn(B, C, H, W, generator=gen))
    ds = bf16_round(torch.randn(B, C, H, W, generator=gen))
    ca, _ = coeffs(0.3)
    xn = uo.normalize(t0, dim=1)
    ((ca * xn * g).sum() + (uo.mp_silu(xn) * ds).sum()).backward()
    dt0 = ops.pixnorm_silu_bwd(nhwc_bf16(g, dev), ca, nhwc_bf16(ds, dev), nhwc_bf16(t0.detach(), dev))
    assert rel_err(to_nchw(dt0), t0.grad) < BF16_GRAD_OP


@pytest.mark.parametrize("mode", ["cat", "up", "plain"])
def test_cat_silu_bwd_vs_autograd(dev, mode):
    from dualdiffusion_b200 import ops
    gen = torch.Generator().manual_seed(43)
    B, H, W, Ca, Cb = 2, 6, 10, 512, (256 if mode == "cat" else 0)
    up = mode == "up"
    Ha, Wa = (H // 2, W // 2) if up else (H, W)
    a_pre = torch.randn(B, Ca, Ha, Wa, generator=gen) * 1.2
    a_pre.view(-1)[::7] *= 3.0                                             # some values beyond the clip
    clip = 2.0
    a_pre = bf16_round(a_pre)
    a_pre[a_pre.abs() == clip] = 0.5     # |x| == clip exactly: clamp passes the gradient, the stored-output mask cannot tell
    a_pre.requires_grad_(True)           # (clip_act = 256 is never hit exactly by O(1) activations)
    a = a_pre.clip(-clip, clip)
    wa, wb = uo.mp_cat_weights(Ca, Cb, 0.5) if Cb else (1.0, 0.0)
    parts = [wa * (uo.resample_2d(a, "up") if up else a)]
    b = None
    if Cb:
        b = bf16_round(torch.randn(B, Cb, H, W, generator=gen)).requires_grad_(True)
        parts.append(wb * b)
    xc = torch.cat(parts, 1)
    d_xc = bf16_round(torch.randn_like(xc))
    d_s = bf16_round(torch.randn_like(xc))
    c1 = 0.9
    ((c1 * xc * d_xc).sum() + (uo.mp_silu(xc) * d_s).sum()).backward()
    xcd, _ = ops.cat_silu(nhwc_bf16(a.detach(), dev), None if b is None else nhwc_bf16(b.detach(), dev), wa, wb, up)
    da, db = ops.cat_silu_bwd(nhwc_bf16(d_xc, dev), c1, nhwc_bf16(d_s, dev), xcd, nhwc_bf16(a.detach(), dev), clip, wa, wb,
                              up, Ca, Cb)
    # the kernel masks with |a| < clip on the (clipped) stored activation: identical to clamp's gradient mask
    assert rel_err(to_nchw(da), a_pre.grad) < BF16_GRAD_OP
    if Cb:
        assert rel_err(to_nchw(db), b.grad) < BF16_GRAD_OP


@pytest.mark.parametrize("down", [False, True])
def test_enc_grad_combine_vs_autograd(dev, down):
    from dualdiffusion_b200 import ops
    gen = torch.Generator().manual_seed(47)
    B, H, W, C = 2, 8, 12, 256
    clip = 1.5
    xp = bf16_round(torch.randn(B, C, H, W, generator=gen))
    xp[xp.abs() == clip] = 0.5           # see test_cat_silu_bwd_vs_autograd
    xp.requires_grad_(True)
    x = xp.clip(-clip, clip)
    nxt = uo.resample_2d(x, "down") if down else x
    dx0 = bf16_round(torch.randn_like(nxt))
    dskip = bf16_round(torch.randn(B, C, H, W, generator=gen))
    ((nxt * dx0).sum() + (x * dskip).sum()).backward()
    out = ops.enc_grad_combine(nhwc_bf16(dx0, dev), down, nhwc_bf16(dskip, dev), nhwc_bf16(x.detach(), dev), clip,
                               (B, H, W, C))
    assert rel_err(to_nchw(out), xp.grad) < BF16_GRAD_OP
    out = ops.enc_grad_combine(nhwc_bf16(dx0, dev), down, None, None, 0.0, (B, H, W, C))
    ref = F.interpolate(dx0, scale_factor=2) * 0.25 if down else dx0
    assert rel_err(to_nchw(out), ref) < BF16_GRAD_OP


def test_attn_in_bwd_vs_autograd(dev):
    from dualdiffusion_b200 import ops
    gen = torch.Generator().manual_seed(53)
    B, H, W, C = 2, 4, 11, 768
    x2 = bf16_round(torch.randn(B, C, H, W, generator=gen)).requires_grad_(True)
    cqk = (torch.randn(B, C, generator=gen) * 0.3 + 1).requires_grad_(True)
    g3, dxv, dxs = (bf16_round(torch.randn(B, C, H, W, generator=gen)) for _ in range(3))
    ca, _ = coeffs(0.3)
    ((ca * x2 * g3).sum() + (x2 * dxv).sum() + (x2 * cqk[:, :, None, None] * dxs).sum()).backward()
    dc = torch.zeros(B, C, device=dev)
    dx2 = ops.attn_in_bwd(nhwc_bf16(g3, dev), ca, nhwc_bf16(dxv, dev), nhwc_bf16(dxs, dev), nhwc_bf16(x2.detach(), dev),
                          cqk.detach().to(dev), dc)
    assert rel_err(to_nchw(dx2), x2.grad) < BF16_GRAD_OP
    assert rel_err(dc, cqk.grad) < 1e-4


@pytest.mark.parametrize("B,H,W,heads", [(2, 4, 86, 4), (1, 2, 43, 20), (1, 8, 8, 2), (1, 1, 1, 1), (2, 3, 37, 3)])
def test_attention_bwd_vs_autograd(dev, B, H, W, heads):
    from dualdiffusion_b200 import ops
    gen = torch.Generator().manual_seed(59 + H * W)
    C = heads * 64
    qk = bf16_round(torch.randn(B, 2 * C, H, W, generator=gen) * 2).requires_grad_(True)      # (head, c, j) channels
    v = bf16_round(torch.randn(B, C, H, W, generator=gen)).requires_grad_(True)
    sv = torch.randn(B, C, generator=gen) * 0.2 + 1
    da = bf16_round(torch.randn(B, C, H, W, generator=gen))
    a = uo.attention_core(qk, v, heads)
    (a * da).sum().backward()

    def split(t):       # oracle channel order (head, c, j) -> kernel order [q | k]
        return t.view(B, heads, 64, 2, H, W).permute(0, 3, 1, 2, 4, 5).reshape(B, 2 * C, H, W)
    qkd, vd = nhwc_bf16(split(qk.detach()), dev), nhwc_bf16(v.detach(), dev)
    y, a_raw = ops.attention_train(qkd, vd, sv.to(dev), heads)
    assert rel_err(to_nchw(a_raw), a) < 1.2e-2
    assert rel_err(to_nchw(y), uo.mp_silu(a.detach() * sv[:, :, None, None])) < 1.2e-2
    dqk, dv = ops.attention_bwd(qkd, vd, a_raw, nhwc_bf16(da, dev), heads)
    if H * W == 1:      # one key: dq = dk = 0 exactly; the kernel's D = <da, bf16(a)> leaves bf16-rounding residue
        assert to_nchw(dqk).abs().max().item() < 5e-3
    else:
        assert rel_err(to_nchw(dqk), split(qk.grad)) < 1.5e-2
    assert rel_err(to_nchw(dv), v.grad) < 1.5e-2


def test_embedding_heads_bwd_vs_autograd(dev):
    from dualdiffusion_b200 import ops
    spec = uo.small_spec()
    sd = uo.synth_state_dict(spec, seed=0)
    gen = torch.Generator().manual_seed(61)
    B = 3
    sigma = torch.tensor([2.0, 0.05, 150.0])
    clap = torch.randn(B, spec.in_channels_emb, generator=gen)
    mask = torch.tensor([True, False, True])
    # get_embeddings
    wl = sd["emb_label.weight"].clone().requires_grad_(True)
    wu = sd["emb_label_unconditional.weight"].clone().requires_grad_(True)
    dout = torch.randn(B, spec.cemb, generator=gen)
    u = uo.mp_conv(torch.ones(1), wu, training=True)
    c = uo.mp_conv(uo.normalize(clap), wl, training=True)
    lab = uo.mp_sum(u, c, mask.unsqueeze(1).float())
    (lab * dout).sum().backward()
    dwl_eff, dwu_eff = ops.label_embedding_bwd(clap.to(dev), mask.float().to(dev), dout.to(dev))
    dwl, dwu = torch.empty_like(wl, device=dev), torch.empty_like(wu, device=dev)
    entries = [dict(w=wl.detach().to(dev), dweff=dwl_eff, dw=dwl, O=wl.shape[0], I_g=wl.shape[1], taps=1, normalize=True),
               dict(w=wu.detach().to(dev), dweff=dwu_eff, dw=dwu, O=wu.shape[0], I_g=1, taps=1, normalize=True)]
    buf, rows = ops.make_wbwd_descs(entries, dev)      # descriptors hold raw pointers: `entries` must outlive the launch
    ops.weight_prep_bwd(buf, 2, rows)
    assert rel_err(dwl, wl.grad) < 1e-4
    assert (dwu.cpu() - wu.grad).abs().max().item() < 1e-5       # normalised scalar rows: gradient ~ eps
    # noise embedding
    wn = sd["emb_noise.weight"].clone()
    lab_in = lab.detach().clone().requires_grad_(True)
    weff = uo.mp_weight(wn, training=True).requires_grad_(True)
    e = uo.mp_fourier(sigma.log() / 4, sd["emb_fourier.freqs"], sd["emb_fourier.phases"]) @ weff.t()
    emb = uo.mp_silu(uo.mp_sum(e, lab_in, spec.label_balance))
    demb = torch.randn(B, spec.cemb, generator=gen)
    (emb * demb).sum().backward()
    dweff, dlabel = ops.noise_embedding_bwd(sigma.to(dev), sd["emb_fourier.freqs"].to(dev), sd["emb_fourier.phases"].to(dev),
                                            wn.to(dev), lab_in.detach().to(dev), spec.label_balance, demb.to(dev), True)
    assert rel_err(dweff, weff.grad) < 1e-4 and rel_err(dlabel, lab_in.grad) < 1e-4
    # emb_linear (grouped) backward
    O, I, G = 512, spec.cemb // 8, 8
    w = uo.normalize(torch.randn(O, I, generator=gen))
    gain = torch.tensor(0.5)
    embv = emb.detach().clone().requires_grad_(True)
    weff2 = uo.mp_weight(w, gain, training=True).requires_grad_(True)
    cc = F.conv2d(embv[:, :, None, None], weff2[:, :, None, None], groups=G)[:, :, 0, 0] + 1
    dc = torch.randn(B, O, generator=gen)
    (cc * dc).sum().backward()
    dweff_d = torch.empty(O, I, device=dev)
    rowscale = torch.empty(O, device=dev)
    demb_d = torch.zeros(B, spec.cemb, device=dev)
    entries = [dict(w=w.to(dev), gain=gain.to(dev).view(1), dout=dc.to(dev), dweff=dweff_d, rowscale=rowscale, groups=G,
                    normalize=True)]
    descs, max_o, max_cols = ops.make_affine_bwd_descs(entries, dev)
    ops.emb_affine_bwd(descs, 1, max_o, max_cols, embv.detach().to(dev), demb_d)
    torch.cuda.synchronize()
    assert rel_err(dweff_d, weff2.grad) < 1e-4 and rel_err(demb_d, embv.grad) < 1e-4
    # logvar head
    wlv = sd["logvar_linear.weight"].clone().requires_grad_(True)
    lv = uo.sigma_loss_logvar({**sd, "logvar_linear.weight": wlv}, sigma)
    dlv = torch.randn(B, generator=gen)
    (lv.flatten() * dlv).sum().backward()
    dw = ops.sigma_logvar_bwd(sigma.to(dev), sd["logvar_fourier.freqs"].to(dev), sd["logvar_fourier.phases"].to(dev),
                              dlv.to(dev))
    assert rel_err(dw, wlv.grad) < 1e-4


def test_head_grad_vs_autograd(dev):
    from dualdiffusion_b200 import ops
    gen = torch.Generator().manual_seed(67)
    B, H, W = 2, 8, 12
    Fx = torch.randn(B, 4, H, W, generator=gen, requires_grad=True)
    x_in = torch.randn(B, 4, H, W, generator=gen)
    sigma = torch.tensor([2.0, 0.3])
    s = sigma.view(-1, 1, 1, 1)
    x_ref = torch.rand(B, 5, H, W, generator=gen)
    dD = torch.randn(B, 4, H, W, generator=gen)
    for xr in (None, x_ref):
        Fx.grad = None
        d = x_in / (1 + s ** 2) + s / (1 + s ** 2).sqrt() * Fx
        if xr is not None:
            d = uo.mp_sum(xr[:, :-1], d, xr[:, -1:])
        (d * dD).sum().backward()
        got = ops.head_grad(dD.to(dev), sigma.to(dev), 1.0, None if xr is None else xr.to(dev), 32)
        assert got.shape == (B, H, W, 32)
        assert rel_err(to_nchw(got)[:, :4], Fx.grad) < 4e-3
        assert got[..., 4:].abs().max().item() == 0


# ------------------------------------------------------------------------------------------
# assembled train step
# ------------------------------------------------------------------------------------------
def make_train_unet(spec, sd, dev):
    from dualdiffusion_b200.modules.unets.unet_edm2_b4 import UNet, UNetConfig
    cfg = UNetConfig(**{k: getattr(spec, k) for k in UNetConfig.__dataclass_fields__ if hasattr(spec, k)})
    net = UNet(cfg)
    net.load_state_dict(sd, strict=True)
    return net.to(dev).train()


def product_loss(net, spec, samples, noise, sigma, clap, mask, dev):
    """unet_trainer.py:239-280 written against the drop-in module (what the reference trainer executes)."""
    emb = net.get_embeddings(clap.to(dev), mask.to(dev))
    sig = sigma.to(dev).view(-1, 1, 1, 1)
    x = samples.to(dev)
    denoised = net(x + noise.to(dev) * sig, sigma.to(dev), None, emb)
    w = (sig ** 2 + spec.sigma_data ** 2) / (sig * spec.sigma_data) ** 2
    wl = (F.mse_loss(denoised, x, reduction="none") * w).mean(dim=(1, 2, 3))
    logvar = net.get_sigma_loss_logvar(sigma.to(dev))
    return (wl / logvar.exp() + logvar).mean(), denoised


def test_train_step_vs_golden_reference_and_oracle(dev):
    spec = uo.small_spec()
    sd = uo.synth_state_dict(spec, seed=0)
    g = load_golden("unet_small_train.pt")
    net = make_train_unet(spec, sd, dev)
    loss, denoised = product_loss(net, spec, g["samples"], g["noise"], g["sigma"], g["clap"], g["mask"], dev)
    assert rel_err(denoised, g["denoised"]) < 3e-2
    assert abs(float(loss) - float(g["loss"])) < 2e-2 * abs(float(g["loss"]))
    loss.backward()
    # (a) reference gradient statistics (norm, projection on a seeded probe) for every parameter
    worst = 0.0
    for name, p in net.named_parameters():
        assert p.grad is not None, name
        n_ref, dot_ref = g["grad_stats"][name]
        gr = p.grad.float().cpu()
        if gr.ndim == 0:        # scalar gains: small sums of large cancelling terms (measured: up to 12 % at |g| = 1e-3)
            assert abs(float(gr) - dot_ref / float(uo.grad_probe(name, gr.shape))) < 0.15 * n_ref + 2e-4, name
            continue
        assert abs(float(gr.norm()) - n_ref) < BF16_GRAD_NET * n_ref + 1e-7, (name, float(gr.norm()), n_ref)
        dot = float((gr * uo.grad_probe(name, gr.shape)).sum())
        assert abs(dot - dot_ref) < 3 * BF16_GRAD_NET * n_ref + 1e-7, (name, dot, dot_ref)   # <err, unit-variance probe> ~ |err|
        worst = max(worst, abs(float(gr.norm()) - n_ref) / (n_ref + 1e-12))
    for name, gref in g["small_grads"].items():
        got = dict(net.named_parameters())[name].grad
        if gref.ndim:
            assert rel_err(got, gref) < BF16_GRAD_NET or (got.cpu() - gref).abs().max().item() < 1e-6, name
    # (b) full gradients against autograd through the oracle
    sdg = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "fourier" not in k else v) for k, v in sd.items()}
    uo.train_loss(sdg, spec, g["samples"], g["noise"], g["sigma"], g["clap"], g["mask"]).backward()
    for name, p in net.named_parameters():
        ref = sdg[name].grad
        got = p.grad.float().cpu()
        if ref.ndim == 0:
            assert abs(float(got) - float(ref)) < 0.15 * abs(float(ref)) + 2e-4, name
            continue
        if ref.norm() < 1e-9:
            assert got.norm() < 1e-6, name
            continue
        assert rel_err(got, ref) < BF16_GRAD_NET, (name, rel_err(got, ref))
        cos = float((got * ref).sum() / (got.norm() * ref.norm()))
        assert cos > 0.998, (name, cos)
    # weight-norm inside the forward projects conv-weight gradients orthogonal to the weights
    w, gw = net.enc["block0_layer0"].conv_res0.weight, net.enc["block0_layer0"].conv_res0.weight.grad
    cosr = (w * gw).flatten(1).sum(1) / (w.flatten(1).norm(dim=1) * gw.flatten(1).norm(dim=1))
    assert cosr.abs().max().item() < 2e-3


def test_train_step_is_linear_in_upstream_gradient_and_accumulates(dev):
    """Size-independent properties at a wider latent: backward is linear in dD; two backward passes through the
    native grad-sync path accumulate; grads equal the autograd path's."""
    from dualdiffusion_b200.ddp import GradAllReducer
    spec = uo.small_spec()
    sd = uo.synth_state_dict(spec, seed=0)
    gen = torch.Generator().manual_seed(71)
    x = torch.randn(2, 4, 32, 80, generator=gen).to(dev)
    sigma = torch.tensor([0.7, 3.0]).to(dev)
    clap = torch.randn(2, spec.in_channels_emb, generator=gen).to(dev)
    mask = torch.tensor([True, True]).to(dev)
    probe = torch.randn(2, 4, 32, 80, generator=gen).to(dev)
    net = make_train_unet(spec, sd, dev)

    def run(scale):
        net.zero_grad(set_to_none=True)
        emb = net.get_embeddings(clap, mask)
        d = net(x, sigma, None, emb)
        (d * probe * scale).sum().backward()
        return {n: p.grad.clone() for n, p in net.named_parameters() if p.grad is not None}

    g1, g2 = run(1.0), run(2.0)
    for n in g1:
        if g1[n].norm() > 1e-8:
            assert rel_err(g2[n], 2 * g1[n]) < 2e-2, n
    net.grad_sync = GradAllReducer()
    g3 = run(1.0)
    for n in g1:
        assert rel_err(g3[n], g1[n]) < 1e-5 or g1[n].norm() < 1e-8, n
    emb = net.get_embeddings(clap, mask)                                  # second micro-step accumulates
    (net(x, sigma, None, emb) * probe).sum().backward()
    for n, p in net.named_parameters():
        if n in g1 and g1[n].norm() > 1e-8:             # incl. emb_label* / logvar_linear (accumulated by autograd)
            assert rel_err(p.grad, 2 * g1[n]) < 1e-4, n


def test_train_step_cuda_graphs_match_eager_and_batch_shards_add_up(dev):
    """(1) The captured forward / per-bucket backward graphs reproduce the eager schedule exactly (same kernels, same
    order; only the atomically-accumulated scale gradients may differ in summation order).  (2) DDP equivalence
    (SURVEY.md section 8(c)): for a loss that is a sum over items, the gradient of a batch equals the sum of the
    gradients of its shards."""
    spec = uo.small_spec()
    sd = uo.synth_state_dict(spec, seed=0)
    gen = torch.Generator().manual_seed(73)
    x = torch.randn(2, 4, 32, 48, generator=gen).to(dev)
    sigma = torch.tensor([0.7, 3.0]).to(dev)
    clap = torch.randn(2, spec.in_channels_emb, generator=gen).to(dev)
    mask = torch.tensor([True, False]).to(dev)
    probe = torch.randn(2, 4, 32, 48, generator=gen).to(dev)
    net = make_train_unet(spec, sd, dev)

    def run(sl, graphs):
        net.use_cuda_graphs = graphs
        net.zero_grad(set_to_none=True)
        emb = net.get_embeddings(clap[sl], mask[sl])
        d = net(x[sl], sigma[sl], None, emb)
        (d * probe[sl]).sum().backward()
        return d.detach().clone(), {n: p.grad.clone() for n, p in net.named_parameters() if p.grad is not None}

    full = slice(0, 2)
    d_e, g_e = run(full, False)
    for _ in range(3):                       # eager warm-up call, capture + replay, replay
        d_g, g_g = run(full, True)
    assert torch.equal(d_g, d_e)
    for n in g_e:
        assert rel_err(g_g[n], g_e[n]) < 1e-3 or g_e[n].norm() < 1e-8, n      # fp32 atomics: summation order only
    _, g0 = run(slice(0, 1), False)
    _, g1 = run(slice(1, 2), False)
    for n in g_e:
        if g_e[n].norm() > 1e-8:
            assert rel_err(g0[n] + g1[n], g_e[n]) < 2e-2, (n, rel_err(g0[n] + g1[n], g_e[n]))


def test_batched_weight_prep_and_transpose_match_single_calls(dev):
    from dualdiffusion_b200 import ops, _lib as L
    gen = torch.Generator().manual_seed(79)
    shapes = [((512, 32, 3, 3), 8, 0, 0), ((256, 6, 3, 3), 1, 0, 64), ((256, 128, 1, 1), 1, 64, 0), ((768, 96, 3, 3), 8, 0, 0)]
    prep, trans, singles = [], [], []
    gain = torch.tensor([0.7], device=dev)
    for shape, groups, qk, rs in shapes:
        w = torch.randn(shape, generator=gen).to(dev)
        O, I, taps = shape[0], shape[1], shape[2] * shape[3]
        out = torch.zeros((O, rs or taps * I), device=dev, dtype=torch.bfloat16)
        prep.append(dict(w=w, out=out, gain=gain, O=O, I_g=I, taps=taps, normalize=True, perm=L.WPERM_QK if qk else 0,
                         head_dim=qk, row_stride=rs))
        ref = ops.weight_prep(w, gain=gain, normalize=True, qk_head_dim=qk, row_stride=rs)
        singles.append((out, ref))
        if not rs:
            dst = torch.empty((groups * I, taps, O // groups), device=dev, dtype=torch.bfloat16)
            trans.append(dict(src=out, dst=dst, cout_g=O // groups, cin_g=I, taps=taps, groups=groups))
            singles.append((dst, ops.weight_transpose(ref, O, I, taps, groups)))
    buf, rows = ops.make_wprep_descs(prep, dev)
    ops.weight_prep_batched(buf, len(prep), rows)
    buf2, tiles = ops.make_wtrans_descs(trans, dev)
    ops.weight_transpose_batched(buf2, len(trans), tiles)
    # the batched kernel sums a row's squares with one warp, the single-tensor kernel with a block: the scale may differ in
    # its last fp32 bit, i.e. an occasional element rounds to the neighbouring bf16 value -- everything else is identical
    for got, ref in singles:
        g, r = got.view(-1).float(), ref.view(-1).float()
        assert rel_err(g, r) < 1e-3
        assert ((g - r).abs() <= r.abs() * 2 ** -7 + 1e-30).all() and (g != r).float().mean().item() < 0.05


def test_batched_normalize_weights_matches_reference_semantics(dev):
    """UNet.normalize_weights() (module.py:185-191 -> mp_tools.py:375-378) as one in-place launch: equals the oracle's
    per-parameter normalize and bumps the parameter versions so prepared weights are refreshed."""
    spec = uo.small_spec()
    sd = uo.synth_state_dict(spec, seed=0)
    net = make_train_unet(spec, sd, dev)
    gen = torch.Generator().manual_seed(83)
    with torch.no_grad():
        for p in net.parameters():
            if p.ndim >= 2:
                p.mul_(1.0 + 0.5 * torch.rand(p.shape[0], *([1] * (p.ndim - 1)), generator=gen).to(dev))
    before = {n: p.detach().cpu().clone() for n, p in net.named_parameters()}
    vers = {n: p._version for n, p in net.named_parameters()}
    net.normalize_weights()
    for n, p in net.named_parameters():
        if p.ndim >= 2 and n != "logvar_linear.weight":
            assert rel_err(p, uo.normalize(before[n])) < 1e-6, n
            assert p._version > vers[n], n
        else:
            assert torch.equal(p.detach().cpu(), before[n]), n
