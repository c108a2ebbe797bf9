"""Developer GPU check (not a test, not the product): compares each C-ABI op with torch on the GPU.
Usage: python tools/dev_check_ops.py <case>   — run each case in its own process under `timeout`."""
import math, sys, time
import torch
import torch.nn.functional as F
sys.path.insert(0, ".")
from dualdiffusion_b200 import ops, _lib as L

dev = "cuda"
torch.manual_seed(0)

def rel(a, b):
    a = a.float(); b = b.float()
    return ((a - b).norm() / (b.norm() + 1e-12)).item(), (a - b).abs().max().item()

def nhwc(x):  # NCHW -> NHWC bf16
    return x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)

def ref_conv(x_nhwc, w, groups, gain=1.0):
    xf = x_nhwc.float().permute(0, 3, 1, 2)
    wf = (w.float() * (gain / math.sqrt(w[0].numel()))).to(torch.bfloat16).float()
    y = F.conv2d(xf, wf, padding=w.shape[-1] // 2, groups=groups)
    return y.permute(0, 2, 3, 1)

def conv_case(B, H, W, Cin, Cout, k, g, naive=False, timing=False):
    x = nhwc(torch.randn(B, Cin, H, W, device=dev))
    w = torch.randn(Cout, Cin // g, k, k, device=dev)
    wp = ops.weight_prep(w)
    fn = ops.mpconv_naive if naive else ops.mpconv
    y = fn(x, wp, k, g)
    torch.cuda.synchronize()
    yr = ref_conv(x, w, g)
    r, m = rel(y, yr)
    msg = f"conv{'_naive' if naive else ''} B{B} {H}x{W} {Cin}->{Cout} k{k} g{g}: rel {r:.2e} max {m:.2e}"
    if timing:
        for _ in range(3): fn(x, wp, k, g)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 20
        e0.record()
        for _ in range(n): fn(x, wp, k, g)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        fl = 2 * B * H * W * Cout * (Cin // g) * k * k
        msg += f"  {ms*1e3:.1f} us  {fl/ms/1e9:.1f} TFLOP/s"
    print(msg, "OK" if r < 1e-2 else "FAIL", flush=True)

case = sys.argv[1]
if case == "naive":
    conv_case(1, 8, 16, 64, 64, 1, 1, naive=True)
    conv_case(2, 6, 10, 64, 128, 3, 2, naive=True)
elif case == "gemm_small":
    conv_case(1, 8, 16, 64, 64, 1, 1)
elif case == "gemm_k":
    conv_case(1, 16, 16, 256, 256, 1, 1)
    conv_case(1, 16, 16, 96, 64, 1, 1)
elif case == "conv3":
    conv_case(1, 16, 24, 256, 512, 3, 8)
    conv_case(2, 5, 43, 512, 256, 3, 8)
    conv_case(2, 2, 43, 1280, 2560, 3, 8)
    conv_case(1, 4, 4, 1280, 1280, 1, 1)
elif case == "conv_big":
    conv_case(2, 32, 688, 256, 512, 3, 8, timing=True)
    conv_case(2, 32, 688, 512, 256, 3, 8, timing=True)
    conv_case(2, 32, 688, 512, 1024, 3, 8, timing=True)
    conv_case(2, 32, 688, 512, 512, 1, 1, timing=True)
    conv_case(2, 16, 344, 1024, 1024, 3, 8, timing=True)
    conv_case(2, 8, 172, 1536, 1536, 3, 8, timing=True)
    conv_case(2, 4, 86, 2048, 2048, 3, 8, timing=True)
    conv_case(2, 2, 43, 2560, 2560, 3, 8, timing=True)
    conv_case(2, 2, 43, 1280, 2560, 1, 1, timing=True)
elif case == "epilogue":
    B, H, W, Cin, Cout = 2, 8, 20, 256, 512
    x = nhwc(torch.randn(B, Cin, H, W, device=dev)); w = torch.randn(Cout, Cin // 8, 3, 3, device=dev)
    wp = ops.weight_prep(w); yr = ref_conv(x, w, 8)
    sc = torch.randn(B, Cout, device=dev) * 0.3 + 1
    y = ops.mpconv(x, wp, 3, 8, epi=L.EPI_SCALE_SILU, scale=sc)
    ref = F.silu(yr * sc[:, None, None, :]) / 0.596
    print("epi scale_silu", rel(y, ref))
    res = nhwc(torch.randn(B, Cout, H, W, device=dev))
    sc2 = torch.randn(B, Cout, device=dev)
    y, y2 = ops.mpconv(x, wp, 3, 8, epi=L.EPI_RESIDUAL, alpha=0.4, beta=0.9, clip=1.5, residual=res, epi2=L.EPI2_SCALE, scale2=sc2)
    ref = (0.4 * yr + 0.9 * res.float()).clamp(-1.5, 1.5)
    print("epi residual", rel(y, ref), "out2 scale", rel(y2, ref * sc2[:, None, None, :]))
    y, y2 = ops.mpconv(x, wp, 3, 8, epi=L.EPI_RESIDUAL, alpha=0.4, beta=0.9, residual=res, epi2=L.EPI2_SILU)
    ref = (0.4 * yr + 0.9 * res.float())
    print("epi residual noclip", rel(y, ref), "out2 silu", rel(y2, F.silu(ref) / 0.596))
elif case == "elementwise":
    B, H, W, Cc = 2, 6, 10, 768
    t = nhwc(torch.randn(B, Cc, H, W, device=dev) * 3)
    x, s = ops.pixnorm_silu(t)
    tf = t.float(); n = tf.norm(dim=-1, keepdim=True); xr = tf / (1e-4 + n / math.sqrt(Cc))
    print("pixnorm", rel(x, xr), "silu", rel(s, F.silu(xr) / 0.596))
    a = nhwc(torch.randn(B, 512, H // 2, W // 2, device=dev))
    xc, s = ops.cat_silu(a, None, 1.0, 0.0, True)
    ar = F.interpolate(a.float().permute(0, 3, 1, 2), scale_factor=2, mode="nearest").permute(0, 2, 3, 1)
    print("up", rel(xc, ar), rel(s, F.silu(ar) / 0.596))
    a = nhwc(torch.randn(B, 512, H, W, device=dev)); b = nhwc(torch.randn(B, 256, H, W, device=dev))
    xc, s = ops.cat_silu(a, b, 1.3, 0.7, False)
    cr = torch.cat([1.3 * a.float(), 0.7 * b.float()], -1)
    print("cat", rel(xc, cr), rel(s, F.silu(cr) / 0.596))
    p = ops.avgpool2(a)
    pr = F.avg_pool2d(a.float().permute(0, 3, 1, 2), 2).permute(0, 2, 3, 1)
    print("avgpool", rel(p, pr))
    # weight prep variants
    w = torch.randn(512, 32, 3, 3, device=dev); g = torch.tensor(0.7, device=dev)
    wp = ops.weight_prep(w, gain=g, normalize=True)
    wn = w / (1e-4 + w.flatten(1).norm(dim=1).view(-1, 1, 1, 1) / math.sqrt(288)) * 0.7 / math.sqrt(288)
    print("wprep norm", rel(wp, wn.permute(0, 2, 3, 1).reshape(512, 9, 32)))
    w = torch.randn(256, 128, 1, 1, device=dev)
    wp = ops.weight_prep(w, qk_head_dim=64)
    wr = (w / math.sqrt(128)).view(2, 64, 2, 128).permute(2, 0, 1, 3).reshape(256, 1, 128)
    print("wprep qk perm", rel(wp, wr))
    wp = ops.weight_prep(w.to(torch.bfloat16), fmt=L.WFMT_F32_OIT)
    print("wprep f32 from bf16", rel(wp, w.to(torch.bfloat16).float().view(256, 128, 1) / math.sqrt(128)))
elif case == "emb":
    B, cemb, cn = 3, 768, 256
    sigma = torch.tensor([2.0, 0.05, 150.0], device=dev)
    fr = math.pi * torch.linspace(0, 1 - 1e-3, cn).erfinv().to(dev); ph = (math.pi / 2 * (torch.arange(cn) % 2 == 0).float()).to(dev)
    w = torch.randn(cemb, cn, device=dev); lab = torch.randn(B, cemb, device=dev)
    e = ops.noise_embedding(sigma, fr, ph, w, lab, 0.5)
    f = (torch.outer(sigma.log() / 4, fr) + ph).cos() * math.sqrt(2)
    en = f @ (w / math.sqrt(cn)).t(); m = torch.lerp(en, lab, 0.5) / math.sqrt(0.5)
    er = F.silu(m) / 0.596
    print("noise_emb", rel(e, er))
    w1 = torch.randn(512, 96, 1, 1, device=dev); g1 = torch.tensor(0.5, device=dev); o1 = torch.empty(B, 512, device=dev)
    w2 = torch.randn(256, 768, 1, 1, device=dev).to(torch.bfloat16); o2 = torch.empty(B, 256, device=dev)
    descs, mo = ops.make_affine_descs([dict(w=w1.view(512, 96), gain=g1, out=o1, groups=8, bias=1.0),
                                       dict(w=w2.view(256, 768), gain=None, out=o2, groups=1, bias=1.0, normalize=True)], dev)
    ops.emb_affine(descs, 2, mo, e)
    r1 = F.conv2d(er[:, :, None, None], w1 * 0.5 / math.sqrt(96), groups=8)[:, :, 0, 0] + 1
    w2f = w2.float().view(256, 768); w2n = w2f / (1e-4 + w2f.norm(dim=1, keepdim=True) / math.sqrt(768))
    r2 = er @ (w2n / math.sqrt(768)).t() + 1
    print("emb_affine grouped", rel(o1, r1), "dense bf16 norm", rel(o2, r2))
elif case == "attention":
    for (B, H, W, heads) in [(2, 4, 86, 16), (2, 2, 43, 20), (1, 8, 8, 12), (1, 4, 4, 2)]:
        Cc = heads * 64; N = H * W
        qk = nhwc(torch.randn(B, 2 * Cc, H, W, device=dev) * 2); v = nhwc(torch.randn(B, Cc, H, W, device=dev))
        sv = torch.randn(B, Cc, device=dev) * 0.2 + 1
        y = ops.attention(qk, v, sv, heads)
        def nrm(t): t = t.float(); return t / (1e-4 + t.norm(dim=-1, keepdim=True) / 8)
        q = nrm(qk.view(B, N, 2, heads, 64)[:, :, 0]).permute(0, 2, 1, 3); k = nrm(qk.view(B, N, 2, heads, 64)[:, :, 1]).permute(0, 2, 1, 3)
        vv = nrm(v.view(B, N, heads, 64)).permute(0, 2, 1, 3)
        o = F.scaled_dot_product_attention(q, k, vv).permute(0, 2, 1, 3).reshape(B, H, W, Cc)
        ref = F.silu(o * sv[:, None, None, :]) / 0.596
        print(f"attention B{B} N{N} heads{heads}", rel(y, ref))
elif case == "sampler":
    n = (1, 4, 32, 688)
    d = torch.randn(2, *n[1:], device=dev); s = torch.randn(n, device=dev); cfg = torch.empty(n, device=dev); xh = torch.empty(n, device=dev)
    ops.sampler_cfg_lerp(d, s, 1.5, 0.8, cfg, xh)
    cr = d[1:].lerp(d[:1], 1.5); xr = torch.lerp(cr, s, 0.8)
    print("cfg", rel(cfg, cr), rel(xh, xr))
    d2 = torch.randn(2, *n[1:], device=dev); nz = torch.randn(n, device=dev); s2 = s.clone(); co = torch.empty(n, device=dev)
    ops.sampler_update(cfg, d2, 1.5, True, 0.7, 0.3, nz, s2, co)
    c2 = torch.lerp(cr, d2[1:].lerp(d2[:1], 1.5), 0.5); sr = torch.lerp(c2, s, 0.7) + 0.3 * nz
    print("update", rel(s2, sr), rel(co, c2))
torch.cuda.synchronize()
print("done", case)
