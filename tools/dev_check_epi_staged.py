"""A/B check of the two MPConv epilogues (run on a B200: `python tools/dev_check_epi_staged.py`).

1. bit-exactness: the shared-memory-staged epilogue (DD_EPI_STAGED=1) must produce exactly the bytes of the direct one
   for every epilogue mode, shape class (per-tap igemm / halo, ragged tiles, n_tile 16..128) and epilogue warp count;
2. timing of both on BASELINE-size layers (CUDA events, 20 launches after 3 warm-ups).
The environment switches are read by the launcher on every call, so one process covers all combinations."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from dualdiffusion_b200 import ops, _lib as L   # noqa: E402

dev = torch.device("cuda:0")

CASES = [
    # B, H, W, Cin, Cout, k, groups
    (1, 8, 16, 64, 64, 1, 1), (1, 16, 16, 256, 256, 1, 1), (2, 7, 13, 96, 64, 1, 1), (1, 16, 24, 256, 512, 3, 8),
    (2, 5, 43, 512, 256, 3, 8), (2, 2, 43, 1280, 2560, 3, 8), (1, 4, 4, 1280, 1280, 1, 1), (1, 1, 1, 64, 32, 3, 1),
    (1, 16, 8, 64, 64, 3, 1), (2, 17, 20, 768, 512, 3, 8), (1, 32, 24, 1280, 1024, 3, 8), (1, 16, 344, 512, 256, 3, 8),
    (2, 8, 172, 512, 1024, 3, 8), (3, 8, 20, 128, 64, 3, 2), (1, 12, 9, 64, 32, 3, 1),
    (1, 16, 40, 64, 48, 3, 1), (1, 16, 40, 64, 80, 3, 1), (2, 9, 33, 128, 96, 1, 1), (1, 16, 40, 128, 112, 3, 1),
    (1, 16, 40, 64, 16, 3, 1), (2, 32, 688, 512, 1024, 3, 8), (2, 32, 688, 512, 512, 1, 1),
]

MODES = [
    ("none", dict()),
    ("raw2", dict(epi2=L.EPI2_RAW)),
    ("silu", dict(epi=L.EPI_SCALE_SILU, _scale=True)),
    ("silu+raw2", dict(epi=L.EPI_SCALE_SILU, _scale=True, epi2=L.EPI2_RAW)),
    ("res", dict(epi=L.EPI_RESIDUAL, alpha=0.6, beta=0.8, clip=2.0, _res=True)),
    ("res+silu2", dict(epi=L.EPI_RESIDUAL, alpha=0.6, beta=0.8, _res=True, epi2=L.EPI2_SILU)),
    ("res+scale2", dict(epi=L.EPI_RESIDUAL, alpha=0.6, beta=0.8, clip=1.0, _res=True, epi2=L.EPI2_SCALE, _scale2=True)),
    ("none+silu2", dict(epi2=L.EPI2_SILU)),
]


def setenv(staged, ew, prefetch=0):
    os.environ["DD_EPI_STAGED"] = "1" if staged else "0"
    os.environ["DD_EPI_PREFETCH"] = "1" if prefetch else "0"
    if ew:
        os.environ["DD_FORCE_EPI_WARPS"] = str(ew)
    else:
        os.environ.pop("DD_FORCE_EPI_WARPS", None)


def run(x, wp, k, g, kw, staged, ew, prefetch=0):
    setenv(staged, ew, prefetch)
    B, H, W, _ = x.shape
    Cout = wp.shape[0]
    # poison the outputs so rows the kernel must not touch / forgets to write show up
    out = torch.full((B, H, W, Cout), 7.0, device=dev, dtype=torch.bfloat16)
    out2 = torch.full((B, H, W, Cout), 9.0, device=dev, dtype=torch.bfloat16) if kw.get("epi2") else None
    r = ops.mpconv(x, wp, k, g, out=out, out2=out2, **kw)
    torch.cuda.synchronize()
    return r if isinstance(r, tuple) else (r,)


def main():
    bad = 0
    n = 0
    for (B, H, W, Cin, Cout, k, g) in CASES:
        gen = torch.Generator().manual_seed(B * 1000 + H * 100 + Cin + Cout + k)
        x = torch.randn(B, H, W, Cin, generator=gen).to(dev).to(torch.bfloat16)
        wp = ops.weight_prep(torch.randn(Cout, Cin // g, k, k, generator=gen).to(dev))
        big = B * H * W > 20000
        for name, spec in MODES:
            kw = {a: b for a, b in spec.items() if not a.startswith("_")}
            if spec.get("_scale"):
                kw["scale"] = (torch.randn(B, Cout, generator=gen) * 0.3 + 1).to(dev)
            if spec.get("_scale2"):
                kw["scale2"] = torch.randn(B, Cout, generator=gen).to(dev)
            if spec.get("_res"):
                kw["residual"] = torch.randn(B, H, W, Cout, generator=gen).to(dev).to(torch.bfloat16)
            for ew in ((0,) if big else (0, 4, 8, 12)):
                ref = run(x, wp, k, g, kw, False, ew)
                variants = [("staged", run(x, wp, k, g, kw, True, ew))]
                if "residual" in kw:        # next-tile residual prefetch (no effect on results)
                    variants.append(("staged+prefetch", run(x, wp, k, g, kw, True, ew, 1)))
                    variants.append(("direct+prefetch", run(x, wp, k, g, kw, False, ew, 1)))
                for vname, got in variants:
                    n += 1
                    for i, (a, b) in enumerate(zip(ref, got)):
                        if not torch.equal(a.view(torch.int16), b.view(torch.int16)):
                            bad += 1
                            d = (a.float() - b.float()).abs()
                            print(f"MISMATCH {(B, H, W, Cin, Cout, k, g)} {name} {vname} ew={ew} out{i}: "
                                  f"{int((d > 0).sum())} / {d.numel()} differ, max {float(d.max()):.4g}, "
                                  f"nan {int(torch.isnan(b.float()).sum())}; first (b,h,w,c): "
                                  f"{(d > 0).nonzero()[:6].tolist()}", flush=True)
    print(f"bit-exactness: {n} combinations, {bad} mismatching outputs", flush=True)

    shapes = [(2, 32, 688, 512, 256, 3, 8), (2, 32, 688, 256, 512, 3, 8), (2, 32, 688, 512, 1024, 3, 8),
              (2, 16, 344, 1024, 512, 3, 8), (2, 16, 344, 1024, 2048, 3, 8), (2, 32, 688, 512, 512, 1, 1),
              (2, 8, 172, 2048, 2048, 3, 8)]
    for (B, H, W, Cin, Cout, k, g) in shapes:
        x = torch.randn(B, H, W, Cin, device=dev).to(torch.bfloat16)
        wp = ops.weight_prep(torch.randn(Cout, Cin // g, k, k, device=dev))
        res = torch.randn(B, H, W, Cout, device=dev).to(torch.bfloat16)
        sc = torch.ones(B, Cout, device=dev)
        for name, kw in (("none", {}), ("silu", dict(epi=L.EPI_SCALE_SILU, scale=sc)),
                         ("res", dict(epi=L.EPI_RESIDUAL, alpha=0.5, beta=0.5, residual=res)),
                         ("res+silu2", dict(epi=L.EPI_RESIDUAL, alpha=0.5, beta=0.5, residual=res, epi2=L.EPI2_SILU)),
                         ("silu+raw2", dict(epi=L.EPI_SCALE_SILU, scale=sc, epi2=L.EPI2_RAW))):
            line = f"{(B, H, W, Cin, Cout, k, g)} {name:10s}"
            full = os.environ.get("DD_CHECK_FULL_SWEEP") is not None
            combos = [(st, ew, 0) for st in (0, 1) for ew in (0, 4, 8, 12)] if full else \
                     [(st, 0, pf) for st in (0, 1) for pf in ((0, 1) if "residual" in kw else (0,))]
            for staged, ew, pf in combos:
                if True:
                    setenv(staged, ew, pf)
                    out = torch.empty(B, H, W, Cout, device=dev, dtype=torch.bfloat16)
                    out2 = torch.empty_like(out) if kw.get("epi2") else None
                    for _ in range(3):
                        ops.mpconv(x, wp, k, g, out=out, out2=out2, **kw)
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for _ in range(20):
                        ops.mpconv(x, wp, k, g, out=out, out2=out2, **kw)
                    e1.record()
                    torch.cuda.synchronize()
                    line += f" | s{staged} ew{ew or 'auto'} pf{pf} {e0.elapsed_time(e1) / 20 * 1e3:6.1f}"
            print(line, flush=True)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
