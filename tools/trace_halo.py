"""Role timeline of the 3x3 halo conv kernel (round-2 diagnostic for the per-tile fixed cost).

    DD_CONV_TRACE=1 python tools/trace_halo.py [B H W Cin Cout groups [epi]]      (defaults: the two level-0 res layers)

Runs the traced instantiation (`conv3x3_halo_kernel<EW, true>`), reads the clock64 stamps of the first 4 CTAs through
`dd_conv_trace_read` and prints, per CTA, where each warp role spends a tile: producer (waiting for a free stage / issuing),
MMA warp (waiting for an accumulator / for operands / issuing), epilogue group (waiting for the accumulator / working), and
the steady-state period per tile of each role.  The role whose busy time equals the period is the limiter; a role that
mostly waits names the role it waits for.  epi: 0 none, 1 scale+silu, 2 residual."""
import ctypes
import os
import sys

os.environ.setdefault("DD_CONV_TRACE", "1")
import torch  # noqa: E402

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dualdiffusion_b200 import _lib as L, ops  # noqa: E402

CTAS, SLOTS, TILES = 4, 12, 64
NAMES = ["tma_free", "tma_issued", "mma_acc", "mma_operands", "mma_issued", "epi_full", "epi_done", "epi_start"]


def read_trace():
    stamps = (ctypes.c_ulonglong * (CTAS * SLOTS * TILES))()
    meta = (ctypes.c_int * 8)()
    L.check(L.load().dd_conv_trace_read(stamps, CTAS * SLOTS * TILES, meta))
    t = torch.tensor(list(stamps), dtype=torch.int64).view(CTAS, SLOTS, TILES)
    return t, list(meta)


def report(t, meta, mhz):
    num_tiles, grid, n_tile, a_stages, nbuf, ew, kchunks, staged = meta
    print(f"tiles {num_tiles} grid {grid} n_tile {n_tile} a_stages {a_stages} nbuf {nbuf} epilogue warps {ew} "
          f"kchunks {kchunks} staged {staged}")
    us = lambda c: c / mhz          # cycles -> microseconds at the SM clock
    for cta in range(CTAS):
        s = t[cta]
        n = int((s[4, :62] > 0).sum())            # (items 62 / 63 carry per-CTA stamps, never tiles)
        if n < 4:
            continue
        t0 = int(s[:, :n][s[:, :n] > 0].min())
        if meta[7] == 2 and int((s[1, :n] > 0).sum()) < n:     # two sub-tiles per activation box: producer stamps on even items
            for row in (0, 1):
                for i in range(1, n, 2):
                    s[row, i] = s[row, i - 1]
        r = (s[:, :n] - t0).double()
        print(f"-- CTA {cta}: {n} tiles, {us(float(r.max())):.1f} us from first stamp to last")
        period = lambda row: float((r[row, n - 1] - r[row, 2]) / max(1, n - 3))
        print(f"   period per tile: producer {us(period(1)):.2f} us, MMA {us(period(4)):.2f} us, epilogue {us(period(6)):.2f} us"
              f"   ({ew // 4} epilogue group(s) take alternate tiles: a group may spend {ew // 4} periods on one tile)")
        tma_wait = (r[0, 1:n] - r[1, :n - 1]).clamp(min=0)      # stage-free wait after the previous tile's issue
        tma_issue = r[1, :n] - r[0, :n]
        mma_acc_wait = (r[2, 1:n] - r[4, :n - 1]).clamp(min=0)
        mma_opnd_wait = r[3, :n] - r[2, :n]
        mma_issue = r[4, :n] - r[3, :n]
        epi_wait = r[5, :n] - r[7, :n]
        epi_work = r[6, :n] - r[5, :n]
        lat = r[3, :n] - r[1, :n]                               # last box issued -> first box seen by the MMA warp
        mid = slice(2, n - 1)
        for name, v in (("producer waits for a free stage", tma_wait), ("producer issues", tma_issue),
                        ("MMA waits for an accumulator", mma_acc_wait), ("MMA waits for operands", mma_opnd_wait),
                        ("MMA issues", mma_issue), ("epilogue waits for the accumulator", epi_wait),
                        ("epilogue works", epi_work), ("TMA issue -> operands visible", lat)):
            w = v[mid] if v.numel() > 3 else v
            print(f"   {name:38s} mean {us(float(w.mean())):6.2f} us   max {us(float(w.max())):6.2f} us")
        if int((s[9, :n] > 0).sum()) == n:                      # epilogue phases (conv3x3_dx_kernel only)
            ph = [("wait for the slab (previous TMA store read)", r[8, :n] - r[5, :n]), ("TMEM loads + shuffle-add + math + staging", r[9, :n] - r[8, :n]),
                  ("proxy fence + TMA store issue", r[10, :n] - r[9, :n]), ("release accumulator", r[6, :n] - r[10, :n])]
            for name, v in ph:
                print(f"      epilogue: {name:45s} mean {us(float(v[mid].mean())):6.2f} us")
        if int(s[0, 63]) > 0 and int(s[6, 63]) > 0:            # entry / set-up / exit stamps (conv3x3_dx_kernel only)
            e0, e1, e2 = float(s[0, 63] - t0), float(s[1, 63] - t0), float(s[6, 63] - t0)
            print(f"   CTA entry at {us(e0):.2f} us, set-up done {us(e1):.2f} us, first UMMAs issued {us(float(r[4, 0])):.2f} us, "
                  f"last epilogue done {us(float(r[6, :n].max())):.2f} us, exit {us(e2):.2f} us   (CTA lifetime {us(e2 - e0):.2f} us)")
            if int(s[5, 63]) > 0:                              # finer set-up stamps
                f = lambda slot: us(float(s[slot, 63] - t0))
                print(f"   set-up: barriers initialised {f(2):.2f}, first weight panel requested {f(3):.2f}, TMEM alloc "
                      f"{f(4):.2f} -> {f(5):.2f}, previous grid complete (griddepcontrol.wait) {f(7):.2f}")
        print("   first tiles (us since first stamp): " + " | ".join(
            f"{i}: tma {us(float(r[1, i])):.1f} mma {us(float(r[4, i])):.1f} epi {us(float(r[6, i])):.1f}" for i in range(min(n, 6))))


def main():
    dev = torch.device("cuda:0")
    args = [int(a) for a in sys.argv[1:]]
    shapes = [tuple(args[:6]) + ((args[6],) if len(args) > 6 else (0,))] if len(args) >= 6 else [
        (2, 32, 688, 256, 512, 8, 1), (2, 32, 688, 512, 256, 8, 2), (2, 32, 688, 512, 256, 8, 0),
        (2, 16, 344, 512, 1024, 8, 1), (2, 16, 344, 1024, 512, 8, 2)]
    mhz = 1965                                     # SM clock under load on this pool's B200s (nvidia-smi max); stamps are SM cycles
    for (B, H, W, Cin, Cout, g, epi) in shapes:
        x = torch.randn(B, H, W, Cin, device=dev).to(torch.bfloat16)
        wp = ops.weight_prep(torch.randn(Cout, Cin // g, 3, 3, device=dev))
        kw = {}
        if epi == 1:
            kw = dict(epi=1, scale=torch.ones(B, Cout, device=dev))
        elif epi == 2:
            kw = dict(epi=2, residual=torch.randn(B, H, W, Cout, device=dev).to(torch.bfloat16), alpha=0.7, beta=0.3, clip=256.0)
        for _ in range(3):
            ops.mpconv(x, wp, 3, g, **kw)
        torch.cuda.synchronize()
        t, meta = read_trace()
        print(f"=== B{B} {H}x{W} {Cin}->{Cout} g{g} epi {epi} (SM clock ~{mhz} MHz)")
        report(t, meta, float(mhz))


if __name__ == "__main__":
    main()
