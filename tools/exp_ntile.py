import os, subprocess, sys
shapes = ["2 2 43 1280 1280 1 1", "2 4 86 1024 1024 1 1", "2 2 43 2560 1280 3 8", "2 8 172 1536 768 3 8", "2 4 86 2048 1024 3 8", "2 2 43 1280 2560 1 1"]
code = r'''
import sys, torch
sys.path.insert(0, ".")
from dualdiffusion_b200 import ops
B, H, W, Cin, Cout, k, g = map(int, sys.argv[1:8])
x = torch.randn(B, H, W, Cin, device="cuda").to(torch.bfloat16)
wp = ops.weight_prep(torch.randn(Cout, Cin // g, k, k, device="cuda"))
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
for _ in range(3): ops.mpconv(x, wp, k, g)
ts = []
for _ in range(7):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ops.mpconv(x, wp, k, g); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e3)
warm = []
for _ in range(7):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ops.mpconv(x, wp, k, g); e1.record(); torch.cuda.synchronize()
    warm.append(e0.elapsed_time(e1) * 1e3)
print("cold %.1f us warm %.1f us" % (sorted(ts)[3], sorted(warm)[3]))
'''
open("/tmp/_one.py", "w").write(code)
for sh in shapes:
    for n in ("auto", "16", "32", "64", "128", "256"):
        env = dict(os.environ, DD_DEBUG_CONV="1")
        if n != "auto": env["DD_FORCE_NTILE"] = n
        r = subprocess.run([sys.executable, "/tmp/_one.py"] + sh.split(), env=env, capture_output=True, text=True)
        cfg = [l for l in r.stderr.splitlines() if l.startswith("[conv]")]
        print(sh, "| n", n, "|", r.stdout.strip(), "|", cfg[-1][7:] if cfg else r.stderr[-200:], flush=True)
