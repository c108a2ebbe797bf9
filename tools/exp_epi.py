import os, subprocess, sys
code = r'''
import sys, torch
sys.path.insert(0, ".")
from dualdiffusion_b200 import ops
B, H, W, Cin, Cout, k, g, epi = map(int, sys.argv[1:9])
x = torch.randn(B, H, W, Cin, device="cuda").to(torch.bfloat16)
wp = ops.weight_prep(torch.randn(Cout, Cin // g, k, k, device="cuda"))
kw = {}
if epi == 1: kw = dict(epi=1, scale=torch.ones(B, Cout, device="cuda"))
if epi == 2: kw = dict(epi=2, alpha=0.5, beta=0.5, residual=torch.randn(B, H, W, Cout, device="cuda").to(torch.bfloat16))
for _ in range(3): ops.mpconv(x, wp, k, g, **kw)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): ops.mpconv(x, wp, k, g, **kw)
e1.record(); torch.cuda.synchronize()
print("%.1f us" % (e0.elapsed_time(e1) / 20 * 1e3))
'''
open("/tmp/_e.py", "w").write(code)
for sh in ["2 32 688 512 256 3 8", "2 32 688 256 512 3 8", "2 32 688 512 1024 3 8", "2 16 344 1024 512 3 8", "2 32 688 512 512 1 1"]:
    for epi in ("1", "2"):
        for ew in ("8", "12"):
            env = dict(os.environ, DD_FORCE_EPI_WARPS=ew)
            r = subprocess.run([sys.executable, "/tmp/_e.py"] + sh.split() + [epi], env=env, capture_output=True, text=True)
            print(sh, "epi", epi, "ew", ew, r.stdout.strip(), r.stderr[-200:] if r.returncode else "", flush=True)
