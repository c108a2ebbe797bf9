import os, subprocess, sys
code = r'''
import sys, torch
sys.path.insert(0, ".")
from dualdiffusion_b200 import ops
B, H, W, Cin, Cout, k, g = map(int, sys.argv[1:8])
x = torch.randn(B, H, W, Cin, device="cuda").to(torch.bfloat16)
wp = ops.weight_prep(torch.randn(Cout, Cin // g, k, k, device="cuda"))
for _ in range(3): ops.mpconv(x, wp, k, g)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): ops.mpconv(x, wp, k, g)
e1.record(); torch.cuda.synchronize()
print("%.1f us" % (e0.elapsed_time(e1) / 20 * 1e3))
'''
open("/tmp/_t.py", "w").write(code)
for sh in ["2 32 688 512 256 3 8", "2 32 688 512 1024 3 8", "2 32 688 256 512 3 8"]:
    for taps in ("9", "5", "1"):
        for nacc in ("1", "nostore"):
            env = dict(os.environ, DD_DBG_TAPS=taps, DD_FORCE_NACC="1")
            if nacc == "nostore": env["DD_DBG_NOSTORE"] = "1"
            r = subprocess.run([sys.executable, "/tmp/_t.py"] + sh.split(), env=env, capture_output=True, text=True)
            print(sh, "taps", taps, "nacc", nacc, r.stdout.strip(), r.stderr[-200:] if r.returncode else "", flush=True)
