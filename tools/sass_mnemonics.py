"""Counts the Blackwell-specific SASS mnemonics per kernel of the built library (cuobjdump -sass): UTCHMMA (tcgen05.mma),
UTMALDG / UTMASTG (TMA load / store), LDTM (tcgen05.ld), HMMA (mma.sync), UTCBAR (tcgen05.commit) and the
BRA.U.ANY lane-serialisation loops that a single-lane TMA producer loop compiles to (must be 0).
    python tools/sass_mnemonics.py > profiles/r02_sass_mnemonics.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "dualdiffusion_b200", "lib", "libdualdiffusion_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
keys = ["UTCHMMA", "UTMALDG", "UTMASTG", "LDTM", "HMMA", "UTCBAR", "BRA.U.ANY"]
counts, name = collections.OrderedDict(), None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = m.group(1)
        counts[name] = collections.Counter()
        continue
    if name is None:
        continue
    for k in keys:
        if re.search(r"\b" + re.escape(k) + r"\b", line) or (k == "HMMA" and " HMMA." in line):
            counts[name][k] += 1
print(f"{'kernel (sm_100a SASS of libdualdiffusion_b200.so)':110s}" + "".join(f"{k:>9s}" for k in keys))
for n, c in sorted(counts.items()):
    if any(c[k] for k in keys):
        short = re.sub(r"^_ZN\d+_GLOBAL__N__[0-9a-f]+_\d+_", "", n)
        print(f"{short[:108]:110s}" + "".join(f"{c[k]:9d}" for k in keys))
