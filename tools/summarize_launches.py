"""Summarises an ncu `--metrics gpu__time_duration.sum --csv` launch list per kernel name."""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1], errors="replace")))
hdr, agg = None, collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    if "Kernel Name" in r:
        hdr = r
        continue
    if hdr is None or len(r) != len(hdr):
        continue
    d = dict(zip(hdr, r))
    if d.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = d["Kernel Name"].split("(")[0].replace("void ", "").replace("<unnamed>::", "")[:70]
    v = float(d["Metric Value"].replace(",", ""))
    u = d["Metric Unit"]
    v = v / 1e3 if u.startswith("n") else (v * 1e3 if u.startswith("m") else v)
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(v[1] for v in agg.values())
print(f"total {tot:.1f} us over {sum(v[0] for v in agg.values())} launches (cold-cache, serialised: compare shares)")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{v[1]:10.1f} us {v[0]:5d}  {100 * v[1] / tot:5.1f}%  {k}")
