"""Aggregates an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel for the pass that contains a given kernel.
  python tools/launch_summary.py <launch_csv> [marker-substring, default k_det]"""
import collections
import csv
import re
import sys

path = sys.argv[1]
marker = sys.argv[2] if len(sys.argv) > 2 else "k_combine"
lines = [l for l in open(path) if not l.startswith("==")]
rows = [(r["Kernel Name"], float(r["Metric Value"].replace(",", ""))) for r in csv.DictReader(lines)
        if r.get("Metric Name") == "gpu__time_duration.sum"]
idx = [i for i, r in enumerate(rows) if "k_features" in r[0]]
for a, b in zip(idx, idx[1:] + [len(rows)]):
    seg = rows[a:b]
    if not any(marker in n and "<1>" not in n for n, _ in seg) or not any("k_pair_stream<3>" in n for n, _ in seg):
        continue
    d = collections.OrderedDict()
    for n, t in seg:
        n = re.sub(r"\(.*", "", n).replace("void ", "").replace("dpe::", "")
        d.setdefault(n, [0, 0.0]); d[n][0] += 1; d[n][1] += t
    tot = sum(v[1] for v in d.values())
    print(f"launches {a}..{b}: {tot / 1e6:.3f} ms")
    for n, (c, t) in sorted(d.items(), key=lambda kv: -kv[1][1]):
        if t > 2e3:
            print(f"  {n[:60]:60s} {c:4d} {t / 1e6:8.3f}")
    break
