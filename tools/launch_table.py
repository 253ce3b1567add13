"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals for the launches between two markers."""
import collections
import csv
import sys


def load(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    h = rows[hi]
    ki, vi, ui, gi = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit"), h.index("Grid Size")
    out = []
    for r in rows[hi + 2:]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v = v / 1e6 if r[ui] in ("ns", "nsecond") else (v / 1e3 if r[ui] in ("us", "usecond") else v)
        out.append((int(r[0]), r[ki], r[gi], v))
    return out


if __name__ == "__main__":
    out = load(sys.argv[1])
    marker = sys.argv[2] if len(sys.argv) > 2 else "k_features"
    which = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [2]
    thr = float(sys.argv[4]) if len(sys.argv) > 4 else 0.3
    feat = [i for i, o in enumerate(out) if marker in o[1]]
    for idx in which:
        if len(feat) > idx + 1:
            seg = out[feat[idx]:feat[idx + 1]]
            agg = collections.defaultdict(lambda: [0, 0.0])
            for o in seg:
                agg[o[1][:48]][0] += 1
                agg[o[1][:48]][1] += o[3]
            print("call", idx, "total ms", round(sum(o[3] for o in seg), 3), "launches", len(seg))
            for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:14]:
                print(f"   {t:8.3f} ms {n:4d} {k}")
            for o in seg:
                if o[3] > thr:
                    print("      ", o[0], o[1][:44], o[2], round(o[3], 3))
