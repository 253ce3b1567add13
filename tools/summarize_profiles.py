"""Turns the ncu outputs under gpurun_out/ into the tracked summaries under profiles/.
  python tools/summarize_profiles.py <launch_csv> <tag> [<ncu-rep> ...]"""
import collections
import csv
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
OUT = ROOT / "profiles"


def launch_table(path, tag):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = [(int(r["ID"]), r["Kernel Name"], float(r["Metric Value"].replace(",", ""))) for r in csv.DictReader(lines)
            if r.get("Metric Name") == "gpu__time_duration.sum"]
    idx = [i for i, r in enumerate(rows) if "k_features" in r[1]]
    out = [f"# ncu launch list `{Path(path).name}` ({tag})\n",
           "Command: `ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 330 --csv "
           "python bench.py --steps 1 --warmup 1 --burn-in 2 --no-cpu-baseline --no-e2e` (N2, 4096 walkers, 1 B200).",
           "Per-launch times under ncu are serialised: compare SHARES, not absolutes.\n"]

    def agg(seg, title):
        d = collections.OrderedDict()
        for _, n, t in seg:
            n = re.sub(r"\(.*", "", n).replace("void ", "").replace("dpe::", "")
            d.setdefault(n, [0, 0.0]); d[n][0] += 1; d[n][1] += t
        tot = sum(v[1] for v in d.values())
        out.append(f"## {title} — {tot / 1e6:.3f} ms, {sum(v[0] for v in d.values())} launches\n")
        out.append("| kernel | launches | total ms | share |\n|---|---:|---:|---:|")
        for n, (c, t) in sorted(d.items(), key=lambda kv: -kv[1][1]):
            out.append(f"| `{n}` | {c} | {t / 1e6:.3f} | {100 * t / tot:.1f}% |")
        out.append("")

    segs = [rows[a:b] for a, b in zip(idx, idx[1:] + [len(rows)])]
    eloc = [sg for sg in segs if any("k_pair_stream<3>" in n for _, n, _ in sg) and any("k_combine" in n for _, n, _ in sg)]
    step = [sg for sg in segs if any("k_accept" in n for _, n, _ in sg) and any("k_pair_stream<1>" in n for _, n, _ in sg)]
    agg(eloc[-1], "forward-Laplacian E_loc pass")
    agg(step[-1], "one Metropolis step (forward pass + propose/accept)")
    (OUT / f"{tag}_launches.md").write_text("\n".join(out))


WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum"]


def metrics_table(reps, tag):
    out = [f"# ncu --set full captures ({tag})\n", "`ncu --set full --clock-control none --import-source on -k regex:... python tools/profile_eloc.py` "
           "(one E_loc pass, N2, 4096 walkers). Values per launch.\n"]
    for rep in reps:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(raw.splitlines()))
        hdr, units = rows[0], rows[1]
        for r in rows[2:]:
            d, u = dict(zip(hdr, r)), dict(zip(hdr, units))
            out.append(f"## `{d['Kernel Name'][:70]}`\n\n| metric | value | unit |\n|---|---:|---|")
            for k in WANT:
                if k in d:
                    out.append(f"| {k} | {d[k]} | {u[k]} |")
            out.append("")
    (OUT / f"{tag}_ncu_metrics.md").write_text("\n".join(out))


if __name__ == "__main__":
    launch_table(sys.argv[1], sys.argv[2])
    if len(sys.argv) > 3:
        metrics_table(sys.argv[3:], sys.argv[2])
    print("wrote", sorted(p.name for p in OUT.glob(sys.argv[2] + "*")))
