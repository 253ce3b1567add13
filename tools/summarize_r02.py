"""Turns the round-2 ncu outputs under gpurun_out/r02/ into the tracked summaries under profiles/.
  python tools/summarize_r02.py"""
import collections
import csv
import re
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "tools"))
from launch_table import load  # noqa: E402

SRC, OUT = ROOT / "gpurun_out" / "r02", ROOT / "profiles"


def short(n):
    return re.sub(r"\(.*", "", n).replace("void ", "").replace("dpe::", "")


def passes(path):
    out = load(path)
    feat = [i for i, o in enumerate(out) if "k_features" in o[1]]
    return [out[a:b] for a, b in zip(feat, feat[1:] + [len(out)])]


def table(seg, title, lines):
    d = collections.OrderedDict()
    for _, n, _, t in seg:
        d.setdefault(short(n), [0, 0.0]); d[short(n)][0] += 1; d[short(n)][1] += t
    tot = sum(v[1] for v in d.values())
    lines.append(f"## {title}: {tot:.3f} ms, {sum(v[0] for v in d.values())} launches\n")
    lines.append("| kernel | launches | total ms | share |\n|---|---:|---:|---:|")
    for n, (c, t) in sorted(d.items(), key=lambda kv: -kv[1][1]):
        if t / tot >= 0.002:
            lines.append(f"| `{n}` | {c} | {t:.3f} | {100 * t / tot:.1f}% |")
    lines.append("")


def launches():
    lines = ["# Round 2: ncu launch lists\n",
             "Command: `ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file <csv> python tools/profile_step.py <molecule> <walkers>`",
             "(one B200; the script runs one Metropolis step with plain launches, one forward-Laplacian E_loc pass and one gradient + KFAC",
             "backward pass, twice; the tables are the SECOND repetition).  Per-launch times under ncu are cold-cache and serialised: compare SHARES",
             "with the live stage timings of `bench.py` (`roofline.eloc_stages_ms`), not absolutes.  Raw lists: `r02_launches_n2.csv`,",
             "`r02_launches_benzene1024.csv`.\n"]
    for name, f, kinds in (("N2 x 4096 walkers", "launches_n2.csv", ("Metropolis step", "(parameter upload) + accept", "E_loc pass", "gradient + KFAC pass")),
                           ("Benzene x 1024 walkers", "launches_benzene1024.csv", ("Metropolis step", "(parameter upload) + accept", "E_loc pass"))):
        ps = passes(SRC / f)
        second = ps[len(ps) // 2:]
        lines.append(f"# {name}\n")
        for seg, kind in zip(second, kinds):
            if "upload" in kind:
                continue
            table(seg, kind, lines)
    (OUT / "r02_launches.md").write_text("\n".join(lines) + "\n")
    for f in ("launches_n2.csv", "launches_benzene1024.csv"):
        (OUT / ("r02_" + f)).write_text((SRC / f).read_text())


KEYS = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM written"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"), ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("launch__registers_per_thread", "regs"), ("launch__shared_mem_per_block_dynamic", "dyn smem"), ("launch__occupancy_limit_registers", "occ lim regs"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts")]


def metrics():
    lines = ["# Round 2: `ncu --set full` key metrics\n",
             "Command per kernel: `ncu --set full --clock-control none --import-source on -k regex:<kernel> -s <skip> -c <n> -o <rep> python tools/profile_step.py N2 4096 <passes>`,",
             "then `ncu -i <rep> --page raw --csv`.  N2 x 4096 walkers, one B200, second repetition of each pass.  One row per captured launch.\n"]
    for f in ("gemm_main", "conv", "det_trace", "mean", "pair_fwd", "det_factor", "fwd_kernels", "grad_atb", "grad_pair", "grad_new_a", "grad_new_b"):
        p = SRC / f"{f}.raw.csv"
        if not p.exists() or p.stat().st_size < 1000:
            continue
        rows = list(csv.reader(open(p)))
        h, units = rows[0], rows[1]
        idx = {k: i for i, k in enumerate(h)}
        cols = [(k, lab) for k, lab in KEYS if k in idx]
        lines.append(f"## {f}\n")
        lines.append("| kernel | grid | " + " | ".join(lab for _, lab in cols) + " |\n|---|---|" + "---:|" * len(cols))
        for r in rows[2:]:
            vals = []
            for k, _ in cols:
                v, u = r[idx[k]], units[idx[k]]
                try:
                    v = f"{float(v.replace(',', '')):.4g}"
                except ValueError:
                    pass
                vals.append(f"{v} {u}".strip())
            lines.append(f"| `{short(r[idx['Kernel Name']])}` | {r[idx['Grid Size']]} | " + " | ".join(vals) + " |")
        lines.append("")
    (OUT / "r02_ncu_metrics.md").write_text("\n".join(lines) + "\n")


if __name__ == "__main__":
    launches()
    metrics()
    print("written", OUT / "r02_launches.md", OUT / "r02_ncu_metrics.md")
