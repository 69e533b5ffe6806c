"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel table
(launch count, total time, share of the profiled GPU time).

    python tools/summarize_launches.py gpurun_out/launches.csv > profiles/rNN_launches.md

The per-launch times under ncu are cold-cache and serialised: compare SHARES with the live
CUDA-event numbers of bench.py, not absolute values."""
import csv
import re
import sys
from collections import defaultdict


def short(name: str) -> str:
    name = re.sub(r"\(.*$", "", name)                       # drop the argument list
    name = re.sub(r"^void\s+", "", name)
    name = name.replace("effq::", "")
    name = re.sub(r"at::native::(\(anonymous namespace\)::)?", "at::", name)
    if len(name) > 90:
        name = name[:87] + "..."
    return name


def main(path):
    rows = []
    with open(path, newline="") as fid:
        lines = [ln for ln in fid if not ln.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        try:
            val = float(r["Metric Value"].replace(",", ""))
        except (KeyError, ValueError):
            continue
        unit = r.get("Metric Unit", "ns")
        scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
        rows.append((short(r["Kernel Name"]), val * scale))
    agg = defaultdict(lambda: [0, 0.0])
    for k, ms in rows:
        agg[k][0] += 1
        agg[k][1] += ms
    total = sum(v[1] for v in agg.values())
    print(f"launches profiled: {len(rows)}   total GPU time under ncu: {total:.1f} ms\n")
    print("| kernel | launches | total ms | share | avg us |")
    print("|---|---:|---:|---:|---:|")
    for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
        print(f"| `{k}` | {n} | {ms:.1f} | {100 * ms / total:.1f} % | {1e3 * ms / n:.1f} |")
    ours = sum(v[1] for k, v in agg.items() if "_kernel" in k and not k.startswith("at::"))
    print(f"\nshare of effq_b200 kernels: {100 * ours / total:.1f} %")


if __name__ == "__main__":
    main(sys.argv[1])
