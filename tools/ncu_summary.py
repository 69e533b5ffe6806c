"""Per-kernel summary of an `ncu --set full` report (read here, without a GPU):
    ncu -i gpurun_out/r02_layer_full.ncu-rep --page raw --csv > /tmp/raw.csv
    python tools/ncu_summary.py /tmp/raw.csv > profiles/r02_layer_ncu.md
One row per kernel (the slowest launch of each): duration, DRAM bytes read / written, DRAM and SM throughput in % of
peak, tensor-pipe activity, issue-slot activity, registers, achieved occupancy."""
import csv
import re
import sys

WANT = {
    "gpu__time_duration.sum": "dur",
    "dram__bytes_read.sum": "rd",
    "dram__bytes_write.sum": "wr",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pct",
    "sm__inst_executed_pipe_tensor.sum": "tensor_inst",
    "sm__issue_active.avg.pct_of_peak_sustained_active": "issue_pct",
    "launch__registers_per_thread": "regs",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "occ_pct",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "bank_conf",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active": "fp64_pct",
}


def num(s):
    try:
        return float(s.replace(",", ""))
    except ValueError:
        return float("nan")


def main(path):
    with open(path, newline="") as fid:
        lines = [ln for ln in fid if not ln.startswith("==")]
    rd = csv.reader(lines)
    header = next(rd)
    units = next(rd)
    idx = {h: i for i, h in enumerate(header)}
    cols = {k: idx[k] for k in WANT if k in idx}
    best = {}
    for row in rd:
        if len(row) < len(header):
            continue
        name = re.sub(r"\(.*$", "", row[idx["Kernel Name"]]).replace("effq::", "").replace("void ", "")
        rec = {WANT[k]: num(row[i]) for k, i in cols.items()}
        rec["unit_dur"] = units[cols["gpu__time_duration.sum"]] if "gpu__time_duration.sum" in cols else "ns"
        rec["unit_rd"] = units[cols["dram__bytes_read.sum"]] if "dram__bytes_read.sum" in cols else "byte"
        rec["unit_wr"] = units[cols["dram__bytes_write.sum"]] if "dram__bytes_write.sum" in cols else "byte"
        rec["grid"] = row[idx["Grid Size"]] if "Grid Size" in idx else ""
        rec["block"] = row[idx["Block Size"]] if "Block Size" in idx else ""
        key = (name, rec["grid"], rec["block"])
        if key not in best or rec.get("dur", 0) > best[key].get("dur", 0):
            best[key] = rec
    scale_b = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    scale_t = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6}
    print("| kernel | grid x block | us | DRAM read MB | DRAM write MB | DRAM % | SM % | tensor pipe % | fp64 pipe % | issue % | regs | warps active % |")
    print("|---|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|")
    for (name, grid, block), r in sorted(best.items(), key=lambda kv: -kv[1].get("dur", 0) * scale_t.get(kv[1]["unit_dur"], 1e-3)):
        us = r.get("dur", float("nan")) * scale_t.get(r["unit_dur"], 1e-3)
        rdm = r.get("rd", float("nan")) * scale_b.get(r["unit_rd"], 1.0) / 1e6
        wrm = r.get("wr", float("nan")) * scale_b.get(r["unit_wr"], 1.0) / 1e6
        print(f"| `{name[:70]}` | {grid} x {block} | {us:.1f} | {rdm:.1f} | {wrm:.1f} | {r.get('dram_pct', float('nan')):.1f} | "
              f"{r.get('sm_pct', float('nan')):.1f} | {r.get('tensor_pct', float('nan')):.1f} | {r.get('fp64_pct', float('nan')):.1f} | "
              f"{r.get('issue_pct', float('nan')):.1f} | {r.get('regs', float('nan')):.0f} | {r.get('occ_pct', float('nan')):.1f} |")


if __name__ == "__main__":
    main(sys.argv[1])
