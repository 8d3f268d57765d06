"""Turn `ncu --set full` captures into the tables of profiles/rNN_summary.md and the traffic file bench.py cites.

    python tools/summarize_ncu.py gpurun_out/ncu_*.ncu-rep --csv-out profiles/r02_ncu_full_raw.csv \
        --traffic-out profiles/r02_traffic.json

Reads each report with `ncu -i ... --page raw --csv` (works without a GPU), keeps the last launch of every kernel
name, prints one markdown row per kernel and writes {kernel: {"dram_bytes": read + write, "source": file}}."""
import argparse
import csv
import io
import json
import re
import subprocess
import sys

KEEP = {
    "ms": "gpu__time_duration.sum",
    "dram_rd_GB": "dram__bytes_read.sum",
    "dram_wr_GB": "dram__bytes_write.sum",
    "lts_sectors": "lts__t_sectors.sum",
    "l2_hit_pct": "lts__t_sector_hit_rate.pct",
    "l1tex_pct": "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts_pct": "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "tensor_pct": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "issue_pct": "sm__inst_issued.avg.pct_of_peak_sustained_active",
    "warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
    "warp_inst": "smsp__inst_executed.sum",
    "regs": "launch__registers_per_thread",
}
UNIT_SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3, "msecond": 1.0,
              "usecond": 1e-3, "nsecond": 1e-6, "second": 1e3}


def read_report(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True)
    if out.returncode != 0:
        print(f"ncu could not read {path}: {out.stderr[-300:]}", file=sys.stderr)
        return [], None, None
    rows = list(csv.reader(io.StringIO(out.stdout)))
    return rows[2:], rows[0], rows[1]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("reports", nargs="+")
    ap.add_argument("--csv-out", default="")
    ap.add_argument("--traffic-out", default="")
    a = ap.parse_args()
    table, traffic, raw_rows, raw_hdr = {}, {}, [], None
    for rep in a.reports:
        rows, hdr, units = read_report(rep)
        if not rows:
            continue
        if raw_hdr is None:
            raw_hdr = (hdr, units)
        name_col = hdr.index("Kernel Name")
        for r in rows:
            raw_rows.append(r)
            name = re.sub(r"\(.*", "", r[name_col]).replace("void ", "").strip()
            ent = {}
            for k, metric in KEEP.items():
                cols = [i for i, h in enumerate(hdr) if h == metric]
                if not cols:
                    continue
                try:
                    v = float(r[cols[0]].replace(",", ""))
                except ValueError:
                    continue
                u = units[cols[0]]
                if k == "ms":
                    v *= UNIT_SCALE.get(u, 1.0)
                elif k.endswith("_GB"):
                    v = v * UNIT_SCALE.get(u, 1.0) / 1e9
                ent[k] = v
            ent["source"] = rep
            table[name] = ent
    cols = ["ms", "dram_rd_GB", "dram_wr_GB", "l2_hit_pct", "l1tex_pct", "lts_pct", "dram_pct", "tensor_pct", "issue_pct",
            "warps_active_pct", "regs"]
    print("| kernel | " + " | ".join(cols) + " | L2 sectors x 32 B / ms (TB/s) | warp instr (M) |")
    print("|---|" + "---|" * (len(cols) + 2))
    for name, e in sorted(table.items(), key=lambda kv: -kv[1].get("ms", 0)):
        l2 = e.get("lts_sectors", 0) * 32 / max(e.get("ms", 1e-9), 1e-9) / 1e9
        print(f"| `{name}` | " + " | ".join(f"{e.get(c, float('nan')):.3g}" for c in cols) +
              f" | {l2:.2f} | {e.get('warp_inst', 0) / 1e6:.0f} |")
        traffic[name] = {"dram_bytes": (e.get("dram_rd_GB", 0) + e.get("dram_wr_GB", 0)) * 1e9, "ms": e.get("ms"),
                         "source": e["source"]}
    if a.csv_out and raw_hdr:
        with open(a.csv_out, "w", newline="") as f:
            w = csv.writer(f)
            w.writerow(raw_hdr[0])
            w.writerow(raw_hdr[1])
            w.writerows(raw_rows)
    if a.traffic_out:
        json.dump(traffic, open(a.traffic_out, "w"), indent=1)


if __name__ == "__main__":
    main()
