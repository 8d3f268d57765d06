"""Per-kernel table of an `ncu --metrics gpu__time_duration.sum --csv` launch list (profiles/rNN_launches.csv):
    python tools/launch_table.py gpurun_out/r02_launches.csv [--skip N] [--steps S]
--skip drops the first N launches (index open, warm-up); shares are of the remaining total."""
import argparse
import csv
import re
from collections import OrderedDict


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--skip", type=int, default=0)
    ap.add_argument("--only-from", default="", help="start at the first launch of this kernel after --skip")
    a = ap.parse_args()
    rows = [r for r in csv.reader(l for l in open(a.csv) if l.startswith('"'))]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    launches = []
    for r in rows[1:]:
        if len(r) <= vi:
            continue
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0}.get(r[ui], 1e-6)
        launches.append((re.sub(r"\(.*", "", r[ki]).replace("void ", "").strip(), v * scale))
    launches = launches[a.skip:]
    if a.only_from:
        for i, (n, _) in enumerate(launches):
            if n.startswith(a.only_from):
                launches = launches[i:]
                break
    agg = OrderedDict()
    for n, ms in launches:
        c, t = agg.get(n, (0, 0.0))
        agg[n] = (c + 1, t + ms)
    total = sum(t for _, t in agg.values())
    print(f"{len(launches)} launches, {total:.3f} ms")
    print("| kernel | launches | avg ms | total ms | share |")
    print("|---|---|---|---|---|")
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{n}` | {c} | {t / c:.3f} | {t:.3f} | {100 * t / total:.1f}% |")


if __name__ == "__main__":
    main()
