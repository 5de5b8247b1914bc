"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel table (markdown).
Usage: python tools/summarize_launches.py gpurun_out/launches.csv "<title>" "<command>" > profiles/<name>.md"""
import csv
import io
import re
import sys
from collections import defaultdict


def main():
    path, title, cmd = sys.argv[1], sys.argv[2], sys.argv[3]
    lines = open(path, errors="replace").read().splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
    rows = list(csv.DictReader(io.StringIO("\n".join(lines[start:]))))
    agg = defaultdict(lambda: [0, 0.0])
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        ms = v / 1e6 if unit in ("ns", "nsecond") else (v / 1e3 if unit in ("us", "usecond") else v)
        name = re.sub(r"\(.*", "", r["Kernel Name"]).strip()
        agg[name][0] += 1
        agg[name][1] += ms
    tot = sum(v[1] for v in agg.values())
    n = sum(v[0] for v in agg.values())
    print(f"# {title}\n\ncommand: `{cmd}`\n")
    print("Launches are cold-cache and serialised under ncu: compare SHARES, not absolutes.\n")
    print(f"launches {n} total kernel ms {tot:.2f}\n")
    print("| kernel | launches | ms | share |\n|---|---|---|---|")
    for k, (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k[:70]}` | {c} | {ms:.2f} | {100 * ms / tot:.2f}% |")


if __name__ == "__main__":
    main()
