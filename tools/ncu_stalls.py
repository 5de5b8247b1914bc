"""Top stall sites of one kernel from an `ncu --set full --import-source on` report (SASS view): instructions ranked by warp
stall samples, with the dominant stall reasons.  python tools/ncu_stalls.py report.ncu-rep [top_n]"""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hi]
    col = {h: i for i, h in enumerate(hdr)}
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    data = rows[hi + 1:]
    tot = sum(int(r[col["# Samples"]] or 0) for r in data)
    agg = {h: sum(int(r[col[h]] or 0) for r in data) for h in stall_cols}
    print(f"total samples {tot}; by reason: " + ", ".join(f"{h[6:]} {100.0 * v / tot:.1f}%" for h, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
    idx = sorted(range(len(data)), key=lambda i: -int(data[i][col["# Samples"]] or 0))[:top]
    for i in sorted(idx):
        r = data[i]
        n = int(r[col["# Samples"]] or 0)
        reasons = sorted(((int(r[col[h]] or 0), h[6:]) for h in stall_cols), reverse=True)[:3]
        print(f"{i:6d} {100.0 * n / tot:5.1f}%  {r[col['Source']].strip()[:90]:90s} " + " ".join(f"{nm}:{v}" for v, nm in reasons if v))


if __name__ == "__main__":
    main()
