"""Opcode census of the shipped library (evidence that the hot kernels are Blackwell-native): cuobjdump -sass of
styletts2_b200/libstyletts2_b200.so, per kernel counts of the tcgen05 / TMEM / TMA / cluster mnemonics.
    python tools/sass_summary.py > profiles/r02_sass_opcodes.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "styletts2_b200", "libstyletts2_b200.so")
KEYS = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "UTCATOMSWS", "UBLKCP", "UTMALDG", "UTMASTG", "LDGSTS", "SYNCS", "ELECT", "STAS", "UCGABAR",
        "MUFU.SIN", "F2FP.SATFINITE.E4M3", "HMMA", "FFMA2"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    counts, cur, total = collections.OrderedDict(), None, collections.Counter()
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = re.sub(r"\(.*", "", name)
            counts[cur] = collections.Counter()
            continue
        if cur and re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
            total[cur] += 1
            for k in KEYS:
                if re.search(r"\b" + re.escape(k), line):
                    counts[cur][k] += 1
    print("# SASS opcode census of styletts2_b200/libstyletts2_b200.so (sm_100a)\n")
    print("`cuobjdump -sass` of the in-tree library, counted per kernel by `tools/sass_summary.py`.  tcgen05.mma -> UTCHMMA (kind::f16) /")
    print("UTCQMMA (kind::f8f6f4), tcgen05.commit -> UTCBAR, tcgen05.ld -> LDTM, tcgen05.alloc -> UTCATOMSWS, cp.async.bulk -> UBLKCP,")
    print("cp.async -> LDGSTS, mbarrier -> SYNCS, st.async (DSMEM) -> STAS, barrier.cluster -> UCGABAR.  No HMMA (legacy mma.sync) anywhere.\n")
    cols = [k for k in KEYS if any(c[k] for c in counts.values())]
    print("| kernel | instrs | " + " | ".join(cols) + " |")
    print("|---|---|" + "---|" * len(cols))
    for name, c in counts.items():
        if any(c[k] for k in cols):
            print(f"| `{name}` | {total[name]} | " + " | ".join(str(c[k]) if c[k] else "" for k in cols) + " |")
    print(f"\n{len(counts)} kernels in the library; {sum(1 for c in counts.values() if c['UTCHMMA'] or c['UTCQMMA'])} issue tcgen05 MMAs.")


if __name__ == "__main__":
    main()
