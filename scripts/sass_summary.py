#!/usr/bin/env python
"""Per-kernel counts of the Blackwell-only SASS mnemonics in libdomainrag_b200.so (SURVEY 5 / 8d evidence):
UTCHMMA / UTCQMMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st), UTMALDG (TMA tensor loads), UBLKCP (cp.async.bulk),
UTCBAR (tcgen05.commit -> mbarrier), SYNCS (mbarrier). Runs on the CPU (cuobjdump).

    python scripts/sass_summary.py [> profiles/r02_sass_summary.txt]
"""
import re
import subprocess
import sys
from collections import Counter, OrderedDict
from pathlib import Path

REPO = Path(__file__).resolve().parents[1]
LIB = REPO / "domain_rag_b200" / "libdomainrag_b200.so"
MNEMONICS = ("UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "UTCATOMSWS", "SYNCS", "HMMA", "FFMA2")


def kernel_counts(lib=LIB):
    sass = subprocess.run(["cuobjdump", "-sass", str(lib)], capture_output=True, text=True, check=True).stdout
    out, name = OrderedDict(), None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
            out[name] = Counter()
            continue
        if name is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+(?:\.[A-Z0-9_.]+)?)", line)
        if m:
            op = m.group(1)
            for mn in MNEMONICS:
                if op.split(".")[0] == mn:
                    out[name][mn] += 1
                    if mn in ("UTCHMMA", "UTMALDG", "UTCBAR", "LDTM") and "." in op:
                        out[name][op] += 1
    return out


def main():
    counts = kernel_counts()
    print(f"# {LIB.name}: Blackwell SASS mnemonics per kernel (cuobjdump -sass, sm_100a)")
    total = Counter()
    for k, c in counts.items():
        if not c:
            continue
        short = re.sub(r"\(.*", "", k)
        print(f"{short}: " + ", ".join(f"{m} x{n}" for m, n in sorted(c.items())))
        for m in MNEMONICS:
            total[m] += c.get(m, 0)
    print("# total: " + ", ".join(f"{m} x{total[m]}" for m in MNEMONICS if total[m]))


if __name__ == "__main__":
    sys.exit(main())
