#!/usr/bin/env python
"""SASS evidence per kernel (B200_PROFILING.md, "What proves a Blackwell-native kernel"): counts of the tcgen05 / TMEM / TMA
mnemonics in every kernel of the built library.  Runs on the CPU box:  python scripts/sass_summary.py > profiles/rNN_sass_summary.md"""
import re
import subprocess
import sys
from collections import OrderedDict
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
LIB = ROOT / "hicdiff_b200" / "lib" / "libhicdiff_b200.so"
MNEMONICS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTCCP", "UTMALDG", "UTMASTG", "UBLKCP", "HMMA", "SYNCS", "MUFU"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    kernels = OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = {k: 0 for k in MNEMONICS}
            kernels[cur]["insts"] = 0
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1)
            kernels[cur]["insts"] += 1
            for k in MNEMONICS:
                if op.startswith(k):
                    kernels[cur][k] += 1
    demangled = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    print("# SASS mnemonic counts per kernel of hicdiff_b200/lib/libhicdiff_b200.so (`cuobjdump -sass`, sm_100a)\n")
    print("UTCHMMA = tcgen05.mma, UTCBAR = tcgen05.commit, LDTM / STTM = tcgen05.ld / st, UTCCP = tcgen05.cp, UTMALDG / UTMASTG = TMA tensor "
          "load / store, UBLKCP = 1-D bulk copy, HMMA = mma.sync (legacy path), SYNCS = mbarrier ops, MUFU = SFU.\n")
    print("| kernel | SASS insts | " + " | ".join(MNEMONICS) + " |")
    print("|---|---:|" + "---:|" * len(MNEMONICS))
    tot = {k: 0 for k in MNEMONICS}
    for (name, c), dn in zip(kernels.items(), demangled):
        short = re.sub(r"\(.*$", "", dn.replace("(anonymous namespace)::", "")).replace("void ", "").replace("hd::", "")
        print(f"| `{short}` | {c['insts']} | " + " | ".join(str(c[k]) if c[k] else "" for k in MNEMONICS) + " |")
        for k in MNEMONICS:
            tot[k] += c[k]
    print(f"| **total ({len(kernels)} kernels)** | | " + " | ".join(str(tot[k]) for k in MNEMONICS) + " |")


if __name__ == "__main__":
    sys.exit(main())
