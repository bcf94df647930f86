"""Opcode evidence for the kernels' claims (tcgen05 / TMEM / TMA / bulk copies): disassembles the in-tree
library with `cuobjdump -sass` and counts, per kernel, the mnemonics that prove them
(B200_PROFILING.md: UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG = cp.async.bulk.tensor,
UBLKCP = cp.async.bulk, UTCBAR = tcgen05.commit, SYNCS = mbarrier ops), plus instruction counts and registers.

    python profiles/sass_opcodes.py > profiles/r02_sass_opcodes.txt
"""

from __future__ import annotations

import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "pikazoo_b200", "csrc", "libpikazoo_b200.so")
WATCH = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UBLKCP", "SYNCS", "HMMA", "LDGSTS", "LDSM", "MUFU", "IMAD",
         "REDUX", "VOTE", "SHFL", "ATOMG", "RED", "BAR", "LDG", "STG", "LDS", "STS", "LDL", "STL"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    regs = dict(re.findall(r"Function (\S+):\n\s*REG:(\d+)", res))
    archs = sorted(set(re.findall(r"arch = (sm_\w+)", sass)))
    kernels = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = kernels.setdefault(m.group(1), collections.Counter())
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and cur is not None:
            cur["_total"] += 1
            cur[m.group(1)] += 1
    names = demangle(list(kernels))
    print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)}: {len(kernels)} kernels, cubin architectures {archs}")
    print("# columns: SASS instructions | registers | watched mnemonics (count > 0 only)")
    for k, c in kernels.items():
        short = re.sub(r"\(.*", "", names.get(k, k))
        watched = " ".join(f"{w}={c[w]}" for w in WATCH if c[w])
        print(f"{short:<64} {c['_total']:>6} | {regs.get(k, '?'):>3} | {watched}")
    tot = collections.Counter()
    for c in kernels.values():
        tot.update(c)
    print("# library totals: " + " ".join(f"{w}={tot[w]}" for w in ("UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UBLKCP")))


if __name__ == "__main__":
    sys.exit(main())
