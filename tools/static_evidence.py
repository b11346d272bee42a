#!/usr/bin/env python
"""Static (no GPU) evidence about the built library, for profiles/.

    python tools/static_evidence.py [out.md]

Reads the in-tree ``libls_b200.so`` with ``cuobjdump`` and writes, per kernel: registers,
stack (spill) bytes, static shared memory, SASS instruction count and the counts of the
mnemonics that tell how the kernel moves data (LDG.E.128 / STG.E.128 vector accesses, UBLKCP =
cp.async.bulk, UTMALDG = cp.async.bulk.tensor, SYNCS = mbarrier, CCTL = prefetch, ACQBULK / PDL
trigger, RED/ATOM).  Nothing here is a measurement; it is what `-Xptxas -v` and
`cuobjdump -sass` say about the shipped cubins (all sm_100a).
"""
from __future__ import annotations

import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "e2e_parking_carla_b200", "libls_b200.so")

MNEMONICS = [
    ("LDG.128", re.compile(r"\bLDG\.E(\.\w+)*\.128")),
    ("STG.128", re.compile(r"\bSTG\.E(\.\w+)*\.128")),
    ("LDS", re.compile(r"\bLDS(\.|\b)")),
    ("STS", re.compile(r"\bSTS(\.|\b)")),
    ("SHFL", re.compile(r"\bSHFL\.")),
    ("ATOM/RED", re.compile(r"\b(ATOMG|ATOMS|ATOM|RED|REDG)\.")),
    ("UBLKCP", re.compile(r"\bUBLKCP")),
    ("UBLKPF", re.compile(r"\bUBLKPF")),
    ("UTMALDG", re.compile(r"\bUTMALDG")),
    ("SYNCS", re.compile(r"\bSYNCS\.")),
    ("CCTL", re.compile(r"\bCCTL\.")),
    ("ACQBULK", re.compile(r"\bACQBULK")),
    ("LDL/STL", re.compile(r"\b(LDL|STL)(\.|\b)")),
]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
    short = {}
    for n, d in zip(names, out):
        d = re.sub(r"\(.*$", "", d)                       # drop the parameter list
        d = re.sub(r"^void ", "", d)
        d = d.replace("__nv_bfloat16", "bf16").replace("(LsOutMode)", "out=").replace("(bool)", "")
        short[n] = d
    return short


def res_usage():
    txt = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True, check=True).stdout
    res = {}
    for m in re.finditer(r"Function (\S+):\n\s+REG:(\d+) STACK:(\d+) SHARED:(\d+)", txt):
        res[m.group(1)] = (int(m.group(2)), int(m.group(3)), int(m.group(4)))
    return res


def sass_counts():
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    counts = {}
    cur = None
    for line in txt.split("\n"):
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur is None or "/*" not in line:
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(.*?);", line)
        if not m:
            continue
        ins = m.group(1)
        counts[cur]["instr"] += 1
        for key, rx in MNEMONICS:
            if rx.search(ins):
                counts[cur][key] += 1
    archs = set(re.findall(r"arch = (\S+)", txt))
    return counts, archs


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else None
    res = res_usage()
    counts, archs = sass_counts()
    names = sorted(res)
    short = demangle(names)
    head = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    lines = ["# Static evidence of the shipped cubins (no GPU involved)", "",
             "`python tools/static_evidence.py` on `e2e_parking_carla_b200/libls_b200.so` built from commit `%s`" % head,
             "(`nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3`; `cuobjdump -res-usage` + `cuobjdump -sass`).",
             "Architectures in the fatbin: %s.  %d kernels." % (", ".join(sorted(archs)), len(names)), "",
             "Columns: registers per thread, stack bytes (non-zero = local-memory spill or indexed local array), static shared",
             "memory bytes, SASS instructions, then counts of the data-movement mnemonics.", "",
             "| kernel | regs | stack | smem | SASS | " + " | ".join(k for k, _ in MNEMONICS) + " |",
             "|---|---|---|---|---|" + "---|" * len(MNEMONICS)]
    spills = []
    for n in sorted(names, key=lambda k: short[k]):
        r, st, sh = res[n]
        c = counts.get(n, collections.Counter())
        lines.append("| `%s` | %d | %d | %d | %d | " % (short[n], r, st, sh, c["instr"])
                     + " | ".join(str(c[k]) if c[k] else "" for k, _ in MNEMONICS) + " |")
        if st:
            spills.append((short[n], st, c["LDL/STL"]))
    lines += ["", "## Kernels with a stack frame", ""]
    if spills:
        for s, st, n in spills:
            lines.append("* `%s`: %d bytes, %d LDL/STL instructions" % (s, st, n))
    else:
        lines.append("none")
    tot = collections.Counter()
    for c in counts.values():
        tot.update(c)
    lines += ["", "## Totals over the library", "",
              ", ".join("%s %d" % (k, tot[k]) for k in ["instr"] + [k for k, _ in MNEMONICS])]
    text = "\n".join(lines) + "\n"
    if out:
        with open(out, "w") as f:
            f.write(text)
    else:
        sys.stdout.write(text)


if __name__ == "__main__":
    main()
