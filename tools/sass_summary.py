#!/usr/bin/env python
"""Opcode histogram per kernel of libmanner_b200.so (cuobjdump -sass), written to profiles/sass_summary.txt so that a reader
need not disassemble the git-ignored .so.  Run after a build:  python tools/sass_summary.py"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "manner_b200", "lib", "libmanner_b200.so")
MARK = ("UTCHMMA", "UTCQMMA", "UTMALDG", "UTMASTG", "LDTM", "STTM", "UTCBAR", "UTCCP", "SYNCS", "HMMA", "LDGSTS", "LDG.E.128", "LDG.E.CONSTANT", "LDS.128",
        "STL", "LDL", "ATOM", "RED", "REDUX", "SHFL", "DFMA", "MUFU", "FFMA", "NANOSLEEP", "CCTL", "MEMBAR", "ERRBAR")


def main() -> None:
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    kernels = collections.OrderedDict()
    name = None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            kernels[name] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m and name:
            kernels[name][m.group(1)] += 1
    lines = [f"SASS summary of {os.path.relpath(LIB, ROOT)} (sm_100a; cuobjdump -sass; own kernels only, cub / thrust instantiations left out)", ""]
    for k, c in kernels.items():
        if "cub::" in k or "thrust::" in k:
            continue
        total = sum(c.values())
        short = re.sub(r"\(.*", "", k)
        lines.append(f"{short}   [{total} instructions]")
        marks = []
        for mk in MARK:
            n = sum(v for op, v in c.items() if op == mk or op.startswith(mk + "."))
            if n:
                marks.append(f"{mk}={n}")
        lines.append("    marks: " + (" ".join(marks) if marks else "-"))
        lines.append("    top:   " + " ".join(f"{op}={n}" for op, n in c.most_common(12)))
    path = os.path.join(ROOT, "profiles", "sass_summary.txt")
    with open(path, "w") as f:
        f.write("\n".join(lines) + "\n")
    print(path, len(kernels), "functions")


if __name__ == "__main__":
    main()
