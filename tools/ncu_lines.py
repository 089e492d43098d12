#!/usr/bin/env python
"""Attribute the per-instruction counters of an ncu SASS source page to CUDA source lines.

    ncu -i rep.ncu-rep --page source --csv --kernel-name regex:K > sass.csv
    python tools/ncu_lines.py sass.csv <lib.so> <mangled-kernel-substring> [top]

The SASS page carries no line numbers; `nvdisasm -g` of the same cubin does, and lists the same
instructions in the same order, so the two are zipped by position.
"""
import csv
import os
import re
import subprocess
import sys
import tempfile


def sass_lines(lib, kernel):
    d = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=d, check=True, capture_output=True)
    cubin = [os.path.join(d, f) for f in os.listdir(d) if f.endswith(".cubin")][0]
    txt = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
    out, cur, inside = [], ("?", 0), False
    for ln in txt:
        if ln.startswith(".text."):
            inside = kernel in ln
            continue
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
            out.append(cur)
    return out


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    lib, kernel = sys.argv[2], sys.argv[3]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    hdr = rows[1]
    ci, sm, src = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Source")
    inst = [(int(r[ci]), int(r[sm]), r[src]) for r in rows[2:] if len(r) > ci and r[ci].isdigit()]
    lines = sass_lines(lib, kernel)
    if len(lines) != len(inst):
        print(f"warning: {len(inst)} profiled instructions vs {len(lines)} disassembled", file=sys.stderr)
    agg = {}
    for (n, s, _), loc in zip(inst, lines):
        a = agg.setdefault(loc, [0, 0])
        a[0] += n
        a[1] += s
    tot = sum(a[0] for a in agg.values())
    smp = sum(a[1] for a in agg.values())
    print("total warp instructions", tot, "samples", smp)
    srcs = {}
    for (f, l), (n, s) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        if f not in srcs:
            p = os.path.join(os.path.dirname(os.path.abspath(lib)), "csrc", f)
            srcs[f] = open(p).read().splitlines() if os.path.exists(p) else []
        text = srcs[f][l - 1].strip()[:90] if 0 < l <= len(srcs[f]) else ""
        print(f"{n:>9} {100 * n / tot:5.1f}%  smp {100 * s / max(smp, 1):5.1f}%  {f}:{l}: {text}")


if __name__ == "__main__":
    main()
