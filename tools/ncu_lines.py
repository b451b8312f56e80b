#!/usr/bin/env python
"""Per-source-line view of an ncu `--page source --csv` dump (SASS rows, no line column) by aligning it with
`nvdisasm -g` of the cubin that holds the kernel (built with -lineinfo).

    python tools/ncu_lines.py <src.csv> <cubin> <mangled-name-substring> [top]

Prints, per source line: instructions executed (warp level), stall samples and the dominant stall reasons.
"""
import csv
import re
import subprocess
import sys
from collections import defaultdict


def sass_lines(cubin, key):
    out = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
    lines, cur, on, res = out.splitlines(), None, False, []
    for ln in lines:
        m = re.match(r"\s*\.text\.(\S+):", ln)
        if m:
            on = key in m.group(1)
            continue
        if not on:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", ln):
            res.append(cur)
    return res


def main():
    src, cubin, key = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    rows = list(csv.reader(open(src)))
    hdr = rows[1]
    body = [r for r in rows[2:] if len(r) == len(hdr)]
    lines = sass_lines(cubin, key)
    if len(lines) != len(body):
        print(f"warning: {len(lines)} SASS instructions in the cubin vs {len(body)} rows in the profile", file=sys.stderr)
    ix = {h: i for i, h in enumerate(hdr)}
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    agg = defaultdict(lambda: defaultdict(float))
    for ln, r in zip(lines, body):
        a = agg[ln]
        a["inst"] += float(r[ix["Instructions Executed"]] or 0)
        a["samples"] += float(r[ix["# Samples"]] or 0)
        for h in stall_cols:
            a[h] += float(r[ix[h]] or 0)
    tot_i = sum(a["inst"] for a in agg.values())
    tot_s = sum(a["samples"] for a in agg.values())
    print(f"total warp instructions {tot_i:.0f}, stall samples {tot_s:.0f}")
    print(f"{'file:line':28s} {'inst%':>6s} {'smp%':>6s}  top stalls")
    for ln, a in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:top]:
        st = sorted(((a[h], h[6:]) for h in stall_cols), reverse=True)[:3]
        name = f"{ln[0]}:{ln[1]}" if ln else "?"
        print(f"{name:28s} {100 * a['inst'] / tot_i:6.2f} {100 * a['samples'] / tot_s:6.2f}  " +
              ", ".join(f"{n} {100 * v / max(a['samples'], 1):.0f}%" for v, n in st if v > 0))


if __name__ == "__main__":
    main()
