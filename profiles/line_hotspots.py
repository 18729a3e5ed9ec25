"""Correlate an ncu `--page source --csv` SASS dump with nvdisasm line info:
per-source-line warp-instruction counts and stall samples for one kernel.

    python profiles/line_hotspots.py <ncu_source.csv> <nvdisasm -g -c output> <mangled kernel substr> [top]
"""
import csv
import re
import sys
from collections import defaultdict


def sass_lines(path, kernel):
    """[(source line or None)] per SASS instruction of `kernel`, in address order."""
    out, cur, inside = [], None, False
    for ln in open(path, errors="replace"):
        if ln.startswith(".text.") or re.match(r"\s*\.section\s+\.text\.", ln):
            inside = kernel in ln
            cur = None
            continue
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", ln):
            out.append(cur)
    return out


def main():
    ncu_csv, sass, kernel = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    rows = list(csv.reader(open(ncu_csv)))
    hdr = rows[1]
    ia, ii, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    inst = [(r[ia], int(r[ii]), int(r[isamp]), r[hdr.index("Source")]) for r in rows[2:] if len(r) > ii]
    lines = sass_lines(sass, kernel)
    print(f"ncu instructions: {len(inst)}, nvdisasm instructions: {len(lines)}")
    n = min(len(inst), len(lines))
    agg, samp = defaultdict(int), defaultdict(int)
    for k in range(n):
        agg[lines[k]] += inst[k][1]
        samp[lines[k]] += inst[k][2]
    tot, tots = sum(agg.values()), sum(samp.values())
    print(f"total warp instructions {tot}, samples {tots}")
    for key, v in sorted(agg.items(), key=lambda kv: -kv[1])[:top]:
        print(f"{str(key):38s} inst {v:10d} {v / tot:6.1%}   samples {samp[key]:6d} {samp[key] / max(tots, 1):6.1%}")


if __name__ == "__main__":
    main()
