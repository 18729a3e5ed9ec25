"""Top SASS instructions by stall samples, with their source lines.
    python profiles/top_stalls.py <ncu --page source --csv dump> <nvdisasm -g -c dump> <kernel substr> [n]"""
import csv
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import line_hotspots as lh  # noqa: E402


def main():
    src, sass, kern = sys.argv[1:4]
    n = int(sys.argv[4]) if len(sys.argv) > 4 else 25
    rows = list(csv.reader(open(src)))
    hdr = rows[1]
    ii, isamp, isrc = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Source")
    inst = [(int(r[ii]), int(r[isamp]), r[isrc]) for r in rows[2:] if len(r) > ii]
    lines = lh.sass_lines(sass, kern)
    tot = sum(x[1] for x in inst)
    print(f"{len(inst)} SASS instructions, {sum(x[0] for x in inst)} executed, {tot} samples")
    for k in sorted(range(len(inst)), key=lambda k: -inst[k][1])[:n]:
        ln = lines[k] if k < len(lines) else None
        print(f"{inst[k][1]:6d} {inst[k][1] / max(tot, 1):5.1%} exec={inst[k][0]:9d} {str(ln):34s} {inst[k][2][:64]}")


if __name__ == "__main__":
    main()
