"""Opcode histogram of a SASS address range: python profiles/sass_hist.py file.sass 0x9a0 0x3650"""
import collections
import re
import sys

path, lo, hi = sys.argv[1], int(sys.argv[2], 16), int(sys.argv[3], 16)
c = collections.Counter()
for ln in open(path):
    m = re.match(r"\s*/\*([0-9a-f]{4,5})\*/\s+(.*?);", ln)
    if not m:
        continue
    a = int(m.group(1), 16)
    if lo <= a < hi:
        t = m.group(2).split()
        op = t[1] if t[0].startswith("@") else t[0]
        c[op.split(".")[0]] += 1
for k, v in c.most_common():
    print(f"{v:5d} {k}")
print(sum(c.values()), "static instructions in range")
