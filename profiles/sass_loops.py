"""Backward branches (loops) of one kernel in a cuobjdump -sass dump, with static instruction counts and
local-memory (spill) traffic inside each loop: python profiles/sass_loops.py lib.so <mangled-substring>"""
import re
import subprocess
import sys

lib, pat = sys.argv[1], sys.argv[2]
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
blocks = re.split(r"\n\s*Function : ", txt)
for b in blocks:
    name = b.split("\n", 1)[0]
    if pat not in name:
        continue
    ins = [(int(m.group(1), 16), m.group(2)) for m in re.finditer(r"/\*([0-9a-f]{4,5})\*/\s+(.*?);", b)]
    print(name, len(ins), "instructions")
    for a, t in ins:
        m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d+,\s*)?(0x[0-9a-f]+)", t)
        if m and int(m.group(1), 16) < a:
            lo = int(m.group(1), 16)
            body = [x for x in ins if lo <= x[0] <= a]
            spill = sum(1 for x in body if re.search(r"\b(LDL|STL)\b", x[1]))
            stg = sum(1 for x in body if "STG" in x[1])
            print(f"  loop {lo:#x}..{a:#x}: {len(body)} static instructions, {spill} LDL/STL, {stg} STG")
