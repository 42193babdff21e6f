"""Static SASS rows of one kernel attributed to the OUTERMOST source line (the call site inside the kernel body) and, optionally,
to the innermost inlined function line.  Developer tool for the instruction-footprint work (the step kernel is instruction-fetch bound).

    cuobjdump -xelf all gym_genesis_b200/libgg_b200.so ; python tools/sass_by_site.py gg_kernels.sm_100a.cubin step_kernelILi7ELi2ELb0 [--inner]
"""
import re
import subprocess
import sys
from collections import Counter

cubin, pat = sys.argv[1], sys.argv[2]
inner = "--inner" in sys.argv
out = subprocess.run(["nvdisasm", "-gi", "-c", cubin], capture_output=True, text=True).stdout
on = False
site = Counter()
chain = []
total = 0
pending = []
for line in out.splitlines():
    if line.startswith("//---") and ".text." in line:
        on = pat in line
        continue
    if not on:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', line)
    if m:
        pending.append((m.group(1).split("/")[-1], int(m.group(2))))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,6}\*/", line):
        if pending:
            chain = pending
            pending = []
        total += 1
        if chain:
            key = chain[0] if inner else chain[-1]
        else:
            key = ("?", 0)
        site[key] += 1
print("total SASS rows", total)
for (f, l), c in sorted(site.items(), key=lambda x: (x[0][0], x[0][1])):
    if c >= (5 if not inner else 15):
        print(f"{c:6d}  {f}:{l}")
