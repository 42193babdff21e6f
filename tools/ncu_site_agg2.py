"""Like ncu_site_agg.py but groups by the line at inline depth `--depth D` counted from the kernel body (0 = outermost)."""
import csv, re, subprocess, sys
from collections import Counter, defaultdict
src_csv, cubin, pat = sys.argv[1:4]
depth = int(sys.argv[sys.argv.index("--depth") + 1]) if "--depth" in sys.argv else 1
minsamp = int(sys.argv[sys.argv.index("--min") + 1]) if "--min" in sys.argv else 10
out = subprocess.run(["nvdisasm", "-gi", "-c", cubin], capture_output=True, text=True).stdout
on = False; keys = []; chain = []; pending = []
for line in out.splitlines():
    if line.startswith("//---") and ".text." in line:
        on = pat in line; continue
    if not on: continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', line)
    if m:
        pending.append((m.group(1).split("/")[-1], int(m.group(2)))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,6}\*/", line):
        if pending: chain, pending = pending, []
        if chain:
            c = list(reversed(chain))          # outermost first
            keys.append(tuple(c[: depth + 1]))
        else:
            keys.append((("?", 0),))
rows = list(csv.reader(open(src_csv))); hdr = rows[1]; body = rows[2:]
assert len(body) == len(keys), (len(body), len(keys))
ci = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = defaultdict(lambda: {"ex": 0, "samp": 0, "static": 0, "st": Counter()})
tot_ex = tot_s = 0
for k, r in zip(keys, body):
    a = agg[k]; ex = int(r[ci["Instructions Executed"]]); s = int(r[ci["# Samples"]])
    a["ex"] += ex; a["samp"] += s; a["static"] += 1; tot_ex += ex; tot_s += s
    for h in stall_cols:
        v = int(r[ci[h]])
        if v: a["st"][h[6:]] += v
print(f"total executed warp-instr {tot_ex}  samples {tot_s}  static rows {len(keys)}")
for k, a in sorted(agg.items()):
    if a["samp"] < minsamp: continue
    name = " < ".join(f"{f.replace('gg_','').replace('.cuh','').replace('.cu','')}:{l}" for f, l in k)
    top = " ".join(f"{n}:{v}" for n, v in a["st"].most_common(4))
    print(f"{name:44s} {a['static']:6d} {a['ex']:9d} {a['samp']:6d} {100 * a['samp'] / tot_s:5.1f}  {top}")
