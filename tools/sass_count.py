"""Static SASS instruction count per kernel of a built library (developer tool): python tools/sass_count.py lib.so"""
import re, subprocess, sys
out = subprocess.run(["cuobjdump", "-sass", sys.argv[1]], capture_output=True, text=True).stdout
name, cnt, ops = None, {}, {}
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = m.group(1); cnt[name] = 0; ops[name] = {}
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
    if m and name:
        cnt[name] += 1
        op = m.group(2).split(".")[0]
        ops[name][op] = ops[name].get(op, 0) + 1
for n, c in sorted(cnt.items(), key=lambda x: -x[1]):
    top = sorted(ops[n].items(), key=lambda x: -x[1])[:12]
    print(c, n[-60:], " ".join(f"{k}:{v}" for k, v in top))
