"""Aggregate an ncu report's source page (cuda,sass view) per enclosing source function: executed warp-instructions and stall samples."""
import collections, csv, re, subprocess, sys, os
rep = sys.argv[1]
nwarps = float(sys.argv[2]) if len(sys.argv) > 2 else 512.0
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
def func_map(path):
    fl = []
    for i, l in enumerate(open(path), 1):
        if l.startswith('GG_HD') or l.startswith('__global__') or l.startswith('template <int NU> __device__') or l.startswith('static'):
            m = re.search(r'(\w+)\(', l)
            if m and m.group(1) not in ('__launch_bounds__',):
                fl.append((i, m.group(1)))
            elif m:
                m2 = re.search(r'\)\s*(\w+)\(', l)
                fl.append((i, m2.group(1) if m2 else 'kernel'))
    return fl
maps = {f: func_map(os.path.join(root, 'gym_genesis_b200/csrc', f)) for f in ('gg_env.cuh', 'gg_math.cuh', 'gg_kernels.cu')}
cur = None; hdr = None
fagg = collections.defaultdict(lambda: [0, 0]); stall = collections.Counter()
for r in rows:
    if r and r[0] == "File Path": cur = r[1].split('/')[-1]; continue
    if r and r[0] == "Line No":
        hdr = r; si = hdr.index("# Samples"); ii = hdr.index("Instructions Executed")
        sc = [(i, h) for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
        continue
    if hdr is None or len(r) <= ii: continue
    try: ln = int(r[0]); s = int(r[si] or 0); ex = int(r[ii] or 0)
    except ValueError: continue
    name = '?'
    for i, n in maps.get(cur, []):
        if i <= ln: name = n
    fagg[(cur, name)][0] += s; fagg[(cur, name)][1] += ex
    for i, h in sc:
        try: stall[h] += int(r[i] or 0)
        except ValueError: pass
ts = sum(v[0] for v in fagg.values()); te = sum(v[1] for v in fagg.values())
print(f"executed warp-instructions {te}  per warp {te / nwarps:.0f}   stall samples {ts}")
print("stalls:", ", ".join(f"{h[6:]} {100 * v / max(ts, 1):.1f}%" for h, v in stall.most_common(7)))
for k, v in sorted(fagg.items(), key=lambda x: -x[1][1])[:22]:
    print(f"{v[1] / nwarps:8.0f} instr/warp {100 * v[1] / te:5.1f}%   samples {100 * v[0] / max(ts, 1):5.1f}%   {k[0]}:{k[1]}")
