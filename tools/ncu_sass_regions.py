"""Split an ncu report's SASS view of the role-specialised step kernel into regions delimited by BAR / EXIT instructions and
print executed warp-instructions and stall samples per region (per block = per 32 envs).  usage: ncu_sass_regions.py rep nblocks"""
import csv, subprocess, sys
rep, nblocks = sys.argv[1], float(sys.argv[2])
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]; si = hdr.index("# Samples"); ii = hdr.index("Instructions Executed")
acc_i = acc_s = 0; n = 0; start = 0
tot_i = sum(int(r[ii] or 0) for r in rows[2:] if len(r) > ii)
for k, r in enumerate(rows[2:]):
    if len(r) <= ii: continue
    ex = int(r[ii] or 0); s = int(r[si] or 0)
    acc_i += ex; acc_s += s; n += 1
    src = r[1].strip()
    if "BAR." in src or src.startswith("EXIT") or " EXIT" in src and "@" not in src:
        if acc_i:
            print(f"sass rows {start:6d}-{k:6d}  {acc_i / nblocks:8.0f} warp-instr/block  samples {acc_s:5d}   ends with: {src[:50]}")
        acc_i = acc_s = 0; start = k + 1
print(f"tail rows {start}-  {acc_i / nblocks:8.0f}   total {tot_i / nblocks:.0f} per block")
