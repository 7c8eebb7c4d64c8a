"""Stall samples between consecutive barriers/branches of a kernel (SASS view) -- coarse phase attribution."""
import csv, subprocess, sys
rep, kernel = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kernel], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]; ci = {h: i for i, h in enumerate(hdr)}
seg, segs, tot, n = 0, [], 0, 0
first = None
for r in rows[2:]:
    if len(r) < len(hdr): continue
    try: s = int(r[ci["# Samples"]])
    except: continue
    src = r[ci["Source"]].strip()
    if first is None: first = src
    seg += s; tot += s; n += 1
    if "BAR.SYNC" in src:
        segs.append((seg, n, src[:40])); seg = 0
segs.append((seg, n, "END"))
print("total", tot)
for i, (s, n, src) in enumerate(segs):
    print(f"seg {i:2d}: {100*s/max(tot,1):5.1f}%  (ends at instr {n}: {src})")
