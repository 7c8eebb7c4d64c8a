"""Dynamic (executed) instruction mix of a kernel from an .ncu-rep source page."""
import csv, subprocess, sys, collections
rep, kernel = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kernel], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]; ci = {h: i for i, h in enumerate(hdr)}
agg = collections.Counter(); tot = 0
for r in rows[2:]:
    if len(r) < len(hdr): continue
    try: n = int(r[ci["Instructions Executed"]])
    except: continue
    src = r[ci["Source"]].strip().split()
    op = src[1] if src and src[0].startswith("@") and len(src) > 1 else (src[0] if src else "?")
    agg[op.split(".")[0]] += n; tot += n
print("total warp-instructions executed:", tot)
for k, v in agg.most_common(28): print(f"{100*v/tot:5.1f}%  {v:>10d}  {k}")
