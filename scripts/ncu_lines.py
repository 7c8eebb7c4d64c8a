"""Per-source-line stall samples of a kernel from an .ncu-rep (needs -lineinfo + --import-source on)."""
import csv, subprocess, sys
rep, kernel = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda", "--csv", "--kernel-name", "regex:" + kernel],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if "# Samples" in r)
hdr = rows[hi]; ci = {h: i for i, h in enumerate(hdr)}
data = []
cur_file = ""
for r in rows[hi + 1:]:
    if len(r) < len(hdr):
        if r and r[0] in ("File Path", "File Name") and len(r) > 1: cur_file = r[1]
        continue
    try: s = int(r[ci["# Samples"]])
    except: continue
    data.append((s, r[ci.get("Line", 0)] if "Line" in ci else r[0], r[ci["Source"]], cur_file))
tot = sum(d[0] for d in data) or 1
print("total samples", tot)
for s, ln, src, f in sorted(data, key=lambda t: -t[0])[:top]:
    print(f"{100*s/tot:5.1f}%  {ln:>5} {src.strip()[:110]}")
