"""Top stall-sample lines of a kernel from an .ncu-rep (source page, SASS view)."""
import csv, subprocess, sys
rep, kernel = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kernel], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
data = []
for r in rows[2:]:
    if len(r) < len(hdr): continue
    try: s = int(r[ci["# Samples"]])
    except: continue
    data.append((s, r))
tot = sum(s for s, _ in data) or 1
print("total samples", tot, "instructions", len(data))
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {h: 0 for h in stalls}
for s, r in data:
    for h in stalls:
        try: agg[h] += int(r[ci[h]])
        except: pass
print("stall totals:", {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v > tot * 0.02})
for s, r in sorted(data, key=lambda t: -t[0])[:top]:
    st = {h[6:]: int(r[ci[h]]) for h in stalls if r[ci[h]] not in ("", "0")}
    st = dict(sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print(f"{100*s/tot:5.1f}%  {r[ci['Source']][:70]:70s} {st}")
