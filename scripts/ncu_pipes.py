"""Per-kernel summary of an `ncu --metrics ... --csv` launch list (long format): python scripts/ncu_pipes.py <csv> <steps>"""
import collections, csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
steps = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
hdr = rows[0]; ik = hdr.index("Kernel Name"); im = hdr.index("Metric Name"); iv = hdr.index("Metric Value"); iid = hdr.index("ID")
agg = collections.OrderedDict(); seen = set()
for r in rows[1:]:
    k = r[ik].split('(')[0].replace('rdfe::', '').replace('void ', '')[:28]
    a = agg.setdefault(k, collections.Counter())
    a[r[im]] += float(r[iv].replace(',', ''))
    if (r[iid], k) not in seen:
        seen.add((r[iid], k)); a['n'] += 1
tot = 0
for k, a in agg.items():
    n = a['n']
    mi = a['smsp__inst_executed.sum'] / 1e6
    tot += mi
    print(f"{k:28s} launches {int(n):3d}  us/launch {a['gpu__time_duration.sum']/1e3/n:7.1f}  Minst/launch {mi/n:6.2f}  "
          f"issue% {a['smsp__issue_active.avg.pct_of_peak_sustained_active']/n:5.1f}  warps% {a['sm__warps_active.avg.pct_of_peak_sustained_active']/n:5.1f}")
print(f"total Minst {tot:.1f} over {steps} steps = {tot/steps:.1f} per step")
