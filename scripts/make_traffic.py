"""Per-kernel DRAM bytes, warp instructions and shared-memory wavefronts per launch from an ncu launch list (long CSV:
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,
l1tex__data_pipe_lsu_wavefronts_mem_shared.sum --clock-control none --csv):
python scripts/make_traffic.py profiles/r2_launches_final.csv profiles/r2_traffic.json"""
import collections, csv, json, sys

src, dst = sys.argv[1], sys.argv[2]
rows = [r for r in csv.reader(open(src)) if len(r) > 5]
hdr = rows[0]
ik, im, iv, iu, iid = (hdr.index(c) for c in ("Kernel Name", "Metric Name", "Metric Value", "Metric Unit", "ID"))
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3}
acc = collections.OrderedDict()
for r in rows[1:]:
    name = r[ik].split("(")[0].replace("rdfe::", "").replace("void ", "").strip()
    a = acc.setdefault(name, {"ids": set(), "rd": 0.0, "wr": 0.0, "us": 0.0, "inst": 0.0, "smem": 0.0})
    a["ids"].add(r[iid])
    v = float(r[iv].replace(",", "")) * scale.get(r[iu], 1.0)
    if r[im] == "dram__bytes_read.sum": a["rd"] += v
    elif r[im] == "dram__bytes_write.sum": a["wr"] += v
    elif r[im] == "gpu__time_duration.sum": a["us"] += v
    elif r[im] == "smsp__inst_executed.sum": a["inst"] += v
    elif r[im] == "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum": a["smem"] += v
out = {"source": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,"
                 "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum --clock-control none (%s), "
                 "bench.py --steps 4 --warmup 3, 64 streams 752x480" % src, "kernels": {}}
for name, a in acc.items():
    n = len(a["ids"])
    out["kernels"][name] = {"launches": n, "dram_read_bytes_per_launch": a["rd"] / n, "dram_write_bytes_per_launch": a["wr"] / n,
                            "mean_us_under_ncu": a["us"] / n, "warp_inst_per_launch": a["inst"] / n,
                            "smem_wavefronts_per_launch": a["smem"] / n}
json.dump(out, open(dst, "w"), indent=1)
# shares of the serialised step: only the steady-state steps (from the first step that tracks, i.e. launches LK)
first_lk = min((int(i) for n, a in acc.items() if n.startswith("lk_track") for i in a["ids"]), default=0) - 9
step_us = collections.OrderedDict()
for r in rows[1:]:
    if r[im] == "gpu__time_duration.sum" and int(r[iid]) >= first_lk:
        name = r[ik].split("(")[0].replace("rdfe::", "").replace("void ", "").strip()
        if not name.startswith("at::"):
            step_us[name] = step_us.get(name, 0.0) + float(r[iv].replace(",", "")) * scale.get(r[iu], 1.0)
tot = sum(step_us.values())
out["share_of_serialised_step"] = {n: v / tot for n, v in step_us.items()}
json.dump(out, open(dst, "w"), indent=1)
for name, a in acc.items():
    if name.startswith("at::"): continue
    print("%-28s launches %3d  us/launch %7.1f  share %5.1f%%  dram MB/launch %7.1f" % (name, len(a["ids"]), a["us"] / len(a["ids"]), 100 * step_us.get(name, 0.0) / tot, (a["rd"] + a["wr"]) / len(a["ids"]) / 1e6))
