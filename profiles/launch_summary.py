"""Reduces an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` log of
`python bench.py --steps 1 --warmup W --no-e2e --no-cpu` to the timed step: one row per kernel name with launch
count, summed duration and DRAM bytes, plus traffic.json (what bench.py reports as roofline.traffic).
usage: python profiles/launch_summary.py <ncu.csv> <out.md> <out_step.csv>"""
import csv, json, sys, collections, os

src, out_md, out_csv = sys.argv[1:4]
rows = [r for r in csv.reader(open(src)) if len(r) >= 15 and r[0].isdigit()]
# one record per launch id
L = collections.OrderedDict()
for r in rows:
    d = L.setdefault(int(r[0]), {"name": r[4], "grid": r[8], "block": r[7]})
    d[r[12]] = float(r[14].replace(",", "")) * ({"msecond": 1e6, "usecond": 1e3, "nsecond": 1.0, "second": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}.get(r[13], 1.0))
ids = list(L)
enc = [i for i in ids if "encode_kernel" in L[i]["name"]]
start = enc[-1]
# the step begins with the capacity-offset scan launches right before the encode kernel
while start - 1 in L and ("scan_u" in L[start - 1]["name"] or "capacity" in L[start - 1]["name"] or "hoff" in L[start - 1]["name"]):
    start -= 1
step = [i for i in ids if i >= start]
agg = collections.OrderedDict()
for i in step:
    d = L[i]
    a = agg.setdefault(d["name"], [0, 0.0, 0.0, 0.0])
    a[0] += 1; a[1] += d.get("gpu__time_duration.sum", 0.0); a[2] += d.get("dram__bytes_read.sum", 0.0); a[3] += d.get("dram__bytes_write.sum", 0.0)
tot = sum(a[1] for a in agg.values())
with open(out_md, "w") as f:
    f.write("# ncu launch list of the timed step (`python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu`, 1 M x 15 kb reads)\n\n")
    f.write("Per-launch times under ncu are serialised and cold-cache; the SHARE of the step is what is compared with bench.py's stage times.\n\n")
    f.write("| kernel | launches | time ms | share | DRAM read MB | DRAM write MB |\n|---|---|---|---|---|---|\n")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write("| `%s` | %d | %.3f | %.1f%% | %.1f | %.1f |\n" % (k[:90], a[0], a[1] / 1e6, 100 * a[1] / tot, a[2] / 1e6, a[3] / 1e6))
    f.write("\ntotal %.3f ms in %d launches\n" % (tot / 1e6, len(step)))
with open(out_csv, "w") as f:
    w = csv.writer(f)
    w.writerow(["id", "kernel", "grid", "block", "time_ns", "dram_read_bytes", "dram_write_bytes"])
    for i in step:
        d = L[i]
        w.writerow([i, d["name"], d["grid"], d["block"], d.get("gpu__time_duration.sum", 0), d.get("dram__bytes_read.sum", 0), d.get("dram__bytes_write.sum", 0)])
ext = {k: a for k, a in agg.items() if any(x in k for x in ("encode_kernel", "scan_kernel", "kmerhash_kernel"))}
json.dump({"extract_dram_bytes_per_launch": sum(a[2] + a[3] for a in ext.values()),
           "per_kernel": {k: {"dram_read_bytes": a[2], "dram_write_bytes": a[3], "ns": a[1]} for k, a in ext.items()},
           "total_ns_under_ncu": sum(a[1] for a in ext.values()),
           "source": os.path.basename(out_csv) + ": ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none on the timed step of python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu (1 M x 15 kb reads)"},
          open(os.path.join(os.path.dirname(out_md), "traffic.json"), "w"), indent=1)
