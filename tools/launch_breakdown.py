"""Per-kernel shares of the timed step from an ncu launch list (`--metrics gpu__time_duration.sum,...  --csv`).
usage: python tools/launch_breakdown.py gpurun_out/x_launches.csv   (second half of the launches = the timed step)"""
import collections, csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
h = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr = rows[h]
ki, mi, vi, ii = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
d = collections.OrderedDict()
for r in rows[h + 1:]:
    if len(r) > vi:
        d.setdefault(r[ii], {"k": r[ki]})[r[mi]] = float(r[vi].replace(",", ""))
ids = list(d)[len(d) // 2:]
agg = collections.OrderedDict()
for i in ids:
    k = re.sub(r"\(.*", "", d[i]["k"]).replace("void ", "").replace("<unnamed>::", "")[:58]
    a = agg.setdefault(k, [0, 0.0, 0.0, 0.0])
    a[0] += 1
    a[1] += d[i]["gpu__time_duration.sum"]
    a[2] = max(a[2], d[i].get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0))
    a[3] = max(a[3], d[i].get("dram__throughput.avg.pct_of_peak_sustained_elapsed", 0))
tot = sum(a[1] for a in agg.values())
print("| kernel | launches | us | share | tensor-pipe active % (max) | DRAM % of peak (max) |\n|---|---|---|---|---|---|")
for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"| `{k}` | {a[0]} | {a[1] / 1e3:.0f} | {100 * a[1] / tot:.1f} % | {a[2]:.1f} | {a[3]:.1f} |")
print(f"| **total** | {len(ids)} | **{tot / 1e3:.0f}** | | | |")
