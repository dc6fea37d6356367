"""Per-kernel table from an ncu launch list (the CSV written by
   ncu --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --clock-control none --csv --log-file X ...):
launches, total and longest duration, share of the listed time and -- when the DRAM counters are in the list -- DRAM GB/s
against the measured HBM peak of MEASURED_PEAKS.json.  Prints markdown.   python tools/launch_table.py X.csv [title]"""
import collections, csv, json, os, re, sys

path = sys.argv[1]
title = sys.argv[2] if len(sys.argv) > 2 else path
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
peak = 6447.8
pk = os.path.join(root, "MEASURED_PEAKS.json")
if os.path.exists(pk):
    peak = float(json.load(open(pk))["hbm_gbs"])
rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
hdr = rows[0]
ki, mi, vi, ii, ui = (hdr.index(k) for k in ("Kernel Name", "Metric Name", "Metric Value", "ID", "Metric Unit"))
per = collections.OrderedDict()
for r in rows[1:]:
    d = per.setdefault(r[ii], {"k": r[ki]})
    v = float(r[vi].replace(",", ""))
    if r[mi].startswith("dram"):
        v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[r[ui]]
    else:
        v *= {"ns": 1e-3, "us": 1, "usecond": 1, "ms": 1e3, "msecond": 1e3, "s": 1e6, "nsecond": 1e-3}.get(r[ui], 1)
    d[r[mi]] = v
agg = collections.OrderedDict()
for d in per.values():
    k = re.sub(r"\(.*", "", d["k"]).replace("void ", "")
    a = agg.setdefault(k, [0, 0.0, 0.0, 0.0])
    a[0] += 1
    a[1] += d.get("gpu__time_duration.sum", 0.0)
    a[2] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
    a[3] = max(a[3], d.get("gpu__time_duration.sum", 0.0))
tot = sum(a[1] for a in agg.values())
have_dram = any(a[2] > 0 for a in agg.values())
print(f"### {title}\n")
print(f"{len(per)} launches, {tot / 1e3:.2f} ms listed (per-launch times under ncu are cold-cache and serialised: compare shares).\n")
print("| kernel | launches | total us | share | longest us |" + (" DRAM GB/s | of HBM peak |" if have_dram else ""))
print("|---|---|---|---|---|" + ("---|---|" if have_dram else ""))
for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
    line = f"| `{k}` | {a[0]} | {a[1]:.1f} | {100 * a[1] / tot:.2f}% | {a[3]:.1f} |"
    if have_dram:
        g = a[2] / a[1] / 1e3 if a[1] > 0 else 0.0
        line += f" {g:.0f} | {100 * g / peak:.1f}% |"
    print(line)
