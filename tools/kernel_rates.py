"""ncu csv (--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv) -> per-kernel achieved
DRAM rate:  python tools/kernel_rates.py launches.csv [label] [peak_GBs]  (prints JSON; committed under profiles/)"""
import collections
import csv
import json
import sys

path = sys.argv[1]
label = sys.argv[2] if len(sys.argv) > 2 else path
peak = float(sys.argv[3]) if len(sys.argv) > 3 else 6541.8
rows = list(csv.DictReader(l for l in open(path) if not l.startswith("==")))
per = collections.OrderedDict()
for r in rows:
    key = (r["ID"], r["Kernel Name"])
    v = float(r["Metric Value"].replace(",", ""))
    u = r["Metric Unit"]
    name = r["Metric Name"]
    if name == "gpu__time_duration.sum":
        v = v * {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}.get(u, 1e-9)
    else:
        v = v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
    per.setdefault(key, {})[name] = v
agg = collections.OrderedDict()
for (_, kname), mtr in per.items():
    short = kname.split("(")[0].replace("ipm::", "").replace("void ", "").replace("<unnamed>::", "")
    a = agg.setdefault(short, dict(launches=0, seconds=0.0, dram_bytes=0.0, best_gbs=0.0))
    t = mtr.get("gpu__time_duration.sum", 0.0)
    by = mtr.get("dram__bytes_read.sum", 0.0) + mtr.get("dram__bytes_write.sum", 0.0)
    a["launches"] += 1
    a["seconds"] += t
    a["dram_bytes"] += by
    if t > 0:
        a["best_gbs"] = max(a["best_gbs"], by / t * 1e-9)
out = dict(label=label, peak_gbs=peak, how="ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum "
           "--clock-control none (serialised, cold cache per launch)", kernels={})
for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["seconds"]):
    if a["seconds"] <= 0:
        continue
    gbs = a["dram_bytes"] / a["seconds"] * 1e-9
    out["kernels"][k] = dict(launches=a["launches"], avg_us=a["seconds"] / a["launches"] * 1e6,
                             dram_mb_per_launch=a["dram_bytes"] / a["launches"] * 1e-6, achieved_gbs=gbs,
                             frac_of_peak=gbs / peak, best_launch_gbs=a["best_gbs"])
print(json.dumps(out, indent=1))
