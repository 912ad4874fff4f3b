"""Summarise an ncu launch list (--metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --csv):
per kernel name: launches, total ms, share, DRAM GB read / written.  python scripts/ncu_launch_summary.py file.csv [--skip-torch]"""
import csv, sys, collections, re
path = sys.argv[1]
rows = []
with open(path, newline="") as f:
    lines = [l for l in f if l.startswith('"')]
rd = csv.DictReader(lines)
per = collections.OrderedDict()
for r in rd:
    per.setdefault(r["ID"], {"name": r["Kernel Name"], "grid": r["Grid Size"], "block": r["Block Size"]})[r["Metric Name"]] = (
        float(r["Metric Value"].replace(",", "")), r["Metric Unit"])
def ms(v):
    val, unit = v
    return val * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
def gb(v):
    val, unit = v
    return val * {"byte": 1e-9, "Kbyte": 1e-6, "Mbyte": 1e-3, "Gbyte": 1.0}.get(unit, 1e-9)
agg = collections.OrderedDict()
for k, v in per.items():
    name = re.sub(r"\(.*", "", v["name"]).replace("void ", "").replace("nc::", "")
    name = re.sub(r"<unnamed>::", "", name)
    if "--skip-torch" in sys.argv and ("at::" in v["name"] or "elementwise" in v["name"]):
        continue
    a = agg.setdefault(name, [0, 0.0, 0.0, 0.0])
    a[0] += 1
    a[1] += ms(v["gpu__time_duration.sum"])
    if "dram__bytes_read.sum" in v:
        a[2] += gb(v["dram__bytes_read.sum"]); a[3] += gb(v["dram__bytes_write.sum"])
tot = sum(a[1] for a in agg.values())
print(f"{'kernel':60s} {'n':>5s} {'ms':>10s} {'share':>7s} {'rd GB':>9s} {'wr GB':>9s} {'GB/s':>8s}")
for n, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{n[:60]:60s} {a[0]:5d} {a[1]:10.3f} {a[1]/tot:7.4f} {a[2]:9.3f} {a[3]:9.3f} {(a[2]+a[3])/max(a[1],1e-9)*1e3:8.1f}")
print(f"{'total':60s} {sum(a[0] for a in agg.values()):5d} {tot:10.3f} {1.0:7.4f} {sum(a[2] for a in agg.values()):9.3f} {sum(a[3] for a in agg.values()):9.3f}")
