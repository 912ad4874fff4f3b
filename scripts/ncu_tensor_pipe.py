"""Time-weighted tensor-pipe activity of the tcgen05 launches of one forward, from an ncu metrics list
(--metrics gpu__time_duration.sum,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active,
dram__bytes_read.sum,dram__bytes_write.sum --csv).  Writes the per-kernel table (stdout) and a JSON summary.
usage: python scripts/ncu_tensor_pipe.py launches.csv out.json "<what was profiled>" """
import csv, sys, json, re, collections
path, out_json, what = sys.argv[1], sys.argv[2], sys.argv[3]
lines = [l for l in open(path, newline="") if l.startswith('"')]
per = collections.OrderedDict()
for r in csv.DictReader(lines):
    d = per.setdefault(r["ID"], {"name": r["Kernel Name"]})
    d[r["Metric Name"]] = (float(r["Metric Value"].replace(",", "")), r["Metric Unit"])
def ms(v): return v[0] * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(v[1], 1e-6)
def gb(v): return v[0] * {"byte": 1e-9, "Kbyte": 1e-6, "Mbyte": 1e-3, "Gbyte": 1.0}.get(v[1], 1e-9)
TP = "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active"
agg = collections.OrderedDict()
for d in per.values():
    name = re.sub(r"\(.*", "", d["name"]).replace("void ", "").replace("nc::", "").replace("(anonymous namespace)::", "")
    a = agg.setdefault(name, {"n": 0, "ms": 0.0, "tp_ms": 0.0, "gb": 0.0})
    t = ms(d["gpu__time_duration.sum"])
    a["n"] += 1; a["ms"] += t; a["tp_ms"] += t * d.get(TP, (0.0, "%"))[0]
    if "dram__bytes_read.sum" in d: a["gb"] += gb(d["dram__bytes_read.sum"]) + gb(d["dram__bytes_write.sum"])
tot = sum(a["ms"] for a in agg.values())
tc = {k: a for k, a in agg.items() if k.startswith(("conv_umma", "conv_ru_fused", "conv_h16"))}
tc_ms = sum(a["ms"] for a in tc.values()); tc_tp = sum(a["tp_ms"] for a in tc.values())
print(f"# {what}")
print(f"{'kernel':48s} {'n':>4s} {'ms':>9s} {'share':>6s} {'tensor pipe %':>13s} {'DRAM GB':>8s} {'GB/s':>7s}")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["ms"]):
    print(f"{k[:48]:48s} {a['n']:4d} {a['ms']:9.3f} {a['ms']/tot:6.3f} {a['tp_ms']/max(a['ms'],1e-9):13.1f} {a['gb']:8.2f} {a['gb']/max(a['ms'],1e-9)*1e3:7.0f}")
print(f"tcgen05 launches: {sum(a['n'] for a in tc.values())}, {tc_ms:.3f} ms = {tc_ms/tot:.3f} of the forward, time-weighted tensor pipe {tc_tp/max(tc_ms,1e-9):.1f} %")
json.dump({"tensor_pipe_pct_time_weighted": round(tc_tp / max(tc_ms, 1e-9), 2), "tcgen05_share_of_forward": round(tc_ms / tot, 4),
           "per_kernel": {k: {"launches": a["n"], "ms": round(a["ms"], 3), "tensor_pipe_pct": round(a["tp_ms"] / max(a["ms"], 1e-9), 2)} for k, a in tc.items()},
           "note": f"ncu {TP}, time-weighted over every tcgen05 launch of {what} (--clock-control none; profiler run, not a bench value)"},
          open(out_json, "w"), indent=1)
