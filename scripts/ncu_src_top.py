"""Top stall-sample SASS lines from `ncu --page source --csv` output (first kernel table, or --table N)."""
import csv, sys
path = sys.argv[1]; table = int(sys.argv[2]) if len(sys.argv) > 2 else 0; frac = float(sys.argv[3]) if len(sys.argv) > 3 else 0.004
rows = list(csv.reader(open(path)))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
s = starts[table]; e = starts[table + 1] if table + 1 < len(starts) else len(rows)
print(rows[s][1][:120])
hdr = rows[s + 1]; data = [r for r in rows[s + 2:e] if len(r) > 6]
isrc = hdr.index("Source"); isamp = hdr.index("# Samples"); iex = hdr.index("Instructions Executed")
tot = sum(int(r[isamp] or 0) for r in data)
print("total samples", tot, "instrs", len(data))
for i, r in enumerate(data):
    sm = int(r[isamp] or 0); t = r[isrc]
    key = any(k in t for k in ("UTCHMMA", "UTCQMMA", "UTCBAR", "UTMALDG", "UTMASTG", "UBLKCP", "LDTM", "BAR.SYNC", "MUFU.SIN"))
    if sm > tot * frac or (key and sm > tot * 0.0005):
        print(f"{i:5d} {sm:7d} {100*sm/tot:5.1f}% ex={r[iex]:>10s} {t[:80]}")
