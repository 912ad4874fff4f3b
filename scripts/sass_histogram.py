"""Tensor-core / TMEM / TMA / barrier opcode histogram of the built library: cuobjdump -sass lib.so | python scripts/sass_histogram.py"""
import sys, re, collections, subprocess
fn = None; per = collections.OrderedDict(); tot = collections.Counter()
KEY = ("UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTMAPF", "SYNCS", "UCGABAR", "MEMBAR", "ERRBAR",
       "FENCE", "MUFU", "BAR.SYNC", "HMMA", "IMMA")
FULL = ("UTCHMMA", "UTMALDG", "UTMASTG", "UTCBAR", "LDTM", "STTM", "UBLKCP")
for line in sys.stdin:
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = m.group(1); per[fn] = collections.Counter(); continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if not (m and fn):
        continue
    op = m.group(1)
    for k in KEY:
        if op.startswith(k):
            base = op if k in FULL else k
            per[fn][base] += 1; tot[base] += 1
            break
print("# SASS opcode histogram of neuralcodecs_b200/libneuralcodecs_cuda.so (cuobjdump -sass, sm_100a); tensor-core / TMEM / TMA / barrier opcodes only.")
print("# UTCHMMA = tcgen05.mma (.2CTA = cta_group::2), LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG = TMA tensor load / store (.2CTA = cta_group::2 form),")
print("# UBLKCP = cp.async.bulk (1-D), UTCBAR = tcgen05.commit (.MULTICAST = to both CTAs of a pair), SYNCS = mbarrier ops, UCGABAR = cluster barrier.")
print("# HMMA / IMMA (mma.sync) would show up here if any kernel used the legacy tensor-core path: none does.")
print("total:", dict(sorted(tot.items())))
seen = collections.OrderedDict()
for f, c in per.items():
    if not c:
        continue
    name = subprocess.run(["c++filt", f], capture_output=True, text=True).stdout.strip().split("(")[0]
    a = seen.setdefault(re.sub(r"<.*", "", name).replace("void ", ""), [0, collections.Counter()])
    a[0] += 1; a[1].update(c)
for k, (n, c) in seen.items():
    if any(x.startswith(("UTCHMMA", "LDTM", "UTMALDG", "UBLKCP")) for x in c):
        print(f"{k} ({n} instantiations):", dict(sorted(c.items())))
