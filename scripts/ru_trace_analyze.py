"""Reads an NC_TRACE_RU dump (CTA 0's clock64() at role events of conv_ru_fused_kernel's first tiles) and prints, per tile,
each event relative to the tile's first A-load issue, plus steady-state per-tile periods."""
import sys
import numpy as np
NAMES = {0: "ld.issue0", 1: "ld.issueN", 20: "tr.begin", 2: "tr.raw0", 21: "tr.done", 3: "m1.begin", 4: "m1.acc1free", 5: "m1.a_full0",
         6: "m1.issued", 18: "e1.begin", 10: "e1.acc1full", 11: "e1.done", 7: "m2.acc2free", 8: "m2.h_full0", 9: "m2.issued",
         19: "e2.begin", 12: "e2.acc2full", 13: "e2.r_full0", 14: "e2.stored", 15: "w.tile", 16: "w.w1done", 17: "w.w2done"}
lines = open(sys.argv[1]).read().splitlines()
print(lines[0])
a = np.array([[int(x) for x in l.split()] for l in lines[1:]], dtype=np.int64)
t0 = a[a > 0].min()
lo, hi = int(sys.argv[2]) if len(sys.argv) > 2 else 20, int(sys.argv[3]) if len(sys.argv) > 3 else 28
order = [0, 1, 20, 2, 21, 3, 4, 5, 6, 18, 10, 11, 7, 8, 9, 19, 12, 13, 14, 15, 16, 17]
print("tile " + " ".join(f"{NAMES[e]:>11s}" for e in order))
for t in range(lo, hi):
    print(f"{t:4d} " + " ".join(f"{(a[t, e] - t0) if a[t, e] else -1:11d}" for e in order))
print("\nsteady-state period per tile (clk), tiles 16..80:")
for e in order:
    col = a[16:80, e]
    if (col > 0).all() and len(col) > 1:
        d = np.diff(col)
        print(f"  {NAMES[e]:>12s}: mean {d.mean():8.0f}  min {d.min():6d}  max {d.max():6d}")
print("\nmean gaps within a tile (clk), tiles 16..80:")
pairs = [(0, 2, "A load issue -> raw tile landed (TMA latency + queue)"), (2, 21, "transform of the tile (all chunks)"),
         (21, 5, "a_full -> MMA thread sees chunk 0 (tile t)"), (4, 6, "m1: acc1 free -> all k7 MMAs issued"),
         (6, 10, "m1 issued -> E1 sees acc1 full (MMA completion + commit)"), (18, 10, "E1 waiting for acc1"), (10, 11, "E1 drain (all groups)"),
         (11, 8, "E1 done -> m2 sees h_full0"), (7, 9, "m2: acc2 free -> issued"), (9, 12, "m2 issued -> E2 sees acc2 full"),
         (19, 12, "E2 waiting for acc2"), (12, 14, "E2 (all groups) until last store issued"), (3, 4, "m1 waiting for acc1 free"),
         (16, 17, "W: w2 tiles of the tile"), (15, 16, "W: w1 tiles of the next tile"),
         # conv_umma_kernel (no m2 / e1): the same event ids, one accumulator
         (6, 12, "MMAs issued -> epilogue sees acc full"), (3, 4, "MMA thread waiting for a free accumulator"),
         (4, 5, "MMA thread waiting for the first A chunk"), (12, 13, "epilogue: acc full -> first stage free / residual landed"),
         (12, 22, "epilogue g0: acc full -> tcgen05.ld done"), (22, 23, "epilogue g0: acc_empty arrive + bias"),
         (23, 13, "epilogue g0: waiting for the stage (e_free / residual landed)"), (13, 24, "epilogue g0: residual add + activation"),
         (24, 25, "epilogue g0: smem stores + proxy fence"), (25, 26, "epilogue g0: bar.sync of the 8 warps"),
         (26, 14, "epilogue: bar.sync(g0) -> last group's store issued + wait_group.read 1")]
for x, y, label in pairs:
    v = a[16:80, y] - a[16:80, x]
    ok = (a[16:80, y] > 0) & (a[16:80, x] > 0)
    if ok.any():
        print(f"  {label:62s}: mean {v[ok].mean():8.0f}  min {v[ok].min():7d}  max {v[ok].max():7d}")
