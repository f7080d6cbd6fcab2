"""Summarise a DVG_TRACE dump of lstm_step_kernel (stderr of a DVG_TC_TRACE=1 run)."""
import sys
import numpy as np

rows = []
for line in open(sys.argv[1]):
    if not line.startswith("cta"):
        continue
    toks = line.split(":", 1)[1].replace("|", " ").replace("#", "").split()
    rows.append([int(t) for t in toks])
a = np.array(rows)
names = {0: "start", 1: "setup", 26: "trigdone", 27: "xpack", 28: "it0 published", 29: "it0 stored", 30: "end", 31: "finalizer", 32: "it0 loopdone", 33: "it0 hpstored", 34: "it0 fenced", 35: "head tanh g0", 36: "head store g0", 37: "head tanh g1", 38: "head store g1", 39: "tail mask known", 40: "tail rsample done", 41: "tail all done", 42: "exit"}
for m in range(3):
    for k, n in enumerate(["pstart", "depok", "stage0", "mmaissued", "accready", "epidone", "item", "requested"]):
        names[2 + m * 8 + k] = f"it{m} {n}"
for m in range(3):
    for i in range(8):
        names[64 + m * 8 + i] = f"it{m} kb{i} full@mma"
        names[88 + m * 8 + i] = f"it{m} kb{i} issued"
for kb in range(4):
    for j, n in enumerate(["poll start", "relaxed ok", "acquire done", "proxy fence done"]):
        names[112 + kb * 4 + j] = f"dep kb{kb} {n}"
for s in range(a.shape[1]):
    col = a[:, s]
    v = col[col >= 0]
    if len(v) and s in names:
        print(f"{s:2d} {names[s]:16s} n={len(v):3d} min={v.min():6d} med={int(np.median(v)):6d} max={v.max():6d}")
if len(sys.argv) > 2:
    for r in a:
        print(" ".join(f"{x:6d}" for x in r))
