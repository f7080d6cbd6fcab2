"""Summarise the chained-launch dump of scripts/chain_trace.py (stderr log)."""
import sys

import numpy as np

names = {0: "start", 1: "setup/wait done", 27: "xpack done", 26: "trig delivered", 31: "mask published", 4: "it0 stage0@mma",
         6: "it0 acc ready", 7: "it0 epi done", 14: "it1 acc ready", 15: "it1 epi done", 22: "it2 acc ready", 23: "it2 epi done",
         30: "items done", 39: "tail mask known", 42: "exit"}
cur, launches = None, {}
for line in open(sys.argv[1]):
    if line.startswith("STEP TRACE chain launch"):
        cur = int(line.split()[-1])
        launches[cur] = []
    elif line.startswith("cta") and cur is not None:
        toks = line.split(":", 1)[1].replace("#", "").split()
        launches[cur].append([int(t) for t in toks])
for li, rows in launches.items():
    a = np.array(rows)
    print("launch", li)
    for s in sorted(names):
        v = a[:, s]
        v = v[v != -1]
        if len(v):
            print("  %2d %-18s n=%3d min=%7d p10=%7d med=%7d p90=%7d max=%7d" % (s, names[s], len(v), v.min(), np.percentile(v, 10), np.median(v), np.percentile(v, 90), v.max()))
