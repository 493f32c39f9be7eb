#!/usr/bin/env python
"""Print the timeline of a row_kernel launch dumped with AIR_ROW_TRACE=<prefix> (air_api.cu: launch_row_path).
    python tools/row_trace.py <prefix>.<seq>.bin [cta]
Unit stamps (MMA warp): 0 reached, 1 A / D waits done, 2 weights landed, 3 issued + committed.
Task stamps (epilogue thread 64): 0 reached, 1 first wait done (a_free for loads, d_full for epilogues), 2 accumulator
released (before the a_free wait), 3 done."""
import sys
import numpy as np

raw = open(sys.argv[1], "rb").read()
n_units, n_tasks, MAXU, MAXT = np.frombuffer(raw[:16], dtype=np.int32)
off = 16
units = np.frombuffer(raw[off:off + 16 * n_units], dtype=np.uint8).reshape(n_units, 16)
off += 16 * n_units
ttype = np.frombuffer(raw[off:off + n_tasks], dtype=np.uint8)
off += n_tasks
a = np.frombuffer(raw[off:], dtype=np.int64).reshape(-1, MAXU + MAXT, 4)
cta = int(sys.argv[2]) if len(sys.argv) > 2 else 0
t = a[cta]
t0 = t[t > 0].min()
rel = lambda x: int(x - t0) if x > 0 else -1
names = ["LOAD_HL", "LOAD_CROP", "ELU", "OUT", "WHAT", "WHERE"]
ev = []
for u in range(n_units):
    r = [rel(x) for x in t[u]]
    n16, ah, d, fl = units[u][10], units[u][11], units[u][12], units[u][13]
    ev.append((r[0], f"unit {u:3d} N{int(n16) * 16:3d} A{ah} D{d} fl{int(fl):02x}  reach {r[0]:7d}  ad_ok {r[1]:7d} (+{r[1] - r[0]:5d})  "
                     f"w_ok {r[2]:7d} (+{r[2] - r[1]:5d})  issued {r[3]:7d} (+{r[3] - r[2]:5d})"))
for k in range(n_tasks):
    r = [rel(x) for x in t[MAXU + k]]
    ev.append((r[0], f"    task {k:3d} {names[ttype[k]]:9s} reach {r[0]:7d}  wait1 {r[1]:7d} (+{r[1] - r[0]:5d})  rel {r[2]:7d}  "
                     f"done {r[3]:7d} (+{r[3] - max(r[1], r[0]):5d} work)"))
for _, line in sorted(ev):
    print(line)
span = a.reshape(a.shape[0], -1).max(axis=1) - np.where(a > 0, a, 1 << 62).reshape(a.shape[0], -1).min(axis=1)
print("kernel span per CTA (clocks): min", int(span.min()), "median", int(np.median(span)), "max", int(span.max()))
