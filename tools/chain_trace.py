#!/usr/bin/env python
"""Print the per-group timeline of a chain_kernel launch dumped with AIR_CHAIN_TRACE=<prefix> (see air_api.cu).
    python tools/chain_trace.py <prefix>.<seq>.bin [cta]
Stamps per group (SM clocks, relative to the CTA's first stamp):
  0 MMA warp saw a_ready   1 first weight tile landed   2 last weight tile landed   3 all MMAs issued
  4 epilogue: A operand written (before a_ready arrive)   5 epilogue: d_full seen (MMAs complete)"""
import sys
import numpy as np

a = np.fromfile(sys.argv[1], dtype=np.int64).reshape(-1, 32, 8)
ctas = [int(sys.argv[2])] if len(sys.argv) > 2 else [0, a.shape[0] // 2, a.shape[0] - 1]
for c in ctas:
    t = a[c]
    t0 = t[0][t[0] > 0].min()
    print(f"CTA {c}")
    prev5 = 0
    for g in range(32):
        if t[g].max() == 0:
            break
        r = [(int(x - t0) if x > 0 else -1) for x in t[g][:6]]
        print(f"  g{g:2d}  a_written {r[4]:7d}  a_seen {r[0]:7d}  w_first {r[1]:7d}  w_last {r[2]:7d}  issued {r[3]:7d}  "
              f"d_full {r[5]:7d} | epi+load {r[4] - prev5:6d}  mma {r[5] - r[0]:6d}")
        prev5 = r[5]
    span = a[:, :, :6].max(axis=(1, 2)) - np.where(a[:, 0, :6] > 0, a[:, 0, :6], 1 << 62).min(axis=1)
print("kernel span per CTA (clocks): min", int(span.min()), "median", int(np.median(span)), "max", int(span.max()))
