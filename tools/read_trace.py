#!/usr/bin/env python
"""Summarise an AIR_READ_TRACE dump: per-phase durations of the where_read CTAs."""
import struct
import sys
import numpy as np

raw = open(sys.argv[1], "rb").read()
B, _, ns, _ = struct.unpack("4i", raw[:16])
t = np.frombuffer(raw[16:], dtype=np.int64).reshape(-1, 8)[:B]
t0 = t[:, 0].min()
print(f"B={B}; kernel span {(t[:, 4].max() - t0) / 1e3:.1f} us")
for i, nm in enumerate(["where codes (+ presence scan)", "tap tables", "image copy wait", "gather + split + store"]):
    d = (t[:, i + 1] - t[:, i]) / 1e3
    print(f"  {nm:30s} mean {d.mean():6.2f} us  p10 {np.percentile(d, 10):6.2f}  p90 {np.percentile(d, 90):6.2f}")
life = (t[:, 4] - t[:, 0]) / 1e3
print(f"CTA lifetime mean {life.mean():.2f} us, p90 {np.percentile(life, 90):.2f}")
start = (t[:, 0] - t0) / 1e3
for lo in range(0, 60, 5):
    m = (start >= lo) & (start < lo + 5)
    if m.any():
        print(f"  CTAs starting in [{lo},{lo + 5}) us: {int(m.sum()):5d}, lifetime {life[m].mean():.2f}")
