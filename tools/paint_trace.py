#!/usr/bin/env python
"""Summarise an AIR_PAINT_TRACE dump: per-phase durations of the paint CTAs, CTA lifetimes, waves per SM."""
import struct
import sys
import numpy as np

raw = open(sys.argv[1], "rb").read()
B, n_prior, ns, _ = struct.unpack("4i", raw[:16])
t = np.frombuffer(raw[16:], dtype=np.int64).reshape(-1, 8)[: B + n_prior]
t0 = t[:, 0][t[:, 0] > 0].min()
prior, paint = t[:n_prior], t[n_prior:]
print(f"B={B}, {n_prior} prior CTAs; kernel span {(t[:, 6].max() - t0) / 1e3:.1f} us")
print(f"prior CTAs: start {((prior[:, 0] - t0) / 1e3).mean():.1f} us (max {((prior[:, 0] - t0) / 1e3).max():.1f}), "
      f"lifetime {((prior[:, 6] - prior[:, 0]) / 1e3).mean():.2f} us (max {((prior[:, 6] - prior[:, 0]) / 1e3).max():.2f})")
names = ["where load + inverse", "tap tables", "glimpse copy wait", "viz + column pass", "row pass", "block sum"]
for i, nm in enumerate(names):
    d = (paint[:, i + 1] - paint[:, i]) / 1e3
    print(f"  {nm:22s} mean {d.mean():6.2f} us  p10 {np.percentile(d, 10):6.2f}  p90 {np.percentile(d, 90):6.2f}")
life = (paint[:, 6] - paint[:, 0]) / 1e3
print(f"paint CTA lifetime mean {life.mean():.2f} us, p90 {np.percentile(life, 90):.2f}")
start = (paint[:, 0] - t0) / 1e3
for lo in range(0, 80, 10):
    m = (start >= lo) & (start < lo + 10)
    if m.any():
        print(f"  CTAs starting in [{lo},{lo + 10}) us: {int(m.sum()):5d}, lifetime {life[m].mean():.2f}")
sm = paint[:, 7]
print(f"SMs used {len(np.unique(sm))}, CTAs per SM min {np.bincount(sm).min()} max {np.bincount(sm).max()}")
