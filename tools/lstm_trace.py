#!/usr/bin/env python
"""Print the timeline of a lstm_cluster_kernel launch dumped with AIR_LSTM_TRACE=<prefix>: python tools/lstm_trace.py <file> [cta]
Epilogue stamps (thread 64): 0 start, 1 e operand written, 2 gx accumulator seen, 3 gx stored + h_init written,
per step t: 4+5t accumulator seen, 5+5t gate math done, 6+5t outputs + own slab written, 7+5t peers' slabs visible,
8+5t operand of the next step written; 30 end.  MMA warp: 32+2g operand seen, 33+2g group issued (g = 0: gx, 1..T: steps)."""
import sys
import numpy as np
a = np.fromfile(sys.argv[1], dtype=np.int64).reshape(-1, 64)
cta = int(sys.argv[2]) if len(sys.argv) > 2 else 0
names = {0: "start", 1: "e written", 2: "gx acc seen", 3: "gx stored, h_init written", 30: "end"}
for t in range(5):
    names.update({4 + 5 * t: f"t{t} acc seen", 5 + 5 * t: f"t{t} gates done", 6 + 5 * t: f"t{t} slab written",
                  7 + 5 * t: f"t{t} peers visible", 8 + 5 * t: f"t{t} next operand written"})
for g in range(6):
    names.update({32 + 2 * g: f"  mma g{g} operand seen", 33 + 2 * g: f"  mma g{g} issued"})
for c in (cta,):
    t = a[c]
    t0 = t[t > 0].min()
    ev = sorted((int(v - t0), names.get(i, str(i))) for i, v in enumerate(t) if v > 0)
    prev = 0
    for v, n in ev:
        print(f"{v:8d} (+{v - prev:6d})  {n}")
        prev = v
span = a.max(axis=1) - np.where(a > 0, a, 1 << 62).min(axis=1)
print("span per CTA: min", int(span.min()), "median", int(np.median(span)), "max", int(span.max()))
