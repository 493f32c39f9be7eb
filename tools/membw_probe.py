#!/usr/bin/env python
"""Write / copy bandwidth of plain torch ops at the paint kernel's sizes (is 123 MB of canvas stores the floor?)."""
import torch
dev = torch.device("cuda", 0)
def t(f, n=20):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
for mb in (41, 123, 492):
    x = torch.empty(mb * 250000, device=dev); y = torch.empty_like(x)
    tz = t(lambda: x.zero_()); tc = t(lambda: y.copy_(x))
    print(f"{mb} MB: fill {tz:.1f} us = {mb / tz * 1e-3:.2f} TB/s written; copy {tc:.1f} us = {2 * mb / tc * 1e-3:.2f} TB/s moved")
