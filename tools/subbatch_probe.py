#!/usr/bin/env python
"""What would sub-batch pipelining buy?  Throughput of S engines of 4096 / S canvases each, every engine on its own stream
(kernels of different sub-batches may co-run), against one engine of 4096."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import attend_infer_repeat_b200 as air
from attend_infer_repeat_b200.cell import _init_flat

dev = torch.device("cuda", 0)
cfg = air.CellConfig(precision=air.AIR_PREC_TC_SPLIT)
T, Btot = 3, 4096
params, _ = _init_flat(air.param_spec(cfg), dev, seed=0)
prior = air.make_prior(dict(loc=0., scale=1.), dict(loc=0., scale=1.), dict(loc=0., scale=1.), 0.5, True)
full = len(sys.argv) > 1 and sys.argv[1] == "full"     # full: S engines of 4096 canvases each (two batches in flight)
for S in ((1, 2, 3, 4, 6, 8) if full else (1, 2, 4)):
    B = Btot if full else Btot // S
    engs = [air.Engine(cfg, B, T, device=dev) for _ in range(S)]
    for e in engs:
        e.cache_weights(True)
        e.set_launch_overlap(S == 1 or os.environ.get("PROBE_KEEP_PDL") is not None)   # what EnginePool does
    streams = [torch.cuda.Stream(device=dev) for _ in range(S)]
    data = [(torch.rand(B, 50, 50, device=dev), torch.randn(T, B, 4, device=dev), torch.randn(T, B, cfg.na, device=dev),
             torch.rand(T, B, 1, device=dev)) for _ in range(S)]
    def step():
        for e, st, d in zip(engs, streams, data):
            with torch.cuda.stream(st):
                e.forward(params, *d, prior)
    for _ in range(5):
        step()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 100
    ev0.record()
    for st in streams:
        st.wait_event(ev0)
    for _ in range(n):
        step()
    for st in streams:
        torch.cuda.current_stream().wait_stream(st)
    ev1.record()
    torch.cuda.synchronize()
    import time
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        step()
    host = (time.perf_counter() - t0) / n / (S if full else 1)
    torch.cuda.synchronize()
    print(f"S={S} x B={B}: {ev0.elapsed_time(ev1) / n / (S if full else 1):.4f} ms per 4096 canvases (host enqueue {host * 1e3:.4f} ms)")
    for e in engs:
        e.close()
