#!/usr/bin/env python
"""Round-2 diagnostic: per-stage CUDA-event times of one B=4096 forward pass and the per-group SM-clock timeline of the
two chain_kernel launches (AIR_CHAIN_TRACE).  Writes gpurun_out/r02_diag.txt."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
out_dir = os.path.join(ROOT, "gpurun_out")
os.makedirs(out_dir, exist_ok=True)
trace = len(sys.argv) > 1 and sys.argv[1] == "trace"
if trace:
    os.environ["AIR_CHAIN_TRACE"] = os.path.join(out_dir, "ctrace")
    os.environ["AIR_ROW_TRACE"] = os.path.join(out_dir, "rtrace")
    os.environ["AIR_LSTM_TRACE"] = os.path.join(out_dir, "ltrace")
    os.environ["AIR_PAINT_TRACE"] = os.path.join(out_dir, "ptrace")
    os.environ["AIR_READ_TRACE"] = os.path.join(out_dir, "wtrace")

import torch  # noqa: E402
import attend_infer_repeat_b200 as air  # noqa: E402
from attend_infer_repeat_b200.cell import _init_flat  # noqa: E402

dev = torch.device("cuda", 0)
cfg = air.CellConfig(precision=air.AIR_PREC_TC_SPLIT)
B, T = int(os.environ.get("DIAG_B", 4096)), 3
eng = air.Engine(cfg, B, T, device=dev)
eng.cache_weights(True)
params, _ = _init_flat(air.param_spec(cfg), dev, seed=0)
prior = air.make_prior(dict(loc=0., scale=1.), dict(loc=0., scale=1.), dict(loc=0., scale=1.), 0.5, True)
img = torch.rand(B, 50, 50, device=dev)
ew, ea, u = torch.randn(T, B, 4, device=dev), torch.randn(T, B, cfg.na, device=dev), torch.rand(T, B, 1, device=dev)
if trace:
    for i in range(2):
        eng.forward(params, img, ew, ea, u, prior)
    torch.cuda.synchronize()
    sys.exit(0)

for i in range(5):
    eng.forward(params, img, ew, ea, u, prior)
torch.cuda.synchronize()
eng.profile(True)
acc = {}
N = 20
for i in range(N):
    eng.forward(params, img, ew, ea, u, prior)
    torch.cuda.synchronize()
    for k, v in eng.stage_times_ms().items():
        acc[k] = acc.get(k, 0.0) + v / N
eng.profile(False)
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
for i in range(50):
    eng.forward(params, img, ew, ea, u, prior)
ev1.record()
torch.cuda.synchronize()
with open(os.path.join(out_dir, "r02_diag.txt"), "w") as f:
    f.write(f"B={B} forward ms/step (50 back to back): {ev0.elapsed_time(ev1) / 50:.4f}\n")
    for k, v in acc.items():
        f.write(f"  stage {k:16s} {v * 1e3:8.1f} us\n")
print(open(os.path.join(out_dir, "r02_diag.txt")).read())
