#!/usr/bin/env python
"""Canvas error of the tensor-core engine against the fp32 oracle at B=4096, binned by the smallest sampled |s| of the
canvas's painted steps (the conditioning rule of tests/test_gpu_full_batch.py is chosen from this table)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import attend_infer_repeat_b200 as air
from oracle import air_oracle as O
from tests import util as U
from tests.test_gpu_full_batch import _ill_conditioned

B = 4096
ocfg = U.oracle_cfg(**U.SCRIPT); pc = O.PriorConfig()
params, img, nums, noise = U.make_problem(ocfg, B, seed=4096)
with torch.no_grad():
    ref = O.forward(ocfg, pc, params, img, *noise, global_step=20000)
    r64 = O.forward(ocfg, pc, {k: v.double() for k, v in params.items()}, img.double(), *(n.double() for n in noise), global_step=20000)
res = {}
for name, prec in (("fp32", air.AIR_PREC_FP32), ("tc", air.AIR_PREC_TC_SPLIT)):
    out = U.run_cuda(ocfg, params, img, noise, pc, 20000, precision=prec)
    T = ocfg.T
    same = (out["presence"].reshape(T, B) == ref["outs"]["presence"].reshape(T, B)).all(0)
    _, s_min = _ill_conditioned(ref, T, B)
    cerr = (out["canvas"].reshape(T, B, -1) - ref["canvas"].reshape(T, B, -1)).abs().amax((0, 2))
    cerr64 = (out["canvas"].reshape(T, B, -1).double() - r64["canvas"].reshape(T, B, -1)).abs().amax((0, 2))
    oerr64 = (ref["canvas"].reshape(T, B, -1).double() - r64["canvas"].reshape(T, B, -1)).abs().amax((0, 2))
    werr = (out["where"] - ref["outs"]["where"]).abs().max()
    werr64 = (out["where"].double() - r64["outs"]["where"]).abs().max()
    owerr64 = (ref["outs"]["where"].double() - r64["outs"]["where"]).abs().max()
    print(f"[{name}] where max err vs fp32 oracle {float(werr):.2e}; vs fp64 {float(werr64):.2e}; oracle fp32 vs fp64 {float(owerr64):.2e}; flips {int((~same).sum())}")
    edges = [0, 1e-3, 3e-3, 1e-2, 2e-2, 5e-2, 1e-1, 2e-1, 5e-1, 1e9]
    for lo, hi in zip(edges[:-1], edges[1:]):
        m = same & (s_min >= lo) & (s_min < hi)
        if int(m.sum()):
            print(f"   |s|min in [{lo:.0e},{hi:.0e}): n={int(m.sum()):5d}  canvas err vs fp32 oracle max {float(cerr[m].max()):.2e}  "
                  f"vs fp64 {float(cerr64[m].max()):.2e}  (oracle fp32 vs fp64 {float(oerr64[m].max()):.2e})")
