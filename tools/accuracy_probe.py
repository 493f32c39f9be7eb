#!/usr/bin/env python
"""Error of each engine / schedule against the float64 oracle (and of the fp32 oracle itself), per output tensor.
Run on a GPU box:  python tools/accuracy_probe.py [weight_gain] [B]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import attend_infer_repeat_b200 as air  # noqa: E402
from oracle import air_oracle as O  # noqa: E402
from tests import util as U  # noqa: E402

gain = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
B = int(sys.argv[2]) if len(sys.argv) > 2 else 48
seed = int(sys.argv[3]) if len(sys.argv) > 3 else 8
ocfg = U.oracle_cfg(**U.SCRIPT)
pc = O.PriorConfig()
params, img, nums, noise = U.make_problem(ocfg, B, seed, gain)
ref32 = O.forward(ocfg, pc, params, img, *noise, global_step=60000)
ref64 = O.forward(ocfg, pc, {k: v.double() for k, v in params.items()}, img.double(), *(n.double() for n in noise),
                  global_step=60000)
keys = ["where", "what", "glimpse", "presence_prob"]


def report(name, out):
    same = (out["presence"].reshape(ocfg.T, B) == ref64["outs"]["presence"].reshape(ocfg.T, B).float()).all(0)
    parts = [f"{k} {float((out[k].double().reshape(-1) - ref64['outs'][k].reshape(-1)).abs().max()):.2e}" for k in keys]
    cv = (out["canvas"].double().reshape(ocfg.T, B, -1) - ref64["canvas"].reshape(ocfg.T, B, -1)).abs()[:, same]
    lp = (out["loss_per_sample"].double() - ref64["loss_per_sample"]).abs()[same]
    print(f"{name:28s} " + "  ".join(parts) + f"  canvas {float(cv.max()):.2e}  loss/sample {float(lp.max()):.2e}"
          f"  presence-agree {int(same.sum())}/{B}")


o32 = {k: ref32["outs"][k] for k in keys}
o32.update(presence=ref32["outs"]["presence"], canvas=ref32["canvas"], loss_per_sample=ref32["loss_per_sample"])
report("oracle fp32 (torch CPU)", o32)
report("cuda fp32 engine", U.run_cuda(ocfg, params, img, noise, pc, 60000, precision=air.AIR_PREC_FP32))
report("cuda tc fused chains", U.run_cuda(ocfg, params, img, noise, pc, 60000, precision=air.AIR_PREC_TC_SPLIT))
os.environ["AIR_NO_CHAIN"] = "1"
report("cuda tc per-layer", U.run_cuda(ocfg, params, img, noise, pc, 60000, precision=air.AIR_PREC_TC_SPLIT))
