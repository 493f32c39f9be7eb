#!/usr/bin/env python
"""Host (Python / ctypes / launch) time of AIRonMNIST.train_op per step against its device time at B=4096: is the public-API
training step bound by the host?  Also prints the top host-side costs (cProfile)."""
import os, sys, time, cProfile, pstats, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import attend_infer_repeat_b200 as air

dev = torch.device("cuda", 0)
B, T = 4096, 3
imgs = torch.rand(B, 50, 50, device=dev)
model = air.AIRonMNIST(imgs, torch.zeros(3, B, 1, device=dev), max_steps=T, explore_eps=1e-3, inpt_encoder_hidden=[256, 256],
                       glimpse_encoder_hidden=[256, 256], glimpse_decoder_hidden=[256, 256],
                       transform_estimator_hidden=[256, 256], steps_pred_hidden=[128, 64], baseline_hidden=[256, 128],
                       transform_var_bias=.5, step_bias=.75, output_multiplier=.5, precision=air.AIR_PREC_TC_SPLIT)
pr = dict(loc=0., scale=1.)
nsp = dict(anneal='exp', init=1. - 1e-15, final=1e-7, steps_div=1e4, steps=1e5, hold_init=1e3, analytic=True)
train_op, _ = model.train_step(1e-5, 0., pr, pr, pr, nsp)
for _ in range(10):
    train_op()
torch.cuda.synchronize()
n = 200
t0 = time.perf_counter()
for _ in range(n):
    train_op()
t_host = time.perf_counter() - t0
torch.cuda.synchronize()
t_all = time.perf_counter() - t0
print(f"host enqueue {t_host / n * 1e3:.3f} ms/step, wall incl. device {t_all / n * 1e3:.3f} ms/step")
prof = cProfile.Profile()
prof.enable()
for _ in range(50):
    train_op()
prof.disable()
torch.cuda.synchronize()
s = io.StringIO()
pstats.Stats(prof, stream=s).sort_stats("cumulative").print_stats(22)
print(s.getvalue()[:3500])
