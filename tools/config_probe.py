"""Forward + ELBO and training-step timings for an arbitrary configuration (BASELINE.json configs[3]: 100x100 canvas,
28x28 glimpse, 5 steps, batch 2048 -- the STN-bandwidth-bound regime), CUDA events on the launching stream."""
import argparse
import json
import sys
import torch
sys.path.insert(0, ".")
import attend_infer_repeat_b200 as air
from attend_infer_repeat_b200.cell import _init_flat
from attend_infer_repeat_b200.data import synthetic_multi_mnist_u8

ap = argparse.ArgumentParser()
ap.add_argument("--H", type=int, default=100)
ap.add_argument("--glimpse", type=int, default=28)
ap.add_argument("--T", type=int, default=5)
ap.add_argument("--batch", type=int, default=2048)
ap.add_argument("--steps", type=int, default=30)
args = ap.parse_args()
dev = torch.device("cuda", 0)
H, g, T, B = args.H, args.glimpse, args.T, args.batch
cfg = air.CellConfig(H=H, W=H, h=g, w=g, precision=air.AIR_PREC_TC_SPLIT)
params, _ = _init_flat(air.param_spec(cfg), dev, seed=0)
prior = air.make_prior(dict(loc=0., scale=1.), dict(loc=0., scale=1.), dict(loc=0., scale=1.),
                       air.functional.anneal_weight(1 - 1e-15, 1e-7, "exp", 20000, 1e5, 1e3, 1e4), True)
u8 = torch.from_numpy(synthetic_multi_mnist_u8(256, H, H, seed=0)[0])
sets = []
for s in range(4):
    idx = torch.randint(0, 256, (B,), generator=torch.Generator().manual_seed(s))
    sets.append((u8[idx].float() / 255).to(dev).contiguous())
eng = air.Engine(cfg, B, T, device=dev)
eng.cache_weights(True)
noise = eng.draw_noise(1)
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(5):
    eng.forward(params, sets[i % 4], *noise, prior)
torch.cuda.synchronize()
ev0.record()
for i in range(args.steps):
    eng.forward(params, sets[i % 4], *noise, prior)
ev1.record()
torch.cuda.synchronize()
fwd_ms = ev0.elapsed_time(ev1) / args.steps
eng.profile(True)
eng.forward(params, sets[0], *noise, prior)
stages = {k: round(v, 4) for k, v in eng.stage_times_ms().items()}
eng.profile(False)
loss = float(eng.scalar("loss"))
eng.close()
teng = air.Engine(cfg, B, T, device=dev)
teng.train_enable(True)
n = params.numel()
p2 = params.clone()
grad, mg, ms, mom = torch.empty(n, device=dev), torch.zeros(n, device=dev), torch.ones(n, device=dev), torch.zeros(n, device=dev)
def step(i):
    teng.forward(p2, sets[i % 4], *noise, prior)
    teng.backward(p2, sets[i % 4], noise[0], noise[1], prior, grad)
    teng.rmsprop_step(p2, grad, mg, ms, mom, 1e-5)
for i in range(3):
    step(i)
torch.cuda.synchronize()
nt = max(5, args.steps // 3)
ev0.record()
for i in range(nt):
    step(i)
ev1.record()
torch.cuda.synchronize()
train_ms = ev0.elapsed_time(ev1) / nt
P, G = H * H, g * g
alg_bytes = B * (P * 4 + T * (4 + cfg.na + 1) * 4 + T * (P + G + 3 * cfg.na + 12 + 2) * 4)
print(json.dumps({"config": f"{H}x{H} canvas, {g}x{g} glimpse, T={T}, B={B}, tcgen05 split engine",
                  "forward_ms": fwd_ms, "forward_cell_steps_per_s": B * T / (fwd_ms * 1e-3),
                  "forward_algorithmic_gbs": alg_bytes / (fwd_ms * 1e-3) / 1e9, "stage_ms": stages,
                  "train_step_ms": train_ms, "train_cell_steps_per_s": B * T / (train_ms * 1e-3), "loss": loss,
                  "train_workspace_mb": round(teng.train_workspace_bytes / 1e6, 1)}))
