"""BASELINE.json configs[4] per-GPU shapes: 256 canvases x K = 5 particles = 1280 rows, T = 3 (batch 1024 over 4 GPUs)."""
import json, sys, torch
sys.path.insert(0, ".")
import attend_infer_repeat_b200 as air
from attend_infer_repeat_b200.cell import _init_flat
from attend_infer_repeat_b200.data import synthetic_multi_mnist_u8
n, K, T = 256, 5, 3
dev = torch.device("cuda", 0)
cfg = air.CellConfig(precision=air.AIR_PREC_TC_SPLIT)
eng = air.Engine(cfg, n * K, T, device=dev)
eng.cache_weights(True)
params, _ = _init_flat(air.param_spec(cfg), dev, seed=0)
prior = air.make_prior(dict(loc=0., scale=1.), dict(loc=0., scale=1.), dict(loc=0., scale=1.), 0.5, True)
u8 = torch.from_numpy(synthetic_multi_mnist_u8(n, 50, 50, seed=0)[0])
img = (u8.float() / 255).to(dev).repeat_interleave(K, 0).contiguous()
noise = eng.draw_noise(3)
def step():
    eng.forward(params, img, *noise, prior)
    return eng.iwae_bound(K, prior)
for _ in range(5): step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(100): mean, bound, lw = step()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 100
print(json.dumps({"config": "256 canvases x K=5 particles x T=3 per GPU (configs[4]: batch 1024 over 4 GPUs)", "ms_per_pass": ms,
                  "canvas_particle_steps_per_s": n * K * T / (ms * 1e-3), "iwae_bound": float(mean),
                  "mean_elbo_of_the_same_rows": -float(eng.scalar("loss"))}))
