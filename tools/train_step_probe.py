"""One B=4096 training step (forward with saved activations + backward + RMSProp) for ncu launch lists / profiles."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import attend_infer_repeat_b200 as air
from attend_infer_repeat_b200.cell import _init_flat
from attend_infer_repeat_b200.data import synthetic_multi_mnist_u8

B, T = int(sys.argv[1]) if len(sys.argv) > 1 else 4096, 3
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda", 0)
cfg = air.CellConfig(precision=air.AIR_PREC_TC_SPLIT if os.environ.get("AIR_PROBE_FP32") is None else air.AIR_PREC_FP32)
eng = air.Engine(cfg, B, T, device=dev)
eng.train_enable(True)
params, _ = _init_flat(air.param_spec(cfg), dev, seed=0)
prior = air.make_prior(dict(loc=0., scale=1.), dict(loc=0., scale=1.), dict(loc=0., scale=1.),
                       air.functional.anneal_weight(1 - 1e-15, 1e-7, "exp", 20000, 1e5, 1e3, 1e4), True)
u8 = torch.from_numpy(synthetic_multi_mnist_u8(256, 50, 50, seed=0)[0])
img = (u8[torch.randint(0, 256, (B,))].float() / 255).to(dev).contiguous()
g = torch.Generator(device=dev).manual_seed(0)
ew, ea, u = torch.randn(T, B, 4, device=dev, generator=g), torch.randn(T, B, 50, device=dev, generator=g), torch.rand(T, B, 1, device=dev, generator=g)
n = params.numel()
grad, mg, ms, mom = torch.empty(n, device=dev), torch.zeros(n, device=dev), torch.ones(n, device=dev), torch.zeros(n, device=dev)
for i in range(steps):
    eng.forward(params, img, ew, ea, u, prior)
    eng.backward(params, img, ew, ea, prior, grad)
    eng.rmsprop_step(params, grad, mg, ms, mom, 1e-5)
    if len(sys.argv) > 3:
        torch.cuda.synchronize()
        print(i, "loss", float(eng.scalar("loss")), "grad nan", int(torch.isnan(grad).sum()), "max|g|", float(grad.abs().max()),
              "params nan", int(torch.isnan(params).sum()))
        eng.check_range()
torch.cuda.synchronize()
print("loss", float(eng.scalar("loss")))
if os.environ.get("AIR_PROBE_TIME"):      # device time per training step (CUDA events on the launching stream)
    n_t = int(os.environ["AIR_PROBE_TIME"])
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    first_bad = torch.full((), -1, device=dev, dtype=torch.int64)      # first step with a non-finite gradient (no host sync)
    track = os.environ.get("AIR_PROBE_NANTRACK") is not None
    e0.record()
    for i in range(n_t):
        eng.forward(params, img, ew, ea, u, prior)
        eng.backward(params, img, ew, ea, prior, grad)
        if track:
            bad = ~torch.isfinite(grad).all()
            first_bad = torch.where(bad & (first_bad < 0), torch.full_like(first_bad, i), first_bad)
        eng.rmsprop_step(params, grad, mg, ms, mom, 1e-5)
    e1.record()
    if track:
        print("first step with a non-finite gradient:", int(first_bad), "max|g|", float(grad.abs().max()))
    torch.cuda.synchronize()
    print(f"train step: {e0.elapsed_time(e1) / n_t:.4f} ms over {n_t} steps, AIR_SIDE_STREAMS={os.environ.get('AIR_SIDE_STREAMS', 'default')}, "
          f"loss {float(eng.scalar('loss')):.4f}, grad nan {int(torch.isnan(grad).sum())}")
