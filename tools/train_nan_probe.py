"""Hunt for the first non-finite value in a B=4096 training loop (same loop as bench.py's train_step: rotating input sets,
fixed noise per set): after every step, synchronise and look at the loss, the pass's outputs, the gradient (per parameter
entry) and the parameters; stop and report at the first non-finite one.

    python tools/train_nan_probe.py [steps] [fresh_noise 0|1] [sets]
"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import attend_infer_repeat_b200 as air
from attend_infer_repeat_b200.cell import _init_flat
from attend_infer_repeat_b200.data import synthetic_multi_mnist_u8

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 400
fresh = int(sys.argv[2]) if len(sys.argv) > 2 else 0
n_sets = int(sys.argv[3]) if len(sys.argv) > 3 else 4
B, T = 4096, 3
dev = torch.device("cuda", 0)
cfg = air.CellConfig(precision=air.AIR_PREC_TC_SPLIT)
eng = air.Engine(cfg, B, T, device=dev)
eng.train_enable(True)
spec = air.param_spec(cfg)
params, _ = _init_flat(spec, dev, seed=0)
prior = air.make_prior(dict(loc=0., scale=1.), dict(loc=0., scale=1.), dict(loc=0., scale=1.),
                       air.functional.anneal_weight(1 - 1e-15, 1e-7, "exp", 20000, 1e5, 1e3, 1e4), True)
base = torch.from_numpy(synthetic_multi_mnist_u8(256, 50, 50, seed=0)[0])
g = torch.Generator(device=dev).manual_seed(1234)
sets = []
for s in range(n_sets):
    idx = torch.randint(0, 256, (B,), generator=torch.Generator().manual_seed(s))
    sets.append(((base[idx].float() / 255).to(dev).contiguous(), torch.randn(T, B, 4, device=dev, generator=g),
                 torch.randn(T, B, cfg.na, device=dev, generator=g), torch.rand(T, B, 1, device=dev, generator=g)))
n = params.numel()
grad, mg, ms, mom = torch.empty(n, device=dev), torch.zeros(n, device=dev), torch.ones(n, device=dev), torch.zeros(n, device=dev)


def report(step, what):
    print(f"step {step}: first non-finite value in {what}")
    off = 0
    for name, (r, c) in spec:
        gb = int((~torch.isfinite(grad[off:off + r * c])).sum())
        pb = int((~torch.isfinite(params[off:off + r * c])).sum())
        if gb or pb:
            print(f"   {name:28s} grad non-finite {gb:8d} / {r * c:8d}   params non-finite {pb}")
        off += r * c
    for k, v in eng.out.items():
        if v is not None and v.is_floating_point():
            bad = ~torch.isfinite(v)
            if bool(bad.any()):
                rows = bad.reshape(-1, v.shape[-1]).any(1).nonzero().reshape(-1) if v.dim() > 1 else bad.nonzero().reshape(-1)
                print(f"   out[{k}] {tuple(v.shape)}: {int(bad.sum())} non-finite, first rows {rows[:6].tolist()}")
    wh = eng.out["where"].reshape(T, B, 4)
    print("   min |sx| %.3g  min |sy| %.3g" % (float(wh[..., 0].abs().min()), float(wh[..., 2].abs().min())))
    for _ in range(2):      # the first failing call reports (and clears) the gradient-GEMM flag, the second the forward's
        try:
            eng.check_range()
            print("   range flags: clear")
            break
        except Exception as e:
            print("   range flag:", e)


prev = None
for i in range(steps):
    img, ew, ea, u = sets[i % n_sets]
    if fresh:
        ew, ea, u = eng.draw_noise(1000 + i)
    eng.forward(params, img, ew, ea, u, prior)
    eng.backward(params, img, ew, ea, prior, grad)
    torch.cuda.synchronize()
    loss = float(eng.scalar("loss"))
    gmax = float(grad.abs().max())
    if not (loss == loss and abs(loss) < 1e30):
        report(i, f"the forward pass (loss {loss}); previous step: {prev}")
        break
    if not bool(torch.isfinite(grad).all()):
        report(i, f"the gradient (loss {loss:.4f} is finite); previous step: {prev}")
        break
    eng.rmsprop_step(params, grad, mg, ms, mom, 1e-5)
    torch.cuda.synchronize()
    if not bool(torch.isfinite(params).all()):
        report(i, f"the parameters after the optimiser step (loss {loss:.4f}, max|g| {gmax:.4g}); previous step: {prev}")
        break
    wh = eng.out["where"].reshape(T, B, 4)
    prev = dict(loss=round(loss, 4), gmax=gmax, min_sx=float(wh[..., 0].abs().min()), min_sy=float(wh[..., 2].abs().min()))
    if i % 100 == 0:
        print(i, prev, flush=True)
else:
    print(f"no non-finite value in {steps} steps (fresh_noise={fresh}); last {prev}")
