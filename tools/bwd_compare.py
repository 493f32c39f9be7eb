"""Debug: tensor-core vs SIMT weight gradients at a large batch (same forward, same inputs)."""
import os, sys
import torch
import attend_infer_repeat_b200 as air
from attend_infer_repeat_b200.cell import _init_flat
from attend_infer_repeat_b200.data import synthetic_multi_mnist_u8

B, T = int(sys.argv[1]) if len(sys.argv) > 1 else 4096, 3
dev = torch.device("cuda", 0)
cfg = air.CellConfig(precision=air.AIR_PREC_FP32)
params, _ = _init_flat(air.param_spec(cfg), dev, seed=0)
prior = air.make_prior(dict(loc=0., scale=1.), dict(loc=0., scale=1.), dict(loc=0., scale=1.),
                       air.functional.anneal_weight(1 - 1e-15, 1e-7, "exp", 20000, 1e5, 1e3, 1e4), True)
u8 = torch.from_numpy(synthetic_multi_mnist_u8(256, 50, 50, seed=0)[0])
img = (u8[torch.randint(0, 256, (B,), generator=torch.Generator().manual_seed(0))].float() / 255).to(dev).contiguous()
g = torch.Generator(device=dev).manual_seed(0)
ew, ea, u = torch.randn(T, B, 4, device=dev, generator=g), torch.randn(T, B, 50, device=dev, generator=g), torch.rand(T, B, 1, device=dev, generator=g)
grads = {}
losses = {}
for mode in ("simt", "tc", "tc_fwd"):
    if mode == "simt":
        os.environ["AIR_NO_TC_BWD"] = "1"
    else:
        os.environ.pop("AIR_NO_TC_BWD", None)
    eng = air.Engine(air.CellConfig(precision=air.AIR_PREC_TC_SPLIT) if mode == "tc_fwd" else cfg, B, T, device=dev)
    eng.train_enable(True)
    eng.forward(params, img, ew, ea, u, prior)
    grads[mode] = eng.backward(params, img, ew, ea, prior).clone()
    losses[mode] = float(eng.scalar("loss"))
    torch.cuda.synchronize()
    try:
        eng.check_range()
    except Exception as e:
        print(mode, "RANGE:", e)
    eng.close()
print("loss", losses)
off = 0
for name, (r, c) in air.param_spec(cfg):
    a, b, f = (grads[m][off:off + r * c] for m in ("simt", "tc", "tc_fwd"))
    off += r * c
    err = float((a - b).abs().max()); sc = float(a.abs().max()); errf = float((a - f).abs().max())
    print(f"{name:28s} max|g| {sc:.3e} err {err:.3e} rel {err / (sc + 1e-30):.2e} nan_tc {int(torch.isnan(b).sum())} tcfwd_rel {errf / (sc + 1e-30):.2e}")
