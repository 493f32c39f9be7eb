import os, sys, torch
sys.path.insert(0, '/root/repo')
os.environ['AIR_CHAIN_TRACE'] = '/root/repo/gpurun_out/ctrace'
import attend_infer_repeat_b200 as air
from attend_infer_repeat_b200.cell import _init_flat
dev = torch.device('cuda', 0)
cfg = air.CellConfig(precision=air.AIR_PREC_TC_SPLIT)
B, T = 4096, 3
eng = air.Engine(cfg, B, T, device=dev)
params, _ = _init_flat(air.param_spec(cfg), dev, seed=0)
prior = air.make_prior(dict(loc=0., scale=1.), dict(loc=0., scale=1.), dict(loc=0., scale=1.), 0.5, True)
img = torch.rand(B, 50, 50, device=dev)
ew, ea, u = torch.randn(T, B, 4, device=dev), torch.randn(T, B, cfg.na, device=dev), torch.rand(T, B, 1, device=dev)
for i in range(3):
    eng.forward(params, img, ew, ea, u, prior)
torch.cuda.synchronize()
