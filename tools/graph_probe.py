import sys, time, torch
sys.path.insert(0, ".")
import attend_infer_repeat_b200 as air
from attend_infer_repeat_b200.cell import _init_flat
B, T = int(sys.argv[1]) if len(sys.argv) > 1 else 64, 3
dev = torch.device("cuda", 0)
cfg = air.CellConfig(precision=air.AIR_PREC_TC_SPLIT)
eng = air.Engine(cfg, B, T, device=dev); eng.train_enable(True)
params, _ = _init_flat(air.param_spec(cfg), dev, seed=0)
prior = air.make_prior(dict(loc=0., scale=1.), dict(loc=0., scale=1.), dict(loc=0., scale=1.), 0.5, True)
img = torch.rand(B, 50, 50, device=dev)
noise = eng.draw_noise(1)
n = params.numel(); grad, mg, ms, mom = torch.empty(n, device=dev), torch.zeros(n, device=dev), torch.ones(n, device=dev), torch.zeros(n, device=dev)
def step():
    eng.forward(params, img, *noise, prior); eng.backward(params, img, noise[0], noise[1], prior, grad); eng.rmsprop_step(params, grad, mg, ms, mom, 1e-5)
for _ in range(5): step()
torch.cuda.synchronize()
p0 = params.clone()
step(); torch.cuda.synchronize(); p_eager = params.clone(); params.copy_(p0)
mg0, ms0, mom0 = mg.clone(), ms.clone(), mom.clone()
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    step(); step()
torch.cuda.current_stream().wait_stream(s)
torch.cuda.synchronize()
params.copy_(p0); mg.copy_(mg0); ms.copy_(ms0); mom.copy_(mom0)
# undo the one eager step's slot updates approximately: just compare graph vs eager from the same state
mg1, ms1, mom1 = mg.clone(), ms.clone(), mom.clone()
step(); torch.cuda.synchronize(); p_e = params.clone()
params.copy_(p0); mg.copy_(mg1); ms.copy_(ms1); mom.copy_(mom1)
with torch.cuda.graph(g):
    step()
params.copy_(p0); mg.copy_(mg1); ms.copy_(ms1); mom.copy_(mom1)
g.replay(); torch.cuda.synchronize()
print("graph vs eager max |dparam|:", float((params - p_e).abs().max()), "loss", float(eng.scalar("loss")))
for _ in range(10): g.replay()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(300): g.replay()
torch.cuda.synchronize(); print("graph replay ms/step", (time.perf_counter() - t0) / 300 * 1e3)
t0 = time.perf_counter()
for _ in range(300): step()
torch.cuda.synchronize(); print("eager ms/step", (time.perf_counter() - t0) / 300 * 1e3)
