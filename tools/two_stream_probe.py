"""Throughput of two AIR handles fed alternately on two streams (batch k+1's encoder / LSTM under batch k's paint)."""
import sys, time, torch
sys.path.insert(0, ".")
import attend_infer_repeat_b200 as air
from attend_infer_repeat_b200.cell import _init_flat
from attend_infer_repeat_b200.data import synthetic_multi_mnist_u8
B, T = 4096, 3
dev = torch.device("cuda", 0)
cfg = air.CellConfig(precision=air.AIR_PREC_TC_SPLIT)
params, _ = _init_flat(air.param_spec(cfg), dev, seed=0)
prior = air.make_prior(dict(loc=0., scale=1.), dict(loc=0., scale=1.), dict(loc=0., scale=1.), 0.5, True)
u8 = torch.from_numpy(synthetic_multi_mnist_u8(256, 50, 50, seed=0)[0])
sets = [(u8[torch.randint(0, 256, (B,), generator=torch.Generator().manual_seed(s))].float() / 255).to(dev).contiguous() for s in range(4)]
for n_eng in (1, 2, 3):
    engs = [air.Engine(cfg, B, T, device=dev) for _ in range(n_eng)]
    for e in engs: e.cache_weights(True)
    noise = engs[0].draw_noise(1)
    streams = [torch.cuda.Stream() for _ in range(n_eng)]
    def run(n):
        for i in range(n):
            k = i % n_eng
            with torch.cuda.stream(streams[k]):
                engs[k].forward(params, sets[i % 4], *noise, prior)
    run(12); torch.cuda.synchronize()
    t0 = time.perf_counter(); run(120); torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 120
    print(n_eng, "engines/streams: ms per pass", round(dt * 1e3, 4), "M cell-steps/s", round(B * T / dt / 1e6, 2))
    for e in engs: e.close()
