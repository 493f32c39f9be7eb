import sys, time, torch
sys.path.insert(0, ".")
import attend_infer_repeat_b200 as air
from attend_infer_repeat_b200.cell import _init_flat
from functools import partial
from oracle import air_oracle as O
B, T = 64, 3
dev = torch.device("cuda", 0)
for prec in (air.AIR_PREC_TC_SPLIT, air.AIR_PREC_FP32):
    cfg = air.CellConfig(precision=prec)
    eng = air.Engine(cfg, B, T, device=dev); eng.train_enable(True)
    params, _ = _init_flat(air.param_spec(cfg), dev, seed=0)
    prior = air.make_prior(dict(loc=0., scale=1.), dict(loc=0., scale=1.), dict(loc=0., scale=1.), 0.5, True)
    img = torch.rand(B, 50, 50, device=dev)
    noise = eng.draw_noise(1)
    n = params.numel(); grad, mg, ms, mom = torch.empty(n, device=dev), torch.zeros(n, device=dev), torch.ones(n, device=dev), torch.zeros(n, device=dev)
    def step():
        eng.forward(params, img, *noise, prior); eng.backward(params, img, noise[0], noise[1], prior, grad); eng.rmsprop_step(params, grad, mg, ms, mom, 1e-5)
    for _ in range(10): step()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(200): step()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 200
    t0 = time.perf_counter()
    for _ in range(200): step()
    cpu = (time.perf_counter() - t0) / 200
    torch.cuda.synchronize()
    print("engine-level B=64 prec", prec, "ms/step", dt * 1e3, "cpu enqueue ms", cpu * 1e3, "launches", eng.launch_count / 410)
    eng.close()
img, nums = O.synthetic_multi_mnist(B, 50, 50, seed=5)
model = air.AIRonMNIST(img.cuda(), nums.cuda(), max_steps=T, explore_eps=1e-3, inpt_encoder_hidden=[256, 256], glimpse_encoder_hidden=[256, 256], glimpse_decoder_hidden=[256, 256], transform_estimator_hidden=[256, 256], steps_pred_hidden=[128, 64], baseline_hidden=[256, 128], transform_var_bias=.5, step_bias=.75, output_multiplier=.5, precision=air.AIR_PREC_TC_SPLIT)
pr = dict(loc=0., scale=1.)
nsp = dict(anneal='exp', init=1. - 1e-15, final=1e-7, steps_div=1e4, steps=1e5, hold_init=1e3)
train_op, gs = model.train_step(1e-5, 0., pr, pr, pr, nsp)
for _ in range(10): train_op()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(200): train_op()
torch.cuda.synchronize(); print("model-level train_op B=64 ms/step", (time.perf_counter() - t0) / 200 * 1e3)
import cProfile, pstats
pr_ = cProfile.Profile(); pr_.enable()
for _ in range(100): train_op()
torch.cuda.synchronize(); pr_.disable()
pstats.Stats(pr_).sort_stats("cumulative").print_stats(18)
