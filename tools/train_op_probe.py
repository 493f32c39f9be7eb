"""The PUBLIC-API training step (AIRonMNIST.train_step -> train_op: forward, BaselineMLP, both backward passes, both
centered-RMSProp updates) at B=4096 for ncu launch lists; eager (cuda_graph=False) so that every launch is listed.
usage: train_op_probe.py [B] [steps]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import attend_infer_repeat_b200 as air
from attend_infer_repeat_b200.data import synthetic_multi_mnist_u8

B, T = int(sys.argv[1]) if len(sys.argv) > 1 else 4096, 3
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda", 0)
u8 = torch.from_numpy(synthetic_multi_mnist_u8(256, 50, 50, seed=0)[0])
img = (u8[torch.randint(0, 256, (B,))].float() / 255).to(dev).contiguous()
model = air.AIRonMNIST(img, torch.zeros(3, B, 1, device=dev), max_steps=T, explore_eps=1e-3, inpt_encoder_hidden=[256, 256],
                       glimpse_encoder_hidden=[256, 256], glimpse_decoder_hidden=[256, 256],
                       transform_estimator_hidden=[256, 256], steps_pred_hidden=[128, 64], baseline_hidden=[256, 128],
                       transform_var_bias=.5, step_bias=.75, output_multiplier=.5, precision=air.AIR_PREC_TC_SPLIT, seed=0)
pr = dict(loc=0., scale=1.)
nsp = dict(anneal='exp', init=1. - 1e-15, final=1e-7, steps_div=1e4, steps=1e5, hold_init=1e3, analytic=True)
train_op, _ = model.train_step(1e-5, 0., pr, pr, pr, nsp, cuda_graph=os.environ.get("AIR_PROBE_GRAPH") is not None)
model.global_step = 20000
g = torch.Generator(device=dev).manual_seed(0)
noise = model.cell.draw_noise(B, T, generator=g)
for i in range(steps):
    train_op(None, None, noise)
    torch.cuda.synchronize()
print("loss", float(model.loss.value))
