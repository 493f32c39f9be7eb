"""Stress loop for functional.linear on the shape of a once-flaky parity case (64 x 2500 x 256, ELU): 300 fresh random problems."""
import sys, torch
sys.path.insert(0, "/root/repo")
import attend_infer_repeat_b200.functional as AF
from oracle import air_oracle as O
torch.manual_seed(0)
bad_runs = 0
for it in range(300):
    M, K, N = 64, 2500, 256
    x = torch.randn(M, K); w = torch.randn(K, N) / K ** 0.5; b = torch.randn(N)
    ref = torch.nn.functional.elu(x @ w + b)
    out = AF.linear(x.cuda(), w.cuda(), b.cuda(), act=1).cpu()
    err = (out - ref).abs()
    nbad = int((err > 2e-5 + 2e-5 * ref.abs()).sum())
    if nbad:
        bad_runs += 1
        print("iter", it, "bad", nbad, "max err", float(err.max()))
print("bad runs", bad_runs, "of 300")
