"""A numpy transliteration of paint_bwd_kernel's index logic (backward_kernels.cuh: tap tables with validity bits, the
footprint rectangle, per-glimpse-column / -row canvas ranges, the two-pass gather, the guarded division by the scale),
checked against autograd through the oracle's stn_paint on square and NON-square shapes, mirrored (negative) scales and
glimpses that fall mostly outside the canvas.  Runs on the CPU: python tools/paint_bwd_model.py"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from oracle import air_oracle as O

def make_btap(coord, n, stride):
    inside = (coord > -1.0) and (coord < n)
    f = np.floor(coord) if np.isfinite(coord) else 0.0
    fi = int(f); ci = fi + 1
    d = (f + 1.0) - coord
    f_ok = 0 <= fi <= n - 1; c_ok = 0 <= ci <= n - 1
    return dict(d=d, i_f=min(max(fi, 0), n - 1) * stride, i_c=min(max(ci, 0), n - 1) * stride, f_ok=f_ok, c_ok=c_ok, inside=inside, aux=coord)

def weight(t, idx):
    w = 0.0
    if t["f_ok"] and t["i_f"] == idx: w += t["d"]
    if t["c_ok"] and t["i_c"] == idx: w += 1.0 - t["d"]
    return w

def paint_bwd(glimpse, where, dC, pres, H, W, h, w):
    """one canvas, one step: glimpse [h,w], where (sx,tx,sy,ty), dC [H,W] -> dglimpse [h,w], dwhere [4]"""
    sx, tx, sy, ty = where
    det = sx * sy; a_inv = sy / det; d_inv = sx / det; ntx = -(a_inv * tx); nty = -(d_inv * ty)
    S_w, S_h = (w - 1) * 0.5, (h - 1) * 0.5
    step_W, step_H = 2.0 / (W - 1), 2.0 / (H - 1)
    tx_t = [make_btap(a_inv * ((-1.0 + j * step_W) * S_w) + ntx * S_w + S_w, w, 1) for j in range(W)]
    ty_t = [make_btap(d_inv * ((-1.0 + r * step_H) * S_h) + nty * S_h + S_h, h, w) for r in range(H)]
    clo, chi = [W] * w, [-1] * w; rlo, rhi = [H] * h, [-1] * h
    c_lo, c_hi, r_lo, r_hi = W, -1, H, -1
    for j, tp in enumerate(tx_t):
        if tp["inside"]:
            c_lo, c_hi = min(c_lo, j), max(c_hi, j)
            if tp["f_ok"]: clo[tp["i_f"]] = min(clo[tp["i_f"]], j); chi[tp["i_f"]] = max(chi[tp["i_f"]], j)
            if tp["c_ok"]: clo[tp["i_c"]] = min(clo[tp["i_c"]], j); chi[tp["i_c"]] = max(chi[tp["i_c"]], j)
    for r, tp in enumerate(ty_t):
        if tp["inside"]:
            r_lo, r_hi = min(r_lo, r), max(r_hi, r)
            if tp["f_ok"]: rlo[tp["i_f"] // w] = min(rlo[tp["i_f"] // w], r); rhi[tp["i_f"] // w] = max(rhi[tp["i_f"] // w], r)
            if tp["c_ok"]: rlo[tp["i_c"] // w] = min(rlo[tp["i_c"] // w], r); rhi[tp["i_c"] // w] = max(rhi[tp["i_c"] // w], r)
    dgl = np.zeros((h, w)); acc = np.zeros(4)
    if pres == 0 or c_hi < c_lo or r_hi < r_lo:
        return dgl, np.zeros(4)
    D = glimpse.reshape(-1)
    for r in range(r_lo, r_hi + 1):
        by = ty_t[r]; dy = by["d"]
        for c in range(c_lo, c_hi + 1):
            bx = tx_t[c]; gv = pres * dC[r, c]; dx = bx["d"]
            ff = D[by["i_f"] + bx["i_f"]] if (bx["f_ok"] and by["f_ok"]) else 0.0
            cc = D[by["i_c"] + bx["i_c"]] if (bx["c_ok"] and by["c_ok"]) else 0.0
            fc = D[by["i_c"] + bx["i_f"]] if (bx["f_ok"] and by["c_ok"]) else 0.0
            cf = D[by["i_f"] + bx["i_c"]] if (bx["c_ok"] and by["f_ok"]) else 0.0
            gx = gv * (((1 - dy) * cc + dy * cf) - (dy * ff + (1 - dy) * fc))
            gy = gv * ((dx * fc + (1 - dx) * cc) - (dx * ff + (1 - dx) * cf))
            acc += [gx * (bx["aux"] - S_w), gx, gy * (by["aux"] - S_h), gy]
    nr = r_hi - r_lo + 1
    U = np.zeros((nr, w))
    for rr in range(nr):
        for i in range(w):
            U[rr, i] = sum(weight(tx_t[c], i) * dC[r_lo + rr, c] for c in range(clo[i], chi[i] + 1))
    for j in range(h):
        for i in range(w):
            dgl[j, i] = pres * sum(weight(ty_t[r], j * w) * U[r - r_lo, i] for r in range(rlo[j], rhi[j] + 1))
    o = np.array([0.0 if acc[0] == 0 else -acc[0] / sx, 0.0 if acc[1] == 0 else -S_w * acc[1] / sx,
                  0.0 if acc[2] == 0 else -acc[2] / sy, 0.0 if acc[3] == 0 else -S_h * acc[3] / sy])
    return dgl, o

rng = np.random.RandomState(0)
worst = 0
for (H, W, h, w) in [(9, 14, 4, 6), (17, 31, 5, 9), (12, 7, 6, 3), (50, 50, 20, 20)]:
    for trial in range(6):
        where = np.array([rng.uniform(0.2, 1.2) * rng.choice([-1, 1]), rng.uniform(-0.8, 0.8), rng.uniform(0.2, 1.2) * rng.choice([-1, 1]), rng.uniform(-0.8, 0.8)])
        if trial == 5: where[0] = 2.5   # mostly outside
        gl = rng.standard_normal((h, w)); dC = rng.standard_normal((H, W)); pres = 1.0
        g_t = torch.tensor(gl[None], dtype=torch.float64, requires_grad=True)
        w_t = torch.tensor(where[None], dtype=torch.float64, requires_grad=True)
        out = O.stn_paint(g_t, w_t, (H, W))
        (out * torch.tensor(dC[None])).sum().backward()
        dgl, dwh = paint_bwd(gl, where, dC, pres, H, W, h, w)
        e1 = np.abs(dgl - g_t.grad[0].numpy()).max() / max(1e-12, np.abs(g_t.grad[0].numpy()).max())
        e2 = np.abs(dwh - w_t.grad[0].numpy()).max() / max(1e-12, np.abs(w_t.grad[0].numpy()).max())
        worst = max(worst, e1, e2)
        print((H, W, h, w), trial, "dglimpse rel err %.2e  dwhere rel err %.2e" % (e1, e2))
print("worst", worst)
