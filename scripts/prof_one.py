"""Run one device-resident call (for ncu captures): python scripts/prof_one.py <func> <dtype> <rows> <n> [window]"""
import sys, numpy as np, torch
sys.path.insert(0, ".")
from numbagg_b200.decorators import run_move, run_move_exp, run_fill
name, dt, rows, n = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
w = int(sys.argv[5]) if len(sys.argv) > 5 else 20
dtype = torch.float64 if dt == "f64" else torch.float32
g = torch.Generator(device="cuda").manual_seed(0)
a = torch.rand((rows, n), generator=g, device="cuda", dtype=dtype); a[a <= 0.3] = float("nan")
b = a * a + 1
for _ in range(3):
    if name.startswith("move_exp"):
        arrs = [a, b] if name in ("move_exp_nancov", "move_exp_nancorr") else [a]
        run_move_exp(name, arrs, 0.1, 0.0, -1)
    elif name in ("ffill", "bfill"):
        run_fill(name, a, n, -1)
    else:
        arrs = [a, b] if name in ("move_cov", "move_corr") else [a]
        run_move(name, arrs, w, max(1, w // 2), -1)
torch.cuda.synchronize()
