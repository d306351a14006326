"""Three calls of one matrix function on the bench shape (for ncu captures): prof_mat.py <func> [f32]"""
import sys
import torch
sys.path.insert(0, ".")
from numbagg_b200 import decorators as D
func = sys.argv[1]
dt = torch.float32 if len(sys.argv) > 2 and sys.argv[2] == "f32" else torch.float64
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(0)
a = torch.empty((200_000, 32), dtype=dt, device=dev).uniform_(0, 1, generator=g)
a[a <= 0.1] = float("nan")
al = torch.full((200_000,), 0.05, dtype=dt, device=dev)
for _ in range(3):
    if "exp" in func:
        D.run_matrix(func, a, alpha=al)
    else:
        D.run_matrix(func, a, window=100, min_count=10)
torch.cuda.synchronize()
