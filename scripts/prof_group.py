import sys, numpy as np, torch
sys.path.insert(0, ".")
from numbagg_b200.decorators import run_group
name, dt, rows, n, K = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
dtype = torch.float64 if dt == "f64" else torch.float32
g = torch.Generator(device="cuda").manual_seed(0)
a = torch.rand((rows, n), generator=g, device="cuda", dtype=dtype); a[a <= 0.1] = float("nan")
lab = torch.from_numpy(np.random.RandomState(0).randint(0, K, size=n).astype(np.int64)).cuda()
for _ in range(3):
    run_group(name, a, lab, K, 1)
torch.cuda.synchronize()
