"""Run 3 device-resident steps of a bench.py workload (for ncu captures): prof_workload.py <workload>"""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
import bench
from numbagg_b200 import decorators as D
family, func, dt, rows, n, params = bench.WORKLOADS[sys.argv[1]]
dev = torch.device("cuda", 0)
tdt = torch.float32 if dt == "f32" else torch.float64
g = torch.Generator(device=dev).manual_seed(0)
shape = (rows, n) if family != "group1d" else (rows * n,)
a = torch.empty(shape, dtype=tdt, device=dev)
flat = a.view(-1)
for s in range(0, flat.numel(), 1 << 28):
    seg = flat[s:s + (1 << 28)]
    seg.uniform_(0, 1, generator=g)
    seg[seg <= bench.nan_frac(family)] = float("nan")
tensors = [a] + ([a * a + 1] if func in bench.TWO_INPUT else [])
labels = None
if family == "group":
    labels = torch.from_numpy(np.random.RandomState(0).randint(0, params["num_labels"], size=n)).to(dev)
elif family == "group1d":
    labels = torch.randint(0, params["num_labels"], (rows * n,), generator=g, device=dev, dtype=torch.int64)
for _ in range(3):
    if family == "move":
        D.run_move(func, tensors, params["window"], params["min_count"], -1)
    elif family == "exp":
        D.run_move_exp(func, tensors, params["alpha"], 0.0, -1)
    elif family == "fill":
        D.run_fill(func, a, n, -1)
    elif family == "matrix":
        D.run_matrix(func, a, window=params["window"], min_count=params["min_count"])
    elif family == "quantile":
        D.run_quantile(a, torch.tensor(np.atleast_1d(params["quantiles"]), dtype=torch.float64, device=dev), (params["axis"] % 2,))
    elif family == "reduce":
        D.run_reduce(func, a, (0, 1) if params["axis"] is None else (params["axis"] % 2,))
    else:
        D.run_group(func, a if family == "group" else a.view(1, -1), labels, params["num_labels"], 1)
torch.cuda.synchronize()
