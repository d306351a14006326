"""One device-resident nbg_reduce call for ncu: prof_reduce.py <func> <f32|f64> <shape e.g. 1000x100000> <axes e.g. 1>"""
import sys, torch
sys.path.insert(0, ".")
from numbagg_b200.decorators import run_reduce
name, dt = sys.argv[1], sys.argv[2]
shape = tuple(int(x) for x in sys.argv[3].split("x"))
axes = tuple(int(x) for x in sys.argv[4].split(","))
dtype = torch.float64 if dt == "f64" else torch.float32
g = torch.Generator(device="cuda").manual_seed(0)
a = torch.rand(shape, generator=g, device="cuda", dtype=dtype); a[a <= 0.1] = float("nan")
for _ in range(3):
    run_reduce(name, a, axes)
torch.cuda.synchronize()
