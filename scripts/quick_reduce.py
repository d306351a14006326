"""Quick device timing of nbg_reduce over the kernel geometries (GB/s of input read)."""
import sys, numpy as np, torch
sys.path.insert(0, ".")
from numbagg_b200.decorators import run_reduce
torch.cuda.set_device(0)
PEAK = 6447.8
INNER = 10
def ev(fn, reps=7):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    s=torch.cuda.Event(enable_timing=True); e=torch.cuda.Event(enable_timing=True)
    ts=[]
    for _ in range(reps):
        s.record()
        for _ in range(INNER): fn()   # queue depth hides the host-side launch path
        e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e) / INNER)
    return float(np.median(ts))
FUNCS = sys.argv[1].split(",") if len(sys.argv) > 1 else ["nansum", "nanmean", "nanvar", "nancount", "nanmax", "nanargmax"]
CASES = [
    ((1000, 100000), (1,)), ((1, 100_000_000), (1,)), ((100_000_000,), (0,)), ((1_000_000, 100), (1,)), ((100_000, 1000), (1,)),
    ((10_000_000, 10), (1,)), ((100000, 1000), (0,)), ((10_000_000, 10), (0,)), ((1000, 1000, 100), (1,)), ((100, 1000, 1000), (0,)),
]
for dt in (torch.float32, torch.float64):
    for shape, axes in CASES:
        g=torch.Generator(device="cuda").manual_seed(0)
        a=torch.rand(shape, generator=g, device="cuda", dtype=dt); a[a<=0.1]=float("nan")
        nbytes=a.numel()*a.element_size()
        line=f"{str(dt)[6:]:8s} {str(shape):22s} ax={axes}: "
        for name in FUNCS:
            ms=ev(lambda: run_reduce(name, a, axes))
            line+=f"{name[3:] if name.startswith('nan') else name}={nbytes/ms/1e6:.0f}({nbytes/ms/1e6/PEAK:.0%}) "
        print(line, flush=True)
        del a
