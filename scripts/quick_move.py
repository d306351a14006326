import time, numpy as np, torch, sys
sys.path.insert(0, ".")
import numbagg_b200 as nb
from numbagg_b200.decorators import run_move
from oracle import oracle
torch.cuda.set_device(0)
def ev(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    s=torch.cuda.Event(enable_timing=True); e=torch.cuda.Event(enable_timing=True)
    ts=[]
    for _ in range(reps):
        s.record(); fn(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
    return np.median(ts), min(ts)
for dt, shape, w in [(torch.float64,(2000,100000),20),(torch.float64,(100,10000),20),(torch.float32,(1000,1000000),1000),(torch.float32,(1000,1000000),20)]:
    g=torch.Generator(device="cuda").manual_seed(0)
    a=torch.rand(shape, generator=g, device="cuda", dtype=dt); a[a<=0.1]=float("nan")
    b=a*a+1
    for name in ["move_mean","move_sum","move_std","move_var","move_cov","move_corr"]:
        arrs=[a,b] if name in("move_cov","move_corr") else [a]
        med,best=ev(lambda: run_move(name, arrs, w, max(1,w//2), -1))
        nbytes=a.numel()*a.element_size()*(len(arrs)+1)
        print(f"{name:10s} {str(dt):14s} {shape} w={w}: {med:.3f} ms  {a.numel()/med/1e6:.1f} Gel/s  {nbytes/med/1e6:.0f} GB/s ({nbytes/med/1e6/6447.8:.2%} of measured copy)", flush=True)
    del a,b
