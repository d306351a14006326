import sys, os, numpy as np, torch
sys.path.insert(0, ".")
from numbagg_b200.decorators import run_group
torch.cuda.set_device(0)
rows, n, K = 10000, 1_000_000, 1000
g=torch.Generator(device="cuda").manual_seed(0)
a=torch.rand((rows,n), generator=g, device="cuda", dtype=torch.float32); a[a<=0.1]=float("nan")
lab=torch.from_numpy(np.random.RandomState(0).randint(0,K,size=n).astype(np.int64)).cuda()
def ev(fn, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    s=torch.cuda.Event(enable_timing=True); e=torch.cuda.Event(enable_timing=True)
    ts=[]
    for _ in range(reps):
        s.record(); fn(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
    return min(ts)
for name in sys.argv[1:]:
    t=ev(lambda: run_group(name, a, lab, K, 1))
    print(f"C={os.environ.get('NBG_RB_C')} NSEG={os.environ.get('NBG_RB_NSEG')} {name}: {t:.3f} ms {4.005e10/t/1e6/6447.8:.2%}", flush=True)
